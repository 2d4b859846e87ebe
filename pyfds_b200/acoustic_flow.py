"""Acoustics in a moving medium. Mirror of ``pyfds/acoustic_flow.py``."""

import logging as lo
import warnings as wn

import numpy as np

from . import acoustics as acs

__all__ = [
    'AcousticFlow2D',
]

logger = lo.getLogger('pyfds')


class AcousticFlow2D(acs.Acoustic2D):
    """Two-dimensional acoustic field in a medium flowing along x: after every leapfrog step each grid
    row is shifted by one cell every ``flow_t_deltas[row]`` steps.
    Reference: ``pyfds/acoustic_flow.py:13-57``. Both the leapfrog step and the row shift run on the
    device (``fds_set_flow``, ``include/fdsb200.h``): ``simulate(n)`` is one device call, the step
    kernels still advance several steps per launch and end a launch where a row has to move.
    ``apply_flow`` remains as the host statement of the shift; if a subclass overrides it, the field is
    stepped through ``sim_step`` once per step and the override is called on the host arrays."""

    def __init__(self, flow, *args, **kwargs):
        """``flow``: velocity of the medium along x, one number or one per grid row; the remaining
        arguments are those of ``Field2D`` (``pyfds/acoustic_flow.py:20-43``)."""
        super().__init__(*args, **kwargs)
        rows = self.y.samples
        scalar = isinstance(flow, (float, int))
        per_row = isinstance(flow, (list, np.ndarray)) and len(flow) == rows
        if not (scalar or per_row):
            raise ValueError('Flow must either be scalar or a vector with length of y_samples.')
        self.flow = np.asarray(flow) if per_row else np.full(rows, flow, dtype=np.float64)
        # a row has moved by one cell after this many steps (truncated towards zero, as astype does)
        self.flow_t_deltas = (self.x.increment / self.flow / self.t.increment).astype(int)
        if np.isin(self.flow_t_deltas, (0, 1)).any():
            message = 'Flow velocity may be to high. Consider reducing t_delta.'
            wn.warn(message, stacklevel=2)
            logger.warning(message)

    def _flow_on_device(self):
        return type(self).apply_flow is AcousticFlow2D.apply_flow

    def _device_flow(self):
        """``|flow_t_deltas|`` per grid row: ``step % f == 0`` (``pyfds/acoustic_flow.py:54``) does not
        depend on the sign of ``f``; ``f == 0`` holds for every step (NumPy evaluates ``step % 0`` to
        0) and the most negative integer (what ``astype(int)`` makes of an infinite period, i.e. of
        zero flow) only divides step 0."""
        if not self._flow_on_device():
            return None
        periods = np.asarray(self.flow_t_deltas).astype(np.int64).reshape(-1)
        if periods.shape[0] != self.y.samples:
            raise ValueError('flow_t_deltas must have one entry per grid row.')
        limits = np.iinfo(np.int64)
        return np.where(periods == limits.min, limits.max, np.abs(periods))

    def _uses_device(self):
        return super()._uses_device() and self._flow_on_device()

    def sim_step(self):
        if self._flow_on_device():
            self._device_step()          # leapfrog step and row shift in one device call
        else:
            super().sim_step()
            self.apply_flow()
    sim_step._on_device = True

    def apply_flow(self):
        """Host statement of the shift (``pyfds/acoustic_flow.py:49-57``): every row whose period
        divides the current step moves one cell towards +x and gets a zero at x = 0. All due rows of a
        component move in one fancy-indexed assignment (its right-hand side is a copy, so the order
        of reads and writes is that of the reference's per-row slices)."""
        periods = np.asarray(self.flow_t_deltas)
        with np.errstate(divide='ignore', invalid='ignore'):
            due = np.mod(self.step, periods) == 0       # a period of 0 is due at every step
        if not due.any():
            return
        for component in (self.pressure, self.velocity_x, self.velocity_y):
            grid = component.values.reshape(self.y.samples, self.x.samples)
            grid[due, 1:] = grid[due, :-1]
            grid[due, 0] = 0
            if not np.shares_memory(grid, component.values):
                component.values = grid.reshape(-1)
