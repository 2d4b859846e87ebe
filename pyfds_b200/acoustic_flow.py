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
        super().__init__(*args, **kwargs)

        if isinstance(flow, (list, np.ndarray)) and len(flow) == self.y.samples:
            self.flow = np.asarray(flow)
        elif isinstance(flow, (float, int)):
            self.flow = np.ones(self.y.samples) * flow
        else:
            raise ValueError('Flow must either be scalar or a vector with length of y_samples.')

        # period (in steps) after which a row has moved by one cell
        self.flow_t_deltas = (self.x.increment / self.flow / self.t.increment).astype(int)

        if np.any(self.flow_t_deltas == 1) or np.any(self.flow_t_deltas == 0):
            wn.warn('Flow velocity may be to high. Consider reducing t_delta.', stacklevel=2)
            logger.warning('Flow velocity may be to high. Consider reducing t_delta.')

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
        nx = self.x.samples
        for component in [self.pressure, self.velocity_x, self.velocity_y]:
            for n, f in enumerate(self.flow_t_deltas):
                if self.step % f == 0:
                    component.values[n * nx + 1: (n + 1) * nx] = \
                        component.values[n * nx: (n + 1) * nx - 1]
                    component.values[n * nx: n * nx + 1] = 0
