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
    Reference: ``pyfds/acoustic_flow.py:13-57``. The leapfrog step runs on the device, the row shift on
    the host arrays (the overridden ``sim_step`` is called once per step by ``Field.simulate``)."""

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

    def sim_step(self):
        super().sim_step()
        self.apply_flow()

    def apply_flow(self):
        nx = self.x.samples
        for component in [self.pressure, self.velocity_x, self.velocity_y]:
            for n, f in enumerate(self.flow_t_deltas):
                if self.step % f == 0:
                    component.values[n * nx + 1: (n + 1) * nx] = \
                        component.values[n * nx: (n + 1) * nx - 1]
                    component.values[n * nx: n * nx + 1] = 0
