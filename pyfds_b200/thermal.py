"""Heat conduction: ``Thermal1D``, ``Thermal2D``, ``Thermal3DAxi`` and ``ThermalMaterial``.

Mirror of the reference interface ``pyfds/thermal.py``. The explicit update (flux from the temperature
gradient, temperature from the flux divergence) runs on the CUDA engine; only the temperature is
state, the heat-flux components are recomputed every step and stored when the host reads them.
"""

import logging as lo

import numpy as np

from . import fields as fld
from .acoustics import _DeviceModel, _operator_property

__all__ = [
    'Thermal1D', 'Thermal2D', 'Thermal3DAxi', 'ThermalMaterial'
]

logger = lo.getLogger('pyfds')


class Thermal1D(_DeviceModel, fld.Field1D):
    """One-dimensional thermal field. Reference: ``pyfds/thermal.py:12-51``."""

    _device_model = 'thermal1d'
    _device_components = ('temperature', 'heat_flux')
    _material_params = ('density', 'heat_capacity', 'thermal_conductivity_x')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.temperature = fld.FieldComponent(self.num_points)
        self.heat_flux = fld.FieldComponent(self.num_points)

    a_t_q = _operator_property('a_t_q')
    a_q_t = _operator_property('a_q_t')

    def _factors(self, rho, cp, kx):
        # pyfds/thermal.py:33-37
        dt, dx = self.t.increment, self.x.increment
        return dt / dx / rho / cp, 1 / dx * kx

    def _coefficient_tables(self, m):
        a, k = self._factors(m['density'][1:], m['heat_capacity'][1:],
                             m['thermal_conductivity_x'][1:])
        return {'tables': {'FX': a, 'GX': k}, 'lossy': False}

    def _build_operators(self):
        a, k = self._factors(self.material_vector('density'),
                             self.material_vector('heat_capacity'),
                             self.material_vector('thermal_conductivity_x'))
        return {'a_t_q': self.d_x(factors=a), 'a_q_t': self.d_x(factors=k, variant='backward')}

    def sim_step(self):
        """One explicit step on the device (``pyfds/thermal.py:40-51``)."""
        self._device_step()
    sim_step._on_device = True


class Thermal2D(_DeviceModel, fld.Field2D):
    """Two-dimensional thermal field. Reference: ``pyfds/thermal.py:54-107``."""

    _device_model = 'thermal2d'
    _device_components = ('temperature', 'heat_flux_x', 'heat_flux_y')
    _material_params = ('density', 'heat_capacity', 'thermal_conductivity_x',
                        'thermal_conductivity_y')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.temperature = fld.FieldComponent(self.num_points)
        self.heat_flux_x = fld.FieldComponent(self.num_points)
        self.heat_flux_y = fld.FieldComponent(self.num_points)

    a_t_qx = _operator_property('a_t_qx')
    a_t_qy = _operator_property('a_t_qy')
    a_qx_t = _operator_property('a_qx_t')
    a_qy_t = _operator_property('a_qy_t')

    def _factors(self, rho, cp, kx, ky):
        # pyfds/thermal.py:78-89
        dt, dx, dy = self.t.increment, self.x.increment, self.y.increment
        return {'ax': dt / dx / rho / cp,
                'ay': dt / dy / rho / cp,
                'kx': 1 / dx * kx,
                'ky': 1 / dy * ky}

    def _coefficient_tables(self, m):
        k = self._factors(m['density'][1:], m['heat_capacity'][1:],
                          m['thermal_conductivity_x'][1:], m['thermal_conductivity_y'][1:])
        return {'tables': {'FX': k['ax'], 'FY': k['ay'], 'GX': k['kx'], 'GY': k['ky']},
                'lossy': False}

    def _build_operators(self):
        k = self._factors(self.material_vector('density'), self.material_vector('heat_capacity'),
                          self.material_vector('thermal_conductivity_x'),
                          self.material_vector('thermal_conductivity_y'))
        return {'a_t_qx': self.d_x(factors=k['ax']), 'a_t_qy': self.d_y(factors=k['ay']),
                'a_qx_t': self.d_x(factors=k['kx'], variant='backward'),
                'a_qy_t': self.d_y(factors=k['ky'], variant='backward')}

    def sim_step(self):
        """One explicit step on the device (``pyfds/thermal.py:92-107``)."""
        self._device_step()
    sim_step._on_device = True


class Thermal3DAxi(_DeviceModel, fld.Field2D):
    """Three-dimensional, axial-symmetric thermal field; x is the radial and y the axial direction.
    Reference: ``pyfds/thermal.py:110-176``."""

    _device_model = 'thermal3daxi'
    _device_components = ('temperature', 'heat_flux_x', 'heat_flux_y')
    _material_params = ('density', 'heat_capacity', 'thermal_conductivity_x',
                        'thermal_conductivity_y')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.temperature = fld.FieldComponent(self.num_points)
        self.heat_flux_x = fld.FieldComponent(self.num_points)
        self.heat_flux_y = fld.FieldComponent(self.num_points)

    a_t_qx = _operator_property('a_t_qx')
    a_t_qy = _operator_property('a_t_qy')
    a_qx_t = _operator_property('a_qx_t')
    a_qy_t = _operator_property('a_qy_t')

    def _radii(self):
        """Radius of every heat-flux point (``pyfds/thermal.py:128-138``)."""
        return np.tile(self.x.vector, self.y.samples) + self.x.increment / 2

    def _factors(self, rho, cp, kx, ky, r):
        # pyfds/thermal.py:141-157
        dt, dx, dy = self.t.increment, self.x.increment, self.y.increment
        return {'ax': dt / dx / rho / cp / r,
                'ay': dt / dy / rho / cp,
                'kx': 1 / dx * kx,
                'ky': 1 / dy * ky}

    def _coefficient_tables(self, m):
        r = self.x.vector + self.x.increment / 2
        k = self._factors(m['density'][1:, None], m['heat_capacity'][1:, None],
                          m['thermal_conductivity_x'][1:, None],
                          m['thermal_conductivity_y'][1:, None], r[None, :])
        return {'tables': {'FY': k['ay'][:, 0], 'GX': k['kx'][:, 0], 'GY': k['ky'][:, 0]},
                'column_tables': {'FX': k['ax']},
                'column_vectors': {'R': r, 'RR': r ** 2},
                'lossy': False}

    def _build_operators(self):
        k = self._factors(self.material_vector('density'), self.material_vector('heat_capacity'),
                          self.material_vector('thermal_conductivity_x'),
                          self.material_vector('thermal_conductivity_y'), self._radii())
        return {'a_t_qx': self.d_x(factors=k['ax']), 'a_t_qy': self.d_y(factors=k['ay']),
                'a_qx_t': self.d_x(factors=k['kx'], variant='backward'),
                'a_qy_t': self.d_y(factors=k['ky'], variant='backward')}

    def sim_step(self):
        """One explicit step on the device (``pyfds/thermal.py:160-176``)."""
        self._device_step()
    sim_step._on_device = True


class ThermalMaterial:
    """Thermal material parameters; the conductivity is a scalar or an (x, y) pair.
    Reference: ``pyfds/thermal.py:179-211``."""

    def __init__(self, heat_capacity, density, thermal_conductivity):
        self.heat_capacity = heat_capacity
        self.density = density
        self.thermal_conductivity_x = None
        self.thermal_conductivity_y = None
        self.thermal_conductivity = thermal_conductivity

    @property
    def thermal_conductivity(self):
        return self.thermal_conductivity_x, self.thermal_conductivity_y

    @thermal_conductivity.setter
    def thermal_conductivity(self, value):
        if isinstance(value, (list, tuple, np.ndarray)) and len(value) == 2:
            self.thermal_conductivity_x = value[0]
            self.thermal_conductivity_y = value[1]
        elif isinstance(value, (float, int)):
            self.thermal_conductivity_x = value
            self.thermal_conductivity_y = value
        else:
            raise ValueError('Thermal conductivity must either be scalar or a 2 element vector.')
