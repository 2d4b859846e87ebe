"""Linear acoustics: ``Acoustic1D``, ``Acoustic2D``, ``Acoustic3DAxi`` and ``AcousticMaterial``.

Mirror of the reference interface ``pyfds/acoustics.py`` (constructor arguments, component names,
``assemble_matrices`` / ``sim_step`` / ``is_stable``, the ``a_*`` operator attributes). The classes do
not step anything themselves: ``assemble_matrices`` freezes the material description and the
coefficient expressions of the reference operators (evaluated once per *material* instead of once per
cell), ``sim_step`` / ``simulate`` run the staggered pressure/velocity leapfrog on the CUDA engine.
"""

import numpy as np

from . import _bake
from . import fields as fld

__all__ = [
    'Acoustic1D', 'Acoustic2D', 'Acoustic3DAxi', 'AcousticAxisymmetric', 'AcousticMaterial',
]


class _DeviceModel:
    """Shared plumbing of the models on the CUDA hot path."""

    _material_params = ()
    _operator_names = ()

    def _init_device_model(self):
        self._baked = None
        self._operators = {}

    def assemble_matrices(self):
        """Freezes what the reference's ``assemble_matrices`` reads (materials per region) for the
        device engine; the scipy ``a_*`` operators are built lazily and only if somebody asks."""
        epoch = (self._baked['epoch'] + 1) if self._baked else 0
        if 'material_vector' in vars(self):
            # material_vector was replaced on the instance (MaterialCoupling): per-point parameters
            snapshot = _bake.DenseSnapshot(self, self._material_params)
        else:
            snapshot = _bake.MaterialSnapshot(self, self._material_params)
        self._baked = {'snapshot': snapshot, 'epoch': epoch, 'lossy': False}
        self._operators = {}
        self.matrices_assembled = True

    def _operator(self, name):
        if not self.matrices_assembled:
            return None
        if name not in self._operators:
            self._operators.update(self._build_operators())
        return self._operators[name]

    def reset(self):
        super().reset()
        from . import _engine
        _engine.reset(self)

    def _cfl_ok(self, limit):
        """``np.all(material_vector('sound_velocity') < limit)`` -- the reference's stability test
        (``pyfds/acoustics.py:62-63``). The painted vector only holds velocities of the material
        regions, so if every region passes, so does every cell, and no N-sized vector is needed
        (2 GB at 16384^2); a failing region may be painted over completely by later ones, so only
        then is the vector itself consulted."""
        if 'material_vector' not in vars(self):
            speeds = [getattr(m, 'sound_velocity') for r in self.material_regions
                      for m in r.materials if hasattr(m, 'sound_velocity')]
            if speeds and all(np.ndim(c) == 0 for c in speeds) and \
                    all(bool(c < limit) for c in speeds):
                return np.True_
        return np.all(self.material_vector('sound_velocity') < limit)


def _operator_property(name):
    def getter(self):
        return self._operator(name)

    def setter(self, value):
        self._operators[name] = value

    return property(getter, setter, doc='scipy operator {} of the reference (lazy).'.format(name))


def _lossy(mu):
    return bool(np.any(mu != 0))


class Acoustic1D(_DeviceModel, fld.Field1D):
    """One-dimensional acoustic field. Reference: ``pyfds/acoustics.py:9-63``."""

    _device_model = 'acoustic1d'
    _device_components = ('pressure', 'velocity')
    _material_params = ('sound_velocity', 'density', 'absorption_coef')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.pressure = fld.FieldComponent(self.num_points)
        self.velocity = fld.FieldComponent(self.num_points)

    a_p_v = _operator_property('a_p_v')
    a_v_p = _operator_property('a_v_p')
    a_v_v = _operator_property('a_v_v')

    def _factors(self, c, rho, mu):
        # the reference expressions, verbatim in structure (pyfds/acoustics.py:30-37)
        dt, dx = self.t.increment, self.x.increment
        return (dt / dx * c ** 2 * rho,
                dt / dx / rho,
                dt / dx ** 2 * mu / rho)

    def _coefficient_tables(self, m):
        f, g, h = self._factors(m['sound_velocity'][1:], m['density'][1:],
                                m['absorption_coef'][1:])
        return {'tables': {'FX': f, 'GX': g, 'VM1': h, 'V0': -2 * h, 'VP1': h},
                'lossy': _lossy(m['absorption_coef'][1:])}

    def _build_operators(self):
        f, g, h = self._factors(self.material_vector('sound_velocity'),
                                self.material_vector('density'),
                                self.material_vector('absorption_coef'))
        return {'a_p_v': self.d_x(factors=f), 'a_v_p': self.d_x(factors=g, variant='backward'),
                'a_v_v': self.d_x2(factors=h)}

    def sim_step(self):
        """One leapfrog step on the device (``pyfds/acoustics.py:40-52``)."""
        self._device_step()
    sim_step._on_device = True

    def is_stable(self):
        """CFL check with 1 % headroom (``pyfds/acoustics.py:54-63``)."""
        return self._cfl_ok(0.99 * self.x.increment / self.t.increment)


class Acoustic2D(_DeviceModel, fld.Field2D):
    """Two-dimensional acoustic field. Reference: ``pyfds/acoustics.py:66-139``."""

    _device_model = 'acoustic2d'
    _device_components = ('pressure', 'velocity_x', 'velocity_y')
    _material_params = ('sound_velocity', 'density', 'absorption_coef')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.pressure = fld.FieldComponent(self.num_points)
        self.velocity_x = fld.FieldComponent(self.num_points)
        self.velocity_y = fld.FieldComponent(self.num_points)

    a_p_vx = _operator_property('a_p_vx')
    a_p_vy = _operator_property('a_p_vy')
    a_vx_p = _operator_property('a_vx_p')
    a_vy_p = _operator_property('a_vy_p')
    a_vx_vx = _operator_property('a_vx_vx')
    a_vy_vy = _operator_property('a_vy_vy')

    def _factors(self, c, rho, mu):
        # pyfds/acoustics.py:92-107
        dt, dx, dy = self.t.increment, self.x.increment, self.y.increment
        return {'fx': dt / dx * c ** 2 * rho,
                'fy': dt / dy * c ** 2 * rho,
                'gx': dt / dx / rho,
                'gy': dt / dy / rho,
                'hx': dt / dx ** 2 * mu / rho,
                'hy': dt / dy ** 2 * mu / rho}

    def _coefficient_tables(self, m):
        k = self._factors(m['sound_velocity'][1:], m['density'][1:], m['absorption_coef'][1:])
        # a_vx_vx = (d_x2 + d_y2).todia(): scipy adds both operands into a zeroed 5-diagonal array
        zero = np.zeros_like(k['hx'])
        return {'tables': {'FX': k['fx'], 'FY': k['fy'], 'GX': k['gx'], 'GY': k['gy'],
                           'VM1': zero + k['hx'], 'VP1': zero + k['hx'],
                           'VMN': zero + k['hy'], 'VPN': zero + k['hy'],
                           'V0': (zero + -2 * k['hx']) + -2 * k['hy']},
                'lossy': _lossy(m['absorption_coef'][1:])}

    def _build_operators(self):
        k = self._factors(self.material_vector('sound_velocity'), self.material_vector('density'),
                          self.material_vector('absorption_coef'))
        a_vv = (self.d_x2(factors=k['hx']) + self.d_y2(factors=k['hy'])).todia()
        return {'a_p_vx': self.d_x(factors=k['fx']), 'a_p_vy': self.d_y(factors=k['fy']),
                'a_vx_p': self.d_x(factors=k['gx'], variant='backward'),
                'a_vy_p': self.d_y(factors=k['gy'], variant='backward'),
                'a_vx_vx': a_vv, 'a_vy_vy': a_vv}

    def sim_step(self):
        """One leapfrog step on the device (``pyfds/acoustics.py:111-128``)."""
        self._device_step()
    sim_step._on_device = True

    def is_stable(self):
        """CFL check with 1 % headroom (``pyfds/acoustics.py:130-139``)."""
        return self._cfl_ok(0.99 * min(self.x.increment, self.y.increment) / self.t.increment)


class Acoustic3DAxi(_DeviceModel, fld.Field2D):
    """Three-dimensional, axial-symmetric acoustic field; x is the radial and y the axial direction.
    Reference: ``pyfds/acoustics.py:142-236``."""

    _device_model = 'acoustic3daxi'
    _device_components = ('pressure', 'velocity_x', 'velocity_y')
    _material_params = ('sound_velocity', 'density', 'absorption_coef')

    def __init__(self, *args, **kwargs):
        self._init_device_model()
        super().__init__(*args, **kwargs)
        self.pressure = fld.FieldComponent(self.num_points)
        self.velocity_x = fld.FieldComponent(self.num_points)
        self.velocity_y = fld.FieldComponent(self.num_points)

    a_p_vx = _operator_property('a_p_vx')
    a_p_vy = _operator_property('a_p_vy')
    a_vx_p = _operator_property('a_vx_p')
    a_vy_p = _operator_property('a_vy_p')
    a_vx_vx = _operator_property('a_vx_vx')
    a_vy_vy = _operator_property('a_vy_vy')

    def _radii(self):
        """Radius of every velocity point (``pyfds/acoustics.py:166-176``)."""
        return np.tile(self.x.vector, self.y.samples) + self.x.increment / 2

    def _column_radii(self):
        return self.x.vector + self.x.increment / 2

    def _factors(self, c, rho, mu, r):
        # pyfds/acoustics.py:181-201; r broadcasts per cell (reference) or per column (tables)
        dt, dx, dy = self.t.increment, self.x.increment, self.y.increment
        return {'fx': dt / dx * c ** 2 * rho / r,
                'fy': dt / dy * c ** 2 * rho,
                'gx': dt / dx / rho,
                'gy': dt / dy / rho,
                'hx': dt / dx ** 2 * mu / rho,
                'hy': dt / dy ** 2 * mu / rho,
                'hc': dt / dx * mu / rho / r}

    def _coefficient_tables(self, m):
        c, rho, mu = m['sound_velocity'][1:], m['density'][1:], m['absorption_coef'][1:]
        r = self._column_radii()
        k = self._factors(c[:, None], rho[:, None], mu[:, None], r[None, :])
        hx, hy = k['hx'][:, 0], k['hy'][:, 0]
        zero = np.zeros_like(hx)
        zero2 = np.zeros_like(k['hc'])
        # ((d_x2 + d_y2) + d_x central).todia(): the +-1 diagonals pick up -+ hc/2
        return {'tables': {'GX': k['gx'][:, 0], 'GY': k['gy'][:, 0], 'FY': k['fy'][:, 0],
                           'VMN': zero + hy, 'VPN': zero + hy,
                           'V0': (zero + -2 * hx) + -2 * hy,
                           'EB': self.t.increment * mu / rho},
                'column_tables': {'FX': k['fx'],
                                  'VM1': (zero2 + k['hx']) + -k['hc'] / 2,
                                  'VP1': (zero2 + k['hx']) + k['hc'] / 2},
                'column_vectors': {'R': r, 'RR': r ** 2},
                'lossy': _lossy(mu)}

    def _build_operators(self):
        k = self._factors(self.material_vector('sound_velocity'), self.material_vector('density'),
                          self.material_vector('absorption_coef'), self._radii())
        a_vv = (self.d_x2(factors=k['hx']) + self.d_y2(factors=k['hy'])
                + self.d_x(factors=k['hc'], variant='central')).todia()
        return {'a_p_vx': self.d_x(factors=k['fx']), 'a_p_vy': self.d_y(factors=k['fy']),
                'a_vx_p': self.d_x(factors=k['gx'], variant='backward'),
                'a_vy_p': self.d_y(factors=k['gy'], variant='backward'),
                'a_vx_vx': a_vv, 'a_vy_vy': a_vv}

    def sim_step(self):
        """One leapfrog step on the device (``pyfds/acoustics.py:205-225``)."""
        self._device_step()
    sim_step._on_device = True

    def is_stable(self):
        """CFL check with 1 % headroom (``pyfds/acoustics.py:227-236``)."""
        return self._cfl_ok(0.99 * min(self.x.increment, self.y.increment) / self.t.increment)


#: name used by BASELINE.json for the axisymmetric model
AcousticAxisymmetric = Acoustic3DAxi


class AcousticMaterial:
    """Acoustic material parameters. Reference: ``pyfds/acoustics.py:239-286``."""

    def __init__(self, sound_velocity, density,
                 shear_viscosity=0, bulk_viscosity=0,
                 thermal_conductivity=0, isobaric_heat_cap=1, isochoric_heat_cap=1,
                 absorption_coef=None):
        self.sound_velocity = sound_velocity
        self.density = density
        self.shear_viscosity = shear_viscosity
        self.bulk_viscosity = bulk_viscosity
        self.thermal_conductivity = thermal_conductivity
        self.isobaric_heat_cap = isobaric_heat_cap
        self.isochoric_heat_cap = isochoric_heat_cap
        self._absorption_coef = absorption_coef

    @property
    def absorption_coef(self):
        """Sum of all losses (mu); an explicit falsy override (0, None) falls back to the derived
        value, as in ``pyfds/acoustics.py:276``."""
        if not self._absorption_coef:
            return (4 / 3 * self.shear_viscosity + self.bulk_viscosity
                    + self.thermal_conductivity
                    * (self.isobaric_heat_cap - self.isochoric_heat_cap)
                    / (self.isobaric_heat_cap * self.isochoric_heat_cap))
        return self._absorption_coef

    @absorption_coef.setter
    def absorption_coef(self, value):
        self._absorption_coef = value
