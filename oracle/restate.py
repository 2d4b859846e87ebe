"""CPU restatement of the pyfds time-stepping path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference legs of ``bench.py`` may
import this package; ``pyfds_b200`` never does (it has no CPU path at all).

What is restated (each function cites the reference lines it follows):

* the operator factories ``d_x / d_y / d_x2 / d_y2`` (``pyfds/fields.py:158-202,273-365``) as plain
  ``(data, offsets)`` DIA pairs;
* the arithmetic of the third-party dependency the reference steps with -- scipy (unpinned in
  ``pyproject.toml:16-20``; 1.18.1 in this image): ``_dia_base._add_sparse`` (``scipy/sparse/_dia.py``
  177-222) and ``dia_matvec`` (``scipy/sparse/sparsetools/dia.h``: ``y = 0; for each stored diagonal,
  in order: y[i] += diag[j] * x[j]``), see ``dia_add`` / ``dia_matvec``;
* ``assemble_matrices`` and ``sim_step`` of Acoustic1D/2D/3DAxi (``pyfds/acoustics.py:27-52,89-128,
  178-225``) and Thermal1D/2D/3DAxi (``pyfds/thermal.py:30-51,75-107,139-176``);
* ``FieldComponent.apply_bounds`` / ``write_outputs`` and ``Boundary.apply``
  (``pyfds/fields.py:591-611``, ``pyfds/regions.py:125-145``).

Parity pinning: ``tests/test_oracle.py`` checks this restatement bit for bit against golden fields and
probe signals produced by running the real reference in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``) and against the known-answer matrices of the
reference's own tests (``test/test_fields.py:57-107``, ``test/test_acoustics.py:19-66``).

A stepper is built from any *field-like* object (the real ``pyfds`` classes or the ``pyfds_b200``
mirrors): it reads ``t/x/y``, ``material_vector(name)``, the components' ``values``, ``boundaries`` and
``outputs``, copies what it needs and never touches the object again.
"""

import numpy as np

# ---------------------------------------------------------------------------------------------
# DIA operators
# ---------------------------------------------------------------------------------------------


class Dia:
    """A square DIA matrix: ``data[k, j]`` is the entry in column j of the diagonal ``offsets[k]``
    (scipy convention: the factor sits on the *column*)."""

    def __init__(self, data, offsets, n):
        self.data = np.asarray(data, dtype=np.float64)
        self.offsets = [int(o) for o in offsets]
        self.n = int(n)

    def toarray(self):
        dense = np.zeros((self.n, self.n))
        for k, off in enumerate(self.offsets):
            for j in range(max(0, off), min(self.n + off, self.n)):
                dense[j - off, j] += self.data[k, j]
        return dense


def dia_matvec(a, x, backend='restated'):
    """``a.dot(x)``: scipy ``dia_matvec`` -- start from zeros, add one diagonal after the other in
    stored order; every ``+=`` is one rounded multiply followed by one rounded add (NumPy never fuses).
    ``backend='scipy'`` runs the same loop in scipy's C++ (the code the reference actually executes,
    ``pyfds/acoustics.py:117-128``) and is what the CPU baseline times; ``backend='c'`` runs the plain C
    restatement ``oracle/dia_matvec.c`` (a third, independent implementation of the loop)."""
    if backend == 'scipy':
        return a.scipy().dot(x)
    if backend == 'c':
        from . import cbuild
        x = np.ascontiguousarray(x, dtype=np.float64)
        data = np.ascontiguousarray(a.data, dtype=np.float64)
        offsets = np.asarray(a.offsets, dtype=np.int64)
        y = np.empty(a.n)
        cbuild.library().fds_oracle_dia_matvec(a.n, len(a.offsets), offsets.ctypes.data,
                                               data.ctypes.data, x.ctypes.data, y.ctypes.data)
        return y
    n = a.n
    y = np.zeros(n)
    for k, off in enumerate(a.offsets):
        i_start, j_start = max(0, -off), max(0, off)
        j_end = min(n + off, n)
        count = j_end - j_start
        if count > 0:
            y[i_start:i_start + count] += a.data[k, j_start:j_end] * x[j_start:j_end]
    return y


def _scipy_of(a):
    import scipy.sparse as sp
    return sp.dia_matrix((a.data, a.offsets), shape=(a.n, a.n))


Dia.scipy = lambda self: self.__dict__.setdefault('_sp', _scipy_of(self))


def dia_add(a, b):
    """``a + b`` for two DIA matrices with full-length diagonals, following the three branches of
    scipy 1.18 ``_dia_base._add_sparse``."""
    if a.offsets == b.offsets:
        return Dia(a.data + b.data, a.offsets, a.n)
    new_offsets = sorted(set(a.offsets) | set(b.offsets))
    a_idx = [new_offsets.index(o) for o in a.offsets]
    b_idx = [new_offsets.index(o) for o in b.offsets]
    if len(new_offsets) == len(a.offsets):
        # result structure equals a's: copy a (reordered), add b in place
        data = np.empty((len(new_offsets), a.n))
        for k, row in zip(a_idx, a.data):
            data[k] = row
        for k, row in zip(b_idx, b.data):
            data[k] += row
    elif len(new_offsets) == len(b.offsets):
        data = np.empty((len(new_offsets), a.n))
        for k, row in zip(b_idx, b.data):
            data[k] = row
        for k, row in zip(a_idx, a.data):
            data[k] += row
    else:
        data = np.zeros((len(new_offsets), a.n))
        for k, row in zip(a_idx, a.data):
            data[k] += row
        for k, row in zip(b_idx, b.data):
            data[k] += row
    return Dia(data, new_offsets, a.n)


def d_first(n, offset, factors=None, variant='forward'):
    """First-difference operator times per-point factors (``pyfds/fields.py:158-184,273-328``):
    forward = diagonals (0, +o), backward = (-o, 0), central = (-o, +o) with halved factors."""
    f = np.ones(n) if factors is None else np.asarray(factors, dtype=np.float64)
    if variant == 'forward':
        return Dia([-f, f], [0, offset], n)
    if variant == 'central':
        return Dia([-f / 2, f / 2], [-offset, offset], n)
    if variant == 'backward':
        return Dia([-f, f], [-offset, 0], n)
    raise ValueError('Unknown difference quotient variant {}.'.format(variant))


def d_second(n, offset, factors=None):
    """Second-difference operator (``pyfds/fields.py:186-202,330-365``): diagonals (-o, 0, +o)."""
    f = np.ones(n) if factors is None else np.asarray(factors, dtype=np.float64)
    return Dia([f, -2 * f, f], [-offset, 0, offset], n)


# ---------------------------------------------------------------------------------------------
# boundaries and probes
# ---------------------------------------------------------------------------------------------


class _Component:
    """Copy of a ``FieldComponent``: values, boundary records and probe index lists."""

    def __init__(self, component):
        self.values = np.array(component.values, dtype=np.float64)
        self.bounds = []
        for b in component.boundaries:
            idx = np.asarray(b.region.indices, dtype=np.int64).reshape(-1)
            self.bounds.append((idx, b.value, b.additive))
        self.probes = [np.asarray(o.region.indices, dtype=np.int64).reshape(-1)
                       for o in component.outputs]
        self.signals = [[[] for _ in idx] for idx in self.probes]

    def apply_bounds(self, step):
        """``pyfds/fields.py:591-600`` with ``Boundary.apply`` (``pyfds/regions.py:136-145``):
        boundaries in list order; scalar, ``signal[step]`` or one signal per point; ``additive``
        multiplies the old value."""
        for idx, value, additive in self.bounds:
            old = self.values[idx]
            if np.ndim(value) == 0:
                new = additive * old + value
            elif isinstance(value, np.ndarray):
                new = additive * old + value[step]
            else:
                new = [additive * old[ii] + signal[step] for ii, signal in enumerate(value)]
            self.values[idx] = new

    def write_outputs(self):
        """``pyfds/fields.py:602-611``: one sample per probe point per step."""
        for idx, signals in zip(self.probes, self.signals):
            for point, signal in zip(idx, signals):
                signal.append(self.values[point])

    def probe_arrays(self):
        return [np.array(s, dtype=np.float64).reshape(len(s), -1) for s in self.signals]


# ---------------------------------------------------------------------------------------------
# steppers
# ---------------------------------------------------------------------------------------------


class _Stepper:
    components = ()

    def __init__(self, field, backend='restated'):
        self.backend = backend
        self.step = int(field.step)
        self.dt = field.t.increment
        self.dx = field.x.increment
        self.nx = int(field.x.samples)
        if hasattr(field, 'y'):
            self.dy = field.y.increment
            self.ny = int(field.y.samples)
        else:
            self.ny = 1
        self.n = self.nx * self.ny
        self.comp = {name: _Component(getattr(field, name)) for name in self.components}
        self.assemble(field)

    def dot(self, a, x):
        return dia_matvec(a, x, self.backend)

    def run(self, n_steps):
        for _ in range(int(n_steps)):
            self.sim_step()
            self.step += 1
        return self

    def values(self, name):
        return self.comp[name].values

    def signals(self, name):
        """Probe signals of a component: list (one per Output) of arrays [n_points][n_steps]."""
        return self.comp[name].probe_arrays()


class Acoustic1D(_Stepper):
    """``pyfds/acoustics.py:27-52``."""
    components = ('pressure', 'velocity')

    def assemble(self, field):
        dt, dx, n = self.dt, self.dx, self.n
        c, rho, mu = (field.material_vector(k) for k in
                      ('sound_velocity', 'density', 'absorption_coef'))
        self.a_p_v = d_first(n, 1, dt / dx * c ** 2 * rho)
        self.a_v_p = d_first(n, 1, dt / dx / rho, 'backward')
        self.a_v_v = d_second(n, 1, dt / dx ** 2 * mu / rho)

    def sim_step(self):
        p, v = self.comp['pressure'], self.comp['velocity']
        p.apply_bounds(self.step)
        p.write_outputs()
        v.values -= (self.dot(self.a_v_p, p.values) - self.dot(self.a_v_v, v.values))
        v.apply_bounds(self.step)
        v.write_outputs()
        p.values -= self.dot(self.a_p_v, v.values)


class Acoustic2D(_Stepper):
    """``pyfds/acoustics.py:89-128``."""
    components = ('pressure', 'velocity_x', 'velocity_y')

    def assemble(self, field):
        dt, dx, dy, n, nx = self.dt, self.dx, self.dy, self.n, self.nx
        c, rho, mu = (field.material_vector(k) for k in
                      ('sound_velocity', 'density', 'absorption_coef'))
        self.a_p_vx = d_first(n, 1, dt / dx * c ** 2 * rho)
        self.a_p_vy = d_first(n, nx, dt / dy * c ** 2 * rho)
        self.a_vx_p = d_first(n, 1, dt / dx / rho, 'backward')
        self.a_vy_p = d_first(n, nx, dt / dy / rho, 'backward')
        self.a_vx_vx = dia_add(d_second(n, 1, dt / dx ** 2 * mu / rho),
                               d_second(n, nx, dt / dy ** 2 * mu / rho))
        self.a_vy_vy = self.a_vx_vx

    def sim_step(self):
        p, vx, vy = (self.comp[k] for k in self.components)
        p.apply_bounds(self.step)
        p.write_outputs()
        vx.values -= (self.dot(self.a_vx_p, p.values) - self.dot(self.a_vx_vx, vx.values))
        vy.values -= (self.dot(self.a_vy_p, p.values) - self.dot(self.a_vy_vy, vy.values))
        vx.apply_bounds(self.step)
        vx.write_outputs()
        vy.apply_bounds(self.step)
        vy.write_outputs()
        p.values -= (self.dot(self.a_p_vx, vx.values) + self.dot(self.a_p_vy, vy.values))


class AcousticFlow2D(Acoustic2D):
    """``pyfds/acoustic_flow.py:44-57``: the Acoustic2D step, then every row n with
    ``step % flow_t_deltas[n] == 0`` moves one cell towards +x in all three components."""

    def assemble(self, field):
        super().assemble(field)
        self.flow_t_deltas = [int(f) for f in field.flow_t_deltas]

    def sim_step(self):
        super().sim_step()
        # Python's % on ints has the sign conventions of the reference's `self.step % f` for f != 0;
        # NumPy evaluates `step % 0` to 0 (pyfds/acoustic_flow.py:54 with an np.int64 period)
        moving = [n for n, f in enumerate(self.flow_t_deltas) if f == 0 or self.step % f == 0]
        if not moving:
            return
        for name in self.components:
            grid = self.comp[name].values.reshape(self.ny, self.nx)
            grid[moving, 1:] = grid[moving, :-1]
            grid[moving, 0] = 0


class Acoustic3DAxi(_Stepper):
    """``pyfds/acoustics.py:166-225``."""
    components = ('pressure', 'velocity_x', 'velocity_y')

    def assemble(self, field):
        dt, dx, dy, n, nx = self.dt, self.dx, self.dy, self.n, self.nx
        c, rho, mu = (field.material_vector(k) for k in
                      ('sound_velocity', 'density', 'absorption_coef'))
        # _radii: x coordinate of every velocity point (pyfds/acoustics.py:166-176)
        self.radii = np.tile(np.arange(nx) * dx, self.ny) + dx / 2
        self.mu, self.rho = mu, rho
        r = self.radii
        self.a_p_vx = d_first(n, 1, dt / dx * c ** 2 * rho / r)
        self.a_p_vy = d_first(n, nx, dt / dy * c ** 2 * rho)
        self.a_vx_p = d_first(n, 1, dt / dx / rho, 'backward')
        self.a_vy_p = d_first(n, nx, dt / dy / rho, 'backward')
        self.a_vx_vx = dia_add(dia_add(d_second(n, 1, dt / dx ** 2 * mu / rho),
                                       d_second(n, nx, dt / dy ** 2 * mu / rho)),
                               d_first(n, 1, dt / dx * mu / rho / r, 'central'))
        self.a_vy_vy = self.a_vx_vx

    def sim_step(self):
        p, vx, vy = (self.comp[k] for k in self.components)
        p.apply_bounds(self.step)
        p.write_outputs()
        vx.values -= (self.dot(self.a_vx_p, p.values) - self.dot(self.a_vx_vx, vx.values)
                      + self.dt * self.mu / self.rho * vx.values / self.radii ** 2)
        vy.values -= (self.dot(self.a_vy_p, p.values) - self.dot(self.a_vy_vy, vy.values))
        vx.apply_bounds(self.step)
        vx.write_outputs()
        vy.apply_bounds(self.step)
        vy.write_outputs()
        p.values -= (self.dot(self.a_p_vx, vx.values * self.radii)
                     + self.dot(self.a_p_vy, vy.values))


class Thermal1D(_Stepper):
    """``pyfds/thermal.py:30-51``."""
    components = ('temperature', 'heat_flux')

    def assemble(self, field):
        dt, dx, n = self.dt, self.dx, self.n
        rho, cp, kx = (field.material_vector(k) for k in
                       ('density', 'heat_capacity', 'thermal_conductivity_x'))
        self.a_t_q = d_first(n, 1, dt / dx / rho / cp)
        self.a_q_t = d_first(n, 1, 1 / dx * kx, 'backward')

    def sim_step(self):
        t, q = self.comp['temperature'], self.comp['heat_flux']
        t.apply_bounds(self.step)
        t.write_outputs()
        q.values = -self.dot(self.a_q_t, t.values)
        q.apply_bounds(self.step)
        q.write_outputs()
        t.values -= self.dot(self.a_t_q, q.values)


class Thermal2D(_Stepper):
    """``pyfds/thermal.py:75-107``."""
    components = ('temperature', 'heat_flux_x', 'heat_flux_y')
    axi = False

    def assemble(self, field):
        dt, dx, dy, n, nx = self.dt, self.dx, self.dy, self.n, self.nx
        rho, cp, kx, ky = (field.material_vector(k) for k in
                           ('density', 'heat_capacity', 'thermal_conductivity_x',
                            'thermal_conductivity_y'))
        if self.axi:
            # pyfds/thermal.py:128-143
            self.radii = np.tile(np.arange(nx) * dx, self.ny) + dx / 2
            self.a_t_qx = d_first(n, 1, dt / dx / rho / cp / self.radii)
        else:
            self.a_t_qx = d_first(n, 1, dt / dx / rho / cp)
        self.a_t_qy = d_first(n, nx, dt / dy / rho / cp)
        self.a_qx_t = d_first(n, 1, 1 / dx * kx, 'backward')
        self.a_qy_t = d_first(n, nx, 1 / dy * ky, 'backward')

    def sim_step(self):
        t, qx, qy = (self.comp[k] for k in self.components)
        t.apply_bounds(self.step)
        t.write_outputs()
        qx.values = -self.dot(self.a_qx_t, t.values)
        qy.values = -self.dot(self.a_qy_t, t.values)
        qx.apply_bounds(self.step)
        qx.write_outputs()
        qy.apply_bounds(self.step)
        qy.write_outputs()
        flux_x = qx.values * self.radii if self.axi else qx.values
        t.values -= (self.dot(self.a_t_qx, flux_x) + self.dot(self.a_t_qy, qy.values))


class Thermal3DAxi(Thermal2D):
    """``pyfds/thermal.py:139-176``."""
    axi = True


STEPPERS = {
    'Acoustic1D': Acoustic1D, 'Acoustic2D': Acoustic2D, 'AcousticFlow2D': AcousticFlow2D,
    'Acoustic3DAxi': Acoustic3DAxi,
    'Thermal1D': Thermal1D, 'Thermal2D': Thermal2D, 'Thermal3DAxi': Thermal3DAxi,
}


def stepper_for(field, backend='restated'):
    """Stepper matching the class name of ``field`` (searching its bases, so subclasses work)."""
    for klass in type(field).__mro__:
        if klass.__name__ in STEPPERS:
            return STEPPERS[klass.__name__](field, backend)
    raise TypeError('No restatement for {}'.format(type(field).__name__))
