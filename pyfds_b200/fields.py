"""Field core: grid axes, field components, region constructors and the ``simulate()`` driver.

Host-side mirror of the reference interface ``pyfds/fields.py`` (same class names, constructor
arguments, return values and exceptions). What is different is *how* a step is executed: the
reference assembles ``scipy.sparse.dia_matrix`` operators (``pyfds/fields.py:158-202,273-365``) and
steps them by sparse mat-vec in a Python loop (``pyfds/fields.py:87-93``); here ``simulate()`` hands
the whole run to the CUDA step engine (``pyfds_b200/_engine.py`` -> ``libfdsb200.so``), which keeps
the state in HBM and applies boundaries, sources and probes inside the step kernels.

There is no CPU fallback: a model on the hot path (``_device_model`` set) needs the CUDA library and a
GPU to ``simulate()``. Subclasses that override ``sim_step()`` in Python (coupling, flow, user models)
are still driven step by step through that override, exactly like the reference template method.
"""

import logging as lo

import numpy as np

from . import regions as reg

logger = lo.getLogger('pyfds')


class Field:
    """Base class for all fields. Reference: ``pyfds/fields.py:9-127``."""

    #: name of the device model ('acoustic2d', ...) for classes on the CUDA hot path, else None
    _device_model = None
    #: component attribute names in device order (scalar, x-vector, y-vector)
    _device_components = ()

    def __init__(self):
        self.material_regions = []
        self.step = 0
        self.matrices_assembled = False
        self.t = None

    # ---- abstract interface (pyfds/fields.py:18-65) ------------------------------------------

    @property
    def num_points(self):
        raise NotImplementedError

    def get_index(self, position):
        raise NotImplementedError

    def assemble_matrices(self):
        raise NotImplementedError

    def sim_step(self):
        raise NotImplementedError

    # ---- materials (pyfds/fields.py:34-57) --------------------------------------------------

    def material_vector(self, mat_parameter):
        """Per-point vector of one material parameter. Regions are painted in list order, later
        regions override earlier ones, a region only paints if its material has the attribute.
        Raises ``KeyError`` if no material has it (``pyfds/fields.py:54-55``)."""

        param_found = False
        mat_vector = np.zeros(self.num_points)

        for mat_reg in self.material_regions:
            for mat in mat_reg.materials:
                if hasattr(mat, mat_parameter):
                    _paint(mat_vector, mat_reg.region, getattr(mat, mat_parameter), self)
                    param_found = True

        if not param_found:
            raise KeyError('Material parameter {} not found in set materials.'.format(mat_parameter))

        return mat_vector

    def add_material_region(self, *args, **kwargs):
        new_material_region = reg.MaterialRegion(*args, **kwargs)
        self.material_regions.append(new_material_region)
        logger.info('Material region {} added.'.format(new_material_region.region.name))

    # ---- time loop (pyfds/fields.py:67-95) --------------------------------------------------

    def _uses_device(self):
        """True if the fused device loop may replace the per-step Python loop: the class is on the
        hot path and nobody overrode ``sim_step`` further down the hierarchy (coupling and flow
        subclasses do, and must keep being called once per step)."""
        return self._device_model is not None and \
            getattr(type(self).sim_step, '_on_device', False)

    def _simulate_on_device(self, num_steps, progress_logger=None):
        """Hook for composite fields (``SynchronizedFields``) that can run a whole ``simulate`` call
        with device-resident state although they are not a device model themselves."""
        return False

    def _device_flow(self):
        """Per-row shift periods if the medium flows and the device applies the shift (only
        ``AcousticFlow2D``), else ``None``."""
        return None

    def simulate(self, num_steps=None):
        """Run ``num_steps`` steps (``self.t.samples`` if falsy, as ``pyfds/fields.py:74-75``)."""

        if not num_steps:
            num_steps = self.t.samples
            progress_logger = ProgressLogger(num_steps)
        else:
            progress_logger = None

        if not self.matrices_assembled:
            self.assemble_matrices()
            logger.info('Matrices created.')

        logger.info('Starting simulation of {} steps.'.format(num_steps))

        if self._uses_device():
            from . import _engine
            _engine.run(self, int(num_steps), progress_logger)
        elif self._simulate_on_device(int(num_steps), progress_logger):
            pass                            # coupled fields with their state resident on the device
        else:
            start_step = self.step
            while self.step < start_step + num_steps:
                self.sim_step()
                if progress_logger:
                    progress_logger.log(self.step)
                self.step += 1

        logger.info('Simulation of {} steps completed.'.format(num_steps))

    def _device_step(self):
        """One device step with host ``values`` coherent before and after; does not advance
        ``self.step`` (the caller does, as in ``pyfds/fields.py:89-93``)."""
        if not self.matrices_assembled:
            self.assemble_matrices()
        from . import _engine
        _engine.run(self, 1, None, advance=False)

    def get_point_region(self, position, name=''):
        return reg.PointRegion([self.get_index(position)], position, name=name)

    def reset(self):
        """Zero every component and the step counter; boundaries, materials, outputs and recorded
        ``signals`` stay (``pyfds/fields.py:121-127``)."""
        # instance attributes only: the lazy scipy operators of the device models are properties of
        # the class, and touching them here would assemble (and cache) every one of them
        for component in list(vars(self).values()):
            if isinstance(component, FieldComponent):
                component.values = np.zeros_like(component.values)
        self.step = 0

    # ---- the device handle must never be pickled or forked (gfx.py:72-86 pickles the Field) ------

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_engine_state', None)
        state.pop('_local_slabs', None)
        if '_operators' in state:       # lazy scipy views: rebuilt on demand, never shipped
            state['_operators'] = {}
        return state


def _paint(vector, region, value, field):
    """``vector[region.indices] = value`` without materialising implicit regions."""
    d = region.descriptor if isinstance(region, reg.Region) else None
    if d is None:
        vector[region.indices] = value
    elif d[0] == 'rect':
        _, x0, x1, y0, y1, nx = d
        vector.reshape(-1, nx)[y0:y1 + 1, x0:x1 + 1] = value
    elif d[0] == 'range':
        vector[d[1]:d[2]] = value
    else:
        vector[region.indices] = value


class Field1D(Field):
    """One-dimensional fields. Reference: ``pyfds/fields.py:130-241``."""

    def __init__(self, x_samples, x_delta, t_samples, t_delta, material):
        super().__init__()
        self.x = Dimension(x_samples, x_delta)
        self.t = Dimension(t_samples, t_delta)

        self.add_material_region(self.get_line_region((0, self.x.last), name='main'), material)

    @property
    def num_points(self):
        return self.x.samples

    # scipy operator factories: kept for API compatibility (tests, coupling code and the lazy
    # ``a_*`` attributes use them); the step engine never calls them.
    def d_x(self, factors=None, variant='forward'):
        return _first_derivative(self.num_points, 1, factors, variant)

    def d_x2(self, factors=None):
        return _second_derivative(self.num_points, 1, factors)

    def get_index(self, position):
        return self.x.get_index(position)

    def get_position(self, index):
        return self.x.vector[index]

    def get_line_region(self, position, name=''):
        """Inclusive index range between two x coordinates (``pyfds/fields.py:228-241``)."""
        start = self.get_index(position[0])
        stop = self.get_index(position[1]) + 1
        return reg.LineRegion._from_descriptor(('range', start, max(stop, start)), position,
                                               name=name)


class Field2D(Field):
    """Two-dimensional fields, flat index ``x + y * x.samples``.
    Reference: ``pyfds/fields.py:244-533``."""

    def __init__(self, x_samples, x_delta, y_samples, y_delta, t_samples, t_delta, material):
        super().__init__()
        self.x = Dimension(x_samples, x_delta)
        self.y = Dimension(y_samples, y_delta)
        self.t = Dimension(t_samples, t_delta)

        self.add_material_region(
            self.get_rect_region((0, 0, self.x.last, self.y.last), name='main'), material)

    @property
    def num_points(self):
        return self.x.samples * self.y.samples

    def d_x(self, factors=None, variant='forward'):
        return _first_derivative(self.num_points, 1, factors, variant)

    def d_y(self, factors=None, variant='forward'):
        return _first_derivative(self.num_points, self.x.samples, factors, variant)

    def d_x2(self, factors=None):
        return _second_derivative(self.num_points, 1, factors)

    def d_y2(self, factors=None):
        return _second_derivative(self.num_points, self.x.samples, factors)

    def get_index(self, position):
        return self.x.get_index(position[0]) + self.y.get_index(position[1]) * self.x.samples

    def get_position(self, index):
        # int(index / nx): float division then truncation, as pyfds/fields.py:389
        return self.x.vector[index % self.x.samples], self.y.vector[int(index / self.x.samples)]

    def _positions(self, index_array):
        """Vectorised ``get_position`` with identical arithmetic."""
        nx = self.x.samples
        return (self.x.vector[index_array % nx],
                self.y.vector[(index_array / nx).astype(np.int64)])

    def get_line_region(self, position, name=''):
        """Rounded-DDA line from (start_x, start_y) to (end_x, end_y), inclusive, running start to
        end (``pyfds/fields.py:391-418``)."""
        nx = self.x.samples
        start_idx = self.get_index(position[:2])
        end_idx = self.get_index(position[2:])

        start_x, start_y = start_idx % nx, int(start_idx / nx)
        x_diff = start_x - end_idx % nx
        y_diff = start_y - int(end_idx / nx)
        num_points = max(abs(x_diff), abs(y_diff))
        if num_points == 0:
            # the reference divides 0 / 0 here and fails converting nan to int
            raise ValueError('cannot convert float NaN to integer')

        if y_diff == 0 or x_diff == 0 or abs(x_diff) == abs(y_diff):
            # axis-parallel or exact diagonal: every rounded offset is an integer already
            stride = -(x_diff // num_points) - nx * (y_diff // num_points)
            return reg.LineRegion._from_descriptor(('stride', start_idx, num_points + 1, stride),
                                                   position, name=name)

        ii = np.arange(num_points + 1)
        frac = ii / np.int64(num_points)
        x_position = start_x - np.round(frac * x_diff)
        y_position = start_y - np.round(frac * y_diff)
        indices = (x_position + nx * y_position).astype(np.int64)
        return reg.LineRegion(indices, position, name=name)

    def get_rect_region(self, position, name=''):
        """Inclusive rectangle (origin_x, origin_y, size_x, size_y); sizes may be negative
        (``pyfds/fields.py:420-441``). Index order is x outer, y inner."""
        x_start = self.x.get_index(position[0])
        y_start = self.y.get_index(position[1])
        x_end = self.x.get_index(position[0] + position[2])
        y_end = self.y.get_index(position[1] + position[3])

        x_start, x_end = min(x_start, x_end), max(x_start, x_end)
        y_start, y_end = min(y_start, y_end), max(y_start, y_end)

        return reg.RectRegion._from_descriptor(
            ('rect', x_start, x_end, y_start, y_end, self.x.samples), position, name)

    def get_tri_region(self, position, name=''):
        """Inclusive triangle through three points, found with edge functions after reordering the
        vertices clockwise (``pyfds/fields.py:443-502``)."""
        if (position[2] - position[0]) * (position[5] - position[1]) - \
                (position[3] - position[1]) * (position[4] - position[0]) > 0:
            position = (position[0], position[1],
                        position[4], position[5],
                        position[2], position[3])

        corner = (self.get_index(position[:2]),
                  self.get_index(position[2:4]),
                  self.get_index(position[4:]))
        candidates = np.arange(min(corner), max(corner) + 1, dtype=np.int64)
        px, py = self._positions(candidates)

        def edge(a, b):
            # >= 0 on the edge and to its inner side
            return (px - position[a]) * (position[b + 1] - position[a + 1]) - \
                   (py - position[a + 1]) * (position[b] - position[a])

        inside = (edge(0, 2) >= 0) & (edge(2, 4) >= 0) & (edge(4, 0) >= 0)
        return reg.TriRegion(candidates[inside], position, name)

    def get_ellipse_region(self, centre, radii, name=''):
        """Points inside an axis-parallel ellipse; the scanned index range excludes its upper end
        (``pyfds/fields.py:504-533``)."""
        if np.isscalar(radii):
            radii = (radii, radii)

        min_point_index = self.get_index((centre[0], centre[1] - radii[1]))
        max_point_index = self.get_index((centre[0], centre[1] + radii[1]))
        candidates = np.arange(min_point_index, max_point_index, dtype=np.int64)
        px, py = self._positions(candidates)
        inside = _scalar_square(centre[0] - px) / radii[0] ** 2 \
            + _scalar_square(centre[1] - py) / radii[1] ** 2 <= 1
        return reg.EllipseRegion(candidates[inside], centre, radii, name)


def _scalar_square(array):
    """``v ** 2`` evaluated per element as a *scalar* power (libm ``pow``), which is what the
    reference's per-point loop does (``pyfds/fields.py:519-520``). NumPy's array ``** 2`` is a
    multiplication and differs from it in the last bit for about 1 in 1000 inputs -- enough to flip
    points that lie exactly on the ellipse."""
    return np.array([v ** 2 for v in np.asarray(array, dtype=np.float64).tolist()],
                    dtype=np.float64).reshape(np.shape(array))


def _unit_factors(factors, num_points):
    if factors is None:
        return np.array(1).repeat(num_points)
    return factors


def _first_derivative(num_points, offset, factors, variant):
    """DIA first-difference operator with the factor on the *column*
    (``pyfds/fields.py:158-184,273-328``)."""
    import scipy.sparse as sp
    factors = _unit_factors(factors, num_points)
    shape = (num_points, num_points)
    if variant == 'forward':
        return sp.dia_matrix((np.array([-factors, factors]), [0, offset]), shape=shape)
    if variant == 'central':
        return sp.dia_matrix((np.array([-factors / 2, factors / 2]), [-offset, offset]),
                             shape=shape)
    if variant == 'backward':
        return sp.dia_matrix((np.array([-factors, factors]), [-offset, 0]), shape=shape)
    raise ValueError('Unknown difference quotient variant {}.'.format(variant))


def _second_derivative(num_points, offset, factors):
    """DIA second-difference operator (``pyfds/fields.py:186-202,330-365``)."""
    import scipy.sparse as sp
    factors = _unit_factors(factors, num_points)
    return sp.dia_matrix((np.array([factors, -2 * factors, factors]), [-offset, 0, offset]),
                         shape=(num_points, num_points))


class Dimension:
    """A space or time axis. Reference: ``pyfds/fields.py:536-571``."""

    def __init__(self, samples, increment):
        self.samples = int(samples)
        self.increment = increment
        self.snap_radius = np.finfo(float).eps * 10

    @property
    def vector(self):
        return np.arange(start=0, stop=self.samples) * self.increment

    @property
    def last(self):
        """``max(self.vector)`` for a positive increment, without building the vector."""
        if self.samples > 0 and self.increment > 0:
            return np.int64(self.samples - 1) * self.increment
        return max(self.vector)

    def get_index(self, value):
        """Index of the sample within the (absolute) snap radius of ``value``; exactly one must
        exist (``pyfds/fields.py:557-571``). O(1): only the neighbourhood of ``value/increment``
        is tested, with the same comparison the reference applies to the whole axis."""

        guess = None
        if self.samples > 8 and self.increment > 1e-9:
            with np.errstate(all='ignore'):
                estimate = np.float64(value) / self.increment
            if np.isfinite(estimate):
                guess = int(np.clip(np.rint(estimate), 0, self.samples - 1))
        if guess is None:
            candidates = np.arange(self.samples)
        else:
            candidates = np.arange(max(guess - 3, 0), min(guess + 4, self.samples))

        hits = candidates[np.abs(candidates * self.increment - value) <= self.snap_radius]
        assert len(hits) < 2, "Multiple points found within snap radius of given value."
        assert len(hits) > 0, "No point found within snap radius of given value."

        return int(hits[0])


class FieldComponent:
    """One component of a field: values plus its boundaries and probes.
    Reference: ``pyfds/fields.py:574-633``."""

    def __init__(self, num_points):
        self.values = np.zeros(num_points)
        self.boundaries = []
        self.outputs = []

    def apply_bounds(self, step):
        """Host-side boundary application in list order (``pyfds/fields.py:591-600``)."""
        for bound in self.boundaries:
            self.values[bound.region.indices] = bound.apply(self.values[bound.region.indices],
                                                            step=step)

    def write_outputs(self):
        """Host-side probe sampling (``pyfds/fields.py:602-611``)."""
        for output in self.outputs:
            if not output.signals:
                output.signals = [[self.values[index]] for index in output.region.indices]
            else:
                for index, signal in zip(output.region.indices, output.signals):
                    signal.append(self.values[index])

    def add_boundary(self, *args, **kwargs):
        new_bound = reg.Boundary(*args, **kwargs)
        self.boundaries.append(new_bound)
        logger.info('Boundary {} added.'.format(new_bound.region.name))

    def add_output(self, *args, **kwargs):
        new_output = reg.Output(*args, **kwargs)
        self.outputs.append(new_output)
        logger.info('Output region {} added.'.format(new_output.region.name))


class ProgressLogger:
    """Logs progress in percent at a fixed increment without repeating a message.
    Reference: ``pyfds/fields.py:636-668``."""

    def __init__(self, num_steps, log_increment=5, logger_instance=None):
        self.num_steps = num_steps
        self.log_increment = log_increment
        self.logger = logger_instance if logger_instance else logger
        self._last_message_at = None

    def log(self, current_step):
        percent = int(current_step / self.num_steps * 100)
        if int(current_step / self.num_steps * 100 % self.log_increment) == 0 and \
                self._last_message_at != percent:
            self.logger.info('Simulating. {} % completed.'.format(percent))
            self._last_message_at = percent
