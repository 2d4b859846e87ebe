"""Region, boundary, probe and material records (host side).

Mirror of the reference interface ``pyfds/regions.py:8-180`` -- same class names, constructor
arguments and semantics -- with one structural difference that the hot path needs: a region may be
held *implicitly* (rectangle = four integers, axis-parallel line = start/stop/stride) and only turned
into an explicit index array when somebody reads ``.indices``. The reference materialises every region
as a Python ``list`` of ints (``pyfds/fields.py:440-441``), which is what stops it at ~8192^2 cells;
the implicit form lets the baking code paint a 32768^2 material map with array slices.

``indices`` always has the same values in the same order as the reference list.
"""

import numpy as np

__all__ = [
    'Boundary', 'MaterialRegion', 'Output',
]


class Region:
    """Set of grid points (flat indices) for which some behaviour applies (boundary, material,
    probe). Reference: ``pyfds/regions.py:8-21``."""

    def __init__(self, indices, name=''):
        self._explicit = indices
        self._implicit = None
        self.name = name

    # -- implicit descriptors ---------------------------------------------------------------
    # ('rect', x0, x1, y0, y1, nx)      inclusive index ranges, flat index = x + y * nx
    # ('range', start, stop)            half-open flat index range, stride 1
    # ('stride', start, count, stride)  start + k * stride, k = 0..count-1 (stride may be < 0)

    @classmethod
    def _from_descriptor(cls, descriptor, *args, **kwargs):
        region = cls(None, *args, **kwargs)
        region._implicit = descriptor
        return region

    @property
    def descriptor(self):
        """Implicit description of the region or None if it only exists as an index list."""
        return self._implicit if self._explicit is None else None

    @property
    def indices(self):
        if self._explicit is None:
            self._explicit = _materialise(self._implicit)
        return self._explicit

    @indices.setter
    def indices(self, value):
        self._explicit = value
        self._implicit = None

    def index_array(self):
        """Indices as an int64 array (no copy if already one)."""
        return np.asarray(self.indices, dtype=np.int64).reshape(-1)

    def __len__(self):
        d = self.descriptor
        if d is None:
            return len(self._explicit)
        if d[0] == 'rect':
            return (d[2] - d[1] + 1) * (d[4] - d[3] + 1)
        if d[0] == 'range':
            return d[2] - d[1]
        return d[2]


def _materialise(descriptor):
    kind = descriptor[0]
    if kind == 'rect':
        _, x0, x1, y0, y1, nx = descriptor
        xs = np.arange(x0, x1 + 1, dtype=np.int64)
        ys = np.arange(y0, y1 + 1, dtype=np.int64)
        # x outer, y inner -- the order of the reference comprehension (pyfds/fields.py:440-441)
        return (xs[:, None] + ys[None, :] * nx).reshape(-1)
    if kind == 'range':
        return np.arange(descriptor[1], descriptor[2], dtype=np.int64)
    if kind == 'stride':
        _, start, count, stride = descriptor
        return start + stride * np.arange(count, dtype=np.int64)
    raise ValueError('Unknown region descriptor {}.'.format(kind))


class PointRegion(Region):
    """Region given by individual points. Reference: ``pyfds/regions.py:24-38``."""

    def __init__(self, indices, coordinates, name=''):
        super().__init__(indices, name)
        self.point_coordinates = coordinates


class LineRegion(Region):
    """Region given by a line of points. Reference: ``pyfds/regions.py:41-55``."""

    def __init__(self, indices, coordinates, name=''):
        super().__init__(indices, name)
        self.line_coordinates = coordinates


class RectRegion(Region):
    """Region given by a rectangle of points. Reference: ``pyfds/regions.py:58-72``."""

    def __init__(self, indices, coordinates, name=''):
        super().__init__(indices, name)
        self.rect_coordinates = coordinates


class TriRegion(Region):
    """Region given by a triangle of points. Reference: ``pyfds/regions.py:75-89``."""

    def __init__(self, indices, coordinates, name=''):
        super().__init__(indices, name)
        self.tri_coordinates = coordinates


class EllipseRegion(Region):
    """Region given by an ellipse of points. Reference: ``pyfds/regions.py:92-103``."""

    def __init__(self, indices, centre, radii, name=''):
        super().__init__(indices, name)
        self.centre = centre
        self.radii = radii


class Boundary:
    """Values forced onto (or added to) a field component at every step: fixed boundaries and
    excitation signals. Reference: ``pyfds/regions.py:106-145``.

    ``value`` is a scalar, one signal (``numpy.ndarray`` indexed by the absolute step) applied to all
    points, or a list with one signal per point. ``additive`` multiplies the old value (a ``False``
    therefore contributes ``0 * old``).
    """

    def __init__(self, region, value=0, additive=False):
        self.region = region
        self.value = value
        self.additive = additive

    def kind(self):
        """'scalar', 'signal' or 'signals' -- the three cases of ``pyfds/regions.py:136-145``."""
        if np.ndim(self.value) == 0:
            return 'scalar'
        if isinstance(self.value, np.ndarray):
            return 'signal'
        return 'signals'

    def apply(self, old_values, step):
        """Host-side application (used by ``FieldComponent.apply_bounds``; the device engine bakes
        the same rule into its boundary table instead of calling this)."""
        kind = self.kind()
        if kind == 'scalar':
            return self.additive * old_values + self.value
        if kind == 'signal':
            return self.additive * old_values + self.value[step]
        return [self.additive * old_values[ii] + signal[step]
                for ii, signal in enumerate(self.value)]


class Output:
    """Probe: records the values of a component at the points of a region after every step.
    ``signals[k][s]`` is point k at step s. Reference: ``pyfds/regions.py:148-165``."""

    def __init__(self, region):
        self.region = region
        self.signals = []

    @property
    def mean_signal(self):
        return np.mean(np.asarray(self.signals), axis=0)


class MaterialRegion:
    """Material assignment for a region. Reference: ``pyfds/regions.py:168-180``."""

    def __init__(self, region, material):
        self.region = region
        self.materials = [material]
