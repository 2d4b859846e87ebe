"""INTEGRATION.md option B, executed: the UNMODIFIED reference package (staged in ``baseline/_ref`` by
``__graft_entry__.build()``) builds the golden scenarios with its own classes, ``Field.simulate`` is
routed through ``libfdsb200.so`` by the standalone ctypes stub ``integration/pyfds_b200_stub.py`` (no
``pyfds_b200`` Python code involved), and the results are compared bit for bit with the goldens the
same reference produced on its scipy path."""

import os
import sys
import types

import numpy as np
import pytest

import scenarios
from conftest import bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

pytestmark = pytest.mark.gpu

NAMES = ['acoustic1d_lossy', 'acoustic1d_lossless', 'acoustic2d_lossless', 'acoustic2d_lossy',
         'acoustic2d_wide', 'acoustic2d_lossy_wide', 'acoustic2d_boundaries',
         'acoustic2d_signal_lines', 'thermal2d', 'thermal2d_wide']


@pytest.fixture(scope='module')
def reference(library):
    if not os.path.isdir(os.path.join(STAGED, 'pyfds')):
        pytest.skip('the reference is not staged in baseline/_ref (run __graft_entry__.build())')
    for name in ('matplotlib', 'matplotlib.patches', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, STAGED)
    sys.path.insert(0, os.path.join(ROOT, 'integration'))
    import warnings
    warnings.simplefilter('ignore', DeprecationWarning)
    import pyfds
    import pyfds_b200_stub
    assert os.path.abspath(pyfds.__file__).startswith(STAGED)
    original = pyfds_b200_stub.install(pyfds, library)
    yield pyfds
    pyfds.fields.Field.simulate = original


@pytest.mark.parametrize('name', NAMES)
def test_real_reference_objects_through_the_c_abi_equal_their_own_goldens(reference, name):
    field, steps = scenarios.SCENARIOS[name](reference)
    assert type(field).__module__.startswith('pyfds.')          # the reference's class, not ours
    first = steps // 3
    field.simulate(first)                                        # segmented, as the goldens were made
    field.simulate(steps - first)
    got = scenarios.collect(field)
    with np.load(os.path.join(GOLDEN, name + '.npz')) as golden:
        for key in golden.files:
            if key == 'versions':
                continue
            assert np.array_equal(bits(np.asarray(got[key])), bits(golden[key])), (name, key)
