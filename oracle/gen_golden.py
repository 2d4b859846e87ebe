"""Generates the golden vectors under ``tests/golden/`` by running the REAL reference
(``/root/reference/pyfds``, unmodified, imported in place) on the scenarios of ``tests/scenarios.py``.

Runs only in the build container (the GPU box has no ``/root/reference``); the resulting ``.npz`` files
are committed. ``import pyfds`` pulls in matplotlib through ``pyfds/gfx.py`` which is not installed,
so empty stand-in modules are registered first -- nothing on the time-stepping path uses them.

    python oracle/gen_golden.py            # all scenarios + region index maps
    python oracle/gen_golden.py NAME ...   # only the named scenarios
"""

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get('PYFDS_REFERENCE', '/root/reference')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    for name in ('matplotlib', 'matplotlib.patches', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REFERENCE)
    import pyfds
    assert os.path.abspath(pyfds.__file__).startswith(os.path.abspath(REFERENCE))
    return pyfds


def main():
    import warnings
    warnings.simplefilter('ignore', DeprecationWarning)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import scenarios
    pyfds = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    import scipy
    versions = 'numpy {} scipy {}'.format(np.__version__, scipy.__version__)

    only = sys.argv[1:]
    for name, builder in scenarios.SCENARIOS.items():
        if only and name not in only:
            continue
        field, steps = builder(pyfds)
        # segmented run: the second call must resume at field.step (pyfds/fields.py:87-93)
        first = steps // 3
        field.simulate(first)
        field.simulate(steps - first)
        data = scenarios.collect(field)
        data['versions'] = np.asarray(versions)
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **data)
        print('{:28s} steps {:4d}  {}'.format(
            name, steps, {k: v.shape for k, v in data.items() if k.startswith('values')}))

    for name, builder in scenarios.COUPLED_SCENARIOS.items():
        if only and name not in only:
            continue
        group, steps = builder(pyfds)
        first = steps // 3
        group.simulate(first)
        group.simulate(steps - first)
        data = scenarios.collect_group(group)
        data['versions'] = np.asarray(versions)
        np.savez_compressed(os.path.join(GOLDEN, 'coupled_' + name + '.npz'), **data)
        print('{:28s} steps {:4d}  {}'.format(
            name, steps, {k: v.shape for k, v in data.items() if '/values/' in k}))

    if only:
        return
    regions = {k: np.asarray(r.indices, dtype=np.int64)
               for k, r in scenarios.region_cases(pyfds).items()}
    np.savez_compressed(os.path.join(GOLDEN, 'regions.npz'), **regions)
    print('regions', {k: v.shape[0] for k, v in regions.items()})


if __name__ == '__main__':
    main()
