"""CPU emulation of a ``SynchronizedFields`` group (test infrastructure): every member field is stepped
by the CPU restatement (``oracle/restate.py``), the group's own interaction objects -- the classes of
``pyfds_b200/coupling.py`` under test -- are applied to the host arrays in between, exactly as
``SynchronizedFields.sim_step`` prescribes (``pyfds/coupling.py:81-87``). Compared with goldens of the
real reference this pins the host classes; compared with the device session it pins the device path."""

import numpy as np

from oracle import restate


def run_group_on_cpu(group, steps):
    """Advances ``group`` by ``steps`` steps on the CPU; returns the dict ``scenarios.collect_group``
    would give for it."""
    group.assemble_matrices()
    fields = list(group.fields)
    steppers = [restate.stepper_for(field) for field in fields]
    epochs = [field._baked['epoch'] for field in fields]
    for _ in range(steps):
        step = group.step
        for field, stepper in zip(fields, steppers):
            for name in stepper.components:       # interactions rebind / modify the host arrays
                stepper.comp[name].values = np.array(getattr(field, name).values, dtype=np.float64)
            stepper.step = step
            stepper.run(1)
            for name in stepper.components:
                getattr(field, name).values = stepper.values(name).copy()
        for interaction in group.interactions:
            interaction.apply(step)
        for k, (field, stepper) in enumerate(zip(fields, steppers)):
            if field._baked['epoch'] != epochs[k]:     # a MaterialCoupling re-assembled the field
                stepper.assemble(field)
                epochs[k] = field._baked['epoch']
        group.step = step + 1
    out = {}
    for f, (field, stepper) in enumerate(zip(fields, steppers)):
        for name in stepper.components:
            out['field{}/values/{}'.format(f, name)] = np.asarray(getattr(field, name).values,
                                                                  dtype=np.float64)
            for k, signals in enumerate(stepper.signals(name)):
                out['field{}/signals/{}/{}'.format(f, name, k)] = signals
        out['field{}/step'.format(f)] = np.asarray(field.step)
    return out
