"""Coupled fields: lock-step simulation of several fields and interactions between them.

Mirror of the reference interface ``pyfds/coupling.py``. The interactions are arbitrary Python
transfer functions of one component's ``values``, so they stay on the host; what changes is how much
travels for them:

* ``SynchronizedFields.sim_step`` is the reference statement (``pyfds/coupling.py:81-87``): every
  field's ``sim_step()`` (one device step each, host ``values`` coherent before and after), then the
  interactions on the host arrays.
* ``SynchronizedFields.simulate`` runs a *device session* when every field is on the device path and
  every interaction is a plain ``BoundaryCoupling``: the state of all fields stays in HBM for the whole
  call, boundary / probe tables are uploaded once, and per step only the components an interaction
  actually touches cross the bus -- the source when the interaction reads it (every step if it
  accumulates, else every ``stepping``-th), the target when it is written. Anything else
  (``MaterialCoupling``, subclasses with their own ``sim_step``/``apply``) falls back to the per-step
  loop. Set ``device_session = False`` on the instance if a transfer function looks at components
  other than the source it is given.
"""

import logging as lo

import numpy as np

from . import fields as fld

__all__ = [
    'SynchronizedFields', 'BoundaryCoupling', 'MaterialCoupling', 'MaterialCouplingExponential',
    'MaterialCouplingPowerLaw',
]

logger = lo.getLogger('pyfds')


class SynchronizedFields(fld.Field):
    """Several fields with the same time stepping, stepped together.
    Reference: ``pyfds/coupling.py:14-87``."""

    def __init__(self, fields, interactions):
        self.fields = fields
        self.interactions = interactions

        # expose the components of all fields as attributes (the reference does this for gfx)
        for field in self.fields:
            for name, value in vars(field).items():
                if isinstance(value, fld.FieldComponent):
                    if not hasattr(self, name):
                        self.__setattr__(name, value)
                    else:
                        raise RuntimeError("Coupling of fields with identically named components "
                                           "is currently not possible")

        self.t = self.fields[0].t
        self.x = self.fields[0].x
        if hasattr(self.fields[0], 'y'):
            self.y = self.fields[0].y

    @property
    def step(self):
        return self.fields[0].step

    @step.setter
    def step(self, value):
        for field in self.fields:
            field.step = value

    @property
    def num_points(self):
        return self.fields[0].num_points

    @property
    def material_regions(self):
        regions = []
        for field in self.fields:
            regions += field.material_regions
        return regions

    def assemble_matrices(self):
        for field in self.fields:
            field.assemble_matrices()

    @property
    def matrices_assembled(self):
        return all([field.matrices_assembled for field in self.fields])

    def sim_step(self):
        for field in self.fields:
            field.sim_step()
        for interaction in self.interactions:
            interaction.apply(self.step)

    # ---- device session -------------------------------------------------------------------------

    #: allow ``simulate`` to keep the fields' state on the device between steps (module docstring)
    device_session = True

    def _session_plan(self):
        """``[(interaction, source key, target key)]`` with keys ``(field number, component number)``
        if the device session applies, else ``None``."""
        if not self.device_session or type(self).sim_step is not SynchronizedFields.sim_step:
            return None
        owners = {}
        for f, field in enumerate(self.fields):
            if not field._uses_device():
                return None
            for c, name in enumerate(field._device_components):
                owners[id(getattr(field, name))] = (f, c)
        plan = []
        for interaction in self.interactions:
            if type(interaction) is not BoundaryCoupling:
                return None
            source = owners.get(id(interaction.source_component))
            target = owners.get(id(interaction.target_component))
            if source is None or target is None:
                return None
            plan.append((interaction, source, target))
        return plan

    def _simulate_on_device(self, num_steps, progress_logger=None):
        """``num_steps`` x ``sim_step`` with the state of all fields resident on the device; returns
        ``False`` (and does nothing) if the session does not apply."""
        plan = self._session_plan()
        if plan is None:
            return False
        from . import _engine
        first_step = self.step
        engines, tables, components = [], [], []
        for field in self.fields:
            if not field.matrices_assembled:
                field.assemble_matrices()
            engine = _engine.prepare(field)
            engines.append(engine)
            tables.append(_engine.upload_run_tables(field, engine, first_step, num_steps))
            components.append(_engine._components(field))
            _engine.upload_values(field, engine)
        # host copy of (field, component) equals the device copy
        fresh = {(f, c): True for f in range(len(self.fields))
                 for c in range(len(components[f]))}

        def to_host(key):
            if not fresh[key]:
                f, c = key
                _engine.download_values(self.fields[f], engines[f], only=(c,))
                fresh[key] = True

        try:
            for step in range(first_step, first_step + num_steps):
                for f, engine in enumerate(engines):
                    n_slots, layout = tables[f]
                    records = engine.step(step, 1, n_slots)
                    if n_slots:
                        _engine._append_signals(layout, records)
                    for c in range(len(components[f])):
                        fresh[(f, c)] = False
                for interaction, source, target in plan:
                    writes = step % interaction.stepping == 0
                    if writes or interaction.accumulate is True:
                        to_host(source)
                    if writes and interaction.additive is True:
                        to_host(target)
                    interaction.apply(step)
                    if writes:
                        f, c = target
                        engines[f].upload_state(c, _engine._host_values(components[f][c],
                                                                        self.fields[f].num_points))
                        fresh[target] = True
                if progress_logger:
                    progress_logger.log(step)
                self.step = step + 1
        finally:
            for key in fresh:
                to_host(key)
        return True


class BoundaryCoupling():
    """Feeds a function of one component into another component, every ``stepping``-th step,
    optionally accumulating in between. Reference: ``pyfds/coupling.py:90-140``."""

    def __init__(self, source_component, target_component, transfer_function,
                 additive=True, accumulate=False, stepping=1):
        self.source_component = source_component
        self.target_component = target_component
        self.transfer_function = transfer_function
        self.additive = additive
        self.accumulate = accumulate
        self.stepping = stepping

        self.accumulated_transfer = 0

    def apply(self, step):
        if self.accumulate is True:
            self.accumulated_transfer += self.transfer_function(self.source_component.values)

        if step % self.stepping == 0:
            if self.accumulate is False:
                transfer = self.transfer_function(self.source_component.values)
            else:
                transfer = self.accumulated_transfer
                self.accumulated_transfer = 0
            if self.additive is True:
                self.target_component.values += transfer
            else:
                self.target_component.values = transfer


class MaterialCoupling():
    """Scales one material parameter of a field by a function of another field's component and
    re-assembles the target field when the factors changed enough.
    Reference: ``pyfds/coupling.py:143-215``.

    The target field's ``material_vector`` is replaced on the instance, as in the reference; the device
    engine then bakes its material ids from those per-point vectors (distinct value combinations become
    materials, at most 31 of them -- see ``_bake.DenseSnapshot``)."""

    def __init__(self, source_component, target_field, target_parameter,
                 transfer_function, rel_change_threshold=None, stepping=1):
        self.source_component = source_component
        self.target_field = target_field
        self.target_parameter = target_parameter
        self.transfer_function = transfer_function
        self.rel_change_threshold = rel_change_threshold
        self.stepping = stepping

        self.last_used_factors = 0

        self.target_field.static_material_vector = self.target_field.material_vector
        self.target_field.material_vector = self._material_vector

    def _material_vector(self, mat_parameter):
        if mat_parameter == self.target_parameter:
            return self.target_field.static_material_vector(mat_parameter) \
                * self.transfer_function(self.source_component.values)
        return self.target_field.static_material_vector(mat_parameter)

    def apply(self, step):
        if step % self.stepping == 0:
            transfer_factors = self.transfer_function(self.source_component.values)
            rel_change = max(abs((transfer_factors - self.last_used_factors) / transfer_factors))
            if self.rel_change_threshold is None or rel_change > self.rel_change_threshold:
                self.target_field.assemble_matrices()
                self.last_used_factors = transfer_factors
                if self.rel_change_threshold is not None:
                    logger.info(f"Relative change in parameters is {rel_change}.")
                    logger.info(f"Matrices reassembled in step {step}.")


class MaterialCouplingExponential(MaterialCoupling):
    """``p * (a + (1 - a) * exp(b * q))``. Reference: ``pyfds/coupling.py:218-259``."""

    def __init__(self, source_component, target_field, target_parameter, a, b,
                 rel_change_threshold=None, stepping=1):
        self.a = a
        self.b = b
        super().__init__(source_component=source_component, target_field=target_field,
                         target_parameter=target_parameter,
                         transfer_function=self.transfer_function,
                         rel_change_threshold=rel_change_threshold, stepping=stepping)

    def transfer_function(self, values):
        return self.a + (1 - self.a) * np.exp(self.b * values)


class MaterialCouplingPowerLaw(MaterialCoupling):
    """``p * (1 + factor * q ** power)``. Reference: ``pyfds/coupling.py:262-300``."""

    def __init__(self, source_component, target_field, target_parameter, power, factor,
                 rel_change_threshold=None, stepping=1):
        self.power = power
        self.factor = factor
        super().__init__(source_component=source_component, target_field=target_field,
                         target_parameter=target_parameter,
                         transfer_function=self.transfer_function,
                         rel_change_threshold=rel_change_threshold, stepping=stepping)

    def transfer_function(self, values):
        return 1 + self.factor * values ** self.power
