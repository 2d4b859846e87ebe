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


class linear:
    """Transfer function ``values -> scale * values`` for a ``BoundaryCoupling``. Behaves like the
    lambda it replaces; being an object, it lets ``SynchronizedFields.simulate`` see what the coupling
    computes and run it on the device (an extension: the reference takes any callable, and so does this
    package -- an opaque callable is simply evaluated on the host)."""

    def __init__(self, scale):
        self.scale = scale

    def __call__(self, values):
        return self.scale * values

logger = lo.getLogger('pyfds')


class SynchronizedFields(fld.Field):
    """Several fields on the same axes and time base, advanced in lock step, with ``interactions``
    applied after every common step. Public surface of ``pyfds/coupling.py:14-87``: ``fields``,
    ``interactions``, the axes ``t`` / ``x`` (/ ``y``) of the first field, every field's components as
    attributes of the group, ``step`` / ``num_points`` / ``material_regions`` / ``matrices_assembled``
    as views of the members."""

    def __init__(self, fields, interactions):
        self.fields = fields
        self.interactions = interactions

        # The components of all members become attributes of the group (animators and user scripts
        # address ``group.pressure``); two members with a component of the same name cannot be told
        # apart that way, which the reference refuses with this very error.
        for member in fields:
            components = {name: value for name, value in vars(member).items()
                          if isinstance(value, fld.FieldComponent)}
            clash = [name for name in components if hasattr(self, name)]
            if clash:
                raise RuntimeError("Coupling of fields with identically named components "
                                   "is currently not possible")
            vars(self).update(components)

        leader = fields[0]
        self.t, self.x = leader.t, leader.x
        if hasattr(leader, 'y'):
            self.y = leader.y

    # the group has no state of its own: these are views of its members
    step = property(lambda self: self.fields[0].step)

    @step.setter
    def step(self, value):
        for member in self.fields:
            member.step = value

    num_points = property(lambda self: self.fields[0].num_points)
    material_regions = property(
        lambda self: [region for member in self.fields for region in member.material_regions])
    matrices_assembled = property(
        lambda self: all(member.matrices_assembled for member in self.fields))

    def assemble_matrices(self):
        for member in self.fields:
            member.assemble_matrices()

    def sim_step(self):
        """One common step (``pyfds/coupling.py:81-87``): every member's own ``sim_step`` -- one device
        step each, host ``values`` coherent afterwards -- then the interactions in list order."""
        for member in self.fields:
            member.sim_step()
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

    def _group_plan(self):
        """What the device needs to run every interaction itself (``fds_group_*``): a list of
        ``(kind, interaction, ...)`` if all members are 1-D device fields and every interaction is one
        the engine knows -- a ``BoundaryCoupling`` with a ``linear`` transfer function, the viscous
        heating of ``ThermoAcoustic1D``, or one of the two built-in material laws -- else ``None``."""
        if not self.device_session or type(self).sim_step is not SynchronizedFields.sim_step:
            return None
        owners = {}
        for f, field in enumerate(self.fields):
            if getattr(field, '_device_model', None) not in ('acoustic1d', 'thermal1d') or \
                    not field._uses_device() or getattr(field, 'device_kernel', 0) == 1:
                return None
            for c, name in enumerate(field._device_components):
                owners[id(getattr(field, name))] = (f, c)
        if len({field.num_points for field in self.fields}) != 1:
            return None
        from . import coupled_fields
        plan, law_targets, heated = [], set(), set()
        for interaction in self.interactions:
            kind = type(interaction)
            if kind is BoundaryCoupling:
                source = owners.get(id(interaction.source_component))
                target = owners.get(id(interaction.target_component))
                if source is None or target is None or int(interaction.stepping) < 1 or \
                        interaction.additive not in (True, False) or \
                        interaction.accumulate not in (True, False):
                    return None
                function = interaction.transfer_function
                if isinstance(function, linear) and np.ndim(function.scale) == 0:
                    plan.append(('linear', interaction, source, target))
                elif getattr(function, '__func__', None) is \
                        coupled_fields.ThermoAcoustic1D._viscous_heating and \
                        function.__self__ is self and source == (0, 1) and target == (1, 0):
                    plan.append(('heating', interaction, source, target))
                    heated.add(0)
                else:
                    return None
            elif kind in (MaterialCouplingExponential, MaterialCouplingPowerLaw) and \
                    getattr(interaction.transfer_function, '__func__', None) is \
                    kind.transfer_function:
                source = owners.get(id(interaction.source_component))
                targets = [f for f, field in enumerate(self.fields)
                           if field is interaction.target_field]
                if source is None or not targets or targets[0] in law_targets or \
                        int(interaction.stepping) < 1:
                    return None
                field = self.fields[targets[0]]
                if interaction.target_parameter not in field._material_params or \
                        vars(field).get('material_vector') != interaction._material_vector:
                    return None          # somebody else shadows material_vector as well
                law_targets.add(targets[0])
                plan.append(('law', interaction, source, targets[0]))
            else:
                return None
        if law_targets & heated:
            return None      # the heating term reads the operators of a field that is re-assembled
        return plan

    def _simulate_on_group(self, plan, num_steps, progress_logger=None):
        """``num_steps`` x ``sim_step`` entirely on the device: members and interactions as kernels on
        one stream (``fds_group_step``), values and probe records cross the bus once per call."""
        from . import _bake, _engine
        first_step = self.step
        law_targets = {entry[3]: entry[1] for entry in plan if entry[0] == 'law'}
        engines, layouts = [], []
        for f, field in enumerate(self.fields):
            if not field.matrices_assembled:
                field.assemble_matrices()
            law = law_targets.get(f)
            lossy = law is not None and law.target_parameter == 'absorption_coef'
            engine = _engine.prepare(field, device=int(getattr(field, 'device_index', 0)),
                                     per_cell=law is not None, lossy=lossy or None)
            engines.append(engine)
            layouts.append(_engine.upload_run_tables(field, engine, first_step, num_steps))
            _engine.upload_values(field, engine)
        group = _engine.Group(engines)
        try:
            for kind, interaction, source, target in plan:
                if kind == 'linear':
                    group.add_linear(source, target, interaction.transfer_function.scale,
                                     interaction.additive, interaction.accumulate,
                                     interaction.stepping, interaction.accumulated_transfer)
                elif kind == 'heating':
                    sound, heat = self.fields
                    density = sound.material_vector('density')
                    _, gradient_factor, _ = sound._factors(
                        sound.material_vector('sound_velocity'), density,
                        sound.material_vector('absorption_coef'))
                    gain = sound.material_vector('absorption_coef') / \
                        heat.material_vector('density') / heat.material_vector('heat_capacity')
                    group.add_viscous_heating(0, 1, density, gradient_factor, gain, self.t.increment,
                                              interaction.accumulate, interaction.stepping,
                                              interaction.accumulated_transfer)
                else:
                    field = self.fields[target]
                    statics = np.stack([np.asarray(field.static_material_vector(p), dtype=np.float64)
                                        for p in field._material_params])
                    dt, dx = field.t.increment, field.x.increment
                    if type(interaction) is MaterialCouplingExponential:
                        law, p0, p1 = 0, interaction.a, interaction.b
                    else:
                        law, p0, p1 = 1, interaction.factor, interaction.power
                    group.add_material_law(
                        source, target, field._material_params.index(interaction.target_parameter),
                        law, p0, p1, interaction.rel_change_threshold, interaction.stepping, statics,
                        interaction.last_used_factors, [dt / dx, dt / dx ** 2, 1 / dx])
            slots = [n_slots for n_slots, _ in layouts]
            chunk = num_steps if progress_logger is None else max(1, -(-num_steps // 20))
            done = 0
            while done < num_steps:
                count = min(chunk, num_steps - done)
                for (n_slots, layout), records in zip(
                        layouts, group.step(first_step + done, count, slots)):
                    if n_slots:
                        _engine._append_signals(layout, records)
                if progress_logger is not None:
                    for s in range(first_step + done, first_step + done + count):
                        progress_logger.log(s)
                done += count
            for field, engine in zip(self.fields, engines):
                _engine.download_values(field, engine)
            # what the interactions carry from one call to the next
            for number, (kind, interaction, source, target) in enumerate(plan):
                values, count = group.read(number)
                if kind == 'law':
                    if count:
                        field = self.fields[target]
                        interaction.last_used_factors = values
                        snapshot = field._baked['snapshot']
                        if isinstance(snapshot, _bake.DenseSnapshot):
                            name = interaction.target_parameter
                            snapshot.vectors[name] = np.asarray(
                                field.static_material_vector(name), dtype=np.float64) * values
                        field._baked['epoch'] += 1
                        field._operators = {}
                elif interaction.accumulate is True:
                    interaction.accumulated_transfer = values if values.any() else 0
        finally:
            group.close()
        self.step = first_step + num_steps
        return True

    def _simulate_on_device(self, num_steps, progress_logger=None):
        """``num_steps`` x ``sim_step`` with the state of all fields resident on the device; returns
        ``False`` (and does nothing) if no device session applies."""
        group_plan = self._group_plan()
        if group_plan is not None:
            self._last_session = 'device'
            return self._simulate_on_group(group_plan, num_steps, progress_logger)
        plan = self._session_plan()
        if plan is None:
            self._last_session = 'per step'
            return False
        self._last_session = 'host interactions'
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


class BoundaryCoupling:
    """``target.values (+)= f(source.values)`` every ``stepping``-th step; with ``accumulate`` the
    function is evaluated after every step and the sum since the last delivery is handed over.
    Attributes and behaviour of ``pyfds/coupling.py:90-140`` (``additive`` / ``accumulate`` are compared
    by identity with ``True`` / ``False`` there, so that is what decides here as well)."""

    def __init__(self, source_component, target_component, transfer_function,
                 additive=True, accumulate=False, stepping=1):
        self.source_component, self.target_component = source_component, target_component
        self.transfer_function = transfer_function
        self.additive, self.accumulate, self.stepping = additive, accumulate, stepping
        self.accumulated_transfer = 0      # sum of f(source) since the last delivery

    def _evaluate(self):
        return self.transfer_function(self.source_component.values)

    def apply(self, step):
        if self.accumulate is True:
            self.accumulated_transfer += self._evaluate()
        if step % self.stepping:
            return
        if self.accumulate is False:
            delivery = self._evaluate()
        else:
            delivery, self.accumulated_transfer = self.accumulated_transfer, 0
        target = self.target_component
        if self.additive is True:
            target.values += delivery      # in place, like the reference: views stay valid
        else:
            target.values = delivery


class MaterialCoupling:
    """Scales one material parameter of ``target_field`` point by point with a function of another
    field's component, and re-assembles the target when the factors have moved by more than
    ``rel_change_threshold`` (always, if that is ``None``). Public surface of
    ``pyfds/coupling.py:143-215``.

    As in the reference the target's ``material_vector`` is shadowed on the instance (the original stays
    reachable as ``static_material_vector``), so whoever assembles the target sees the scaled parameter;
    the device engine bakes its material ids from those per-point vectors (distinct value combinations
    become materials, at most 31 of them -- see ``_bake.DenseSnapshot``)."""

    def __init__(self, source_component, target_field, target_parameter,
                 transfer_function, rel_change_threshold=None, stepping=1):
        self.source_component, self.target_field = source_component, target_field
        self.target_parameter, self.transfer_function = target_parameter, transfer_function
        self.rel_change_threshold, self.stepping = rel_change_threshold, stepping
        self.last_used_factors = 0         # factors of the last re-assembly
        target_field.static_material_vector = target_field.material_vector
        target_field.material_vector = self._material_vector

    def _factors(self):
        return self.transfer_function(self.source_component.values)

    def _material_vector(self, mat_parameter):
        plain = self.target_field.static_material_vector(mat_parameter)
        return plain * self._factors() if mat_parameter == self.target_parameter else plain

    def apply(self, step):
        if step % self.stepping:
            return
        factors = self._factors()
        rel_change = max(abs((factors - self.last_used_factors) / factors))
        threshold = self.rel_change_threshold
        if threshold is not None and not rel_change > threshold:
            return
        self.target_field.assemble_matrices()
        self.last_used_factors = factors
        if threshold is not None:
            logger.info(f"Relative change in parameters is {rel_change}.")
            logger.info(f"Matrices reassembled in step {step}.")


class _ParametricMaterialCoupling(MaterialCoupling):
    """A ``MaterialCoupling`` whose transfer function is a method of the instance."""

    def __init__(self, source_component, target_field, target_parameter, rel_change_threshold,
                 stepping):
        super().__init__(source_component, target_field, target_parameter, self.transfer_function,
                         rel_change_threshold=rel_change_threshold, stepping=stepping)


class MaterialCouplingExponential(_ParametricMaterialCoupling):
    """Factor ``a + (1 - a) * exp(b * q)`` of the source values ``q``
    (``pyfds/coupling.py:218-259``)."""

    def __init__(self, source_component, target_field, target_parameter, a, b,
                 rel_change_threshold=None, stepping=1):
        self.a, self.b = a, b
        super().__init__(source_component, target_field, target_parameter, rel_change_threshold,
                         stepping)

    def transfer_function(self, values):
        return self.a + (1 - self.a) * np.exp(self.b * values)


class MaterialCouplingPowerLaw(_ParametricMaterialCoupling):
    """Factor ``1 + factor * q ** power`` of the source values ``q``
    (``pyfds/coupling.py:262-300``)."""

    def __init__(self, source_component, target_field, target_parameter, power, factor,
                 rel_change_threshold=None, stepping=1):
        self.power, self.factor = power, factor
        super().__init__(source_component, target_field, target_parameter, rel_change_threshold,
                         stepping)

    def transfer_function(self, values):
        return 1 + self.factor * values ** self.power
