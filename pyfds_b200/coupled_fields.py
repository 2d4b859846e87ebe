"""Preset coupled simulations on the device path: the public surface of ``pyfds/coupled_fields.py``
(``ThermoAcoustic1D``), written against this package's ``SynchronizedFields`` device session."""

from . import acoustics, coupling, thermal

__all__ = ['ThermoAcoustic1D']

#: 1-D helpers of the acoustic field that the preset exposes as its own (pyfds/coupled_fields.py:48-51)
_FORWARDED = ('get_index', 'get_position', 'get_line_region')


class ThermoAcoustic1D(coupling.SynchronizedFields):
    """An ``Acoustic1D`` and a ``Thermal1D`` field on the same line and time axis, stepped together;
    what the sound wave loses to viscosity warms the medium (energy conservation).

    Same constructor, attributes and arithmetic as ``pyfds/coupled_fields.py:10-65``: ``fields[0]`` is
    the acoustic field, ``fields[1]`` the thermal one, and the only interaction is an additive
    ``BoundaryCoupling`` from ``velocity`` to ``temperature`` applied every ``stepping``-th step
    (accumulating in between when ``stepping > 1``). ``simulate()`` keeps both fields on the device and
    moves only ``velocity`` down and ``temperature`` up per application (``coupling.py`` here).
    """

    def __init__(self, x_samples, x_delta, t_samples, t_delta,
                 thermal_material, acoustic_material, stepping=1):
        axes = (x_samples, x_delta, t_samples, t_delta)
        sound = acoustics.Acoustic1D(*axes, acoustic_material)
        heat = thermal.Thermal1D(*axes, thermal_material)
        heating = coupling.BoundaryCoupling(sound.velocity, heat.temperature, self._viscous_heating,
                                            additive=True, accumulate=stepping > 1,
                                            stepping=stepping)
        super().__init__([sound, heat], [heating])
        for name in _FORWARDED:
            setattr(self, name, getattr(sound, name))

    def _viscous_heating(self, velocity):
        """Temperature rise of one time increment caused by the velocity field ``velocity``:
        ``mu / (rho_thermal * c_p) * (dv/dx)**2 * dt``.

        ``dv/dx`` comes from the acoustic field's backward-difference operator ``a_v_p`` (built lazily
        from the baked coefficients), whose ``1 / rho`` factor is cancelled by scaling the velocity
        with the acoustic density first. The operations run in the order of
        ``pyfds/coupled_fields.py:58-65`` -- quotient by quotient, then the square, then ``dt`` -- so
        every sample has the reference's bits.
        """
        sound, heat = self.fields
        gradient = sound.a_v_p.dot(velocity * sound.material_vector('density'))
        gain = sound.material_vector('absorption_coef') / heat.material_vector('density')
        gain = gain / heat.material_vector('heat_capacity')
        return gain * gradient ** 2 * self.t.increment

    #: the reference's name for the transfer function
    _loss_coupling = _viscous_heating
