"""Preset coupled simulations. Mirror of ``pyfds/coupled_fields.py``."""

from . import acoustics as ac
from . import coupling as cp
from . import thermal as th

__all__ = [
    'ThermoAcoustic1D',
]


class ThermoAcoustic1D(cp.SynchronizedFields):
    """Thermal and acoustic 1-D co-simulation: viscous acoustic losses heat the medium.
    Reference: ``pyfds/coupled_fields.py:10-65``."""

    def __init__(self, x_samples, x_delta, t_samples, t_delta,
                 thermal_material, acoustic_material, stepping=1):
        acoustic_field = ac.Acoustic1D(x_samples, x_delta, t_samples, t_delta, acoustic_material)
        thermal_field = th.Thermal1D(x_samples, x_delta, t_samples, t_delta, thermal_material)

        acoustic_loss = cp.BoundaryCoupling(
            source_component=acoustic_field.velocity,
            target_component=thermal_field.temperature,
            transfer_function=self._loss_coupling,
            additive=True,
            accumulate=True if stepping > 1 else False,
            stepping=stepping
        )
        super().__init__([acoustic_field, thermal_field], [acoustic_loss])

        self.get_index = acoustic_field.get_index
        self.get_position = acoustic_field.get_position
        self.get_line_region = acoustic_field.get_line_region

    def _loss_coupling(self, velocity):
        """Temperature increment from the viscous loss: the spatial derivative of the velocity comes
        from the acoustic field's ``a_v_p`` operator with its density factor removed
        (``pyfds/coupled_fields.py:54-65``; the operator is the lazily built scipy matrix)."""
        velocity_derivative = self.fields[0].a_v_p.dot(
            velocity * self.fields[0].material_vector('density')
        )

        return self.fields[0].material_vector('absorption_coef') \
            / self.fields[1].material_vector('density') \
            / self.fields[1].material_vector('heat_capacity') \
            * velocity_derivative ** 2 * self.t.increment
