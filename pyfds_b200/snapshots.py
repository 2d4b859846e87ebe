"""Field snapshots while a simulation runs: the output side of the hot path.

The reference animates a field by running ``field.simulate(steps_per_frame)`` in a child process and
pickling the observed component's whole ``values`` array through a queue after every frame
(``Animator._sim_function``, ``pyfds/gfx.py:72-86``). ``FrameStream`` produces the same sequence of
messages -- ``(time, values)`` after every ``steps_per_frame`` steps -- from the device: the state stays
in HBM for the whole sequence, a frame is gathered on the device (optionally every ``decimate``-th sample
per axis, which is what a plot can show anyway), and travels to page-locked host memory on a second
stream while the next frame's steps are computed.

    for time, frame in FrameStream(field, 'pressure', steps_per_frame=10, decimate=(4, 4)):
        image.set_data(frame)               # frame.shape == (ceil(ny / 4), ceil(nx / 4))

After the stream is exhausted (or closed) the field is exactly where ``simulate(steps_per_frame)``
called once per frame would have left it: component ``values``, ``step`` and all ``Output.signals``.
"""

import numpy as np

__all__ = ['FrameStream']


class FrameStream:
    """Iterator over ``(time, frame)`` pairs of one component of a field on the CUDA hot path.

    ``time`` is ``field.t.vector[step - 1]`` as in ``pyfds/gfx.py:80``; ``frame`` is a new
    ``(rows, columns)`` array (1-D fields: one row) holding every ``decimate[1]``-th row and every
    ``decimate[0]``-th column of the component after the frame's last step.

    ``num_frames`` defaults to ``int(t.samples / steps_per_frame)`` (``pyfds/gfx.py:77``).
    ``lookahead=True`` enqueues the steps of the next frame before waiting for the current one, so
    that the copy to the host overlaps the computation; the field's boundaries and outputs must then
    not be modified between frames (the reference's child process does not see such changes either).
    """

    def __init__(self, field, observed_component, steps_per_frame, num_frames=None,
                 decimate=(1, 1), lookahead=True):
        if not field._uses_device():
            raise TypeError('{} is stepped through a Python sim_step() override; snapshots need a '
                            'model on the device path.'.format(type(field).__name__))
        if observed_component not in field._device_components:
            raise KeyError('{} has no component {!r}.'.format(type(field).__name__,
                                                             observed_component))
        steps_per_frame = int(steps_per_frame)
        if steps_per_frame < 1:
            raise ValueError('steps_per_frame must be at least 1.')
        if np.ndim(decimate) == 0:
            decimate = (decimate, decimate)
        self.stride_x, self.stride_y = int(decimate[0]), int(decimate[1])
        if self.stride_x < 1 or self.stride_y < 1:
            raise ValueError('decimate must be at least 1 along each axis.')
        self.field = field
        self.component = field._device_components.index(observed_component)
        self.steps_per_frame = steps_per_frame
        self.num_frames = int(field.t.samples / steps_per_frame) if num_frames is None \
            else int(num_frames)
        self.lookahead = bool(lookahead)
        self._generator = None

    def __iter__(self):
        if self._generator is None:
            self._generator = self._run()
        return self._generator

    def __next__(self):
        return next(iter(self))

    def close(self):
        """Stops early; the field is left after the last frame whose steps were enqueued."""
        if self._generator is not None:
            self._generator.close()

    def _run(self):
        from . import _engine
        field = self.field
        if not field.matrices_assembled:
            field.assemble_matrices()
        engine = _engine.prepare(field)
        total = self.num_frames * self.steps_per_frame
        first_step = field.step
        n_slots, layout = _engine.upload_run_tables(field, engine, first_step, total)
        _engine.upload_values(field, engine)
        shape = engine.frame_shape(self.stride_x, self.stride_y)

        def advance(frame):
            records = engine.step(first_step + frame * self.steps_per_frame, self.steps_per_frame,
                                  n_slots)
            if n_slots:
                _engine._append_signals(layout, records)
            field.step += self.steps_per_frame
            engine.snapshot_async(self.component, self.stride_x, self.stride_y, frame % 2)
            return field.t.vector[field.step - 1]

        try:
            pending = None                      # (frame number, time) enqueued but not yet yielded
            for frame in range(self.num_frames):
                time = advance(frame)
                if self.lookahead:
                    if pending is not None:
                        yield pending[1], engine.snapshot_wait(pending[0] % 2, shape)
                    pending = (frame, time)
                else:
                    yield time, engine.snapshot_wait(frame % 2, shape)
            if pending is not None:
                yield pending[1], engine.snapshot_wait(pending[0] % 2, shape)
        finally:
            _engine.download_values(field, engine)
