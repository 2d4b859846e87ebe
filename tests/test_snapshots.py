"""Field snapshots: FrameStream must deliver what the reference's animation loop puts on
its queue (pyfds/gfx.py:72-86: after every `simulate(steps_per_frame)` the time of the last step and the
observed component's values) and leave the field exactly where those `simulate` calls leave it."""

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits


def test_frame_stream_arguments():
    field, _ = scenarios.acoustic2d_lossless(fds)
    with pytest.raises(KeyError):
        fds.FrameStream(field, 'temperature', 5)
    with pytest.raises(ValueError):
        fds.FrameStream(field, 'pressure', 0)
    with pytest.raises(ValueError):
        fds.FrameStream(field, 'pressure', 5, decimate=(0, 1))
    stream = fds.FrameStream(field, 'velocity_y', 7, decimate=3)
    assert (stream.component, stream.stride_x, stream.stride_y) == (2, 3, 3)
    assert stream.num_frames == int(field.t.samples / 7)           # gfx.py:77

    class Custom(fds.Acoustic2D):
        def sim_step(self):
            pass
    custom = Custom(t_delta=1e-7, t_samples=5, x_delta=1e-3, x_samples=8, y_delta=1e-3, y_samples=8,
                    material=fds.AcousticMaterial(1500, 1000))
    with pytest.raises(TypeError):
        fds.FrameStream(custom, 'pressure', 1)


def _reference_messages(builder, component, steps_per_frame, num_frames):
    """The queue messages of Animator._sim_function, produced by plain simulate() calls."""
    field, _ = builder(fds)
    messages = []
    for _ in range(num_frames):
        field.simulate(steps_per_frame)
        messages.append((field.t.vector[field.step - 1],
                         np.array(getattr(field, component).values)))
    return field, messages


@pytest.mark.gpu
@pytest.mark.parametrize('name,component,decimate,lookahead', [
    ('acoustic2d_wide', 'pressure', (1, 1), True),
    ('acoustic2d_wide', 'velocity_x', (4, 3), True),
    ('acoustic2d_lossy', 'velocity_y', (5, 7), False),
    ('thermal2d', 'temperature', (2, 2), True),
    ('acoustic_flow2d_wide', 'pressure', (3, 1), True),
    ('acoustic1d_lossy', 'pressure', (2, 1), True),
])
def test_frames_equal_simulate_loop(library, name, component, decimate, lookahead):
    builder = scenarios.SCENARIOS[name]
    steps_per_frame, num_frames = 7, 6
    twin, messages = _reference_messages(builder, component, steps_per_frame, num_frames)
    field, _ = builder(fds)
    nx = field.x.samples
    ny = field.y.samples if hasattr(field, 'y') else 1
    stream = fds.FrameStream(field, component, steps_per_frame, num_frames, decimate=decimate,
                             lookahead=lookahead)
    count = 0
    for (time, frame), (want_time, want_values) in zip(stream, messages):
        want = want_values.reshape(ny, nx)[::decimate[1], ::decimate[0]]
        assert time == want_time
        assert frame.shape == want.shape
        assert np.array_equal(bits(frame), bits(want)), (name, count)
        count += 1
    assert count == num_frames
    got, expected = scenarios.collect(field), scenarios.collect(twin)
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), key


@pytest.mark.gpu
def test_closing_a_stream_early_leaves_a_consistent_field(library):
    builder = scenarios.acoustic2d_wide
    field, _ = builder(fds)
    stream = fds.FrameStream(field, 'pressure', 5, 8, decimate=(2, 2))
    for k, (time, frame) in enumerate(stream):
        if k == 1:
            break
    stream.close()
    # with lookahead the steps of frame 2 were already enqueued when frame 1 was handed out
    assert field.step == 15
    twin, _ = builder(fds)
    twin.simulate(15)
    got, expected = scenarios.collect(field), scenarios.collect(twin)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), key
