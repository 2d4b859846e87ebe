"""``fds_simulate`` with its transfers overlapped by row bands (time-skewed launches over row ranges,
DESIGN.md 4.5), forced on grids small enough for the CPU restatement to follow: every model family,
step counts that are and are not multiples of the kernel's steps per launch (the short launch goes
first), 2 to 9 bands -- bitwise against the restatement, fields and probe signals. The full-size
configurations (tests/test_gpu_configs.py) take the same path without being forced."""

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from oracle import restate

pytestmark = pytest.mark.gpu


def _build(kind, steps):
    if kind == 'lossless':
        return scenarios._acoustic2d(fds, lossy=False, nx=256, ny=700, steps=steps, seed=51)[0]
    if kind == 'lossy':
        return scenarios._acoustic2d(fds, lossy=True, nx=192, ny=640, steps=steps, seed=52)[0]
    if kind == 'axi_lossy':
        return scenarios._acoustic2d(fds, lossy=True, nx=192, ny=600, steps=steps, seed=53,
                                     klass='Acoustic3DAxi')[0]
    if kind == 'axi_lossless':
        return scenarios._acoustic2d(fds, lossy=False, nx=256, ny=620, steps=steps, seed=54,
                                     klass='Acoustic3DAxi')[0]
    if kind == 'thermal':
        return scenarios._thermal2d(fds, 'Thermal2D', 256, 660, steps, seed=55)[0]
    raise ValueError(kind)


@pytest.mark.parametrize('kind,steps,bands', [
    ('lossless', 13, 8), ('lossless', 16, 3), ('lossless', 3, 9), ('lossy', 7, 8), ('lossy', 6, 2),
    ('axi_lossy', 5, 4), ('axi_lossless', 9, 5), ('thermal', 11, 6), ('thermal', 8, 8)])
def test_band_pipelined_call_equals_cpu_restatement_bitwise(library, monkeypatch, kind, steps, bands):
    monkeypatch.setenv('FDS_PIPELINE_FORCE', '1')
    monkeypatch.setenv('FDS_PIPELINE_BANDS', str(bands))
    field = _build(kind, 2 * steps)
    stepper = restate.stepper_for(field).run(2 * steps)
    field.simulate(steps)
    engine = field.__dict__['_engine_state'].engine
    assert engine.last_pipeline_bands() >= 2, 'the call was not pipelined'
    field.simulate(steps)                    # same shape again: task tables come from the arena
    assert engine.last_pipeline_bands() >= 2
    got, expected = scenarios.collect(field), scenarios.collect_stepper(stepper)
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), key


def test_pipelined_and_plain_calls_agree(library, monkeypatch):
    """The same field stepped with and without the band pipeline: identical bits."""
    results = []
    for force in (True, False):
        if force:
            monkeypatch.setenv('FDS_PIPELINE_FORCE', '1')
        else:
            monkeypatch.delenv('FDS_PIPELINE_FORCE', raising=False)
            monkeypatch.setenv('FDS_NO_PIPELINE', '1')
        field = _build('lossless', 10)
        field.simulate(10)
        bands = field.__dict__['_engine_state'].engine.last_pipeline_bands()
        assert (bands >= 2) == force
        results.append(scenarios.collect(field))
    for key in results[0]:
        assert np.array_equal(bits(np.asarray(results[0][key])), bits(np.asarray(results[1][key])))
