"""The quotient sequence of the axisymmetric streaming kernel (``sv_quotient``,
pyfds_b200/csrc/fds_streamv.cuh) against the IEEE division the reference uses
(numpy ``/``, pyfds/acoustics.py:210-212): ``oracle/quotient_check.c`` runs both on the CPU with the
same operations (one rounded multiply, two fused multiply-adds) over the ranges the kernel admits."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SOURCE = os.path.join(ROOT, 'oracle', 'quotient_check.c')


def _build():
    from oracle import cbuild
    try:
        return cbuild.build_quotient_check()
    except RuntimeError as error:
        if 'gcc not found' in str(error):
            pytest.skip('gcc not found')
        raise


def test_quotient_sequence_equals_ieee_division():
    binary = _build()
    result = subprocess.run([binary, '2000000'], capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout
    cases, mismatches = [int(result.stdout.split()[k]) for k in (-3, -1)]
    assert cases > 30_000_000 and mismatches == 0, result.stdout


def test_admitted_ranges_match_the_kernel_constants():
    """The limits the C check draws numerators and divisors from are the ones in the kernel."""
    with open(os.path.join(ROOT, 'pyfds_b200', 'csrc', 'fds_streamv.cuh')) as handle:
        kernel = handle.read()
    assert 'kDivHiLo = (1023u - 493u) << 20' in kernel
    assert 'kDivHiSpan = (493u + 477u) << 20' in kernel
    assert 'sv_moderate(k.rr[c], 256u)' in kernel and 'sv_moderate(k.eb[c], 200u)' in kernel
    with open(SOURCE) as handle:
        check = handle.read()
    # |a| = eb |vx|: 2^-200 * 2^-493 and 2^200 * 2^477
    assert '(1023 - 693) + rnd() % (693 + 677 + 1)' in check
    assert '1023 - 256 + rnd() % 513' in check
