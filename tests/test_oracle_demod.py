"""CPU tests of the soft demapper restatement (oracle/pusch_oracle.c: orc_demodulate_soft) against the golden vectors
produced by the compiled reference (tests/golden/ref_demod.npz), against the reference itself where
oracle/_ref/libsrsref.so exists, and against self-checking properties (hard decisions of a clean constellation,
quantiser range, the remainder split of a call)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.vectors import DEMOD_MODS, demod_inputs, modulate

GOLD = np.load(Path(__file__).parent / "golden" / "ref_demod.npz")


def golden_cases():
    p_s = p_l = 0
    for mod, kind, n in GOLD["cases"]:
        mod, n = int(mod), int(n)
        q = max(mod, 1)
        yield mod, int(kind), GOLD["symbols"][p_s:p_s + n], GOLD["noise_vars"][p_s:p_s + n], GOLD["llrs"][p_l:p_l + n * q]
        p_s += n
        p_l += n * q


def test_demod_golden(orc):
    n = 0
    for mod, kind, s, nv, want in golden_cases():
        got = orc.demodulate_soft(s, nv, mod)
        assert (got == want).all(), (mod, kind, s.size)
        n += want.size
    assert n > 20000


def test_demod_vs_reference(orc, ref_available):
    if not ref_available:
        pytest.skip("compiled reference not present")
    ref = po.Reference()
    rng = np.random.default_rng(77)
    for it in range(600):
        mod = int(rng.choice(DEMOD_MODS))
        s, nv = demod_inputs(rng, int(rng.integers(1, 200)), mod, it % 4)
        assert (orc.demodulate_soft(s, nv, mod) == ref.demodulate_soft(s, nv, mod)).all(), (it, mod)


@pytest.mark.parametrize("qm", [2, 4, 6, 8])
def test_hard_decisions_of_a_clean_constellation(orc, qm):
    """Soft bit > 0 means bit 0 (TS 38.211 5.1 mapping): a noiseless constellation demaps to the bits that made it, in
    the SIMD blocks and in the scalar remainder alike."""
    rng = np.random.default_rng(qm)
    bits = rng.integers(0, 2, 203 * qm).astype(np.uint8)
    s = modulate(bits, qm)
    nv = np.full(s.size, 0.05, np.float32)
    for simd in (True, False):
        llr = orc.demodulate_soft(s, nv, qm, simd)
        assert ((llr < 0).astype(np.uint8) == bits).all()
        assert np.abs(llr.astype(int)).max() <= 120


def test_bpsk_and_pi2_bpsk(orc):
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, 101).astype(np.uint8)
    bpsk = ((1 - 2 * bits.astype(np.float64)) * (1 + 1j) / np.sqrt(2)).astype(np.complex64)
    nv = np.full(bits.size, 0.1, np.float32)
    assert ((orc.demodulate_soft(bpsk, nv, 1) < 0).astype(np.uint8) == bits).all()
    pi2 = bpsk.copy()
    pi2[1::2] *= 1j  # odd symbols rotated by +90 degrees at the transmitter
    assert ((orc.demodulate_soft(pi2, nv, 0) < 0).astype(np.uint8) == bits).all()


def test_remainder_split_follows_the_call(orc):
    """The SIMD / scalar split is relative to the call: demapping a call in two pieces is not demapping it whole, but
    equals demapping the pieces (what makes the call boundaries part of the input)."""
    rng = np.random.default_rng(9)
    for mod, block in ((2, 16), (4, 8), (6, 16), (8, 4)):
        s, nv = demod_inputs(rng, 5 * block + 3, mod, 2)
        whole = orc.demodulate_soft(s, nv, mod)
        cut = 2 * block + 1
        pieces = np.concatenate([orc.demodulate_soft(s[:cut], nv[:cut], mod), orc.demodulate_soft(s[cut:], nv[cut:], mod)])
        all_scalar = orc.demodulate_soft(s, nv, mod, simd=False)
        # the last three symbols are the scalar remainder of the whole call
        assert (whole[-3 * mod:] == all_scalar[-3 * mod:]).all()
        assert pieces.size == whole.size


def test_ill_formed_inputs_give_zero(orc):
    """demodulation_mapper.h:56-58: NaN / infinite / negative noise variances and NaN symbols give zero soft bits."""
    for mod in (2, 4, 6, 8):
        n = 32 + 3
        s = np.full(n, 0.3 - 0.2j, np.complex64)
        for bad in (np.nan, -1.0, 0.0, np.inf):
            llr = orc.demodulate_soft(s, np.full(n, bad, np.float32), mod)
            assert not llr.any(), (mod, bad)
        llr = orc.demodulate_soft(np.full(n, np.nan + 0j, np.complex64), np.full(n, 0.1, np.float32), mod)
        assert not llr.reshape(-1, mod)[:, 0::2].any(), mod
