"""CPU tests of the codeword front end restatement (oracle/pusch_oracle.c: pseudo-random sequence, descrambling, UL-SCH
demultiplexing) against the golden vectors produced by the compiled reference (tests/golden/ref_frontend.npz) and,
where oracle/_ref/libsrsref.so exists, against the reference itself on fresh random configurations."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.vectors import ulsch_case

GOLD = np.load(Path(__file__).parent / "golden" / "ref_frontend.npz")


def test_prg_golden(orc):
    pos = 0
    for c_init, offset, n in GOLD["prg_cases"]:
        nbytes = (int(n) + 7) // 8
        want = np.unpackbits(GOLD["prg_bits"][pos:pos + nbytes])[:n]
        pos += nbytes
        assert (orc.prg_bits(int(c_init), int(offset), int(n)) == want).all()


def test_prg_first_bits_known_answer(orc):
    # TS 38.211 5.2.1 by hand for c_init = 0: x2 stays zero, c(n) = x1(n + 1600).
    x1 = [1] + [0] * 30
    for n in range(1600 + 64):
        x1.append(x1[n + 3] ^ x1[n])
    assert orc.prg_bits(0, 0, 64).tolist() == x1[1600:1664]


def test_ulsch_demux_golden(orc):
    p_llr = p_seq = p_out = 0
    for cfg_arr, lens in zip(GOLD["cfgs"], GOLD["lens"]):
        cfg = dict(zip(po.ULSCH_CFG_FIELDS, (int(v) for v in cfg_arr)))
        n = int(lens[0])
        llr = GOLD["llrs"][p_llr:p_llr + n]
        seq = np.unpackbits(GOLD["seq_bits"][p_seq:p_seq + (n + 7) // 8])[:n]
        p_llr += n
        p_seq += (n + 7) // 8
        rc, outs = orc.ulsch_demux(cfg, llr, seq)
        assert rc == 0
        for k in range(4):
            want = GOLD["outs"][p_out:p_out + int(lens[1 + k])]
            p_out += int(lens[1 + k])
            assert outs[k].size == want.size and (outs[k] == want).all(), (cfg, k)


def test_ulsch_demux_without_uci_is_identity(orc):
    rng = np.random.default_rng(3)
    cfg = dict(qm=8, nof_layers=4, nof_prb=273, start_symbol_index=0, nof_symbols=14, dmrs_type=1,
               dmrs_symbol_mask=1 << 2, nof_cdm_groups_without_data=2)
    n = orc.ulsch_codeword_length(cfg)
    assert n == 1362816  # config 3 of BASELINE.json
    llr = rng.integers(-120, 121, n).astype(np.int8)
    rc, outs = orc.ulsch_demux(cfg, llr, np.zeros(n, np.uint8))
    assert rc == 0 and (outs[0] == llr).all() and all(o.size == 0 for o in outs[1:])


def test_revert_scrambling(orc):
    llr = np.array([5, -5, 0, 120, -120, 127, -127, -128], np.int8)
    seq = np.array([1, 1, 1, 1, 1, 1, 1, 1], np.uint8)
    assert orc.revert_scrambling(llr, seq).tolist() == [-5, 5, 0, -120, 120, -127, 127, -128]
    assert (orc.revert_scrambling(llr, 1 - seq) == llr).all()


@pytest.mark.skipif(not po.Reference.available(), reason="oracle/_ref/libsrsref.so not built")
def test_frontend_vs_reference(orc):
    ref = po.Reference("auto")
    rng = np.random.default_rng(77)
    for c_init, offset, n in [(0, 0, 100), (98765, 0, 5000), (0x7FFFFFFF, 31, 999), (7 << 15 | 3, 250000, 4096)]:
        assert (orc.prg_bits(c_init, offset, n) == ref.prg_bits(c_init, offset, n)).all()
    checked = 0
    while checked < 150:
        cfg, llr, seq, outs = ulsch_case(orc, rng)
        rc, ro = ref.ulsch_demux(cfg, llr, seq, int(rng.choice([0, 0, 7, 50])))
        assert rc == 0
        for a, b in zip(outs, ro):
            assert a.size == b.size and (a == b).all(), cfg
        checked += 1
