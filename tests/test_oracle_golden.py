"""CPU tests: the oracle (oracle/pusch_oracle.c) against the committed golden fixtures of tests/golden/.

The fixtures were produced by tests/golden/make_golden.py from (a) the reference tree's own on-disk LDPC vectors and
(b) the compiled reference itself (oracle/_ref/libsrsref.so), so these tests pin the oracle where /root/reference does
not exist."""
import zlib
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.vectors import LIFTING_SIZES

G = Path(__file__).resolve().parent / "golden"


def test_examples_encode_and_decode(orc):
    # Semantics of the reference's ldpc_enc_dec_test.cpp:237-315: noiseless LLR = +-10 (fillers +10 -> here +127 like the
    # dematcher), ONE iteration, no CRC, decoded bits == message, for every base graph and lifting size.
    ex = np.load(G / "ldpc_examples.npz")
    for bg in (1, 2):
        for Z in LIFTING_SIZES:
            K = (22 if bg == 1 else 10) * Z
            N = (66 if bg == 1 else 50) * Z
            msgs = np.unpackbits(ex[f"bg{bg}_z{Z}_msg"], axis=1)[:, :K]
            cws = np.unpackbits(ex[f"bg{bg}_z{Z}_cw"], axis=1)[:, :N]
            for m, c, F in zip(msgs, cws, ex[f"bg{bg}_z{Z}_filler"]):
                assert (orc.ldpc_encode(bg, Z, m) == c).all(), (bg, Z)
                llr = (10 - 20 * c.astype(np.int16)).astype(np.int8)
                if F:
                    llr[K - 2 * Z - F:K - 2 * Z] = 127
                for n in sorted({K + 2 * Z, (K + 2 * Z + N) // 2, N}):
                    it, bits = orc.ldpc_decode(bg, Z, llr[:n], int(F), po.CRC_NONE, 1)
                    assert it == 0 and (np.unpackbits(bits)[:K] == m).all(), (bg, Z, n)


@pytest.mark.parametrize("variant,scale", [("auto", po.SCALE_X86), ("generic", po.SCALE_GENERIC)])
def test_decoder_vs_reference_vectors(orc, variant, scale):
    d = np.load(G / "ref_decoder.npz")
    off = boff = 0
    n_pass = 0
    for (bg, Z, F, crc_kind, mi, n), it_ref in zip(d["cases"], d[f"iters_{variant}"]):
        K = (22 if bg == 1 else 10) * Z
        kb = (K + 7) // 8
        llr = d["llrs"][off:off + n]
        bits_ref = d[f"bits_{variant}"][boff:boff + kb]
        off += n
        boff += kb
        it, bits = orc.ldpc_decode(int(bg), int(Z), llr, int(F), int(crc_kind), int(mi), scale)
        assert it == it_ref and (bits == bits_ref).all(), (bg, Z, F, crc_kind, mi, n)
        n_pass += it > 0
    assert n_pass > 10  # the vectors exercise the early stop


def test_dematcher_vs_reference_vectors(orc):
    d = np.load(G / "ref_dematcher.npz")
    width = int(d["simd_width"])
    bo = lo = 0
    for N, E, new_data, rv, qm, nref, F in d["cases"]:
        buf = d["buf0"][bo:bo + N].copy()
        orc.rate_dematch(buf, d["llrs"][lo:lo + E], bool(new_data), int(rv), int(qm), int(nref), int(F), width)
        assert (buf == d["buf1"][bo:bo + N]).all(), (N, E, new_data, rv, qm, nref, F)
        bo += N
        lo += E


def test_pusch_decoder_vs_reference_vectors(orc):
    d = np.load(G / "ref_pusch.npz")
    to = lo = so = oo = 0
    n_ok = 0
    for bg, tb_bytes, qm, nl, n_llr, nref, es, mi, fill, C in d["cases"]:
        tb = d["tbs"][to:to + tb_bytes]
        to += tb_bytes
        harq = po.Harq(int(C), int(fill))
        N = orc.segment_rx(int(tb_bytes) * 8, int(bg), int(qm), int(nl), int(n_llr))[0].full_length
        for t, rv in enumerate([0, 2, 3, 1]):
            llr = d["llrs"][lo:lo + n_llr]
            lo += n_llr
            out, st = orc.pusch_decode(harq, llr, int(tb_bytes), int(bg), rv, int(qm), int(nref), int(nl), int(mi),
                                       bool(es), t == 0)
            assert (st[:5] == d["stats"][so]).all(), (bg, tb_bytes, t, st, d["stats"][so])
            if st[0]:
                assert (out == d["tb_out"][oo:oo + tb_bytes]).all() and (out == tb).all()
                n_ok += 1
            crc = [zlib.crc32(harq.soft[cb][:N].tobytes()) for cb in range(C)]
            assert crc == list(d["soft_crc32"][so][:C])
            so += 1
            oo += tb_bytes
    assert n_ok > 5


def test_crc_known_answers(orc):
    # Self-checking: appending the CRC makes the remainder zero (crc_calculator_test.cpp idea), plus fixed values.
    rng = np.random.default_rng(0)
    for kind, bits in ((po.CRC16, 16), (po.CRC24A, 24), (po.CRC24B, 24)):
        for n in (1, 8, 40, 333, 8424):
            msg = rng.integers(0, 2, n).astype(np.uint8)
            c = orc.crc(kind, np.packbits(msg), n)
            ext = np.concatenate([msg, [(c >> (bits - 1 - i)) & 1 for i in range(bits)]]).astype(np.uint8)
            assert orc.crc(kind, np.packbits(ext), n + bits) == 0
    assert orc.crc(po.CRC24A, np.array([0x80], np.uint8), 1) == 0x864CFB
    assert orc.crc(po.CRC24B, np.array([0x80], np.uint8), 1) == 0x800063
    assert orc.crc(po.CRC16, np.array([0x80], np.uint8), 1) == 0x1021


def test_llr_edge_cases(orc):
    # All-zero input is not decodable; without CRC the output is all ones (ldpc_enc_dec_test.cpp:334-343).
    it, bits = orc.ldpc_decode(1, 8, np.zeros(66 * 8, np.int8), 0, po.CRC_NONE, 2)
    assert it == 0 and (np.unpackbits(bits)[:176] == 1).all()
    # The all-zero codeword (all LLRs positive) has CRC 0 for every CRC: one iteration with early stop.
    for kind in (po.CRC16, po.CRC24A, po.CRC24B):
        it, bits = orc.ldpc_decode(2, 16, np.full(50 * 16, 9, np.int8), 0, kind, 6)
        assert it == 1 and not bits.any()
