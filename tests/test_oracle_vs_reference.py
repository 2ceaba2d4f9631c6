"""CPU tests: the oracle against the compiled reference (oracle/_ref/libsrsref.so) on fresh random inputs.
Skipped where the library has not been built (it is built by __graft_entry__.build() wherever /root/reference exists
and shipped with the snapshot)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.vectors import LIFTING_SIZES, awgn_llr, random_message

pytestmark = pytest.mark.skipif(not po.Reference.available(), reason="oracle/_ref/libsrsref.so not built")


def test_crc(orc):
    ref = po.Reference("auto")
    rng = np.random.default_rng(1)
    for kind in (po.CRC16, po.CRC24A, po.CRC24B, po.CRC24C, po.CRC11, po.CRC6):
        for n in (1, 7, 8, 24, 25, 100, 1000, 8448, 30000):
            d = rng.integers(0, 256, (n + 7) // 8 + 1).astype(np.uint8)
            assert orc.crc(kind, d, n) == ref.crc(kind, d, n)


def test_encoder_all_graphs(orc):
    ref = po.Reference("auto")
    rng = np.random.default_rng(2)
    for bg in (1, 2):
        for Z in LIFTING_SIZES:
            msg = rng.integers(0, 2, (22 if bg == 1 else 10) * Z).astype(np.uint8)
            assert (orc.ldpc_encode(bg, Z, msg) == ref.ldpc_encode(bg, Z, msg)).all(), (bg, Z)


@pytest.mark.parametrize("variant,scale", [("auto", po.SCALE_X86), ("avx2", po.SCALE_X86), ("generic", po.SCALE_GENERIC)])
def test_decoder(orc, variant, scale):
    try:
        ref = po.Reference(variant)
    except RuntimeError:
        pytest.skip(f"{variant} not supported by this CPU")
    rng = np.random.default_rng(3)
    for trial in range(150):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice(LIFTING_SIZES))
        kb = 22 if bg == 1 else 10
        K = kb * Z
        crc_kind = int(rng.integers(0, 4))
        crc_bits = 16 if crc_kind == 1 else 24
        F = int(rng.integers(0, max(1, min(Z, K - crc_bits - 1)))) if rng.random() < 0.5 else 0
        if K - F - crc_bits <= 0:
            continue
        cw = orc.ldpc_encode(bg, Z, random_message(orc, bg, Z, F, crc_kind, rng))
        nodes = int(rng.integers(kb + 2, (66 if bg == 1 else 50) + 1))
        rate = K / (nodes * Z)
        llr = awgn_llr(cw, 10 * np.log10(2 ** (2 * rate) - 1) + 1.5 + rng.uniform(-1.5, 1.5), rng)
        llr[K - 2 * Z - F:K - 2 * Z] = 127
        llr[nodes * Z:] = 0
        if rng.random() < 0.1:
            llr[rng.integers(0, llr.size, 5)] = rng.choice([-128, -127, 127, 121, -121], 5)
        mi = int(rng.integers(1, 9))
        a = ref.ldpc_decode(bg, Z, llr, F, crc_kind, mi)
        b = orc.ldpc_decode(bg, Z, llr, F, crc_kind, mi, scale)
        assert a[0] == b[0] and (a[1] == b[1]).all(), (variant, bg, Z, F, crc_kind, mi)


@pytest.mark.parametrize("variant,width", [("avx512", 64), ("avx2", 32), ("generic", 0)])
def test_dematcher(orc, variant, width):
    try:
        ref = po.Reference(variant)
    except RuntimeError:
        pytest.skip(f"{variant} not supported by this CPU")
    rng = np.random.default_rng(4)
    for trial in range(500):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice(LIFTING_SIZES))
        kb = 22 if bg == 1 else 10
        N = (66 if bg == 1 else 50) * Z
        Ksys = (kb - 2) * Z
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        F = int(rng.integers(0, min(Ksys - 1, 2 * Z))) if rng.random() < 0.6 else 0
        nref = int(rng.integers(Ksys + 2 * Z, N + 50)) if rng.random() < 0.4 else 0
        E = int(rng.integers(1, max(2, 3 * N // qm))) * qm
        rv = int(rng.integers(0, 4))
        buf0 = (rng.integers(-120, 121, N) if rng.random() < 0.6 else rng.integers(-128, 128, N)).astype(np.int8)
        llr = (rng.integers(-120, 121, E) if rng.random() < 0.8 else rng.integers(-128, 128, E)).astype(np.int8)
        new_data = bool(rng.integers(0, 2))
        a, b = buf0.copy(), buf0.copy()
        ref.rate_dematch(a, llr, new_data, rv, qm, nref, F)
        orc.rate_dematch(b, llr, new_data, rv, qm, nref, F, width)
        assert (a == b).all(), (variant, bg, Z, qm, F, nref, E, rv, new_data)


def test_segmentation_and_tb_encode(orc):
    ref = po.Reference("auto")
    rng = np.random.default_rng(5)
    for trial in range(150):
        bg = int(rng.integers(1, 3))
        tb_bytes = int(rng.integers(1, 3000 if bg == 2 else 20000)) if rng.random() < 0.9 else int(
            rng.integers(100000, 159749))
        if bg == 2 and tb_bytes * 8 > 30000:
            tb_bytes = 3000
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.integers(1, 5))
        nsym = int(np.ceil(tb_bytes * 8 / rng.uniform(0.15, 0.92) / qm / nl)) * nl
        n_llr = nsym * qm
        rv = int(rng.integers(0, 4))
        nref = 0 if rng.random() < 0.5 else 25344
        mo = orc.segment_rx(tb_bytes * 8, bg, qm, nl, n_llr)
        a = np.array([[m.Z, m.full_length, m.rm_length, m.nof_filler, m.cw_offset, m.nof_crc_bits] for m in mo])
        assert (a == ref.segment_rx(tb_bytes * 8, bg, rv, qm, nref, nl, n_llr)).all()
        if tb_bytes < 4000:
            tb = rng.integers(0, 256, tb_bytes).astype(np.uint8)
            co, no = orc.tb_encode(tb, bg, rv, qm, nref, nl, n_llr)
            cr, nr = ref.tb_encode(tb, bg, rv, qm, nref, nl, n_llr)
            assert no == nr and (co == cr).all()


def test_pusch_decoder_harq(orc):
    rng = np.random.default_rng(6)
    for trial in range(12):
        bg = int(rng.integers(1, 3))
        tb_bytes = int(rng.integers(20, 900 if bg == 2 else 6000))
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.integers(1, 3))
        rate = rng.uniform(0.5, 0.9) if bg == 1 else rng.uniform(0.2, 0.6)
        nsym = int(np.ceil(tb_bytes * 8 / rate / qm / nl)) * nl
        n_llr = nsym * qm
        nref = 0 if rng.random() < 0.5 else int(rng.integers(8000, 25344))
        tb = rng.integers(0, 256, tb_bytes).astype(np.uint8)
        metas = orc.segment_rx(tb_bytes * 8, bg, qm, nl, n_llr)
        C = len(metas)
        fill = int(rng.integers(-120, 121))
        H = po.Harq(C, fill)
        RP = po.ReferencePusch(C, "auto", fill)
        es, mi = bool(rng.integers(0, 2)), int(rng.integers(2, 7))
        snr = (8 if bg == 1 else 3) * rate / 0.8 + rng.uniform(-4, -1)
        for t, rv in enumerate([0, 2, 3, 1]):
            cw, _ = orc.tb_encode(tb, bg, rv, qm, nref, nl, n_llr)
            llr = awgn_llr(cw, snr, rng)
            tbo, so = orc.pusch_decode(H, llr, tb_bytes, bg, rv, qm, nref, nl, mi, es, t == 0)
            tbr, sr = RP.decode(llr, tb_bytes, bg, rv, qm, nref, nl, mi, es, t == 0)
            assert (so[:5] == sr[:5]).all()
            if so[0]:
                assert (tbo == tbr).all() and (tbo == tb).all()
            for cb in range(C):
                s, c = RP.get_cb(cb, metas[cb].full_length)
                assert (s == H.soft[cb][:metas[cb].full_length]).all() and c == bool(H.crc_ok[cb])
            if so[0]:
                break


def test_codeblock_decoder_on_reused_buffers(orc):
    """pusch_codeblock_decoder (rate dematcher + decoder on the buffer it keeps, pusch_codeblock_decoder.cpp:35-71) on
    buffers that codeblocks of OTHER sizes used before and nobody cleared (rx_buffer_pool_impl.cpp:44): what an older,
    longer codeblock left behind shows in the stale stretch of a limited-buffer transmission and decides how many rows
    the decoder uses (ldpc_decoder_impl.cpp:86-114). The oracle against the compiled reference, buffers and results -
    the GPU is held to the same histories in tests/test_gpu_parity.py (test_harq_sequences_random,
    test_entry_taken_over_by_other_shapes)."""
    from tests.vectors import awgn_llr, random_message
    ref = po.Reference("auto")
    rng = np.random.default_rng(8)
    sizes = [z for z in LIFTING_SIZES if z <= 96]
    for entry in range(12):
        slot_r = np.zeros(po.MAX_CB_SIZE, np.int8)  # one pool entry, as the reference and as the oracle see it
        slot_o = np.zeros(po.MAX_CB_SIZE, np.int8)
        for owner in range(6):
            bg = int(rng.integers(1, 3))
            Z = int(rng.choice(sizes))
            kb = 22 if bg == 1 else 10
            N, K, Ksys = (66 if bg == 1 else 50) * Z, kb * Z, (kb - 2) * Z
            qm = int(rng.choice([2, 4, 6, 8]))
            crc = po.CRC24B if K > 64 else po.CRC16
            F = 4 * int(rng.integers(0, max(1, min(Z // 2, (K - 26) // 4)))) if rng.random() < 0.5 else 0
            nref = int(rng.integers(Ksys + 2 * Z, N + 1)) if rng.random() < 0.6 else 0
            cw = orc.ldpc_encode(bg, Z, random_message(orc, bg, Z, F, crc, rng))
            for t in range(int(rng.integers(1, 4))):
                ncb = min(nref, N) if nref else N
                E = int(rng.integers(max(1, (Ksys - F) // qm), max(2, (ncb - F) // qm + 1))) * qm
                rv = 0 if t == 0 else int(rng.integers(0, 4))
                llr = awgn_llr(orc.rate_match(cw, E, rv, qm, nref, F), float(rng.uniform(-2.0, 8.0)), rng)
                es, mi = bool(rng.integers(0, 2)), int(rng.integers(1, 7))
                a = ref.cb_decode(slot_r[:N], llr, t == 0, rv, qm, nref, F, crc, es, mi)
                b = orc.cb_decode(slot_o[:N], llr, t == 0, rv, qm, nref, F, crc, es, mi)
                case = (entry, owner, t, bg, Z, qm, F, nref, E, rv, es, mi)
                assert (slot_r == slot_o).all(), case
                assert a[0] == b[0] and (a[1] == b[1]).all(), case
