"""GPU parity of the downlink twin (LDPC encoding + rate matching, pdc_ldpc_encode / pdc_encode) against the oracle's
orc_ldpc_encode / orc_rate_match / orc_tb_encode, which are pinned against the compiled reference's ldpc_encoder and
ldpc_rate_matcher (tests/test_oracle_vs_reference.py). GF(2) work: bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi, ldpc
from srsran_edgeric_5g_b200.ldpc import LIFTING_SIZES

pytestmark = pytest.mark.gpu


def test_ldpc_encode_every_base_graph_and_lifting_size(ctx, orc):
    rng = np.random.default_rng(71)
    for bg in (1, 2):
        for Z in LIFTING_SIZES:
            K = (22 if bg == 1 else 10) * Z
            msg = rng.integers(0, 2, K).astype(np.uint8)
            got = ctx.ldpc_encode(bg, Z, msg)
            assert (got == orc.ldpc_encode(bg, Z, msg)).all(), (bg, Z)


def test_encode_and_rate_match_random_codeblocks(ctx, orc):
    """Batches of mixed shapes: every redundancy version and modulation order, limited buffers, filler bits, rate-matched
    lengths from a fraction of the codeblock to several laps of the circular buffer."""
    rng = np.random.default_rng(72)
    for trial in range(12):
        n_cb = int(rng.integers(1, 24))
        cbs = np.zeros(n_cb, capi.ENC_DESC_DTYPE)
        msgs, want = [], []
        msg_off = out_off = 0
        for i in range(n_cb):
            bg = int(rng.integers(1, 3))
            Z = int(rng.choice(LIFTING_SIZES))
            kb, N = (22, 66 * Z) if bg == 1 else (10, 50 * Z)
            K = kb * Z
            F = int(rng.integers(0, Z)) if rng.random() < 0.5 else 0
            qm = int(rng.choice([1, 2, 4, 6, 8]))
            rv = int(rng.integers(0, 4))
            nref = 0 if rng.random() < 0.5 else int(rng.integers((kb - 2) * Z + 2 * Z, N + 1))
            E = int(rng.integers(max(qm, N // 5), 3 * N)) // qm * qm
            msg = rng.integers(0, 2, K).astype(np.uint8)
            msg[K - F:] = 0
            cw = orc.ldpc_encode(bg, Z, msg)
            rm = orc.rate_match(cw, E, rv, qm, nref, F)
            # Every other codeblock or so asks for packed output (eight bits per byte, MSB first, zero padded).
            pk = capi.ENC_PACKED if rng.random() < 0.5 else 0
            w = np.packbits(rm) if pk else rm
            want.append((out_off, w))
            cbs[i] = (msg_off, out_off, E, nref, Z, F, bg, qm, rv, pk)
            packed = np.packbits(msg)
            msgs.append(packed)
            msg_off += packed.size
            out_off += w.size + int(rng.integers(0, 5))
        out = ctx.encode(cbs, np.concatenate(msgs), out_capacity=out_off + 8)
        for off, w in want:
            assert (out[off:off + w.size] == w).all(), trial


@pytest.mark.parametrize("bg,qm,nl,tb_bytes,rate", [(1, 8, 4, 159749, 0.9378), (1, 6, 2, 12000, 0.6), (2, 2, 1, 400, 0.3),
                                                   (1, 4, 1, 3000, 0.45)])
def test_transport_block_tx_chain(ctx, orc, bg, qm, nl, tb_bytes, rate):
    """The whole TX chain of a transport block: segmentation and CRC attachment on the host (ldpc.py, mirror of
    ldpc_segmenter_tx), encoding + rate matching on the device, against orc_tb_encode; then the device's own receiver
    decodes what the device encoded (rv 0, no noise)."""
    rng = np.random.default_rng(73 + qm)
    tbs_bits = tb_bytes * 8
    n_llr = int(np.ceil(tbs_bits / rate / qm / nl)) * nl * qm
    C = ldpc.compute_nof_codeblocks(tbs_bits, bg)
    nref = ldpc.compute_N_ref(tb_bytes, C)
    tb = rng.integers(0, 256, tb_bytes).astype(np.uint8)
    for rv in (0, 2, 3, 1):
        want, _ = orc.tb_encode(tb, bg, rv, qm, nref, nl, n_llr)
        metas = ldpc.segment_rx(tbs_bits, bg, rv, qm, nref, nl, n_llr)
        segs = ldpc.segment_tx(ctx, tb, bg)
        assert len(segs) == C == len(metas)
        cbs = np.zeros(C, capi.ENC_DESC_DTYPE)
        msgs, off = [], 0
        for k, (m, seg) in enumerate(zip(metas, segs)):
            packed = np.packbits(seg)
            cbs[k] = (off, m.cw_offset, m.rm_length, nref, m.lifting_size, m.nof_filler_bits, bg, qm, rv, 0)
            msgs.append(packed)
            off += packed.size
        got = ctx.encode(cbs, np.concatenate(msgs), out_capacity=n_llr)
        assert (got == want).all(), rv
    # loopback: rv 0 codeword as ideal soft bits through the device receiver
    want0, _ = orc.tb_encode(tb, bg, 0, qm, nref, nl, n_llr)
    llrs = np.where(want0 == 0, 40, -40).astype(np.int8)
    metas = ldpc.segment_rx(tbs_bits, bg, 0, qm, nref, nl, n_llr)
    crc_kind = capi.CRC24B if C > 1 else (capi.CRC24A if tbs_bits > 3824 else capi.CRC16)
    rx = np.zeros(C, capi.CB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    for k, m in enumerate(metas):
        ctx.harq_write(900 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
        rx[k] = (m.cw_offset, m.rm_length, 900 + k, nref, m.lifting_size, m.nof_filler_bits, bg, qm, 0, crc_kind, 6, flags, 0)
    tbd = np.zeros(1, capi.TB_DESC_DTYPE)
    tbd[0] = (0, C, tbs_bits, 0, 0)
    ctx.submit(rx, llrs, tbd, stream=0)
    out = ctx.wait(0)
    assert out["cb_results"]["crc_ok"].all()
    if C > 1:
        assert out["tb_results"][0]["tb_crc_ok"] == 1
    assert (out["tb_bytes"][:tb_bytes] == tb).all()


def test_invalid_descriptors_are_rejected(ctx):
    cbs = np.zeros(1, capi.ENC_DESC_DTYPE)
    cbs[0] = (0, 0, 100, 0, 17, 0, 1, 2, 0, 0)  # 17 is not a lifting size
    with pytest.raises(capi.PdcError):
        ctx.encode(cbs, np.zeros(64, np.uint8), out_capacity=200)
    cbs[0] = (0, 0, 101, 0, 16, 0, 1, 2, 0, 0)  # E is not a multiple of qm
    with pytest.raises(capi.PdcError):
        ctx.encode(cbs, np.zeros(64, np.uint8), out_capacity=200)
    cbs[0] = (0, 0, 100, 0, 16, 0, 1, 2, 0, 2)  # unknown flag
    with pytest.raises(capi.PdcError):
        ctx.encode(cbs, np.zeros(64, np.uint8), out_capacity=200)
    cbs[0] = (0, 190, 100, 0, 16, 0, 1, 2, 0, capi.ENC_PACKED)  # 13 packed bytes at 190 do not fit in 200
    with pytest.raises(capi.PdcError):
        ctx.encode(cbs, np.zeros(64, np.uint8), out_capacity=200)
    cbs[0] = (0, 187, 100, 0, 16, 0, 1, 2, 0, capi.ENC_PACKED)  # ... at 187 they do
    ctx.encode(cbs, np.zeros(64, np.uint8), out_capacity=200)
