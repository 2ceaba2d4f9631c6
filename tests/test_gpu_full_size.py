"""GPU tests at the full sizes of BASELINE.json (the batch the headline metric is measured on, and the 16-cell slot of
config 4), through properties that do not need the oracle to decode thousands of codeblocks: encode -> noise -> decode
round trips, agreement of every copy of a tiled input (determinism over CTAs, codeblock pairs and the work counter),
idempotence of a decode-only pass, and HARQ combining that rescues a failed first transmission. A sample of the batch is
still compared with the oracle bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi, ldpc
from tests.vectors import CbBatch, make_cb_batch, make_tb_llrs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_ctx():
    c = capi.Context(device=0, max_cbs=8192, max_llrs=8192 * 25344, harq_entries=8192, max_tbs=16,
                     max_tb_bytes=16 * 160000, nof_streams=1)
    yield c
    c.close()


@pytest.mark.parametrize("early_stop", [True, False])
def test_headline_batch_8192_codeblocks(big_ctx, orc, early_stop):
    """8192 codeblocks BG1 Z = 384 rate 1/3 (46 rows), 6 iterations: the workload bench.py quotes."""
    distinct, copies = 64, 128
    small = make_cb_batch(orc, 1, 384, distinct, 25344, 2, 0, 0.5, seed=808)
    llrs = np.tile(small.llrs, (copies, 1))
    batch = CbBatch(1, 384, 25344, 2, 0, 0, 0, po.CRC24B, llrs, np.tile(small.msgs, (copies, 1)))
    cbs = batch.descriptors(capi, 6, early_stop)
    big_ctx.submit(cbs, np.ascontiguousarray(llrs.reshape(-1)), None, stream=0, want_bits=True)
    out = big_ctx.wait(0)
    res, bits = out["cb_results"], out["cb_bits"][:, :1056]
    # (a) round trip: every codeblock decodes to the message that was encoded
    assert res["crc_ok"].all() and (res["status"] == 0).all()
    want = np.packbits(batch.msgs, axis=1)
    assert (bits == want).all()
    # (b) every copy of an input gives the same iteration count, rows in use and bits
    for f in ("iters", "nlayers", "crc_ok"):
        assert (res[f].reshape(copies, distinct) == res[f][:distinct]).all(), f
    assert (res["nlayers"] == 46).all()
    # (c) the distinct codeblocks against the oracle: iteration counts and bits
    ref = small.run_oracle(orc, 6, early_stop)
    assert (res["iters"][:distinct] == ref["iters"]).all() and (bits[:distinct] == ref["bits"]).all()
    # (d) idempotence: decoding the HARQ entries again without dematching changes nothing
    cbs2 = cbs.copy()
    cbs2["flags"] &= ~np.uint8(capi.CB_DEMATCH)
    big_ctx.submit(cbs2, np.zeros(16, np.int8), None, stream=0, want_bits=True)  # no LLRs are read
    again = big_ctx.wait(0)
    assert (again["cb_results"] == res).all() and (again["cb_bits"][:, :1056] == bits).all()
    # (e) the HARQ soft buffers of a sample are the oracle's
    for i in (0, 63, 64, 4097, 8191):
        assert (big_ctx.harq_read(i, 25344) == ref["harq"][i % distinct]).all(), i


def test_config4_slot_16_cells_with_harq_rescue(big_ctx, orc):
    """16 x (273 PRB, 256QAM, 4 layers) = 2432 codeblocks in one batch with device TB assembly. Half of the cells get a
    first transmission too noisy to decode; their retransmission (rv 2) is combined in the device HARQ arena and must
    rescue them, while the good cells deliver their transport blocks at once."""
    rng = np.random.default_rng(4040)
    tbs_bits, n_llr, qm, nl, cells = 1277992, 1362816, 8, 4, 16
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tbs = [rng.integers(0, 256, tbs_bits // 8).astype(np.uint8) for _ in range(2)]  # two distinct TBs, alternating
    tb_stride = (tbs_bits + 24 + 31) // 32 * 4

    def slot(rv, new_data, snrs, active):
        cbs, tbd, llrs = [], [], []
        for k, c in enumerate(active):
            l, _ = make_tb_llrs(orc, tbs[c % 2], 1, rv, qm, nref, nl, n_llr, snrs[c], rng)
            llrs.append(l)
            flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_EARLY_STOP | (capi.CB_NEW_DATA if new_data else 0)
            for i, m in enumerate(ldpc.segment_rx(tbs_bits, 1, rv, qm, nref, nl, n_llr)):
                cbs.append((k * n_llr + m.cw_offset, m.rm_length, c * C + i, nref, m.lifting_size, m.nof_filler_bits, 1, qm,
                            rv, capi.CRC24B, 6, flags, k))
            tbd.append((k * C, C, tbs_bits, k * tb_stride, 0))
        big_ctx.submit(np.array(cbs, capi.CB_DESC_DTYPE), np.concatenate(llrs), np.array(tbd, capi.TB_DESC_DTYPE), stream=0,
                       want_bits=False)
        return big_ctx.wait(0)

    for i in range(cells * C):
        big_ctx.harq_write(i, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
    snrs = [9.0 if c % 2 == 0 else 5.5 for c in range(cells)]
    out = slot(0, True, snrs, list(range(cells)))
    ok = out["tb_results"]["tb_crc_ok"].astype(bool)
    assert ok[0::2].all() and not ok[1::2].any()
    for c in range(0, cells, 2):
        o = c * tb_stride
        assert (out["tb_bytes"][o:o + tbs_bits // 8] == tbs[c % 2]).all()
    # retransmission of the failed cells only, combined with what the arena holds
    bad = list(range(1, cells, 2))
    out2 = slot(2, False, snrs, bad)
    assert out2["tb_results"]["tb_crc_ok"].all()
    for k, c in enumerate(bad):
        o = k * tb_stride
        assert (out2["tb_bytes"][o:o + tbs_bits // 8] == tbs[c % 2]).all()
