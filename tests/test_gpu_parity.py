"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Integer/byte work: the bar is bit-exact - decoded bits, CRC flag, iteration count and HARQ buffer contents."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi
from tests.vectors import LIFTING_SIZES, awgn_llr, make_cb_batch, random_message

pytestmark = pytest.mark.gpu


def _assert_same(out, ref, what=""):
    assert (out["crc_ok"] == ref["crc_ok"]).all(), f"{what}: crc flags differ"
    assert (out["iters"] == ref["iters"]).all(), f"{what}: iteration counts differ"
    assert (out["bits"] == ref["bits"]).all(), f"{what}: decoded bits differ"
    if "harq" in out:
        assert (out["harq"] == ref["harq"]).all(), f"{what}: HARQ buffers differ"


def test_library_is_native(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] > 100


@pytest.mark.parametrize("early_stop", [True, False])
def test_config1_bg1_z384_rate13(ctx, orc, early_stop):
    # Config 1: BG1, Z=384, full 66Z input, across the waterfall of the 6-iteration decoder (-0.4 dB: a few codeblocks
    # pass, all need 6 iterations; 0 dB: all pass in 5-6; +0.6 dB: 4-5), so that CRC flags AND iteration counts vary.
    passes, total, iters = 0, 0, set()
    for k, snr in enumerate((-0.4, 0.0, 0.6)):
        b = make_cb_batch(orc, 1, 384, n_cb=16, E=66 * 384, qm=2, rv=0, snr_db=snr, seed=11 + k)
        out = b.run_gpu(ctx, 6, early_stop)
        ref = b.run_oracle(orc, 6, early_stop)
        _assert_same(out, ref, f"config1 at {snr} dB")
        assert (out["nlayers"] == 46).all()
        passes += int(ref["crc_ok"].sum())
        total += b.n_cb
        iters |= set(ref["iters"][ref["crc_ok"]].tolist())
    # the vectors really straddle the waterfall (otherwise this test compares failing decodes only)
    assert 0 < passes < total, (passes, total)
    if early_stop:
        assert len(iters) >= 3, iters


@pytest.mark.parametrize("nodes,snr", [(24, 7.4), (26, 6.2), (33, 3.6), (44, 1.2)])
def test_config1_truncated_inputs(ctx, orc, nodes, snr):
    Z = 384
    b = make_cb_batch(orc, 1, Z, n_cb=12, E=nodes * Z, qm=2, rv=0, snr_db=snr, seed=nodes)
    out = b.run_gpu(ctx, 6, True)
    ref = b.run_oracle(orc, 6, True)
    _assert_same(out, ref, f"{nodes} nodes")
    assert (out["nlayers"] == nodes + 2 - 22).all()


def test_config2_bg2_all_small_lifting_sizes(ctx, orc):
    # Config 2: BG2, every lifting size up to 64, all shift tables, QPSK.
    for Z in [z for z in LIFTING_SIZES if z <= 64]:
        b = make_cb_batch(orc, 2, Z, n_cb=6, E=2 * (25 * Z), qm=2, rv=0, snr_db=-2.0, seed=100 + Z, crc_kind=po.CRC16,
                          nof_filler=min(Z, 4))
        out = b.run_gpu(ctx, 6, True)
        ref = b.run_oracle(orc, 6, True)
        _assert_same(out, ref, f"BG2 Z={Z}")


def test_every_base_graph_and_lifting_size(ctx, orc):
    rng = np.random.default_rng(5)
    for bg in (1, 2):
        for Z in LIFTING_SIZES:
            n_short = 66 if bg == 1 else 50
            E = int(rng.integers(30 if bg == 1 else 16, n_short + 10)) * Z
            E -= E % 4
            b = make_cb_batch(orc, bg, Z, n_cb=2, E=E, qm=4, rv=int(rng.integers(0, 4)), snr_db=float(rng.uniform(-2, 5)),
                              seed=1000 * bg + Z, crc_kind=po.CRC24B if Z > 3 else po.CRC16,
                              nof_filler=int(rng.integers(0, Z)))
            mi = int(rng.integers(1, 8))
            es = bool(rng.integers(0, 2))
            out = b.run_gpu(ctx, mi, es)
            ref = b.run_oracle(orc, mi, es)
            _assert_same(out, ref, f"BG{bg} Z={Z} E={E}")


def test_generic_scale_mode(orc):
    from srsran_edgeric_5g_b200 import capi
    c = capi.Context(device=0, max_cbs=64, harq_entries=64, max_tbs=1, max_tb_bytes=4096, scale_mode=capi.SCALE_GENERIC,
                     combine_simd_width=0)
    b = make_cb_batch(orc, 1, 96, n_cb=16, E=40 * 96, qm=2, rv=0, snr_db=2.0, seed=3)
    out = b.run_gpu(c, 6, True)
    ref = b.run_oracle(orc, 6, True, scale=po.SCALE_GENERIC, simd_width=0)
    _assert_same(out, ref, "generic scale")
    c.close()


def test_rate_dematcher_random(ctx, orc):
    rng = np.random.default_rng(4)
    for trial in range(400):
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
        mode = rng.random()
        if mode < 0.5:
            buf0 = rng.integers(-120, 121, N).astype(np.int8)
        elif mode < 0.8:
            buf0 = rng.integers(-128, 128, N).astype(np.int8)
        else:
            buf0 = np.zeros(N, np.int8)
        llr = rng.integers(-120, 121, E).astype(np.int8) if rng.random() < 0.8 else rng.integers(-128, 128, E).astype(
            np.int8)
        new_data = bool(rng.integers(0, 2))
        a, b = buf0.copy(), buf0.copy()
        ctx.rate_dematch(a, llr, new_data, rv, qm, nref, F)
        orc.rate_dematch(b, llr, new_data, rv, qm, nref, F, 64)
        assert (a == b).all(), (bg, Z, qm, F, nref, E, rv, new_data, np.nonzero(a != b)[0][:8])


def test_rate_dematcher_express_combine(ctx, orc):
    """Retransmissions that qualify for the dematcher's express combine (one lap at most, everything a multiple of four
    soft bits): every redundancy version, limited buffers with and without a wrap, filler bits, non-finite soft bits on
    either side (those words take the general path), and a second retransmission on top of the first."""
    rng = np.random.default_rng(44)
    hits = 0
    for trial in range(300):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice([z for z in LIFTING_SIZES if z % 4 == 0]))
        kb = 22 if bg == 1 else 10
        N = (66 if bg == 1 else 50) * Z
        Ksys = (kb - 2) * Z
        qm = int(rng.choice([2, 4, 6, 8]))
        F = 4 * int(rng.integers(0, min(Ksys - 4, 2 * Z) // 4)) if rng.random() < 0.5 else 0
        if rng.random() < 0.5:
            nref = 0
        else:
            nref = int(rng.integers(Ksys + 2 * Z, N))
            nref -= nref % 4 if rng.random() < 0.6 else 0  # a multiple of four lets the walk wrap in the express path
        ncb = min(nref, N) if nref else N
        dn = ncb - F
        e_max = dn // (4 * qm)
        if e_max < 1:
            continue
        E = int(rng.integers(1, e_max + 1)) * 4 * qm
        buf0 = (rng.integers(-120, 121, N) if rng.random() < 0.7 else rng.integers(-128, 128, N)).astype(np.int8)
        a, b = buf0.copy(), buf0.copy()
        for rv in rng.permutation(4)[:2]:
            llr = (rng.integers(-120, 121, E) if rng.random() < 0.8 else rng.integers(-128, 128, E)).astype(np.int8)
            ctx.rate_dematch(a, llr, False, int(rv), qm, nref, F)
            orc.rate_dematch(b, llr, False, int(rv), qm, nref, F, 64)
            assert (a == b).all(), (bg, Z, qm, F, nref, E, int(rv), np.nonzero(a != b)[0][:8])
            hits += 1
    assert hits > 400


def test_express_combine_batch_and_rows_in_use(ctx, orc):
    """A batch large enough for one dematcher CTA per codeblock (where a retransmission takes the entry's previous
    "last non-zero soft bit" instead of reading what it does not touch): sparse soft bits so that the last non-zero
    position moves around - including back down when a retransmission cancels it; whole HARQ entries against the oracle
    and the decoder's number of rows in use (ldpc_decoder_impl.cpp:86-114) against the entry's contents."""
    rng = np.random.default_rng(91)
    n_cb = 1300
    shapes = []
    for i in range(n_cb):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice([8, 12, 16, 20, 24, 28, 32, 36, 40]))
        kb = 22 if bg == 1 else 10
        N = (66 if bg == 1 else 50) * Z
        Ksys = (kb - 2) * Z
        qm = int(rng.choice([2, 4, 6, 8]))
        F = 4 * int(rng.integers(0, Z // 4)) if rng.random() < 0.4 else 0
        nref = 0 if rng.random() < 0.6 else (int(rng.integers(Ksys + 2 * Z, N)) & ~3)
        shapes.append((bg, Z, kb, N, Ksys, qm, F, nref))
    # (the reference leaves parts of a limited buffer stale on a new transmission: start from what the arena holds)
    bufs = [ctx.harq_read(i, s[3]) for i, s in enumerate(shapes)]
    for rnd in range(5):
        cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
        llrs, off = [], 0
        for i, (bg, Z, kb, N, Ksys, qm, F, nref) in enumerate(shapes):
            ncb = nref if nref else N
            e_max = (ncb - F) // (4 * qm)
            new = rnd in (0, 2)  # (round 2: a new transmission over whatever the first one left, stale stretch included)
            E = int(rng.integers(max(1, (Ksys - F) // (4 * qm) + 1), e_max + 1)) * 4 * qm if new else \
                int(rng.integers(1, e_max + 1)) * 4 * qm
            rv = 0 if new else int(rng.integers(0, 4))
            mode = rng.random()
            if mode < 0.4:
                l = rng.integers(-1, 2, E).astype(np.int8)
            elif mode < 0.6:
                l = np.zeros(E, np.int8)
                l[:int(rng.integers(0, E))] = rng.integers(-3, 4, 1)[0]
            else:
                l = rng.integers(-120, 121, E).astype(np.int8) * (rng.random(E) < 0.3)
                l = l.astype(np.int8)
            flags = capi.CB_DEMATCH | capi.CB_DECODE | (capi.CB_NEW_DATA if new else 0)
            cbs[i] = (off, E, i, nref, Z, F, bg, qm, rv, capi.CRC16, 1, flags, 0xffff)
            llrs.append(l)
            off += E
            orc.rate_dematch(bufs[i], l, new, rv, qm, nref, F, 64)
        ctx.submit(cbs, np.concatenate(llrs), None, stream=0, want_bits=False)
        out = ctx.wait(0)
        for i, (bg, Z, kb, N, Ksys, qm, F, nref) in enumerate(shapes):
            got = ctx.harq_read(i, N)
            assert (got == bufs[i]).all(), (rnd, i, shapes[i], np.nonzero(got != bufs[i])[0][:8])
            nz = np.nonzero(bufs[i])[0]
            res = out["cb_results"][i]
            if nz.size == 0:
                assert res["status"] == 1
                continue
            cb_len = max(int(nz[-1]) + 1 + 2 * Z, (kb + 4) * Z)
            cb_len = (cb_len + Z - 1) // Z * Z
            assert res["nlayers"] == cb_len // Z - kb, (rnd, i, shapes[i], int(nz[-1]), res)


def test_harq_sequences_random(ctx, orc):
    """Randomised HARQ histories through pdc_submit (tests/vectors.py: harq_sequence_rounds): few entries (several
    dematcher CTAs per codeblock) and more than a thousand (one CTA per codeblock, where the express paths rely on the
    entry's record of its last non-zero soft bit), the second large population taking over the entries of the first with
    other codeblock shapes - what the reference's buffer pool does when it hands codeblock buffers out again
    (rx_buffer_pool_impl.cpp:44: nothing is cleared)."""
    from tests.vectors import harq_sequence_rounds
    n = harq_sequence_rounds(ctx, orc, np.random.default_rng(2024), n_ent=40, rounds=6, harq_base=0, max_Z=384)
    n += harq_sequence_rounds(ctx, orc, np.random.default_rng(2025), n_ent=1250, rounds=4, harq_base=0, max_Z=48)
    n += harq_sequence_rounds(ctx, orc, np.random.default_rng(2026), n_ent=1250, rounds=4, harq_base=0, max_Z=48)
    assert n > 8000


def test_entry_taken_over_by_other_shapes(ctx, orc):
    """An entry's record of its last non-zero soft bit is only good for the codeblock length it was written for: a long
    codeblock fills the entries, a short one takes them over (its record describes 800 positions), then a codeblock with a
    limited buffer is received whose stale stretch [1568, 1960) still holds what the FIRST one left there - the decoder
    must see those soft bits (rows in use, ldpc_decoder_impl.cpp:86-114) like the reference, which scans the buffer."""
    rng = np.random.default_rng(77)
    n_cb = 1250  # one dematcher CTA per codeblock: the express paths consult the record
    steps = ((1, 32, 0, 2112, 2), (2, 16, 500, 480, 4), (1, 32, 1960, 1568, 8))
    bufs = [ctx.harq_read(i, 2112) for i in range(n_cb)]
    for step, (bg, Z, nref, E, qm) in enumerate(steps):
        N, kb = (66 if bg == 1 else 50) * Z, (22 if bg == 1 else 10)
        cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
        llrs = rng.integers(-120, 121, (n_cb, E)).astype(np.int8)
        llrs[llrs == 0] = 1
        want = []
        for i in range(n_cb):
            cbs[i] = (i * E, E, i, nref, Z, 0, bg, qm, 0, capi.CRC24B, 2, capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA,
                      0xffff)
            want.append(orc.cb_decode(bufs[i][:N], llrs[i], True, 0, qm, nref, 0, po.CRC24B, False, 2))
        ctx.submit(cbs, np.ascontiguousarray(llrs.reshape(-1)), None, stream=0, want_bits=True)
        out = ctx.wait(0)
        for i in range(n_cb):
            assert (ctx.harq_read(i, 2112) == bufs[i]).all(), (step, i)
            last = int(np.nonzero(bufs[i][:N])[0][-1])
            cb_len = (max(last + 1 + 2 * Z, (kb + 4) * Z) + Z - 1) // Z * Z
            assert out["cb_results"]["nlayers"][i] == cb_len // Z - kb, (step, i, last, out["cb_results"][i])
            assert (out["cb_bits"][i, :(kb * Z + 7) // 8] == want[i][1]).all(), (step, i)
        if step == 2:
            assert last >= 1568  # the stale stretch did hold soft bits of the first codeblock


def test_high_rate_hint_never_changes_results(ctx, orc):
    """PDC_LAUNCH_HIGH_RATE only selects a decoder instantiation: a batch that does NOT fit the hint (all 46 rows in use,
    and a mix of row counts) and one that does (four rows) decode identically with and without it, and like the oracle."""
    import torch
    stream = torch.cuda.current_stream()
    for E, qm, snr, seed in ((66 * 384, 2, 0.5, 21), (30 * 384, 4, 4.0, 22), (8960, 8, 8.4, 23)):
        b = make_cb_batch(orc, 1, 384, n_cb=10, E=E, qm=qm, rv=0, snr_db=snr, seed=seed)
        ref = b.run_oracle(orc, 6, True)
        cbs = b.descriptors(capi, 6, True, harq_base=100)
        for i in range(b.n_cb):
            ctx.harq_write(100 + i, np.zeros(b.N, np.int8))
        d_cbs = torch.from_numpy(cbs.view(np.uint8)).cuda()
        d_llr = torch.from_numpy(np.ascontiguousarray(b.llrs.reshape(-1))).cuda()
        flags = int(np.bitwise_or.reduce(cbs["flags"]))
        outs = []
        for hint in (0, capi.LAUNCH_HIGH_RATE):
            d_res = torch.zeros(b.n_cb * 4, dtype=torch.uint8, device="cuda")
            d_bits = torch.zeros(b.n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
            ctx.launch_device(d_cbs.data_ptr(), b.n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384,
                              flags | hint, True, cuda_stream=stream.cuda_stream)
            torch.cuda.synchronize()
            res = d_res.cpu().numpy().view(capi.CB_RESULT_DTYPE)
            bits = d_bits.cpu().numpy().reshape(b.n_cb, capi.PDC_MAX_CB_BYTES)[:, :(b.K + 7) // 8]
            outs.append((res.copy(), bits.copy()))
            assert (res["crc_ok"].astype(bool) == ref["crc_ok"]).all() and (bits == ref["bits"]).all(), (E, hint)
            assert (np.where(res["crc_ok"] == 1, res["iters"], 6) == ref["iters"]).all(), (E, hint)
        assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()


def test_crc(ctx, orc):
    rng = np.random.default_rng(6)
    for kind in (po.CRC16, po.CRC24A, po.CRC24B, po.CRC24C, po.CRC11, po.CRC6):
        for n in (1, 7, 8, 24, 31, 32, 33, 100, 1000, 8448, 30000, 1277992):
            d = rng.integers(0, 256, (n + 7) // 8 + 1).astype(np.uint8)
            assert ctx.crc(kind, d, n) == orc.crc(kind, d, n), (kind, n)


def test_ldpc_decoder_single_cb_api(ctx, orc):
    rng = np.random.default_rng(8)
    # all-zero input (ldpc_enc_dec_test.cpp:334-343) and a clean +-10 codeword decoded in one iteration without CRC.
    it, out = ctx.ldpc_decode(1, 8, np.zeros(66 * 8, np.int8), 0, po.CRC_NONE, 2)
    assert it == 0 and (np.unpackbits(out)[:22 * 8] == 1).all()
    for bg, Z in ((1, 384), (2, 7), (2, 64), (1, 13)):
        K = (22 if bg == 1 else 10) * Z
        msg = rng.integers(0, 2, K).astype(np.uint8)
        cw = orc.ldpc_encode(bg, Z, msg)
        n = int(rng.integers(K + 2 * Z, cw.size + 1))  # any length, also not a multiple of Z
        llr = (10 - 20 * cw.astype(np.int16)).astype(np.int8)[:n]
        it, out = ctx.ldpc_decode(bg, Z, llr, 0, po.CRC_NONE, 1)
        assert it == 0 and (np.unpackbits(out)[:K] == msg).all(), (bg, Z, n)
        it_o, out_o = orc.ldpc_decode(bg, Z, llr, 0, po.CRC_NONE, 1)
        assert (out == out_o).all()


def test_harq_retransmissions_cb_level(ctx, orc):
    # rv sequence {0,2,3,1} combined in one HARQ entry; noise such that rv0 alone fails for most codeblocks.
    Z, bg, qm = 384, 1, 8
    rng = np.random.default_rng(21)
    n_cb, E = 8, 9216
    msgs = [random_message(orc, bg, Z, 16, po.CRC24B, rng) for _ in range(n_cb)]
    cws = [orc.ldpc_encode(bg, Z, m) for m in msgs]
    harq_o = np.full((n_cb, 66 * Z), 55, np.int8)  # sentinel: stale regions must match too
    for i in range(n_cb):
        ctx.harq_write(300 + i, harq_o[i])
    from tests.vectors import CbBatch
    for t, rv in enumerate([0, 2, 3, 1]):
        llrs = np.stack([awgn_llr(orc.rate_match(cw, E, rv, qm, 12611, 16), 5.3, rng) for cw in cws])
        b = CbBatch(bg, Z, E, qm, rv, 12611, 16, po.CRC24B, llrs, np.stack(msgs))
        out = b.run_gpu(ctx, 6, True, new_data=(t == 0), harq_base=300, harq_init=None)
        ref = b.run_oracle(orc, 6, True, new_data=(t == 0), harq=harq_o)
        _assert_same(out, ref, f"retx {t}")


def test_mixed_batch_pairs_and_lone_codeblocks(ctx, orc):
    # One batch mixing lifting sizes, base graphs, iteration counts and an odd number of codeblocks: consecutive
    # codeblocks that cannot share a CTA pair are decoded one after the other; results must not depend on the pairing.
    from srsran_edgeric_5g_b200 import capi
    rng = np.random.default_rng(31)
    shapes = [(1, 384, 30 * 384, 6), (1, 384, 30 * 384, 6), (1, 384, 30 * 384, 4), (2, 384, 20 * 384, 6),
              (1, 96, 40 * 96, 6), (1, 96, 26 * 96, 6), (2, 15, 30 * 15 + 2, 5), (1, 384, 66 * 384, 6),
              (1, 384, 66 * 384, 6), (2, 7, 30 * 7 + 4, 6), (1, 208, 33 * 208, 6)]
    batches = [make_cb_batch(orc, bg, Z, 1, E - E % 2, 2, 0, float(rng.uniform(1, 6)), 500 + i, nof_filler=i % 3)
               for i, (bg, Z, E, mi) in enumerate(shapes)]
    n = len(batches)
    cbs = np.zeros(n, capi.CB_DESC_DTYPE)
    off = 0
    llrs = []
    for i, (b, (bg, Z, E, mi)) in enumerate(zip(batches, shapes)):
        flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | (capi.CB_EARLY_STOP if i % 2 == 0 else 0)
        cbs[i] = (off, b.E, 700 + i, 0, Z, b.nof_filler, bg, 2, 0, b.crc_kind, mi, flags, 0xffff)
        ctx.harq_write(700 + i, np.zeros(b.N, np.int8))
        llrs.append(b.llrs[0])
        off += b.E
    ctx.submit(cbs, np.concatenate(llrs), None, stream=0, want_bits=True)
    out = ctx.wait(0)
    for i, (b, (bg, Z, E, mi)) in enumerate(zip(batches, shapes)):
        ref = b.run_oracle(orc, mi, i % 2 == 0)
        r = out["cb_results"][i]
        assert bool(r["crc_ok"]) == bool(ref["crc_ok"][0]) and r["iters"] == ref["iters"][0], (i, r, ref["iters"])
        assert (out["cb_bits"][i, :(b.K + 7) // 8] == ref["bits"][0]).all(), i
        assert (ctx.harq_read(700 + i, b.N) == ref["harq"][0]).all(), i


def test_pair_with_different_rows_in_use(ctx, orc):
    # Same shape, but the second codeblock ends in zero LLRs that span a whole node: the reference trims them and uses
    # fewer base-graph rows (ldpc_decoder_impl.cpp:86-114), so the pair cannot be decoded in lock step.
    b = make_cb_batch(orc, 1, 64, n_cb=4, E=40 * 64, qm=2, rv=0, snr_db=2.5, seed=77)
    b.llrs[1, -130:] = 0
    b.llrs[2, -1:] = 0
    out = b.run_gpu(ctx, 6, True, harq_base=800)
    ref = b.run_oracle(orc, 6, True)
    _assert_same(out, ref, "unequal rows")
    assert out["nlayers"][0] != out["nlayers"][1]


def test_all_zero_and_dematch_only(ctx, orc):
    from srsran_edgeric_5g_b200 import capi
    b = make_cb_batch(orc, 2, 32, n_cb=3, E=20 * 32, qm=2, rv=0, snr_db=3.0, seed=5, crc_kind=po.CRC16)
    b.llrs[1] = 0  # all-zero input: not decodable
    cbs = b.descriptors(capi, 6, True, True, harq_base=900)
    cbs["flags"][2] = int(cbs["flags"][2]) & (0xff ^ capi.CB_DECODE)  # dematch / combine only
    for i in range(3):
        ctx.harq_write(900 + i, np.zeros(b.N, np.int8))
    ctx.submit(cbs, np.ascontiguousarray(b.llrs.reshape(-1)), None, stream=0, want_bits=True)
    out = ctx.wait(0)
    ref = b.run_oracle(orc, 6, True)
    r = out["cb_results"]
    assert r["status"][1] == 1 and r["crc_ok"][1] == 0 and r["iters"][1] == 6
    assert bool(r["crc_ok"][0]) == bool(ref["crc_ok"][0]) and r["iters"][0] == ref["iters"][0]
    assert r["iters"][2] == 0 and r["crc_ok"][2] == 0  # not decoded
    assert (ctx.harq_read(902, b.N) == ref["harq"][2]).all()


def test_general_kernel_forced(orc, monkeypatch):
    # PDC_FORCE_SCALAR=1 routes every batch through the general kernel: both kernels must agree with the oracle.
    from srsran_edgeric_5g_b200 import capi
    monkeypatch.setenv("PDC_FORCE_SCALAR", "1")
    c = capi.Context(device=0, max_cbs=64, harq_entries=64, max_tbs=1, max_tb_bytes=4096)
    monkeypatch.delenv("PDC_FORCE_SCALAR")
    for bg, Z, E in ((1, 384, 30 * 384), (2, 36, 40 * 36), (1, 11, 50 * 11 + 1)):
        b = make_cb_batch(orc, bg, Z, n_cb=5, E=E - E % 2, qm=2, rv=0, snr_db=3.0, seed=Z)
        _assert_same(b.run_gpu(c, 6, True), b.run_oracle(orc, 6, True), f"general kernel BG{bg} Z={Z}")
    c.close()


def test_persistent_ctas_across_changing_shapes(ctx, orc):
    """More codeblock pairs than resident CTAs, shapes changing from pair to pair: a persistent CTA keeps the lifted
    graph and the CRC weights of its previous pair in shared memory when the shape repeats and must rebuild exactly
    what changed (base graph / lifting size; filler bits or CRC polynomial only; iteration count). One Z = 384 codeblock
    puts the batch on the 384-thread kernel (296 CTAs); 720 codeblocks give every CTA at least one more pair."""
    from srsran_edgeric_5g_b200 import capi
    rng = np.random.default_rng(4242)
    kinds = [(1, 384, po.CRC24B, 0)]
    for _ in range(40):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice([4, 7, 10, 16, 24, 36, 52, 64]))
        kinds.append((bg, Z, [po.CRC16, po.CRC24A, po.CRC24B][int(rng.integers(0, 3))] if Z > 3 else po.CRC16,
                      int(rng.integers(0, Z))))
    protos = []
    for i, (bg, Z, crc, F) in enumerate(kinds):
        n_short = 66 if bg == 1 else 50
        E = int(rng.integers(28 if bg == 1 else 14, n_short + 1)) * Z
        E -= E % 2
        protos.append(make_cb_batch(orc, bg, Z, 1, E, 2, 0, float(rng.uniform(0.5, 6)), 9000 + i, crc_kind=crc, nof_filler=F))
    n = 720
    pick = np.concatenate([[0], rng.integers(1, len(protos), n - 1)])
    # pairs repeat their shape now and then (reuse path) and otherwise change it (rebuild path)
    for i in range(2, n, 2):
        if rng.random() < 0.3:
            pick[i], pick[i + 1] = pick[i - 2], pick[i - 1]
        elif rng.random() < 0.5:
            pick[i + 1] = pick[i]
    cbs = np.zeros(n, capi.CB_DESC_DTYPE)
    llrs, off = [], 0
    iters = rng.integers(1, 7, n)
    for i, p in enumerate(pick):
        b = protos[p]
        flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | (capi.CB_EARLY_STOP if i % 3 else 0)
        cbs[i] = (off, b.E, 1000 + i, 0, b.Z, b.nof_filler, b.bg, 2, 0, b.crc_kind, int(iters[i]), flags, 0xffff)
        llrs.append(b.llrs[0])
        off += b.E
    for i in range(n):
        ctx.harq_write(1000 + i, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
    ctx.submit(cbs, np.concatenate(llrs), None, stream=0, want_bits=True)
    out = ctx.wait(0)
    cache = {}
    for i, p in enumerate(pick):
        key = (int(p), int(iters[i]), bool(i % 3))
        if key not in cache:
            cache[key] = protos[p].run_oracle(orc, int(iters[i]), bool(i % 3))
        ref, b, r = cache[key], protos[p], out["cb_results"][i]
        assert bool(r["crc_ok"]) == bool(ref["crc_ok"][0]) and r["iters"] == ref["iters"][0], (i, kinds[p], r, ref["iters"])
        assert (out["cb_bits"][i, :(b.K + 7) // 8] == ref["bits"][0]).all(), (i, kinds[p])
