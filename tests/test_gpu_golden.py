"""GPU parity against the committed golden fixtures (-m gpu), straight through the C ABI - no oracle in between.

tests/golden/*.npz hold (a) the reference tree's own on-disk LDPC vectors (examplesBG{1,2}.dat) and (b) outputs of the
compiled reference (oracle/_ref/libsrsref.so), made by tests/golden/make_golden.py. The CPU suite pins the oracle to
them (tests/test_oracle_golden.py); here the CUDA path is pinned to them directly, on both decoder kernels:
  * the single-codeblock call (pdc_ldpc_decode: general kernel, any input length),
  * the batched call (pdc_submit with decode-only descriptors over HARQ entries: throughput kernel) for inputs that are a
    whole number of variable nodes - what pusch_decoder feeds it."""
import zlib
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi
from tests.vectors import LIFTING_SIZES

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _decode_batch(ctx, cases):
    """cases: list of (bg, Z, llr, F, crc_kind, max_iter, early_stop). One decode-only batch over HARQ entries 0..n-1
    (written with the LLRs, zero beyond them): what the throughput kernel sees after a rate dematcher."""
    cbs = np.zeros(len(cases), capi.CB_DESC_DTYPE)
    for i, (bg, Z, llr, F, kind, mi, es) in enumerate(cases):
        buf = np.zeros(capi.PDC_MAX_CB_SOFT, np.int8)
        buf[:llr.size] = llr
        ctx.harq_write(i, buf)
        flags = capi.CB_DECODE | (capi.CB_EARLY_STOP if es else 0)
        cbs[i] = (0, 0, i, 0, Z, F, bg, 2, 0, kind, mi, flags, 0xffff)
    ctx.submit(cbs, np.zeros(16, np.int8), None, stream=0, want_bits=True)
    return ctx.wait(0)


def test_examples_encode_and_decode(ctx):
    # ldpc_enc_dec_test.cpp:237-315 on the reference's own message / codeword pairs: noiseless LLR = +-10, fillers
    # +127 (what the dematcher writes), ONE iteration, no CRC: decoded bits == message for every base graph and
    # lifting size; the encoder reproduces the codeword.
    ex = np.load(G / "ldpc_examples.npz")
    for bg in (1, 2):
        batch, want = [], []
        for Z in LIFTING_SIZES:
            K = (22 if bg == 1 else 10) * Z
            N = (66 if bg == 1 else 50) * Z
            msgs = np.unpackbits(ex[f"bg{bg}_z{Z}_msg"], axis=1)[:, :K]
            cws = np.unpackbits(ex[f"bg{bg}_z{Z}_cw"], axis=1)[:, :N]
            for m, c, F in zip(msgs, cws, ex[f"bg{bg}_z{Z}_filler"]):
                assert (ctx.ldpc_encode(bg, Z, m) == c).all(), (bg, Z)
                llr = (10 - 20 * c.astype(np.int16)).astype(np.int8)
                if F:
                    llr[K - 2 * Z - F:K - 2 * Z] = 127
                for n in sorted({K + 2 * Z, (K + 2 * Z + N) // 2, N}):
                    it, bits = ctx.ldpc_decode(bg, Z, llr[:n], int(F), po.CRC_NONE, 1)
                    assert it == 0 and (np.unpackbits(bits)[:K] == m).all(), (bg, Z, n)
                batch.append((bg, Z, llr, int(F), po.CRC_NONE, 1, False))
                want.append(m)
        out = _decode_batch(ctx, batch)
        for i, m in enumerate(want):
            assert (np.unpackbits(out["cb_bits"][i])[:m.size] == m).all(), (bg, batch[i][1])


@pytest.mark.parametrize("variant,scale", [("auto", capi.SCALE_X86), ("generic", capi.SCALE_GENERIC)])
def test_decoder_vs_reference_vectors(orc, variant, scale):
    # Outputs of the compiled reference decoders ("auto" = AVX512/AVX2, and generic): iteration count and bits.
    d = np.load(G / "ref_decoder.npz")
    c = capi.Context(device=0, max_cbs=256, harq_entries=256, max_tbs=1, max_tb_bytes=4096, scale_mode=scale,
                     combine_simd_width=0 if variant == "generic" else 64)
    off = boff = 0
    batch, want = [], []
    n_pass = 0
    for (bg, Z, F, crc_kind, mi, n), it_ref in zip(d["cases"], d[f"iters_{variant}"]):
        K = (22 if bg == 1 else 10) * Z
        kb = (K + 7) // 8
        llr = d["llrs"][off:off + n]
        bits_ref = d[f"bits_{variant}"][boff:boff + kb]
        off += n
        boff += kb
        it, bits = c.ldpc_decode(int(bg), int(Z), llr, int(F), int(crc_kind), int(mi))
        assert it == it_ref and (bits == bits_ref).all(), (bg, Z, F, crc_kind, mi, n)
        n_pass += it > 0
        if n % Z == 0:
            # (the reference's tail handling of a partial node is state dependent, SURVEY 8a R8: single-codeblock call only)
            batch.append((int(bg), int(Z), llr, int(F), int(crc_kind), int(mi), crc_kind != po.CRC_NONE))
            want.append((int(it_ref), bits_ref, int(crc_kind), int(mi)))
    assert n_pass > 10  # the vectors exercise the early stop
    assert len(batch) > 20
    out = _decode_batch(c, batch)
    for i, (it_ref, bits_ref, kind, mi) in enumerate(want):
        r = out["cb_results"][i]
        assert (out["cb_bits"][i][:bits_ref.size] == bits_ref).all(), (i, batch[i][:2])
        if kind != po.CRC_NONE:
            # ldpc_decoder::decode returns the iteration count iff the CRC passed (0 in the fixture otherwise)
            assert bool(r["crc_ok"]) == (it_ref > 0) and int(r["iters"]) == (it_ref if it_ref > 0 else mi), (i, r)
    c.close()


def test_dematcher_vs_reference_vectors(ctx):
    # HARQ buffer before / after ldpc_rate_dematcher of the compiled reference (incl. stale regions and LBRM).
    d = np.load(G / "ref_dematcher.npz")
    assert int(d["simd_width"]) == 64
    bo = lo = 0
    for N, E, new_data, rv, qm, nref, F in d["cases"]:
        buf = d["buf0"][bo:bo + N].copy()
        ctx.rate_dematch(buf, d["llrs"][lo:lo + E], bool(new_data), int(rv), int(qm), int(nref), int(F))
        assert (buf == d["buf1"][bo:bo + N]).all(), (N, E, new_data, rv, qm, nref, F)
        bo += N
        lo += E


def test_pusch_decoder_vs_reference_vectors(ctx, orc):
    # pusch_decoder_impl of the compiled reference over rv 0-2-3-1: TB CRC, TB bytes, statistics, and a CRC32 of every
    # codeblock's soft buffer after each (re)transmission - against the batched GPU pusch_decoder.
    from srsran_edgeric_5g_b200.pusch_decoder import (PuschDecoderBatch, pusch_decoder_configuration,
                                                      pusch_decoder_notifier_spy, rx_buffer, rx_buffer_pool)
    d = np.load(G / "ref_pusch.npz")
    to = lo = so = oo = 0
    n_ok = 0
    for case, (bg, tb_bytes, qm, nl, n_llr, nref, es, mi, fill, C) in enumerate(d["cases"]):
        tb = d["tbs"][to:to + tb_bytes]
        to += tb_bytes
        # Like the fixture's driver, ONE buffer serves all four transmissions whatever the CRC says (its codeblock CRC
        # flags persist: once a codeblock passed it is only combined), so "release" just unlocks.
        pool = rx_buffer_pool(ctx, first_entry=0, nof_entries=256)
        buf = rx_buffer(pool, ("case", case), list(range(int(C))))
        buf.release = buf.unlock
        batch = PuschDecoderBatch(ctx)
        N = orc.segment_rx(int(tb_bytes) * 8, int(bg), int(qm), int(nl), int(n_llr))[0].full_length
        for t, rv in enumerate([0, 2, 3, 1]):
            llr = d["llrs"][lo:lo + n_llr]
            lo += n_llr
            buf.lock()
            if t == 0:
                for k in range(C):
                    ctx.harq_write(buf.get_absolute_codeblock_id(k), np.full(capi.PDC_MAX_CB_SOFT, fill, np.int8))
            spy = pusch_decoder_notifier_spy()
            rx = np.zeros(int(tb_bytes), np.uint8)
            dec = batch.create()
            cfg = pusch_decoder_configuration(int(bg), rv, int(qm), int(nref), int(nl), int(mi), bool(es), t == 0)
            b = dec.new_data(rx, buf, spy, cfg)
            b.on_new_softbits(llr)
            b.on_end_softbits()
            batch.flush()
            r = spy.get_entries()[0]
            st = d["stats"][so]  # {tb_crc_ok, nof_codeblocks_total, nof observations, min, max}
            assert r.tb_crc_ok == bool(st[0]) and r.nof_codeblocks_total == st[1], (case, t)
            assert r.ldpc_decoder_stats.get_nof_observations() == st[2], (case, t)
            if st[2]:
                assert r.ldpc_decoder_stats.get_min() == st[3] and r.ldpc_decoder_stats.get_max() == st[4], (case, t)
            if st[0]:
                assert (rx == d["tb_out"][oo:oo + tb_bytes]).all() and (rx == tb).all(), (case, t)
                n_ok += 1
            crc = [zlib.crc32(ctx.harq_read(buf.get_absolute_codeblock_id(k), N).tobytes()) for k in range(C)]
            assert crc == list(d["soft_crc32"][so][:C]), (case, t)
            so += 1
            oo += tb_bytes
    assert n_ok > 5
