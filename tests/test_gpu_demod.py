"""GPU parity of the soft demapper through the C ABI (pdc_demodulate_soft, pdc_launch_demod_device, pdc_submit_symbols)
against the oracle restatement of demodulation_mapper_impl (x86 build: SIMD blocks + scalar remainder per call) and the
golden vectors of the compiled reference. Bit-exact: the soft bits are int8."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi, ldpc
from tests.test_gpu_frontend import cw_desc, source_index_of_sch
from tests.vectors import DEMOD_MODS, demod_codeword, demod_inputs, make_tb_llrs, modulate, ofdm_symbol_sizes, ulsch_case

pytestmark = pytest.mark.gpu

GOLD = np.load(Path(__file__).parent / "golden" / "ref_demod.npz")


def test_demodulate_soft_golden_vectors(ctx):
    p_s = p_l = 0
    for mod, kind, n in GOLD["cases"]:
        mod, n = int(mod), int(n)
        q = max(mod, 1)
        got = ctx.demodulate_soft(GOLD["symbols"][p_s:p_s + n], GOLD["noise_vars"][p_s:p_s + n], mod)
        assert (got == GOLD["llrs"][p_l:p_l + n * q]).all(), (mod, int(kind), n)
        p_s += n
        p_l += n * q


def test_demodulate_soft_vs_oracle(ctx, orc):
    """Every modulation, call sizes from one symbol to several tiles, all four input families (incl. NaN / infinity /
    denormals / interval boundaries / rounding ties)."""
    rng = np.random.default_rng(31)
    sizes = [1, 2, 3, 4, 5, 15, 16, 17, 31, 255, 256, 257, 1023, 1024, 1025, 3276 * 4, 5000]
    for it in range(240):
        mod = DEMOD_MODS[it % 6]
        n = sizes[it % len(sizes)] if it < 200 else int(rng.integers(1, 9000))
        s, nv = demod_inputs(rng, n, mod, (it // 6) % 4)
        got = ctx.demodulate_soft(s, nv, mod)
        want = orc.demodulate_soft(s, nv, mod)
        assert (got == want).all(), (it, mod, n, np.nonzero(got != want)[0][:5])


def test_scalar_build_of_the_reference(orc):
    """PDC_DEMOD_SCALAR reproduces the portable build (scalar loop for every symbol)."""
    c2 = capi.Context(device=0, max_cbs=8, harq_entries=8, max_tbs=1, max_tb_bytes=1024, demod_mode=capi.DEMOD_SCALAR)
    rng = np.random.default_rng(32)
    for it in range(48):
        mod = DEMOD_MODS[it % 6]
        s, nv = demod_inputs(rng, int(rng.integers(1, 700)), mod, it % 4)
        assert (c2.demodulate_soft(s, nv, mod) == orc.demodulate_soft(s, nv, mod, simd=False)).all(), (it, mod)
    c2.close()


def test_batch_of_calls_on_device_buffers(ctx, orc):
    """pdc_launch_demod_device: many calls of mixed modulation in one launch, soft-bit offsets that are not multiples of
    16 (byte-wise store path) and calls that are."""
    import torch
    rng = np.random.default_rng(33)
    calls, syms, nvs, want = [], [], [], []
    sym_pos = llr_pos = 0
    for k in range(40):
        mod = int(rng.choice(DEMOD_MODS))
        n = int(rng.integers(1, 2600))
        s, nv = demod_inputs(rng, n, mod, k % 3)
        llr_pos += int(rng.integers(0, 9)) if k % 2 else (-llr_pos) % 16
        calls.append((sym_pos, n, llr_pos, mod))
        want.append((llr_pos, orc.demodulate_soft(s, nv, mod)))
        syms.append(s)
        nvs.append(nv)
        sym_pos += n
        llr_pos += n * max(mod, 1)
    d_sym = torch.from_numpy(np.concatenate(syms).view(np.float32)).cuda()
    d_nv = torch.from_numpy(np.concatenate(nvs)).cuda()
    d_llr = torch.full((llr_pos + 16,), 77, dtype=torch.int8, device="cuda")
    torch.cuda.synchronize()
    ctx.launch_demod_device(np.array(calls, capi.DEMOD_CALL_DTYPE), d_sym.data_ptr(), d_nv.data_ptr(), sym_pos,
                            d_llr.data_ptr(), llr_pos, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_llr.cpu().numpy()
    covered = np.zeros(got.size, bool)
    for off, w in want:
        assert (got[off:off + w.size] == w).all()
        covered[off:off + w.size] = True
    assert (got[~covered] == 77).all()  # nothing written outside the calls


def test_more_calls_than_one_launch_carries_and_empty_calls(ctx, orc):
    """The call table travels in the kernel parameters, 224 calls per launch: a batch of 500 calls (some of them empty)
    goes out as three launches; an empty demodulate_soft call is a no-op."""
    import torch
    rng = np.random.default_rng(35)
    calls, syms, nvs, want = [], [], [], []
    sym_pos = llr_pos = 0
    for k in range(500):
        mod = int(rng.choice(DEMOD_MODS))
        n = 0 if k % 37 == 5 else int(rng.integers(1, 70))
        s, nv = demod_inputs(rng, n, mod, k % 3) if n else (np.zeros(0, np.complex64), np.zeros(0, np.float32))
        calls.append((sym_pos, n, llr_pos, mod))
        want.append(orc.demodulate_soft(s, nv, mod) if n else np.zeros(0, np.int8))
        syms.append(s)
        nvs.append(nv)
        sym_pos += n
        llr_pos += n * max(mod, 1)
    d_sym = torch.from_numpy(np.concatenate(syms).view(np.float32)).cuda()
    d_nv = torch.from_numpy(np.concatenate(nvs)).cuda()
    d_llr = torch.zeros(llr_pos + 16, dtype=torch.int8, device="cuda")
    before = ctx.launch_count()
    ctx.launch_demod_device(np.array(calls, capi.DEMOD_CALL_DTYPE), d_sym.data_ptr(), d_nv.data_ptr(), sym_pos,
                            d_llr.data_ptr(), llr_pos, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert ctx.launch_count() - before == 3
    assert (d_llr.cpu().numpy()[:llr_pos] == np.concatenate(want)).all()
    assert ctx.demodulate_soft(np.zeros(0, np.complex64), np.zeros(0, np.float32), 4).size == 0


def test_symbols_bpsk_flag(ctx, orc):
    """qm = 1 through pdc_submit_symbols: pi/2-BPSK unless the codeword says plain BPSK (PDC_CW_PLAIN_BPSK)."""
    rng = np.random.default_rng(36)
    cfg = dict(qm=1, nof_layers=1, nof_prb=7, start_symbol_index=0, nof_symbols=14, dmrs_type=1, dmrs_symbol_mask=1 << 3,
               nof_cdm_groups_without_data=2)
    n = orc.ulsch_codeword_length(cfg)
    s, nv = demod_inputs(rng, n, 0, 0)
    c_init = 4242
    seq = orc.prg_bits(c_init, 0, n)
    for flag, mod in ((0, 0), (capi.CW_PLAIN_BPSK, 1)):
        raw = demod_codeword(orc, cfg, s, nv, mod=mod)
        want = orc.revert_scrambling(raw, seq)
        d = cw_desc(cfg, c_init=c_init, flags=capi.CW_SCRAMBLED | flag)
        ctx.harq_write(520, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
        ctx.submit_symbols(np.array([d]), np.zeros(1, np.uint32), s, nv, stream=0)
        cbs = np.zeros(1, capi.CB_DESC_DTYPE)
        cbs[0] = (0, n, 520, 0, 384, 0, 1, 1, 0, capi.CRC24B, 1, capi.CB_DEMATCH | capi.CB_NEW_DATA, 0xffff)
        ctx.submit(cbs, None, None, stream=0, want_bits=False)
        ctx.wait(0)
        buf = np.zeros(66 * 384, np.int8)
        orc.rate_dematch(buf, want, True, 0, 1)
        assert (ctx.harq_read(520) == buf).all(), mod


def test_invalid_calls_are_rejected(ctx):
    s = np.zeros(8, np.complex64)
    nv = np.ones(8, np.float32)
    with pytest.raises(capi.PdcError):
        ctx.demodulate_soft(s, nv, 3)
    with pytest.raises(capi.PdcError):
        ctx.demodulate_soft(s, nv, 10)


def test_symbols_through_random_codewords(ctx, orc):
    """pdc_submit_symbols on random PUSCH allocations (every modulation, 1-4 layers, UCI multiplexed in): the UL-SCH and
    UCI streams equal oracle demapper (one call per OFDM symbol) -> descrambling -> oracle demultiplexer. The UL-SCH
    stream stays on the device; it is observed through dematch-only codeblocks (HARQ buffer = oracle rate dematcher of
    the oracle's stream)."""
    rng = np.random.default_rng(34)
    for trial in range(40):
        cfg, _, _, _ = ulsch_case(orc, rng)
        qm = cfg["qm"]
        mod = 0 if qm == 1 else qm
        n_llr = orc.ulsch_codeword_length(cfg)
        s, nv = demod_inputs(rng, n_llr // qm, mod, trial % 3)
        raw = demod_codeword(orc, cfg, s, nv, mod=mod)
        c_init = int(rng.integers(0, 1 << 31))
        seq = orc.prg_bits(c_init, 0, n_llr)
        rc, outs = orc.ulsch_demux(cfg, orc.revert_scrambling(raw, seq), seq)
        assert rc == 0
        pad = 5 * (trial % 2)
        d = cw_desc(cfg, c_init=c_init, flags=capi.CW_SCRAMBLED, in_offset=16 * (trial % 3))
        ctx.submit_symbols(np.array([d]), np.array([pad], np.uint32), np.concatenate([np.zeros(pad, np.complex64), s]),
                           np.concatenate([np.ones(pad, np.float32), nv]), stream=0)
        n_sch = outs[0].size
        chunk = 20000 // qm * qm
        starts = list(range(0, n_sch, chunk))
        cbs = np.zeros(max(1, len(starts)), capi.CB_DESC_DTYPE)
        for k, st in enumerate(starts):
            cbs[k] = (st, min(chunk, n_sch - st), 500 + k, 0, 384, 0, 1, qm, 0, capi.CRC24B, 1,
                      capi.CB_DEMATCH | capi.CB_NEW_DATA, 0xffff)
        for k in range(len(starts)):
            ctx.harq_write(500 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))  # the oracle starts from an empty buffer
        if n_sch:
            ctx.submit(cbs, None, None, stream=0, want_bits=False)
        out = ctx.wait(0)
        fe = out["codewords"] if n_sch else out
        res = fe["cw_results"][0]
        assert [int(res["n_sch"]), int(res["n_harq_ack"]), int(res["n_csi_part1"]), int(res["n_csi_part2"])] == \
               [o.size for o in outs]
        want_uci = np.concatenate(outs[1:])
        assert (fe["uci"][:want_uci.size] == want_uci).all(), trial
        for k, st in enumerate(starts):
            buf = np.zeros(66 * 384, np.int8)
            orc.rate_dematch(buf, outs[0][st:st + min(chunk, n_sch - st)], True, 0, qm)
            assert (ctx.harq_read(500 + k) == buf).all(), (trial, k)


@pytest.mark.parametrize("qm,nl,nprb,ack_bits", [(8, 4, 60, 2), (6, 2, 33, 0), (4, 1, 51, 7), (2, 1, 25, 0)])
def test_symbols_feed_the_decoder(ctx, orc, qm, nl, nprb, ack_bits):
    """The whole device chain from the equaliser's output: modulated transport block + noise -> pdc_submit_symbols ->
    pdc_submit(llrs = NULL). Every codeblock result, the decoded bits and the transport block must equal those of the
    same chain entered one step later with the ORACLE's soft bits (pdc_submit_codewords), and the TB must decode."""
    rng = np.random.default_rng(50 + qm)
    cfg = dict(qm=qm, nof_layers=nl, nof_prb=nprb, start_symbol_index=0, nof_symbols=14, dmrs_type=1,
               dmrs_symbol_mask=(1 << 2) | (1 << 11), nof_cdm_groups_without_data=2, nof_harq_ack_bits=ack_bits,
               nof_enc_harq_ack_bits=(40 * qm * nl if ack_bits else 0),
               nof_harq_ack_rvd=(60 * qm * nl if 0 < ack_bits <= 2 else 0), nof_csi_part1_bits=(11 if ack_bits else 0),
               nof_enc_csi_part1_bits=(30 * qm * nl if ack_bits else 0))
    n = orc.ulsch_codeword_length(cfg)
    src = source_index_of_sch(orc, cfg, n)
    n_sch = src.size
    tbs_bits = int(n_sch * 0.45) // 8 * 8
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
    cw_bits, _ = orc.tb_encode(tb, 1, 0, qm, nref, nl, n_sch)
    # codeword bits -> resource-element order (UCI elements carry random bits), scrambled, modulated, noise added
    c_init = 0x4601 * 32768 + 91
    seq = orc.prg_bits(c_init, 0, n)
    bits = rng.integers(0, 2, n).astype(np.uint8)
    bits[src[src >= 0]] = cw_bits[src >= 0]
    tx = modulate(bits ^ seq, qm)
    sigma2 = 10 ** (-{2: 10.0, 4: 17.0, 6: 23.0, 8: 29.0}[qm] / 10)
    noise = (rng.standard_normal(tx.size) + 1j * rng.standard_normal(tx.size)) * np.sqrt(sigma2 / 2)
    sym = capi.PinnedBuffer(tx.size * 8, np.complex64)
    sym.array[:] = (tx + noise).astype(np.complex64)
    nv = capi.PinnedBuffer(tx.size * 4, np.float32)
    nv.array[:] = (sigma2 * (1 + 0.2 * rng.random(tx.size))).astype(np.float32)
    assert sum(ofdm_symbol_sizes(cfg)) == tx.size

    metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_sch)
    cbs = np.zeros(C, capi.CB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    # a lone codeblock carries the transport-block CRC (ldpc_segmenter_rx: CRC16 up to 3824 bits, CRC24A above)
    crc_kind = capi.CRC24B if C > 1 else (capi.CRC16 if tbs_bits <= 3824 else capi.CRC24A)
    for k, m in enumerate(metas):
        cbs[k] = (m.cw_offset, m.rm_length, 200 + k, nref, m.lifting_size, m.nof_filler_bits, 1, qm, 0, crc_kind, 6,
                  flags, 0)
    tbd = np.zeros(1, capi.TB_DESC_DTYPE)
    tbd[0] = (0, C, tbs_bits, 0, 0)
    d = cw_desc(cfg, c_init=c_init, flags=capi.CW_SCRAMBLED)
    for k in range(C):  # entries reused from other tests: start from empty soft buffers like a fresh rx_buffer
        ctx.harq_write(200 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
        ctx.harq_write(400 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
    ctx.submit_symbols(np.array([d]), np.zeros(1, np.uint32), sym.array, nv.array, stream=0)
    ctx.submit(cbs, None, tbd, stream=0)
    out = ctx.wait(0)
    # the same slot entered with the oracle's soft bits
    raw = demod_codeword(orc, cfg, sym.array, nv.array)
    cbs2 = cbs.copy()
    cbs2["harq_id"] += 200
    ctx.submit_codewords(np.array([d]), raw, stream=1)
    ctx.submit(cbs2, None, tbd, stream=1)
    want = ctx.wait(1)
    assert (out["codewords"]["cw_results"] == want["codewords"]["cw_results"]).all()
    assert (out["codewords"]["uci"] == want["codewords"]["uci"]).all()
    assert (out["cb_results"] == want["cb_results"]).all()
    assert (out["cb_bits"] == want["cb_bits"]).all()
    for k in range(C):
        assert (ctx.harq_read(200 + k) == ctx.harq_read(400 + k)).all()
    assert out["tb_results"][0]["tb_crc_ok"] == 1, (out["cb_results"], np.mean((raw < 0) != (bits ^ seq)))
    assert (out["tb_bytes"][:tbs_bits // 8] == tb).all()


def test_symbols_with_deferred_descrambling(ctx, orc):
    """PDC_CW_DEFER_DESCRAMBLING after the demapper: the soft bits stay scrambled where the demapper wrote them and the
    rate dematcher descrambles - same results as the materialised path."""
    rng = np.random.default_rng(61)
    qm, nl, nprb = 8, 2, 40
    cfg = dict(qm=qm, nof_layers=nl, nof_prb=nprb, start_symbol_index=0, nof_symbols=14, dmrs_type=1,
               dmrs_symbol_mask=1 << 3, nof_cdm_groups_without_data=2)
    n = orc.ulsch_codeword_length(cfg)
    tbs_bits = int(n * 0.5) // 8 * 8
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
    cw_bits, _ = orc.tb_encode(tb, 1, 0, qm, nref, nl, n)
    c_init = 0x1234 * 32768 + 5
    seq = orc.prg_bits(c_init, 0, n)
    tx = modulate(cw_bits ^ seq, qm)
    sigma2 = 10 ** (-2.6)
    noise = (rng.standard_normal(tx.size) + 1j * rng.standard_normal(tx.size)) * np.sqrt(sigma2 / 2)
    sym = (tx + noise).astype(np.complex64)
    nv = np.full(tx.size, sigma2, np.float32)
    metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n)
    cbs = np.zeros(C, capi.CB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    for k, m in enumerate(metas):
        cbs[k] = (m.cw_offset, m.rm_length, 300 + k, nref, m.lifting_size, m.nof_filler_bits, 1, qm, 0, capi.CRC24B, 6,
                  flags, 0)
    tbd = np.zeros(1, capi.TB_DESC_DTYPE)
    tbd[0] = (0, C, tbs_bits, 0, 0)
    outs = []
    for fl in (capi.CW_SCRAMBLED, capi.CW_SCRAMBLED | capi.CW_DEFER_DESCRAMBLING):
        for k in range(C):
            ctx.harq_write(300 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
        d = cw_desc(cfg, c_init=c_init, flags=fl)
        ctx.submit_symbols(np.array([d]), np.zeros(1, np.uint32), sym, nv, stream=0)
        ctx.submit(cbs, None, tbd, stream=0)
        outs.append(ctx.wait(0))
        outs[-1]["harq"] = [ctx.harq_read(300 + k) for k in range(C)]
    assert (outs[0]["cb_results"] == outs[1]["cb_results"]).all()
    assert (outs[0]["cb_bits"] == outs[1]["cb_bits"]).all()
    assert all((a == b).all() for a, b in zip(outs[0]["harq"], outs[1]["harq"]))
    assert outs[0]["tb_results"][0]["tb_crc_ok"] == 1 and (outs[1]["tb_bytes"][:tbs_bits // 8] == tb).all()
    # and the HARQ buffers are those of the oracle chain
    raw = demod_codeword(orc, cfg, sym, nv)
    llr = orc.revert_scrambling(raw, seq)
    for k, m in enumerate(metas):
        buf = np.zeros(66 * m.lifting_size, np.int8)
        orc.rate_dematch(buf, llr[m.cw_offset:m.cw_offset + m.rm_length], True, 0, qm, nref, m.nof_filler_bits)
        assert (outs[1]["harq"][k][:buf.size] == buf).all(), k
