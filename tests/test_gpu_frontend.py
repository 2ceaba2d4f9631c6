"""GPU parity of the codeword front end (scrambling sequence, descrambling, UL-SCH demultiplexing) through the C ABI
(pdc_scrambling_sequence, pdc_ulsch_demux, pdc_submit_codewords + pdc_submit) against the oracle restatement of
pseudo_random_generator_impl / pusch_demodulator_impl::revert_scrambling / ulsch_demultiplex_impl."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi, ldpc
from tests.vectors import make_tb_llrs, ulsch_case

pytestmark = pytest.mark.gpu


def cw_desc(cfg, in_offset=0, sch_offset=0, uci_offset=0, c_init=0, flags=0):
    d = np.zeros(1, capi.CW_DESC_DTYPE)[0]
    d["in_offset"], d["sch_offset"], d["uci_offset"], d["c_init"], d["flags"] = in_offset, sch_offset, uci_offset, c_init, flags
    for k in po.ULSCH_CFG_FIELDS:
        d[k] = cfg.get(k, 0)
    return d


def test_scrambling_sequence(ctx, orc):
    for c_init, offset, n in [(0, 0, 64), (1, 0, 1), (12345, 0, 4000), (0x7FFFFFFF, 777, 3001), (1 << 30, 100000, 2000),
                              (4660 * 32768 + 321, 0, 1362816), (99 << 15, 2000000, 90000)]:
        got = ctx.scrambling_sequence(c_init, offset, n)
        assert (got == orc.prg_bits(c_init, offset, n)).all(), (c_init, offset, n)


def check_streams(res, sch, uci, d, outs):
    assert [int(res["n_sch"]), int(res["n_harq_ack"]), int(res["n_csi_part1"]), int(res["n_csi_part2"])] == \
           [o.size for o in outs]
    so, uo = int(d["sch_offset"]), int(d["uci_offset"])
    assert (sch[so:so + outs[0].size] == outs[0]).all()
    want_uci = np.concatenate(outs[1:])
    assert (uci[uo:uo + want_uci.size] == want_uci).all()


def test_ulsch_demux_random_configurations(ctx, orc):
    """Both entry conditions: descrambled input with the sequence supplied (what ulsch_demultiplex::on_new_block
    receives), and scrambled input with the sequence generated on the device from c_init."""
    rng = np.random.default_rng(5)
    for trial in range(120):
        cfg, llr, _, _ = ulsch_case(orc, rng)
        c_init = int(rng.integers(0, 1 << 31))
        seq = orc.prg_bits(c_init, 0, llr.size)
        rc, outs = orc.ulsch_demux(cfg, llr, seq)
        assert rc == 0
        # (a) descrambled input + caller-supplied sequence
        res, sch, uci = ctx.ulsch_demux(np.array([cw_desc(cfg)]), llr, np.packbits(seq))
        check_streams(res[0], sch, uci, cw_desc(cfg), outs)
        # (b) scrambled input, sequence from c_init
        raw = orc.revert_scrambling(llr, seq)
        d = cw_desc(cfg, c_init=c_init, flags=capi.CW_SCRAMBLED)
        res, sch, uci = ctx.ulsch_demux(np.array([d]), raw)
        check_streams(res[0], sch, uci, d, outs)


def test_ulsch_demux_batch_of_codewords(ctx, orc):
    """Several codewords in one call, each at its own offsets; bytes between the codewords' outputs stay untouched."""
    rng = np.random.default_rng(6)
    descs, raws, wants = [], [], []
    in_off = sch_off = uci_off = 0
    for k in range(12):
        cfg, llr, _, _ = ulsch_case(orc, rng, max_prb=30)
        c_init = int(rng.integers(0, 1 << 31))
        seq = orc.prg_bits(c_init, 0, llr.size)
        rc, outs = orc.ulsch_demux(cfg, llr, seq)
        d = cw_desc(cfg, in_off, sch_off, uci_off, c_init, capi.CW_SCRAMBLED)
        descs.append(d)
        raws.append(orc.revert_scrambling(llr, seq))
        wants.append(outs)
        in_off += llr.size
        sch_off += (outs[0].size + 3) // 4 * 4 + 8
        uci_off += sum(o.size for o in outs[1:]) + 3
    res, sch, uci = ctx.ulsch_demux(np.array(descs), np.concatenate(raws), sch_capacity=sch_off, uci_capacity=uci_off + 1)
    for d, r, outs in zip(descs, res, wants):
        check_streams(r, sch, uci, d, outs)
    # gaps were not written
    for d, outs in zip(descs, wants):
        end = int(d["sch_offset"]) + outs[0].size
        assert (sch[end:(end + 3) // 4 * 4 + 8][:8] == 0).all()


def test_ulsch_demux_golden_vectors(ctx):
    gold = np.load(Path(__file__).parent / "golden" / "ref_frontend.npz")
    p_llr = p_seq = p_out = 0
    for cfg_arr, lens in zip(gold["cfgs"], gold["lens"]):
        cfg = dict(zip(po.ULSCH_CFG_FIELDS, (int(v) for v in cfg_arr)))
        n = int(lens[0])
        llr = gold["llrs"][p_llr:p_llr + n]
        nb = (n + 7) // 8
        seq_packed = gold["seq_bits"][p_seq:p_seq + nb]
        p_llr += n
        p_seq += nb
        outs = []
        for k in range(4):
            outs.append(gold["outs"][p_out:p_out + int(lens[1 + k])])
            p_out += int(lens[1 + k])
        res, sch, uci = ctx.ulsch_demux(np.array([cw_desc(cfg)]), llr, seq_packed)
        check_streams(res[0], sch, uci, cw_desc(cfg), outs)


def source_index_of_sch(orc, cfg, n):
    """For every UL-SCH output soft bit, the input soft bit it comes from (-1 = punctured), found by pushing the digits
    of the input index through the oracle demultiplexer."""
    idx = np.arange(n)
    src = np.zeros(0, np.int64)
    digits = []
    for k in range(4):
        d = ((idx >> (6 * k)) & 63).astype(np.int8) + 1  # 1..64, never 0 (punctured elements read 0)
        rc, outs = orc.ulsch_demux(cfg, d, np.zeros(n, np.uint8))
        assert rc == 0
        digits.append(outs[0].astype(np.int64))
    punct = digits[0] == 0
    src = sum((dg - 1) << (6 * k) for k, dg in enumerate(digits))
    src[punct] = -1
    return src


@pytest.mark.parametrize("ack_bits", [0, 2, 7])
def test_front_end_feeds_the_decoder(ctx, orc, ack_bits):
    """pdc_submit_codewords + pdc_submit(llrs = NULL): a 4-layer 256QAM slot with HARQ-ACK (punctured or rate matched
    around) and CSI Part 1 multiplexed in, scrambled; the transport block must come out of the device chain and every
    stream must equal the oracle's."""
    rng = np.random.default_rng(40 + ack_bits)
    qm, nl, nprb = 8, 4, 60
    cfg = dict(qm=qm, nof_layers=nl, nof_prb=nprb, start_symbol_index=0, nof_symbols=14, dmrs_type=1,
               dmrs_symbol_mask=1 << 2, nof_cdm_groups_without_data=2, nof_harq_ack_bits=ack_bits,
               nof_enc_harq_ack_bits=(40 * qm * nl if ack_bits else 0),
               nof_harq_ack_rvd=(60 * qm * nl if ack_bits <= 2 else 0), nof_csi_part1_bits=11,
               nof_enc_csi_part1_bits=90 * qm * nl)
    n = orc.ulsch_codeword_length(cfg)
    src = source_index_of_sch(orc, cfg, n)
    n_sch = src.size
    tbs_bits = int(n_sch * 0.8) // 8 * 8
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
    sch_llr, _ = make_tb_llrs(orc, tb, 1, 0, qm, nref, nl, n_sch, 9.5, rng)
    llr = rng.integers(-100, 101, n).astype(np.int8)  # UCI elements carry arbitrary soft bits
    llr[src[src >= 0]] = sch_llr[src >= 0]
    c_init = 0x4601 * 32768 + 77
    seq = orc.prg_bits(c_init, 0, n)
    rc, outs = orc.ulsch_demux(cfg, llr, seq)
    assert rc == 0 and outs[0].size == n_sch
    raw = capi.PinnedBuffer(n)
    raw.array[:] = orc.revert_scrambling(llr, seq)

    metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_sch)
    cbs = np.zeros(C, capi.CB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    for k, m in enumerate(metas):
        cbs[k] = (m.cw_offset, m.rm_length, 100 + k, nref, m.lifting_size, m.nof_filler_bits, 1, qm, 0, capi.CRC24B, 6,
                  flags, 0)
    tbd = np.zeros(1, capi.TB_DESC_DTYPE)
    tbd[0] = (0, C, tbs_bits, 0, 0)
    d = cw_desc(cfg, c_init=c_init, flags=capi.CW_SCRAMBLED)
    for k in range(C):  # entries other tests may have used: start from empty soft buffers like a fresh rx_buffer
        ctx.harq_write(100 + k, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
    ctx.submit_codewords(np.array([d]), raw.array, stream=0)
    ctx.submit(cbs, None, tbd, stream=0)
    out = ctx.wait(0)
    fe = out["codewords"]
    want_uci = np.concatenate(outs[1:])
    assert int(fe["cw_results"][0]["n_sch"]) == n_sch
    assert (fe["uci"][:want_uci.size] == want_uci).all()
    # the decoder saw exactly the oracle's UL-SCH stream: same codeblock results as decoding that stream directly
    ctx.submit(cbs, outs[0], tbd, stream=1)
    direct = ctx.wait(1)
    assert (out["cb_results"] == direct["cb_results"]).all()
    assert (out["cb_bits"] == direct["cb_bits"]).all()
    assert out["tb_results"][0]["tb_crc_ok"] == 1, out["cb_results"]
    assert (out["tb_bytes"][:tbs_bits // 8] == tb).all()


@pytest.mark.parametrize("qm,nl,nprb,rate", [(1, 1, 30, 0.3), (2, 1, 25, 0.2), (2, 2, 50, 0.1), (4, 2, 40, 0.5),
                                             (6, 3, 33, 0.7), (8, 4, 60, 0.85)])
def test_deferred_descrambling_in_the_dematcher(ctx, orc, qm, nl, nprb, rate):
    """PDC_CW_DEFER_DESCRAMBLING: codewords without UCI stay scrambled and the rate dematcher descrambles while it
    stages the codeblocks. Two transmissions (new data, then a retransmission that combines) of two codewords in one
    batch - one without UCI (deferred), one with CSI Part 1 (materialised) - must leave exactly the codeblock results and
    HARQ soft bits of decoding the descrambled UL-SCH streams directly."""
    rng = np.random.default_rng(qm * 10 + nl)
    base = dict(qm=qm, nof_layers=nl, nof_prb=nprb, start_symbol_index=0, nof_symbols=14, dmrs_type=1,
                dmrs_symbol_mask=1 << 2, nof_cdm_groups_without_data=2)
    cfgs = [dict(base), dict(base, nof_csi_part1_bits=7, nof_enc_csi_part1_bits=30 * qm * nl)]
    bg = 1 if rate > 0.3 else 2
    state = []
    for k, cfg in enumerate(cfgs):
        n = orc.ulsch_codeword_length(cfg)
        src = source_index_of_sch(orc, cfg, n)
        n_sch = src.size
        tbs_bits = max(24, int(n_sch * rate) // 8 * 8)
        if bg == 2:
            tbs_bits = min(tbs_bits, 3824)
        C = ldpc.compute_nof_codeblocks(tbs_bits, bg)
        tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
        state.append(dict(cfg=cfg, n=n, src=src, n_sch=n_sch, tbs_bits=tbs_bits, C=C, tb=tb,
                          c_init=int(rng.integers(0, 1 << 31))))
    first_id = {0: 300, 1: 700}  # HARQ entries of the chain / of the direct decode
    for base in first_id.values():  # both sets start from the same (empty) soft buffers whatever ran before
        for i in range(sum(st["C"] for st in state)):
            ctx.harq_write(base + i, np.zeros(capi.PDC_MAX_CB_SOFT, np.int8))
    for tx, rv in enumerate([0, 2]):
        raws, schs, cws, cb_chain, cb_direct = [], [], [], [], []
        in_off = sch_off = uci_off = cb_pos = 0
        for k, st in enumerate(state):
            sch_llr, _ = make_tb_llrs(orc, st["tb"], bg, rv, qm, 0, nl, st["n_sch"], 2.0 + 6 * rate, rng)
            llr = rng.integers(-100, 101, st["n"]).astype(np.int8)
            llr[st["src"]] = sch_llr
            seq = orc.prg_bits(st["c_init"], 0, st["n"])
            rc, outs = orc.ulsch_demux(st["cfg"], llr, seq)
            assert rc == 0 and (outs[0] == sch_llr).all()
            raws.append(orc.revert_scrambling(llr, seq))
            schs.append((sch_off, sch_llr))
            cws.append(cw_desc(st["cfg"], in_off, sch_off, uci_off, st["c_init"],
                               capi.CW_SCRAMBLED | capi.CW_DEFER_DESCRAMBLING))
            flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_EARLY_STOP | (capi.CB_NEW_DATA if tx == 0 else 0)
            for i, m in enumerate(ldpc.segment_rx(st["tbs_bits"], bg, rv, qm, 0, nl, st["n_sch"])):
                crc = capi.CRC24B if st["C"] > 1 else (capi.CRC24A if st["tbs_bits"] > 3824 else capi.CRC16)
                row = (sch_off + m.cw_offset, m.rm_length, 0, 0, m.lifting_size, m.nof_filler_bits, bg, qm, rv, crc, 6,
                       flags, 0xffff)
                cb_chain.append((row, first_id[0] + cb_pos))
                cb_direct.append((row, first_id[1] + cb_pos))
                cb_pos += 1
            in_off += st["n"]
            sch_off += (st["n_sch"] + 15) // 16 * 16
            uci_off += 4096
        def descs(rows):
            a = np.zeros(len(rows), capi.CB_DESC_DTYPE)
            for i, (row, hid) in enumerate(rows):
                a[i] = row
                a[i]["harq_id"] = hid
            return a
        raw = np.concatenate(raws)
        ctx.submit_codewords(np.array(cws), raw, stream=0)
        ctx.submit(descs(cb_chain), None, stream=0)
        chain = ctx.wait(0)
        sch_space = np.zeros(sch_off, np.int8)
        for off, v in schs:
            sch_space[off:off + v.size] = v
        ctx.submit(descs(cb_direct), sch_space, stream=1)
        direct = ctx.wait(1)
        assert (chain["cb_results"] == direct["cb_results"]).all(), (tx, chain["cb_results"], direct["cb_results"])
        assert (chain["cb_bits"] == direct["cb_bits"]).all()
        for i in range(cb_pos):
            assert (ctx.harq_read(first_id[0] + i) == ctx.harq_read(first_id[1] + i)).all(), (tx, i)
    assert chain["cb_results"]["crc_ok"].all()
