"""CPU tests of the host-side logic: segmentation arithmetic, HARQ buffer pool, decoder state machine, sharding."""
import numpy as np
import pytest

from srsran_edgeric_5g_b200 import ldpc, sharding
from srsran_edgeric_5g_b200.channel_coding import (create_crc_calculator_factory_sw, create_ldpc_decoder_factory_sw,
                                                   create_ldpc_rate_dematcher_factory_sw)
from srsran_edgeric_5g_b200.pusch_decoder import PuschDecoderBatch, pusch_decoder_configuration, rx_buffer_pool


def test_segmentation_matches_oracle(orc):
    rng = np.random.default_rng(7)
    for trial in range(300):
        bg = int(rng.integers(1, 3))
        tb_bytes = int(rng.integers(1, 3000 if bg == 2 else 20000)) if rng.random() < 0.9 else int(
            rng.integers(100000, 159749))
        if bg == 2 and tb_bytes * 8 > 30000:
            tb_bytes = 3000
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        nl = int(rng.integers(1, 5))
        nsym = int(np.ceil(tb_bytes * 8 / rng.uniform(0.15, 0.92) / qm / nl)) * nl
        n_llr = nsym * qm
        a = ldpc.segment_rx(tb_bytes * 8, bg, 0, qm, 0, nl, n_llr)
        b = orc.segment_rx(tb_bytes * 8, bg, qm, nl, n_llr)
        assert len(a) == len(b) == ldpc.compute_nof_codeblocks(tb_bytes * 8, bg)
        for x, y in zip(a, b):
            assert (x.lifting_size, x.full_length, x.rm_length, x.nof_filler_bits, x.cw_offset, x.nof_crc_bits) == (
                y.Z, y.full_length, y.rm_length, y.nof_filler, y.cw_offset, y.nof_crc_bits)


def test_config3_numbers():
    # 273 PRB x 13 symbols x 12 x 4 layers x 256QAM.
    tbs, n_llr = 1277992, 1362816
    metas = ldpc.segment_rx(tbs, ldpc.BG1, 0, 8, 0, 4, n_llr)
    assert len(metas) == 152 and metas[0].lifting_size == 384 and metas[0].nof_filler_bits == 16
    assert sorted({m.rm_length for m in metas}) == [8960, 8992]
    assert sum(m.rm_length == 8960 for m in metas) == 124
    assert ldpc.compute_N_ref(tbs // 8, 152) == 12611


def test_unknown_factory_types_return_none():
    # The reference returns nullptr for unknown types (channel_coding_factories.cpp:100-124, :166-192).
    assert create_ldpc_decoder_factory_sw("avx2") is None
    assert create_ldpc_rate_dematcher_factory_sw("generic") is None
    assert create_crc_calculator_factory_sw("lut") is None


class _FakeCtx:
    class cfg:
        harq_entries = 8

    def harq_free(self, i):
        pass


def test_rx_buffer_pool_semantics():
    pool = rx_buffer_pool(_FakeCtx())
    a = pool.reserve(None, (1, 0), 3, True)
    assert a is not None and a.get_nof_codeblocks() == 3 and a.locked
    assert pool.reserve(None, (1, 0), 3, False) is None  # locked
    a.unlock()
    assert pool.reserve(None, (2, 0), 6, True) is None  # not enough codeblocks left
    b = pool.reserve(None, (1, 0), 3, False)
    assert b is a  # retransmission gets the same entries: HARQ state is sticky
    ids = [b.get_absolute_codeblock_id(k) for k in range(3)]
    b.unlock()
    assert pool.reserve(None, (1, 0), 2, False) is None  # different number of codeblocks on a retransmission
    c = pool.reserve(None, (1, 0), 2, True)  # new data may change the size
    assert c is not None and c.get_nof_codeblocks() == 2
    c.release()
    assert pool.reserve(None, (1, 0), 2, False) is None  # released buffers are gone
    d = pool.reserve(None, (3, 1), 8, True)
    assert d is not None and sorted(d.get_absolute_codeblock_id(k) for k in range(8)) == list(range(8))
    assert len(set(ids)) == 3


def test_pusch_decoder_state_machine():
    batch = PuschDecoderBatch(ctx=None)
    dec = batch.create()
    pool = rx_buffer_pool(_FakeCtx())
    buf = pool.reserve(None, (1, 0), 1, True)
    cfg = pusch_decoder_configuration(ldpc.BG2, 0, 2, 0, 1)
    with pytest.raises(RuntimeError):
        dec.on_new_softbits(np.zeros(4, np.int8))  # not configured yet
    with pytest.raises(ValueError):
        dec.new_data(np.zeros(1000, np.uint8), buf, None, cfg)  # 3 codeblocks needed, buffer has 1
    b = dec.new_data(np.zeros(20, np.uint8), buf, None, cfg)
    with pytest.raises(RuntimeError):
        dec.new_data(np.zeros(20, np.uint8), buf, None, cfg)  # already collecting
    b.on_new_softbits(np.zeros(300, np.int8))
    dec.set_nof_softbits(400)
    with pytest.raises(ValueError):
        b.on_end_softbits()  # 300 != 400
    b.on_new_softbits(np.zeros(100, np.int8))
    b.on_end_softbits()
    assert batch.pending() == 1


def test_sharding_is_a_sticky_partition():
    tbs = [dict(cell=c, rnti=0x4601 + u) for c in range(16) for u in range(5)]
    for world in (1, 2, 4, 8):
        sharding.check_partition(tbs, world)
        parts = [sharding.shard_transport_blocks(tbs, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(len(tbs)))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 5  # 16 cells -> 2 per GPU at 8


# ---- host-side plan of the UL-SCH demultiplexing (csrc/ulsch_plan.h) -------------------------------------------------

def _plan_check_lib():
    """tests/host/ulsch_plan_check.cpp: the plan executed on the CPU with the index arithmetic of the device gathers."""
    import ctypes
    import subprocess
    from pathlib import Path
    here = Path(__file__).resolve().parent
    out = here / "_build" / "libulsch_plan_check.so"
    src = here / "host" / "ulsch_plan_check.cpp"
    hdr = here.parent / "srsran_edgeric_5g_b200" / "csrc" / "ulsch_plan.h"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        out.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    return ctypes.CDLL(str(out))


def test_ulsch_plan_matches_oracle(orc):
    import ctypes
    from srsran_edgeric_5g_b200 import capi
    from oracle import pyoracle as po
    from tests.vectors import ulsch_case
    L = _plan_check_lib()
    rng = np.random.default_rng(8)
    vp = ctypes.c_void_p
    for _ in range(400):
        cfg, llr, seq, outs = ulsch_case(orc, rng)
        d = np.zeros(1, capi.CW_DESC_DTYPE)
        for k in po.ULSCH_CFG_FIELDS:
            d[0][k] = cfg.get(k, 0)
        sch, uci, n = np.zeros(llr.size, np.int8), np.zeros(llr.size, np.int8), np.zeros(4, np.uint32)
        rc = L.plan_demux_cpu(d.ctypes.data_as(vp), llr.ctypes.data_as(vp), seq.ctypes.data_as(vp), sch.ctypes.data_as(vp),
                              uci.ctypes.data_as(vp), n.ctypes.data_as(vp))
        assert rc == 0 and list(n) == [o.size for o in outs], cfg
        want = np.concatenate(outs[1:])
        assert (sch[:n[0]] == outs[0]).all() and (uci[:want.size] == want).all(), cfg


def test_ulsch_plan_rejects_inconsistent_descriptions():
    import ctypes
    from srsran_edgeric_5g_b200 import capi
    L = _plan_check_lib()
    d = np.zeros(1, capi.CW_DESC_DTYPE)
    d[0]["qm"], d[0]["nof_layers"], d[0]["nof_prb"], d[0]["nof_symbols"], d[0]["dmrs_type"] = 2, 1, 4, 14, 1
    d[0]["nof_cdm_groups_without_data"] = 2
    buf = np.zeros(4096, np.int8)
    n = np.zeros(4, np.uint32)
    vp = ctypes.c_void_p
    args = [buf.ctypes.data_as(vp)] * 4 + [n.ctypes.data_as(vp)]
    assert L.plan_demux_cpu(d.ctypes.data_as(vp), *args) == -1            # no DM-RS symbol
    d[0]["dmrs_symbol_mask"] = 1 << 2
    d[0]["nof_harq_ack_bits"], d[0]["nof_enc_harq_ack_bits"] = 4, 1 << 20  # HARQ-ACK that cannot fit
    assert L.plan_demux_cpu(d.ctypes.data_as(vp), *args) == -1
    d[0]["qm"] = 3                                                          # not a modulation order
    assert L.plan_demux_cpu(d.ctypes.data_as(vp), *args) == -1


def test_tbs_and_base_graph_known_answers():
    """TS 38.214 5.1.3.2 / TS 38.212 6.2.2 through sch.py: the largest NR transport block (what BASELINE config 3 decodes:
    273 PRB, 4 layers, 256QAM MCS 27 -> 1 277 992 bits = MAX_TBS, ldpc_segmenter_impl.cpp:38) and the table end points."""
    from srsran_edgeric_5g_b200 import sch
    a = sch.pusch_allocation("qam256", 27, 273, nof_layers=4)
    assert a["tbs_bits"] == 1277992 and a["base_graph"] == 1 and a["n_llr"] == 1362816 and a["qm"] == 8
    assert sch.tbs_calculate(14, 12, 0, 2, 120, 1, 1) == 32          # one PRB at the lowest MCS
    assert sch.get_ldpc_base_graph(0.9, 292) == 2 and sch.get_ldpc_base_graph(0.9, 296) == 1
    assert sch.get_ldpc_base_graph(0.67, 3824) == 2 and sch.get_ldpc_base_graph(0.68, 3824) == 1
    assert sch.get_ldpc_base_graph(0.25, 100000) == 2 and sch.get_ldpc_base_graph(0.26, 100000) == 1
    # every size is byte aligned and above 3824 bits leaves equal codeblocks (step 4 of the procedure)
    for mcs in range(29):
        for nprb in (1, 6, 25, 52, 106, 273):
            t = sch.pusch_allocation("qam64", mcs, nprb)["tbs_bits"]
            assert t % 8 == 0 and t >= 24
            if t > 3824:
                C = ldpc.compute_nof_codeblocks(t, sch.pusch_allocation("qam64", mcs, nprb)["base_graph"])
                assert (t + 24) % C == 0


def test_tbs_matches_the_reference(ref_available):
    """sch.py against pusch_mcs_get_config + tbs_calculator_calculate + get_ldpc_base_graph of the compiled reference."""
    if not ref_available:
        pytest.skip("compiled reference not present")
    from oracle import pyoracle as po
    from srsran_edgeric_5g_b200 import sch
    ref = po.Reference()
    n = 0
    for table, T in (("qam64", sch.MCS_TABLE_QAM64), ("qam256", sch.MCS_TABLE_QAM256)):
        for mcs in range(len(T)):
            for nprb in (1, 2, 3, 5, 11, 24, 25, 26, 51, 52, 53, 79, 106, 133, 217, 273):
                for nsym, dmrs in ((14, 12), (14, 36), (12, 24), (7, 6)):
                    for nl in (1, 2, 4):
                        qm, r1024 = sch.pusch_mcs_get_config(table, mcs)
                        tbs = sch.tbs_calculate(nsym, dmrs, 0, qm, r1024, nl, nprb)
                        got = (tbs, sch.get_ldpc_base_graph(r1024 / 1024.0, tbs), qm, float(r1024))
                        assert got == ref.pusch_tbs(table, mcs, nsym, dmrs, 0, nl, nprb), (table, mcs, nprb, nsym, nl)
                        n += 1
    assert n > 10000
