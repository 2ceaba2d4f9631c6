"""GPU parity of the batched pusch_decoder (-m gpu): transport-block level, HARQ retransmissions, several UEs per batch.
Mirrors the assertions of the reference's pusch_decoder_vectortest.cpp:327-395 (TB CRC, TB bytes, one statistic per
decoded codeblock, iteration bounds) and compares everything with the oracle's restatement of pusch_decoder_impl."""
import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi
from srsran_edgeric_5g_b200.ldpc import compute_N_ref, compute_nof_codeblocks
from srsran_edgeric_5g_b200.pusch_decoder import (PuschDecoderBatch, pusch_decoder_configuration,
                                                  pusch_decoder_notifier_spy, rx_buffer_pool)
from tests.vectors import make_tb_llrs

pytestmark = pytest.mark.gpu


def _run_tb_sequence(ctx, orc, pool, batch, ues, rv_seq, max_iter, early_stop, rng, fill):
    """ues: list of dicts(tb, bg, qm, nl, n_llr, nref, snr). Every (re)transmission of all UEs goes in one GPU batch.
    fill: value the UEs' HARQ entries are set to before the first transmission; None = left as they are."""
    harqs = []
    for u in ues:
        C = compute_nof_codeblocks(u["tb"].size * 8, u["bg"])
        harqs.append(po.Harq(C, 0 if fill is None else fill))
    done = [False] * len(ues)
    for t, rv in enumerate(rv_seq):
        spies, rx_tbs, expected = [], [], []
        for i, u in enumerate(ues):
            if done[i]:
                spies.append(None)
                rx_tbs.append(None)
                expected.append(None)
                continue
            llrs, C = make_tb_llrs(orc, u["tb"], u["bg"], rv, u["qm"], u["nref"], u["nl"], u["n_llr"], u["snr"], rng)
            cfg = pusch_decoder_configuration(u["bg"], rv, u["qm"], u["nref"], u["nl"], max_iter, early_stop, t == 0)
            buf = pool.reserve(None, ("ue", i), C, t == 0)
            assert buf is not None
            if t == 0 and fill is None:
                # the entries are taken as the previous owner left them (the reference's pool does not clear them)
                for k in range(C):
                    harqs[i].soft[k, :capi.PDC_MAX_CB_SOFT] = ctx.harq_read(buf.get_absolute_codeblock_id(k))
            elif t == 0:
                for k in range(C):
                    ctx.harq_write(buf.get_absolute_codeblock_id(k), np.full(capi.PDC_MAX_CB_SOFT, fill, np.int8))
            spy = pusch_decoder_notifier_spy()
            rx = np.zeros(u["tb"].size, np.uint8)
            dec = batch.create()
            b = dec.new_data(rx, buf, spy, cfg)
            half = llrs.size // 2
            b.on_new_softbits(llrs[:half])
            dec.set_nof_softbits(llrs.size)
            b.on_new_softbits(llrs[half:])
            b.on_end_softbits()
            spies.append((spy, buf, C))
            rx_tbs.append(rx)
            expected.append(orc.pusch_decode(harqs[i], llrs, u["tb"].size, u["bg"], rv, u["qm"], u["nref"], u["nl"],
                                             max_iter, early_stop, t == 0))
        assert all(s is None or not s[0].get_entries() for s in spies)  # nothing is notified before the flush
        batch.flush()
        for i, u in enumerate(ues):
            if done[i]:
                continue
            spy, buf, C = spies[i]
            assert len(spy.get_entries()) == 1
            r = spy.get_entries()[0]
            tb_o, st = expected[i]
            assert r.tb_crc_ok == bool(st[0]), (i, t)
            assert r.nof_codeblocks_total == st[1]
            assert r.ldpc_decoder_stats.get_nof_observations() == st[2]
            if st[2]:
                assert r.ldpc_decoder_stats.get_min() == st[3] and r.ldpc_decoder_stats.get_max() == st[4]
                assert sum(r.ldpc_decoder_stats._v) == st[5]
            assert (buf.get_codeblocks_crc() == harqs[i].crc_ok.astype(bool)).all()
            for k in range(C):
                N = (66 if u["bg"] == 1 else 50) * orc.segment_rx(u["tb"].size * 8, u["bg"], u["qm"], u["nl"],
                                                                   u["n_llr"])[0].Z
                assert (ctx.harq_read(buf.get_absolute_codeblock_id(k), N) == harqs[i].soft[k][:N]).all(), (i, t, k)
            if r.tb_crc_ok:
                assert (rx_tbs[i] == tb_o).all() and (rx_tbs[i] == u["tb"]).all()
                done[i] = True
    return done


@pytest.mark.parametrize("early_stop,max_iter", [(True, 6), (False, 2)])
def test_multi_ue_harq_batch(ctx, orc, early_stop, max_iter):
    rng = np.random.default_rng(17 + max_iter)
    pool = rx_buffer_pool(ctx, first_entry=512, nof_entries=1024)
    batch = PuschDecoderBatch(ctx)
    ues = []
    for i in range(10):
        bg = 1 if i % 3 else 2
        tb_bytes = int(rng.integers(20, 900 if bg == 2 else 5000))
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.integers(1, 3))
        rate = rng.uniform(0.55, 0.9) if bg == 1 else rng.uniform(0.25, 0.6)
        nsym = int(np.ceil(tb_bytes * 8 / rate / qm / nl)) * nl
        C = compute_nof_codeblocks(tb_bytes * 8, bg)
        nref = 0 if i % 2 else compute_N_ref(tb_bytes + 40, C)
        snr = (8 if bg == 1 else 3) * rate / 0.8 + rng.uniform(-4, -1)
        ues.append(dict(tb=rng.integers(0, 256, tb_bytes).astype(np.uint8), bg=bg, qm=qm, nl=nl, n_llr=nsym * qm,
                        nref=nref, snr=snr))
    done = _run_tb_sequence(ctx, orc, pool, batch, ues, [0, 2, 3, 1], max_iter, early_stop, rng, fill=-7)
    assert any(done)


def random_ues(rng, n_ue, max_tb_bytes_bg1=5000):
    ues = []
    for i in range(n_ue):
        bg = 1 if i % 3 else 2
        tb_bytes = int(rng.integers(20, 900 if bg == 2 else max_tb_bytes_bg1))
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.integers(1, 3))
        rate = rng.uniform(0.55, 0.9) if bg == 1 else rng.uniform(0.25, 0.6)
        nsym = int(np.ceil(tb_bytes * 8 / rate / qm / nl)) * nl
        C = compute_nof_codeblocks(tb_bytes * 8, bg)
        nref = 0 if rng.random() < 0.5 else compute_N_ref(tb_bytes + 40, C)
        snr = (8 if bg == 1 else 3) * rate / 0.8 + rng.uniform(-4, -1)
        ues.append(dict(tb=rng.integers(0, 256, tb_bytes).astype(np.uint8), bg=bg, qm=qm, nl=nl, n_llr=nsym * qm,
                        nref=nref, snr=snr))
    return ues


def tb_generations(ctx, orc, rng, pool, n_gen, n_ue, max_tb_bytes_bg1=5000):
    """Generations of UEs following each other on the SAME buffer-pool keys: every generation has transport blocks of
    other sizes (codeblock count, lifting size, limited buffer or not), so the pool's entries go from codeblock shape to
    codeblock shape without ever being cleared, as in the reference (rx_buffer_pool_impl.cpp:44). Each generation runs
    its HARQ process (rv 0-2-3-1) with all UEs of a transmission in one batch. Returns the transport blocks decoded."""
    batch = PuschDecoderBatch(ctx)
    n_done = 0
    for gen in range(n_gen):
        ues = random_ues(rng, n_ue, max_tb_bytes_bg1)
        early_stop, max_iter = bool(rng.integers(0, 2)), int(rng.integers(2, 9))
        n_done += sum(_run_tb_sequence(ctx, orc, pool, batch, ues, [0, 2, 3, 1], max_iter, early_stop, rng, fill=None))
    return n_done


def test_generations_of_ues_on_the_same_entries(ctx, orc):
    rng = np.random.default_rng(4242)
    pool = rx_buffer_pool(ctx, first_entry=512, nof_entries=1024)
    assert tb_generations(ctx, orc, rng, pool, n_gen=4, n_ue=14) > 10


def test_config3_slot_273prb_256qam_4layers(ctx, orc):
    # Config 3: one TB of 1 277 992 bits, 152 codeblocks, Z=384, F=16, E=8960/8992, gNB-style Nref, rv sequence.
    rng = np.random.default_rng(33)
    tb_bytes = 1277992 // 8
    n_llr = 273 * 12 * 13 * 4 * 8
    C = compute_nof_codeblocks(tb_bytes * 8, 1)
    assert C == 152 and n_llr == 1362816
    nref = compute_N_ref(tb_bytes, C)
    assert nref == 12611
    pool = rx_buffer_pool(ctx, first_entry=0, nof_entries=512)
    batch = PuschDecoderBatch(ctx)
    ue = dict(tb=rng.integers(0, 256, tb_bytes).astype(np.uint8), bg=1, qm=8, nl=4, n_llr=n_llr, nref=nref, snr=7.9)
    done = _run_tb_sequence(ctx, orc, pool, batch, [ue], [0, 2, 3, 1], 6, True, rng, fill=3)
    assert done[0]


@pytest.mark.parametrize("bw_prb,table", [(52, "qam64"), (106, "qam64"), (106, "qam256")])
def test_config5_edgeric_multi_ue_slots(ctx, orc, bw_prb, table):
    """BASELINE config 5 (5G/configs/zmq-mode-multi-ue.yml: 10 MHz = 52 PRB, 20 MHz = 106 PRB at 15 kHz, one layer, 2-4
    UEs sharing the carrier): transport block sizes from the MCS tables and TS 38.214 5.1.3.2 (sch.py), base graph by
    TS 38.212 6.2.2, 1-7 codeblocks per UE, BG1 / BG2 mix, all UEs of a slot in ONE batch, HARQ over rv 0-2-3-1, gNB-style
    limited buffer. Every result, CRC flag and HARQ soft bit against the oracle."""
    from srsran_edgeric_5g_b200 import sch
    rng = np.random.default_rng(500 + bw_prb + len(table))
    pool = rx_buffer_pool(ctx, first_entry=1536, nof_entries=512)
    batch = PuschDecoderBatch(ctx)
    seen_bg, seen_cb = set(), set()
    for slot in range(6):
        n_ue = int(rng.integers(2, 5))
        cuts = np.sort(rng.choice(np.arange(1, bw_prb), n_ue - 1, replace=False))
        shares = np.diff(np.concatenate([[0], cuts, [bw_prb]]))
        ues = []
        for n_prb in shares:
            n_mcs = len(sch.MCS_TABLE_QAM64 if table == "qam64" else sch.MCS_TABLE_QAM256)
            a = sch.pusch_allocation(table, int(rng.integers(0, n_mcs)), int(n_prb), nof_dmrs_symbols=int(rng.integers(1, 4)))
            C = compute_nof_codeblocks(a["tbs_bits"], a["base_graph"])
            nref = compute_N_ref(a["tbs_bits"] // 8, C)  # TBS_LBRM of the UE = this transport block
            # per-soft-bit SNR around the decoding threshold of the code rate, so that some UEs need retransmissions
            snr = (8 if a["base_graph"] == 1 else 3) * a["rate"] / 0.8 + rng.uniform(-4, 0)
            ues.append(dict(tb=rng.integers(0, 256, a["tbs_bits"] // 8).astype(np.uint8), bg=a["base_graph"], qm=a["qm"],
                            nl=1, n_llr=a["n_llr"], nref=nref, snr=float(snr)))
            seen_bg.add(a["base_graph"])
            seen_cb.add(C)
        _run_tb_sequence(ctx, orc, pool, batch, ues, [0, 2, 3, 1], 6, True, rng, fill=0)
        for i in range(len(ues)):  # end of the HARQ processes: hand the entries back
            b = pool._buffers.get(("ue", i))
            if b is not None:
                b.release()
    assert seen_bg == {1, 2} or bw_prb == 106
    assert max(seen_cb) >= 2
