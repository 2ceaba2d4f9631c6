"""Host-side robustness of the C ABI (-m gpu): queues driven from different threads at the same time, synchronous
single-codeblock calls from many threads while batches are in flight, and two contexts on two devices in one process."""
import threading

import numpy as np
import pytest

from oracle import pyoracle as po
from srsran_edgeric_5g_b200 import capi
from tests.vectors import make_cb_batch

pytestmark = pytest.mark.gpu


def _same(out, ref):
    return bool((out["crc_ok"] == ref["crc_ok"]).all() and (out["iters"] == ref["iters"]).all() and
                (out["bits"] == ref["bits"]).all())


def _run_on_queue(c, b, q, harq_base, rounds, errors):
    try:
        cbs = b.descriptors(capi, 6, True, True, harq_base)
        kb = (b.K + 7) // 8
        for _ in range(rounds):
            c.submit(cbs, np.ascontiguousarray(b.llrs.reshape(-1)), None, stream=q, want_bits=True)
            out = c.wait(q)
            got = {"crc_ok": out["cb_results"]["crc_ok"].astype(bool), "iters": out["cb_results"]["iters"].astype(np.int32),
                   "bits": out["cb_bits"][:, :kb].copy()}
            if not _same(got, b.ref):
                errors.append(("queue", q))
    except Exception as e:  # noqa: BLE001 - reported by the test
        errors.append(("queue", q, repr(e)))


def test_two_threads_two_queues_and_sync_calls(orc):
    """Thread A drives queue 0, thread B queue 1 (first use of both at the same time: the decoder's per-stream scratch
    must not be shared or reallocated under them), while four more threads hammer the synchronous single-codeblock
    calls; every result is compared with the oracle."""
    c = capi.Context(device=0, max_cbs=64, max_llrs=64 * 25344, harq_entries=256, max_tbs=1, max_tb_bytes=4096,
                     nof_streams=2)
    batches = []
    for q, (bg, Z, E, snr) in enumerate(((1, 384, 66 * 384, 0.2), (2, 96, 40 * 96, 1.5))):
        b = make_cb_batch(orc, bg, Z, n_cb=24, E=E, qm=2, rv=0, snr_db=snr, seed=50 + q)
        b.ref = b.run_oracle(orc, 6, True)
        batches.append(b)
    small = make_cb_batch(orc, 2, 52, n_cb=8, E=50 * 52, qm=2, rv=0, snr_db=3.0, seed=9, crc_kind=po.CRC16)
    small_ref = [orc.ldpc_decode(2, 52, small.run_oracle(orc, 6, True)["harq"][i][:50 * 52], 0, po.CRC16, 6)
                 for i in range(small.n_cb)]
    harq = small.run_oracle(orc, 6, True)["harq"]
    errors = []

    def sync_calls(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(40):
                i = int(rng.integers(0, small.n_cb))
                it, bits = c.ldpc_decode(2, 52, harq[i][:50 * 52], 0, po.CRC16, 6)
                if it != small_ref[i][0] or not (bits == small_ref[i][1]).all():
                    errors.append(("sync decode", i))
                d = rng.integers(0, 256, 100).astype(np.uint8)
                if c.crc(po.CRC24A, d, 800) != orc.crc(po.CRC24A, d, 800):
                    errors.append(("sync crc", i))
        except Exception as e:  # noqa: BLE001
            errors.append(("sync", repr(e)))

    threads = [threading.Thread(target=_run_on_queue, args=(c, batches[q], q, 100 * q, 6, errors)) for q in range(2)]
    threads += [threading.Thread(target=sync_calls, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    c.close()
    assert not errors, errors[:5]


def test_submit_rejects_inconsistent_transport_blocks(ctx):
    """Descriptors the assembly kernel would index out of range with are refused on the host (PDC_ERR_INVALID)."""
    cbs = np.zeros(2, capi.CB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA
    for i in range(2):
        cbs[i] = (i * 1000, 1000, i, 0, 96, 0, 1, 2, 0, capi.CRC24B, 2, flags, 0)
    llrs = np.zeros(2000, np.int8)
    good = np.array([(0, 2, 2 * (22 * 96 - 24) - 24, 0, 0)], capi.TB_DESC_DTYPE)
    ctx.submit(cbs, llrs, good, stream=0, want_bits=False)
    ctx.wait(0)
    for label, mutate in (("payload larger than the codeblocks", lambda c, t: t.__setitem__(0, (0, 2, 2 * 22 * 96, 0, 0))),
                          ("codeblocks of different lifting size", lambda c, t: c.__setitem__(1, (1000, 1000, 1, 0, 88, 0, 1, 2, 0, capi.CRC24B, 2, flags, 0))),
                          ("filler bits beyond the message", lambda c, t: [c.__setitem__(i, (i * 1000, 1000, i, 0, 96, 22 * 96 - 10, 1, 2, 0, capi.CRC24B, 2, flags, 0)) for i in range(2)]),
                          ("transport block index outside the batch", lambda c, t: c.__setitem__(0, (0, 1000, 0, 0, 96, 0, 1, 2, 0, capi.CRC24B, 2, flags, 5)))):
        c2, t2 = cbs.copy(), good.copy()
        mutate(c2, t2)
        with pytest.raises(capi.PdcError):
            ctx.submit(c2, llrs, t2, stream=0, want_bits=False)
        assert ctx.poll(0), label  # nothing was queued


def test_two_contexts_on_two_devices(orc):
    """One context per GPU in one process: kernel attributes and constant tables are per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    b = make_cb_batch(orc, 1, 384, n_cb=8, E=66 * 384, qm=2, rv=0, snr_db=0.3, seed=4)
    ref = b.run_oracle(orc, 6, True)
    ctxs = [capi.Context(device=d, max_cbs=64, harq_entries=64, max_tbs=1, max_tb_bytes=4096) for d in (0, 1)]
    for c in ctxs + ctxs[::-1]:
        out = b.run_gpu(c, 6, True)
        assert _same(out, ref)
        assert (out["harq"] == ref["harq"]).all()
    for c in ctxs:
        c.close()


def test_pipelined_submit_equals_single_piece(orc):
    """A large pdc_submit batch is pipelined inside the library (soft bits in groups on a copy stream, kernels per group on
    lane streams, results as they exist). Same batch through a context with PDC_NO_PIPELINE=1: every output byte equal.
    Covers groups of whole transport blocks, one large block cut into groups (assembled after the join), codeblocks
    outside any transport block, and pageable as well as page-locked output buffers."""
    import os
    from srsran_edgeric_5g_b200 import ldpc
    rng = np.random.default_rng(77)

    def make(layout):
        """layout: list of (tbs_bytes or None for loose codeblocks, n_llr)."""
        cbs, tbd, llrs = [], [], []
        off = tb_off = 0
        flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
        for tb_bytes, n_llr in layout:
            if tb_bytes is None:
                for _ in range(3):
                    cbs.append((off, 2000, len(cbs), 0, 96, 7, 2, 2, 0, capi.CRC16, 3, flags, 0xffff))
                    off += 2000
                    llrs.append(rng.integers(-40, 41, 2000).astype(np.int8))
                continue
            metas = ldpc.segment_rx(tb_bytes * 8, 1, 0, 4, 0, 2, n_llr)
            tbd.append((len(cbs), len(metas), tb_bytes * 8, tb_off, 0))
            tb_off += (tb_bytes * 8 + 24 + 31) // 32 * 4
            kind = capi.CRC24B if len(metas) > 1 else capi.CRC24A
            for m in metas:
                cbs.append((off + m.cw_offset, m.rm_length, len(cbs), 0, m.lifting_size, m.nof_filler_bits, 1, 4, 0, kind,
                            3, flags, len(tbd) - 1))
            llrs.append(rng.integers(-60, 61, n_llr).astype(np.int8))
            off += n_llr
        return (np.array(cbs, capi.CB_DESC_DTYPE), np.array(tbd, capi.TB_DESC_DTYPE) if tbd else None,
                np.concatenate(llrs))

    layouts = {
        "several transport blocks": [(30000, 400000), (12000, 200000), (None, 0), (52000, 700000), (30000, 400000),
                                     (4000, 60000), (40000, 500000)],
        "one large transport block": [(150000, 1600000)],
        "codeblocks only": [(None, 0)] * 200,
    }
    outs = {}
    for mode in ("pipelined", "single"):
        os.environ["PDC_NO_PIPELINE"] = "1" if mode == "single" else "0"
        c = capi.Context(device=0, max_cbs=1024, max_llrs=4 << 20, harq_entries=1024, max_tbs=16, max_tb_bytes=1 << 20,
                         nof_streams=1)
        os.environ.pop("PDC_NO_PIPELINE")
        rng = np.random.default_rng(77)
        for name, layout in layouts.items():
            cbs, tbd, llrs = make(layout)
            for pinned in (False, True):
                src = llrs
                if pinned:
                    buf = capi.PinnedBuffer(llrs.size)
                    buf.array[:] = llrs
                    src = buf.array
                c.submit(cbs, src, tbd, stream=0, want_bits=True)
                o = c.wait(0)
                harq = np.stack([c.harq_read(i, 25344) for i in range(0, cbs.size, max(1, cbs.size // 16))])
                outs[(mode, name, pinned)] = (o["cb_results"].copy(), o["cb_bits"].copy(),
                                              None if tbd is None else o["tb_results"].copy(),
                                              None if tbd is None else o["tb_bytes"].copy(), harq)
        c.close()
    for name in layouts:
        for pinned in (False, True):
            a, b = outs[("pipelined", name, pinned)], outs[("single", name, pinned)]
            for x, y, what in zip(a, b, ("codeblock results", "codeblock bits", "TB results", "TB bytes", "HARQ")):
                assert (x is None and y is None) or (x.tobytes() == y.tobytes()), (name, pinned, what)
