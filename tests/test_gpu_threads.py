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
