#!/usr/bin/env python3
"""Where a lone decode CTA spends its time: runs small batches against a library built with -DH2_PHASE_TIMING
(tools/phase_probe.sh builds it) and prints the global-timer differences between the phase boundaries of CTA 0."""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

NAMES = ["entry", "descriptors", "wait for dematcher", "tables + weights", "last non-zero + layers", "soft bits in",
         "state zeroed", "iteration 0", "iterations 1..n-1", "check + publish", "exit"]


def main():
    import torch
    from oracle.pyoracle import Oracle
    from srsran_edgeric_5g_b200 import capi
    from tests.vectors import make_cb_batch
    orc = Oracle()
    ctx = capi.Context(device=0, max_cbs=512, max_llrs=1 << 23, harq_entries=512)
    lib = ctypes.CDLL(os.environ["PDC_LIBRARY"])
    lib.pdc_debug_read_phases.argtypes = [ctypes.c_void_p]
    stream = torch.cuda.current_stream()
    for label, E, n_cb, iters in (("4 rows", 8960, 8, 6), ("4 rows", 8960, 152, 6), ("46 rows", 25344, 8, 6)):
        b = make_cb_batch(orc, bg=1, Z=384, n_cb=n_cb, E=E, qm=8 if E == 8960 else 2, rv=0, snr_db=8.4 if E == 8960 else -1.0,
                          seed=5, crc_kind=capi.CRC24B)
        cbs = b.descriptors(capi, iters, False)
        flags = int(cbs["flags"][0])
        d_cbs = torch.from_numpy(cbs.view(np.uint8)).cuda()
        d_llr = torch.from_numpy(np.ascontiguousarray(b.llrs.reshape(-1))).cuda()
        d_res = torch.zeros(n_cb * 4, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
        acc = np.zeros(11)
        reps = 20
        for r in range(reps + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.launch_device(d_cbs.data_ptr(), n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384, flags, True,
                              cuda_stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ph = (ctypes.c_ulonglong * 16)()
            lib.pdc_debug_read_phases(ph)
            t = np.array([ph[i] for i in range(11)], dtype=np.float64)
            if r >= 5:
                acc[1:] += np.diff(t) / 1e3
                acc[0] += e0.elapsed_time(e1) * 1e3
        acc /= reps
        print(f"{label}, {n_cb} codeblocks, {iters} iterations: chain {acc[0]:.1f} us (events); decode CTA 0: "
              f"{sum(acc[1:]):.1f} us")
        for k in range(1, 11):
            print(f"    {NAMES[k]:28s} {acc[k]:7.2f} us")
    ctx.close()


if __name__ == "__main__":
    main()
