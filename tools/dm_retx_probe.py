#!/usr/bin/env python3
"""Rate dematcher on retransmissions: time and HBM fraction of the combine path (rv 2 / 3 / 1 combined into entries that
hold a first transmission; 8192 codeblocks BG1 Z = 384, E = 25344 and E = 8960), next to the new-data path."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from srsran_edgeric_5g_b200 import capi
    n_cb = 8192
    ctx = capi.Context(device=0, max_cbs=n_cb, max_llrs=n_cb * 25344, harq_entries=n_cb)
    stream = torch.cuda.current_stream()
    peaks = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text()).get("hbm_gbs", 6650.0) if peaks.exists() else 6650.0
    rng = np.random.default_rng(1)
    for E, qm, nref in ((25344, 2, 0), (8960, 8, 12611), (16896, 6, 0)):
        llr = rng.integers(-60, 60, n_cb * E).astype(np.int8)
        d_llr = torch.from_numpy(llr).cuda()
        d_res = torch.zeros(n_cb * 4, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
        for label, rv, new in (("new data rv0", 0, True), ("combine rv0", 0, False), ("combine rv2", 2, False),
                               ("combine rv3", 3, False), ("combine rv1", 1, False)):
            cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
            fl = capi.CB_DEMATCH | (capi.CB_NEW_DATA if new else 0)
            for i in range(n_cb):
                cbs[i] = (i * E, E, i, nref, 384, 0, 1, qm, rv, capi.CRC24B, 6, fl, 0xffff)
            d_cbs = torch.from_numpy(cbs.view(np.uint8)).cuda()

            def step():
                ctx.launch_device(d_cbs.data_ptr(), n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384, fl, True,
                                  cuda_stream=stream.cuda_stream)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            ncb = nref if nref else 25344
            # algorithmic bytes: E soft bits in; new data: Ncb written (copy + zero fill); combine: min(E, Ncb) read + written
            touched = min(E, ncb)
            byts = E + (ncb if new else 2 * touched)
            print(f"E={E} qm={qm} Ncb={ncb}  {label:14s} {ms * 1e3:7.1f} us  {n_cb * byts / ms / 1e6:7.0f} GB/s = "
                  f"{n_cb * byts / ms / 1e6 / peak:.2f} of HBM", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
