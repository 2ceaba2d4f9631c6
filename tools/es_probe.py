#!/usr/bin/env python3
"""Measurement aid: cost of one decoder iteration and of one early-stop check on a config-3 slot (152 codeblocks, 4 rows
each), device resident. Fixed iteration counts 1..6 give the iteration cost; early stop on noise-only input (never
stops) adds one check per iteration."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from srsran_edgeric_5g_b200 import capi  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_cb, E = 152 * cells, 8960
ctx = capi.Context(device=0, max_cbs=n_cb, max_llrs=n_cb * E, harq_entries=n_cb, max_tbs=1, max_tb_bytes=4096)
rng = np.random.default_rng(0)
llr = rng.integers(-30, 31, n_cb * E).astype(np.int8)  # noise: the CRC never passes
d_llr = torch.from_numpy(llr).cuda()
d_res = torch.zeros(n_cb * 4, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()


def run(max_iter, early, dematch=True, decode=True):
    cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
    flags = (capi.CB_DEMATCH if dematch else 0) | (capi.CB_DECODE if decode else 0) | capi.CB_NEW_DATA | \
        (capi.CB_EARLY_STOP if early else 0)
    for k in range(n_cb):
        cbs[k] = (k * E, E, k, 12611, 384, 16, 1, 8, 0, capi.CRC24B, max_iter, flags, 0xffff)
    d_cbs = torch.from_numpy(cbs.view(np.uint8)).cuda()

    def step():
        ctx.launch_device(d_cbs.data_ptr(), n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384, flags, True,
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
    return e0.elapsed_time(e1) * 50


t_dm = run(1, False, True, False)
print("dematch only: %.1f us" % t_dm)
for it in (1, 2, 3, 6):
    print("fixed %d iterations: %.1f us   early-stop checks every iteration (never passes): %.1f us" %
          (it, run(it, False), run(it, True)))
ctx.close()
