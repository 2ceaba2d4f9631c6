#!/usr/bin/env python3
"""Profiling aid: the codeword front end of one 16-cell slot (BASELINE config 4) on device buffers, a few launches.

    ncu --set full -k regex:"prg_kernel|ulsch_sch_kernel" ... python tools/frontend_profile.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from srsran_edgeric_5g_b200 import capi  # noqa: E402

cells, n_llr = 16, 1362816
ctx = capi.Context(device=0, max_cbs=64, harq_entries=64, max_tbs=1, max_tb_bytes=4096)
cws = np.zeros(cells, capi.CW_DESC_DTYPE)
for c in range(cells):
    cws[c]["in_offset"], cws[c]["sch_offset"], cws[c]["c_init"] = c * n_llr, c * n_llr, (0x4601 + c) * 32768 + 17 * c
    cws[c]["flags"] = capi.CW_SCRAMBLED
    for k, v in (("qm", 8), ("nof_layers", 4), ("nof_prb", 273), ("nof_symbols", 14), ("dmrs_type", 1),
                 ("dmrs_symbol_mask", 1 << 2), ("nof_cdm_groups_without_data", 2)):
        cws[c][k] = v
rng = np.random.default_rng(0)
d_raw = torch.from_numpy(rng.integers(-120, 121, cells * n_llr).astype(np.int8)).cuda()
d_sch = torch.zeros(cells * n_llr + 16, dtype=torch.int8, device="cuda")
stream = torch.cuda.current_stream()
for _ in range(4):
    ctx.launch_codewords_device(cws, d_raw.data_ptr(), cells * n_llr, d_sch.data_ptr(), cells * n_llr,
                                cuda_stream=stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10):
    ctx.launch_codewords_device(cws, d_raw.data_ptr(), cells * n_llr, d_sch.data_ptr(), cells * n_llr,
                                cuda_stream=stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
print("front end, 16 codewords of %d soft bits: %.1f us per slot" % (n_llr, e0.elapsed_time(e1) * 100))
ctx.close()
