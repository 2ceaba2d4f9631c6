#!/usr/bin/env python3
"""Profiling aid: the soft demapper of one 16-cell slot (BASELINE config 4: 16 x 170 352 256QAM symbols, one
demodulate_soft call per OFDM symbol) on device buffers, a few launches.

    ncu --set full -k regex:demod_kernel ... python tools/demod_profile.py [qm]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from srsran_edgeric_5g_b200 import capi  # noqa: E402

qm = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cells, n_sym, per_call = 16, 170352, 13104
ctx = capi.Context(device=0, max_cbs=64, harq_entries=64, max_tbs=1, max_tb_bytes=4096)
rng = np.random.default_rng(0)
sym = ((rng.standard_normal(cells * n_sym) + 1j * rng.standard_normal(cells * n_sym)) * 0.7).astype(np.complex64)
nv = (0.01 + 0.01 * rng.random(cells * n_sym)).astype(np.float32)
calls = np.array([(c * n_sym + k * per_call, per_call, (c * n_sym + k * per_call) * qm, qm)
                  for c in range(cells) for k in range(n_sym // per_call)], capi.DEMOD_CALL_DTYPE)
d_sym = torch.from_numpy(sym.view(np.float32)).cuda()
d_nv = torch.from_numpy(nv).cuda()
d_llr = torch.zeros(cells * n_sym * qm + 16, dtype=torch.int8, device="cuda")
stream = torch.cuda.current_stream()


def run():
    ctx.launch_demod_device(calls, d_sym.data_ptr(), d_nv.data_ptr(), sym.size, d_llr.data_ptr(), cells * n_sym * qm,
                            cuda_stream=stream.cuda_stream)


for _ in range(4):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20):
    run()
e1.record(stream)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 50
print("soft demapper, %d symbols (Qm = %d) in %d calls: %.1f us per slot = %.0f GB/s" %
      (sym.size, qm, calls.size, us, sym.size * (12 + qm) / us / 1e3))
ctx.close()
