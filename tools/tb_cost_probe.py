#!/usr/bin/env python3
"""What the TB assembly kernel costs in the slot chain: the config-3 / config-5-like slots with and without transport
blocks in the launch (resident, CUDA events)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from oracle.pyoracle import Oracle
    from srsran_edgeric_5g_b200 import capi, ldpc
    from tests.vectors import make_tb_llrs
    orc = Oracle()
    ctx = capi.Context(device=0, max_cbs=2432, max_llrs=1 << 20, harq_entries=2432, max_tbs=16, max_tb_bytes=16 * 160000)
    stream = torch.cuda.current_stream()
    rng = np.random.default_rng(3)
    for label, tbs_bits, n_llr, qm, nl, cells in (("config3", 1277992, 1362816, 8, 4, 1), ("config4", 1277992, 1362816, 8, 4, 16),
                                                  ("small", 30000 * 8, 399996, 6, 1, 4)):
        C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
        nref = ldpc.compute_N_ref(tbs_bits // 8, C)
        tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
        llrs, _ = make_tb_llrs(orc, tb, 1, 0, qm, nref, nl, n_llr, 8.4, rng)
        metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_llr)
        n_cb = C * cells
        cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
        tbd = np.zeros(cells, capi.TB_DESC_DTYPE)
        flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA
        tb_stride = (tbs_bits + 24 + 31) // 32 * 4
        for c in range(cells):
            tbd[c] = (c * C, C, tbs_bits, c * tb_stride, 0)
            for k, m in enumerate(metas):
                cbs[c * C + k] = (c * n_llr + m.cw_offset, m.rm_length, c * C + k, nref, m.lifting_size, m.nof_filler_bits, 1,
                                  qm, 0, capi.CRC24B, 6, flags, c)
        d_cbs = torch.from_numpy(cbs.view(np.uint8)).cuda()
        d_tbs = torch.from_numpy(tbd.view(np.uint8)).cuda()
        d_llr = torch.from_numpy(np.tile(llrs, cells)).cuda()
        d_res = torch.zeros(n_cb * 4, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
        d_tres = torch.zeros(cells * 4, dtype=torch.uint8, device="cuda")
        d_tb = torch.zeros(cells * tb_stride + 16, dtype=torch.uint8, device="cuda")
        out = {}
        for with_tb in (True, False):
            def step():
                ctx.launch_device(d_cbs.data_ptr(), n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384, flags,
                                  True, cuda_stream=stream.cuda_stream, d_tbs=d_tbs.data_ptr() if with_tb else 0,
                                  n_tb=cells if with_tb else 0, d_tb_results=d_tres.data_ptr() if with_tb else 0,
                                  d_tb_bytes=d_tb.data_ptr() if with_tb else 0)
            for _ in range(5):
                step()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(30):
                    step()
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 30 * 1e3)
            out["with TB assembly" if with_tb else "without"] = round(best, 1)
        print(label, n_cb, "codeblocks:", out)
    ctx.close()


if __name__ == "__main__":
    main()
