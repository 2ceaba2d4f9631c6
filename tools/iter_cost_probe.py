#!/usr/bin/env python3
"""Fixed cost and per-iteration cost of the decode kernel on a config-3 slot (152 codeblocks, four rows in use) and on a
low-rate slot (all 46 rows): chain time without TB assembly at max_iter = 1, 2, 4, 6, 8 (resident, CUDA events)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from oracle.pyoracle import Oracle
    from srsran_edgeric_5g_b200 import capi, ldpc
    from tests.vectors import make_tb_llrs, make_cb_batch
    orc = Oracle()
    ctx = capi.Context(device=0, max_cbs=2432, max_llrs=1 << 23, harq_entries=2432, max_tbs=16, max_tb_bytes=16 * 160000)
    stream = torch.cuda.current_stream()
    rng = np.random.default_rng(3)
    tbs_bits, n_llr, qm, nl = 1277992, 1362816, 8, 4
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
    llrs, _ = make_tb_llrs(orc, tb, 1, 0, qm, nref, nl, n_llr, 8.4, rng)
    metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_llr)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA
    cases = []
    for n_use in (152, 148, 76, 8, 2):
        cbs = np.zeros(n_use, capi.CB_DESC_DTYPE)
        for k, m in enumerate(metas[:n_use]):
            cbs[k] = (m.cw_offset, m.rm_length, k, nref, m.lifting_size, m.nof_filler_bits, 1, qm, 0, capi.CRC24B, 6, flags, 0xffff)
        cases.append(("high rate (4 rows), %d codeblocks" % n_use, cbs, llrs))
    # low rate: E = 25344 (rate 1/3): all 46 rows
    for n_use in (148, 8):
        b = make_cb_batch(orc, bg=1, Z=384, n_cb=n_use, E=25344, qm=2, rv=0, snr_db=-1.0, seed=5, crc_kind=capi.CRC24B)
        cases.append(("rate 1/3 (46 rows), %d codeblocks" % n_use, b.descriptors(capi, 6, False), b.llrs.reshape(-1)))
    for label, cbs, ll in cases:
        n_cb = cbs.size
        d_llr = torch.from_numpy(np.ascontiguousarray(ll)).cuda()
        d_res = torch.zeros(n_cb * 4, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(n_cb * capi.PDC_MAX_CB_BYTES, dtype=torch.uint8, device="cuda")
        out = {}
        for it in (1, 2, 4, 6, 8):
            c2 = cbs.copy()
            c2["max_iter"] = it
            d_cbs = torch.from_numpy(c2.view(np.uint8)).cuda()

            def step():
                ctx.launch_device(d_cbs.data_ptr(), n_cb, d_llr.data_ptr(), d_res.data_ptr(), d_bits.data_ptr(), 384, flags,
                                  True, cuda_stream=stream.cuda_stream)
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
            out[it] = round(best, 1)
        print(label, "us at max_iter", out, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
