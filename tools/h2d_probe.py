#!/usr/bin/env python3
"""Host-to-device copy ceiling of a multi-GPU box, rank by rank and all ranks at once (torchrun, one rank per GPU):
page-locked vs write-combined staging, with and without pinning the rank to the CPUs next to its GPU. The e2e figure
of bench.py --gpus N has to be read against the concurrent number printed here (bench.py prints it too)."""
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    from srsran_edgeric_5g_b200 import capi
    import bench
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.load()
    n = 208 << 20
    out = {}
    for pin in (False, True):
        aff = bench.pin_to_gpu_numa_node(torch, local) if pin else None
        for kind in ("pinned", "write_combined"):
            buf = capi.PinnedBuffer(n, np.uint8, input_only=(kind == "write_combined"))
            buf.array[:] = 1
            src = torch.from_numpy(buf.array)
            dst = torch.empty(n, dtype=torch.uint8, device="cuda")
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            for mode in ("alone", "concurrent"):
                if mode == "alone":
                    # one rank at a time
                    ms = 0.0
                    for r in range(world):
                        if world > 1:
                            dist.barrier()
                        if r == rank:
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            for _ in range(6):
                                dst.copy_(src, non_blocking=True)
                            e1.record()
                            torch.cuda.synchronize()
                            ms = e0.elapsed_time(e1)
                else:
                    if world > 1:
                        dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(6):
                        dst.copy_(src, non_blocking=True)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                t = torch.tensor([6 * n / (ms * 1e-3) / 1e9], device="cuda")
                if world > 1:
                    g = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(g, t)
                    vals = [float(x.item()) for x in g]
                else:
                    vals = [float(t.item())]
                out["%s/%s/%s" % ("numa" if pin else "free", kind, mode)] = {
                    "per_rank_gbs": [round(v, 1) for v in vals], "sum_gbs": round(sum(vals), 1), "affinity": aff}
            del src, dst, buf
    if rank == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
