"""CPU test of the N>1 path: two processes (gloo), each decodes its shard of a slot's transport blocks with the oracle
standing in for the GPU, results are gathered on the host; the union equals the single-process result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slot():
    rng = np.random.default_rng(5)
    tbs = []
    for cell in range(4):
        for ue in range(2):
            tbs.append(dict(cell=cell, rnti=0x4601 + ue, tb=rng.integers(0, 256, int(rng.integers(20, 200))).astype(np.uint8)))
    return tbs


def _decode(tb):
    sys.path.insert(0, str(ROOT))
    from oracle import pyoracle as po
    from tests.vectors import awgn_llr
    orc = po.Oracle()
    rng = np.random.default_rng(int(tb["cell"]) * 100 + int(tb["rnti"]))
    n_llr = int(np.ceil(tb["tb"].size * 8 / 0.5 / 2)) * 2
    cw, C = orc.tb_encode(tb["tb"], 2, 0, 2, 0, 1, n_llr)
    out, st = orc.pusch_decode(po.Harq(C), awgn_llr(cw, 6.0, rng), tb["tb"].size, 2, 0, 2, 0, 1, 6, True, True)
    return dict(ok=bool(st[0]), tb=out.tobytes())


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from srsran_edgeric_5g_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tbs = _slot()
    mine = sharding.shard_transport_blocks(tbs, world, rank)
    local = {i: _decode(tbs[i]) for i in mine}
    dist.barrier()
    allr = sharding.gather_slot_results(local, world)
    # the fixed-size per-slot gather of the hot path (one row per owned transport block, padded)
    import torch
    rows = torch.full((len(tbs), 2), -1, dtype=torch.int64)
    for k, i in enumerate(mine):
        rows[k, 0], rows[k, 1] = i, int(local[i]["ok"])
    table = sharding.gather_slot_flags(rows, world)
    assert sorted(int(r[0]) for r in table if r[0] >= 0) == list(range(len(tbs)))
    assert all(int(r[1]) == 1 for r in table if r[0] >= 0)
    assert sharding.cells_of_rank(4, world, rank) == [c for c in range(4) if c % world == rank]
    if rank == 0:
        q.put((sorted(allr), [allr[i]["ok"] for i in sorted(allr)], [allr[i]["tb"] for i in sorted(allr)], len(mine)))
    dist.destroy_process_group()


def test_two_rank_sharded_slot():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys, oks, outs, n_mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tbs = _slot()
    assert keys == list(range(len(tbs))) and n_mine == len(tbs) // 2
    single = [_decode(tb) for tb in tbs]
    assert oks == [s["ok"] for s in single] and all(oks)
    assert outs == [s["tb"] for s in single] == [tb["tb"].tobytes() for tb in tbs]
