"""GPU test of the N > 1 path (-m gpu): two processes (gloo for the plumbing, both on cuda:0 - the data path has no
collective), each decoding ITS cells of a multi-cell slot on the GPU through the C ABI (pdc_submit_codewords +
pdc_submit on host buffers); the per-slot gather (sharding.gather_slot_flags) must return every cell exactly once with
the transport block that was sent, over two transmissions of a HARQ process that stays on its rank."""
import os
import socket
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
N_CELLS = 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cell(orc, cell, rv, snr):
    """Transport block + scrambled soft bits of one cell (BG1, 16QAM, 2 layers, 5 codeblocks)."""
    from tests.vectors import make_tb_llrs
    rng = np.random.default_rng(1000 + cell)
    tb = rng.integers(0, 256, 4000).astype(np.uint8)
    n_llr = 13 * 12 * 50 * 2 * 4
    llrs, C = make_tb_llrs(orc, tb, 1, rv, 4, 0, 2, n_llr, snr, np.random.default_rng(7 * cell + rv))
    c_init = (0x4601 + cell) * 32768 + cell
    return tb, orc.revert_scrambling(llrs, orc.prg_bits(c_init, 0, n_llr)), c_init, n_llr, C


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from srsran_edgeric_5g_b200 import capi, ldpc, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    mine = sharding.cells_of_rank(N_CELLS, world, rank)
    per = max(len(sharding.cells_of_rank(N_CELLS, world, r)) for r in range(world))
    ctx = capi.Context(device=0, max_cbs=64, max_llrs=1 << 20, harq_entries=64, max_tbs=8, max_tb_bytes=1 << 16)
    tables = []
    for t, (rv, snr) in enumerate(((0, 1.0), (2, 4.0))):  # the first transmission is too noisy on purpose
        cells = [_cell(orc, c, rv, snr) for c in mine]
        n_llr, C = cells[0][3], cells[0][4]
        tbs_bits = cells[0][0].size * 8
        tb_stride = (tbs_bits + 24 + 31) // 32 * 4
        cws = np.zeros(len(mine), capi.CW_DESC_DTYPE)
        cbs = np.zeros(len(mine) * C, capi.CB_DESC_DTYPE)
        tbd = np.zeros(len(mine), capi.TB_DESC_DTYPE)
        flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_EARLY_STOP | (capi.CB_NEW_DATA if t == 0 else 0)
        for k, (tb, raw, c_init, _, _) in enumerate(cells):
            cws[k]["in_offset"], cws[k]["sch_offset"], cws[k]["c_init"] = k * n_llr, k * n_llr, c_init
            cws[k]["flags"] = capi.CW_SCRAMBLED | capi.CW_DEFER_DESCRAMBLING
            for key, v in (("qm", 4), ("nof_layers", 2), ("nof_prb", 50), ("nof_symbols", 14), ("dmrs_type", 1),
                           ("dmrs_symbol_mask", 1 << 2), ("nof_cdm_groups_without_data", 2)):
                cws[k][key] = v
            tbd[k] = (k * C, C, tbs_bits, k * tb_stride, 0)
            for i, m in enumerate(ldpc.segment_rx(tbs_bits, 1, rv, 4, 0, 2, n_llr)):
                # HARQ entries are per rank (its own arena): the cell's entries are the same in both transmissions
                cbs[k * C + i] = (k * n_llr + m.cw_offset, m.rm_length, k * C + i, 0, m.lifting_size, m.nof_filler_bits, 1,
                                  4, rv, capi.CRC24B, 6, flags, k)
        ctx.submit_codewords(cws, np.concatenate([c[1] for c in cells]), stream=0)
        ctx.submit(cbs, None, tbd, stream=0, want_bits=False)
        out = ctx.wait(0)
        local = torch.full((per, 3), -1, dtype=torch.int64)
        for k, cell in enumerate(mine):
            local[k, 0], local[k, 1] = cell, int(out["tb_results"][k]["tb_crc_ok"])
            local[k, 2] = zlib.crc32(out["tb_bytes"][k * tb_stride:k * tb_stride + tbs_bits // 8].tobytes())
        tables.append(sharding.gather_slot_flags(local, world).numpy().tolist())
    if rank == 0:
        q.put((tables, ctx.launch_count()))
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_decode_their_cells_on_the_gpu(orc):
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tables, launches = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert launches >= 4  # the rank really ran kernels
    want = {c: zlib.crc32(_cell(orc, c, 0, 1.0)[0].tobytes()) for c in range(N_CELLS)}
    first, second = tables
    for table in (first, second):
        assert sorted(r[0] for r in table if r[0] >= 0) == list(range(N_CELLS))  # every cell exactly once
    assert not all(r[1] == 1 for r in first if r[0] >= 0)  # the noisy first transmission leaves cells undecoded
    for r in second:  # combining with rv 2 on the SAME rank's HARQ entries rescues them
        if r[0] >= 0:
            assert r[1] == 1 and r[2] == want[r[0]], r
