"""Multi-GPU sharding of the PUSCH decode path: one process per GPU, cells/UEs partitioned on the host, no data-path
collective. Codeblocks, transport blocks, UEs and cells are independent; the only requirement is that a
UE's HARQ process always lands on the same GPU, because its soft bits live in that GPU's HARQ arena
(rx_buffer_pool_impl.cpp:44 keys buffers by (rnti, harq_id)). torch.distributed is used only for plumbing: a barrier and
a host-side gather of the small per-slot results (NCCL for GPU runs, gloo for CPU tests)."""
from typing import Dict, Iterable, List, Sequence


def owner_of_cell(cell_id: int, world_size: int) -> int:
    """Static cell -> rank map. Sticky by construction: it depends on nothing but the cell id."""
    return cell_id % world_size


def owner_of_ue(cell_id: int, rnti: int, world_size: int, by_cell: bool = True) -> int:
    """Rank that owns every HARQ process of this UE. by_cell=False hashes the UE instead (cells larger than a GPU)."""
    if by_cell:
        return owner_of_cell(cell_id, world_size)
    return (cell_id * 65537 + rnti) % world_size


def shard_transport_blocks(tbs: Sequence[dict], world_size: int, rank: int, by_cell: bool = True) -> List[int]:
    """Indices of the transport blocks of a slot this rank decodes. Each tb is a dict with 'cell' and 'rnti'."""
    return [i for i, tb in enumerate(tbs) if owner_of_ue(tb["cell"], tb["rnti"], world_size, by_cell) == rank]


def gather_slot_results(local: Dict[int, dict], world_size: int, group=None) -> Dict[int, dict]:
    """Host-side gather of per-TB results {tb index: result} from every rank (a few bytes per TB, once per slot)."""
    if world_size == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, local, group=group)
    out: Dict[int, dict] = {}
    for p in parts:
        for k, v in p.items():
            assert k not in out, "a transport block was decoded by two ranks"
            out[k] = v
    return out


def gather_slot_flags(local, world_size: int, group=None):
    """Per-slot gather on the hot path: every rank contributes one fixed-size tensor (e.g. {cell, tb_crc_ok, checksum} per
    owned cell, padded to the largest share) and gets the concatenation in rank order back. One small collective per slot
    instead of the pickled objects of gather_slot_results; NCCL for GPU tensors, gloo for CPU tensors."""
    if world_size == 1:
        return local.clone()
    import torch
    import torch.distributed as dist
    out = torch.empty((world_size,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out.view((world_size * local.shape[0],) + tuple(local.shape[1:]))


def cells_of_rank(n_cells: int, world_size: int, rank: int) -> List[int]:
    """Cells this rank owns (owner_of_cell), in increasing order."""
    return [c for c in range(n_cells) if owner_of_cell(c, world_size) == rank]


def check_partition(tbs: Sequence[dict], world_size: int, by_cell: bool = True) -> None:
    """Every transport block has exactly one owner and all HARQ processes of a UE share it."""
    owners = {}
    for tb in tbs:
        o = owner_of_ue(tb["cell"], tb["rnti"], world_size, by_cell)
        key = (tb["cell"], tb["rnti"])
        assert owners.setdefault(key, o) == o
        assert 0 <= o < world_size
