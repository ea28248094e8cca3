"""Host-side plumbing for the sharded (one process per GPU) mode, SURVEY.md §8(e).

torch.distributed is used only to move harness data between ranks (graph slices,
the NCCL unique id, per-shard result lists); the engine's own per-sweep exchange
runs inside libspaghetti_gpu over NCCL.  Works with the `nccl` backend on GPUs
and with `gloo` on CPU (tests)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def doc_shard(rank: int, world: int, n_docs: int) -> Tuple[int, int]:
    """Contiguous doc range of a rank (document-sharded index)."""
    return n_docs * rank // world, n_docs * (rank + 1) // world


def row_slice(rank: int, world: int, n_nodes: int) -> Tuple[int, int]:
    """Rows of the out-edge CSR a rank generates/exports before the exchange."""
    return n_nodes * rank // world, n_nodes * (rank + 1) // world


def assemble_graph(n_nodes: int, part_row_ptr: np.ndarray, part_col_idx: np.ndarray, device="cpu"):
    """Every rank holds rows row_slice(rank) of the CSR (row_ptr local to the slice);
    returns the full (row_ptr uint64 [N+1], col_idx uint32 [E]) on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = row_slice(rank, world, n_nodes)
    assert len(part_row_ptr) == hi - lo + 1
    deg = torch.from_numpy(np.diff(part_row_ptr.astype(np.int64))).to(device)
    cnt = torch.tensor([int(part_row_ptr[-1])], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    degs = []
    for r in range(world):
        a, b = row_slice(r, world, n_nodes)
        degs.append(torch.zeros(b - a, dtype=torch.int64, device=device))
    # all_gather needs equal shapes on gloo; pad to the largest slice
    pad = max(max(d.numel() for d in degs), 1)
    pdeg = torch.zeros(pad, dtype=torch.int64, device=device)
    pdeg[: deg.numel()] = deg
    gd = [torch.zeros(pad, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(gd, pdeg)
    padc = max(max(cnts), 1)
    pcol = torch.zeros(padc, dtype=torch.int32, device=device)
    pcol[: cnts[rank]] = torch.from_numpy(part_col_idx.view(np.int32)).to(device)
    gc = [torch.zeros(padc, dtype=torch.int32, device=device) for _ in range(world)]
    dist.all_gather(gc, pcol)
    all_deg = torch.cat([gd[r][: degs[r].numel()] for r in range(world)])
    row_ptr = np.zeros(n_nodes + 1, dtype=np.uint64)
    row_ptr[1:] = torch.cumsum(all_deg, 0).cpu().numpy().astype(np.uint64)
    col_idx = torch.cat([gc[r][: cnts[r]] for r in range(world)]).cpu().numpy().view(np.uint32)
    return row_ptr, np.ascontiguousarray(col_idx)


def share_unique_id(make_id):
    """Rank 0 creates the NCCL unique id (ss_comm_unique_id); everyone gets the bytes."""
    import torch.distributed as dist
    obj = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def engine_grid(rank: int, row_groups: int, topic_groups: int, n_topics: int):
    """Place of a rank in the row x topic grid of PageRank engines: ranks of one topic group are
    consecutive.  -> (rank inside its row group, topic group, first topic, number of topics)."""
    assert n_topics % topic_groups == 0
    per = n_topics // topic_groups
    return rank % row_groups, rank // row_groups, (rank // row_groups) * per, per


def share_group_unique_id(make_id, group_rank: int, group_size: int):
    """One NCCL unique id per group of `group_size` consecutive ranks: the first rank of every group creates
    one (ss_comm_unique_id), everybody receives the id of its own group.  Collective over the whole world."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = make_id() if group_rank == 0 else None
    ids = [None] * world
    dist.all_gather_object(ids, uid)
    return ids[rank - group_rank]


def gather_result_lists(docs: np.ndarray, finals: np.ndarray, prs: np.ndarray, counts: np.ndarray, device="cpu"):
    """All ranks' [Q][k] lists -> [world][Q][k] on every rank (input of ss_merge_topk)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    out = []
    for a, dt in ((docs.view(np.int32), torch.int32), (finals, torch.float64), (prs, torch.float64),
                  (counts.view(np.int32), torch.int32)):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        out.append(torch.stack(g).cpu().numpy())
    return out[0].view(np.uint32), out[1], out[2], out[3].view(np.uint32)
