"""Multi-GPU plumbing: one process per GPU, batches sharded by envmap ownership, one all-gather of the rendered
refmaps.  The reference does the same implicitly (each Lightning DDP rank renders its own sampler shard inside
get_input, main.py:554 / models/drmnet.py:561-569); there is no exchange step inside either kernel, so no collective
is needed on the data path -- the all-gather exists for callers that want every refmap on one rank (sampling,
benchmark).  24 MB envmaps never cross NVLink: rank r loads the maps it owns."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def owned_envmaps(num_envmaps: int, world_size: int, rank: int) -> List[int]:
    """Envmap b belongs to rank b % world_size."""
    return list(range(rank, num_envmaps, world_size))


def shard_renders(env_index: torch.Tensor, world_size: int, rank: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Renders follow their envmap.  Returns (ids of the renders this rank executes, their local envmap slots)."""
    env_index = env_index.to(torch.int64)
    mine = torch.nonzero(env_index % world_size == rank).flatten()
    return mine, env_index[mine] // world_size


def all_gather_refmaps(local: torch.Tensor, render_ids: torch.Tensor, total: int) -> torch.Tensor:
    """Gather [n_r, ...] blocks of every rank into [total, ...] ordered by global render id.

    Ranks may own different counts: blocks are padded to the maximum so a single equal-count all_gather is used
    (NCCL all-gather over NVLink on GPUs, gloo on CPU in tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = local.new_zeros((total,) + tuple(local.shape[1:]))
        out[render_ids.to(local.device)] = local
        return out
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c) for c in counts))
    pad = local.new_zeros((nmax,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    ids = torch.full((nmax,), -1, dtype=torch.int64, device=local.device)
    ids[: local.shape[0]] = render_ids.to(local.device)
    blocks = [torch.empty_like(pad) for _ in range(world)]
    id_blocks = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(blocks, pad)
    dist.all_gather(id_blocks, ids)
    out = local.new_zeros((total,) + tuple(local.shape[1:]))
    for blk, idb in zip(blocks, id_blocks):
        keep = idb >= 0
        out[idb[keep]] = blk[keep]
    return out
