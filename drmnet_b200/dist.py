"""Multi-GPU plumbing: one process per GPU, batches sharded by envmap ownership, one all-gather of the rendered
refmaps.  The reference does the same implicitly (each Lightning DDP rank renders its own sampler shard inside
get_input, main.py:554 / models/drmnet.py:561-569); there is no exchange step inside either kernel, so no collective
is needed on the data path -- the all-gather exists for callers that want every refmap on one rank (sampling,
benchmark).  24 MB envmaps never cross NVLink: rank r loads the maps it owns.

The gather is ONE ``all_gather_into_tensor`` with equal counts (the ids a rank owns are a pure function of its rank,
so no count exchange is needed), issued on a side stream so the next render batch can start while the blocks travel;
nothing in it synchronises with the host.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

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


def shard_by_cost(cost: torch.Tensor, world_size: int) -> List[torch.Tensor]:
    """Greedy longest-processing-time partition of renders over ranks by a predicted cost (e.g. the footprint class of
    ``renderer.predicted_cost``): ids per rank with near-equal cost sums, instead of ``b mod world``.  Deterministic, so
    every rank computes the same partition from the same host-side costs."""
    cost = torch.as_tensor(cost, dtype=torch.float64).flatten()
    order = torch.argsort(cost, descending=True, stable=True).tolist()
    load = [0.0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        parts[r].append(i)
        load[r] += float(cost[i])
    return [torch.tensor(sorted(p), dtype=torch.int64) for p in parts]


class RefmapGather:
    """Equal-count all-gather of per-rank blocks [n_max, ...] into [world * n_max, ...] on a side stream.

    ``ids_per_rank[r]`` are the global render ids rank r produces (known to every rank); blocks are padded to the
    largest count.  ``launch(local)`` enqueues the collective behind the kernels that wrote ``local`` and returns at
    once; ``result()`` makes the caller's stream wait for it and returns the refmaps ordered by global render id.
    """

    def __init__(self, ids_per_rank: List[torch.Tensor], total: int, device, group=None):
        self.group = group
        self.world = len(ids_per_rank)
        self.total = int(total)
        self.n_max = max(1, max(int(i.numel()) for i in ids_per_rank))
        self.device = torch.device(device)
        # scatter map: where row j of the gathered buffer goes (padding rows go to a dump slot at index `total`)
        dest = torch.full((self.world * self.n_max,), self.total, dtype=torch.int64)
        for r, ids in enumerate(ids_per_rank):
            dest[r * self.n_max: r * self.n_max + ids.numel()] = ids.to(torch.int64)
        self.dest = dest.to(self.device)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._gathered: Optional[torch.Tensor] = None
        self._event = None

    def launch(self, local: torch.Tensor) -> None:
        pad = local
        if local.shape[0] != self.n_max:
            pad = local.new_zeros((self.n_max,) + tuple(local.shape[1:]))
            pad[: local.shape[0]] = local
        out = local.new_empty((self.world * self.n_max,) + tuple(local.shape[1:]))
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(out, pad.contiguous(), group=self.group)
                self._event = torch.cuda.Event()
                self._event.record(self.stream)
            pad.record_stream(self.stream)
            out.record_stream(self.stream)
        else:
            dist.all_gather_into_tensor(out, pad.contiguous(), group=self.group)
        self._gathered = out

    def result(self) -> torch.Tensor:
        assert self._gathered is not None, "launch() first"
        if self._event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._event)
        g = self._gathered
        full = g.new_zeros((self.total + 1,) + tuple(g.shape[1:]))
        full.index_copy_(0, self.dest, g)  # a fixed scatter map: no boolean mask, no host synchronisation
        return full[: self.total]


def all_gather_refmaps(local: torch.Tensor, render_ids: torch.Tensor, total: int,
                       ids_per_rank: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """Gather [n_r, ...] blocks of every rank into [total, ...] ordered by global render id.

    ``ids_per_rank`` (ids of every rank, known everywhere) makes this a single equal-count collective without any
    exchange of counts.  When it is omitted the counts and ids are exchanged first (two small extra collectives)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = local.new_zeros((total,) + tuple(local.shape[1:]))
        out.index_copy_(0, render_ids.to(local.device, torch.int64), local)
        return out
    world = dist.get_world_size()
    if ids_per_rank is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        counts = torch.empty((world,), dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(counts, n)
        n_max = int(counts.max())  # the one host read of this fallback path
        ids = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
        ids[: local.shape[0]] = render_ids.to(local.device)
        all_ids = torch.empty((world * n_max,), dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(all_ids, ids)
        all_ids = all_ids.view(world, n_max).cpu()
        ids_per_rank = [row[row >= 0] for row in all_ids]
    g = RefmapGather(ids_per_rank, total, local.device)
    g.launch(local)
    return g.result()
