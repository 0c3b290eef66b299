"""Multi-process host logic on CPU: world_size 2, gloo.  Covers envmap ownership, render sharding and the padded
all-gather that returns rendered refmaps ordered by global render id (what bench.py --gpus N does over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drmnet_b200.dist import RefmapGather, all_gather_refmaps, owned_envmaps, shard_by_cost, shard_renders


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n_env, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        env_index = torch.randint(0, n_env, (total,), generator=g)  # same on every rank
        mine, local_slot = shard_renders(env_index, world, rank)
        owned = owned_envmaps(n_env, world, rank)
        assert all(owned[int(s)] == int(env_index[i]) for i, s in zip(mine, local_slot))
        # stand-in for the render: a block whose content encodes the global render id
        local = torch.stack([torch.full((3, 4, 4), float(i)) for i in mine]) if len(mine) else torch.zeros(0, 3, 4, 4)
        full = all_gather_refmaps(local, mine, total)
        expect = torch.arange(total, dtype=torch.float32)[:, None, None, None].expand(total, 3, 4, 4)
        assert torch.equal(full, expect)
        # the fixed-count path bench.py uses: every rank knows every rank's ids, one collective, no count exchange
        ids_per_rank = [shard_renders(env_index, world, r)[0] for r in range(world)]
        g = RefmapGather(ids_per_rank, total, "cpu")
        g.launch(local)
        assert torch.equal(g.result(), expect)
        # cost-balanced sharding: a partition of the renders, the same on every rank
        cost = torch.rand(total, generator=torch.Generator().manual_seed(1)) ** 4
        parts = shard_by_cost(cost, world)
        assert sorted(int(i) for p in parts for i in p) == list(range(total))
        local2 = torch.stack([torch.full((3, 4, 4), float(i)) for i in parts[rank]]) if len(parts[rank]) else torch.zeros(0, 3, 4, 4)
        g2 = RefmapGather(parts, total, "cpu")
        g2.launch(local2)
        assert torch.equal(g2.result(), expect)
        torch.save(full, os.path.join(result_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,n_env", [(10, 4), (7, 3), (3, 8)])
def test_sharded_render_gather_world2(tmp_path, total, n_env):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, n_env, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(a, b) and a.shape == (total, 3, 4, 4)


def test_single_process_gather_is_a_scatter_by_id():
    local = torch.arange(3, dtype=torch.float32)[:, None].expand(3, 2).contiguous()
    out = all_gather_refmaps(local, torch.tensor([2, 0, 1]), 3)
    assert out[:, 0].tolist() == [1.0, 2.0, 0.0]


def test_ownership_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in owned_envmaps(13, world, r))
        assert seen == list(range(13))


def test_cost_balanced_sharding_is_balanced():
    cost = torch.tensor([16.0] * 4 + [4.0] * 8 + [1.0] * 52)  # a footprint mix: a few expensive renders
    for world in (2, 4, 8):
        parts = shard_by_cost(cost, world)
        loads = [float(cost[p].sum()) for p in parts]
        assert sorted(int(i) for p in parts for i in p) == list(range(64))
        assert max(loads) - min(loads) <= 16.0 and max(loads) <= 1.15 * sum(loads) / world + 1e-9
