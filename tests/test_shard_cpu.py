"""CPU tests of the multi-GPU host logic: the (batch x heads) shard plan, and a world_size-2
`gloo` run in which each rank computes its shard with the oracle and the gathered result equals
the unsharded one (no data-path collective is needed by the product; the all_gather here is the
test's own checker)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flash_attention_from_scratch_b200.shard import plan_shards, shard_for_rank


@pytest.mark.parametrize("B,H,W", [(8, 32, 1), (8, 32, 2), (8, 32, 4), (8, 32, 8), (4, 32, 8),
                                   (1, 16, 8), (3, 16, 2), (6, 4, 4)])
def test_plan_is_a_partition(B, H, W):
    shards = plan_shards(B, H, W)
    assert len(shards) == W
    seen = torch.zeros(B, H, dtype=torch.int32)
    for s in shards:
        seen[s.b0:s.b1, s.h0:s.h1] += 1
        assert s.units == B * H // W
    assert (seen == 1).all()


def test_plan_rejects_uneven():
    with pytest.raises(ValueError):
        plan_shards(1, 3, 2)
    with pytest.raises(ValueError):
        plan_shards(2, 2, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, H, N, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sdpa_ref

    torch.manual_seed(0)  # every rank regenerates the same global problem, keeps only its shard
    q, k, v = (torch.randn(B, N, H, 128).bfloat16() for _ in range(3))
    sh = shard_for_rank(B, H, world, rank)
    o_local = sdpa_ref(sh.take(q).contiguous(), sh.take(k).contiguous(), sh.take(v).contiguous())
    gathered = [torch.empty_like(o_local) for _ in range(world)]
    dist.all_gather(gathered, o_local.contiguous())
    if rank == 0:
        full = torch.empty(B, N, H, 128, dtype=torch.bfloat16)
        for r, g in enumerate(gathered):
            s = shard_for_rank(B, H, world, r)
            full[s.b0:s.b1, :, s.h0:s.h1] = g
        ref = sdpa_ref(q, k, v)
        ret["maxdiff"] = (full.float() - ref.float()).abs().max().item()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,H", [(2, 2), (1, 4)])
def test_world_size_2_gloo_sharded_equals_unsharded(B, H):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, B, H, 128, ret), nprocs=2, join=True)
        # per-(b,h) problems are independent: sharding must not change any value
        assert ret["maxdiff"] <= 2 ** -7
