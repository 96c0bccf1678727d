"""CPU, world_size 2, gloo: the host-side multi-GPU logic (env sharding, packed SUM all-reduce of additive
statistics, flat-bucket gradient all-reduce, parameter broadcast).  The per-shard moments that the CUDA kernel
would produce are computed here with plain torch (test-only) so the EXACTNESS of the cross-rank merge is what is
being checked: merged stats over 2 shards == stats of the union batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bez_isaacgym_b200 import dist as bdist
from oracle import rl_games_oracle as rg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pivoted_moments(x, pivot):
    d = x.double() - pivot
    return torch.cat([torch.tensor([float(x.shape[0])], dtype=torch.float64), d.sum(0), (d * d).sum(0)])


def _merge(acc, pivot, mean, var, count):
    """Same arithmetic as bezk_rms_merge (csrc/bezk_learner.cu rms_merge_kernel)."""
    c = mean.numel()
    b = acc[0]
    s, ss = acc[1:1 + c], acc[1 + c:]
    mean_b = pivot + s / b
    var_b = (ss - s * s / b) / (b - 1.0)
    return rg.RunningMeanStd.merge(mean, var, count, mean_b, var_b, b)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert bdist.is_distributed()
        n_envs, c = 1001, 54
        g = torch.Generator().manual_seed(0)
        x_all = torch.randn(n_envs, c, generator=g) * 3 + 1          # the union batch (same on both ranks)
        lo, hi = bdist.shard_range(n_envs, rank, world)
        mean = torch.randn(c, dtype=torch.float64, generator=g); var = torch.rand(c, dtype=torch.float64, generator=g) + 0.5
        count = torch.tensor(77.0, dtype=torch.float64)
        pivot = mean.clone()
        acc = _pivoted_moments(x_all[lo:hi], pivot)
        adv_acc = torch.tensor([hi - lo, float(x_all[lo:hi, 0].double().sum()), float((x_all[lo:hi, 0].double() ** 2).sum())],
                               dtype=torch.float64)
        flat, (acc_v, adv_v) = bdist.pack([acc, adv_acc])            # ONE collective for both statistics
        bdist.allreduce_sum_(flat, dist.group.WORLD)
        local_only = acc.clone()
        assert torch.equal(bdist.allreduce_sum_(local_only), acc)       # group=None -> local statistics, no collective
        whole = _pivoted_moments(x_all, pivot)
        assert torch.allclose(acc_v, whole, rtol=1e-13, atol=1e-9)
        assert adv_v[0].item() == n_envs
        got = _merge(acc_v, pivot, mean, var, count)
        want = _merge(whole, pivot, mean, var, count)
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=1e-13, atol=1e-12)
        # ... and equals the reference update on the union batch up to its fp32 batch moments
        ref = rg.RunningMeanStd.merge(mean, var, count, x_all.mean(0), x_all.var(0), n_envs)
        assert torch.allclose(got[0], ref[0], rtol=1e-5, atol=1e-6) and torch.allclose(got[1], ref[1], rtol=1e-5, atol=1e-6)

        # gradient all-reduce on one flat bucket + parameter broadcast
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(54, 8), torch.nn.ELU(), torch.nn.Linear(8, 18))
        bdist.broadcast_parameters(net, src=0)
        ref_w = [p.detach().clone() for p in net.parameters()]
        gathered = [torch.empty_like(ref_w[0]) for _ in range(world)]
        dist.all_gather(gathered, ref_w[0])
        assert torch.equal(gathered[0], gathered[1])
        net(x_all[lo:hi]).pow(2).mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        bdist.allreduce_grads_(list(net.parameters()))
        for p, l in zip(net.parameters(), local):
            both = [torch.empty_like(l) for _ in range(world)]
            dist.all_gather(both, l)
            assert torch.allclose(p.grad, (both[0] + both[1]) / 2, rtol=1e-6, atol=1e-7)
        # the learner's group convention: None resolves to WORLD under torchrun, so statistics are shared exactly like
        # the gradients; the low-level helper keeps "None = local"
        assert bdist.resolve_group(None) is dist.group.WORLD and bdist.resolve_group(dist.group.WORLD) is dist.group.WORLD
        shared = acc.clone()
        bdist.allreduce_sum_(shared, bdist.resolve_group(None))
        assert torch.allclose(shared, whole, rtol=1e-13, atol=1e-9)
        # env-sharded ranks key their Philox noise by global env ids: shard offsets tile [0, n) without gaps
        bases = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(bases, torch.tensor([lo]))
        assert [int(b) for b in bases] == [bdist.shard_range(n_envs, r, world)[0] for r in range(world)]
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_and_gradient_exchange(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_partitions_exactly():
    for n in (1, 7, 4096, 262144, 1000003):
        for world in (1, 2, 4, 8):
            blocks = [bdist.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        bdist.shard_range(10, 3, 2)


def test_single_process_collectives_are_noops():
    assert bdist.resolve_group(None) is None                 # no process group: the learner stays single-process
    t = torch.arange(5, dtype=torch.float64)
    assert torch.equal(bdist.allreduce_sum_(t.clone()), t)
    flat, views = bdist.pack([torch.ones(3, dtype=torch.float64), torch.zeros((), dtype=torch.float64)])
    assert flat.numel() == 4 and views[0].shape == (3,) and views[1].shape == ()
    views[0][1] = 5.0
    assert flat[1] == 5.0                                    # views alias the packed buffer
