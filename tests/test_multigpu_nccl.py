"""GPU, 2+ devices (skipped otherwise): env-sharded KickEnv ranks give the same per-env results as one rank over the
union (Philox is keyed by the GLOBAL env id only through the per-rank seed... here each rank owns its own simulator
shard, so equality is checked for the statistics exchange: RunningMeanStd over NCCL == single-GPU on the union)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bez_isaacgym_b200 import dist as bdist
        from bez_isaacgym_b200.learner import RunningMeanStd, normalize_advantages
        g = torch.Generator().manual_seed(0)
        x_all = torch.randn(65536, 54, generator=g) * 2 + 1
        lo, hi = bdist.shard_range(x_all.shape[0], rank, world)
        mod = RunningMeanStd(54, process_group=dist.group.WORLD).cuda()
        single = RunningMeanStd(54).cuda()
        for it in range(3):
            mod(x_all[lo:hi].cuda() + it)
            single(x_all.cuda() + it)
        assert mod.count.item() == single.count.item() == 1 + 3 * 65536
        assert torch.allclose(mod.running_mean, single.running_mean, rtol=1e-12, atol=1e-12)
        assert torch.allclose(mod.running_var, single.running_var, rtol=1e-11, atol=1e-12)
        r, v = x_all[:, 0].contiguous(), x_all[:, 1].contiguous()
        glob = normalize_advantages(r[lo:hi].cuda(), v[lo:hi].cuda(), process_group=dist.group.WORLD)
        whole = normalize_advantages(r.cuda(), v.cuda())
        assert torch.allclose(glob, whole[lo:hi], rtol=1e-6, atol=1e-6)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)          # a wedged NCCL rendezvous must fail the test, not hang the GPU box
def test_sharded_statistics_equal_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
