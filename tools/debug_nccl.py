import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from bez_isaacgym_b200 import dist as bdist, ops
from bez_isaacgym_b200.learner import RunningMeanStd
g = torch.Generator().manual_seed(0)
x_all = torch.randn(65536, 54, generator=g) * 2 + 1
lo, hi = bdist.shard_range(x_all.shape[0], rank, world)
mod = RunningMeanStd(54, process_group=dist.group.WORLD).cuda()
single = RunningMeanStd(54).cuda()
for it in range(2):
    xs = x_all[lo:hi].cuda() + it
    mod._workspace(xs.device)
    mod._pivot.copy_(mod.running_mean)
    ops.rms_moments(xs, mod._pivot, mod._acc, mod._scratch)
    torch.cuda.synchronize()
    print(rank, it, "local acc0", mod._acc[0].item(), "dev", mod._acc.device, flush=True)
    bdist.allreduce_sum_(mod._acc, mod.process_group)
    torch.cuda.synchronize()
    print(rank, it, "reduced acc0", mod._acc[0].item(), flush=True)
    ops.rms_merge(mod._acc, mod._pivot, mod.running_mean, mod.running_var, mod.count.view(1))
    single(x_all.cuda() + it)
    torch.cuda.synchronize()
    print(rank, it, "counts", mod.count.item(), single.count.item(), (mod.running_mean - single.running_mean).abs().max().item(), flush=True)
dist.destroy_process_group()
