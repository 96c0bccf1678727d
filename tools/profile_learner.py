"""Small driver for ncu: runs the learner kernels a few times at sizes where they are HBM-bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bez_isaacgym_b200 import ops, synthetic_gym as sg
dev = torch.device("cuda:0")
m = 2097152
x = torch.randn(m, 54, device=dev)
mean = torch.zeros(54, dtype=torch.float64, device=dev); var = torch.ones(54, dtype=torch.float64, device=dev)
count = torch.ones(1, dtype=torch.float64, device=dev)
acc = torch.empty(109, dtype=torch.float64, device=dev)
scratch = torch.empty(ops.rms_scratch_doubles(54), dtype=torch.float64, device=dev)
y = torch.empty_like(x)
mb = {k: t.to(dev).contiguous() for k, t in sg.make_minibatch(1048576).items()}
cfgp = ops.make_ppo_cfg()
stats = torch.empty(8, dtype=torch.float64, device=dev)
part = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device=dev)
gmu = torch.empty(1048576, 18, device=dev); gv = torch.empty(1048576, device=dev); gls = torch.empty(18, device=dev)
for _ in range(3):
    ops.rms_moments(x, mean, acc, scratch)
    ops.rms_merge(acc, mean, mean, var, count)
    ops.rms_normalize(x, mean, var, y)
    ops.ppo_loss(mb["actions"], mb["mu"], mb["logstd"], mb["old_mu"], mb["old_sigma"], mb["values"].view(-1),
                 mb["old_values"].view(-1), mb["returns"].view(-1), mb["old_neglogp"], mb["advantages"], cfgp, stats, part,
                 grad_mu=gmu, grad_values=gv, grad_logstd=gls)
torch.cuda.synchronize()
print("ok")
