import os, sys, json, torch
sys.path.insert(0, "/root/repo")
from bez_isaacgym_b200 import bez_model as bm, synthetic_gym as sg, tasks as T
from bez_isaacgym_b200.synthetic_sim import SyntheticGym
dev = torch.device("cuda:0")
class Sim(SyntheticGym):
    owns_root_reset = True
n = 262144
for task, cls in (("walk", T.WalkEnv), ("orient", T.OrientEnv)):
    cfg = bm.default_task_cfg(n, task=task); cfg["env"]["imuPrevVelAliasing"] = False
    env = cls(cfg, "cuda:0", 0, True, sim=Sim(n, device="cuda:0", task=task))
    p, r = sg.make_bookkeeping(n, device=dev, max_episode_length=600); env.progress_buf.copy_(p); env.reset_buf.copy_(r)
    for _ in range(5): env.post_physics_step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(32): env.post_physics_step()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 160 * 1e3
    print(task, round(us, 2), "us  frac", round(432 * n / (us * 1e-6) / 1e9 / 6546.6, 3))
    del env, g
