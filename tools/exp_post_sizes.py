"""post-physics kernel time (graph of 32 back-to-back launches) at several env counts, for the library in BEZK_LIB."""
import os, sys, json, torch
sys.path.insert(0, "/root/repo")
from bez_isaacgym_b200 import bez_model as bm, synthetic_gym as sg
from bez_isaacgym_b200.synthetic_sim import SyntheticGym
from bez_isaacgym_b200.tasks.kick_env import KickEnv
dev = torch.device("cuda:0")
class Sim(SyntheticGym):
    owns_root_reset = True
out = {}
for n in [int(x) for x in os.environ.get("SIZES", "4096,65536,262144,1048576").split(",")]:
    cfg = bm.default_task_cfg(n); cfg["env"]["imuPrevVelAliasing"] = False
    env = KickEnv(cfg, "cuda:0", 0, True, sim=Sim(n, device="cuda:0", filler=(n <= 262144)))
    p, r = sg.make_bookkeeping(n, device=dev); env.progress_buf.copy_(p); env.reset_buf.copy_(r)
    vals = torch.randn(n, device=dev); sh = torch.empty(n, device=dev); dn = torch.empty(n, dtype=torch.uint8, device=dev)
    env.set_rollout_targets(values=vals, shaped_rewards=sh, dones_u8=dn)
    for _ in range(5): env.post_physics_step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(32): env.post_physics_step()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); 
        for _ in range(5): g.replay()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 160 * 1e3)
    ts.sort()
    out[n] = round(ts[len(ts)//2], 2)
    del env, g
    torch.cuda.empty_cache()
print(os.environ.get("TAG", ""), json.dumps(out), "frac@262144", round(545*262144/(out.get(262144, 1e9)*1e-6)/1e9/6546.6, 4) if 262144 in out else "")
