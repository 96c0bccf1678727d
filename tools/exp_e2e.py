#!/usr/bin/env python
"""Host-pipeline (e2e) experiment: KickEnv.step with the simulator tensors in pinned host memory, per-step wall time and the
device time of K0 / the post-physics kernel, for the zero-copy and the staged pipelines.  Knobs come from the environment
(BEZK_SMART_GRANULE=0|1|2).   python tools/exp_e2e.py [--envs 65536]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bez_isaacgym_b200 import bez_model as bm, synthetic_gym as sg  # noqa: E402
from bez_isaacgym_b200.synthetic_sim import SyntheticGym  # noqa: E402
from bez_isaacgym_b200.tasks.kick_env import KickEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=64)
    args = ap.parse_args()
    n = args.envs
    out = {"envs": n, "smart_granule": os.environ.get("BEZK_SMART_GRANULE", "1")}

    class OwnedRootSim(SyntheticGym):
        owns_root_reset = True

    for mode in ("zero_copy", "staged"):
        cfg = bm.default_task_cfg(n, use_gpu_pipeline=False, rl_device="cpu")
        cfg["env"]["imuPrevVelAliasing"] = False
        cfg["env"]["hostPipeline"] = mode
        env = KickEnv(cfg, "cuda:0", 0, True, sim=OwnedRootSim(n, device="cuda:0", seed=1, host=True, filler=True))
        act = sg.make_actions(n, seed=1).pin_memory()
        for _ in range(5):
            env.step(act)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            env.step(act)
        torch.cuda.synchronize()
        out[f"{mode}_ms_per_step"] = round(1e3 * (time.perf_counter() - t0) / args.steps, 4)
        # device time of the two kernels alone
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        k0 = post = 0.0
        for _ in range(16):
            ev[0].record(); env.pre_physics_step(act); ev[1].record(); env.post_physics_step(); ev[2].record()
            torch.cuda.synchronize()
            k0 += ev[0].elapsed_time(ev[1]); post += ev[1].elapsed_time(ev[2])
        out[f"{mode}_k0_ms"] = round(k0 / 16, 4)
        out[f"{mode}_post_ms"] = round(post / 16, 4)
        del env
    print(json.dumps(out))


if __name__ == "__main__":
    main()
