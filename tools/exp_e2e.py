#!/usr/bin/env python
"""Host-pipeline (e2e) experiment: KickEnv.step with the simulator tensors in pinned host memory -- wall time per env step
(one stream sync per step, results on the host) for the zero-copy, staged and copy-engine (staged_ce, by chunk count)
pipelines.  One JSON line per configuration.   python tools/exp_e2e.py [--envs 262144] [--steps 64]"""
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
    ap.add_argument("--envs", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--modes", default="zero_copy,staged_ce:4,staged_ce:4s,staged_pack:2,staged_pack:4,staged_pack:8")
    ap.add_argument("--timeline", action="store_true", help="add the per-chunk device / host timeline of one step (staged_ce / staged_pack)")
    ap.add_argument("--pack-threads", type=int, default=0)
    ap.add_argument("--pack-dof", type=int, default=0)
    ap.add_argument("--pack-spin-us", type=int, default=-1)
    args = ap.parse_args()
    n = args.envs
    # under torchrun: one process per GPU, a gloo barrier before each timed loop (all ranks load the host link together)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    class OwnedRootSim(SyntheticGym):
        owns_root_reset = True

    sim = OwnedRootSim(n, device=dev, seed=1 + rank, host=True, filler=True)
    act = sg.make_actions(n, seed=1).pin_memory()
    for spec in args.modes.split(","):
        mode, _, chunks = spec.partition(":")
        split = chunks.endswith("s")
        chunks = chunks.rstrip("s")
        cfg = bm.default_task_cfg(n, use_gpu_pipeline=False, rl_device="cpu")
        cfg["env"]["imuPrevVelAliasing"] = False
        cfg["env"]["hostPipeline"] = mode
        if chunks:
            cfg["env"]["hostPipelineChunks"] = int(chunks) if chunks.isdigit() else [float(w) for w in chunks.split("-")]
        cfg["env"]["hostPipelineSplitSparse"] = split
        cfg["env"]["hostPipelineTimeline"] = args.timeline and mode in ("staged_ce", "staged_pack")
        cfg["env"]["hostPackThreads"] = args.pack_threads
        cfg["env"]["hostPackDof"] = bool(args.pack_dof)
        cfg["env"]["hostPackSpinUs"] = args.pack_spin_us
        env = KickEnv(cfg, dev, 0, True, sim=sim)
        for _ in range(5):
            env.step(act)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            env.step(act)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / args.steps
        barrier()
        # the two halves of a step
        t_pre = t_post = 0.0
        for _ in range(8):
            torch.cuda.synchronize(); a = time.perf_counter()
            env.pre_physics_step(act)
            torch.cuda.synchronize(); b = time.perf_counter()
            env._ce_copy_out = mode in ("staged_ce", "staged_pack")
            env.post_physics_step()
            env._ce_copy_out = False
            torch.cuda.synchronize(); c = time.perf_counter()
            t_pre += b - a; t_post += c - b
        tl = None
        if cfg["env"]["hostPipelineTimeline"]:
            env.step(act)
            torch.cuda.synchronize()
            tl = env.host_timeline()
        print(json.dumps({"rank": rank, "world": world, "pack_threads": getattr(env, "host_pack_threads", None), "mode": mode, "pack_dof": args.pack_dof if mode == "staged_pack" else None, "chunks": chunks or None, "split_sparse": split, "envs": n, "ms_per_step": round(ms, 4),
                          "env_steps_per_s_M": round(n / ms / 1e3, 2), "pre_ms": round(1e3 * t_pre / 8, 4),
                          "post_ms": round(1e3 * t_post / 8, 4), "timeline": tl}), flush=True)
        del env


if __name__ == "__main__":
    main()
