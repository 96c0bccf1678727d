#!/usr/bin/env python
"""Launch-latency model of the fused post-physics kernel: t(n) = t0 + n / rate, with and without programmatic
dependent launch (BEZK_PDL, read once per process by libbezk.so -> one child process per setting).

    python tools/exp_post.py            # driver: sweeps the matrix, prints one JSON line per configuration
    python tools/exp_post.py --child    # one configuration (taken from the environment)
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from bez_isaacgym_b200 import ops, synthetic_gym as sg
    dev = torch.device("cuda:0")
    out = {"pdl": os.environ.get("BEZK_PDL", "1")}
    for n in (4096, 65536, 262144, 524288, 1048576, 2097152):
        st = sg.make_state(n, seed=1, device=dev, filler=(n <= 262144))
        goal, ball_init, default, lo, hi = sg.make_constants(n, dev)
        cfg = ops.make_task_cfg(reset_root_states=False)
        actions = sg.make_actions(n, device=dev)
        targets = torch.empty(n, 18, device=dev)
        obs = torch.empty(n, 54, device=dev); rew = torch.empty(n, device=dev)
        progress, reset = sg.make_bookkeeping(n, device=dev)
        timeout = torch.empty(n, dtype=torch.long, device=dev)
        prev = torch.zeros(n, 3, device=dev)

        def post():
            ops.post_physics(st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, ball_init, None,
                             reset, progress, timeout, cfg, obs, rew, prev_lin_vel=prev)

        def both():
            ops.pre_physics(actions, targets, cfg)
            post()

        def timeit(body, reps=32, iters=15):
            for _ in range(3):
                body()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    body()
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            return ts[len(ts) // 2] * 1e3 / reps

        out[f"post_us_{n}"] = round(timeit(post), 2)
        out[f"step_us_{n}"] = round(timeit(both), 2)
        del st
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.child:
        return child()
    lines = []
    for pdl in (0, 1):
        env = dict(os.environ, BEZK_PDL=str(pdl))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"pdl": pdl, "error": r.stderr[-400:]})
        print(line, flush=True)
        lines.append(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
