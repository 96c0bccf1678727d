#!/usr/bin/env python
"""Per-kernel timing of every libbezk kernel at the BASELINE sizes (CUDA events, warm-up, L2 flushed between
iterations by a 256 MB memset when the working set is smaller than L2).  Prints one JSON object per kernel:
algorithmic GB/s and fraction of the measured HBM peak.  usage: python tools/bench_kernels.py [--out profiles/x.jsonl]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bez_isaacgym_b200 import ops, synthetic_gym as sg  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, iters=20, warm=3, flush=None, reps=10):
    """Median time of one `fn()`: `reps` back-to-back calls captured in a CUDA graph (so the CPU-side ctypes launch
    cost is not in the number), replayed `iters` times between CUDA events.  `flush` (a buffer larger than L2) is
    zeroed inside the graph before every call when the working set would otherwise stay L2-resident; the time of the
    flush alone is measured the same way and subtracted."""
    def graph_of(body):
        for _ in range(warm):
            body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                body()
        return g

    def run(g):
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2] * 1e-3 / reps

    if flush is None:
        return run(graph_of(fn))
    t_flush = run(graph_of(lambda: flush.zero_()))
    return max(run(graph_of(lambda: (flush.zero_(), fn()))) - t_flush, 1e-9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="all", choices=["all", "task", "learner", "rollout"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def report(name, units, bytes_per_unit, secs, note=""):
        gbs = units * bytes_per_unit / secs / 1e9
        rows.append(dict(kernel=name, units=units, bytes_per_unit=bytes_per_unit, us=secs * 1e6, algorithmic_gbs=gbs,
                         frac_of_measured_peak=gbs / pk, note=note))
        print(json.dumps(rows[-1]))

    # ---- task kernels across env counts
    for n in (4096, 65536, 262144, 1048576) if args.only in ("all", "task") else ():
        st = sg.make_state(n, seed=1, device=dev, filler=(n <= 262144))
        goal, ball_init, default, lo, hi = sg.make_constants(n, dev)
        cfg = ops.make_task_cfg(reset_root_states=False)
        actions = sg.make_actions(n, device=dev)
        targets = torch.empty(n, 18, device=dev)
        obs = torch.empty(n, 54, device=dev); rew = torch.empty(n, device=dev)
        progress, reset = sg.make_bookkeeping(n, device=dev)
        timeout = torch.empty(n, dtype=torch.long, device=dev)
        prev = torch.zeros(n, 3, device=dev)
        fl = flush if n < 262144 else None
        report("pre_physics", n, 144, timeit(lambda: ops.pre_physics(actions, targets, cfg), flush=fl), f"n={n}")
        report("post_physics_fused", n, 536,
               timeit(lambda: ops.post_physics(st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, ball_init, None,
                                               reset, progress, timeout, cfg, obs, rew, prev_lin_vel=prev), flush=fl), f"n={n}")
        report("obs_kernel(parts=3)", n, 472,
               timeit(lambda: ops.post_physics(st.dof_state, st.rigid_body, st.root_states, st.net_contact, goal, ball_init, None,
                                               reset, progress, timeout, cfg, obs, None, prev_lin_vel=prev, parts=3), flush=fl), f"n={n}")
        report("reward_kernel(parts=4)", n, 184,
               timeit(lambda: ops.post_physics(st.dof_state, st.rigid_body, st.root_states, None, goal, ball_init, None,
                                               reset, progress, None, cfg, None, rew, parts=4), flush=fl), f"n={n}")
        del st
        if n in (4096, 262144):
            # sibling tasks: fused step, 432 B/env = R dof 144 + root 52 + imu 40 + feet 24 + goal 8 + reset/progress 16
            # (+ prev 12) ; W obs 208 + rew 4 + reset/progress/timeout 24 (+ prev 12)
            for task in ("walk", "orient"):
                sw = sg.make_state(n, seed=1, device=dev, task=task)
                cfgw = ops.make_task_cfg(num_bodies=sw.num_bodies, max_episode_length=600, reset_root_states=False)
                gw = torch.tensor([[2.0, 0.0]], device=dev).repeat(n, 1); ga = torch.full((n,), 1.5708, device=dev)
                obw = torch.empty(n, 52, device=dev)
                pw, rw = sg.make_bookkeeping(n, device=dev, max_episode_length=600)
                report(f"post_physics_fused[{task}]", n, 432,
                       timeit(lambda: ops.post_physics_task(task, sw.dof_state, sw.rigid_body, sw.root_states, sw.net_contact, gw, None,
                                                            rw, pw, timeout, cfgw, obw, rew, goal_angle=ga, prev_lin_vel=prev), flush=fl),
                       f"n={n}")
                del sw
    # ---- GAE
    for n in (4096, 262144) if args.only in ("all", "learner") else ():
        r, v, d, lv, ld = sg.make_rollout(n, 32, device=dev)
        adv, ret = torch.empty_like(r), torch.empty_like(r)
        report("gae(T=32)", n * 32, 17, timeit(lambda: ops.gae(r, v, d, lv, ld, 0.99, 0.95, adv, ret),
                                               flush=flush if n < 262144 else None), f"n={n}")
    # ---- RunningMeanStd train forward, advantage normalisation, PPO loss
    for m in (32768, 131072, 1048576, 8388608) if args.only in ("all", "learner") else ():
        x = torch.randn(m, 54, device=dev)
        mean = torch.zeros(54, dtype=torch.float64, device=dev); var = torch.ones(54, dtype=torch.float64, device=dev)
        count = torch.ones(1, dtype=torch.float64, device=dev)
        acc = torch.empty(109, dtype=torch.float64, device=dev)
        scratch = torch.empty(ops.rms_scratch_doubles(54), dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        fl = flush if m * 216 < (200 << 20) else None

        def rms_train():
            ops.rms_moments(x, mean, acc, scratch)
            ops.rms_merge(acc, mean, mean, var, count)
            ops.rms_normalize(x, mean, var, y)
        report("rms_train_forward(obs), separate entries", m, 432, timeit(rms_train, flush=fl), f"m={m} (bezk_rms_moments + merge + normalize, 4 launches)")
        report("rms_train_forward(obs)", m, 432, timeit(lambda: ops.rms_train_forward(x, mean, var, count, y, scratch), flush=fl),
               f"m={m} (bezk_rms_train_forward: moments, fold + snapshot, merge + normalise = 3 launches)")
        report("rms_moments(obs)", m, 216, timeit(lambda: ops.rms_moments(x, mean, acc, scratch), flush=fl), f"m={m}")
        report("rms_normalize(obs)", m, 432, timeit(lambda: ops.rms_normalize(x, mean, var, y), flush=fl), f"m={m}")
        del x, y
        if m > 1048576:
            continue
        mb = {k: t.to(dev).contiguous() for k, t in sg.make_minibatch(m).items()}
        cfgp = ops.make_ppo_cfg()
        stats = torch.empty(8, dtype=torch.float64, device=dev)
        part = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device=dev)
        gmu = torch.empty(m, 18, device=dev); gv = torch.empty(m, device=dev); gls = torch.empty(18, device=dev)
        report("ppo_loss_fwd_bwd", m, 384,
               timeit(lambda: ops.ppo_loss(mb["actions"], mb["mu"], mb["logstd"], mb["old_mu"], mb["old_sigma"], mb["values"].view(-1),
                                           mb["old_values"].view(-1), mb["returns"].view(-1), mb["old_neglogp"], mb["advantages"],
                                           cfgp, stats, part, grad_mu=gmu, grad_values=gv, grad_logstd=gls), flush=fl), f"m={m}")
        ret, val = torch.randn(m, device=dev), torch.randn(m, device=dev)
        a3 = torch.empty(3, dtype=torch.float64, device=dev); sc1 = torch.empty(ops.rms_scratch_doubles(1), dtype=torch.float64, device=dev)
        out = torch.empty(m, device=dev)

        def advn():
            ops.adv_moments(ret, val, a3, sc1)
            ops.adv_normalize(ret, val, a3, out)
        report("adv_normalize(2 pass)", m, 20, timeit(advn, flush=fl), f"m={m}")
    # ---- rollout storage: flattening pass vs slab-addressed minibatches, policy head
    if args.only in ("all", "rollout"):
        T = 32
        for n in (4096, 262144):
            obses = torch.randn(T, n, 54, device=dev)
            flat = torch.empty(n * T, 54, device=dev)
            fl = flush if n * T * 216 < (200 << 20) else None
            report("swap_and_flatten01(obses)", n * T, 432, timeit(lambda: ops.swap_and_flatten01(obses, out=flat), flush=fl),
                   f"(32,{n},54) f32: 216 B read + 216 B written per sample")
            acts = torch.randn(T, n, 18, device=dev); flat_a = torch.empty(n * T, 18, device=dev)
            report("swap_and_flatten01(actions)", n * T, 144, timeit(lambda: ops.swap_and_flatten01(acts, out=flat_a), flush=fl),
                   f"(32,{n},18) f32")
            vals = torch.randn(T, n, 1, device=dev); flat_v = torch.empty(n * T, 1, device=dev)
            report("swap_and_flatten01(values)", n * T, 8, timeit(lambda: ops.swap_and_flatten01(vals, out=flat_v), flush=fl),
                   f"(32,{n},1) f32")
            t0 = timeit(lambda: flat.copy_(obses.transpose(0, 1).reshape(n * T, 54)), flush=fl)
            report("torch transpose+reshape copy (obses)", n * T, 432, t0, "what rl_games' swap_and_flatten01 costs in torch")
            # one minibatch (32768 samples = 1024 envs x 32): flattened-contiguous vs slab view, normalise and moments
            mb, E = 32768, 1024
            mean = torch.zeros(54, dtype=torch.float64, device=dev); var = torch.ones(54, dtype=torch.float64, device=dev)
            acc = torch.empty(109, dtype=torch.float64, device=dev)
            scratch = torch.empty(ops.rms_scratch_doubles(54), dtype=torch.float64, device=dev)
            y = torch.empty(mb, 54, device=dev)
            cont = flat[:mb]
            view = obses[:, :E]
            report("rms_normalize(minibatch, contiguous)", mb, 432, timeit(lambda: ops.rms_normalize(cont, mean, var, y), flush=flush), f"n={n}")
            report("rms_normalize_slabs(minibatch view)", mb, 432, timeit(lambda: ops.rms_normalize_slabs(view, mean, var, y), flush=flush), f"n={n}")
            report("rms_moments(minibatch, contiguous)", mb, 216, timeit(lambda: ops.rms_moments(cont, mean, acc, scratch), flush=flush), f"n={n}")
            report("rms_moments_slabs(minibatch view)", mb, 216, timeit(lambda: ops.rms_moments_slabs(view, mean, acc, scratch), flush=flush), f"n={n}")
            del obses, flat, acts, flat_a
            # policy head: mu 72 + value 4 read; actions, mus, sigmas, targets 4 x 72 + neglogp 4 + values 4 written
            mu = torch.randn(n, 18, device=dev); logstd = torch.zeros(18, device=dev); vn = torch.randn(n, device=dev)
            vm = torch.zeros(1, dtype=torch.float64, device=dev); vv = torch.ones(1, dtype=torch.float64, device=dev)
            o = [torch.empty(n, 18, device=dev) for _ in range(4)]
            nl = torch.empty(n, device=dev); vo = torch.empty(n, device=dev)
            cfg = ops.make_task_cfg()
            xa = torch.randn(n, 54, device=dev); ca = torch.randn(n, 54, device=dev)
            ncfg = ops.make_noise_cfg("gaussian", "additive", a=0.002)
            report("dr_noise(obs, philox)", n * 54, 12, timeit(lambda: ops.dr_noise(xa, ncfg, corr=ca, seed=1, step=2), flush=flush if n < 262144 else None),
                   f"n={n}: x + corr read, y written, white noise generated in registers")
            fl2 = flush if n < 262144 else None
            report("policy_head(philox)", n, 372,
                   timeit(lambda: ops.policy_head(mu, logstd, vn, vm, vv, 1e-5, noise=None, seed=1, step=2, actions=o[0], neglogp=nl,
                                                  values=vo, mus=o[1], sigmas=o[2], task_cfg=cfg, targets=o[3]), flush=fl2), f"n={n}")
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
