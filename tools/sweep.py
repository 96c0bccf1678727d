#!/usr/bin/env python
"""BASELINE configs[1]/[2]/[4]: env-count sweep 4096 -> 1M of the BezKick step (K0 + fused post-physics) and of GAE, this
repo's CUDA path vs the reference's torch ops (oracle port) on the SAME GPU and on the host CPU; plus the PPO epoch math
(RunningMeanStd + GAE + advantage normalisation + loss) on the 4096 x 32 rollout.  One JSON line per row.

    python tools/sweep.py [--out profiles/r01_sweep.jsonl] [--no-cpu] [--max-envs 1048576]

Measurement tool, not product code: ``oracle/`` is imported here only as the BASELINE BEING TIMED (the same role as
``bench.py``'s ``cpu_baseline`` leg); every "bezk" row runs the C-ABI kernels and nothing else.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bez_isaacgym_b200 import bez_model as bm, ops, synthetic_gym as sg  # noqa: E402
from bez_isaacgym_b200.synthetic_sim import SyntheticGym  # noqa: E402
from bez_isaacgym_b200.tasks import KickEnv  # noqa: E402
from oracle import rl_games_oracle as rg  # noqa: E402
from oracle import task_oracle as to  # noqa: E402


def gpu_time(fn, reps=10, iters=10, graph=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        run = g.replay
    else:
        def run():
            for _ in range(reps):
                fn()
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def wall_time(fn, iters):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def make_oracle(n, device):
    st = sg.make_state(n, seed=1, device=device, filler=False)
    goal, ball_init, default, lo, hi = sg.make_constants(n, device)
    orc = to.KickStepOracle(n, st.root_states, st.dof_state, st.rigid_body, st.net_contact, default, lo, hi, goal, ball_init,
                            torch.tensor([0.0, 0.0], device=device), st.root_states.clone(), alias_prev_lin_vel=False)
    orc.prev_lin_vel = torch.zeros(n, 3, device=device)
    progress, reset = sg.make_bookkeeping(n, device=device)
    orc.progress_buf[:] = progress
    orc.reset_buf[:] = reset
    actions = sg.make_actions(n, device=device)

    def step():
        orc.pre_physics_step(actions)
        orc.post_physics_step()
    return step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--max-envs", type=int, default=1048576)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []

    def emit(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    sizes = [n for n in (4096, 16384, 65536, 262144, 1048576) if n <= args.max_envs]
    for n in sizes:
        class OwnedRootSim(SyntheticGym):
            owns_root_reset = True
        cfg = bm.default_task_cfg(n)
        cfg["env"]["imuPrevVelAliasing"] = False
        env = KickEnv(cfg, "cuda:0", 0, True, sim=OwnedRootSim(n, device="cuda:0", filler=(n <= 262144)))
        progress, reset = sg.make_bookkeeping(n, device=dev)
        env.progress_buf.copy_(progress); env.reset_buf.copy_(reset)
        actions = sg.make_actions(n, device=dev)

        def step():
            env.pre_physics_step(actions)
            env.post_physics_step()
        t_graph = gpu_time(step)
        t_eager = gpu_time(step, graph=False)
        t_api = wall_time(lambda: (env.step(actions), torch.cuda.synchronize()), 20)
        emit(what="task_step", impl="bezk CUDA (graph replay)", envs=n, us=t_graph * 1e6, env_steps_per_s=n / t_graph)
        emit(what="task_step", impl="bezk CUDA (eager launches)", envs=n, us=t_eager * 1e6, env_steps_per_s=n / t_eager)
        emit(what="task_step", impl="bezk KickEnv.step() + sync (python API, wall clock)", envs=n, us=t_api * 1e6, env_steps_per_s=n / t_api)
        del env
        torch.cuda.empty_cache()
        ref_gpu = make_oracle(n, dev)
        t_ref_gpu = gpu_time(ref_gpu, reps=2, iters=5, graph=False)
        emit(what="task_step", impl="reference torch ops on the same B200 (oracle port, eager)", envs=n, us=t_ref_gpu * 1e6,
             env_steps_per_s=n / t_ref_gpu)
        del ref_gpu
        torch.cuda.empty_cache()
        if not args.no_cpu and n <= 262144:
            torch.set_num_threads(os.cpu_count())
            ref_cpu = make_oracle(n, "cpu")
            t_cpu = wall_time(ref_cpu, 5 if n <= 65536 else 3)
            emit(what="task_step", impl=f"reference torch ops on the host CPU ({os.cpu_count()} threads)", envs=n, us=t_cpu * 1e6,
                 env_steps_per_s=n / t_cpu)
        # GAE
        r, v, d, lv, ld = sg.make_rollout(n, 32, device=dev)
        adv, ret = torch.empty_like(r), torch.empty_like(r)
        t = gpu_time(lambda: ops.gae(r, v, d, lv, ld, 0.99, 0.95, adv, ret))
        emit(what="gae_T32", impl="bezk CUDA (graph replay)", envs=n, us=t * 1e6, samples_per_s=n * 32 / t)
        df, ldf = d.float(), ld.float()
        t = gpu_time(lambda: rg.discount_values(ldf, lv, df, v, r, 0.99, 0.95) + v, reps=2, iters=5, graph=False)
        emit(what="gae_T32", impl="reference torch ops on the same B200 (eager)", envs=n, us=t * 1e6, samples_per_s=n * 32 / t)
        if not args.no_cpu and n <= 262144:
            rc, vc, dc, lvc, ldc = r.cpu(), v.cpu(), df.cpu(), lv.cpu(), ldf.cpu()
            t = wall_time(lambda: rg.discount_values(ldc, lvc, dc, vc, rc, 0.99, 0.95) + vc, 5)
            emit(what="gae_T32", impl=f"reference torch ops on the host CPU ({os.cpu_count()} threads)", envs=n, us=t * 1e6,
                 samples_per_s=n * 32 / t)
        del r, v, d, lv, ld, adv, ret
        torch.cuda.empty_cache()

    # ---- configs[2]: PPO epoch math on the 4096 x 32 rollout: GAE + value RMS x2 + advantage normalisation +
    #      5 mini-epochs x 4 minibatches x (obs RMS train forward + PPO loss fwd/bwd)
    from bez_isaacgym_b200 import learner as L
    n, T, mbs = 4096, 32, 32768
    r, v, d, lv, ld = sg.make_rollout(n, T, device=dev)
    obs = torch.randn(n * T, 54, device=dev)
    mb = {k: t.to(dev) for k, t in sg.make_minibatch(n * T).items()}
    obs_rms, val_rms = L.RunningMeanStd(54).to(dev), L.RunningMeanStd(1).to(dev)
    cfgp = L.PPOLossConfig()
    stats = torch.empty(8, dtype=torch.float64, device=dev)
    part = torch.empty(ops.ppo_scratch_doubles(), dtype=torch.float64, device=dev)
    kc = ops.make_ppo_cfg()
    gmu = torch.empty(mbs, 18, device=dev); gv = torch.empty(mbs, device=dev); gls = torch.empty(18, device=dev)
    norm_obs = torch.empty(mbs, 54, device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    vflat = torch.empty(n * T, 1, device=dev)

    def epoch_bezk():
        ops.gae(r, v, d, lv, ld, 0.99, 0.95, adv, ret)
        flat_ret, flat_val = L.swap_and_flatten01(ret), L.swap_and_flatten01(v)
        a_n = L.normalize_advantages(flat_ret, flat_val)
        val_rms(flat_val, out=vflat); val_rms(flat_ret, out=vflat)
        for _ in range(5):
            for i in range(n * T // mbs):
                sl = slice(i * mbs, (i + 1) * mbs)
                obs_rms(obs[sl], out=norm_obs)
                ops.ppo_loss(mb["actions"][sl], mb["mu"][sl], mb["logstd"], mb["old_mu"][sl], mb["old_sigma"][sl],
                             mb["values"].view(-1)[sl], mb["old_values"].view(-1)[sl], mb["returns"].view(-1)[sl],
                             mb["old_neglogp"][sl], a_n[sl], kc, stats, part, grad_mu=gmu, grad_values=gv, grad_logstd=gls)
    t = gpu_time(epoch_bezk, reps=1, iters=10)
    emit(what="ppo_epoch_math_4096x32", impl="bezk CUDA (graph replay)", us=t * 1e6)
    t = gpu_time(epoch_bezk, reps=1, iters=10, graph=False)
    emit(what="ppo_epoch_math_4096x32", impl="bezk CUDA (eager launches)", us=t * 1e6)

    o_obs, o_val = rg.RunningMeanStd(54), rg.RunningMeanStd(1)
    for o in (o_obs, o_val):
        o.running_mean, o.running_var, o.count = o.running_mean.to(dev), o.running_var.to(dev), o.count.to(dev)
    df, ldf = d.float(), ld.float()

    def epoch_ref():
        a = rg.discount_values(ldf, lv, df, v, r, 0.99, 0.95)
        rt = a + v
        flat_ret, flat_val = rg.swap_and_flatten01(rt), rg.swap_and_flatten01(v)
        a_n, _, _ = rg.prepare_dataset(flat_ret, flat_val, o_val)
        for _ in range(5):
            for i in range(n * T // mbs):
                sl = slice(i * mbs, (i + 1) * mbs)
                o_obs(obs[sl])
                mu = mb["mu"][sl].clone().requires_grad_(True); val = mb["values"][sl].clone().requires_grad_(True)
                ls = mb["logstd"].clone().requires_grad_(True)
                out = rg.ppo_loss(dict(mu=mu, values=val, logstd=ls, actions=mb["actions"][sl], old_mu=mb["old_mu"][sl],
                                       old_sigma=mb["old_sigma"][sl], old_values=mb["old_values"][sl], returns=mb["returns"][sl],
                                       old_neglogp=mb["old_neglogp"][sl], advantages=a_n[sl]))
                out["loss"].backward()
    t = gpu_time(epoch_ref, reps=1, iters=5, graph=False)
    emit(what="ppo_epoch_math_4096x32", impl="reference torch ops + autograd on the same B200 (eager)", us=t * 1e6)
    if args.out:
        with open(args.out, "w") as f:
            for row in rows:
                f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    main()
