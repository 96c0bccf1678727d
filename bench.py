#!/usr/bin/env python
"""BezKick hot-path benchmark: task step (K0 + fused post-physics kernel) x horizon + one GAE scan.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference torch path on the host CPU

A "step" is one rollout segment of the hot path on one batch of synthetic Isaac-Gym-layout state:
``horizon`` (32) env steps through ``KickEnv.step`` (2 kernel launches each) followed by one GAE scan over the
(32, envs) rollout.  ``value`` = env-steps/s over all ranks with everything resident in HBM; ``e2e`` = the same
metric through ``KickEnv.step`` in host-pipeline mode (simulator tensors and actions in pinned HOST memory,
H2D/D2H copies inside the timed region).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "task+GAE env-steps/s"
UNIT = "env-steps/s"
TASK_BYTES_PER_ENV_STEP = 680          # K0 144 + post-physics 536 (prev_lin_vel buffer mode)
K0_BYTES = 144
POST_BYTES = 536
GAE_BYTES_PER_SAMPLE = 17


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--envs-per-gpu", type=int, default=262144)
    p.add_argument("--horizon", type=int, default=32)
    p.add_argument("--e2e-envs", type=int, default=65536, help="envs per GPU for the host-pipeline (e2e) leg")
    p.add_argument("--e2e-mode", default="zero_copy", choices=["zero_copy", "staged"])
    p.add_argument("--e2e-steps", type=int, default=5)
    p.add_argument("--cpu-sample-envs", type=int, default=65536)
    p.add_argument("--cpu-rollouts", type=int, default=24, help="rollouts of the bounded CPU-baseline sample (~10-30 s)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--fusion", default="fused", choices=["fused", "split"])
    p.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay of the step")
    p.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity hint (0 = leave default)")
    return p.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:                       # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:                   # noqa: BLE001
                pass
            time.sleep(0.004)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_rate(n, horizon, rollouts, threads=None):
    """The reference torch path on the host CPU (oracle port of the reference's own functions, op for op):
    `rollouts` x (horizon task steps + one GAE scan) at n envs.  Returns (env-steps/s, seconds, threads)."""
    from bez_isaacgym_b200 import synthetic_gym as sg
    from oracle import rl_games_oracle as rg
    from oracle import task_oracle as to
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    st = sg.make_state(n, seed=1234, filler=False)
    goal, ball_init, default, lower, upper = sg.make_constants(n)
    init_root = torch.zeros(n * 2, 13)
    orc = to.KickStepOracle(n, st.root_states, st.dof_state, st.rigid_body, st.net_contact, default, lower, upper,
                            goal, ball_init, torch.tensor([0.0, 0.0]), init_root.clone(), alias_prev_lin_vel=False)
    orc.initial_root_states = st.root_states.clone()      # a real simulator owns the root reset: rows are unchanged
    orc.prev_lin_vel = torch.zeros(n, 3)
    progress, reset = sg.make_bookkeeping(n, seed=99)
    orc.progress_buf[:] = progress
    orc.reset_buf[:] = reset
    actions = sg.make_actions(n)
    rewards, values, dones, last_values, last_dones = sg.make_rollout(n, horizon, seed=7)

    def rollout():
        for _ in range(horizon):
            orc.pre_physics_step(actions)
            orc.post_physics_step()
        adv = rg.discount_values(last_dones.float(), last_values, dones.float(), values, rewards, 0.99, 0.95)
        return adv + values

    orc.pre_physics_step(actions); orc.post_physics_step()         # warm-up (jit / allocator)
    t0 = time.perf_counter()
    for _ in range(rollouts):
        rollout()
    dt = time.perf_counter() - t0
    return n * horizon * rollouts / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_sample_envs
    rate, secs, threads = cpu_reference_rate(n, args.horizon, 1)              # warm-up rollout
    t0 = time.perf_counter()
    rate, secs, threads = cpu_reference_rate(n, args.horizon, max(1, args.steps))
    sample = f"{max(1, args.steps)} rollouts of {args.horizon} task steps + GAE at {n} envs on the host CPU"
    out = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"bez_kick task step x{args.horizon} + GAE, reference torch ops on CPU (oracle port)",
                      "envs": n, "horizon": args.horizon},
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: this image exports NCCL_DEBUG=VERSION, which prints a banner on rank 0
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    ge.build()
    from bez_isaacgym_b200 import bez_model as bm, ops, synthetic_gym as sg
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks.kick_env import KickEnv

    n, T = args.envs_per_gpu, args.horizon
    l2_fetch = ops.set_l2_fetch_granularity(args.l2_fetch) if args.l2_fetch else None
    cfg = bm.default_task_cfg(n, rl_device=str(dev))
    cfg["env"]["imuPrevVelAliasing"] = False                # the general (680 B/env-step) path with a prev_lin_vel buffer
    cfg["seed"] = 42 + rank

    class OwnedRootSim(SyntheticGym):
        owns_root_reset = True                              # as with Isaac Gym: the simulator restores root states

    sim = OwnedRootSim(n, device=str(dev), seed=1234 + rank, filler=True)
    env = KickEnv(cfg, str(dev), 0, True, sim=sim, fusion=args.fusion)
    progress, reset = sg.make_bookkeeping(n, seed=99 + rank, device=dev)
    env.progress_buf.copy_(progress); env.reset_buf.copy_(reset)
    actions = sg.make_actions(n, seed=4321 + rank, device=dev)
    rewards, values, dones, last_values, last_dones = sg.make_rollout(n, T, seed=7 + rank, device=dev)
    advs, rets = torch.empty_like(rewards), torch.empty_like(rewards)
    stream = torch.cuda.current_stream(dev)
    launches_per_step = T * (2 if args.fusion == "fused" else 3) + 1

    post_events = []

    def bench_step(record):
        for _ in range(T):
            env.pre_physics_step(actions)
            env.sim.simulate()
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                env.post_physics_step()
                e1.record(stream)
                post_events.append((e0, e1))
            else:
                env.post_physics_step()
        ops.gae(rewards, values, dones, last_values, last_dones, 0.99, 0.95, advs, rets)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        bench_step(False)
    barrier()
    graph = None
    if not args.no_graph:
        # the step is 65 dependent launches of 7..45 us kernels: replaying it as ONE CUDA graph removes the CPU-side
        # launch cost and most of the inter-kernel gaps (the kernels, arguments and work are identical)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            bench_step(False)
        stream = torch.cuda.current_stream(dev)
        graph.replay()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            bench_step(True)
    end.record(stream)
    barrier()
    ms = start.elapsed_time(end)
    # Duration of the dominant kernel.  Eager timed region: per-launch CUDA events recorded inside it.  Graph-replayed timed
    # region (default): events cannot be read back from inside a replayed graph, so right after it (a) the same steps run
    # eagerly with per-launch events (includes the event-record gaps) and (b) a graph holding only the T post-physics
    # launches of one step is replayed between two events (back-to-back launch duration); (b) is what `achieved` uses.
    eager_ms = None
    post_graph_ms = None
    if graph is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(min(args.steps, 10)):
            bench_step(True)
        e1.record(stream)
        barrier()
        eager_ms = e0.elapsed_time(e1) / min(args.steps, 10)
        pg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(pg):
            for _ in range(T):
                env.post_physics_step()
        pg.replay()
        barrier()
        e0.record(stream)
        for _ in range(min(args.steps, 10)):
            pg.replay()
        e1.record(stream)
        barrier()
        post_graph_ms = e0.elapsed_time(e1) / (min(args.steps, 10) * T)
    sampler.stop_flag = True
    sampler.join()
    reset_rate = float(env.reset_buf.float().mean())
    post_events_ms = sum(a.elapsed_time(b) for a, b in post_events) / len(post_events)
    post_ms = post_graph_ms if post_graph_ms is not None else post_events_ms
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    env_steps = n * world * T * args.steps
    value = env_steps / (ms * 1e-3)

    peak, peak_src = peaks()
    achieved = POST_BYTES * n / (post_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "bezk::task_tile_kernel<7> (fused post-physics)" if args.fusion == "fused"
                else "bezk::task_tile_kernel<3>+<4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_env": POST_BYTES, "envs_per_launch": n,
                "avg_launch_ms": post_ms, "peak_source": peak_src,
                "avg_launch_ms_eager_events": post_events_ms,
                "timing": "CUDA events around a replayed graph of the step's 32 post-physics launches, taken right after the "
                          "graph-replayed timed region (avg_launch_ms); per-launch events of an eager pass in avg_launch_ms_eager_events"
                if graph is not None else "per-launch CUDA events inside the (eager) timed region",
                "whole_step_gbs": (TASK_BYTES_PER_ENV_STEP * n * T + GAE_BYTES_PER_SAMPLE * n * T) * args.steps / (ms * 1e-3) / 1e9}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            roofline["traffic"] = json.load(f).get("post_physics_bytes_per_launch")

    # ---- e2e: host pipeline (simulator tensors + actions in pinned host memory) through KickEnv.step ----
    e2e = None
    if not args.no_e2e:
        ne = args.e2e_envs
        hcfg = bm.default_task_cfg(ne, use_gpu_pipeline=False, rl_device="cpu")
        hcfg["env"]["imuPrevVelAliasing"] = False
        hcfg["env"]["hostPipeline"] = args.e2e_mode
        hsim = OwnedRootSim(ne, device=str(dev), seed=1234 + rank, host=True, filler=True)
        henv = KickEnv(hcfg, f"cuda:{local}", 0, True, sim=hsim, fusion=args.fusion)
        hact = sg.make_actions(ne, seed=1).pin_memory()
        hr, hv, hd, hlv, hld = [t.pin_memory() for t in sg.make_rollout(ne, T, seed=3)]
        d_r, d_v, d_d, d_lv, d_ld = [torch.empty_like(t, device=dev) for t in (hr, hv, hd, hlv, hld)]
        d_adv, d_ret = torch.empty_like(d_r), torch.empty_like(d_r)
        h_adv, h_ret = torch.empty_like(hr).pin_memory(), torch.empty_like(hr).pin_memory()

        def e2e_step():
            for _ in range(T):
                henv.step(hact)
            for d, h in ((d_r, hr), (d_v, hv), (d_d, hd), (d_lv, hlv), (d_ld, hld)):
                d.copy_(h, non_blocking=True)
            ops.gae(d_r, d_v, d_d, d_lv, d_ld, 0.99, 0.95, d_adv, d_ret)
            h_adv.copy_(d_adv, non_blocking=True); h_ret.copy_(d_ret, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nb = hsim.num_bodies
        if args.e2e_mode == "staged":
            h2d_env = 4 * (26 + 36 + 13 * nb + 3 * nb + 18)       # all four simulator tensors + actions copied
            d2h_env = 4 * 54 + 4 + 8 + 8 + 4 * 36 + 4 * 18        # obs, rew, reset, timeouts, dof_state back, targets
        else:
            # zero-copy: the kernels gather over PCIe -- dense dof_state + root_states + actions, and the sparse
            # rows at the 64 B fetch granularity (IMU link ~96 B, two feet ~144 B, measured: profiles/r01_fetch_granularity.md)
            h2d_env = 4 * (26 + 36 + 18) + 96 + 144
            d2h_env = 4 * 54 + 4 + 8 + 8 + 4 * 18                 # obs, rew, reset, timeouts, targets (+ rare reset rows)
        e2e = {"value": ne * world * T * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": (h2d_env * T + 9 * T + 5) * ne, "d2h_bytes_per_step": (d2h_env * T + 8 * T) * ne,
               "envs_per_gpu": ne, "host_pipeline": args.e2e_mode,
               "note": "KickEnv.step with use_gpu_pipeline=False (simulator tensors, actions, targets in pinned HOST memory; "
                       "obs/rew/reset/timeouts returned on the host every env step); rollout H2D + adv/returns D2H per GAE"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn = args.cpu_sample_envs
        rate, secs, threads = cpu_reference_rate(cn, T, args.cpu_rollouts)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_rollouts} rollouts of {T} task steps + GAE at {cn} envs ({secs:.1f} s), reference torch ops "
                                  f"(oracle port) with torch.set_num_threads({threads})"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": f"bez_kick {n} envs/GPU: {T} x (K0 pre-physics + fused post-physics) + 1 GAE scan "
                                      f"(BASELINE configs[3] shard size; configs[1] is the 4096-env case)",
                          "envs_per_gpu": n, "horizon": T, "parallelism": f"env-sharded x{world}, no data-path collective",
                          "fusion": args.fusion, "l2": "inputs larger than L2 (state footprint ~%.0f MB/GPU)" % (
                              n * 4 * (26 + 36 + 16 * bm.BODIES_NO_CLEATS) / 1e6),
                          "launch": "cuda_graph_replay" if graph is not None else "eager", "ms_per_step_eager_with_events": eager_ms,
                          "reset_rate_per_step": reset_rate, "l2_fetch_granularity": l2_fetch, "imu_prev_lin_vel": "buffer (680 B/env-step path)"},
               "clocks": sampler.summary(), "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
               "e2e": e2e, "cpu_baseline": cpu_baseline}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
