#!/usr/bin/env python
"""BASELINE configs[2]/[3]: one full PPO epoch of BezKick on N GPUs (one process per GPU, envs sharded), every piece of
the hot path on the libbezk kernels and every exchange the path has on NCCL:

    rollout   T x [ obs RMS (eval) -> policy MLP (torch) -> policy_head kernel (sample, neglogp, value un-norm,
                    experience slots, PD targets) -> simulate (synthetic no-op) -> fused post-physics kernel writing
                    the next obses slot -> reward shaping ]
    GAE       one scan kernel over the (T, N) rollout
    dataset   value RMS x2 (moments -> ALL-REDUCE -> merge -> normalise), advantage moments -> ALL-REDUCE -> normalise
    learn     mini_epochs x minibatches x [ obs RMS train on the SLAB view (moments -> ALL-REDUCE -> merge -> normalise),
                    MLP forward (torch), fused PPO loss fwd+bwd kernel on slab views, MLP backward (torch),
                    gradient ALL-REDUCE on one flat bucket, clip, Adam ]

The policy MLP GEMMs / Adam stay in torch (north_star).  Prints ONE JSON line (rank 0): env-steps/s for the whole epoch,
per-phase device milliseconds (max over ranks), and the collective count / bytes.

    python tools/epoch_bench.py --envs-per-gpu 4096                                  # configs[2]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
           tools/epoch_bench.py --envs-per-gpu 262144                                # configs[3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch import nn  # noqa: E402


class Policy(nn.Module):
    """cfg/train/bez_kickPPO.yaml:10-32: MLP 54-400-200-100 (ELU), mu head 18, value head 1, fixed sigma parameter."""

    def __init__(self):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(54, 400), nn.ELU(), nn.Linear(400, 200), nn.ELU(), nn.Linear(200, 100), nn.ELU())
        self.mu = nn.Linear(100, 18)
        self.value = nn.Linear(100, 1)
        self.sigma = nn.Parameter(torch.zeros(18))

    def forward(self, x):
        h = self.body(x)
        return self.mu(h), self.value(h)


class Phases:
    def __init__(self, stream):
        self.stream, self.acc, self.open = stream, {}, None

    def start(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        self.open = (name, e)

    def stop(self):
        name, e0 = self.open
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(self.stream)
        self.acc.setdefault(name, []).append((e0, e1))

    def totals(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.acc.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=262144)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--minibatches", type=int, default=4, help="minibatches per mini-epoch (the yaml's 131072 / 32768)")
    ap.add_argument("--mini-epochs", type=int, default=5)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--fp32", action="store_true", help="MLP in fp32 instead of bf16 autocast")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    import __graft_entry__ as ge
    ge.build()
    from bez_isaacgym_b200 import bez_model as bm, dist as bdist, learner as L
    from bez_isaacgym_b200.learner import experience as ex
    from bez_isaacgym_b200.synthetic_sim import SyntheticGym
    from bez_isaacgym_b200.tasks.kick_env import KickEnv

    N, T = args.envs_per_gpu, args.horizon
    mb = N * T // args.minibatches
    cfg = bm.default_task_cfg(N, rl_device=str(dev))
    cfg["seed"] = 42 + rank

    class OwnedRootSim(SyntheticGym):
        owns_root_reset = True

    env = KickEnv(cfg, str(dev), 0, True, sim=OwnedRootSim(N, device=str(dev), seed=1234 + rank, filler=False))
    buf = ex.ExperienceBuffer(dict(observation_space=env.observation_space, action_space=env.action_space),
                              dict(num_actors=N, horizon_length=T), dev)
    torch.manual_seed(0)
    model = Policy().to(dev)
    bdist.broadcast_parameters(model, 0, group)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4, eps=1e-8)
    obs_rms = L.RunningMeanStd(54, process_group=group).to(dev)
    val_rms = L.RunningMeanStd(1, process_group=group).to(dev)
    rewards = torch.empty(T, N, 1, device=dev)
    returns = torch.empty(T, N, 1, device=dev)
    advs = torch.empty(T, N, 1, device=dev)
    adv_n = torch.empty(T, N, device=dev)
    vals_n = torch.empty(T, N, 1, device=dev)
    rets_n = torch.empty(T, N, 1, device=dev)
    nparams = sum(p.numel() for p in model.parameters())
    bucket = torch.empty(nparams, device=dev)
    amp = dict(device_type="cuda", dtype=torch.bfloat16, enabled=not args.fp32)
    stream = torch.cuda.current_stream(dev)
    obs = env.reset()["obs"].clone()
    dones = torch.zeros(N, dtype=torch.uint8, device=dev)
    collectives = {"count": 0, "bytes": 0}
    frame = [0]

    def epoch(ph):
        nonlocal obs, dones
        # ------------------------------------------------ rollout
        ph.start("rollout")
        obs_rms.eval(); val_rms.eval()
        buf.slot("obses", 0).copy_(obs)
        env.set_obs_target(buf.slot("obses", 0))
        for t in range(T):
            buf.update_data("dones", t, dones)
            with torch.no_grad(), torch.autocast(**amp):
                mu, v = model(obs_rms(buf.slot("obses", t)))
            res = L.policy_head(mu.float(), model.sigma, v.float(), val_rms, experience=buf, t=t, seed=cfg["seed"], step=frame[0], env=env)
            frame[0] += 1
            if t + 1 < T:
                env.set_obs_target(buf.slot("obses", t + 1))          # the step kernel writes the next slot in place
            else:
                env.set_obs_target(obs)
            env.set_rollout_targets(values=res["values"], shaped_rewards=rewards[t], dones_u8=dones, gamma=0.99, scale_value=0.01)
            o, rew, done, info = env.step_precomputed_targets(res["env_actions"])
        with torch.no_grad(), torch.autocast(**amp):
            _, last_v = model(obs_rms(obs))
        last_values = val_rms(last_v.float(), unnorm=True)
        ph.stop()
        # ------------------------------------------------ GAE
        ph.start("gae")
        L.discount_values(dones, last_values, buf.tensor_dict["dones"], buf.tensor_dict["values"], rewards, 0.99, 0.95,
                          out_advs=advs, out_returns=returns)
        ph.stop()
        # ------------------------------------------------ dataset statistics (prepare_dataset)
        ph.start("dataset")
        obs_rms.train(); val_rms.train()
        L.normalize_advantages(returns, buf.tensor_dict["values"], process_group=group, out=adv_n.view(-1))
        val_rms(buf.tensor_dict["values"], out=vals_n)               # two train-mode updates per epoch (ckpt: 1 + 2*frame)
        val_rms(returns, out=rets_n)
        collectives["count"] += 3; collectives["bytes"] += 8 * (3 + 3 + 3)
        ph.stop()
        # ------------------------------------------------ mini-epochs
        sds = ex.SlabDataset(buf, mb, extra=dict(returns=rets_n, advantages=adv_n, old_values=vals_n))
        for _ in range(args.mini_epochs):
            for i in range(len(sds)):
                s = sds[i]
                ph.start("learn_kernels")
                x = obs_rms(s["obses"])                              # train mode: slab moments -> all-reduce -> merge -> normalise
                ph.stop()
                ph.start("mlp_fwd")
                with torch.autocast(**amp):
                    mu, v = model(x)
                mu32, v32 = mu.float(), v.float()
                ph.stop()
                ph.start("learn_kernels")
                loss, info = L.ppo_loss(mu32, v32, model.sigma, s["actions"], s["mus"], s["sigmas"], s["old_values"], s["returns"],
                                        s["neglogpacs"], s["advantages"])
                ph.stop()
                ph.start("mlp_bwd_adam")
                opt.zero_grad(set_to_none=False)
                loss.backward()
                ph.stop()
                ph.start("grad_allreduce")
                bdist.allreduce_grads_(list(model.parameters()), group, True, bucket)
                ph.stop()
                ph.start("mlp_bwd_adam")
                nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()
                ph.stop()
                collectives["count"] += 2; collectives["bytes"] += 8 * 109 + 4 * nparams

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        epoch(Phases(stream))
    barrier()
    collectives["count"] = collectives["bytes"] = 0
    ph = Phases(stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.epochs):
        epoch(ph)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    tot = ph.totals()
    vec = torch.tensor([ms] + [tot[k] for k in sorted(tot)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    vec = vec.tolist()
    if rank == 0:
        per_epoch = {k: vec[1 + i] / args.epochs for i, k in enumerate(sorted(tot))}
        out = {"what": "full PPO epoch (BASELINE configs[2]/[3])", "n_gpus": world, "envs_per_gpu": N, "horizon": T,
               "minibatch": mb, "mini_epochs": args.mini_epochs, "epochs": args.epochs,
               "ms_per_epoch": vec[0] / args.epochs, "env_steps_per_s": N * world * T * args.epochs / (vec[0] * 1e-3),
               "phase_ms_per_epoch": per_epoch, "mlp_dtype": "fp32" if args.fp32 else "bf16 autocast",
               "collectives_per_epoch": collectives["count"] / args.epochs if world > 1 else 0,
               "collective_bytes_per_epoch": collectives["bytes"] / args.epochs if world > 1 else 0,
               "final_obs_count": float(obs_rms.count), "expected_obs_count": 1.0 + (args.epochs + args.warmup) * args.mini_epochs * N * T * world}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
