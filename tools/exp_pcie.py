#!/usr/bin/env python
"""Which side of the zero-copy host pipeline costs what: the fused post-physics kernel at 65536 envs with the sparse
simulator tensors (rigid_body, net_contact), the dense ones (dof_state, root_states) and the outputs (obs, rew, timeout)
each placed either in HBM or in pinned host memory.  One JSON line per combination."""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bez_isaacgym_b200 import ops, synthetic_gym as sg  # noqa: E402


def main():
    n = 65536
    dev = torch.device("cuda:0")
    st = sg.make_state(n, seed=1, filler=False)
    goal, ball_init, default, lo, hi = sg.make_constants(n, dev)
    cfg = ops.make_task_cfg(reset_root_states=False, write_contact_filter=False)
    progress, reset = sg.make_bookkeeping(n, device=dev)
    prev = torch.zeros(n, 3, device=dev)
    place = lambda t, host: t.clone().pin_memory() if host else t.to(dev)      # noqa: E731
    for sparse_host, dense_host, out_host in itertools.product((0, 1), repeat=3):
        rb, cf = place(st.rigid_body, sparse_host), place(st.net_contact, sparse_host)
        dof, root = place(st.dof_state, dense_host), place(st.root_states, dense_host)
        obs = place(torch.empty(n, 54), out_host); rew = place(torch.empty(n), out_host)
        timeout = place(torch.empty(n, dtype=torch.long), out_host)

        def run():
            ops.post_physics(dof, rb, root, cf, goal, ball_init, None, reset, progress, timeout, cfg, obs, rew, prev_lin_vel=prev)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print(json.dumps({"sparse": "host" if sparse_host else "hbm", "dense": "host" if dense_host else "hbm",
                          "outputs": "host" if out_host else "hbm", "ms": round(ts[len(ts) // 2], 4)}), flush=True)


if __name__ == "__main__":
    main()
