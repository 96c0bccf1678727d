#!/usr/bin/env python
"""Host packer experiment (no GPU needed): time ``bezk_host_pack_begin`` / ``_wait`` gathering the sparse Isaac Gym rows of
``--envs`` envs, issued as ``--chunks`` jobs like one step of the ``staged_pack`` host pipeline does, after evicting the caches.
One JSON line per run: when each begin() returned and when each wait() returned (ms after the first begin).
    python tools/exp_host_pack.py --threads 15 --chunks 4 [--spin-us 2000] [--pin 0]"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bez_isaacgym_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262144)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--spin-us", type=int, default=-1)
    ap.add_argument("--pin", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--dof", type=int, default=1, help="copy the dense dof_state rows along (chunk layout [dof rows | records])")
    ap.add_argument("--gap-ms", type=float, default=3.0, help="idle time between two steps (the K0 phase + simulate())")
    args = ap.parse_args()
    lib = _lib.load()
    n, nb = args.envs, 22
    cfg = ops.make_task_cfg(num_bodies=nb)
    rb = np.random.default_rng(0).standard_normal((n, nb, 13), dtype=np.float32)
    cf = np.random.default_rng(1).standard_normal((n, nb, 3), dtype=np.float32)
    root = np.random.default_rng(2).standard_normal((n, 2, 13), dtype=np.float32)
    dof = np.random.default_rng(3).standard_normal((n, 36), dtype=np.float32)
    rs = lib.bezk_host_pack_record_floats(0, ctypes.byref(cfg))
    pw = rs + (36 if args.dof else 0)
    rec = np.zeros(n * pw, np.float32)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)              # noqa: E731
    workers = lib.bezk_host_pack_config(args.threads, args.spin_us, args.pin)
    junk = np.zeros((96 << 20) // 4, np.float32)
    c = -(-n // args.chunks)
    runs = []
    for _ in range(args.reps):
        junk += 1                                             # evict the caches: the simulator has just rewritten the tensors
        t_gap = time.perf_counter()
        while time.perf_counter() - t_gap < args.gap_ms * 1e-3:
            pass
        t0 = time.perf_counter()
        tickets, begins, waits = [], [], []
        for lo in range(0, n, c):
            tickets.append(lib.bezk_host_pack_begin(0, P(rb), P(cf), P(root), P(dof) if args.dof else None, None, ctypes.byref(cfg),
                                                    ctypes.c_void_p(rec.ctypes.data + 4 * lo * pw), lo, min(c, n - lo)))
            begins.append(round(1e3 * (time.perf_counter() - t0), 3))
        for t in tickets:
            lib.bezk_host_pack_wait(t)
            waits.append(round(1e3 * (time.perf_counter() - t0), 3))
        runs.append({"begin_ms": begins, "wait_ms": waits})
    k0 = min(c, n)
    first = rec[(k0 * 36 if args.dof else 0):k0 * pw].reshape(k0, rs)
    assert np.array_equal(first[:, :10], rb[:k0, cfg.imu_body, 3:13])
    best = min(r["wait_ms"][-1] for r in runs)
    print(json.dumps({"envs": n, "workers": workers, "chunks": args.chunks, "dof": args.dof, "spin_us": args.spin_us, "pin": args.pin,
                      "cpus": os.cpu_count(), "best_total_ms": best, "runs": runs[1:]}), flush=True)


if __name__ == "__main__":
    main()
