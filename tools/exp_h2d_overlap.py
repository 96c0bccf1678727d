#!/usr/bin/env python
"""What overlaps on the host link: copy-engine transfers of one BezKick step's host traffic at n envs,
alone and concurrently on separate streams.
  dense   H2D  dof_state 144 + root_states 104 + actions 72 B/env      (cudaMemcpyAsync)
  sparse  H2D  IMU slice 40 B of 1144, two foot rows 12 B of 264       (cudaMemcpy2DAsync)
  out     D2H  obs 216 + rew 4 + reset 8 + timeouts 8 + targets 72 B/env
Each case is timed as wall clock around (launch all, synchronize), median of 7.  Tool only (cuda-python)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuda.bindings import runtime as rt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda:0")
    torch.cuda.init()
    pin = lambda *s: torch.randn(*s).pin_memory()                  # noqa: E731
    rb, cf, dof, root, act = pin(n, 286), pin(n, 66), pin(n, 36), pin(n, 26), pin(n, 18)
    d_dof, d_root, d_act = (torch.empty(n, w, device=dev) for w in (36, 26, 18))
    d_imu, d_fl, d_fr = torch.empty(n, 10, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    d_obs, d_rew, d_tgt = torch.randn(n, 54, device=dev), torch.randn(n, device=dev), torch.randn(n, 18, device=dev)
    d_reset, d_to = torch.zeros(n, dtype=torch.long, device=dev), torch.zeros(n, dtype=torch.long, device=dev)
    h_obs, h_rew, h_tgt = torch.empty(n, 54).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, 18).pin_memory()
    h_reset, h_to = torch.empty(n, dtype=torch.long).pin_memory(), torch.empty(n, dtype=torch.long).pin_memory()
    s_dense, s_sparse, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    H2D = rt.cudaMemcpyKind.cudaMemcpyHostToDevice

    def rows(lo, hi):
        return slice(lo, hi)

    def dense(lo, hi):
        with torch.cuda.stream(s_dense):
            for d, h in ((d_dof, dof), (d_root, root), (d_act, act)):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)

    def sparse(lo, hi, stream=None):
        st = (stream or s_sparse).cuda_stream
        k = hi - lo
        for dst, dp, src, off, sp, w in ((d_imu, 40, rb, (13 + 3) * 4, 1144, 40), (d_fl, 12, cf, 36 * 4, 264, 12),
                                        (d_fr, 12, cf, 60 * 4, 264, 12)):
            (err,) = rt.cudaMemcpy2DAsync(dst.data_ptr() + lo * dp, dp, src.data_ptr() + lo * sp + off, sp, w, k, H2D, st)
            assert err == rt.cudaError_t.cudaSuccess, err

    def out(lo, hi):
        with torch.cuda.stream(s_out):
            for h, d in ((h_obs, d_obs), (h_rew, d_rew), (h_tgt, d_tgt), (h_reset, d_reset), (h_to, d_to)):
                h[lo:hi].copy_(d[lo:hi], non_blocking=True)

    def chunked(fns):
        step = (n + chunks - 1) // chunks
        for lo in range(0, n, step):
            for f in fns:
                f(lo, min(n, lo + step))

    cases = {
        "dense_h2d": [dense], "sparse_h2d": [sparse], "out_d2h": [out],
        "dense+sparse (2 streams)": [dense, sparse],
        "dense+sparse (1 stream)": [dense, lambda lo, hi: sparse(lo, hi, s_dense)],
        "dense+out": [dense, out], "sparse+out": [sparse, out],
        "dense+sparse+out (3 streams)": [dense, sparse, out],
    }
    for name, fns in cases.items():
        ts = []
        for _ in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            chunked(fns)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = sorted(ts[1:])
        ms = ts[len(ts) // 2]
        print(json.dumps({"case": name, "envs": n, "chunks": chunks, "ms": round(ms, 3), "ns_per_env": round(ms * 1e6 / n, 2),
                          "env_steps_per_s_M": round(n / ms / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
