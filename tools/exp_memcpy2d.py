#!/usr/bin/env python
"""A/B for the host pipeline: pull the sparse Isaac Gym rows out of PINNED HOST memory with the copy
engines (cudaMemcpy2DAsync, strided source -> dense device staging) instead of 64-byte zero-copy reads issued by the SMs.

Measures, at n envs, the wall-clock-free device time (CUDA events) of
  * IMU slice   : src pitch 1144 B (22 bodies x 13 floats), width 40 B            -> dst pitch 40
  * feet, 2 rows: src pitch  264 B (22 bodies x 3 floats),  width 12 B, twice     -> dst pitch 12
  * feet, 1 span: src pitch  264 B, width 108 B (left foot .. right foot)        -> dst pitch 108
  * dense       : plain cudaMemcpyAsync of dof_state (144 B/env) for the link's own rate
One JSON line per case.  Tool only (uses cuda-python); the product's copy path lives in libbezk.so."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuda.bindings import runtime as rt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda:0")
    torch.cuda.init()
    rb = torch.randn(n, 22 * 13).pin_memory()
    cf = torch.randn(n, 22 * 3).pin_memory()
    dof = torch.randn(n, 36).pin_memory()
    stream = torch.cuda.current_stream(dev)
    H2D = rt.cudaMemcpyKind.cudaMemcpyHostToDevice

    def copy2d(dst, dpitch, src_ptr, spitch, width, rows):
        (err,) = rt.cudaMemcpy2DAsync(dst.data_ptr(), dpitch, src_ptr, spitch, width, rows, H2D, stream.cuda_stream)
        assert err == rt.cudaError_t.cudaSuccess, err

    d_imu = torch.empty(n, 10, device=dev)
    d_fl, d_fr = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    d_span = torch.empty(n, 27, device=dev)
    d_dof = torch.empty(n, 36, device=dev)
    d_imu64 = torch.empty(n, 16, device=dev)
    cases = {
        "imu_40B_of_1144": (lambda: copy2d(d_imu, 40, rb.data_ptr() + (13 + 3) * 4, 1144, 40, n), 40),
        "imu_64B_of_1144": (lambda: copy2d(d_imu64, 64, rb.data_ptr() + (13 + 3) * 4, 1144, 64, n), 64),
        "feet_2x12B_of_264": (lambda: (copy2d(d_fl, 12, cf.data_ptr() + 12 * 3 * 4, 264, 12, n),
                                       copy2d(d_fr, 12, cf.data_ptr() + 20 * 3 * 4, 264, 12, n)), 24),
        "feet_span_108B_of_264": (lambda: copy2d(d_span, 108, cf.data_ptr() + 12 * 3 * 4, 264, 108, n), 108),
        "dense_dof_144B": (lambda: d_dof.copy_(dof, non_blocking=True), 144),
        "whole_net_contact_264B": (lambda: torch.empty(n, 66, device=dev).copy_(cf, non_blocking=True), 264),
    }
    for name, (fn, nbytes) in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(json.dumps({"case": name, "envs": n, "ms": round(ms, 4), "payload_gbs": round(nbytes * n / ms / 1e6, 2),
                          "ns_per_env": round(ms * 1e6 / n, 2)}), flush=True)
    assert torch.equal(d_imu.cpu(), rb.view(n, 22, 13)[:, 1, 3:13])
    assert torch.equal(d_span.cpu()[:, 0:3], cf.view(n, 22, 3)[:, 12]) and torch.equal(d_span.cpu()[:, 24:27], cf.view(n, 22, 3)[:, 20])


if __name__ == "__main__":
    main()
