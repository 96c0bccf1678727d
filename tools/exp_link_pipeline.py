#!/usr/bin/env python
"""Host-link experiment behind the staged_pack pipeline: the post-physics phase as a chunked copy pipeline WITHOUT kernels --
per chunk [s_in] H2D of the chunk's inputs -> event -> [s_out] D2H of its results -- for different ways of splitting the same
bytes (240 B/env in, 236 B/env out) into cudaMemcpyAsync calls and different chunk counts.  Wall clock around (issue all,
synchronize), median of 9.  Answers: what does a copy call cost under bidirectional load, and how many chunks pay?
    python tools/exp_link_pipeline.py [envs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda:0")
    torch.cuda.init()
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def bufs(widths, to_dev):
        # one (n, w) byte tensor pair per width; H2D: pinned -> device, D2H: device -> pinned
        out = []
        for w in widths:
            h = torch.empty(n, w, dtype=torch.uint8).pin_memory()
            d = torch.empty(n, w, dtype=torch.uint8, device=dev)
            out.append((d, h) if to_dev else (h, d))
        return out

    splits = {
        "in 144+96, out 216+4+8+8 (staged_pack today)": ([144, 96], [216, 4, 8, 8]),
        "in 144+96, out 216+20": ([144, 96], [216, 20]),
        "in 144+96, out 236": ([144, 96], [236]),
        "in 240, out 236": ([240], [236]),
        "in 144+96, out 216+4+1": ([144, 96], [216, 4, 1]),
        "in 240 only": ([240], []),
        "out 236 only": ([], [236]),
    }
    for name, (win, wout) in splits.items():
        ins, outs = bufs(win, True), bufs(wout, False)
        for chunks in (1, 2, 4, 8, 16):
            step = -(-n // chunks)
            evs = [torch.cuda.Event() for _ in range(chunks)]
            ts = []
            for _ in range(10):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for c, lo in enumerate(range(0, n, step)):
                    hi = min(n, lo + step)
                    with torch.cuda.stream(s_in):
                        for dst, src in ins:
                            dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                        evs[c].record(s_in)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(evs[c])
                        for dst, src in outs:
                            dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                t_issue = time.perf_counter()
                torch.cuda.synchronize()
                ts.append(((time.perf_counter() - t0) * 1e3, (t_issue - t0) * 1e3))
            ts = sorted(ts[1:])
            ms, issue = ts[len(ts) // 2]
            gb = n * (sum(win) + sum(wout)) / 1e9
            print(json.dumps({"split": name, "envs": n, "chunks": chunks, "ms": round(ms, 3), "issue_ms": round(issue, 3),
                              "combined_GBps": round(gb / (ms * 1e-3), 1)}), flush=True)
        del ins, outs


if __name__ == "__main__":
    main()
