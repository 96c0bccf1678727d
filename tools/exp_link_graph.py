#!/usr/bin/env python
"""Is the per-call cost of the chunked copy pipeline (tools/exp_link_copies.py: ~40 us per extra H2D call under bidirectional
load) a submission cost that a CUDA graph removes?  The staged_pack post-physics phase's copies (per chunk: H2D 144 + 96 B/env on
one stream, then D2H 216 + 4 + 8 + 8 B/env on a second) issued eagerly vs replayed as ONE captured graph.  Median of 9."""
import json
import sys
import time

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda:0")
    torch.cuda.init()
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def bufs(widths, to_dev):
        out = []
        for w in widths:
            h = torch.empty(n, w, dtype=torch.uint8).pin_memory()
            d = torch.empty(n, w, dtype=torch.uint8, device=dev)
            out.append((d, h) if to_dev else (h, d))
        return out

    ins, outs = bufs([144, 96], True), bufs([216, 4, 8, 8], False)
    for chunks in (4, 8):
        step = -(-n // chunks)

        def issue(cur):
            s_in.wait_stream(cur)
            evs = []
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                with torch.cuda.stream(s_in):
                    for dst, src in ins:
                        dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev)
                    for dst, src in outs:
                        dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                evs.append(ev)
            cur.wait_stream(s_out)
            cur.wait_stream(s_in)

        def timed(fn):
            ts = []
            for _ in range(10):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            ts = sorted(ts[1:])
            return ts[len(ts) // 2]

        cur = torch.cuda.current_stream(dev)
        eager = timed(lambda: issue(cur))
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        with torch.cuda.stream(cap):
            with torch.cuda.graph(g, stream=cap):
                issue(cap)
        graph = timed(g.replay)
        print(json.dumps({"envs": n, "chunks": chunks, "eager_ms": round(eager, 3), "graph_ms": round(graph, 3)}), flush=True)


if __name__ == "__main__":
    main()
