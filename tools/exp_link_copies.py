#!/usr/bin/env python
"""Follow-up to exp_link_pipeline.py: is the cost of splitting a chunk's H2D bytes into two cudaMemcpyAsync calls a per-call cost
or a property of the source buffers?  All regions are carved from ONE pinned and ONE device allocation; the same 240 B/env go
H2D per chunk as (a) one call, (b) two calls on adjacent halves of the same region, (c) two calls from two separate regions
(144 + 96), each followed by one 236 B/env D2H call on a second stream.  The matrix is run twice.  Median of 9, wall clock."""
import json
import sys
import time

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    dev = torch.device("cuda:0")
    torch.cuda.init()
    s_in, s_out, s_in2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for trial in range(2):
        host = torch.empty(n * (240 + 240 + 236) + 4096, dtype=torch.uint8).pin_memory()
        devb = torch.empty(n * (240 + 240 + 236) + 4096, dtype=torch.uint8, device=dev)
        host.fill_(1)

        def region(buf, off, w):
            return buf[off:off + n * w].view(n, w)
        h_one, d_one = region(host, 0, 240), region(devb, 0, 240)
        h_a, d_a = region(host, n * 240, 144), region(devb, n * 240, 144)
        h_b, d_b = region(host, n * 384, 96), region(devb, n * 384, 96)
        h_out, d_out = region(host, n * 480, 236), region(devb, n * 480, 236)
        variants = {
            "one call 240": lambda lo, hi: [(d_one[lo:hi], h_one[lo:hi])],
            "two calls, adjacent halves (120+120 of the chunk)": lambda lo, hi: [
                (d_one[lo:(lo + hi) // 2], h_one[lo:(lo + hi) // 2]), (d_one[(lo + hi) // 2:hi], h_one[(lo + hi) // 2:hi])],
            "two calls, two regions (144 + 96)": lambda lo, hi: [(d_a[lo:hi], h_a[lo:hi]), (d_b[lo:hi], h_b[lo:hi])],
        }
        variants["two regions, two H2D streams"] = variants["two calls, two regions (144 + 96)"]
        for name, mk in variants.items():
            for chunks in (4, 8):
                step = -(-n // chunks)
                evs = [torch.cuda.Event() for _ in range(chunks)]
                evs2 = [torch.cuda.Event() for _ in range(chunks)]
                ts = []
                for _ in range(10):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for c, lo in enumerate(range(0, n, step)):
                        hi = min(n, lo + step)
                        if "two H2D streams" in name:
                            (d0, h0), (d1, h1) = mk(lo, hi)
                            with torch.cuda.stream(s_in2):
                                d1.copy_(h1, non_blocking=True)
                                evs2[c].record(s_in2)
                            with torch.cuda.stream(s_in):
                                d0.copy_(h0, non_blocking=True)
                                s_in.wait_event(evs2[c])
                                evs[c].record(s_in)
                        else:
                            with torch.cuda.stream(s_in):
                                for dst, src in mk(lo, hi):
                                    dst.copy_(src, non_blocking=True)
                                evs[c].record(s_in)
                        with torch.cuda.stream(s_out):
                            s_out.wait_event(evs[c])
                            h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
                    torch.cuda.synchronize()
                    ts.append((time.perf_counter() - t0) * 1e3)
                ts = sorted(ts[1:])
                print(json.dumps({"trial": trial, "h2d": name, "chunks": chunks, "ms": round(ts[len(ts) // 2], 3),
                                  "min_ms": round(ts[0], 3)}), flush=True)
        del host, devb


if __name__ == "__main__":
    main()
