"""What the copy part of the host-buffer pipeline can do on this box: N chunks go host -> device on one stream, each comes
back device -> host on a second stream as soon as it has arrived (no kernels).  Prints ms for the whole round trip and the
aggregate GB/s, for several chunk counts and total sizes, next to one-directional copies of the same bytes.
    python profiles/probe_copy_pipeline.py
"""
import json
import sys
import time

import torch

dev = torch.device("cuda:0")
for total_mb in (109, 218):
    n = total_mb * (1 << 20) // 8
    host_in = torch.empty(n, dtype=torch.float64).pin_memory()
    host_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for chunks in (1, 8, 16, 32):
        edges = [n * i // chunks for i in range(chunks + 1)]
        best = None
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            evs = []
            with torch.cuda.stream(s_in):
                for c in range(chunks):
                    d[edges[c]:edges[c + 1]].copy_(host_in[edges[c]:edges[c + 1]], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record()
                    evs.append(e)
            with torch.cuda.stream(s_out):
                for c in range(chunks):
                    s_out.wait_event(evs[c])
                    host_out[edges[c]:edges[c + 1]].copy_(d[edges[c]:edges[c + 1]], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best):
                best = dt
        print(json.dumps({"total_MB_each_way": total_mb, "chunks": chunks, "round_trip_ms": round(best * 1e3, 3),
                          "aggregate_GBps": round(2 * n * 8 / best / 1e9, 1)}), flush=True)
