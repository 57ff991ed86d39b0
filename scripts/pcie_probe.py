#!/usr/bin/env python
"""What bounds bench.py's e2e line: pinned-memory copy bandwidth of one step's blobs (3.6 MB each way), alone and
both directions at once, and the host time of one iteration of the e2e loop's stream bookkeeping."""
import time
import torch
dev = torch.device("cuda:0")
for mb in (3.6, 64.0):
    n = int(mb * 1e6 / 4)
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    d_in = torch.empty(n, dtype=torch.float32, device=dev)
    d_out = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = 50

    def timed(fn):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            for _ in range(reps):
                d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for _ in range(reps):
                h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
        t = timed(fn)
        print("%5.1f MB %-5s %7.1f us per copy  %6.1f GB/s per direction" % (mb, name, t * 1e6, mb * 1e-3 / t))

# host cost of the e2e loop's bookkeeping (no GPU work behind it)
s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ev = [torch.cuda.Event() for _ in range(8)]
for e in ev:
    e.record(s_cmp)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev[0]); ev[1].record(s_in)
    with torch.cuda.stream(s_cmp):
        s_cmp.wait_event(ev[1]); s_cmp.wait_event(ev[2]); ev[0].record(s_cmp); ev[3].record(s_cmp)
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev[3]); ev[2].record(s_out)
    ev[2].synchronize()
print("e2e loop bookkeeping: %.1f us of host time per iteration" % ((time.perf_counter() - t0) / 200 * 1e6))
