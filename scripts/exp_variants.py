"""Kernel-tuning experiment (not part of the product): time the step kernel with subsets of the
output tensors, next to plain fill / copy bandwidth on the same GPU."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200.vector_env import VectorEnv, _ptr

dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 1 << 20))
K = 200

def timeit(fn, k=K, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

# plain bandwidth references
buf = torch.empty(670_000_000, dtype=torch.uint8, device=dev)
buf2 = torch.empty_like(buf)
ms = timeit(lambda: buf.fill_(1), 50, 5); print("fill 670MB      %.1f us  %.0f GB/s (write only)" % (ms*1e3, 670e6/ms/1e6))
ms = timeit(lambda: buf2.copy_(buf), 50, 5); print("copy 670MB      %.1f us  %.0f GB/s (r+w)" % (ms*1e3, 2*670e6/ms/1e6))
s8 = buf.view(torch.int64)
ms = timeit(lambda: s8.sum(), 50, 5); print("read 670MB      %.1f us  %.0f GB/s (read only)" % (ms*1e3, 670e6/ms/1e6))
del buf, buf2, s8

for name, kw in [("island_navigation_ex", {}), ("boat_race_ex", {"level": 3})]:
    for wb, wc, wv in [(1,1,0),(0,0,0),(1,0,0),(0,1,0),(0,0,1),(1,1,1)]:
        env = VectorEnv(name, N, device=dev, autoreset_mode=1, want_board=bool(wb), want_cube=bool(wc), want_value_board=bool(wv), **kw)
        ring = [env.random_actions(0, r).clone() for r in range(8)]
        ptrs = [_ptr(r) for r in ring]
        i = [0]
        def step():
            env.step_raw(ptrs[i[0] & 7]); i[0] += 1
        ms = timeit(step)
        b = env.bytes_per_env_step() * N
        print("%-22s board=%d cube=%d value=%d  %.1f us  %4d B/env  %.0f GB/s  %.2fe9 steps/s" % (name, wb, wc, wv, ms*1e3, env.bytes_per_env_step(), b/ms/1e6, N/ms/1e6))
        env.close()
