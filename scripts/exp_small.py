"""Small-batch regime of the fused step kernel (BASELINE config 2: 65,536 boat_race_ex environments): persistent TMA kernel vs
the direct-store variant (GWSIM_STEP_IMPL=direct), plain launches vs a CUDA graph of the same launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.vector_env import VectorEnv, _ptr
dev = torch.device("cuda", 0)
name, kw = os.environ.get("ENV", "boat_race_ex"), ({"level": 3} if os.environ.get("ENV", "boat_race_ex") == "boat_race_ex" else {})
for n in (1 << 16, 1 << 17, 1 << 18, 1 << 20):
    spec = make_spec(name, autoreset_mode=1, **kw)
    env = VectorEnv(spec, n, device=dev, autoreset_mode=1, want_value_board=False)
    ring = torch.empty((8, n), dtype=torch.int32, device=dev)
    for r in range(8):
        env.random_actions(seed=0, step=r, out=ring[r])
    ptrs = [_ptr(ring[r]) for r in range(8)]
    for t in range(50):
        env.step_raw(ptrs[t & 7])
    torch.cuda.synchronize()
    K = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        env.step_raw(ptrs[t & 7])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    # the same launches captured once in a CUDA graph and replayed
    s = torch.cuda.Stream(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        for t in range(8):
            env.step_raw(ptrs[t & 7])
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for t in range(40):
                env.step_raw(ptrs[t & 7])
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    msg = e0.elapsed_time(e1) / 400
    b = env.bytes_per_env_step() * n
    print("%s impl=%s n=%7d  launches %.2f us/step (%.0f GB/s)   graph %.2f us/step (%.0f GB/s)   working set %.0f MB" %
          (name, os.environ.get("GWSIM_STEP_IMPL", "tma"), n, ms * 1e3, b / ms / 1e6, msg * 1e3, b / msg / 1e6, b / 1e6))
    env.close()
