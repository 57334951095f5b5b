"""A/B timing of alternative builds of libgwsim (GWSIM_LIB): headline island workload + boat."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200.vector_env import VectorEnv, _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 1 << 20))
for name, kw, wv in [("island_navigation_ex", {}, False), ("boat_race_ex", {"level": 3}, False)]:
    env = VectorEnv(name, N, device=dev, autoreset_mode=1, want_value_board=wv, **kw)
    if os.environ.get("PINGPONG"):
        env.state = torch.zeros((2 * env.state_words, N, 4), dtype=torch.int32, device=dev)
        env.reset()
    ring = [env.random_actions(0, r).clone() for r in range(8)]
    ptrs = [_ptr(r) for r in ring]
    for i in range(30): env.step_raw(ptrs[i & 7])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): env.step_raw(ptrs[i & 7])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200)
    b = env.bytes_per_env_step() * N
    print("%-10s %-22s %.1f us  %.0f GB/s  %.2fe9 steps/s" % (os.environ.get("TAG", ""), name, best * 1e3, b / best / 1e6, N / best / 1e6))
    env.close()
