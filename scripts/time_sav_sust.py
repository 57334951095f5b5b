"""Device-timed parallel steps of the sustainability-challenge instantiation of gw_sav_kernel (food_sustainability experiment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 1 << 17))
for name, kw in (("food_sustainability", {}), ("aintelope_savanna", dict(observation_direction_mode=2, action_direction_mode=2))):
    spec = make_spec(name, autoreset_mode=1, **kw)
    env = SavannaVectorEnv(N, device=dev, seed=1, autoreset_mode=1, spec=spec)
    hi = 9 if kw else 5
    acts = [torch.randint(0, hi, (N, 2), dtype=torch.int32, device=dev) for _ in range(8)]
    for i in range(50): env.step_raw(_ptr(acts[i & 7]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300): env.step_raw(_ptr(acts[i & 7]))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 300
    tiles = float((env.board == ord("F")).sum(dim=(1, 2)).float().mean())
    print("%s %s: %.3f ms per %d-environment parallel step, %.3g steps/s, %.0f B/step -> %.0f GB/s; mean F tiles on the board %.2f" % (
        name, kw, ms, N, N / ms * 1e3, env.bytes_per_env_step(), env.bytes_per_env_step() * N / ms / 1e6, tiles))
    env.close()
