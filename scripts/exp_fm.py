"""Timing of the firemaker_ex_ma kernel (tuning experiment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import _abi
if os.environ.get("GWSIM_LIB"):
    _abi.LIB_PATH = os.environ["GWSIM_LIB"]          # A/B against another build of the library
from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 262144))
for cube, crops, lcrops in [(True, True, True), (False, False, False)]:
    env = FiremakerVectorEnv(N, device=dev, seed=1, autoreset_mode=1, want_cube=cube, want_crops=crops, want_layer_crops=lcrops)
    acts = [torch.randint(0, 5, (N, 3), dtype=torch.int32, device=dev) for _ in range(8)]
    ptrs = [_ptr(a) for a in acts]
    for phase, nsteps in (("early (few fires)", 30), ("later", 100), ("late", 200)):
        for i in range(nsteps): env.step_raw(ptrs[i & 7])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): env.step_raw(ptrs[i & 7])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ex = env.observe()
        b = env.bytes_per_env_step() * N
        print("cube=%d crops=%d lcrops=%d %-18s %.2f ms/step  %.2fe6 parallel steps/s  %.0f GB/s  mean external fires %.1f" %
              (cube, crops, lcrops, phase, ms, N / ms / 1e3, b / ms / 1e6, float(ex["ext_fires"].float().mean())))
    env.close()
