"""ncu target: a few launches of the aintelope_savanna kernel (gw_sav_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 1 << 17))
env = SavannaVectorEnv(N, device=dev, seed=1, autoreset_mode=1)
acts = [torch.randint(0, 5, (N, 2), dtype=torch.int32, device=dev) for _ in range(8)]
for i in range(30): env.step_raw(_ptr(acts[i & 7]))
torch.cuda.synchronize()
print("ok", float(env.observe()["frame"].float().mean()))
