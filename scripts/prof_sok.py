"""ncu target: a few launches of the side_effects_sokoban big-map kernel (gw_sok_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 1 << 20))
env = SokobanVectorEnv(make_spec("side_effects_sokoban", level=1), N, device=dev, autoreset_mode=1, want_value_board=False)
acts = [torch.randint(1, 5, (N,), dtype=torch.int32, device=dev) for _ in range(8)]
for i in range(30): env.step_raw(_ptr(acts[i & 7]))
torch.cuda.synchronize()
print("ok", env.stats()["episodes"])
