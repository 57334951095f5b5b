import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = int(os.environ.get("N", 65536))
env = FiremakerVectorEnv(N, device=dev, seed=1, autoreset_mode=1)
acts = [torch.randint(0, 5, (N, 3), dtype=torch.int32, device=dev) for _ in range(8)]
for i in range(60): env.step_raw(_ptr(acts[i & 7]))
torch.cuda.synchronize()
print("ok", float(env.observe()["ext_fires"].float().mean()))
