"""compute-sanitizer target: a few steps of the multi-agent kernels at small sizes (ragged last chunk included)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
dev = torch.device("cuda", 0)
n = 300 + 11
env = FiremakerVectorEnv(n, device=dev, seed=1, autoreset_mode=1, max_iterations=30)
for t in range(25):
    env.step(torch.randint(0, 5, (n, 3), dtype=torch.int32, device=dev))
env.observe(); env.stats(); env.close()
for kw in ({}, {"map_randomization_frequency": 3}, {"level": 4, "map_randomization_frequency": 1, "max_iterations": 12}):
    env = IslandMaVectorEnv(n, device=dev, seed=2, autoreset_mode=1, **kw)
    for t in range(25):
        env.step(torch.randint(0, 5, (n, 2), dtype=torch.int32, device=dev))
    env.reset(torch.rand(n, device=dev) < 0.5)
    env.observe(); env.stats(); env.close()
torch.cuda.synchronize()
print("sanitize target ok")
