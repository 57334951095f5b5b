import sys, os
sys.path.insert(0, "/root/repo")
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
from ai_safety_gridworlds_b200.vector_env import _ptr
dev = torch.device("cuda", 0)
N = 1 << 16
for dm in (0, 1):
    env = FiremakerVectorEnv(N, device=dev, seed=1, autoreset_mode=1, spec=make_spec("firemaker_ex_ma", autoreset_mode=1, amount_agents=3, observation_direction_mode=dm, action_direction_mode=dm))
    acts = [torch.randint(0, 5, (N, 3), dtype=torch.int32, device=dev) for _ in range(8)]
    for i in range(150): env.step_raw(_ptr(acts[i & 7]))
    torch.cuda.synchronize()
    F = (env.board == ord("F")).sum(dim=(1, 2)).float()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): env.step_raw(_ptr(acts[i & 7]))
    e1.record(); torch.cuda.synchronize()
    print("dm", dm, "mean fires", float(F.mean()), "burning games", float((F > 0).float().mean()), "mean fires of burning", float(F[F > 0].mean()), "ms/step (65536 games)", e0.elapsed_time(e1) / 20)
    env.close()
