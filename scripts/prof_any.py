"""ncu / timing target for the secondary kernels: WHICH = fm | fm_dm1 | fm_dm2 | sav | sav_pred | sav_sust | ima | ima_randmap, N environments, STEPS launches.
With TIME=1 the launches are timed with CUDA events (never under ncu) and one line is printed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec
from ai_safety_gridworlds_b200.vector_env import _ptr

dev = torch.device("cuda", 0)
which = os.environ.get("WHICH", "fm")
steps = int(os.environ.get("STEPS", 30))
hi = 5
if which.startswith("fm"):
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    N = int(os.environ.get("N", 1 << 18))
    dm = int(which[-1]) if which.startswith("fm_dm") else 0
    hi = 9 if dm == 2 else 5
    env = FiremakerVectorEnv(N, device=dev, seed=1, autoreset_mode=1, spec=make_spec("firemaker_ex_ma", autoreset_mode=1, amount_agents=3,
                                                                                   observation_direction_mode=dm, action_direction_mode=dm))
    na = 3
elif which in ("sav", "sav_sust", "sav_pred"):
    from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
    N = int(os.environ.get("N", 1 << 17))
    spec = make_spec("food_sustainability" if which == "sav_sust" else "aintelope_savanna", autoreset_mode=1,
                     **(dict(amount_predators=4, amount_agents=2) if which == "sav_pred" else {}))
    env = SavannaVectorEnv(N, device=dev, seed=1, autoreset_mode=1, spec=spec)
    na = 2
else:
    from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
    N = int(os.environ.get("N", 1 << 20))
    kw = dict(map_randomization_frequency=3) if which == "ima_randmap" else {}
    env = IslandMaVectorEnv(N, device=dev, seed=1, autoreset_mode=1, spec=make_spec("island_navigation_ex_ma", autoreset_mode=1, **kw))
    na = 2
acts = [torch.randint(0, hi, (N, na), dtype=torch.int32, device=dev) for _ in range(8)]
for i in range(steps):
    env.step_raw(_ptr(acts[i & 7]))
torch.cuda.synchronize()
if os.environ.get("TIME"):
    reps = int(os.environ.get("REPS", 200))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        env.step_raw(_ptr(acts[i & 7]))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b = env.bytes_per_env_step()
    print("%s: %.4f ms per %d-environment parallel step, %.4g steps/s, %d B/step -> %.0f GB/s = %.1f%% of 6553.6" % (
        which, ms, N, N / ms * 1e3, b, b * N / ms / 1e6, b * N / ms / 1e6 / 65.536))
else:
    print("ok", which, N)
