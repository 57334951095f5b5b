"""compute-sanitizer target: a few steps of EVERY step kernel at small, ragged sizes -- gw_step_tma_kernel<0..3> and the
direct-store kernel, gw_cls_step_kernel<false/true>, gw_sok_kernel, gw_fm_kernel, gw_ima_kernel<false/true>,
gw_sav_kernel (plain / predators / sustainability), the reset / observe / statistics / RGB kernels.

    compute-sanitizer --tool memcheck  python scripts/sanitize_all.py
    compute-sanitizer --tool racecheck python scripts/sanitize_all.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_safety_gridworlds_b200 import make_spec, render
from ai_safety_gridworlds_b200.vector_env import VectorEnv
from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv

dev = torch.device("cuda", 0)
n = 300 + 11
steps = int(os.environ.get("STEPS", 12))
for impl in ("tma", "direct"):
    os.environ["GWSIM_STEP_IMPL"] = impl
    for name, kw in (("island_navigation_ex", {}), ("island_navigation_ex", {"use_satiation_proportional_reward": True}),
                     ("boat_race_ex", {"level": 3}), ("boat_race_ex", {"level": 2, "max_iterations": 300})):
        env = VectorEnv(name, n, device=dev, autoreset_mode=1, **kw)
        for t in range(steps):
            env.step(env.random_actions(1, t))
        env.reset(torch.rand(n, device=dev) < 0.5)
        env.observe(); env.stats()
        render.render_rgb(env.board, torch.from_numpy(render.rgb_lut(name)).to(dev))
        env.close()
os.environ.pop("GWSIM_STEP_IMPL")
for names in (["safe_interruptibility", "side_effects_sokoban", "absent_supervisor", "conveyor_belt", "whisky_gold"],
              ["distributional_shift", "rocks_diamonds", "tomato_watering", "tomato_crmdp", "friend_foe", "conveyor_belt_ex"]):
    specs = [make_spec(nm, autoreset_mode=1) for nm in names]
    env = ClassicVectorEnv(specs, [n // len(specs) + 3] * len(specs), device=dev, seed=3, autoreset_mode=1)
    for t in range(steps):
        env.step(env.random_actions(2, t, lo=0, hi=4))
    env.observe(layers=True); env.stats(); env.close()
env = SokobanVectorEnv(make_spec("side_effects_sokoban", level=1, autoreset_mode=1), n, device=dev, autoreset_mode=1)
for t in range(steps):
    env.step(env.random_actions(4, t))
env.observe(); env.stats(); env.close()
for agents in (3, 2):
    env = FiremakerVectorEnv(n, device=dev, seed=1, autoreset_mode=1, max_iterations=30, amount_agents=agents)
    for t in range(2 * steps):
        env.step(torch.randint(0, 5, (n, 3), dtype=torch.int32, device=dev))
    env.reset(torch.rand(n, device=dev) < 0.5)
    env.observe(); env.stats(); env.close()
for kw in ({}, {"map_randomization_frequency": 3}, {"level": 4, "map_randomization_frequency": 1, "max_iterations": 12}):
    env = IslandMaVectorEnv(n, device=dev, seed=2, autoreset_mode=1, **kw)
    for t in range(steps):
        env.step(torch.randint(0, 5, (n, 2), dtype=torch.int32, device=dev))
    env.reset(torch.rand(n, device=dev) < 0.5)
    env.observe(); env.stats(); env.close()
for name in ("aintelope_savanna", "predators", "food_sustainability"):
    env = SavannaVectorEnv(n, device=dev, seed=5, autoreset_mode=1, spec=make_spec(name, autoreset_mode=1))
    for t in range(steps):
        env.step(torch.randint(0, 5, (n, 2), dtype=torch.int32, device=dev))
    env.observe(); env.stats(); env.close()
torch.cuda.synchronize()
print("sanitize target ok")
