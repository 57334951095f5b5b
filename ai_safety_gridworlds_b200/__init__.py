"""B200-native batched simulator for AI Safety Gridworlds (island_navigation_ex, boat_race_ex, ...).

The hot path (pycolab Engine.play + reward/termination accounting + observation rendering of the
reference) runs as hand-written CUDA for sm_100a in csrc/libgwsim.so behind the C ABI of
include/gwsim.h; this package is the Python host side that mirrors the reference's environment
API.  There is no CPU fallback.
"""
from . import _abi  # noqa: F401
from .envs import make_spec, ENVIRONMENTS  # noqa: F401

__all__ = ["make_spec", "ENVIRONMENTS", "VectorEnv", "ClassicVectorEnv", "FiremakerVectorEnv", "IslandMaVectorEnv", "GridworldGymEnv",
           "GridworldZooParallelEnv", "GridworldZooAecEnv", "SokobanVectorEnv", "get_environment_obj"]


def __getattr__(name):
    # torch is imported lazily so that spec compilation works without it
    if name == "VectorEnv":
        from .vector_env import VectorEnv
        return VectorEnv
    if name == "ClassicVectorEnv":
        from .classic_env import ClassicVectorEnv
        return ClassicVectorEnv
    if name == "FiremakerVectorEnv":
        from .firemaker_env import FiremakerVectorEnv
        return FiremakerVectorEnv
    if name == "IslandMaVectorEnv":
        from .island_ma_env import IslandMaVectorEnv
        return IslandMaVectorEnv
    if name == "GridworldZooParallelEnv":
        from .helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
        return GridworldZooParallelEnv
    if name == "GridworldZooAecEnv":
        from .helpers.gridworld_zoo_aec_env import GridworldZooAecEnv
        return GridworldZooAecEnv
    if name == "SokobanVectorEnv":
        from .sokoban_env import SokobanVectorEnv
        return SokobanVectorEnv
    if name == "get_environment_obj":
        from .helpers.factory import get_environment_obj
        return get_environment_obj
    if name == "GridworldGymEnv":
        from .helpers.gridworld_gym_env import GridworldGymEnv
        return GridworldGymEnv
    raise AttributeError(name)
