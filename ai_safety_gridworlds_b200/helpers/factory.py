"""Environment factory: name -> environment object, the counterpart of the reference's helpers/factory.py:100-277.

`get_environment_obj(name, **kwargs)` returns the reference-style environment for EVERY game the CUDA backend serves, under the
names the reference registers (the module name of the environment or experiment): the single-agent games as
`helpers/safety_env.SafetyEnvironment` objects (reset / step -> TimeStep), the multi-agent games (firemaker_ex_ma,
island_navigation_ex_ma, aintelope_savanna and the aintelope experiment overlays) as `SafetyEnvironmentMa` objects whose
TimeStep fields are dicts keyed by the agent character.  An unknown name raises NotImplementedError like factory.py:199-201.
`register_with_gym()` registers the Gym ids of factory.py:204-274 when gymnasium / gym is importable."""
from . import safety_env
from ..envs import experiments, savanna_experiments

_package_name = __name__.rsplit(".helpers", 1)[0]
_environment_classes = dict(safety_env.ENVIRONMENT_CLASSES)
_environment_classes.update(safety_env.MA_ENVIRONMENT_CLASSES)
for _name in experiments.OVERLAYS:          # experiments/*.py: IslandNavigationEnvironmentExExperiment subclasses, one per module
    _environment_classes[_name] = type("IslandNavigationEnvironmentExExperiment", (safety_env.SafetyEnvironment,),
                                       {"ENV_NAME": _name, "__doc__": "experiments/%s.py on the CUDA backend." % _name})
for _name in savanna_experiments.OVERLAYS:  # experiments/aintelope/*.py: AIntelopeSavannaEnvironmentMaExperiment subclasses
    _environment_classes[_name] = type("AIntelopeSavannaEnvironmentMaExperiment", (safety_env.SafetyEnvironmentMa,),
                                       {"ENV_NAME": _name, "__doc__": "experiments/aintelope/%s.py on the CUDA backend." % _name})


def environment_names():
    return sorted(_environment_classes)


def get_environment_obj(name, *args, **kwargs):
    """Instantiate an environment by name (factory.py:184-201)."""
    environment_class = _environment_classes.get(name.lower(), None)
    if environment_class:
        return environment_class(*args, **kwargs)
    raise NotImplementedError("The requested environment is not available.")


def to_gym_id(env_name):
    """factory.py:226-241: the camel-cased id prefix ('island_navigation_ex' -> 'IslandNavigationEx')"""
    result, next_upper = [], True
    for char in env_name:
        if next_upper:
            result.append(char.upper())
            next_upper = False
        elif char == ".":
            result.append(char)
            next_upper = True
        elif char == "_":
            next_upper = True
        else:
            result.append(char)
    return "".join(result)


def gym_registrations():
    """[(id, entry_point, kwargs)] exactly as factory.py:248-272 registers them (both naming conventions; the conveyor belt
    variants get their own ids) -- for the single-agent games, which are what GridworldGymEnv serves."""
    entry = _package_name + ".helpers.gridworld_gym_env:GridworldGymEnv"
    out = []
    for env_name in sorted(_environment_classes):
        if issubclass(_environment_classes[env_name], safety_env.SafetyEnvironmentMa):
            continue
        prefix = to_gym_id(str(env_name))
        if prefix == "ConveyorBelt":
            for variant in ["vase", "sushi", "sushi_goal", "sushi_goal2"]:
                out.append((to_gym_id(variant) + "-v0", entry, {"env_name": env_name, "variant": variant}))
        else:
            out.append((prefix + "-v0", entry, {"env_name": env_name}))
        out.append(("ai_safety_gridworlds." + env_name + "-v0", entry, {"env_name": env_name}))
    return out


_register_with_gym_done = False


def register_with_gym():
    """factory.py:204-274.  Needs gymnasium or gym (neither is part of this image: ImportError then)."""
    global _register_with_gym_done
    if _register_with_gym_done:
        return
    try:
        from gymnasium.envs.registration import register
    except ImportError:
        from gym.envs.registration import register
    _register_with_gym_done = True
    for gym_id, entry, kwargs in gym_registrations():
        register(id=gym_id, entry_point=entry, kwargs=kwargs)
