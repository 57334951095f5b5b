"""Environment factory: name -> environment object, the counterpart of the reference's helpers/factory.py:100-201.

`get_environment_obj(name, **kwargs)` returns the reference-style single environment (`helpers/safety_env.py`: reset / step ->
TimeStep) for every single-agent game the CUDA backend serves, under the names the reference registers (the module name of the
environment or experiment); an unknown name raises NotImplementedError like factory.py:199-201.  The multi-agent games are
served through their PettingZoo wrappers (`GridworldZooParallelEnv`, `GridworldZooAecEnv`)."""
from . import safety_env
from ..envs import experiments

_environment_classes = dict(safety_env.ENVIRONMENT_CLASSES)
for _name in experiments.OVERLAYS:          # experiments/*.py: IslandNavigationEnvironmentExExperiment subclasses, one per module
    _environment_classes[_name] = type("IslandNavigationEnvironmentExExperiment", (safety_env.SafetyEnvironment,),
                                       {"ENV_NAME": _name, "__doc__": "experiments/%s.py on the CUDA backend." % _name})


def environment_names():
    return sorted(_environment_classes)


def get_environment_obj(name, *args, **kwargs):
    """Instantiate an environment by name (factory.py:184-201)."""
    environment_class = _environment_classes.get(name.lower(), None)
    if environment_class:
        return environment_class(*args, **kwargs)
    raise NotImplementedError("The requested environment is not available.")
