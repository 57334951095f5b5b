"""Environment registry: name -> spec compiler (the counterpart of helpers/factory.py:100-201)."""
from . import aintelope_savanna, boat_race_ex, classic, experiments, firemaker_ex_ma, island_navigation_ex, island_navigation_ex_ma, savanna_experiments

ENVIRONMENTS = {
    island_navigation_ex.NAME: island_navigation_ex.compile_spec,
    boat_race_ex.NAME: boat_race_ex.compile_spec,
}
ENVIRONMENTS.update(classic.COMPILERS)
ENVIRONMENTS[firemaker_ex_ma.NAME] = firemaker_ex_ma.compile_spec
ENVIRONMENTS[island_navigation_ex_ma.NAME] = island_navigation_ex_ma.compile_spec
ENVIRONMENTS[aintelope_savanna.NAME] = aintelope_savanna.compile_spec
ENVIRONMENTS.update(experiments.COMPILERS)
ENVIRONMENTS.update(savanna_experiments.COMPILERS)


def make_spec(env_name, **kwargs):
    try:
        compiler = ENVIRONMENTS[env_name.lower()]
    except KeyError:
        raise NotImplementedError("The requested environment is not available.")  # factory.py:199-201
    return compiler(**kwargs)
