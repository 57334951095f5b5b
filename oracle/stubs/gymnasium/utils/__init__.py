from . import seeding, env_checker  # noqa: F401
