"""Minimal offline stand-in for `pettingzoo` (TEST INFRASTRUCTURE ONLY); see gymnasium stub."""


class ParallelEnv(object):
    pass


class AECEnv(object):
    pass
