"""B200-native batched Stratego engine (hot path of JBLanier/stratego_env, rebuilt as sm_100a CUDA).

Public names mirror the reference package (``stratego_env/__init__.py:1-2``): ``StrategoMultiAgentEnv``,
``ObservationComponents``, ``ObservationModes``, ``GameVersions``.  ``BatchedStrategoEnv`` is the batched,
device-resident variant.  Importing the environments needs torch; the CUDA extension is loaded (and
must exist) as soon as an environment or engine is constructed -- there is no CPU fallback.
"""
from .enums import (GameVersions, ObservationComponents, ObservationModes, RecentMoves, SP)  # noqa: F401

_LAZY = {
    "StrategoMultiAgentEnv": ".stratego_multiagent_env",
    "make_stratego_env": ".stratego_multiagent_env",
    "StrategoProceduralEnv": ".stratego_procedural_env",
    "BatchedStrategoEnv": ".batched_env",
    "StrategoEngine": ".engine",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(_LAZY[name], __name__), name)
    raise AttributeError(name)
