"""B200-native batched Stratego engine (hot path of JBLanier/stratego_env, rebuilt as sm_100a CUDA)."""
from .enums import (GameVersions, ObservationComponents, ObservationModes, RecentMoves, SP)  # noqa: F401
