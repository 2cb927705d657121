"""predpreygrass_b200 — B200-native batched PredPreyGrass environment step."""
from .config import BASE_CONFIG, ECO_CONFIG, SEASONAL_CONFIG, STAG_CONFIG, make_config  # noqa: F401

_LAZY = {
    "BatchedPredPreyGrass": "batched",
    "PredPreyGrass": "env", "PredPreyGrassDenseRewards": "env", "PredPreyGrassDenseRewardsAdditive": "env",
    "PredPreyGrassSparseRewards": "env", "PredPreyGrassSparseRewardsPlusEating": "env",
    "PredPreyGrassSparseRewardsPlusKickback": "env", "PredPreyGrassSeasonal": "env",
    "PredPreyGrassEco": "env_evolutionary", "PredPreyGrassStag": "env_evolutionary",
}
__all__ = ["BASE_CONFIG", "ECO_CONFIG", "SEASONAL_CONFIG", "STAG_CONFIG", "make_config"] + sorted(_LAZY)


def __getattr__(name):  # torch / the CUDA library are loaded only when a class that needs them is asked for
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    raise AttributeError(name)
