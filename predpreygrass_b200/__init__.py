"""predpreygrass_b200 — B200-native batched PredPreyGrass environment step."""
from .config import BASE_CONFIG, make_config  # noqa: F401

__all__ = ["BASE_CONFIG", "make_config", "BatchedPredPreyGrass", "PredPreyGrass"]


def __getattr__(name):
    if name == "BatchedPredPreyGrass":
        from .batched import BatchedPredPreyGrass

        return BatchedPredPreyGrass
    if name == "PredPreyGrass":
        from .env import PredPreyGrass

        return PredPreyGrass
    raise AttributeError(name)
