"""B200-native batched humanoid rollout (reset / step / obs / rewards / auto-reset / GAE) behind a C-ABI.

Importing the package does not touch CUDA; `HumanoidBatch`, `B200HumanoidVecEnv` and `HumanoidEnv` load
libb2h.so on construction and fail loudly without it or without a GPU (there is no CPU path).
"""
from .mjcf import compile_mjcf  # noqa: F401

__all__ = ["compile_mjcf", "HumanoidBatch", "B200HumanoidVecEnv", "HumanoidEnv", "gae"]


def __getattr__(name):  # lazy: keeps `import mujocoposelearning_b200` torch-free
    if name in ("HumanoidBatch", "gae"):
        from . import batch
        return getattr(batch, name)
    if name in ("B200HumanoidVecEnv", "HumanoidEnv"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
