"""gym_dockauv_b200 -- B200-native batched simulator for the step path of the gym_dockauv docking envs.

Host side in Python (this package), arithmetic in hand-written sm_100a CUDA kernels behind a C ABI
(include/dockauv.h, gym_dockauv_b200/csrc/).  There is no CPU fallback.
"""
from .config import BASE_CONFIG, MANUAL_CONFIG, PREDICT_CONFIG, RADAR_64, REGISTRATION_DICT, TRAIN_CONFIG  # noqa: F401

__all__ = ["BASE_CONFIG", "TRAIN_CONFIG", "PREDICT_CONFIG", "MANUAL_CONFIG", "RADAR_64", "REGISTRATION_DICT",
           "envs", "make_gym"]


def __getattr__(name):   # torch is only imported when an env is actually needed
    if name in ("envs", "vec_env", "stats"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if name == "make_gym":
        from .envs import make_gym
        return make_gym
    raise AttributeError(name)
