"""Import shims that let the UNMODIFIED reference (/root/reference) be imported in this container.

Test infrastructure only (used by tests/golden/make_golden.py, which records the golden traces committed
under tests/golden/).  Nothing in the product package imports this file, and nothing on the GPU box needs
/root/reference: the recorded ``.npz`` fixtures travel instead.

The reference pulls in ``gym``, ``skimage.measure.block_reduce`` and ``matplotlib`` at import time
(gym_dockauv/__init__.py:1, objects/sensor.py:5, envs/docking3d.py:11-12); none of them is installed here
and there is no network.  Only ``block_reduce`` touches hot-path arithmetic (sensor.py:137, func=np.max,
block_size=2, default cval=0); it is restated below as pad-with-cval-then-reduce, which is scikit-image's
documented behaviour (scikit-image ~=0.19.3, requirements.txt:35).
"""
import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("DOCKAUV_REFERENCE_ROOT", "/root/reference")


def _block_reduce(image, block_size=2, func=np.sum, cval=0, func_kwargs=None):
    image = np.asarray(image)
    if np.isscalar(block_size):
        block_size = (block_size,) * image.ndim
    pad = []
    for dim, b in zip(image.shape, block_size):
        rem = dim % b
        pad.append((0, (b - rem) if rem else 0))
    image = np.pad(image, pad, mode="constant", constant_values=cval)
    new_shape = []
    for dim, b in zip(image.shape, block_size):
        new_shape += [dim // b, b]
    blocked = image.reshape(new_shape)
    axes = tuple(range(1, 2 * image.ndim, 2))
    return func(blocked, axis=axes)


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.dtype = dtype
        self.shape = self.low.shape


class _Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass


def install():
    """Populate sys.modules with the shims and put the reference on sys.path (idempotent)."""
    if "gym_dockauv" in sys.modules:
        return
    gym = types.ModuleType("gym")
    gym.Env = _Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = _Box
    gym.spaces = spaces
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    gym.envs = envs
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    gym.utils = utils
    gym.make = None
    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": registration,
        "gym.utils": utils, "gym.utils.seeding": seeding,
    })
    skimage = types.ModuleType("skimage")
    measure = types.ModuleType("skimage.measure")
    measure.block_reduce = _block_reduce
    skimage.measure = measure
    sys.modules.update({"skimage": skimage, "skimage.measure": measure})
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
                 "matplotlib.widgets", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.figure",
                 "matplotlib.axes", "matplotlib.lines", "matplotlib.ticker",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
                 "mpl_toolkits.mplot3d.axes3d"]:
        sys.modules.setdefault(name, mock.MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_dockauv"))
