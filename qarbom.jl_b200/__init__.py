"""qarbom_b200: B200-native classical RBM training hot path of NITeQ/QARBoM.jl behind a C ABI.

The directory is called ``qarbom.jl_b200`` (not importable by name); ``import qarbom_b200``
works through the loader module at the repository root.
"""
from . import _ffi, build as _build  # noqa: F401
from ._ffi import QbError, lib  # noqa: F401
from .device import DeviceRBM  # noqa: F401
from .rbm import GRBM, RBM, AbstractRBM, GRBMClassifier, RBMClassifier, copy_rbm  # noqa: F401
from .train import CD, PCD, FastPCD, Accuracy, CrossEntropy, FreeEnergy, MeanSquaredError, reconstruct, train  # noqa: F401

build = _build.build
