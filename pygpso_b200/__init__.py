"""
pygpso_b200 -- B200-native (sm_100a) Gaussian-process surrogate hot path behind pyGPSO's public API.

    from pygpso_b200 import ParameterSpace, GPSOptimiser, GPRSurrogate

drops in for ``from gpso import ...`` of jajcayn/pygpso: the GPflow/TensorFlow arithmetic (hyper-parameter fit,
``predict_y``, UCB scoring of the ternary tree's leaf candidates) runs in hand-written CUDA through the C ABI declared
in ``include/gpso_b200.h``.  There is no CPU fallback: without the built library and a B200 the compute calls raise
``pygpso_b200.backend.GpsoBackendError``.
"""
from .gp_surrogate import GPListOfPoints, GPPoint, GPRSurrogate, GPSurrogate
from .optimisation import CallbackTypes, GPSOCallback, GPSOptimiser
from .param_space import LeafNode, ParameterSpace
from .utils import PointLabels, set_logger

__version__ = "0.1.0"
__all__ = [
    "CallbackTypes", "GPListOfPoints", "GPPoint", "GPRSurrogate", "GPSOCallback", "GPSOptimiser", "GPSurrogate",
    "LeafNode", "ParameterSpace", "PointLabels", "set_logger",
]
