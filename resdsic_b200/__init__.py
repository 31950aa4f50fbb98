"""resdsic_b200 -- B200-native (sm_100a) forward path of the ResDSIC / WACNN learned image codec.

Python host (nn.Module API identical to the reference's) over a plain-C-ABI CUDA
library (include/resdsic_b200.h).  No CPU / ATen / Triton fallback: everything on
the hot path is a hand-written kernel, and a missing library raises.
"""
from . import _lib
from .models import WACNN, SymmetricalTransFormer, configure_model, models

__version__ = "0.1.0"
__all__ = ["models", "configure_model", "WACNN", "SymmetricalTransFormer", "_lib"]
