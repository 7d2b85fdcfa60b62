"""Drop-in replacement for the reference package ``wildlifemapper/segment_anything`` (same import names, ctor
kwargs, forward signatures and ``state_dict`` keys), backed by the sm_100a kernels of
``wildlifemapper_b200``.  Put the directory that CONTAINS this package first on ``sys.path`` (the reference
scripts import ``segment_anything`` from their working directory, train.py:19-31).
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:  # make ``wildlifemapper_b200`` importable when only the package dir is on the path
    sys.path.insert(0, _ROOT)

from .build_sam import (  # noqa: E402,F401
    build_sam,
    build_sam_vit_h,
    build_sam_vit_l,
    build_sam_vit_b,
    sam_model_registry,
)
from .predictor import SamPredictor  # noqa: E402,F401
