"""feta_tmlr_b200 -- B200-native (sm_100a) drop-in for FeTA's spectral hot path.

Module layout mirrors the reference's ``transformer/`` package for the path only
(SURVEY.md section 8): ``ChebNetDynamic.ChebConvDynamic``, ``layers.DiffTransformerEncoderLayer``,
``models.DiffTransformerEncoderGenGCN`` + the three model heads the BASELINE configs use,
``data`` (collate / GPU batch builder) and ``utils``.  All computation on the path goes through
the C ABI of ``libfeta_b200.so`` (``include/feta_b200.h``); there is no CPU fallback.
"""
from . import utils  # noqa: F401
from .ChebNetDynamic import ARMAConvDynamic, ChebConvDynamic  # noqa: F401
from .layers import DiffTransformerEncoderLayer  # noqa: F401
from .models import (DiffTransformerEncoderGenGCN, DiffGraphTransformerGenGCN,  # noqa: F401
                     DiffGraphTransformerGenGCNSBM, DiffGraphTransformerGenGCNMolHiv, GlobalAvg1D)

__version__ = "0.1.0"
