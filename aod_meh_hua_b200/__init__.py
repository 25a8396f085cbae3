"""B200-native MEH alpha -> Dirichlet uncertainty -> HUA -> pool top-k scoring path.

Drop-in for the active-learning pool-scoring path of MoonLab-YH/AOD_MEH_HUA.
The computation lives in libmehhua.so (hand-written sm_100a CUDA behind a C ABI, include/mehhua.h);
this package is the Python host side mirroring the reference's head / train-loop interface.
"""
from .specs import (AGG_AVG, AGG_MAX, AGG_SUM, HEAD_RETINA, HEAD_SSD, SPECS, DetectorSpec,  # noqa: F401
                    ScoringParams, get_spec, parse_agg_spec)

__version__ = "0.1.0"
