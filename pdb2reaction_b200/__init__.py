"""pdb2reaction_b200 -- B200-native backend for pdb2reaction's ``uma_pysis`` calculator.

Only the hot path lives here: the calculator surface (``calculator.py``), the ctypes
binding of the C-ABI CUDA library (``engine.py``, ``csrc/``), weight preparation / MoLE merge
(``weights.py``), image sharding (``sharding.py``) and synthetic inputs (``synth.py``).
"""
from .calculator import (uma_pysis, UMAcore, CALC_KW, GEOM_KW_DEFAULT, EV2AU, F_EVAA_2_AU,  # noqa: F401
                         H_EVAA_2_AU, run_pysis)
from .arch import UMAArch  # noqa: F401

__all__ = ["uma_pysis", "UMAcore", "CALC_KW", "GEOM_KW_DEFAULT", "EV2AU", "F_EVAA_2_AU", "H_EVAA_2_AU",
           "run_pysis", "UMAArch"]
