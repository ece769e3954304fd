"""Import-path alias: ``from pdb2reaction_b200.uma_pysis import uma_pysis, CALC_KW`` mirrors
``from pdb2reaction.uma_pysis import ...`` in the reference's callers (e.g. ``path_opt.py``,
``freq.py``, ``tsopt.py``).  The implementation is in ``calculator.py``."""
from .calculator import *  # noqa: F401,F403
from .calculator import uma_pysis, UMAcore, CALC_KW, GEOM_KW_DEFAULT, EV2AU, F_EVAA_2_AU, H_EVAA_2_AU, run_pysis  # noqa: F401
