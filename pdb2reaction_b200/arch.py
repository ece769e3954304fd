"""uma-s-1p1 architecture constants (eSCN-MD backbone + MoLE), shared by host code.

The reference never states these itself: it loads them with the checkpoint via
``pretrained_mlip.get_predict_unit(model, device=...)`` (reference
``pdb2reaction/uma_pysis.py:246-250``) and only reads ``backbone.max_neighbors`` and
``backbone.cutoff`` back (``uma_pysis.py:301-302``).  The values below restate the
published fairchem-core ``uma-s-1p1`` ("K4L2") configuration.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class UMAArch:
    sphere_channels: int = 128      # C
    hidden_channels: int = 128      # H
    edge_channels: int = 128        # Ce
    lmax: int = 2                   # == mmax, 9 coefficients
    num_layers: int = 4
    num_distance_basis: int = 64    # B, Gaussian smearing
    cutoff: float = 6.0             # Angstrom
    max_neighbors: int = 300
    max_num_elements: int = 100
    num_experts: int = 32
    num_datasets: int = 5
    edge_degree_rescale: float = 5.0
    norm_eps: float = 1e-5
    envelope_exponent: int = 5

    @property
    def n_coeff(self) -> int:
        return (self.lmax + 1) ** 2

    @property
    def x_edge_dim(self) -> int:
        return self.num_distance_basis + 2 * self.edge_channels

    def as_dict(self):
        return asdict(self)


DATASET_LIST = ("oc20", "omol", "omat", "odac", "omc")

# l-primary index (l*l + l + m) of every row of the m-primary layout
# [m=0: l=0,1,2 | m=+1: l=1,2 | m=-1: l=1,2 | m=+2: l=2 | m=-2: l=2]
M_PRIMARY_TO_L_PRIMARY = (0, 2, 6, 3, 7, 1, 5, 8, 4)
# l of every l-primary coefficient
L_OF_COEFF = (0, 1, 1, 1, 2, 2, 2, 2, 2)

SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn "
    "Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm"
).split()
Z_OF_SYMBOL = {s: i for i, s in enumerate(SYMBOLS)}


def atomic_numbers(elem) -> list:
    """Element symbols (any case, as the reference accepts: ``uma_pysis.py:266``) -> Z."""
    out = []
    for e in elem:
        s = str(e).capitalize()
        if s not in Z_OF_SYMBOL or s == "X":
            raise ValueError(f"unknown element symbol {e!r}")
        out.append(Z_OF_SYMBOL[s])
    return out
