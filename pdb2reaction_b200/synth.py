"""Deterministic synthetic cluster models and strings for BASELINE.json's configs.

The reference gets its ~300/500/1500-atom cluster models from ``extract.py`` on real PDB
files; none are shipped, so benches and parity tests use protein-like random clusters:
N atoms in a sphere at 0.10 atoms/A^3, minimum pair distance 0.95 A (H-X) / 1.2 A (X-X),
elements H/C/N/O/S = 50/31/8/10/1 %.
"""
from __future__ import annotations

import numpy as np

_ELEMS = np.array(["H", "C", "N", "O", "S"])
_PROBS = np.array([0.50, 0.31, 0.08, 0.10, 0.01])

CONFIGS = {
    "C1": dict(n_atoms=20, n_images=1, seed=1),
    "C2": dict(n_atoms=300, n_images=12, seed=2),
    "C3": dict(n_atoms=500, n_images=1, seed=3),
    "C4": dict(n_atoms=1500, n_images=32, seed=4),
    "C5": dict(n_atoms=10000, n_images=8, seed=5),
}


def make_cluster(n_atoms: int, seed: int, density: float = 0.10):
    """-> (elem list[str], coords [N,3] float64 Angstrom)."""
    rng = np.random.default_rng(seed)
    radius = (3.0 * n_atoms / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    elem = rng.choice(_ELEMS, size=n_atoms, p=_PROBS)
    if n_atoms <= 20:
        elem = np.where(elem == "S", "C", elem)
    is_h = elem == "H"
    cell = 1.2
    grid = {}
    coords = np.zeros((n_atoms, 3))
    placed = 0
    tries = 0
    while placed < n_atoms:
        tries += 1
        if tries > 2000 * n_atoms:
            raise RuntimeError("cluster generator failed to place atoms")
        p = rng.uniform(-radius, radius, size=3)
        if p @ p > radius * radius:
            continue
        key = tuple(np.floor(p / cell).astype(int))
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for q in grid.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                        dmin = 0.95 if (is_h[placed] or is_h[q]) else 1.2
                        d = coords[q] - p
                        if d @ d < dmin * dmin:
                            ok = False
                            break
                    if not ok:
                        break
                if not ok:
                    break
            if not ok:
                break
        if not ok:
            continue
        coords[placed] = p
        grid.setdefault(key, []).append(placed)
        placed += 1
    return [str(e) for e in elem], coords


def make_string(n_atoms: int, n_images: int, seed: int, n_moved: int = 12, jitter: float = 0.02):
    """Images = linear interpolation between a base cluster and a copy with ``n_moved`` atoms
    displaced by U(0.3, 1.5) A, plus N(0, jitter) noise per image.
    -> (elem, coords [n_images, N, 3] float64 Angstrom)."""
    elem, base = make_cluster(n_atoms, seed)
    rng = np.random.default_rng(seed + 1000)
    end = base.copy()
    moved = rng.choice(n_atoms, size=min(n_moved, n_atoms), replace=False)
    direction = rng.normal(size=(len(moved), 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    end[moved] += direction * rng.uniform(0.3, 1.5, size=(len(moved), 1))
    imgs = np.zeros((n_images, n_atoms, 3))
    for k in range(n_images):
        t = 0.0 if n_images == 1 else k / (n_images - 1)
        imgs[k] = (1 - t) * base + t * end + rng.normal(scale=jitter, size=base.shape)
    return elem, imgs


def make_config(name: str):
    cfg = CONFIGS[name]
    return make_string(cfg["n_atoms"], cfg["n_images"], cfg["seed"])
