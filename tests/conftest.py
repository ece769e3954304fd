import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def arch4():
    from pdb2reaction_b200.arch import UMAArch
    return UMAArch(num_experts=4)


@pytest.fixture(scope="session")
def state4(arch4):
    """Un-merged random-init weights, 4 experts (fast); same architecture otherwise."""
    from pdb2reaction_b200 import weights as W
    return W.init_uma_weights(arch4, seed=0)


@pytest.fixture(scope="session")
def hyper4():
    from oracle import uma_ref
    return uma_ref.Hyper(num_experts=4)


def merged_for(state, arch, elem, charge=0, spin=1, task="omol"):
    from pdb2reaction_b200 import weights as W
    from pdb2reaction_b200.arch import atomic_numbers
    z = atomic_numbers(elem)
    return z, W.merge_mole(state, arch, z, charge, spin, task)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built on demand; nvcc cross-compiles without a GPU)."""
    from pdb2reaction_b200.csrc import build as b
    return b.build()


@pytest.fixture()
def small_model(state4, arch4, monkeypatch):
    """Route model='test-4x' to the 4-expert test weights (the default ID means 32 experts)."""
    from pdb2reaction_b200 import calculator as calc_mod
    from pdb2reaction_b200.checkpoint import EnergyTransform
    monkeypatch.setattr(calc_mod, "load_model_state", lambda model, arch, task_name="omol": (state4, EnergyTransform()))
    orig = calc_mod.CudaBackend.__init__

    def init(self, elem, **kw):
        kw["arch"] = arch4
        orig(self, elem, **kw)

    monkeypatch.setattr(calc_mod.CudaBackend, "__init__", init)
    calc_mod._engine_cache.clear()
    yield
    calc_mod._engine_cache.clear()
