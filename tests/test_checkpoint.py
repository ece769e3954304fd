"""Real-checkpoint loader: fairchem-style names -> this package's state dict.

The key table is a restatement from recall (no fairchem / checkpoint offline), so these tests pin what CAN be
pinned here: the mapping is a bijection on the restated architecture, a checkpoint file pickled with classes of
packages that are not installed can still be read, mismatches fail loudly, and the energy post-processing
(normaliser + element references) is applied by the calculator backend.
"""
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from pdb2reaction_b200 import checkpoint as ck
from pdb2reaction_b200.arch import UMAArch, atomic_numbers


def test_round_trip_every_tensor(state4, arch4):
    fc = ck.export_fairchem_style(state4)
    assert all(k.startswith(("backbone.", "output_heads.")) for k in fc)
    # MoLE expert tensors carry fairchem's plural name
    assert any(k.endswith("so2_conv_1.fc_m0.weights") for k in fc)
    back, rep = ck.convert_state_dict(fc, arch4)
    assert rep.ok() and rep.mapped >= len(state4)
    assert set(back) == set(state4)
    for k in state4:
        assert torch.equal(back[k], state4[k]), k


def test_ignores_buffers_and_reports_unknown_or_missing(state4, arch4):
    fc = ck.export_fairchem_style(state4)
    fc["backbone.blocks.0.edge_wise.so2_conv_1.rad_func.net.0.offset"] = torch.zeros(3)      # buffer-like
    fc["backbone.SO3_grid.lmax_lmax.to_grid_mat"] = torch.zeros(2, 2)
    back, rep = ck.convert_state_dict(fc, arch4)
    assert rep.ok() and len(rep.ignored) == 2
    fc["backbone.some_new_module.weight"] = torch.zeros(3)
    del fc["backbone.norm.affine_weight"]
    with pytest.raises(ValueError) as ei:
        ck.convert_state_dict(fc, arch4)
    assert "some_new_module" in str(ei.value) and "norm.affine_weight" in str(ei.value)
    _, rep = ck.convert_state_dict(fc, arch4, strict=False)
    assert not rep.ok()


def test_wrong_shape_is_reported(state4, arch4):
    fc = ck.export_fairchem_style(state4)
    k = "backbone.sphere_embedding.weight"
    fc[k] = fc[k][:, :64]
    with pytest.raises(ValueError, match="sphere_embedding.weight: got"):
        ck.convert_state_dict(fc, arch4)


def test_reads_pickle_with_uninstalled_classes(tmp_path, state4, arch4):
    """A fairchem MLIPInferenceCheckpoint references fairchem / omegaconf classes; the stub unpickler
    must get the tensors out without those packages."""
    mod = types.ModuleType("fairchem_fake_pkg.units.api")

    class MLIPInferenceCheckpoint:
        pass

    class DictConfig:
        pass
    MLIPInferenceCheckpoint.__module__ = DictConfig.__module__ = mod.__name__
    MLIPInferenceCheckpoint.__qualname__, DictConfig.__qualname__ = "MLIPInferenceCheckpoint", "DictConfig"
    mod.MLIPInferenceCheckpoint, mod.DictConfig = MLIPInferenceCheckpoint, DictConfig
    sys.modules[mod.__name__] = mod
    sys.modules["fairchem_fake_pkg"] = types.ModuleType("fairchem_fake_pkg")
    sys.modules["fairchem_fake_pkg.units"] = types.ModuleType("fairchem_fake_pkg.units")
    try:
        obj = MLIPInferenceCheckpoint()
        obj.model_state_dict = {"junk": torch.zeros(1)}
        obj.ema_state_dict = {"module": ck.export_fairchem_style(state4, prefix="module.backbone."),
                              "n_averaged": torch.tensor(7)}
        # heads are stored outside the averaged backbone prefix in our export: move them in
        obj.ema_state_dict["module"] = {("module." + k if k.startswith("output_heads.") else k): v
                                        for k, v in obj.ema_state_dict["module"].items()}
        cfg = DictConfig()
        refs = np.zeros(100)
        refs[1], refs[6], refs[8] = -13.6, -1029.1, -2041.3
        cfg._content = [{"name": "omol_energy", "property": "energy", "datasets": ["omol"],
                         "normalizer": {"mean": 0.25, "rmsd": 1.75},
                         "element_references": {"element_references": refs.tolist()}},
                        {"name": "omol_forces", "property": "forces", "datasets": ["omol"]}]
        obj.tasks_config = cfg
        path = tmp_path / "uma-fake.pt"
        torch.save(obj, path)
    finally:
        for k in ("fairchem_fake_pkg.units.api", "fairchem_fake_pkg.units", "fairchem_fake_pkg"):
            sys.modules.pop(k, None)
    with pytest.raises(Exception):
        torch.load(path, map_location="cpu", weights_only=False)          # the package really is gone
    sd, tr = ck.load_checkpoint(str(path), arch4, task_name="omol")
    assert set(sd) == set(state4) and all(torch.equal(sd[k], state4[k]) for k in state4)
    assert tr.scale == 1.75 and tr.shift == 0.25
    z = atomic_numbers(["C", "H", "H", "O"])
    assert tr.constant_for(z) == pytest.approx(0.25 - 1029.1 - 2 * 13.6 - 2041.3)
    assert ck.load_checkpoint(str(path), arch4, task_name="oc20")[1].is_identity      # no such task: identity


def test_own_naming_passes_through(tmp_path, state4, arch4):
    path = tmp_path / "own.pt"
    torch.save(state4, path)
    sd, tr = ck.load_checkpoint(str(path), arch4)
    assert tr.is_identity and all(torch.equal(sd[k], state4[k]) for k in state4)


def test_backend_applies_energy_transform(tmp_path, state4, arch4, monkeypatch):
    """E_out = E*rmsd + mean + sum refs[Z], F_out = F*rmsd -- applied on top of whatever the engines return."""
    from pdb2reaction_b200 import calculator as cm
    tr = ck.EnergyTransform(scale=2.0, shift=1.0, element_refs=torch.arange(100, dtype=torch.float64))

    class Eng:
        device = 0

        def energy_forces_host(self, pos, forces):
            b, n = pos.shape[:2]
            return np.full(b, 3.0), (np.ones((b, n, 3), np.float32) if forces else None)

    be = cm.CudaBackend.__new__(cm.CudaBackend)
    be.engines, be._pool, be.transform = [Eng()], None, tr
    be.z = atomic_numbers(["H", "C", "O"])
    be._e_const = tr.constant_for(be.z)
    e, f = be.evaluate(np.zeros((2, 3, 3)), forces=True)
    assert np.allclose(e, 3.0 * 2.0 + 1.0 + (1 + 6 + 8)) and np.allclose(f, 2.0)
