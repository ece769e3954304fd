"""fairchem-core v2 UMA checkpoint  ->  the un-merged state dict this package evaluates.

The reference obtains its model with ``pretrained_mlip.get_predict_unit(model, device=...)``
(``pdb2reaction/uma_pysis.py:228-250``), i.e. it unpickles a fairchem ``MLIPInferenceCheckpoint``
(``model_config``, ``model_state_dict``, ``ema_state_dict``, ``tasks_config``) downloaded from the gated
``facebook/UMA`` repository.  Neither fairchem nor the checkpoint exists offline, so:

* the key table below restates the parameter names of fairchem's ``eSCNMDMoeBackbone`` /
  ``MLP_EFS_Head`` **from recall** -- it is UNVERIFIED against a real file.  Every rule is a regex with
  alternatives for the names I am unsure of, and ``convert_state_dict`` fails loudly with the list of unmapped
  / missing / mis-shaped keys instead of guessing, so a maintainer with the real checkpoint can fix the table
  in minutes;
* the unpickler stubs every ``fairchem.*`` / ``omegaconf.*`` class it meets, so a real checkpoint can be read
  WITHOUT fairchem installed (only ``torch`` tensors and plain containers are needed from it);
* ``export_fairchem_style`` is the inverse mapping, used by the tests to prove the round trip
  (random-init weights -> fairchem-style checkpoint file -> loader -> identical state dict).

Energy post-processing of the prediction unit (task normaliser ``E * rmsd + mean`` and per-element linear
references) is returned as an ``EnergyTransform`` and applied by the calculator backend.
"""
from __future__ import annotations

import io
import pickle
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch

from .arch import DATASET_LIST, UMAArch
from . import weights as _weights

# ----------------------------------------------------------------------------------------------------------
# key table:  (regex over the fairchem backbone / head key, template of this package's key)
# RadialMLP = nn.Sequential(Linear, LayerNorm, SiLU, Linear, LayerNorm, SiLU, Linear) -> indices 0 1 3 4 6
# ----------------------------------------------------------------------------------------------------------
_RAD = {"0": "lin1", "1": "ln1", "3": "lin2", "4": "ln2", "6": "lin3"}
_W = r"(?:weights?|weight)"          # MoLE modules hold ``weights`` [experts, out, in], plain Linear ``weight``

_RULES: List[Tuple[str, str]] = [
    (r"sphere_embedding\.weight", "sphere_embedding.weight"),
    (r"source_embedding\.weight", "source_embedding.weight"),
    (r"target_embedding\.weight", "target_embedding.weight"),
    (r"charge_embedding\.(?:rand_emb\.|embedding\.|emb\.)?weight", "charge_embedding.weight"),
    (r"spin_embedding\.(?:rand_emb\.|embedding\.|emb\.)?weight", "spin_embedding.weight"),
    (r"mix_csd\.(weight|bias)", r"mix_csd.\1"),
    (r"(?:composition_embedding|mole_composition_embedding)\.weight", "composition_embedding.weight"),
    (r"routing_mlp\.([024])\.(weight|bias)", r"routing_mlp.\1.\2"),
    (r"edge_degree_embedding\.rad_func\.net\.([01346])\.(weight|bias)", r"edge_degree.rad.{rad\1}.\2"),
    (r"blocks\.(\d+)\.norm_([12])\.affine_(weight|bias)", r"blocks.\1.norm_\2.affine_\3"),
    (r"blocks\.(\d+)\.edge_wise\.so2_conv_1\.rad_func\.net\.([01346])\.(weight|bias)",
     r"blocks.\1.edge.conv1.rad.{rad\2}.\3"),
    (r"blocks\.(\d+)\.edge_wise\.so2_conv_([12])\.fc_m0\." + _W, r"blocks.\1.edge.conv\2.fc_m0.weight"),
    (r"blocks\.(\d+)\.edge_wise\.so2_conv_([12])\.fc_m0\.bias", r"blocks.\1.edge.conv\2.fc_m0.bias"),
    (r"blocks\.(\d+)\.edge_wise\.so2_conv_([12])\.so2_m_conv\.(\d+)\.fc\." + _W,
     r"blocks.\1.edge.conv\2.fc_m{m\3}.weight"),
    (r"blocks\.(\d+)\.(?:atom_wise|ffn)\.scalar_mlp\.(?:0\.)?(weight|bias)", r"blocks.\1.ffn.scalar_mlp.\2"),
    (r"blocks\.(\d+)\.(?:atom_wise|ffn)\.so3_linear_([12])\.(weight|bias)", r"blocks.\1.ffn.so3_\2.\3"),
    (r"norm\.affine_(weight|bias)", r"norm.affine_\1"),
    (r"(?:head\.)?energy_block\.([024])\.(weight|bias)", r"head.\1.\2"),
]
_COMPILED = [(re.compile(r"^" + pat + r"$"), tpl) for pat, tpl in _RULES]
_DATASET_RE = re.compile(r"^dataset_embedding\.(?:dataset_emb_dict\.)?(\w+)\.weight$")
_PREFIXES = ("module.", "model.", "backbone.", "output_heads.energyandforcehead.", "output_heads.energy.",
             "heads.energyandforcehead.", "head.")
# buffers / bookkeeping tensors of fairchem modules that carry no parameter of the restated model
_IGNORABLE = re.compile(r"(?:^|\.)(?:Jd|Jd_\d+|Jd_list\.\d+|mappingReduced\.[\w.]+|SO3_grid\.[\w.]+|to_m|m_size|l_harmonic|"
                        r"m_harmonic|m_complex|mask_indices_cache|rotate_inv_rescale_cache|coefficient_idx\w*|"
                        r"num_batches_tracked|n_averaged|offset|coeff|expand_index|balance_degree_weight|"
                        r"distance_expansion\.\w+|envelope\.\w+|global_mole_tensors\.\w+)$")


def _strip(key: str) -> str:
    changed = True
    while changed:
        changed = False
        for p in _PREFIXES:
            if key.startswith(p) and not key.startswith("head.") or (p != "head." and key.startswith(p)):
                key = key[len(p):]
                changed = True
    return key


def _map_key(key: str) -> Optional[str]:
    for rx, tpl in _COMPILED:
        m = rx.match(key)
        if m:
            out = m.expand(re.sub(r"\{rad\\(\d)\}", r"{rad\\g<\1>}", re.sub(r"\{m\\(\d)\}", r"{m\\g<\1>}", tpl)))
            out = re.sub(r"\{rad(\d)\}", lambda mm: _RAD[mm.group(1)], out)
            out = re.sub(r"\{m(\d+)\}", lambda mm: str(int(mm.group(1)) + 1), out)     # so2_m_conv.i <-> m = i + 1
            return out
    return None


@dataclass
class EnergyTransform:
    """E_out = E_model * scale + shift + sum_atoms element_refs[Z];  F_out = F_model * scale."""
    scale: float = 1.0
    shift: float = 0.0
    element_refs: Optional[torch.Tensor] = None        # [max_num_elements] eV or None

    def constant_for(self, z) -> float:
        c = float(self.shift)
        if self.element_refs is not None:
            c += float(self.element_refs[torch.as_tensor(list(z), dtype=torch.long)].sum())
        return c

    @property
    def is_identity(self) -> bool:
        return self.scale == 1.0 and self.shift == 0.0 and self.element_refs is None


@dataclass
class ConversionReport:
    mapped: int = 0
    ignored: List[str] = field(default_factory=list)
    unmapped: List[str] = field(default_factory=list)
    missing: List[str] = field(default_factory=list)
    bad_shape: List[str] = field(default_factory=list)

    def ok(self) -> bool:
        return not (self.unmapped or self.missing or self.bad_shape)

    def __str__(self):
        def head(xs):
            return ", ".join(xs[:8]) + (f", ... (+{len(xs) - 8})" if len(xs) > 8 else "")
        return (f"mapped {self.mapped}; ignored {len(self.ignored)}; unmapped [{head(self.unmapped)}]; "
                f"missing [{head(self.missing)}]; bad shape [{head(self.bad_shape)}]")


def expected_shapes(arch: UMAArch) -> Dict[str, Tuple[int, ...]]:
    """Name -> shape of every tensor of the un-merged state dict (from the documented random init)."""
    small = UMAArch(**{**arch.as_dict(), "num_experts": 1})
    sd = _weights.init_uma_weights(small, seed=0)
    out = {}
    for k, v in sd.items():
        shp = tuple(v.shape)
        if v.dim() == 3 and ".edge." in k:
            shp = (arch.num_experts,) + shp[1:]
        elif k.startswith("routing_mlp.4"):
            shp = (arch.num_experts,) + shp[1:]
        out[k] = shp
    return out


def convert_state_dict(sd_fc: Dict[str, torch.Tensor], arch: UMAArch = UMAArch(), strict: bool = True):
    """fairchem-named tensors -> (state dict in this package's naming, ConversionReport)."""
    out: Dict[str, torch.Tensor] = {}
    rep = ConversionReport()
    ds_rows: Dict[str, torch.Tensor] = {}
    for k_raw, v in sd_fc.items():
        if not isinstance(v, torch.Tensor):
            continue
        k = _strip(k_raw)
        m = _DATASET_RE.match(k)
        if m:
            ds_rows[m.group(1)] = v.reshape(-1)
            continue
        if k == "dataset_embedding.weight":
            out[k] = v
            rep.mapped += 1
            continue
        ours = _map_key(k)
        if ours is None:
            (rep.ignored if _IGNORABLE.search(k) else rep.unmapped).append(k_raw)
            continue
        out[ours] = v.detach().to(torch.float32)
        rep.mapped += 1
    if ds_rows:
        miss = [d for d in DATASET_LIST if d not in ds_rows]
        if miss:
            rep.missing += [f"dataset_embedding[{d}]" for d in miss]
        else:
            out["dataset_embedding.weight"] = torch.stack([ds_rows[d] for d in DATASET_LIST]).to(torch.float32)
            rep.mapped += 1
    # SO3 linear weights may be stored [lmax+1, out, in] (as here) -- anything else is reported below
    exp = expected_shapes(arch)
    for k, shp in exp.items():
        if k not in out:
            rep.missing.append(k)
        elif tuple(out[k].shape) != shp:
            if out[k].numel() == int(torch.tensor(shp).prod()) and out[k].dim() == len(shp) + 1 and out[k].shape[0] == 1:
                out[k] = out[k].reshape(shp)                   # e.g. an Embedding(1, C) row
            else:
                rep.bad_shape.append(f"{k}: got {tuple(out[k].shape)}, expected {shp}")
    extra = [k for k in out if k not in exp]
    for k in extra:
        rep.unmapped.append(k + " (no such parameter in the restated architecture)")
    if strict and not rep.ok():
        raise ValueError("fairchem checkpoint does not match the restated uma-s-1p1 architecture -- fix the key table in "
                         "pdb2reaction_b200/checkpoint.py: " + str(rep))
    return out, rep


def export_fairchem_style(sd: Dict[str, torch.Tensor], prefix: str = "backbone.") -> Dict[str, torch.Tensor]:
    """Inverse of convert_state_dict with the primary spelling of every fairchem name (tests, tooling)."""
    inv_rad = {v: k for k, v in _RAD.items()}
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        m = re.match(r"^(edge_degree|blocks\.(\d+)\.edge\.conv1)\.rad\.(lin\d|ln\d)\.(weight|bias)$", k)
        if m:
            base = "edge_degree_embedding" if m.group(1) == "edge_degree" else f"blocks.{m.group(2)}.edge_wise.so2_conv_1"
            out[f"{prefix}{base}.rad_func.net.{inv_rad[m.group(3)]}.{m.group(4)}"] = v
            continue
        m = re.match(r"^blocks\.(\d+)\.edge\.conv([12])\.fc_m(\d)\.(weight|bias)$", k)
        if m:
            l, c, mm, wb = m.groups()
            name = "weights" if wb == "weight" else "bias"
            mod = "fc_m0" if mm == "0" else f"so2_m_conv.{int(mm) - 1}.fc"
            out[f"{prefix}blocks.{l}.edge_wise.so2_conv_{c}.{mod}.{name}"] = v
            continue
        m = re.match(r"^blocks\.(\d+)\.ffn\.scalar_mlp\.(weight|bias)$", k)
        if m:
            out[f"{prefix}blocks.{m.group(1)}.atom_wise.scalar_mlp.0.{m.group(2)}"] = v
            continue
        m = re.match(r"^blocks\.(\d+)\.ffn\.so3_([12])\.(weight|bias)$", k)
        if m:
            out[f"{prefix}blocks.{m.group(1)}.atom_wise.so3_linear_{m.group(2)}.{m.group(3)}"] = v
            continue
        m = re.match(r"^head\.([024])\.(weight|bias)$", k)
        if m:
            out[f"output_heads.energyandforcehead.head.energy_block.{m.group(1)}.{m.group(2)}"] = v
            continue
        if k in ("charge_embedding.weight", "spin_embedding.weight"):
            out[prefix + k.replace(".weight", ".rand_emb.weight")] = v
            continue
        if k == "dataset_embedding.weight":
            for i, d in enumerate(DATASET_LIST):
                out[f"{prefix}dataset_embedding.dataset_emb_dict.{d}.weight"] = v[i:i + 1].clone()
            continue
        out[prefix + k] = v
    return out


# ----------------------------------------------------------------------------------------------------------
# reading a checkpoint file without fairchem installed
# ----------------------------------------------------------------------------------------------------------
class _Stub:
    """Stands in for any class of a package that is not installed (fairchem, omegaconf, ...)."""

    def __init__(self, *a, **kw):
        self._args, self._kwargs = a, kw

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state

    def __reduce_ex__(self, protocol):          # never re-pickled
        raise pickle.PicklingError("stub object")


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return type(name, (_Stub,), {"__module__": module})


class _stub_pickle:                               # the ``pickle_module`` interface torch.load expects
    __name__ = "pdb2reaction_b200_stub_pickle"
    Unpickler = _StubUnpickler
    load = staticmethod(lambda f, **kw: _StubUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _StubUnpickler(io.BytesIO(b), **kw).load())


def _get(obj: Any, name: str, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def _to_plain(x: Any) -> Any:
    """omegaconf / stub containers -> plain python (best effort)."""
    if isinstance(x, dict):
        return {k: _to_plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_to_plain(v) for v in x]
    if isinstance(x, _Stub):
        d = {k: v for k, v in x.__dict__.items() if not k.startswith("__")}
        for key in ("_content", "_val", "_value"):
            if key in d:
                return _to_plain(d[key])
        return {k: _to_plain(v) for k, v in d.items()}
    return x


def energy_transform_from_tasks(tasks_config: Any, task_name: str) -> EnergyTransform:
    """Normaliser + element references of the energy task of ``task_name`` (identity when absent)."""
    tasks = _to_plain(tasks_config)
    if not tasks:
        return EnergyTransform()
    items = list(tasks.values()) if isinstance(tasks, dict) else list(tasks)
    for t in items:
        if not isinstance(t, dict):
            continue
        name = str(t.get("name", ""))
        datasets = t.get("datasets") or []
        if t.get("property", "energy") != "energy" and "energy" not in name:
            continue
        if task_name not in name and task_name not in [str(d) for d in datasets]:
            continue
        norm = t.get("normalizer") or {}
        scale = float(norm.get("rmsd", norm.get("std", 1.0)) or 1.0)
        shift = float(norm.get("mean", 0.0) or 0.0)
        refs = t.get("element_references")
        ref_t = None
        if isinstance(refs, dict):
            refs = refs.get("element_references", refs.get("lin_ref"))
        if refs is not None:
            ref_t = torch.as_tensor(refs, dtype=torch.float64).reshape(-1)
        return EnergyTransform(scale=scale, shift=shift, element_refs=ref_t)
    return EnergyTransform()


def load_checkpoint(path: str, arch: UMAArch = UMAArch(), task_name: str = "omol", strict: bool = True):
    """-> (un-merged state dict in this package's naming, EnergyTransform).

    Accepts (a) a ``torch.save``d state dict already in this package's naming, (b) a plain fairchem-named
    state dict, (c) a fairchem ``MLIPInferenceCheckpoint`` pickle (EMA weights preferred, as ``predict`` uses).
    """
    obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_stub_pickle)
    if isinstance(obj, dict) and "sphere_embedding.weight" in obj and "blocks.0.edge.conv1.fc_m0.weight" in obj:
        return obj, EnergyTransform()
    sd = None
    for name in ("ema_state_dict", "model_state_dict", "state_dict"):
        cand = _get(obj, name)
        if isinstance(cand, dict) and cand:
            sd = cand
            break
    if sd is None and isinstance(obj, dict):
        sd = obj
    if sd is None:
        raise ValueError(f"{path}: no state dict found in the checkpoint object ({type(obj).__name__})")
    if "module" in sd and isinstance(sd["module"], dict):       # AveragedModel wrapper
        sd = sd["module"]
    converted, _ = convert_state_dict(sd, arch, strict=strict)
    return converted, energy_transform_from_tasks(_get(obj, "tasks_config"), task_name)
