"""ctypes binding of the umab C ABI (include/umab.h) + the host-side weight preparation.

The engine evaluates a BATCH of images of one composition per call -- the data-parallel
replacement of the reference's one-image-per-call ``UMAcore.compute``
(``pdb2reaction/uma_pysis.py:330-419``).  There is no CPU fallback: if the library cannot be
loaded or no CUDA device exists, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .arch import UMAArch

# UMAB_LIB: load another build of the same library (A/B measurements of compile-time kernel variants)
_LIB_PATH = os.environ.get("UMAB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libumab.so")
_lib = None

ABI_VERSION = 7
UMAB_RETRY = 2          # umab_last_call: the sync-free graph build overflowed its edge capacity, repeat the call
# "simt": fp32 FFMA GEMMs; "tc": tcgen05 bf16x3 tensor-core GEMMs; "auto": tc for images of >= 100 atoms
DEFAULT_GEMM = "auto"
GEMM_MODES = {"simt": 0, "tc": 1, "auto": 2}

# every symbol include/umab.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "umab_abi_version", "umab_last_error", "umab_create", "umab_destroy", "umab_set_weight",
    "umab_finalize_weights", "umab_set_option", "umab_get_option", "umab_set_system", "umab_build_graph", "umab_graph_counts",
    "umab_graph_copy", "umab_energy_forces", "umab_energy_forces_host", "umab_last_call", "umab_forces_jvp", "umab_gemm", "umab_gemm_bench",
    "umab_debug_tensor", "umab_stats", "umab_profile", "umab_profile_read", "umab_profile_name",
    "umab_hessian_fd_columns", "umab_hessian_mw_workspace", "umab_hessian_mw_project", "umab_release_workspace",
)


class UmabConfig(ctypes.Structure):
    _fields_ = [
        ("sphere_channels", ctypes.c_int32),
        ("hidden_channels", ctypes.c_int32),
        ("num_distance_basis", ctypes.c_int32),
        ("num_layers", ctypes.c_int32),
        ("max_neighbors", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("debug", ctypes.c_int32),
        ("gemm_mode", ctypes.c_int32),
        ("cutoff", ctypes.c_float),
        ("edge_degree_rescale", ctypes.c_float),
        ("workspace_bytes", ctypes.c_int64),
        ("store_bytes", ctypes.c_int64),
    ]


def load_library(path: Optional[str] = None):
    """Load libumab.so (built in-tree by ``pdb2reaction_b200/csrc/build.py``).  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build the CUDA extension first (python -m pdb2reaction_b200.csrc.build "
            "or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.umab_abi_version.restype = i32
    lib.umab_last_error.restype = ctypes.c_char_p
    lib.umab_create.argtypes = [ctypes.POINTER(UmabConfig), ctypes.POINTER(vp)]
    lib.umab_destroy.argtypes = [vp]
    lib.umab_destroy.restype = None
    lib.umab_set_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t]
    lib.umab_finalize_weights.argtypes = [vp]
    lib.umab_set_system.argtypes = [vp, vp, i32]
    lib.umab_build_graph.argtypes = [vp, vp, i32, vp]
    lib.umab_graph_counts.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.umab_graph_copy.argtypes = [vp, vp, vp, vp, vp]
    lib.umab_energy_forces.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.umab_energy_forces_host.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.umab_last_call.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.umab_forces_jvp.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    lib.umab_gemm.argtypes = [i32, vp, vp, vp, vp, i64, i32, i32, vp]
    lib.umab_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.umab_get_option.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
    lib.umab_gemm_bench.argtypes = [i32, vp, vp, vp, i64, i32, i32, i32, ctypes.POINTER(ctypes.c_double), vp]
    lib.umab_debug_tensor.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    lib.umab_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.umab_profile.argtypes = [vp, i32]
    lib.umab_profile_read.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64),
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.umab_hessian_fd_columns.argtypes = [vp, vp, i32, i32, ctypes.c_double, vp, i64, i32, vp]
    lib.umab_hessian_mw_workspace.argtypes = [i32, i32]
    lib.umab_hessian_mw_workspace.restype = i64
    lib.umab_hessian_mw_project.argtypes = [vp, i32, vp, vp, i32, vp, i64, vp]
    lib.umab_release_workspace.argtypes = [vp]
    lib.umab_profile_name.argtypes = [i32]
    lib.umab_profile_name.restype = ctypes.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("umab_last_error", "umab_destroy", "umab_abi_version", "umab_profile_name",
                        "umab_hessian_mw_workspace"):
            fn.restype = i32
    if lib.umab_abi_version() != ABI_VERSION:
        raise RuntimeError("libumab.so ABI version mismatch: rebuild the extension")
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        msg = lib.umab_last_error().decode("utf-8", "replace")
        if "out of memory" in msg.lower():
            raise torch.cuda.OutOfMemoryError("CUDA out of memory (umab): " + msg)
        raise RuntimeError("umab: " + msg)


def prepare_engine_weights(merged: Dict[str, torch.Tensor], arch: UMAArch) -> Dict[str, torch.Tensor]:
    """Merged state dict -> the flat set of fp32 arrays the kernels consume.

    Adds, for every GEMM weight W [out, in], its transpose ``*_t`` (the backward contracts with
    W instead of W^T and the kernels only implement  A . W^T), and for every radial MLP the
    split first layer: ``w1g`` (Gaussian columns) + per-element tables ``t_src``/``t_tgt``
    (= embedding @ W1_part^T), so the [E, 320] x_edge matrix is never materialised.  The m > 0 SO(2)
    weights are expanded to their complex block form [[W_r, -W_i], [W_i, W_r]].
    """
    nb, ce = arch.num_distance_basis, arch.edge_channels
    out: Dict[str, torch.Tensor] = {}

    def put(name, t):
        out[name] = t.detach().to(torch.float32).contiguous().cpu()

    def put_t(name, t):
        put(name, t)
        put(name + "_t", t.transpose(-1, -2))

    def radial(prefix):
        w1 = merged[prefix + ".lin1.weight"].to(torch.float32)
        put_t(prefix + ".w1g", w1[:, :nb])
        put(prefix + ".t_src", merged["source_embedding.weight"].float() @ w1[:, nb:nb + ce].T)
        put(prefix + ".t_tgt", merged["target_embedding.weight"].float() @ w1[:, nb + ce:].T)
        put(prefix + ".lin1.bias", merged[prefix + ".lin1.bias"])
        for k in ("ln1", "ln2"):
            put(f"{prefix}.{k}.weight", merged[f"{prefix}.{k}.weight"])
            put(f"{prefix}.{k}.bias", merged[f"{prefix}.{k}.bias"])
        for k in ("lin2", "lin3"):
            put_t(f"{prefix}.{k}.weight", merged[f"{prefix}.{k}.weight"])
            put(f"{prefix}.{k}.bias", merged[f"{prefix}.{k}.bias"])

    put("sphere_embedding.weight", merged["sphere_embedding.weight"])
    put("csd", merged["csd"])
    put("norm.affine_weight", merged["norm.affine_weight"])
    put("norm.affine_bias", merged["norm.affine_bias"])
    put_t("head.0.weight", merged["head.0.weight"])
    put("head.0.bias", merged["head.0.bias"])
    put_t("head.2.weight", merged["head.2.weight"])
    put("head.2.bias", merged["head.2.bias"])
    put("head.4.weight", merged["head.4.weight"])
    put("head.4.bias", merged["head.4.bias"])
    radial("edge_degree.rad")
    for l in range(arch.num_layers):
        p = f"blocks.{l}"
        for n in ("norm_1", "norm_2"):
            put(f"{p}.{n}.affine_weight", merged[f"{p}.{n}.affine_weight"])
            put(f"{p}.{n}.affine_bias", merged[f"{p}.{n}.affine_bias"])
        radial(p + ".edge.conv1.rad")
        for conv in ("conv1", "conv2"):
            for m in range(arch.lmax + 1):
                key = f"{p}.edge.{conv}.fc_m{m}"
                w = merged[key + ".weight"]
                if w.dim() != 2:
                    raise ValueError(f"{key}.weight is not merged (shape {tuple(w.shape)})")
                if m > 0:
                    # fold the SO(2) (+m, -m) combination into the weight: [x(+m) | x(-m)] @ Wc^T = [o_r | o_i]
                    # (same FLOPs, half the conv output to write and re-read; twin: oracle/staged.so2_complex_weight)
                    ho = w.shape[0] // 2
                    wr, wi = w[:ho].float(), w[ho:].float()
                    w = torch.cat([torch.cat([wr, -wi], 1), torch.cat([wi, wr], 1)], 0)
                put_t(key + ".weight", w)
                if m == 0:
                    put(key + ".bias", merged[key + ".bias"])
        put_t(p + ".ffn.scalar_mlp.weight", merged[p + ".ffn.scalar_mlp.weight"])
        put(p + ".ffn.scalar_mlp.bias", merged[p + ".ffn.scalar_mlp.bias"])
        for k in ("so3_1", "so3_2"):
            put_t(f"{p}.ffn.{k}.weight", merged[f"{p}.ffn.{k}.weight"])
            put(f"{p}.ffn.{k}.bias", merged[f"{p}.ffn.{k}.bias"])
    return out


class _DevPtr:
    """Minimal __cuda_array_interface__ view of a raw fp32 device pointer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


MAX_ATOMS_PER_CALL = 49152        # node-state memory bound of one library call (~100 KB per atom)
# kept for the backward per edge and layer: conv outputs 1408 + 1152 floats, radial weights 1536, radial pre-activations 2 x 128
STORE_BYTES_PER_EDGE = (2560 + 1536 + 256) * 4 * 4   # 17.4 KB per edge and layer, 4 layers


class UmabEngine:
    """One engine = one (weights, composition, charge, spin, task) on one GPU."""

    def __init__(self, merged_weights: Dict[str, torch.Tensor], z: Sequence[int], arch: UMAArch = UMAArch(), *,
                 device: int = 0, cutoff: Optional[float] = None, max_neighbors: Optional[int] = None,
                 debug: bool = False, gemm_mode: Optional[int] = None, workspace_bytes: int = 0,
                 store_bytes: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("pdb2reaction_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.lib = load_library()
        self.arch = arch
        self.device = int(device)
        self.n_atoms = len(z)
        if gemm_mode is None:
            gemm_mode = GEMM_MODES[os.environ.get("UMAB_GEMM", DEFAULT_GEMM)]
        cfg = UmabConfig(arch.sphere_channels, arch.hidden_channels, arch.num_distance_basis, arch.num_layers,
                         int(max_neighbors if max_neighbors is not None else arch.max_neighbors), self.device,
                         int(bool(debug)), int(gemm_mode),
                         float(cutoff if cutoff is not None else arch.cutoff), float(arch.edge_degree_rescale),
                         int(workspace_bytes), int(store_bytes))
        self.cfg = cfg
        h = ctypes.c_void_p()
        _check(self.lib, self.lib.umab_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        for name, t in prepare_engine_weights(merged_weights, arch).items():
            a = t.numpy()
            _check(self.lib, self.lib.umab_set_weight(self._h, name.encode(), a.ctypes.data, a.size))
        _check(self.lib, self.lib.umab_finalize_weights(self._h))
        zz = np.ascontiguousarray(np.asarray(list(z), dtype=np.int32))
        _check(self.lib, self.lib.umab_set_system(self._h, zz.ctypes.data, int(zz.size)))
        self._store_budget = (-1 if store_bytes < 0 else store_bytes if store_bytes > 0 else
                              0.45 * torch.cuda.get_device_properties(self.device).total_memory)
        self._edges_per_atom = 85.0          # refined from the measured graphs
        self.last_call_edges = 0             # edges summed over the sub-batches of the last public call
        self.last_call_subcalls = 0

    def images_per_call(self, forces: bool = True) -> int:
        """How many images one library call should take: bounded by the node state, and -- when
        forces are wanted -- sized so the per-layer conv outputs fit the store budget, which lets
        the backward skip the recomputation of the SO(2) convolutions."""
        n = self.n_atoms
        cap = max(1, MAX_ATOMS_PER_CALL // n)
        if forces and self._store_budget > 0:
            per_image = n * self._edges_per_atom * STORE_BYTES_PER_EDGE * self.arch.num_layers / 4
            fit = int(self._store_budget // per_image)
            if fit >= 1:
                cap = min(cap, fit)
        return cap

    def close(self):
        if getattr(self, "_h", None):
            self.lib.umab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream_ptr(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _last_call(self) -> bool:
        """Collect the last call's counters; False = its sync-free graph build overflowed (repeat the call)."""
        ni, ne, ns = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        rc = self.lib.umab_last_call(self._h, ctypes.byref(ni), ctypes.byref(ne), ctypes.byref(ns))
        if rc == UMAB_RETRY:
            return False
        _check(self.lib, rc)
        self.last_call_edges, self.last_call_subcalls = ne.value, ns.value
        if ni.value > 0 and ne.value > 0:
            self._edges_per_atom = max(20.0, 1.03 * ne.value / (ni.value * self.n_atoms))
        return True

    # ------------------------------------------------------------------ device-resident API
    def energy_forces(self, pos: torch.Tensor, forces: bool = True):
        """pos [B, N, 3] float32 CUDA tensor (Angstrom) -> (E [B] float64, F [B,N,3] float32 | None).
        ONE library call; batches larger than the per-layer stores allow run as sub-batches inside it."""
        assert pos.is_cuda and pos.dtype == torch.float32 and pos.dim() == 3 and pos.shape[1] == self.n_atoms
        pos = pos.contiguous()
        b = pos.shape[0]
        e = torch.empty(b, dtype=torch.float64, device=pos.device)
        f = torch.empty_like(pos) if forces else None
        for _ in range(4):
            _check(self.lib, self.lib.umab_energy_forces(self._h, pos.data_ptr(), b, e.data_ptr(),
                                                         f.data_ptr() if forces else None, self._stream_ptr()))
            if self._last_call():
                return e, f
        raise RuntimeError("umab: edge capacity exceeded repeatedly")

    # ------------------------------------------------------------------ host-buffer API (e2e path)
    def energy_forces_host(self, pos: np.ndarray, forces: bool = True):
        """pos [B, N, 3] float32 numpy (Angstrom) -> (E [B] float64 numpy, F [B,N,3] float32 numpy | None)."""
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        assert pos.ndim == 3 and pos.shape[1] == self.n_atoms
        b = pos.shape[0]
        e = np.empty(b, dtype=np.float64)
        f = np.empty_like(pos) if forces else None
        _check(self.lib, self.lib.umab_energy_forces_host(self._h, pos.ctypes.data, b, e.ctypes.data,
                                                          f.ctypes.data if forces else None, self._stream_ptr()))
        self._last_call()
        return e, f

    # ------------------------------------------------------------------ analytic Hessian columns
    def forces_jvp(self, pos: torch.Tensor, tangent: torch.Tensor):
        """pos, tangent [B, N, 3] float32 CUDA -> (F [B,N,3], dF [B,N,3]) with
        dF = d(forces)/d(eps) at pos + eps * tangent  (= -H . tangent), computed by running the
        forward and the hand-written backward on dual numbers (no finite differences)."""
        assert pos.is_cuda and pos.dtype == torch.float32 and pos.dim() == 3 and pos.shape[1] == self.n_atoms
        assert tangent.shape == pos.shape and tangent.dtype == torch.float32 and tangent.is_cuda
        pos, tangent = pos.contiguous(), tangent.contiguous()
        b = pos.shape[0]
        f = torch.empty_like(pos)
        df = torch.empty_like(pos)
        for _ in range(4):
            _check(self.lib, self.lib.umab_forces_jvp(self._h, pos.data_ptr(), tangent.data_ptr(), b, None,
                                                      f.data_ptr(), df.data_ptr(), self._stream_ptr()))
            if self._last_call():
                return f, df
        raise RuntimeError("umab: edge capacity exceeded repeatedly")

    def graph(self, pos: torch.Tensor):
        """Neighbour search only -> edge_index [2, E] int64 (row 0 source, row 1 target), CPU."""
        assert pos.is_cuda and pos.dtype == torch.float32 and pos.dim() == 3
        pos = pos.contiguous()
        _check(self.lib, self.lib.umab_build_graph(self._h, pos.data_ptr(), pos.shape[0], self._stream_ptr()))
        nn_, ne_ = ctypes.c_int64(), ctypes.c_int64()
        _check(self.lib, self.lib.umab_graph_counts(self._h, ctypes.byref(nn_), ctypes.byref(ne_)))
        src = torch.empty(ne_.value, dtype=torch.int32, device=pos.device)
        tgt = torch.empty(ne_.value, dtype=torch.int32, device=pos.device)
        _check(self.lib, self.lib.umab_graph_copy(self._h, src.data_ptr(), tgt.data_ptr(), None, self._stream_ptr()))
        torch.cuda.current_stream(self.device).synchronize()
        return torch.stack([src.long(), tgt.long()]).cpu()

    NEIGHBOR_MODES = {"auto": 0, "brute": 1, "cell": 2}

    def set_neighbor_mode(self, mode: str):
        """'auto' (shared-memory cell list from 128 atoms per image), 'brute' or 'cell'; identical edge lists."""
        _check(self.lib, self.lib.umab_set_option(self._h, b"neighbor_mode", self.NEIGHBOR_MODES[mode]))

    def release_workspace(self):
        """Give the per-call device memory (edge workspace, per-layer stores, node state) back; weights stay and the
        buffers re-grow on the next call."""
        if getattr(self, "_h", None):
            _check(self.lib, self.lib.umab_release_workspace(self._h))

    def set_option(self, name: str, value: int):
        """Run-time switches of include/umab.h: "nosync", "cuda_graphs", "neighbor_mode", ..."""
        _check(self.lib, self.lib.umab_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        """Option values and counters: "graph_replays", "graph_captures", "overflow_retries", "edges_per_image_seen"."""
        v = ctypes.c_int64()
        _check(self.lib, self.lib.umab_get_option(self._h, name.encode(), ctypes.byref(v)))
        return v.value

    def graph_counts(self):
        nn_, ne_ = ctypes.c_int64(), ctypes.c_int64()
        _check(self.lib, self.lib.umab_graph_counts(self._h, ctypes.byref(nn_), ctypes.byref(ne_)))
        return nn_.value, ne_.value

    def debug_tensor(self, name: str) -> torch.Tensor:
        ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.lib, self.lib.umab_debug_tensor(self._h, name.encode(), ctypes.byref(ptr), ctypes.byref(n)))
        torch.cuda.synchronize()
        if n.value == 0:
            return torch.empty(0, dtype=torch.float32)
        return torch.as_tensor(_DevPtr(ptr.value, n.value), device=f"cuda:{self.device}").clone().cpu()

    def profile(self, enable: bool):
        """Enable/disable (and reset) per-kernel-family CUDA-event timing."""
        _check(self.lib, self.lib.umab_profile(self._h, int(bool(enable))))

    def profile_read(self):
        """{family: {"ms": device ms, "launches": n, "work": FLOPs (gemm) or bytes}}."""
        out, cat = {}, 0
        while True:
            name = self.lib.umab_profile_name(cat)
            if not name:
                return out
            ms, n, wk, by = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
            _check(self.lib, self.lib.umab_profile_read(self._h, cat, ctypes.byref(ms), ctypes.byref(n),
                                                        ctypes.byref(wk), ctypes.byref(by)))
            out[name.decode()] = {"ms": ms.value, "launches": n.value, "work": wk.value, "bytes": by.value}
            cat += 1

    def stats(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _check(self.lib, self.lib.umab_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return {"kernel_launches": a.value, "device_bytes": b.value}


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, mode: int = 0) -> torch.Tensor:
    """C = A @ W^T (+ bias) through the library's GEMM kernels (unit tests / roofline)."""
    lib = load_library()
    assert a.is_cuda and w.is_cuda and a.dtype == torch.float32 and w.dtype == torch.float32
    a, w = a.contiguous(), w.contiguous()
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    _check(lib, lib.umab_gemm(mode, a.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                              c.data_ptr(), m, n, k, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return c


def gemm_bench(a: torch.Tensor, w: torch.Tensor, mode: int, iters: int = 5) -> float:
    """Kernel-only milliseconds per launch of one GEMM shape (CUDA events inside the library)."""
    lib = load_library()
    a, w = a.contiguous(), w.contiguous()
    m, k = a.shape
    n = w.shape[0]
    c = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ms = ctypes.c_double(0.0)
    _check(lib, lib.umab_gemm_bench(mode, a.data_ptr(), w.data_ptr(), c.data_ptr(), m, n, k, iters, ctypes.byref(ms),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return ms.value
