"""Test-only evaluators that can be injected into the calculator (``_backend=``)."""
import numpy as np
import torch


class OracleBackend:
    """The CPU oracle behind the calculator's backend interface (tests only)."""

    def __init__(self, oracle):
        self.oracle = oracle
        self.calls = []

    def evaluate(self, coords_ang, forces=True):
        self.calls.append((coords_ang.shape[0], forces))
        e, f = self.oracle.energy_forces(coords_ang, forces=forces)
        return e.double().numpy(), (None if f is None else f.float().numpy())


class SpringBackend:
    """E = 1/2 k sum_{i<j} (|r_i - r_j| - r0)^2: analytic forces and Hessian are known."""

    def __init__(self, k=3.0, r0=1.1):
        self.k, self.r0 = k, r0
        self.calls = []

    def _e(self, x):
        iu = torch.triu_indices(x.shape[0], x.shape[0], 1)
        d = (x[iu[0]] - x[iu[1]]).norm(dim=1)
        return 0.5 * self.k * ((d - self.r0) ** 2).sum()

    def evaluate(self, coords_ang, forces=True):
        self.calls.append((coords_ang.shape[0], forces))
        es, fs = [], []
        for c in coords_ang.astype(np.float32).astype(np.float64):
            x = torch.tensor(c, requires_grad=True)
            e = self._e(x)
            es.append(e.item())
            if forces:
                g, = torch.autograd.grad(e, x)
                fs.append((-g).numpy().astype(np.float32))
        return np.array(es), (np.stack(fs) if forces else None)

    def hessian(self, coord_ang):
        x = torch.tensor(np.asarray(coord_ang, dtype=np.float64))
        h = torch.autograd.functional.hessian(lambda v: self._e(v.view(-1, 3)), x.reshape(-1))
        return h.numpy()


class SpringBackendAnalytic(SpringBackend):
    """SpringBackend plus the analytic-column entry point of CudaBackend (a backend WITHOUT it makes the calculator fall
    back to finite differences with a warning: test_calculator_contract.py)."""

    def hessian_columns(self, coord_ang, dofs):
        """Analytic-mode stand-in of ``CudaBackend.hessian_columns``: rows k of the exact Hessian at the fp32-rounded
        geometry, as float32 [len(dofs), 3N]."""
        h = self.hessian(np.asarray(coord_ang, dtype=np.float32).astype(np.float64))
        return np.ascontiguousarray(h[[int(k) for k in dofs]], dtype=np.float32)

