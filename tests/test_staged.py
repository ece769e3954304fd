"""The hand-written backward the CUDA kernels implement (oracle/staged.py) against autograd of
the oracle, in float64 (derivation check) and float32 (kernel-comparison precision)."""
import torch

from oracle import staged, uma_ref
from pdb2reaction_b200 import synth
from conftest import merged_for


def test_staged_manual_backward_matches_oracle_autograd(state4, arch4, hyper4):
    elem, imgs = synth.make_string(24, 2, 7)
    z, merged = merged_for(state4, arch4, elem)
    orc = uma_ref.OracleUMA(merged, z, dtype=torch.float64, hyper=hyper4)
    e, f = orc.energy_forces(imgs)
    pos, zz, nat, ei = orc._prep(imgs)
    es, fs, _ = staged.energy_forces(merged, pos.detach(), zz, nat, ei)
    # csd is merged in float32 on the host, the oracle recomputes it in float64: 1e-7 relative
    assert (e - es).abs().max() / 24 < 1e-6
    assert (f.reshape(-1, 3) - fs).abs().max() < 1e-6
    es32, fs32, _ = staged.energy_forces(merged, pos.detach().float(), zz, nat, ei)
    assert ((es32.double() - e) / 24).abs().max() < 1e-5
    assert (fs32.double() - f.reshape(-1, 3)).abs().max() < 1e-4


def test_staged_geometry_adjoint_matches_autograd():
    torch.manual_seed(0)
    pos = (torch.randn(12, 3, dtype=torch.float64) * 2.0).requires_grad_(True)
    src = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
    tgt = torch.tensor([1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10])
    geo = staged.geometry_fwd(pos, src, tgt)
    g_gauss, g_env, g_wig = (torch.randn_like(geo[k]) for k in ("gauss", "env", "wig"))
    loss = (geo["gauss"] * g_gauss).sum() + (geo["env"] * g_env).sum() + (geo["wig"] * g_wig).sum()
    gp, = torch.autograd.grad(loss, pos)
    geo_d = {k: v.detach() for k, v in geo.items()}
    g_vec = staged.geometry_bwd(geo_d, g_gauss, g_env, g_wig)
    ref = torch.zeros_like(pos).index_add(0, src, g_vec).index_add(0, tgt, -g_vec)
    assert (gp - ref).abs().max() < 1e-10
