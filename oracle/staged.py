"""Stage-by-stage restatement of the CUDA pipeline with a HAND-WRITTEN backward (no autograd).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``oracle/uma_ref.py`` is the oracle proper
(autograd forces, fairchem-style Euler-angle Wigner matrices).  This module mirrors, array for
array, what the sm_100a kernels in ``pdb2reaction_b200/csrc`` compute -- closed-form l<=2
Wigner blocks from the edge unit vector, merged MoLE weights, radial first layer split into a
Gaussian GEMM + per-element tables, explicit adjoints of every stage -- so that

  (1) the hand-derived backward is validated on the CPU against autograd of the oracle, and
  (2) each CUDA kernel can be compared with the intermediate of the same name.

Every function takes/returns torch tensors of one dtype (float64 for derivation checks,
float32 for kernel comparisons).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

TO_M = [0, 2, 6, 3, 7, 1, 5, 8, 4]
L_OF = [0, 1, 1, 1, 2, 2, 2, 2, 2]
SQ3 = math.sqrt(3.0)


def _q_forms(dtype):
    q = torch.zeros(5, 3, 3, dtype=dtype)
    q[0, 0, 2] = q[0, 2, 0] = SQ3 / 2
    q[1, 0, 1] = q[1, 1, 0] = SQ3 / 2
    q[2] = torch.diag(torch.tensor([-0.5, 1.0, -0.5], dtype=dtype))
    q[3, 1, 2] = q[3, 2, 1] = SQ3 / 2
    q[4] = torch.diag(torch.tensor([-SQ3 / 2, 0.0, SQ3 / 2], dtype=dtype))
    return q


# ------------------------------------------------------------------ geometry
def geometry_fwd(pos, src, tgt, cutoff=6.0, nbasis=64):
    """-> dict(vec [E,3], d [E], gauss [E,B], env [E], wig [E,34] = [D1 row-major 9 | D2 row-major 25])."""
    vec = pos[src] - pos[tgt]
    d = vec.pow(2).sum(1).sqrt()
    n = vec / d[:, None]
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    s = (x * x + z * z).sqrt()
    pole = s < 1e-12
    ss = torch.where(pole, torch.ones_like(s), s)
    ca = torch.where(pole, torch.ones_like(s), z / ss)
    sa = torch.where(pole, torch.zeros_like(s), x / ss)
    zero = torch.zeros_like(s)
    # R_e = Rx(-beta) Ry(-alpha): rows (ca, 0, -sa), (x, y, z), (y sa, -s, y ca)
    rot = torch.stack([torch.stack([ca, zero, -sa], 1),
                       torch.stack([x, y, z], 1),
                       torch.stack([y * sa, -s, y * ca], 1)], 1)
    q = _q_forms(pos.dtype)
    rq = torch.einsum("eca,mcd,edb->emab", rot, q, rot)
    d2 = (2.0 / 3.0) * torch.einsum("emab,nab->emn", rq, q)
    wig = torch.cat([rot.reshape(-1, 9), d2.reshape(-1, 25)], 1)
    offs = torch.linspace(0.0, cutoff, nbasis, dtype=pos.dtype)
    coeff = -0.5 / (2.0 * (cutoff / (nbasis - 1))) ** 2
    gauss = torch.exp(coeff * (d[:, None] - offs[None, :]) ** 2)
    u = d / cutoff
    env = 1.0 - 21.0 * u**5 + 35.0 * u**6 - 15.0 * u**7
    env = torch.where(u < 1.0, env, torch.zeros_like(env))
    return dict(vec=vec, d=d, gauss=gauss, env=env, wig=wig)


def geometry_bwd(geo, g_gauss, g_env, g_wig, cutoff=6.0, nbasis=64):
    """Adjoint of geometry_fwd -> g_vec [E,3]."""
    vec, d, gauss = geo["vec"], geo["d"], geo["gauss"]
    dt = vec.dtype
    n = vec / d[:, None]
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    s = (x * x + z * z).sqrt()
    pole = s < 1e-12
    ss = torch.where(pole, torch.ones_like(s), s)
    ca = torch.where(pole, torch.ones_like(s), z / ss)
    sa = torch.where(pole, torch.zeros_like(s), x / ss)
    rot = geo["wig"][:, :9].reshape(-1, 3, 3)
    q = _q_forms(dt)
    g1 = g_wig[:, :9].reshape(-1, 3, 3)
    g2 = g_wig[:, 9:].reshape(-1, 5, 5)
    # dL/dR = G1 + (4/3) sum_mn G2[m,n] Q_m R Q_n
    g_rot = g1 + (4.0 / 3.0) * torch.einsum("emn,mab,ebc,ncd->ead", g2, q, rot, q)
    # R rows: (ca,0,-sa), (x,y,z), (y sa, -s, y ca)
    g_ca = g_rot[:, 0, 0] + y * g_rot[:, 2, 2]
    g_sa = -g_rot[:, 0, 2] + y * g_rot[:, 2, 0]
    g_s = -g_rot[:, 2, 1]
    gx = g_rot[:, 1, 0].clone()
    gy = g_rot[:, 1, 1] + sa * g_rot[:, 2, 0] + ca * g_rot[:, 2, 2]
    gz = g_rot[:, 1, 2].clone()
    # ca = z/s, sa = x/s, s = sqrt(x^2+z^2)
    inv_s = torch.where(pole, torch.zeros_like(s), 1.0 / ss)
    g_s_tot = g_s - (g_ca * ca + g_sa * sa) * inv_s
    gx = gx + g_sa * inv_s + g_s_tot * sa
    gz = gz + g_ca * inv_s + g_s_tot * ca
    g_n = torch.stack([gx, gy, gz], 1)
    # n = vec/d
    g_vec = (g_n - n * (g_n * n).sum(1, keepdim=True)) / d[:, None]
    # radial part
    offs = torch.linspace(0.0, cutoff, nbasis, dtype=dt)
    coeff = -0.5 / (2.0 * (cutoff / (nbasis - 1))) ** 2
    g_d = (g_gauss * gauss * (2.0 * coeff) * (d[:, None] - offs[None, :])).sum(1)
    u = d / cutoff
    denv = (-105.0 * u**4 + 210.0 * u**5 - 105.0 * u**6) / cutoff
    denv = torch.where(u < 1.0, denv, torch.zeros_like(denv))
    g_d = g_d + g_env * denv
    return g_vec + g_d[:, None] * n


def wig_full(wig):
    """[E,34] -> block-diagonal [E,9,9] (l-primary)."""
    e = wig.shape[0]
    w = wig.new_zeros(e, 9, 9)
    w[:, 0, 0] = 1.0
    w[:, 1:4, 1:4] = wig[:, :9].reshape(e, 3, 3)
    w[:, 4:9, 4:9] = wig[:, 9:].reshape(e, 5, 5)
    return w


def wig_grad_pack(g_full):
    return torch.cat([g_full[:, 1:4, 1:4].reshape(-1, 9), g_full[:, 4:9, 4:9].reshape(-1, 25)], 1)


# ------------------------------------------------------------------ small math
def silu(x):
    return x * torch.sigmoid(x)


def dsilu(x):
    s = torch.sigmoid(x)
    return s * (1.0 + x * (1.0 - s))


def ln_silu_fwd(u, gamma, beta, eps=1e-5):
    mu = u.mean(1, keepdim=True)
    var = (u - mu).pow(2).mean(1, keepdim=True)
    xh = (u - mu) * (var + eps).rsqrt()
    return silu(xh * gamma + beta)


def ln_silu_bwd(u, gamma, beta, g_out, eps=1e-5):
    mu = u.mean(1, keepdim=True)
    var = (u - mu).pow(2).mean(1, keepdim=True)
    rstd = (var + eps).rsqrt()
    xh = (u - mu) * rstd
    g_ln = g_out * dsilu(xh * gamma + beta)
    g_xh = g_ln * gamma
    return rstd * (g_xh - g_xh.mean(1, keepdim=True) - xh * (g_xh * xh).mean(1, keepdim=True))


def radial_tables(w, name, nbasis=64, ce=128):
    """Split lin1 = [W_gauss | W_src | W_tgt]: per-element tables T = emb @ W_part^T."""
    w1 = w[name + ".lin1.weight"]
    t_src = w["source_embedding.weight"] @ w1[:, nbasis:nbasis + ce].T
    t_tgt = w["target_embedding.weight"] @ w1[:, nbasis + ce:].T
    return w1[:, :nbasis].contiguous(), t_src, t_tgt


def radial_fwd(w, name, gauss, zsrc, ztgt):
    w1g, t_src, t_tgt = radial_tables(w, name, gauss.shape[1])
    u1 = gauss @ w1g.T + t_src[zsrc] + t_tgt[ztgt] + w[name + ".lin1.bias"]
    h1 = ln_silu_fwd(u1, w[name + ".ln1.weight"], w[name + ".ln1.bias"])
    u2 = h1 @ w[name + ".lin2.weight"].T + w[name + ".lin2.bias"]
    h2 = ln_silu_fwd(u2, w[name + ".ln2.weight"], w[name + ".ln2.bias"])
    rad = h2 @ w[name + ".lin3.weight"].T + w[name + ".lin3.bias"]
    return rad, dict(u1=u1, h1=h1, u2=u2, h2=h2)


def radial_bwd(w, name, saved, g_rad, nbasis=64):
    g_h2 = g_rad @ w[name + ".lin3.weight"]
    g_u2 = ln_silu_bwd(saved["u2"], w[name + ".ln2.weight"], w[name + ".ln2.bias"], g_h2)
    g_h1 = g_u2 @ w[name + ".lin2.weight"]
    g_u1 = ln_silu_bwd(saved["u1"], w[name + ".ln1.weight"], w[name + ".ln1.bias"], g_h1)
    return g_u1 @ w[name + ".lin1.weight"][:, :nbasis]          # g_gauss


# ------------------------------------------------------------------ node ops
def bal_w(dtype):
    return torch.tensor([1.0 / ((2 * l + 1) * 3.0) for l in L_OF], dtype=dtype)


def rms_fwd(x, w_aff, b_aff, add0=None, eps=1e-5):
    f = x.clone()
    f[:, 0, :] -= x[:, 0, :].mean(1, keepdim=True)
    s = (f.pow(2) * bal_w(x.dtype).view(1, 9, 1)).sum(1).mean(1)
    r = (s + eps).rsqrt()
    y = f * r.view(-1, 1, 1) * w_aff[L_OF].unsqueeze(0)
    y[:, 0, :] += b_aff
    if add0 is not None:
        y[:, 0, :] += add0
    return y


def rms_bwd(x, w_aff, g_y, eps=1e-5):
    c = x.shape[2]
    f = x.clone()
    f[:, 0, :] -= x[:, 0, :].mean(1, keepdim=True)
    bw = bal_w(x.dtype).view(1, 9, 1)
    s = (f.pow(2) * bw).sum(1).mean(1)
    r = (s + eps).rsqrt()
    gw = g_y * w_aff[L_OF].unsqueeze(0)
    dot = (gw * f).sum((1, 2))
    g_f = gw * r.view(-1, 1, 1) - (r.pow(3) * dot / c).view(-1, 1, 1) * bw * f
    g_x = g_f.clone()
    g_x[:, 0, :] -= g_f[:, 0, :].mean(1, keepdim=True)
    return g_x


def so3_lin_fwd(x, weight, bias):
    y = torch.einsum("nmi,moi->nmo", x, weight[L_OF])
    y[:, 0, :] += bias
    return y


def so3_lin_bwd(g_y, weight):
    return torch.einsum("nmo,moi->nmi", g_y, weight[L_OF])


GATE_L = [0, 0, 0, 1, 1, 1, 1, 1]


def ffn_fwd(w, p, x):
    h = w[p + ".ffn.so3_1.bias"].shape[0]
    gp = x[:, 0, :] @ w[p + ".ffn.scalar_mlp.weight"].T + w[p + ".ffn.scalar_mlp.bias"]
    gs = silu(gp)
    gate = torch.sigmoid(gs).reshape(-1, 2, h)
    y1 = so3_lin_fwd(x, w[p + ".ffn.so3_1.weight"], w[p + ".ffn.so3_1.bias"])
    a = torch.cat([silu(y1[:, 0:1, :]), y1[:, 1:, :] * gate[:, GATE_L, :]], 1)
    y2 = so3_lin_fwd(a, w[p + ".ffn.so3_2.weight"], w[p + ".ffn.so3_2.bias"])
    return y2, dict(gp=gp, y1=y1)


def ffn_bwd(w, p, x, saved, g_y2):
    h = w[p + ".ffn.so3_1.bias"].shape[0]
    gp, y1 = saved["gp"], saved["y1"]
    gs = silu(gp)
    sg = torch.sigmoid(gs)
    gate = sg.reshape(-1, 2, h)
    g_a = so3_lin_bwd(g_y2, w[p + ".ffn.so3_2.weight"])
    g_y1 = torch.cat([g_a[:, 0:1, :] * dsilu(y1[:, 0:1, :]), g_a[:, 1:, :] * gate[:, GATE_L, :]], 1)
    g_gate = torch.zeros_like(gate)
    g_gate.index_add_(1, torch.tensor(GATE_L), g_a[:, 1:, :] * y1[:, 1:, :])
    g_gs = g_gate.reshape(-1, 2 * h) * sg * (1.0 - sg)
    g_gp = g_gs * dsilu(gp)
    g_x = so3_lin_bwd(g_y1, w[p + ".ffn.so3_1.weight"])
    g_x[:, 0, :] += g_gp @ w[p + ".ffn.scalar_mlp.weight"]
    return g_x


# ------------------------------------------------------------------ edgewise stages
def gather_rotate_scale(x, src, tgt, wig, rad):
    """-> A0 [E,3,2C], A1 [E,2,2,2C], A2 [E,2,1,2C] (already multiplied by the radial weights)."""
    e = src.shape[0]
    c = x.shape[2]
    d = wig_full(wig)[:, TO_M, :]                                  # to_m . D
    msg = torch.bmm(d, torch.cat([x[src], x[tgt]], 2))             # [E,9,2C] m-primary
    a0 = msg[:, 0:3, :] * rad[:, : 6 * c].reshape(e, 3, 2 * c)
    a1 = msg[:, 3:7, :].reshape(e, 2, 2, 2 * c) * rad[:, 6 * c:10 * c].reshape(e, 1, 2, 2 * c)
    a2 = msg[:, 7:9, :].reshape(e, 2, 1, 2 * c) * rad[:, 10 * c:12 * c].reshape(e, 1, 1, 2 * c)
    return a0, a1, a2, msg


def so2_complex_weight(w):
    """SO(2) m>0 weight W [2*ho, i] (rows: real | imaginary outputs) -> Wc [2*ho, 2*i] = [[W_r, -W_i], [W_i, W_r]].

    fairchem's SO2_m_Conv applies W to the +m and the -m rows separately and then combines
    (o_r = y_r(+) - y_i(-), o_i = y_r(-) + y_i(+)).  Folding the combination into the weight gives the same
    FLOPs in ONE contraction  [x(+m) | x(-m)] @ Wc^T = [o_r | o_i]  and halves the conv output the kernels have to
    write and re-read (this is what the CUDA pipeline does; engine.prepare_engine_weights builds the same Wc)."""
    ho = w.shape[0] // 2
    wr, wi = w[:ho], w[ho:]
    return torch.cat([torch.cat([wr, -wi], 1), torch.cat([wi, wr], 1)], 0)


def so2_split(ym, half):
    """ym [E, 2*half] = [o_r | o_i] -> (o_r, o_i)."""
    return ym[:, :half], ym[:, half:]


def combine_gate_fwd(y0, y1, y2, h):
    """conv1 outputs -> conv2 inputs B0 [E,3,H], B1 [E,2,2,H], B2 [E,2,1,H]."""
    e = y0.shape[0]
    g = torch.sigmoid(y0[:, : 2 * h]).reshape(e, 2, h)
    t = y0[:, 2 * h:].reshape(e, 3, h)
    b0 = torch.stack([silu(t[:, 0]), t[:, 1] * g[:, 0], t[:, 2] * g[:, 1]], 1)
    o_r, o_i = so2_split(y1, 2 * h)
    b1 = torch.stack([o_r.reshape(e, 2, h) * g, o_i.reshape(e, 2, h) * g], 1)
    p_r, p_i = so2_split(y2, h)
    b2 = torch.stack([p_r * g[:, 1], p_i * g[:, 1]], 1).reshape(e, 2, 1, h)
    return b0, b1, b2


def combine_gate_bwd(y0, y1, y2, h, g_b0, g_b1, g_b2):
    e = y0.shape[0]
    sg = torch.sigmoid(y0[:, : 2 * h]).reshape(e, 2, h)
    t = y0[:, 2 * h:].reshape(e, 3, h)
    o_r, o_i = so2_split(y1, 2 * h)
    o_r, o_i = o_r.reshape(e, 2, h), o_i.reshape(e, 2, h)
    p_r, p_i = so2_split(y2, h)
    g_gate = torch.zeros_like(sg)
    g_gate[:, 0] = g_b0[:, 1] * t[:, 1] + g_b1[:, 0, 0] * o_r[:, 0] + g_b1[:, 1, 0] * o_i[:, 0]
    g_gate[:, 1] = (g_b0[:, 2] * t[:, 2] + g_b1[:, 0, 1] * o_r[:, 1] + g_b1[:, 1, 1] * o_i[:, 1]
                    + g_b2[:, 0, 0] * p_r + g_b2[:, 1, 0] * p_i)
    g_t = torch.stack([g_b0[:, 0] * dsilu(t[:, 0]), g_b0[:, 1] * sg[:, 0], g_b0[:, 2] * sg[:, 1]], 1)
    g_y0 = torch.cat([(g_gate * sg * (1 - sg)).reshape(e, 2 * h), g_t.reshape(e, 3 * h)], 1)
    g_y1 = torch.cat([(g_b1[:, 0] * sg).reshape(e, 2 * h), (g_b1[:, 1] * sg).reshape(e, 2 * h)], 1)
    g_y2 = torch.cat([g_b2[:, 0, 0] * sg[:, 1], g_b2[:, 1, 0] * sg[:, 1]], 1)
    return g_y0, g_y1, g_y2


def z_rows(z0, z1, z2, c):
    """conv2 outputs -> message [E,9,C] in m-primary order."""
    e = z0.shape[0]
    o_r, o_i = so2_split(z1, 2 * c)
    p_r, p_i = so2_split(z2, c)
    return torch.cat([z0.reshape(e, 3, c), o_r.reshape(e, 2, c), o_i.reshape(e, 2, c),
                      p_r.reshape(e, 1, c), p_i.reshape(e, 1, c)], 1)


def z_rows_bwd(g_zm, c):
    e = g_zm.shape[0]
    g_z0 = g_zm[:, 0:3].reshape(e, 3 * c)
    g_z1 = torch.cat([g_zm[:, 3:5].reshape(e, 2 * c), g_zm[:, 5:7].reshape(e, 2 * c)], 1)
    g_z2 = torch.cat([g_zm[:, 7].reshape(e, c), g_zm[:, 8].reshape(e, c)], 1)
    return g_z0, g_z1, g_z2


def rotate_back_reduce(zm, wig, env, tgt, n_nodes, scale=1.0):
    """out[i] = sum_{e -> i} scale * env_e * D_e^T (to_m^T z_e)."""
    zl = torch.zeros_like(zm)
    zl[:, TO_M, :] = zm
    y = torch.bmm(wig_full(wig).transpose(1, 2), zl) * (env * scale).view(-1, 1, 1)
    return zm.new_zeros(n_nodes, 9, zm.shape[2]).index_add(0, tgt, y)


def rotate_back_bwd(zm, wig, env, tgt, g_out, scale=1.0):
    """-> g_zm [E,9,C], g_env [E], g_wig [E,34]."""
    d = wig_full(wig)
    g = g_out[tgt]                                               # [E,9,C] l-primary
    zl = torch.zeros_like(zm)
    zl[:, TO_M, :] = zm
    g_zl = torch.bmm(d, g) * (env * scale).view(-1, 1, 1)
    g_zm = g_zl[:, TO_M, :]
    g_env = scale * (torch.bmm(d.transpose(1, 2), zl) * g).sum((1, 2))
    # y_a = env sum_b D[b,a] zl_b  ->  dD[b,a] = env sum_c zl[b,c] g[a,c]
    g_d = torch.bmm(zl, g.transpose(1, 2)) * (env * scale).view(-1, 1, 1)
    return g_zm, g_env, wig_grad_pack(g_d)


def edgewise_fwd(w, p, x, geo, src, tgt, zsrc, ztgt, c, h):
    rad, rs = radial_fwd(w, p + ".edge.conv1.rad", geo["gauss"], zsrc, ztgt)
    a0, a1, a2, _ = gather_rotate_scale(x, src, tgt, geo["wig"], rad)
    e = src.shape[0]
    y0 = a0.reshape(e, -1) @ w[p + ".edge.conv1.fc_m0.weight"].T + w[p + ".edge.conv1.fc_m0.bias"]
    y1 = a1.reshape(e, -1) @ so2_complex_weight(w[p + ".edge.conv1.fc_m1.weight"]).T      # [E, 4H] = [o_r | o_i]
    y2 = a2.reshape(e, -1) @ so2_complex_weight(w[p + ".edge.conv1.fc_m2.weight"]).T      # [E, 2H]
    b0, b1, b2 = combine_gate_fwd(y0, y1, y2, h)
    z0 = b0.reshape(e, -1) @ w[p + ".edge.conv2.fc_m0.weight"].T + w[p + ".edge.conv2.fc_m0.bias"]
    z1 = b1.reshape(e, -1) @ so2_complex_weight(w[p + ".edge.conv2.fc_m1.weight"]).T
    z2 = b2.reshape(e, -1) @ so2_complex_weight(w[p + ".edge.conv2.fc_m2.weight"]).T
    zm = z_rows(z0, z1, z2, c)
    out = rotate_back_reduce(zm, geo["wig"], geo["env"], tgt, x.shape[0])
    saved = dict(rad=rad, rs=rs, y0=y0, y1=y1, y2=y2, zm=zm)
    return out, saved


def edgewise_bwd(w, p, x, geo, src, tgt, saved, g_out, c, h):
    """-> g_x [N,9,C], g_gauss [E,B], g_env [E], g_wig [E,34]."""
    e = src.shape[0]
    g_zm, g_env, g_wig = rotate_back_bwd(saved["zm"], geo["wig"], geo["env"], tgt, g_out)
    g_z0, g_z1, g_z2 = z_rows_bwd(g_zm, c)
    g_b0 = (g_z0 @ w[p + ".edge.conv2.fc_m0.weight"]).reshape(e, 3, h)
    g_b1 = (g_z1 @ so2_complex_weight(w[p + ".edge.conv2.fc_m1.weight"])).reshape(e, 2, 2, h)
    g_b2 = (g_z2 @ so2_complex_weight(w[p + ".edge.conv2.fc_m2.weight"])).reshape(e, 2, 1, h)
    g_y0, g_y1, g_y2 = combine_gate_bwd(saved["y0"], saved["y1"], saved["y2"], h, g_b0, g_b1, g_b2)
    g_a0 = (g_y0 @ w[p + ".edge.conv1.fc_m0.weight"]).reshape(e, 3, 2 * c)
    g_a1 = (g_y1 @ so2_complex_weight(w[p + ".edge.conv1.fc_m1.weight"])).reshape(e, 2, 2, 2 * c)
    g_a2 = (g_y2 @ so2_complex_weight(w[p + ".edge.conv1.fc_m2.weight"])).reshape(e, 2, 1, 2 * c)
    # gather-rotate-scale adjoint
    rad = saved["rad"]
    dm = wig_full(geo["wig"])[:, TO_M, :]
    xcat = torch.cat([x[src], x[tgt]], 2)
    msg = torch.bmm(dm, xcat)
    r0 = rad[:, : 6 * c].reshape(e, 3, 2 * c)
    r1 = rad[:, 6 * c:10 * c].reshape(e, 1, 2, 2 * c)
    r2 = rad[:, 10 * c:].reshape(e, 1, 1, 2 * c)
    g_rad = torch.cat([(g_a0 * msg[:, 0:3]).reshape(e, -1),
                       (g_a1 * msg[:, 3:7].reshape(e, 2, 2, 2 * c)).sum(1).reshape(e, -1),
                       (g_a2 * msg[:, 7:9].reshape(e, 2, 1, 2 * c)).sum(1).reshape(e, -1)], 1)
    g_msg = torch.cat([g_a0 * r0, (g_a1 * r1).reshape(e, 4, 2 * c), (g_a2 * r2).reshape(e, 2, 2 * c)], 1)
    g_dm = torch.bmm(g_msg, xcat.transpose(1, 2))                 # [E,9(m),9(l)]
    g_d = torch.zeros_like(g_dm)
    g_d[:, TO_M, :] = g_dm
    g_wig = g_wig + wig_grad_pack(g_d)
    g_xcat = torch.bmm(dm.transpose(1, 2), g_msg)                 # [E,9,2C]
    g_x = torch.zeros_like(x)
    g_x.index_add_(0, src, g_xcat[:, :, :c])
    g_x.index_add_(0, tgt, g_xcat[:, :, c:])
    g_gauss = radial_bwd(w, p + ".edge.conv1.rad", saved["rs"], g_rad, geo["gauss"].shape[1])
    return g_x, g_gauss, g_env, g_wig


# ------------------------------------------------------------------ full pipeline
def energy_forces(w: Dict[str, torch.Tensor], pos, z, natoms, edge_index, *, num_layers=4,
                  cutoff=6.0, rescale=5.0, keep=False):
    """Merged weights (incl. ``csd``) -> (E [n_img], F [N,3], intermediates if keep)."""
    dt = pos.dtype
    w = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in w.items()}
    c = w["sphere_embedding.weight"].shape[1]
    h = w["head.0.weight"].shape[0]
    nb = w["edge_degree.rad.lin1.weight"].shape[1] - 2 * w["source_embedding.weight"].shape[1]
    src, tgt = edge_index[0], edge_index[1]
    zsrc, ztgt = z[src], z[tgt]
    n = pos.shape[0]
    csd = w["csd"]
    geo = geometry_fwd(pos, src, tgt, cutoff, nb)
    inter = {"geo": geo}

    x = pos.new_zeros(n, 9, c)
    x[:, 0, :] = w["sphere_embedding.weight"][z] + csd
    rad_ed, rs_ed = radial_fwd(w, "edge_degree.rad", geo["gauss"], zsrc, ztgt)
    zm_ed = torch.cat([rad_ed.reshape(-1, 3, c), rad_ed.new_zeros(rad_ed.shape[0], 6, c)], 1)
    x = x + rotate_back_reduce(zm_ed, geo["wig"], geo["env"], tgt, n, 1.0 / rescale)
    inter["x0"] = x

    xs, n1s, x1s, n2s, esaved, fsaved = [], [], [], [], [], []
    for l in range(num_layers):
        p = f"blocks.{l}"
        xs.append(x)
        n1 = rms_fwd(x, w[p + ".norm_1.affine_weight"], w[p + ".norm_1.affine_bias"], csd)
        e_out, es = edgewise_fwd(w, p, n1, geo, src, tgt, zsrc, ztgt, c, h)
        x1 = x + e_out
        n2 = rms_fwd(x1, w[p + ".norm_2.affine_weight"], w[p + ".norm_2.affine_bias"])
        f_out, fs = ffn_fwd(w, p, n2)
        x = x1 + f_out
        n1s.append(n1); x1s.append(x1); n2s.append(n2); esaved.append(es); fsaved.append(fs)
        if keep:
            inter[f"l{l}"] = dict(n1=n1, e_out=e_out, x1=x1, n2=n2, f_out=f_out, x=x, **es)
    xf = rms_fwd(x, w["norm.affine_weight"], w["norm.affine_bias"])
    s0 = xf[:, 0, :]
    p1 = s0 @ w["head.0.weight"].T + w["head.0.bias"]
    p2 = silu(p1) @ w["head.2.weight"].T + w["head.2.bias"]
    node_e = (silu(p2) @ w["head.4.weight"].T + w["head.4.bias"]).view(-1)
    img = torch.repeat_interleave(torch.arange(len(natoms)), torch.as_tensor(list(natoms)))
    e_img = node_e.new_zeros(len(natoms)).index_add(0, img, node_e)
    inter["node_e"] = node_e

    # ---------------- backward (dE_total/dpos)
    g_p2 = w["head.4.weight"].view(1, -1) * dsilu(p2)
    g_p1 = (g_p2 @ w["head.2.weight"]) * dsilu(p1)
    g_xf = torch.zeros_like(xf)
    g_xf[:, 0, :] = g_p1 @ w["head.0.weight"]
    g_x = rms_bwd(x, w["norm.affine_weight"], g_xf)
    g_gauss = torch.zeros_like(geo["gauss"])
    g_env = torch.zeros_like(geo["env"])
    g_wig = torch.zeros_like(geo["wig"])
    for l in reversed(range(num_layers)):
        p = f"blocks.{l}"
        g_n2 = ffn_bwd(w, p, n2s[l], fsaved[l], g_x)
        g_x1 = g_x + rms_bwd(x1s[l], w[p + ".norm_2.affine_weight"], g_n2)
        g_n1, gg, ge, gw = edgewise_bwd(w, p, n1s[l], geo, src, tgt, esaved[l], g_x1, c, h)
        g_gauss += gg; g_env += ge; g_wig += gw
        g_x = g_x1 + rms_bwd(xs[l], w[p + ".norm_1.affine_weight"], g_n1)
        if keep:
            inter[f"l{l}"]["g_x_in"] = g_x
    # edge-degree embedding adjoint
    g_zm, ge, gw = rotate_back_bwd(zm_ed, geo["wig"], geo["env"], tgt, g_x, 1.0 / rescale)
    g_env += ge; g_wig += gw
    g_gauss += radial_bwd(w, "edge_degree.rad", rs_ed, g_zm[:, 0:3].reshape(-1, 3 * c), nb)
    g_vec = geometry_bwd(geo, g_gauss, g_env, g_wig, cutoff, nb)
    g_pos = torch.zeros_like(pos)
    g_pos.index_add_(0, src, g_vec)
    g_pos.index_add_(0, tgt, -g_vec)
    inter.update(g_gauss=g_gauss, g_env=g_env, g_wig=g_wig, g_vec=g_vec)
    return e_img, -g_pos, inter
