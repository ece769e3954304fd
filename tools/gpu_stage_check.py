"""Run on the GPU box: compare every named intermediate of the CUDA engine (debug mode) with the
staged CPU twin (oracle/staged.py, float32) and with the float64 oracle.  Prints one line per
tensor so a single gpurun call localises a broken kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200.arch import UMAArch, atomic_numbers
from pdb2reaction_b200 import weights as W, synth
from pdb2reaction_b200.engine import UmabEngine
from oracle import uma_ref, staged

def rel(a, b):
    a = a.double().reshape(-1); b = b.double().reshape(-1)
    if a.numel() != b.numel():
        return f"SIZE MISMATCH {a.numel()} vs {b.numel()}"
    d = (a - b).abs().max().item() if a.numel() else 0.0
    s = b.abs().max().item() if b.numel() else 0.0
    return f"maxabs {d:.3e}  scale {s:.3e}  rel {d / (s + 1e-30):.2e}"

def main(n_atoms=24, n_img=2, seed=7, gemm_mode=0, ws=0):
    arch = UMAArch(num_experts=4)
    sd = W.init_uma_weights(arch, 0)
    elem, coords = synth.make_string(n_atoms, n_img, seed)
    z = atomic_numbers(elem)
    m = W.merge_mole(sd, arch, z, 0, 1, "omol")
    hp = uma_ref.Hyper(num_experts=4)
    orc = uma_ref.OracleUMA(m, z, dtype=torch.float64, hyper=hp)
    E64, F64 = orc.energy_forces(coords)
    pos, zz, nat, ei = orc._prep(coords)
    Es, Fs, inter = staged.energy_forces(m, pos.detach().float(), zz, nat, ei, keep=True)
    eng = UmabEngine(m, z, arch, debug=True, gemm_mode=gemm_mode, workspace_bytes=ws)
    pos_d = torch.from_numpy(coords.astype(np.float32)).cuda()
    gi = eng.graph(pos_d)
    print("graph equal:", gi.shape == ei.shape and bool((gi == ei).all()), tuple(gi.shape), tuple(ei.shape))
    e, f = eng.energy_forces(pos_d)
    torch.cuda.synchronize()
    print("E cuda", e.cpu().numpy(), "E staged", Es.numpy(), "E fp64", E64.numpy())
    print("dE/atom vs fp64:", ((e.cpu() - E64).abs().max() / n_atoms).item())
    print("F vs fp64 oracle:", rel(f.cpu(), F64), " | F vs staged32:", rel(f.cpu().reshape(-1, 3), Fs))
    geo = inter["geo"]
    def cmp(name, ref):
        try:
            t = eng.debug_tensor(name)
        except RuntimeError as ex:
            print(f"{name:12s} unavailable ({ex})"); return
        print(f"{name:12s} {rel(t, ref)}")
    cmp("gauss", geo["gauss"]); cmp("env", geo["env"])
    wig = eng.debug_tensor("wig").reshape(-1, 36)[:, :34]
    print(f"{'wig':12s} {rel(wig, geo['wig'])}")
    cmp("x0", inter["x0"])
    for l in range(4):
        d = inter[f"l{l}"]
        cmp(f"l{l}.n1", d["n1"])
        for k in ("rad", "y0", "y1", "y2"):
            cmp(f"l{l}.{k}", d[k])
        cmp(f"l{l}.x1", d["x1"]); cmp(f"l{l}.x", d["x"])
    cmp("node_e", inter["node_e"])
    for l in reversed(range(4)):
        cmp(f"l{l}.g_x", inter[f"l{l}"]["g_x_in"])
    cmp("g_gauss", inter["g_gauss"]); cmp("g_env", inter["g_env"])
    gw = eng.debug_tensor("g_wig").reshape(-1, 36)[:, :34]
    print(f"{'g_wig':12s} {rel(gw, inter['g_wig'])}")
    cmp("g_vec", inter["g_vec"])
    print("stats", eng.stats())

if __name__ == "__main__":
    gm = int(os.environ.get("GEMM_MODE", "0"))
    main(gemm_mode=gm)
    print("---- multi-chunk (tiny workspace) ----")
    main(n_atoms=60, n_img=3, seed=11, gemm_mode=gm, ws=9600 * 4 * 1500)
