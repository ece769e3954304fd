"""Per-kernel roofline table from an `ncu --set full` raw CSV (ncu -i rep --page raw --csv > raw.csv).

For every kernel name: launches captured, mean duration, DRAM bytes read+written per launch, achieved
DRAM GB/s (bytes / duration), fraction of the measured HBM peak (MEASURED_PEAKS.json hbm_gbs, else the
6548.5 GB/s of this pool), tensor-pipe active %, occupancy, registers.
usage: python tools/ncu_kernel_table.py raw.csv out.json [out.md] [note]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "%": 1.0}


def main():
    raw, out_json = sys.argv[1], sys.argv[2]
    out_md = sys.argv[3] if len(sys.argv) > 3 else None
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    peak = 6548.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = list(csv.reader(open(raw)))
    while rows and (not rows[0] or rows[0][0] != "ID"):
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def col(r, name, default=None):
        if name not in idx:
            return default
        try:
            return float(r[idx[name]].replace(",", "")) * SCALE.get(units[idx[name]], 1.0)
        except ValueError:
            return default

    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "")
        a = agg.setdefault(name, collections.defaultdict(float))
        a["n"] += 1
        a["t"] += col(r, "gpu__time_duration.sum", 0.0)
        a["rd"] += col(r, "dram__bytes_read.sum", 0.0)
        a["wr"] += col(r, "dram__bytes_write.sum", 0.0)
        a["tensor"] += col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) or 0.0
        a["dram_pct"] += col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) or 0.0
        a["occ"] += col(r, "sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) or 0.0
        a["l1"] += col(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) or 0.0
        a["l2"] += col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) or 0.0
        a["regs"] = col(r, "launch__registers_per_thread", 0.0)
    table = []
    for name, a in agg.items():
        n = a["n"]
        t = a["t"] / n
        by = (a["rd"] + a["wr"]) / n
        table.append({
            "kernel": name, "launches_captured": int(n), "ms_per_launch_under_ncu": t * 1e3,
            "dram_read_bytes": a["rd"] / n, "dram_write_bytes": a["wr"] / n,
            "achieved_dram_gbs": by / t / 1e9 if t > 0 else None,
            "frac_of_hbm_peak": by / t / 1e9 / peak if t > 0 else None,
            "ncu_dram_throughput_pct": a["dram_pct"] / n, "tensor_pipe_active_pct": a["tensor"] / n,
            "l1tex_throughput_pct": a["l1"] / n, "lts_throughput_pct": a["l2"] / n,
            "warps_active_pct": a["occ"] / n, "registers_per_thread": int(a["regs"] or 0)})
    table.sort(key=lambda d: -d["ms_per_launch_under_ncu"] * d["launches_captured"])
    json.dump({"note": note, "hbm_peak_gbs": peak, "source": os.path.basename(raw), "kernels": table},
              open(out_json, "w"), indent=1)
    if out_md:
        with open(out_md, "w") as f:
            f.write(f"# ncu --set full per-kernel summary\n\n{note}\n\nHBM peak used: {peak:.1f} GB/s (measured copy bandwidth).\n\n")
            f.write("| kernel | n | ms/launch | DRAM rd GB | DRAM wr GB | achieved GB/s | frac of HBM peak | tensor pipe % | l1tex % | L2 % | warps active % | regs |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for d in table:
                f.write("| {kernel} | {launches_captured} | {ms_per_launch_under_ncu:.3f} | {r:.3f} | {w:.3f} | {g:.0f} | {fr:.2f} | "
                        "{tensor_pipe_active_pct:.1f} | {l1tex_throughput_pct:.1f} | {lts_throughput_pct:.1f} | {warps_active_pct:.1f} | "
                        "{registers_per_thread} |\n".format(r=d["dram_read_bytes"] / 1e9, w=d["dram_write_bytes"] / 1e9,
                                                              g=d["achieved_dram_gbs"] or 0, fr=d["frac_of_hbm_peak"] or 0, **d))
    for d in table:
        print(f"{d['kernel'][:44]:44s} n={d['launches_captured']:3d} {d['ms_per_launch_under_ncu']:8.3f} ms "
              f"{d['achieved_dram_gbs'] or 0:7.0f} GB/s ({100 * (d['frac_of_hbm_peak'] or 0):5.1f}% HBM) tensor {d['tensor_pipe_active_pct']:5.1f}%")


if __name__ == "__main__":
    main()
