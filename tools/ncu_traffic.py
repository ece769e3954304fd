"""Summarise an `ncu --set full` capture of one kernel family into profiles/traffic.json:
average DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) and pipe utilisation.
usage: python tools/ncu_traffic.py <report.ncu-rep | raw.csv> <kernel-key> <algorithmic-bytes-per-launch or 0> <note>"""
import csv, io, json, os, subprocess, sys

rep, key, alg, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
while rows and (not rows[0] or rows[0][0] != "ID"):
    rows.pop(0)
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def col(r, name):
    v = float(r[idx[name]].replace(",", ""))
    return v * scale.get(units[idx[name]], 1.0)


data = [r for r in rows[2:] if len(r) == len(hdr)]
n = len(data)
rd = sum(col(r, "dram__bytes_read.sum") for r in data) / n
wr = sum(col(r, "dram__bytes_write.sum") for r in data) / n
tens = sum(float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) for r in data) / n
dram = sum(float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) for r in data) / n
dur = sum(float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(units[idx["gpu__time_duration.sum"]], 1.0) for r in data) / n
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
allj = json.load(open(out_path)) if os.path.exists(out_path) else {}
allj[key] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
             "algorithmic_bytes_per_launch": alg or None, "launches_captured": n, "ms_per_launch_under_ncu": dur,
             "tensor_pipe_active_pct": tens, "dram_throughput_pct": dram, "report": os.path.basename(rep), "note": note}
json.dump(allj, open(out_path, "w"), indent=1)
print(json.dumps(allj[key], indent=1))
