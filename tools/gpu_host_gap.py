"""GPU: is the step host-bound?  Host time to ENQUEUE one C4 step (library call returns when the last kernel is
launched; one internal sync after the neighbour count) vs the device time of the step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import synth, weights as W
from pdb2reaction_b200.arch import UMAArch, atomic_numbers
from pdb2reaction_b200.engine import UmabEngine

arch = UMAArch()
elem, imgs = synth.make_string(1500, 32, 4)
z = atomic_numbers(elem)
merged = W.merge_mole(W.init_uma_weights(arch, seed=0), arch, z, 0, 1, "omol")
store = int(float(os.environ.get("STORE_GB", "0")) * 1e9)
eng = UmabEngine(merged, z, arch, store_bytes=store)
pos = torch.from_numpy(imgs.astype(np.float32)).cuda()
for _ in range(3):
    eng.energy_forces(pos)
torch.cuda.synchronize()
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    eng.energy_forces(pos)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
    tmp = pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)
    print(f"sm {clk} MHz  {pw:.0f} W  {tmp} C  calls {eng.images_per_call(True)}  ", end="")
    print(f"host enqueue {1e3 * (t1 - t0):7.1f} ms   device {e0.elapsed_time(e1):7.1f} ms   wall {1e3 * (t2 - t0):7.1f} ms", flush=True)
