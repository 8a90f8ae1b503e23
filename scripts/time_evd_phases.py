"""Kernel time of k_evd_fast on a 300 x 20000 x 30 slice of the bench stack with the power
iteration capped (FRINGE_EVD_DEBUG_SHORT): separates the covariance + hand-off + epilogue share
from the per-iteration cost.  Timing experiment only; capped runs give wrong phases."""
import os
import sys

import torch

sys.path.insert(0, ".")
from fringe_b200 import synth  # noqa: E402
from fringe_b200.engine import Context  # noqa: E402

lines, cols, bands = 300, 20000, 30
dev = torch.device("cuda", 0)
ctx = Context(0)
slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
count = torch.empty((lines, cols), dtype=torch.int32, device=dev)
wts = torch.empty((lines, cols, 2), dtype=torch.int32, device=dev)
out = torch.zeros((bands, lines, cols), dtype=torch.complex64, device=dev)
tcorr = torch.zeros((lines, cols), dtype=torch.float32, device=dev)
comp = torch.zeros((lines, cols), dtype=torch.complex64, device=dev)
ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05, count=count, wts=wts)
torch.cuda.synchronize()
print("nmap ms", ctx.last_kernel_ms("nmap"), "amp_sort ms", ctx.last_kernel_ms("amp_sort"))
for cap in (sys.argv[1:] or ["1", "5", "9", "0"]):
    if cap == "0":
        os.environ.pop("FRINGE_EVD_DEBUG_SHORT", None)
    else:
        os.environ["FRINGE_EVD_DEBUG_SHORT"] = cap
    ts = []
    for rep in range(3):
        ctx.evd_block_device(slc, wts, 5, 2, "EVD", out=out, tcorr=tcorr, comp=comp)
        torch.cuda.synchronize()
        ts.append(ctx.last_kernel_ms("evd"))
    st = ctx.evd_stats()
    print("cap", cap, "evd ms", min(ts), "its/pixel", st["power_iterations"] / max(st["pixels"], 1), "capped", st["capped"], flush=True)
    import ctypes as C
    from fringe_b200._lib import lib
    cyc = (C.c_int64 * 8)()
    lib.fringe_evd_phase_cycles(ctx._h, cyc)
    tot = sum(cyc) or 1
    if sum(cyc):
        names = ["lists", "covariance", "normalise+handoff", "row load", "iterations", "epilogue"]
        print("   phase share of warp cycles:", {n: round(c / tot, 3) for n, c in zip(names, cyc)}, flush=True)
