"""Kernel time and per-phase cycles of k_evd_cta (evd_cta.cu) on a slice of configs[2] (100 dates, phase_link).
Needs the profiling build for the phase shares:  python -m fringe_b200.build --phase-clocks ; python scripts/cta_phases.py"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from fringe_b200 import synth  # noqa: E402
from fringe_b200._lib import lib  # noqa: E402
from fringe_b200.engine import Context  # noqa: E402

bands = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lines, cols = (int(sys.argv[2]) if len(sys.argv) > 2 else 200), 2000
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
ts = []
for rep in range(3):
    ctx.evd_block_device(slc, wts, 5, 2, "MLE", variant=1, min_neighbors=5, out=out, tcorr=tcorr, comp=comp)
    torch.cuda.synchronize()
    ts.append(ctx.last_kernel_ms("evd"))
st = ctx.evd_stats()
px = max(st["pixels"], 1)
print(f"bands {bands}: evd ms {min(ts):.2f} for {px} solved pixels = {px / min(ts) * 1e3:.3e} px/s; iterations/pixel {st['power_iterations'] / px:.1f}")
cyc = (C.c_int64 * 8)()
lib.fringe_evd_phase_cycles(ctx._h, cyc)
if sum(cyc):
    names = ["draw+list", "staging", "covariance", "coherence+|C|", "LDLt", "iteration", "epilogue", "B+G+v"]
    print("   cycles per solved pixel (thread 0):", {n: round(c / px) for n, c in zip(names, cyc)})
