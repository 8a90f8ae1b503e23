"""Warp-cycle shares of the generic FP64 kernel (MLE), needs python -m fringe_b200.build --phase-clocks."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from fringe_b200 import synth
from fringe_b200.engine import Context
from fringe_b200._lib import lib
dev = torch.device("cuda", 0); ctx = Context(0)
for bands, lines, cols, kw in ((20, 256, 512, dict(method="MLE")), (100, 16, 1024, dict(method="MLE", variant=1, min_neighbors=5))):
    slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
    count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
    for _ in range(2):
        out, tcorr, comp = ctx.evd_block_device(slc, wts, 5, 2, **kw)
        torch.cuda.synchronize()
    cyc = (C.c_int64 * 8)(); lib.fringe_evd_phase_cycles(ctx._h, cyc)
    tot = sum(cyc) or 1
    names = ["covariance", "coherence+gate1", "|C| gate+inverse", "smallest eigenpair", "power iteration", "post"]
    print(bands, "dates:", round(ctx.last_kernel_ms("evd"), 2), "ms", {n: round(c / tot, 3) for n, c in zip(names, cyc)}, ctx.evd_stats(), flush=True)
