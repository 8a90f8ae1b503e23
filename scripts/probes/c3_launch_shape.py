#!/usr/bin/env python3
"""C3 (100 dates, phase_link): how the generic FP64 kernel's time depends on the number of resident warps, whose
per-warp workspaces (485 kB each at N = 100) live in a global scratch buffer.  Uses the profiling override of
fringe_prof_force_generic (bits 8-15 CTAs per SM, bits 16-23 warps per CTA); prints ms per launch for a 200-line strip."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fringe_b200 import synth
from fringe_b200._lib import lib
from fringe_b200.engine import Context

dev = torch.device("cuda", 0)
lines, cols, bands = 200, 2000, 100
slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
with Context(0) as ctx:
    count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
    for ctas, warps in ((0, 0), (1, 4), (1, 2), (1, 1), (2, 2), (2, 8), (4, 4), (4, 8), (8, 4)):
        lib.fringe_prof_force_generic(ctx._h, (ctas << 8) | (warps << 16))
        for _ in range(2):
            ctx.evd_block_device(slc, wts, 5, 2, "MLE", variant=1, min_neighbors=5)
        torch.cuda.synchronize()
        print(f"CTAs/SM {ctas or 'default(2)'} warps/CTA {warps or 'default(4)'}: {ctx.last_kernel_ms('evd'):9.1f} ms per {lines * cols} px", flush=True)
