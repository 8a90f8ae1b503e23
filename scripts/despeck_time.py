"""Kernel time of the despeck pass (prep + average) in its four modes on a 300 x 20000 slice, 11x5 and 23x11 windows."""
import sys

import torch

sys.path.insert(0, ".")
from fringe_b200 import synth  # noqa: E402
from fringe_b200.engine import Context  # noqa: E402

ctx = Context(0)
dev = torch.device("cuda", 0)
lines, cols = 300, 20000
slc = synth.make_stack_torch(8, lines, cols, seed=2, device=dev)
for Nx, Ny in ((5, 2), (11, 5)):
    nul = ((2 * Nx + 1) * (2 * Ny + 1) + 31) // 32
    count = torch.empty((lines, cols), dtype=torch.int32, device=dev)
    wts = torch.empty((lines, cols, nul), dtype=torch.int32, device=dev)
    ctx.nmap_block_device(slc, Nx, Ny, "KS2", 0.05, count=count, wts=wts)
    out = torch.zeros((lines, cols), dtype=torch.complex64, device=dev)
    for name, kw in (("amplitude", {}), ("interferogram", dict(z2=slc[3])), ("coherence", dict(z2=slc[3], coherence=True))):
        ts = []
        for rep in range(4):
            ctx.despeck_block_device(slc[0], wts, Nx, Ny, out=out, **kw)
            torch.cuda.synchronize()
            ts.append(ctx.last_kernel_ms("despeck"))
        px = lines * cols
        print(f"window {2*Nx+1}x{2*Ny+1} {name:14s}: {min(ts):.3f} ms per {px/1e6:.0f} M pixels = {px / min(ts) / 1e6:.1f} G px/s; "
              f"HBM view {(px * (8 * (2 if kw else 1) + 4 * nul + 8 + 16)) / min(ts) / 1e6:.0f} GB/s")
