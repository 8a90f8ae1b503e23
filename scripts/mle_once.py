"""One MLE launch (for ncu): python scripts/mle_once.py [bands lines cols]"""
import sys
import torch
sys.path.insert(0, ".")
from fringe_b200 import synth
from fringe_b200.engine import Context
bands, lines, cols = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (20, 256, 512)
dev = torch.device("cuda", 0); ctx = Context(0)
slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
for _ in range(2):
    out, tcorr, comp = ctx.evd_block_device(slc, wts, 5, 2, method="MLE")
    torch.cuda.synchronize()
print(ctx.last_kernel_ms("evd"), ctx.evd_stats())
