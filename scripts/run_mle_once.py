import sys, torch
sys.path.insert(0, ".")
from fringe_b200 import synth
from fringe_b200.engine import Context
dev = torch.device("cuda", 0); ctx = Context(0)
slc = synth.make_stack_torch(20, 256, 512, seed=2, device=dev)
count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
out, tcorr, comp = ctx.evd_block_device(slc, wts, 5, 2, method="MLE")
torch.cuda.synchronize()
print(ctx.last_kernel_ms("evd"))
