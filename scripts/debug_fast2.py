import sys, numpy as np, os
sys.path.insert(0, '.')
from fringe_b200 import synth
from fringe_b200.engine import Context
ctx = Context(0)
bands, lines, cols = 20, 4, 8
slc = synth.make_stack(bands, lines, cols, seed=1, region=8, zero_fraction=0)
wts = np.full((lines, cols, 2), 0xffffffff, np.uint32)
print("launching", flush=True)
gpu = ctx.evd_block(slc, wts, 5, 2, method="EVD")
print("done", ctx.evd_stats(), gpu[1], flush=True)
