import sys, numpy as np, os, faulthandler
faulthandler.dump_traceback_later(25, exit=True)
sys.path.insert(0, '.')
import oracle
from fringe_b200 import synth
from fringe_b200.engine import Context
O = oracle.load(); ctx = Context(0)
print("ctx ok", flush=True)
for bands, lines, cols in ((20, 8, 16), (30, 12, 33), (10, 9, 20)):
    slc = synth.make_stack(bands, lines, cols, seed=1, region=8)
    wts = O.nmap_block(slc, 5, 2)[1]
    print("nmap oracle ok", flush=True)
    ref = O.evd_block(slc, wts, 5, 2, method=0)
    print("evd oracle ok", flush=True)
    gpu = ctx.evd_block(slc, wts, 5, 2, method="EVD")
    print("gpu ok", flush=True)
    st = ctx.evd_stats()
    ok = ref[1] > 0
    dph = np.abs(np.angle(ref[0] * np.conj(gpu[0]))).max(axis=0)
    print(bands, st, 'tcorr maxdiff', np.abs(ref[1] - gpu[1]).max(), 'phase maxdiff', dph[ok & (ref[1] > 0.3)].max() if (ok & (ref[1]>0.3)).any() else None,
          'kernel ms', ctx.last_kernel_ms('evd'), flush=True)
