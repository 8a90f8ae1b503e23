import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle
from fringe_b200 import synth
from fringe_b200.engine import Context
from test_gpu_sequential import oracle_chain, dilate
o = oracle.load(); ctx = Context(0)
slc = synth.make_stack(200, 64, 512, seed=4, region=64)
wts = o.nmap_block(slc, 5, 2)[1]
res = ctx.sequential_block(slc, wts, 5, 2, 10)
stages, d_out, d_tc, adjusted = oracle_chain(o, slc, wts, 5, 2, 10)
tainted = np.zeros(slc.shape[1:], bool)
for k, (o_ref, t_ref, c_ref) in enumerate(stages):
    t_gpu = res["tcorr_mini"][k]; o_gpu = res["out_mini"][k*10:k*10+o_ref.shape[0]]
    cr, cg = np.where(t_ref < 0, t_ref, 0), np.where(t_gpu < 0, t_gpu, 0)
    bad = cr != cg
    newbad = bad & ~tainted
    tainted |= dilate(bad, 5, 2)
    ok = (t_ref > 0) & (t_gpu > 0) & ~tainted
    good = ok & (t_ref > 0.3)
    d = np.abs(np.angle(o_ref[:, good] * np.conj(o_gpu[:, good])))
    dc = np.abs(c_ref - res["comp"][k])[good]
    print(f"stage {k+1:2d} N={k+10:2d} new code mismatches {int(newbad.sum()):3d} tainted {tainted.mean()*100:5.2f}% good {int(good.sum()):6d} "
          f"phase max {d.max():.2e} q99.9 {np.quantile(d, 0.999):.2e} n>1e-3 {int((d>1e-3).sum())} of {d.size}; dtcorr max {np.abs(t_ref-t_gpu)[ok].max():.2e}; comp max {dc.max():.2e}")
