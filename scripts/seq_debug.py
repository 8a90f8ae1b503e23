import sys, numpy as np
sys.path.insert(0, ".")
import oracle
from fringe_b200 import synth
from fringe_b200.engine import Context
o = oracle.load(); ctx = Context(0)
slc = synth.make_stack(200, 32, 256, seed=4, region=64)
wts = o.nmap_block(slc, 5, 2)[1]
res = ctx.sequential_block(slc, wts, 5, 2, 10)
comps = []
for k, d0 in enumerate(range(0, 200, 10), start=1):
    own = slc[d0:d0 + 10]
    bands = np.concatenate([np.array(comps), own]) if comps else own
    bands = np.ascontiguousarray(bands, np.complex64)
    out, tcorr, comp, npix = o.evd_block(bands, wts, 5, 2, method=1, mini_stack_count=k, want_npix=True)
    tg = res["tcorr_mini"][k - 1]
    cr, cg = np.where(tcorr < 0, tcorr, 0), np.where(tg < 0, tg, 0)
    pairs, cnt = np.unique(np.stack([cr.ravel(), cg.ravel()]), axis=1, return_counts=True)
    zero_power = (np.abs(bands) == 0).all(axis=0).sum()
    print(k, bands.shape[0], "code pairs (ref,gpu):", {(float(a), float(b)): int(c) for (a, b), c in zip(pairs.T, cnt)},
          "nan tcorr ref", int(np.isnan(tcorr).sum()), "gpu", int(np.isnan(tg).sum()))
    bad = cr != cg
    if bad.any():
        y, x = np.argwhere(bad)[0]
        print("   first mismatch", y, x, "npix", npix[y, x], "ref", tcorr[y, x], "gpu", tg[y, x],
              "zero bands at pixel", int((bands[:, y, x] == 0).sum()))
    comps.append(res["comp"][k - 1])       # feed OUR compressed SLC forward: stage-wise comparison
