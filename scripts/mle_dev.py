"""Development check of the register-row MLE kernel: parity against the oracle on a few band counts,
then C1 / ministack timings (device resident).  usage: python scripts/mle_dev.py [quick]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle  # noqa: E402
from fringe_b200 import synth  # noqa: E402
from fringe_b200.engine import Context  # noqa: E402

o = oracle.load()
ctx = Context(0)
dev = torch.device("cuda", 0)


def compare(name, ref, gpu):
    o_ref, t_ref, c_ref = ref
    o_gpu, t_gpu, c_gpu = gpu
    code_ref = np.where(t_ref < 0, t_ref, 0)
    code_gpu = np.where(t_gpu < 0, t_gpu, 0)
    bad = code_ref != code_gpu
    solved = (t_ref > 0) & ~bad
    dt = np.abs(t_ref - t_gpu)[solved]
    good = solved & (t_ref > 0.3)
    dphi = np.abs(np.angle(o_ref[:, good] * np.conj(o_gpu[:, good])))
    scale = np.abs(c_ref[solved]).max() if solved.any() else 1
    dc = np.abs(c_ref - c_gpu)[good].max() / scale if good.any() else 0
    print(f"{name}: pixels {t_ref.size} solved {int(solved.sum())} sentinel-mismatch {int(bad.sum())} "
          f"codes ref {dict(zip(*np.unique(code_ref, return_counts=True)))} "
          f"max dtcorr {dt.max() if dt.size else 0:.2e} max dphi {dphi.max() if dphi.size else 0:.2e} "
          f"n(dphi>1e-3) {int((dphi > 1e-3).sum())} comp {dc:.2e}", flush=True)
    if bad.sum():
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:5]:
            print("   mismatch at", y, x, "ref", t_ref[y, x], "gpu", t_gpu[y, x])


quick = len(sys.argv) > 1
for bands, variant, kw in [(20, 0, {}), (10, 0, {}), (13, 0, dict(mini_stack_count=4)), (24, 0, {}), (29, 0, dict(mini_stack_count=20)),
                           (32, 0, {}), (20, 1, dict(min_neighbors=5)), (30, 1, dict(min_neighbors=5)), (5, 0, {}), (2, 0, {})]:
    slc = synth.make_stack(bands, 48, 96, seed=bands, region=32)
    wts = o.nmap_block(slc, 5, 2)[1]
    t0 = time.time()
    ref = o.evd_block(slc, wts, 5, 2, method=1, variant=variant, **kw)
    t1 = time.time()
    gpu = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=variant, **kw)
    compare(f"N={bands} variant={variant} {kw} (oracle {t1 - t0:.1f}s)", ref, gpu)
    st = ctx.evd_stats()
    print("   stats", st, flush=True)

for bands, lines, cols in [(20, 512, 512), (10, 512, 512), (29, 256, 512), (32, 256, 512)]:
    slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
    count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
    for generic in ((False, True) if not quick else (False,)):
        if generic:
            os.environ["FRINGE_EVD_GENERIC"] = "1"
        else:
            os.environ.pop("FRINGE_EVD_GENERIC", None)
        for _ in range(3):
            out, tcorr, comp = ctx.evd_block_device(slc, wts, 5, 2, method="MLE")
            torch.cuda.synchronize()
        t = ctx.last_kernel_ms("evd")
        st = ctx.evd_stats()
        print(f"MLE N={bands} {lines}x{cols} generic={generic}: {t:.2f} ms = {lines * cols / t / 1e3:.2f} M px/s "
              f"sentinels {float((tcorr < 0).float().mean()):.3f} stats {st}", flush=True)
    os.environ.pop("FRINGE_EVD_GENERIC", None)
