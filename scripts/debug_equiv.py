import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from fringe_b200 import synth
from fringe_b200.engine import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
NX, NY, BANDS = 5, 2, 30
slc = synth.make_stack_torch(BANDS, 1500, 20000, seed=2, device=dev, row_range=(698, 766))
count = torch.empty((68, 20000), dtype=torch.int32, device=dev)
wts = torch.empty((68, 20000, 2), dtype=torch.int32, device=dev)
ctx.nmap_block_device(slc, NX, NY, "KS2", 0.05, count=count, wts=wts)
o0, t0, _ = ctx.evd_block_device(slc, wts, NX, NY, "EVD", first_line=2, n_lines=64)
rot = slc.clone(); phi = 0.7
rot[7] *= complex(np.cos(phi), np.sin(phi))
o1, t1, _ = ctx.evd_block_device((rot * 3.7).contiguous(), wts, NX, NY, "EVD", first_line=2, n_lines=64)
torch.cuda.synchronize()
ok = t0[2:66] > 0.3
d = torch.angle(o1[:, 2:66] * torch.conj(o0[:, 2:66]))
expect = torch.zeros(BANDS, device=dev); expect[7] = phi
err = torch.angle(torch.exp(1j * (d - expect[:, None, None]))).abs()
err = err * ok[None]
print("entries > 1e-3:", int((err > 1e-3).sum()), "pixels:", int((err > 1e-3).any(0).sum()), "max", float(err.max()))
idx = torch.nonzero(err > 1e-3)
for b, y, x in idx[:10].tolist():
    print("band", b, "pix", (y + 2, x), "err", float(err[b, y, x]), "tcorr", float(t0[y + 2, x]), float(t1[y + 2, x]), "count", int(count[y + 2, x]))
# conditioning of the worst pixel in float64
b, y, x = idx[torch.argmax(err[idx[:, 0], idx[:, 1], idx[:, 2]])].tolist()
y += 2
w = wts[y, x].cpu().numpy().astype(np.uint32)
sl = slc.cpu().numpy()
samples = []
for f in range(55):
    if (w[f >> 5] >> (f & 31)) & 1:
        yy, xx = y + f // 11 - 2, x + f % 11 - 5
        if 0 <= yy < 68 and 0 <= xx < 20000:
            samples.append(sl[:, yy, xx])
Z = np.stack(samples, 1).astype(np.complex128)
C = Z @ Z.conj().T
dd = np.sqrt(np.real(np.diag(C))); C = C / np.outer(dd, dd)
ev, vec = np.linalg.eigh(C)
print("worst pixel", (y, x), "band", b, "nshp", len(samples), "top eigenvalues", ev[-3:], "rel gap", (ev[-1] - ev[-2]) / ev[-1], "|v_band|", abs(vec[b, -1]), "min|v|", np.abs(vec[:, -1]).min())
print(ctx.evd_stats())
