"""Kernel times of the other BASELINE.json configs (parity-test cases, not the bench line), device
resident, on strips small enough to finish in seconds:
    C1  20 dates 512 x 512, KS2 11x5 -> evd EVD and MLE
    C3  100 dates, phase_link MLE, min_neighbors 5 (strip 32 x 2048)
    C5  30 dates, AD2 21x21 (strip 256 x 2048) -> evd EVD
    C2' 30 dates, KS2 with the half-window reading of docs/workflows.md (23 x 11 window, strip 128 x 4096)
usage: python scripts/bench_configs.py"""
import sys

import torch

sys.path.insert(0, ".")
from fringe_b200 import synth  # noqa: E402
from fringe_b200.engine import Context  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0)


def run(name, bands, lines, cols, Nx, Ny, nmap_method, evd_kwargs_list, reps=2):
    slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=dev)
    for _ in range(reps):
        count, wts = ctx.nmap_block_device(slc, Nx, Ny, nmap_method, 0.05)
        torch.cuda.synchronize()
    npx = lines * cols
    t_sort, t_nmap = ctx.last_kernel_ms("amp_sort"), ctx.last_kernel_ms("nmap")
    print(f"{name}: {bands} dates {lines}x{cols} window {2*Nx+1}x{2*Ny+1} {nmap_method}: amp_sort {t_sort:.2f} ms, "
          f"nmap {t_nmap:.2f} ms = {npx / t_nmap / 1e3:.1f} M px/s, mean SHP {float(count.float().mean()):.1f}", flush=True)
    for kw in evd_kwargs_list:
        for _ in range(reps):
            out, tcorr, comp = ctx.evd_block_device(slc, wts, Nx, Ny, **kw)
            torch.cuda.synchronize()
        t = ctx.last_kernel_ms("evd")
        st = ctx.evd_stats()
        neg = float((tcorr < 0).float().mean())
        print(f"    evd {kw}: {t:.2f} ms = {npx / t / 1e3:.2f} M px/s, solved {st['pixels']}, fp64 pixels {st['fp64_pixels']}, "
              f"sentinel fraction {neg:.4f}", flush=True)


run("C1", 20, 512, 512, 5, 2, "KS2", [dict(method="EVD"), dict(method="MLE")])
run("C3", 100, 32, 2048, 5, 2, "KS2", [dict(method="MLE", variant=1, min_neighbors=5)])
run("C5", 30, 256, 2048, 10, 10, "AD2", [dict(method="EVD")])
run("C2'", 30, 128, 4096, 11, 5, "KS2", [dict(method="EVD")])

# datum adjustment product: HBM-bound, 24 bytes per pixel
a = torch.randn(1500 * 20000, dtype=torch.complex64, device=dev)
b = torch.randn(1500 * 20000, dtype=torch.complex64, device=dev)
o = torch.empty_like(a)
for _ in range(3):
    ctx.cmul_device(a, b, out=o)
    torch.cuda.synchronize()
t = ctx.last_kernel_ms("cmul")
print(f"cmul: 30 M pixels {t:.3f} ms = {a.numel() * 24 / t / 1e6:.0f} GB/s (algorithmic 24 B/pixel)", flush=True)

# despeck: 30-date stack not needed, two bands of the 1500 x 20000 image, 11x5 window
lines, cols = 1500, 20000
slc = synth.make_stack_torch(30, lines, cols, seed=2, device=dev, row_range=(0, 300))
count, wts = ctx.nmap_block_device(slc, 5, 2, "KS2", 0.05)
z1, z2 = slc[0].contiguous(), slc[7].contiguous()
for coh in (False, True):
    for _ in range(3):
        out = ctx.despeck_block_device(z1, wts, 5, 2, z2=z2, coherence=coh)
        torch.cuda.synchronize()
    t = ctx.last_kernel_ms("despeck")
    npx = z1.numel()
    print(f"despeck coherence={coh}: {npx / 1e6:.1f} M pixels {t:.3f} ms = {npx / t / 1e3:.0f} M px/s, "
          f"{npx * (16 + 8 + 8) / t / 1e6:.0f} GB/s of the 32 algorithmic B/pixel (2 bands in, mask, 1 out)", flush=True)

# ampdispersion: one streaming pass, 8 N + 8 bytes per pixel
for _ in range(3):
    da, mean = ctx.ampdispersion_block_device(slc)
    torch.cuda.synchronize()
t = ctx.last_kernel_ms("ampdispersion")
npx = slc.shape[1] * slc.shape[2]
print(f"ampdispersion: 30 dates {npx / 1e6:.1f} M pixels {t:.3f} ms = {npx / t / 1e3:.0f} M px/s, "
      f"{npx * (8 * 30 + 8) / t / 1e6:.0f} GB/s (algorithmic 248 B/pixel)", flush=True)
