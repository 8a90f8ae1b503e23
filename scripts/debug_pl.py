import sys, numpy as np
sys.path.insert(0, '.')
import oracle
from fringe_b200 import synth
from fringe_b200.engine import Context
O = oracle.load(); ctx = Context(0)
slc = synth.make_stack(20, 64, 96, seed=5, region=32)
wts = O.nmap_block(slc, 5, 2)[1]
ref = O.evd_block(slc, wts, 5, 2, method=1, variant=1, min_neighbors=5, want_npix=True)
gpu = ctx.evd_block(slc, wts, 5, 2, method="MLE", variant=1, min_neighbors=5)
print(ctx.evd_stats())
evd_ref = O.evd_block(slc, wts, 5, 2, method=0)
ok = ref[1] > 0.3
dph = np.abs(np.angle(ref[0] * np.conj(gpu[0]))).max(axis=0) * ok
idx = np.argsort(dph.ravel())[::-1][:8]
for i in idx:
    r, c = divmod(i, 96)
    dph_evd = np.abs(np.angle(evd_ref[0][:, r, c] * np.conj(ref[0][:, r, c]))).max()
    print(r, c, 'npix', ref[3][r, c], 't_ref', ref[1][r, c], 't_gpu', gpu[1][r, c], 'dphase', dph[r, c], 'ref_vs_evd', dph_evd)
    if i == idx[0]:
        # rebuild this pixel's matrices in numpy to look at conditioning
        N = 20
        C = np.zeros((N, N), complex); P = np.zeros(N)
        for dy in range(-2, 3):
            for dx in range(-5, 6):
                f = (dy + 2) * 11 + dx + 5
                if (wts[r, c, f // 32] >> np.uint32(f % 32)) & 1 and 0 <= r + dy < 64 and 0 <= c + dx < 96:
                    z = slc[:, r + dy, c + dx].astype(complex)
                    C += np.outer(z, z.conj()); P += np.abs(z) ** 2
        C = C / np.sqrt(np.outer(P, P)); np.fill_diagonal(C, 1)
        A = np.abs(C); M = np.linalg.inv(A) * C
        w = np.linalg.eigvalsh(M); print('eig(M)[:4]', w[:4], 'cond(A)', np.linalg.cond(A), 'eig(A) min', np.linalg.eigvalsh(A)[0])
