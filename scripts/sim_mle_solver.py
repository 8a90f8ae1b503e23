"""Offline replay of the FP64 smallest-eigenpair solver of the MLE path (smallest_eigen_mle, MODE 0 in
evd_kernels.cu) on M = inv(|C|) o C matrices of a synthetic stack: how many Cholesky factorisations and
inverse-iteration steps it spends per pixel, against a single-factorisation variant.  Cost model per
pixel in complex MACs: factorisation N^3/3, one step (two triangular solves + one product) 3 N^2.

    python scripts/sim_mle_solver.py [bands]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fringe_b200 import synth  # noqa: E402
from oracle import load  # noqa: E402


def mle_matrices(bands=20, n=150, seed=3):
    slc = synth.make_stack(bands, 48, 192, seed=seed)
    count, wts = load().nmap_block(slc, 5, 2, 0, 0.05)[:2]
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        y, x = rng.integers(2, 46), rng.integers(5, 186)
        if count[y, x] < 2:
            continue
        idx = [(y + f // 11 - 2, x + f % 11 - 5) for f in range(55) if (wts[y, x, f >> 5] >> (f & 31)) & 1]
        Z = np.stack([slc[:, a, b] for a, b in idx], 1).astype(np.complex128)
        C = Z @ Z.conj().T
        d = np.sqrt(np.real(np.diag(C)))
        C = C / np.outer(d, d)
        A = np.abs(C)
        np.fill_diagonal(A, 1.0)
        try:
            np.linalg.cholesky(A - 1e-6 * np.eye(bands))
            np.linalg.cholesky(C - 1e-6 * np.eye(bands))
        except np.linalg.LinAlgError:
            continue                                    # gated out before the solver, as in the kernel
        out.append(np.linalg.inv(A) * C)
    return out


def chol_ok(M):
    try:
        return np.linalg.cholesky(M)
    except np.linalg.LinAlgError:
        return None


def shipped(M):
    """The kernel's loop: certified shift just below zero, tighten when the residual allows."""
    n = M.shape[0]
    dmax = np.abs(np.diag(M)).max()
    nchol = steps = 0
    back, L = 1e-12 * dmax, None
    for _ in range(6):
        sigma = -back
        L = chol_ok(M - sigma * np.eye(n)); nchol += 1
        if L is not None:
            break
        back *= 1e3
    if L is None:
        return None
    sigma_ok = sigma
    x = np.ones(n, complex)
    since = 0
    for it in range(200):
        x = np.linalg.solve(L.conj().T, np.linalg.solve(L, x)); steps += 1
        x /= np.linalg.norm(x)
        y = M @ x
        rho = np.real(np.vdot(x, y))
        res = np.linalg.norm(y - rho * x)
        if res <= 1e-11 * dmax:
            break
        since += 1
        prop = rho - 2 * res
        if since >= 2 and res > 1e-8 * dmax and prop > sigma_ok + 0.25 * (rho - sigma_ok):
            L2 = chol_ok(M - prop * np.eye(n)); nchol += 1
            if L2 is not None:
                sigma_ok, L = prop, L2
            else:
                mid = 0.5 * (sigma_ok + prop)
                L2 = chol_ok(M - mid * np.eye(n)); nchol += 1
                if L2 is not None:
                    sigma_ok, L = mid, L2
                else:
                    L = chol_ok(M - sigma_ok * np.eye(n)); nchol += 1
            since = 0
    return nchol, steps


def single_factorisation(M, tol=1e-11):
    n = M.shape[0]
    dmax = np.abs(np.diag(M)).max()
    L = chol_ok(M + 1e-12 * dmax * np.eye(n))
    if L is None:
        return None
    x = np.ones(n, complex)
    for it in range(2000):
        x = np.linalg.solve(L.conj().T, np.linalg.solve(L, x))
        x /= np.linalg.norm(x)
        y = M @ x
        rho = np.real(np.vdot(x, y))
        if np.linalg.norm(y - rho * x) <= tol * dmax:
            return 1, it + 1
    return 1, 2000


if __name__ == "__main__":
    bands = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    mats = mle_matrices(bands)
    cost = lambda c, s: c * bands ** 3 / 3 + s * 3 * bands ** 2
    for name, fn in (("shipped (certified shifts)", shipped), ("one factorisation, plain inverse iteration", single_factorisation)):
        res = [r for r in (fn(M) for M in mats) if r is not None]
        c, s = np.mean([r[0] for r in res]), np.mean([r[1] for r in res])
        print(f"{name:45s} factorisations {c:5.2f}  steps {s:7.2f}  (p90 {np.quantile([r[1] for r in res], 0.9):.0f})  "
              f"cost {cost(c, s) / 1e3:7.1f} k cMAC")
    gaps = []
    for M in mats:
        w = np.linalg.eigvalsh(M)
        gaps.append(w[0] / w[1])
    print("lambda_min / lambda_2 of M, quantiles 10/50/90 %:", np.quantile(gaps, [0.1, 0.5, 0.9]).round(3))
