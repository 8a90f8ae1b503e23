"""Offline replay of k_evd_mma's power iteration (momentum switch-on, residual-test schedule) on
coherence matrices of the bench stack: how the constants in evd_mma.cu were chosen.

    python scripts/sim_power_iteration.py            # needs oracle/ (CPU nmap for the SHP masks)

Prints mean iterations / residual tests per pixel for a few settings; "exact r" rows use the true
lambda2/lambda1 from numpy and bound what any estimator could reach."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fringe_b200 import synth  # noqa: E402
from oracle import load  # noqa: E402


def matrices(n=300, seed=1):
    slc = synth.make_stack(30, 48, 256, seed=2)
    count, wts = load().nmap_block(slc, 5, 2, 0, 0.05)[:2]
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        y, x = rng.integers(2, 46), rng.integers(5, 250)
        if count[y, x] < 2:
            continue
        idx = [(y + f // 11 - 2, x + f % 11 - 5) for f in range(55) if (wts[y, x, f >> 5] >> (f & 31)) & 1]
        Z = np.stack([slc[:, a, b] for a, b in idx], 1).astype(np.complex64)
        C = Z @ Z.conj().T
        d = np.sqrt(np.real(np.diag(C)))
        C = (C / np.outer(d, d)).astype(np.complex64)
        np.fill_diagonal(C, 1)
        out.append(C)
    return out


def run(C, first_chk=3, first_gap=2, safety=0.575, cap=0.2, tol2=4e-12, mfac=1.25, exact_r=None):
    """The kernel's loop: Rayleigh quotient at next_chk-1, residual test at next_chk."""
    x = np.conj(C[0, :]).copy()
    x = x / np.abs(x)
    x = (x / np.linalg.norm(x)).astype(np.complex64)
    xp = np.zeros_like(x)
    lam = inv_lam = 1.0
    beta = 0.0 if exact_r is None else (0.5 * exact_r) ** 2
    rho_prev, next_chk, gap, checks = -1.0, first_chk, first_gap, 0
    for it in range(1000):
        y = (C @ x).astype(np.complex64)
        if it < next_chk - 1:
            x, xp = y * inv_lam - beta * xp, x
        elif it == next_chk - 1:
            xx = np.real(np.vdot(x, x))
            lam = np.real(np.vdot(x, y)) / xx
            inv_lam, sc = 1 / lam, 1 / np.sqrt(xx)
            x, xp = (y * inv_lam - beta * xp) * sc, x * sc
        else:
            checks += 1
            r = y - lam * x
            rho2 = np.real(np.vdot(r, r)) / (lam * lam * np.real(np.vdot(x, x)))
            if rho2 <= tol2:
                return it + 1, checks
            rate_l2 = -0.15
            if rho_prev > 0 and rho2 < rho_prev:
                rate_l2 = np.log2(rho2 / rho_prev) * (0.5 / gap)
                if beta == 0.0:
                    rr = 2 ** rate_l2
                    beta = min((safety * rr) ** 2, cap)
                    rate_l2 = np.log2(rr / (1 + np.sqrt(max(1 - rr * rr, 0))))
            elif rho_prev > 0:
                beta *= 0.5
            m = mfac * 0.5 * np.log2(tol2 / rho2) / min(rate_l2, -0.01) + 0.5
            gap = min(max(int(m), 2), 12) if rho_prev > 0 else first_gap
            next_chk = it + gap
            rho_prev = rho2
            x, xp = y * inv_lam - beta * xp, x
    return 1000, checks


if __name__ == "__main__":
    mats = matrices()
    r = [np.linalg.eigvalsh(C.astype(np.complex128)) for C in mats]
    r = np.array([w[-2] / w[-1] for w in r])
    print("lambda2/lambda1 quantiles 5/50/95 %:", np.quantile(r, [0.05, 0.5, 0.95]).round(3))
    cases = [("round-1 first version: first test 3, gap 4, 0.475 r", dict(first_chk=3, first_gap=4, safety=0.475, cap=1.0, mfac=1.1)),
             ("shipped: first test 3, gap 2, 0.575 r, cap 0.2", dict()),
             ("0.5 r", dict(safety=0.5)), ("0.65 r", dict(safety=0.65))]
    for name, kw in cases:
        res = [run(C, **kw) for C in mats]
        print(f"{name:55s} iterations {np.mean([a for a, _ in res]):6.2f}  max {max(a for a, _ in res):4d}  tests {np.mean([b for _, b in res]):.2f}")
    res = [run(C, exact_r=ri, cap=1.0) for C, ri in zip(mats, r)]
    print(f"{'exact r from iteration 0, same test schedule':55s} iterations {np.mean([a for a, _ in res]):6.2f}")
