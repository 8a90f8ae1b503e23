#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE-HEADER build of the oracle
(oracle/_ref/libfringe_ref.so = the reference's KS2sample.hpp / AD2unique.hpp / ulongmask.hpp /
EigenLapack.hpp compiled in place from /root/reference + the restated block loops).
Run in the authoring container only:  python tests/golden/make_golden.py
The fixtures are small on purpose (committed to git)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from fringe_b200 import synth  # noqa: E402


def main():
    ref = oracle.load("reference")
    assert ref.kind == "reference"
    rng = np.random.default_rng(2024)

    # ---- single-pair known answers (KS2 / AD2), ties included --------------------------------
    pairs_a, pairs_b, ks_p, ad_p = [], [], [], []
    for n in (5, 10, 20, 30):
        for t in range(12):
            a = np.sort(rng.rayleigh(rng.choice([1.0, 1.5, 2.5]), n)).astype(np.float32)
            b = np.sort(rng.rayleigh(rng.choice([1.0, 1.5, 2.5]), n)).astype(np.float32)
            if t % 3 == 0:
                a = np.sort(np.round(a * 4) / 4 + 0.25).astype(np.float32)
                b = np.sort(np.round(b * 4) / 4 + 0.25).astype(np.float32)
            pa = np.full(30, np.nan, np.float32); pb = np.full(30, np.nan, np.float32)
            pa[:n] = a; pb[:n] = b
            pairs_a.append(pa); pairs_b.append(pb)
            ks_p.append(ref.ks2_prob(a, b)); ad_p.append(ref.ad2_prob(a, b))
    np.savez_compressed(os.path.join(HERE, "pairs.npz"), a=np.array(pairs_a), b=np.array(pairs_b),
                        ks_p=np.array(ks_p), ad_p=np.array(ad_p),
                        ad_sigma=np.array([ref.ad2_sigma(n) for n in (5, 10, 20, 30, 100)]))

    # ---- block fixtures -----------------------------------------------------------------------
    slc = synth.make_stack(12, 24, 40, seed=77, region=8)
    count_ks, wts_ks = ref.nmap_block(slc, 5, 2, method=oracle.KS2, thresh=0.05)
    count_ad, wts_ad = ref.nmap_block(slc, 3, 3, method=oracle.AD2, thresh=0.05)
    out = {"slc": slc, "count_ks": count_ks, "wts_ks": wts_ks, "count_ad": count_ad, "wts_ad": wts_ad}
    for name, kw in (("evd", dict(method=oracle.EVD)), ("mle", dict(method=oracle.MLE)),
                     ("stbas", dict(method=oracle.STBAS, bandwidth=4)),
                     ("pl", dict(method=oracle.MLE, variant=oracle.VARIANT_PHASE_LINK, min_neighbors=5)),
                     ("seq", dict(method=oracle.MLE, mini_stack_count=3, first_line=2, n_lines=20))):
        o, t, c = ref.evd_block(slc, wts_ks, 5, 2, **kw)
        out[name + "_out"], out[name + "_tcorr"], out[name + "_comp"] = o, t, c
    np.savez_compressed(os.path.join(HERE, "block_12x24x40.npz"), **out)
    make_post(ref, slc, wts_ks)
    print("wrote", os.listdir(HERE))


def make_post(ref, slc, wts_ks):
    """SURVEY 8f rows on the same block: despeck (three modes), ampdispersion with calibration
    constants, datum-adjustment product.  Separate file so the older fixtures stay byte-identical."""
    alpha = np.linspace(1.0, 1.33, slc.shape[0])
    da, mean = ref.ampdispersion_block(slc, alpha)
    post = {"alpha": alpha, "ampdisp_da": da, "ampdisp_mean": mean,
            "despeck_amp": ref.despeck_block(slc[2], wts_ks, 5, 2),
            "despeck_ifg": ref.despeck_block(slc[2], wts_ks, 5, 2, z2=slc[9]),
            "despeck_coh": ref.despeck_block(slc[2], wts_ks, 5, 2, z2=slc[9], coherence=True),
            "cmul": ref.cmul(slc[4], slc[7])}
    np.savez_compressed(os.path.join(HERE, "post_12x24x40.npz"), **post)


if __name__ == "__main__":
    main()
