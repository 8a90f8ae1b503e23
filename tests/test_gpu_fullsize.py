"""GPU, BASELINE.json full size (configs[1]: 30 dates, 1500 x 20000, KS2 11x5 -> EVD): the oracle
cannot run 3*10^7 pixels in test time, so the whole image is checked through size-independent
properties, and random crops of it against the oracle."""
import numpy as np
import pytest

from conftest import wrapped_diff

pytestmark = pytest.mark.gpu

BANDS, LINES, COLS, NX, NY = 30, 1500, 20000, 5, 2


@pytest.fixture(scope="module")
def full(ctx):
    import torch
    from fringe_b200 import synth
    dev = torch.device("cuda", 0)
    slc = synth.make_stack_torch(BANDS, LINES, COLS, seed=2, device=dev)
    count, wts = ctx.nmap_block_device(slc, NX, NY, "KS2", 0.05)
    out, tcorr, comp = ctx.evd_block_device(slc, wts, NX, NY, "EVD")
    torch.cuda.synchronize()
    return slc, count, wts, out, tcorr, comp


def _bit(wts, dy, dx):
    import torch
    f = (dy + NY) * (2 * NX + 1) + dx + NX
    return (wts[..., f // 32] >> (f % 32)) & 1


def test_mask_symmetry_count_and_validity(full):
    import torch
    slc, count, wts, *_ = full
    # bit (dy,dx) of p equals bit (-dy,-dx) of p+(dy,dx): the pair decision is symmetric (nmap.cpp:464-468)
    for dy in range(0, NY + 1):
        for dx in range(-NX, NX + 1):
            if dy == 0 and dx <= 0:
                continue
            a = _bit(wts, dy, dx)[0:LINES - dy, max(0, -dx):COLS - max(0, dx)]
            b = _bit(wts, -dy, -dx)[dy:LINES, max(0, dx):COLS + min(0, dx)]
            assert torch.equal(a, b), (dy, dx)
    # count = popcount of the mask words
    pop = torch.zeros_like(count)
    for w in range(wts.shape[-1]):
        v = wts[..., w].to(torch.int64) & 0xFFFFFFFF
        for s in range(32):
            pop += ((v >> s) & 1).to(torch.int32)
    assert torch.equal(pop, count)
    # centre bit <=> pixel valid (no zero / NaN amplitude in any date, nmap.cpp:376)
    valid = (slc.abs() != 0).all(dim=0)
    assert torch.equal(_bit(wts, 0, 0).bool(), valid)
    assert int(count.max()) <= 55 and 20 < float(count.float().mean()) < 50
    # no bit outside the image
    assert int(_bit(wts, -1, 0)[0].sum()) == 0 and int(_bit(wts, 0, -1)[:, 0].sum()) == 0


def test_phasors_unit_reference_band_and_sentinels(full):
    import torch
    slc, count, wts, out, tcorr, comp = full
    solved = tcorr > 0
    assert float(solved.float().mean()) > 0.95
    mag = out.abs()
    assert float((mag[:, solved] - 1).abs().max()) < 1e-5
    assert bool((out[0][solved] == 1).all())                       # evd.cpp:748: reference band exactly 1+0j
    assert bool((out[:, ~solved] == 0).all()) and bool((comp[~solved] == 0).all())
    assert float(tcorr.max()) <= 1.0 + 1e-6 and float(tcorr.min()) >= 0.0   # EVD never yields a sentinel here
    # pixels with fewer than 2 SHPs are skipped (evd.cpp:566)
    assert bool((tcorr[count < 2] == 0).all())


def test_scale_and_band_rotation_equivariance(ctx, full):
    """Coherence is normalised, so scaling the stack changes nothing; rotating one date by a constant
    phase rotates that date's phasor by the same angle and leaves temporal coherence alone."""
    import torch
    slc, count, wts, out, tcorr, comp = full
    rows = slice(700, 764)                                         # a 64-line strip keeps this cheap
    sub = slc[:, 698:766].contiguous()
    wsub = wts[698:766].contiguous()
    o0, t0, _ = ctx.evd_block_device(sub, wsub, NX, NY, "EVD", first_line=2, n_lines=64)
    torch.cuda.synchronize()
    # block independence: the strip computed alone equals the rows of the full-image run
    assert torch.equal(t0[2:66], tcorr[rows]) and torch.equal(o0[:, 2:66], out[:, rows])
    rot = sub.clone()
    phi = 0.7
    rot[7] *= complex(np.cos(phi), np.sin(phi))
    o1, t1, _ = ctx.evd_block_device((rot * 3.7).contiguous(), wsub, NX, NY, "EVD", first_line=2, n_lines=64)
    torch.cuda.synchronize()
    ok = t0[2:66] > 0.3
    assert float((t1[2:66] - t0[2:66]).abs().max()) < 5e-5          # half the 1e-4 parity gate
    d = torch.angle(o1[:, 2:66] * torch.conj(o0[:, 2:66]))
    expect = torch.zeros(BANDS, device=d.device); expect[7] = phi
    err = torch.angle(torch.exp(1j * (d - expect[:, None, None])))
    # 38 M (band, pixel) entries: a handful belong to eigenvector components of magnitude ~1e-4 (seen in
    # float64: |v_19| = 1.0e-4 at the one offender), whose phase no single-precision solve can pin to
    # 1e-3 rad; everything else must
    e = err[:, ok].abs()
    assert int((e >= 1e-3).sum()) <= 3 and float(e.max()) < 2e-2


def test_random_crops_against_oracle(full, oracle_lib):
    slc, count, wts, out, tcorr, comp = full
    rng = np.random.default_rng(4)
    for _ in range(4):
        r0 = int(rng.integers(0, LINES - 40)); c0 = int(rng.integers(0, COLS - 72))
        crop = slc[:, r0:r0 + 40, c0:c0 + 72].cpu().numpy()
        c_ref, w_ref = oracle_lib.nmap_block(crop, NX, NY)
        inner = (slice(NY, 40 - NY), slice(NX, 72 - NX))           # full windows inside the crop
        assert np.array_equal(count[r0:r0 + 40, c0:c0 + 72].cpu().numpy()[inner], c_ref[inner])
        w_gpu = wts[r0:r0 + 40, c0:c0 + 72].cpu().numpy().view(np.uint32)
        assert np.array_equal(w_gpu[inner], w_ref[inner])
        o_ref, t_ref, _ = oracle_lib.evd_block(crop, w_gpu, NX, NY, method=0, first_line=NY, n_lines=40 - 2 * NY)
        t_gpu = tcorr[r0:r0 + 40, c0:c0 + 72].cpu().numpy()
        o_gpu = out[:, r0:r0 + 40, c0:c0 + 72].cpu().numpy()
        assert np.abs(t_gpu - t_ref)[inner].max() <= 1e-4
        good = np.zeros_like(t_ref, bool); good[inner] = t_ref[inner] > 0.3
        assert wrapped_diff(o_gpu[:, good], o_ref[:, good]).max() <= 1e-3
