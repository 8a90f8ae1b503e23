"""CPU, world_size 2 over gloo: the multi-GPU row partition (fringe_b200/partition.py, used by
bench.py and the host driver) reproduces the whole-image result with no data-path collective.
No GPU here, so each rank's tile is computed by the CPU oracle standing in for the kernels; what
is under test is the host-side tiling / halo / reassembly logic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lines, cols, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from fringe_b200 import synth
    from fringe_b200.partition import row_tile
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.load()
    o.set_threads(2)
    Nx, Ny = 3, 2
    slc = synth.make_stack(8, lines, cols, seed=21, region=8)       # every rank sees the same "disk"
    t = row_tile(lines, rank, world, Ny)
    block = slc[:, t.b0:t.b1]
    count, wts = o.nmap_block(block, Nx, Ny)
    out, tcorr, _ = o.evd_block(block, wts, Nx, Ny, method=oracle.EVD, first_line=t.first_line, n_lines=t.n_lines)
    mine = torch.from_numpy(tcorr[t.first_line:t.first_line + t.n_lines].copy())
    cnt = torch.from_numpy(count[t.first_line:t.first_line + t.n_lines].copy())
    # verification-only gather (not part of the data path): variable tile heights -> pad
    hmax = (lines + world - 1) // world + 1
    pad_t = torch.zeros((hmax, cols)); pad_t[:mine.shape[0]] = mine
    pad_c = torch.zeros((hmax, cols), dtype=torch.int32); pad_c[:cnt.shape[0]] = cnt
    gt = [torch.zeros_like(pad_t) for _ in range(world)]
    gc = [torch.zeros_like(pad_c) for _ in range(world)]
    dist.all_gather(gt, pad_t)
    dist.all_gather(gc, pad_c)
    if rank == 0:
        full_c, full_w = o.nmap_block(slc, Nx, Ny)
        _, full_t, _ = o.evd_block(slc, full_w, Nx, Ny, method=oracle.EVD)
        tiles = [row_tile(lines, r, world, Ny) for r in range(world)]
        asm_t = np.concatenate([gt[r][:tiles[r].n_lines].numpy() for r in range(world)])
        asm_c = np.concatenate([gc[r][:tiles[r].n_lines].numpy() for r in range(world)])
        q.put((bool(np.array_equal(asm_c, full_c)), bool(np.array_equal(asm_t, full_t))))
    dist.destroy_process_group()


@pytest.mark.parametrize("lines", [23, 40])
def test_two_rank_row_partition_matches_whole_image(lines):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + lines
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lines, 20, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_c, ok_t = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_c and ok_t


def test_row_tiles_cover_image_exactly():
    from fringe_b200.partition import row_tile
    for lines in (1, 7, 1500):
        for world in (1, 2, 4, 8):
            tiles = [row_tile(lines, r, world, 2) for r in range(world)]
            assert tiles[0].r0 == 0 and tiles[-1].r1 == lines
            for a, b in zip(tiles, tiles[1:]):
                assert a.r1 == b.r0
            for t in tiles:
                assert t.b0 == max(0, t.r0 - 2) and t.b1 == min(lines, t.r1 + 2)
                assert t.first_line == t.r0 - t.b0 and t.n_lines == t.r1 - t.r0
