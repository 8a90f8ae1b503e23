#!/usr/bin/env python3
"""Headline benchmark: phase-linked pixels/s (N=30 dates, 11x5 window), BASELINE.json configs[1].

One "step" = one pass of the hot path (nmap KS2 -> evd EVD) over the 30-date 1500x20000
synthetic stack.  With --gpus N > 1 (launched under torchrun, one rank per GPU) the image rows
are partitioned across ranks with Ny-line halos taken from the input -- tiles are independent, so
there is no data-path collective; total work is fixed ("strong" scaling).

Printed JSON (rank 0, one line):
  value      whole-job pixels/s with the stack already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the host C ABI (fringe_nmap_evd_block; also the two separate calls) from
             pinned host buffers, H2D and D2H inside the timed region
  roofline   the dominant kernel (k_evd: covariance + eigen + post) against the FP32-FMA peak
             measured in this run (MEASURED_PEAKS.json has no FP32 figure); algorithmic flops per
             SURVEY.md section 8(d)
  cpu_baseline  the CPU oracle (reference headers build when present) on a bounded strip

`--impl reference` times the reference's CPU path (oracle) alone on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BANDS, LINES, COLS = 30, 1500, 20000
NX, NY = 5, 2
WORKLOAD = "configs[1]: 30-date 1500x20000 synthetic stack, nmap KS2 11x5 (Nx=5,Ny=2,p=0.05) -> evd EVD"


def flops_per_pixel(n: int, shp_sum: float, solved: int, mle: bool = False) -> float:
    """F_cov = S(4N(N-1)+4N); F_eig = 16/3 N^3 + 16 N^2; F_post = 10N(N-1)+30N."""
    f = shp_sum * (8 * n * (n - 1) / 2 + 4 * n)
    per = (16.0 / 3.0) * n ** 3 + 16 * n * n + 20 * n * (n - 1) / 2 + 30 * n
    if mle:
        per += n ** 3 + 2 * n * n
    return f + solved * per


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(steps: int, warmup: int, sample_lines: int = 64, sample_cols: int = 2048):
    """Time the CPU oracle (the reference's own headers + restated loops, OpenMP over pixels) on a
    strip of the workload.  Returns (pixels/s, description dict)."""
    import oracle
    from fringe_b200 import synth
    o = oracle.load()
    cores = os.cpu_count() or 1
    o.set_threads(cores)
    slc = synth.make_stack(BANDS, sample_lines, sample_cols, seed=2)
    npx = sample_lines * sample_cols
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, wts = o.nmap_block(slc, NX, NY, method=0, thresh=0.05)
        o.evd_block(slc, wts, NX, NY, method=0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    rate = npx * len(times) / sum(times)
    desc = {"value": rate, "unit": "pixels/s", "cores": cores, "kind": o.kind,
            "sample": f"{sample_lines}x{sample_cols} strip of the same 30-date stack, nmap KS2 11x5 + evd EVD, "
                      f"{len(times)} pass(es), OpenMP over pixels, OpenBLAS single-threaded per call"}
    return rate, desc, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, desc, sec = cpu_reference_rate(max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "phase-linked pixels/sec (N=30 dates, 11x5 window)",
            "value": rate, "unit": "pixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": desc["sample"]},
            "cpu_baseline": desc,
            "e2e": {"value": rate, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lines", type=int, default=LINES)
    ap.add_argument("--cols", type=int, default=COLS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fringe_b200 import synth
    from fringe_b200.engine import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION prints a banner there
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    lines, cols = args.lines, args.cols
    # row partition with halos: rank g owns rows [r0, r1), reads [b0, b1)
    from fringe_b200.partition import row_tile
    r0, r1, b0, b1, first_line, n_lines = row_tile(lines, rank, world, NY)
    blines = b1 - b0
    my_pixels = n_lines * cols

    ctx = Context(local)
    slc = synth.make_stack_torch(BANDS, lines, cols, seed=2, device=dev, row_range=(b0, b1))
    nu = 2
    count = torch.empty((blines, cols), dtype=torch.int32, device=dev)
    wts = torch.empty((blines, cols, nu), dtype=torch.int32, device=dev)
    out = torch.zeros((BANDS, blines, cols), dtype=torch.complex64, device=dev)
    tcorr = torch.zeros((blines, cols), dtype=torch.float32, device=dev)
    comp = torch.zeros((blines, cols), dtype=torch.complex64, device=dev)

    def step_device():
        ctx.nmap_block_device(slc, NX, NY, "KS2", 0.05, count=count, wts=wts)
        ctx.evd_block_device(slc, wts, NX, NY, "EVD", first_line=first_line, n_lines=n_lines,
                             out=out, tcorr=tcorr, comp=comp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    evd_ms, nmap_ms = [], []
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
        evd_ms.append(ctx.last_kernel_ms("evd"))
        nmap_ms.append(ctx.last_kernel_ms("nmap"))
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    total_pixels = lines * cols
    value = total_pixels * args.steps / (max_ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launch) ------------------------------
    cnt_int = count[first_line:first_line + n_lines]
    solved = int((cnt_int >= 2).sum().item())
    shp_sum = float(cnt_int[cnt_int >= 2].sum().item())
    stats = ctx.evd_stats()
    fl = flops_per_pixel(BANDS, shp_sum, solved)
    k_ms = float(np.mean(evd_ms))
    achieved = fl / (k_ms * 1e-3) * 1e-12
    peak = ctx.fp32_peak_tflops()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = 6650.0
    hbm_src = "fallback"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "measured"
        except Exception:
            pass
    evd_bytes = my_pixels * (16 * BANDS + 4 * nu + 12)
    # ncu --set full on a 100-line launch of the same kernel (profiles/r1_evd_mma_ncu_summary.txt):
    # dram__bytes_read + dram__bytes_write = 2.02 GB for 2.0 M pixels -> 1012 B/pixel, scaled to this launch
    traffic = 1012.0 * my_pixels
    roofline = {"kernel": "k_evd_mma (masked Gram product on 3xTF32 mma.sync + FP32 dominant eigenvector + "
                          "phase ref + tcorr + compressed SLC)",
                "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None,
                "peak_source": "FP32 FMA microbenchmark run inside this bench (fringe_fp32_peak); "
                               "MEASURED_PEAKS.json has no FP32 figure.  Algorithmic flops (SURVEY 8d) over "
                               "the FP32 peak, although the Gram product itself runs on the tensor pipe",
                "kernel_ms": k_ms, "algorithmic_flops_per_launch": fl,
                "mean_shp": shp_sum / max(solved, 1), "solved_pixels": solved,
                "hbm_view": {"achieved_gbs": evd_bytes / (k_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                             "peak_source": hbm_src, "algorithmic_bytes_per_pixel": 16 * BANDS + 4 * nu + 12},
                "nmap_kernel_ms": float(np.mean(nmap_ms)),
                "power_iterations_per_pixel": stats["power_iterations"] / max(stats["pixels"], 1),
                "traffic": traffic,
                "traffic_note": "DRAM bytes per launch, 1012 B/pixel from the ncu capture of a 100-line launch "
                                "(algorithmic 500 B/pixel; the excess is the 512 B/pixel hi/lo sample layout read "
                                "through L2)"}

    # ---- end to end through the host C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_slc = torch.empty((BANDS, blines, cols), dtype=torch.complex64, pin_memory=True)
        h_slc.copy_(slc)
        h_count = torch.empty((blines, cols), dtype=torch.int32, pin_memory=True)
        h_wts = torch.empty((blines, cols, nu), dtype=torch.int32, pin_memory=True)
        h_out = torch.empty((BANDS, blines, cols), dtype=torch.complex64, pin_memory=True)
        h_tcorr = torch.empty((blines, cols), dtype=torch.float32, pin_memory=True)
        h_comp = torch.empty((blines, cols), dtype=torch.complex64, pin_memory=True)
        from fringe_b200._lib import lib

        def step_two_calls():
            ctx._check(lib.fringe_nmap_block(ctx._h, h_slc.data_ptr(), None, None, cols, blines, BANDS, NX, NY,
                                             0, 0.05, h_count.data_ptr(), h_wts.data_ptr()))
            ctx._check(lib.fringe_evd_block(ctx._h, h_slc.data_ptr(), h_wts.data_ptr(), cols, blines, BANDS,
                                            NX, NY, first_line, n_lines, 0, -1, 1, 0, 2, h_out.data_ptr(),
                                            h_tcorr.data_ptr(), h_comp.data_ptr()))

        def step_fused():
            ctx._check(lib.fringe_nmap_evd_block(ctx._h, h_slc.data_ptr(), None, None, cols, blines, BANDS, NX, NY,
                                                 0, 0.05, first_line, n_lines, 0, -1, 1, 0, 2, h_count.data_ptr(),
                                                 h_wts.data_ptr(), h_out.data_ptr(), h_tcorr.data_ptr(),
                                                 h_comp.data_ptr()))

        def time_host(step):
            for _ in range(min(args.warmup, 3)):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        npb = blines * cols
        d2h = npb * 4 + npb * nu * 4 + my_pixels * (BANDS * 8 + 4 + 8)
        t_fused = time_host(step_fused)
        t_two = time_host(step_two_calls)
        # what the host link gives this rank while every rank is copying both ways at once
        # (explains the gap between e2e and the device-resident value at N > 1)
        nb = min(1 << 29, h_slc.numel() * 8)
        src = h_slc.view(torch.uint8).reshape(-1)[:nb]
        dst = h_out.view(torch.uint8).reshape(-1)[:nb]
        d_a = torch.empty(nb, dtype=torch.uint8, device=dev)
        d_b = torch.empty(nb, dtype=torch.uint8, device=dev)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def step_link():
            with torch.cuda.stream(s_up):
                d_a.copy_(src, non_blocking=True)
            with torch.cuda.stream(s_dn):
                dst.copy_(d_b, non_blocking=True)
        t_link = time_host(step_link)
        link_gbs = nb * args.steps / t_link * 1e-9
        del d_a, d_b
        # headline: both stages on one upload; the mask and count still come back to the host
        e2e = {"value": total_pixels * args.steps / t_fused, "unit": "pixels/s",
               "h2d_bytes_per_step": npb * BANDS * 8, "d2h_bytes_per_step": d2h,
               "ms_per_step": t_fused * 1e3 / args.steps,
               "api": "fringe_nmap_evd_block (host pointers, pinned; count, mask, phase, tcorr, "
                      "compressed SLC all copied back), per rank",
               "host_link": {"gbs_each_way": link_gbs, "note": "pinned copies of %d MB up and down at the same "
                             "time on every rank, max over ranks" % (nb >> 20),
                             "floor_ms_per_step": max(npb * BANDS * 8, d2h) / (link_gbs * 1e9) * 1e3},
               "two_calls": {"value": total_pixels * args.steps / t_two, "unit": "pixels/s",
                             "h2d_bytes_per_step": 2 * npb * BANDS * 8 + npb * nu * 4,
                             "d2h_bytes_per_step": d2h, "ms_per_step": t_two * 1e3 / args.steps,
                             "api": "fringe_nmap_block + fringe_evd_block, the stack uploaded twice as "
                                    "nmap.py -> evd.py do"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        _, cpu, _ = cpu_reference_rate(1, 0)

    if rank == 0:
        line = {"metric": "phase-linked pixels/sec (N=30 dates, 11x5 window)", "value": value, "unit": "pixels/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "lines": lines, "cols": cols, "bands": BANDS,
                           "partition": f"{world} row tile(s) with {NY}-line halos, no collective",
                           "l2": "inputs (7.2 GB stack) far larger than the 126 MB L2; no explicit flush",
                           "timing": "CUDA events on the launching (torch current) stream, max over ranks"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
