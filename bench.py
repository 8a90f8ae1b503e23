#!/usr/bin/env python3
"""Benchmark of the phase-linking hot path: phase-linked pixels/s.

Default (the headline, BASELINE.json configs[1]): one "step" = one pass of nmap KS2 -> evd EVD over the 30-date
1500 x 20000 synthetic stack.  `--config c1|c3|c4|c5` runs the other BASELINE configs through the same harness and
prints the same JSON shape (they are parity-test cases first; their default sizes are strips chosen to finish in
minutes and are named in `config.workload`).  With --gpus N > 1 (launched under torchrun, one rank per GPU) the image
rows are partitioned across ranks with halo lines taken from the input -- tiles are independent, so there is no
data-path collective; total work is fixed ("strong" scaling).

Printed JSON (rank 0, one line):
  value         whole-job pixels/s with the stack already resident in HBM (CUDA events, max over ranks)
  e2e           the same through the host C ABI (fringe_nmap_evd_block / fringe_sequential_block) from pinned host
                buffers, H2D and D2H inside the timed region
  roofline      the dominant kernel against the FP32 (EVD) or FP64 (MLE / phase_link) FMA peak measured in this run
                (MEASURED_PEAKS.json has no such figure); algorithmic flops per SURVEY.md section 8(d)
  cpu_baseline  the CPU oracle (reference headers build when present) on a bounded sample

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

CONFIGS = {
    # name: bands, lines, cols, Nx, Ny, nmap method, stage, evd method, variant, min_neighbors, ministack size, CPU sample (lines, cols)
    "c1": dict(bands=20, lines=512, cols=512, Nx=5, Ny=2, nmap="KS2", method="MLE", variant=0, minn=2, s=0, sample=(128, 512),
               metric="phase-linked pixels/sec (N=20 dates, 11x5 window, MLE)",
               workload="configs[0]: 20-date 512x512 synthetic stack, nmap KS2 11x5 (Nx=5,Ny=2,p=0.05) -> evd MLE (the binding's default estimator)"),
    "c2": dict(bands=30, lines=1500, cols=20000, Nx=5, Ny=2, nmap="KS2", method="EVD", variant=0, minn=2, s=0, sample=(64, 2048),
               metric="phase-linked pixels/sec (N=30 dates, 11x5 window)",
               workload="configs[1]: 30-date 1500x20000 synthetic stack, nmap KS2 11x5 (Nx=5,Ny=2,p=0.05) -> evd EVD"),
    "c3": dict(bands=100, lines=1500, cols=2000, Nx=5, Ny=2, nmap="KS2", method="MLE", variant=1, minn=5, s=0, sample=(32, 512),
               metric="phase-linked pixels/sec (N=100 dates, 11x5 window, phase_link)",
               workload="configs[2]: 100-date stack, 1500x2000 column strip of the 1500x20000 image, nmap KS2 11x5 -> phase_link "
                        "(MLE with EVD fall-back, min_neighbors 5)"),
    "c4": dict(bands=200, lines=1500, cols=2000, Nx=5, Ny=2, nmap="KS2", method="MLE", variant=0, minn=2, s=10, sample=(16, 512),
               metric="phase-linked pixels/sec (200 dates, sequential estimator, ministacks of 10)",
               workload="configs[3]: 200-date stack, 1500x2000 column strip, sequential estimator: 20 ministacks of 10 (MLE, "
                        "compressed-SLC hand-off on the device) + datum connection + adjustment; SHP mask of the full stack given"),
    "c5": dict(bands=30, lines=1500, cols=2048, Nx=10, Ny=10, nmap="AD2", method="EVD", variant=0, minn=2, s=0, sample=(48, 512),
               metric="phase-linked pixels/sec (N=30 dates, AD2 21x21 window)",
               workload="configs[4]: 30-date 1500x2048 strip, nmap AD2 21x21 (Nx=Ny=10,p=0.05) -> evd EVD"),
}
METHOD_CODE = {"EVD": 0, "MLE": 1, "STBAS": 2}
NMAP_CODE = {"KS2": 0, "AD2": 1}


def flops_per_launch(n: int, shp_sum: float, solved: int, mle: bool = False) -> float:
    """F_cov = S(4N(N-1)+4N); F_eig = 16/3 N^3 + 16 N^2; F_post = 10N(N-1)+30N; MLE: + N^3 + 2N^2."""
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


def bind_to_gpu_numa_node(index: int):
    """Bind this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned staging buffers of
    the end-to-end leg are first-touched on that node.  Returns the node number, or None when sysfs does not say or
    FRINGE_BENCH_NO_NUMA is set (A/B runs)."""
    if os.environ.get("FRINGE_BENCH_NO_NUMA"):
        return None
    try:
        import torch
        props = torch.cuda.get_device_properties(index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate_drivers(cfg, steps: int, warmup: int):
    """Time the reference's OWN block drivers -- src/nmap/nmap.cpp followed by src/evd/evd.cpp (or phase_link.cpp),
    compiled unmodified with OpenMP against the GDAL / Armadillo stand-ins of oracle/shims/ -- file in, file out on a strip
    of the workload in a RAM-backed directory.  None when those builds are absent or the config is the sequential chain."""
    import shutil
    import tempfile
    import oracle
    from fringe_b200 import stackio, synth
    second = "phase_link" if cfg["variant"] == 1 else "evd"
    cores = usable_cores()
    if cfg["s"] or cores < 2 or not (oracle.ref_driver_available("nmap", True) and oracle.ref_driver_available(second, True)):
        return None
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    nmap_lib, evd_lib = oracle.ref_driver("nmap", True), oracle.ref_driver(second, True)
    sl, sc = cfg["sample"]
    slc = synth.make_stack(cfg["bands"], sl, sc, seed=2)
    root = tempfile.mkdtemp(prefix="fringe_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    enc = lambda x: str(x).encode()
    times = []
    try:
        vrt = stackio.make_stack_on_disk(root, slc)
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)                                   # the drivers report on stdout; this process owes it one JSON line
        try:
            for it in range(warmup + steps):
                w, c, out = (os.path.join(root, f"{n}{it}") for n in ("nmap", "count", "evd"))
                t0 = time.perf_counter()
                rc = nmap_lib.ref_nmap(enc(vrt), enc(w), enc(c), None, cfg["Nx"], cfg["Ny"], enc(cfg["nmap"]), 0.05, 8192, 64)
                rc2 = getattr(evd_lib, "ref_" + second)(enc(vrt), enc(w), enc(out), enc(out), b"compslc.bin", cfg["Nx"], cfg["Ny"],
                                                        enc(cfg["method"]), -1, 1, cfg["minn"], 8192, 64)
                dt = time.perf_counter() - t0
                if rc or rc2:
                    return None
                if it >= warmup:
                    times.append(dt)
        finally:
            os.dup2(saved, 1)
            os.close(saved); os.close(devnull)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    rate = sl * sc * len(times) / sum(times)
    desc = {"value": rate, "unit": "pixels/s", "cores": cores, "kind": "reference",
            "sample": f"{sl}x{sc} strip of the same {cfg['bands']}-date workload, {len(times)} pass(es) of the reference's own "
                      f"src/nmap/nmap.cpp + src/{second}/{second}.cpp (compiled unmodified, OpenMP on the {cores} usable cores, "
                      f"OpenBLAS single-threaded per call; GDAL / Armadillo replaced by I/O and storage stand-ins), files in a "
                      f"RAM-backed directory, wall clock incl. that file I/O"}
    return rate, desc, sum(times) / len(times)


def cpu_reference_rate(cfg, steps: int, warmup: int):
    """Time the reference's CPU path on a strip of the workload: its own drivers where they are built
    (cpu_reference_rate_drivers), else the reference's headers inside the restated loops (OpenMP over pixels).
    Returns (pixels/s, description dict, seconds per pass)."""
    try:
        got = cpu_reference_rate_drivers(cfg, steps, warmup)
    except Exception:
        got = None
    if got is not None:
        return got
    import oracle
    from fringe_b200 import synth
    o = oracle.load()
    cores = usable_cores()
    o.set_threads(cores)
    sl, sc = cfg["sample"]
    slc = synth.make_stack(cfg["bands"], sl, sc, seed=2)
    npx = sl * sc
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, wts = o.nmap_block(slc, cfg["Nx"], cfg["Ny"], method=NMAP_CODE[cfg["nmap"]], thresh=0.05)
        if cfg["s"]:
            comps = []
            for k, d0 in enumerate(range(0, cfg["bands"], cfg["s"]), start=1):
                own = slc[d0:d0 + cfg["s"]]
                bands = np.concatenate([np.array(comps), own]) if comps else own
                _, _, comp = o.evd_block(np.ascontiguousarray(bands, np.complex64), wts, cfg["Nx"], cfg["Ny"], method=1, mini_stack_count=k)
                comps.append(comp)
            o.evd_block(np.array(comps, np.complex64), wts, cfg["Nx"], cfg["Ny"], method=1)
        else:
            o.evd_block(slc, wts, cfg["Nx"], cfg["Ny"], method=METHOD_CODE[cfg["method"]], variant=cfg["variant"],
                        min_neighbors=cfg["minn"])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    rate = npx * len(times) / sum(times)
    desc = {"value": rate, "unit": "pixels/s", "cores": cores, "kind": o.kind,
            "sample": f"{sl}x{sc} strip of the same {cfg['bands']}-date workload, {len(times)} pass(es), OpenMP over pixels on the "
                      f"{cores} usable cores (sched_getaffinity), OpenBLAS single-threaded per call"}
    return rate, desc, sum(times) / len(times)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, desc, sec = cpu_reference_rate(cfg, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": cfg["metric"], "value": rate, "unit": "pixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": desc["sample"]}, "cpu_baseline": desc,
            "e2e": {"value": rate, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--lines", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from fringe_b200 import synth
    from fringe_b200._lib import lib
    from fringe_b200.engine import Context, nulong

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (environment or nccl.conf) prints a banner there
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    BANDS, NX, NY = cfg["bands"], cfg["Nx"], cfg["Ny"]
    lines, cols = args.lines or cfg["lines"], args.cols or cfg["cols"]
    seq = cfg["s"] > 0
    nmini = -(-BANDS // cfg["s"]) if seq else 0
    halo = lib.fringe_sequential_halo(BANDS, cfg["s"], NY) if seq else NY
    # row partition with halos: rank g owns rows [r0, r1), reads [b0, b1)
    from fringe_b200.partition import row_tile
    r0, r1, b0, b1, first_line, n_lines = row_tile(lines, rank, world, halo)
    blines = b1 - b0
    my_pixels = n_lines * cols
    npb = blines * cols
    nu = nulong(NX, NY)
    mcode, vcode = METHOD_CODE[cfg["method"]], cfg["variant"]
    is_dp = cfg["method"] == "MLE" or vcode == 1

    ctx = Context(local)
    slc = synth.make_stack_torch(BANDS, lines, cols, seed=2, device=dev, row_range=(b0, b1))
    count = torch.empty((blines, cols), dtype=torch.int32, device=dev)
    wts = torch.empty((blines, cols, nu), dtype=torch.int32, device=dev)
    out = torch.zeros((BANDS, blines, cols), dtype=torch.complex64, device=dev)
    tcorr = torch.zeros((nmini if seq else 1, blines, cols), dtype=torch.float32, device=dev)
    comp = torch.zeros((nmini if seq else 1, blines, cols), dtype=torch.complex64, device=dev)
    if seq:
        datum = torch.zeros((nmini, blines, cols), dtype=torch.complex64, device=dev)
        tdatum = torch.zeros((blines, cols), dtype=torch.float32, device=dev)
        adjusted = torch.zeros((BANDS, blines, cols), dtype=torch.complex64, device=dev)
        ctx.nmap_block_device(slc, NX, NY, cfg["nmap"], 0.05, count=count, wts=wts)       # the chain is given the mask
        torch.cuda.synchronize()
    # stacks smaller than twice the L2 (126 MB) would be timed out of cache: evict between iterations
    flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev) if slc.numel() * 8 < (1 << 28) else None

    def step_device():
        if flush is not None:
            flush.zero_()
        if seq:
            ctx._check(lib.fringe_sequential_block(ctx._h, slc.data_ptr(), wts.data_ptr(), cols, blines, BANDS, NX, NY, first_line,
                                                   n_lines, cfg["s"], mcode, -1, out.data_ptr(), tcorr.data_ptr(), comp.data_ptr(),
                                                   datum.data_ptr(), tdatum.data_ptr(), adjusted.data_ptr()))
        else:
            ctx.nmap_block_device(slc, NX, NY, cfg["nmap"], 0.05, count=count, wts=wts)
            ctx.evd_block_device(slc, wts, NX, NY, cfg["method"], variant=vcode, min_neighbors=cfg["minn"], first_line=first_line,
                                 n_lines=n_lines, out=out, tcorr=tcorr[0], comp=comp[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count
    evd_ms, nmap_ms, flush_ms = [], [], 0.0
    if flush is not None:                                  # time of the eviction writes, taken off the step below
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            flush.zero_()
        e1.record(); torch.cuda.synchronize()
        flush_ms = e0.elapsed_time(e1)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
        evd_ms.append(ctx.last_kernel_ms("evd"))
        nmap_ms.append(ctx.last_kernel_ms("nmap"))
    elapsed_ms = ev0.elapsed_time(ev1) - flush_ms
    launches = ctx.launch_count - launches0
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    total_pixels = lines * cols
    value = total_pixels * args.steps / (max_ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launch) ------------------------------
    rows = slice(first_line, first_line + n_lines)
    cnt_int = count[rows]
    stats = ctx.evd_stats()
    if seq:
        # every ministack and the datum connection: flops of the pixels each stage solved, over the device time of the
        # whole step (the solves are > 95 % of it)
        fl = 0.0
        solved = 0
        for k in range(nmini):
            ok = tcorr[k][rows] > 0
            nk = k + min(cfg["s"], BANDS - k * cfg["s"])
            fl += flops_per_launch(nk, float(cnt_int[ok].sum().item()), int(ok.sum().item()), mle=True)
            solved += int(ok.sum().item())
        ok = tdatum[rows] > 0
        fl += flops_per_launch(nmini, float(cnt_int[ok].sum().item()), int(ok.sum().item()), mle=True)
        k_ms = max_ms / args.steps
        shp_sum = float(cnt_int[cnt_int >= 2].sum().item())
        kernel = "k_mle<NT> over 20 ministacks + datum connection (whole device step: the solves are its bulk)"
    else:
        ok = tcorr[0][rows] > 0
        solved = int(ok.sum().item())
        shp_sum = float(cnt_int[ok].sum().item())
        fl = flops_per_launch(BANDS, shp_sum, solved, mle=(cfg["method"] == "MLE"))
        k_ms = float(np.mean(evd_ms))
        kernel = ("k_mle<NT> (exact covariance, PSD gates, inv(|C|), certified inverse iteration, FP64)" if is_dp and BANDS <= 32 else
                  "k_evd_cta<CP> (phase_link, 32 < bands <= 104: CTA per pixel, exact covariance tiles in registers, LDL^T test of |C|, "
                  "EVD fall-back on the S x S Gram matrix raised to G^16, FP64) + k_evd<HR,DP> for the pixels it defers"
                  if is_dp and cfg.get("variant") == 1 and 32 < BANDS <= 104 else
                  "k_evd<HR,DP> (generic any-N kernel, FP64 path)" if is_dp else
                  "k_evd_mma (masked Gram product on a two-term FP16 split, mma.sync.m16n8k16 + FP32 dominant eigenvector + phase ref + tcorr + compressed SLC)")
    achieved = fl / (k_ms * 1e-3) * 1e-12
    peak = ctx.fp64_peak_tflops() if is_dp else ctx.fp32_peak_tflops()
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        try:
            hbm_peak = float(json.load(open(peaks_file))["hbm_gbs"]); hbm_src = "measured"
        except Exception:
            pass
    bytes_px = 16 * BANDS + 4 * nu + 12
    # DRAM traffic of the dominant kernel: dram__bytes_read + dram__bytes_write of its launch in the committed ncu --set full
    # capture of this config (profiles/r2_traffic_<config>.json, written by scripts/ncu_traffic.py), scaled by pixels
    traffic_px, traffic_src = None, None
    tfile = os.path.join(ROOT, "profiles", f"r2_traffic_{args.config}.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile))
            traffic_px, traffic_src = float(tj["bytes_per_pixel"]), f"profiles/{os.path.basename(tfile)} ({tj['kernel']}, {tj['pixels']} px captured)"
        except Exception:
            pass
    roofline = {"kernel": kernel, "bound": "fp64" if is_dp else "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak > 0 else None,
                "peak_source": ("FP64" if is_dp else "FP32") + " FMA microbenchmark run inside this bench (libfringe_b200_prof.so); "
                               "MEASURED_PEAKS.json has no such figure.  Algorithmic flops (SURVEY 8d) of the pixels that were "
                               "solved, over that peak" + ("" if is_dp else ", although the Gram product itself runs on the tensor pipe"),
                "kernel_ms": k_ms, "algorithmic_flops_per_launch": fl,
                "mean_shp": shp_sum / max(solved, 1) if not seq else None, "solved_pixels": solved,
                "hbm_view": {"achieved_gbs": my_pixels * bytes_px / (k_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                             "peak_source": hbm_src, "algorithmic_bytes_per_pixel": bytes_px},
                "nmap_kernel_ms": float(np.mean(nmap_ms)),
                "solver": {"iterations_per_pixel": stats["power_iterations"] / max(stats["pixels"], 1),
                           "factorisations_per_pixel": stats["factorisations"] / max(stats["pixels"], 1)},
                "traffic": traffic_px * my_pixels if traffic_px else None,
                "traffic_note": ("DRAM bytes (read + write) per launch: bytes per pixel of the ncu --set full capture in " + traffic_src +
                                 ", times the pixels of this launch") if traffic_src else "no ncu capture of this config under profiles/"}

    # ---- end to end through the host C ABI --------------------------------------------------
    e2e = None
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    if not args.no_e2e:
        numa_node = bind_to_gpu_numa_node(local)
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
        h_slc = pin(BANDS, blines, cols, dtype=torch.complex64)
        h_slc.copy_(slc)
        h_out = pin(BANDS, blines, cols, dtype=torch.complex64)
        h_tcorr = pin(nmini if seq else 1, blines, cols, dtype=torch.float32)
        h_comp = pin(nmini if seq else 1, blines, cols, dtype=torch.complex64)
        h_count = pin(blines, cols, dtype=torch.int32)
        h_wts = pin(blines, cols, nu, dtype=torch.int32)
        if seq:
            h_wts.copy_(wts)
            h_datum = pin(nmini, blines, cols, dtype=torch.complex64)
            h_tdatum = pin(blines, cols, dtype=torch.float32)
            h_adj = pin(BANDS, blines, cols, dtype=torch.complex64)

            def step_host():
                ctx._check(lib.fringe_sequential_block(ctx._h, h_slc.data_ptr(), h_wts.data_ptr(), cols, blines, BANDS, NX, NY,
                                                       first_line, n_lines, cfg["s"], mcode, -1, h_out.data_ptr(), h_tcorr.data_ptr(),
                                                       h_comp.data_ptr(), h_datum.data_ptr(), h_tdatum.data_ptr(), h_adj.data_ptr()))
            h2d = npb * BANDS * 8 + npb * nu * 4
            d2h = my_pixels * (2 * BANDS * 8 + nmini * (4 + 8 + 8) + 4)
            api = ("fringe_sequential_block (host pointers, pinned): the stack goes up once; ministack phasors, coherences, compressed "
                   "SLCs, datum phasors and the adjusted series come back, per rank")
        else:
            def step_host():
                ctx._check(lib.fringe_nmap_evd_block(ctx._h, h_slc.data_ptr(), None, None, cols, blines, BANDS, NX, NY,
                                                     NMAP_CODE[cfg["nmap"]], 0.05, first_line, n_lines, mcode, -1, 1, vcode, cfg["minn"],
                                                     h_count.data_ptr(), h_wts.data_ptr(), h_out.data_ptr(), h_tcorr.data_ptr(),
                                                     h_comp.data_ptr()))
            h2d = npb * BANDS * 8
            d2h = npb * 4 + npb * nu * 4 + my_pixels * (BANDS * 8 + 4 + 8)
            api = ("fringe_nmap_evd_block (host pointers, pinned; count, mask, phase, tcorr, compressed SLC all copied back), per rank")

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

        t_host = time_host(step_host)
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
        e2e = {"value": total_pixels * args.steps / t_host, "unit": "pixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": t_host * 1e3 / args.steps, "api": api,
               "pinned_buffers": ("allocated and first touched with the process bound to NUMA node %d (the GPU's)" % numa_node)
               if numa_node is not None else "allocated where the process ran (no NUMA binding)",
               "host_link": {"gbs_each_way": link_gbs, "note": "pinned copies of %d MB up and down at the same time on every rank, "
                             "max over ranks" % (nb >> 20), "floor_ms_per_step": max(h2d, d2h) / (link_gbs * 1e9) * 1e3}}

    if affinity0 is not None:
        os.sched_setaffinity(0, affinity0)          # the CPU baseline below gets every usable core again
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        _, cpu, _ = cpu_reference_rate(cfg, 1, 0)

    if rank == 0:
        line = {"metric": cfg["metric"], "value": value, "unit": "pixels/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64" if is_dp else "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config, "lines": lines, "cols": cols, "bands": BANDS,
                           "partition": f"{world} row tile(s) with {halo}-line halos, no collective",
                           "l2": ("inputs (%.1f GB stack) far larger than the 126 MB L2; no explicit flush" % (slc.numel() * 8e-9))
                           if flush is None else "stack smaller than 2x L2: a 256 MB buffer is overwritten before every step "
                                                 "(its own time measured separately and taken off)",
                           "timing": "CUDA events on the launching (torch current) stream, max over ranks"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
