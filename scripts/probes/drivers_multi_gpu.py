#!/usr/bin/env python3
"""File-level path on all visible GPUs ("drivers.cpp per-GPU workers never ran on > 1 GPU"): a synthetic
stack on disk, nmap then evd through the C++ block drivers (one worker thread + context per visible GPU), timed with
wall clock (file I/O included) and checked against a single-GPU run of the same command.
usage: python scripts/probes/drivers_multi_gpu.py [lines cols bands]      -> one JSON line"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def run(tag, vrt, root, env_gpus):
    """nmap + evd in a child process (CUDA_VISIBLE_DEVICES decides the worker count)."""
    code = f"""
import sys, time, json
sys.path.insert(0, {ROOT!r})
from fringe_b200.cli import nmap as nmap_cli, evd as evd_cli
t0 = time.perf_counter()
nmap_cli.main(["-i", {vrt!r}, "-o", {root!r} + "/{tag}_nmap", "-c", {root!r} + "/{tag}_count", "-x", "5", "-y", "2", "-r", "500"])
t1 = time.perf_counter()
evd_cli.main(["-i", {vrt!r}, "-w", {root!r} + "/{tag}_nmap", "-o", {root!r} + "/{tag}_evd", "-x", "5", "-y", "2", "-m", "EVD", "-r", "500"])
t2 = time.perf_counter()
print(json.dumps({{"nmap_s": t1 - t0, "evd_s": t2 - t1}}))
"""
    env = dict(os.environ)
    if env_gpus is not None:
        env["CUDA_VISIBLE_DEVICES"] = env_gpus
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    last = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not last:
        raise RuntimeError(out.stdout[-2000:] + out.stderr[-2000:])
    return json.loads(last[-1])


def main():
    import torch
    from fringe_b200 import stackio, synth
    lines, cols, bands = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (1500, 4000, 30)))
    ngpu = torch.cuda.device_count()
    root = tempfile.mkdtemp(prefix="fringe_multi_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        slc = synth.make_stack_torch(bands, lines, cols, seed=2, device=torch.device("cuda", 0)).cpu().numpy()
        vrt = stackio.make_stack_on_disk(root, slc)
        del slc
        one = run("g1", vrt, root, "0")
        res = {"lines": lines, "cols": cols, "bands": bands, "gpus": ngpu, "one_gpu": one,
               "one_gpu_px_s": lines * cols / (one["nmap_s"] + one["evd_s"])}
        if ngpu > 1:
            allg = run("gN", vrt, root, None)
            res["all_gpus"] = allg
            res["all_gpus_px_s"] = lines * cols / (allg["nmap_s"] + allg["evd_s"])
            same = True
            for name in ["_nmap", "_count", "_evd/tcorr.bin", "_evd/compslc.bin"]:
                a = stackio.read_envi(os.path.join(root, "g1" + name)); b = stackio.read_envi(os.path.join(root, "gN" + name))
                same = same and np.array_equal(a.view(np.uint8), b.view(np.uint8))
            res["multi_gpu_rasters_equal_single_gpu"] = bool(same)
        print(json.dumps(res), flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
