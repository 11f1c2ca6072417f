#!/usr/bin/env python3
"""bench.py -- the SHLL time-march hot path on N B200s (one process per GPU), and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--mode strict|fast]
    python bench.py --impl reference ...          # the reference's own C program on the host cores
    python bench.py --gpus N ...  (no torchrun)   # N GPUs from ONE process through shll_group_* (the C drop-in's shape)

A "step" is one time step of the fused kernel over the whole grid (one launch per GPU).  Workloads (BASELINE.json):

    2d_o1      configs[2]: base_shll_2d.c scheme, 4096 x 4096 cells PER GPU (implosion IC, reflective walls).
               N=1 is exactly configs[2]; for N>1 the domain is (4096*N) x 4096, slab-decomposed along x, one halo
               row per side per step -> weak scaling.  This is the default and the headline line.
    2d_o2      configs[4]: 2nd_order_base_shll.c scheme, 2048 x 16384 cells per GPU (16384^2 at 8 GPUs), two halo rows.
    1d_o2      configs[3]: derived 1D 2nd-order, 2^26 cells TOTAL split over N GPUs (strong scaling), fixed step count.
    1d_o1      base_shll.c scheme, 2^26 cells total (strong), for completeness.
    1d_o2_64k  configs[1]: derived 1D 2nd-order, 65 536 cells on one GPU -- the launch-bound regime.

The state (256 MiB per ping-pong buffer for 2d_o1) is larger than the 126 MB L2, so consecutive steps stream from HBM
("l2": "inputs larger than L2" in config); workloads that fit in L2 say so.

JSON keys follow the driver contract; `value` is cell-updates/s with the state resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` is the same metric through the C ABI with host buffers: shll_upload_u from
pinned host memory + K steps + shll_download_u inside the timed region (bytes are amortised per step).
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

WORKLOADS = {
    # name: (program, per-GPU nx, ny, scaling, bytes per cell-update, reference binary, description)
    "2d_o1": dict(prog="base_shll_2d", nx=4096, ny=4096, scaling="weak", bpc=32, ref="ref_2d_o1_4096",
                  desc="configs[2]: base_shll_2d.c 4096x4096 per GPU, 1st order, reflective, implosion IC"),
    "2d_o2": dict(prog="2nd_order_base_shll", nx=2048, ny=16384, scaling="weak", bpc=32, ref="ref_2d_o2_4096",
                  desc="configs[4]: 2nd_order_base_shll.c 2048x16384 per GPU (16384^2 at 8 GPUs), 2-row halos"),
    "1d_o2": dict(prog="2nd_order_base_shll_1d", nx=1 << 26, ny=1, scaling="strong", bpc=24, ref="ref_1d_o2_slice_65536",
                  desc="configs[3]: derived 1D 2nd-order Sod, 2^26 cells total, fixed step count"),
    "1d_o1": dict(prog="base_shll", nx=1 << 26, ny=1, scaling="strong", bpc=24, ref="ref_1d_o1_65536",
                  desc="base_shll.c scheme, 2^26 cells total, fixed step count"),
    "1d_o2_64k": dict(prog="2nd_order_base_shll_1d", nx=65536, ny=1, scaling="strong", bpc=24, ref="ref_1d_o2_slice_65536",
                      desc="configs[1]: derived 1D 2nd-order Sod, 65536 cells (launch-bound regime)"),
}
METRIC = "cell-updates/sec"
UNIT = "cell-updates/s"
L2_BYTES = 126e6


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs about a second before its first sample, longer than a short timed region, so the sampler is started
    ahead of the warm-up, `wait_first()` blocks until it delivers, and every row is stamped on arrival: `window(t0, t1)`
    keeps the samples that fall inside the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.proc = None

    def window(self, t0, t1):
        inside = [r for (t, r) in self.rows if t0 <= t <= t1]
        note = None
        if not inside and self.rows:   # region shorter than the sampling period: the sample closest to it
            mid = 0.5 * (t0 + t1)
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - mid))[1]]
            note = "timed region shorter than the sampling period: nearest sample"
        sm, mx, pw, reasons = [], [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
               "power_w_max": float(max(pw)) if pw else None}
        if note:
            out["note"] = note
        return out


def ref_time_per_step(exe, ncomp, ncells, k1, k2, threads=None):
    """(t(k2) - t(k1)) / (k2 - k1) of the compiled reference program: cancels allocation and first touch."""
    from oracle import oracle as O
    r1 = O.run_ref(exe, ncomp, ncells, step_cap=k1, threads=threads, raw=False)
    r2 = O.run_ref(exe, ncomp, ncells, step_cap=k2, threads=threads, raw=False)
    return (r2["seconds"] - r1["seconds"]) / (k2 - k1), r1["seconds"], r2["seconds"]


def _budget(seconds):
    """CPU-baseline sample length; SHLL_BENCH_CPU_BUDGET scales it (tests use a small factor)."""
    return seconds * float(os.environ.get("SHLL_BENCH_CPU_BUDGET", "1.0"))


def cpu_reference_rate(workload, budget_s=20.0):
    """The reference's own C program (oracle/_ref, kind 'reference') on one host core, bounded sample."""
    from oracle import oracle as O
    budget_s = _budget(budget_s)
    w = WORKLOADS[workload]
    exe = w["ref"]
    if not O.ref_available(exe):
        return None
    # cells the reference binary was built for (compile-time constants)
    if exe.startswith("ref_1d_o2_slice_"):
        n = int(exe.rsplit("_", 1)[1]); ncomp, ncells, cells_eff = 4, n * 4, n * 4   # NY=4 y-uniform run: 4 columns of real work
    elif exe.startswith("ref_1d"):
        n = int(exe.rsplit("_", 1)[1]); ncomp, ncells, cells_eff = 3, n, n
    else:
        n = int(exe.rsplit("_", 1)[1]); ncomp, ncells, cells_eff = 4, n * n, n * n
    # calibrate with a tiny cap, then choose caps that fit the budget
    k1, k2 = 1, 3
    per_step, t1, t2 = ref_time_per_step(exe, ncomp, ncells, k1, k2)
    per_step = max(per_step, 1e-7)
    if per_step * 8 < budget_s:  # the probe was short: take a longer, better-averaged sample inside the budget
        k2 = int(max(4, min(200000, budget_s * 0.6 / per_step)))
        k1 = max(1, k2 // 4)
        per_step, t1, t2 = ref_time_per_step(exe, ncomp, ncells, k1, k2)
    return dict(value=cells_eff / per_step, unit=UNIT, cores=1, kind="reference",
                sample=f"oracle/_ref/{exe} (reference source, gcc -O3) step caps {k1} and {k2}: {t1:.3f}s / {t2:.3f}s, "
                       f"{cells_eff} cells per step", seconds_per_step=per_step)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _replica_times(exe, nrep, cap):
    """nrep copies of one compiled reference program side by side, same step cap: mean in-loop seconds (REF_TIMING)."""
    import re
    import tempfile
    from oracle import oracle as O
    env = dict(os.environ, SHLL_REF_STEP_CAP=str(int(cap)))
    with tempfile.TemporaryDirectory() as td:
        procs = [subprocess.Popen([os.path.join(O.REF_DIR, exe)], cwd=td, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                 for _ in range(nrep)]
        secs = []
        for pr in procs:
            _, err = pr.communicate(timeout=600)
            m = re.search(r"REF_TIMING steps=(-?\d+) seconds=([0-9.eE+-]+)", err)
            if pr.returncode != 0 or not m:
                raise RuntimeError(f"{exe} replica failed: {err[-200:]}")
            secs.append(float(m.group(2)))
    return float(np.mean(secs))


def cpu_all_cores_rate(workload, budget_s=15.0, omp_n=4096):
    """Every host core busy with the reference's code for this workload's scheme (BASELINE.json: 'vs CPU base-c/omp').

    2nd-order 2D: the reference's own OpenMP program (base-omp/2nd_order_base_shll.c, MC limiter -- same cost per cell) with
    one thread per core.  The other schemes have no threaded build in the reference: one single-threaded copy of the
    program per core, side by side (1024^2 / 65536-cell builds so that the copies fit in host memory), rates added up."""
    from oracle import oracle as O
    budget_s = _budget(budget_s)
    cores = host_cores()
    if workload == "2d_o2":
        exe, n = f"ref_omp_o2_{omp_n}", omp_n
        if not O.ref_available(exe):
            return None
        cells = n * n
        k1, k2 = 1, 3
        t = lambda k: O.run_ref(exe, 4, cells, step_cap=k, threads=cores, raw=False)["seconds"]
        t1, t2 = t(k1), t(k2)
        per = max((t2 - t1) / (k2 - k1), 1e-7)
        if per * 8 < budget_s:
            k2 = int(max(4, min(20000, budget_s * 0.6 / per))); k1 = max(1, k2 // 4)
            t1, t2 = t(k1), t(k2)
            per = max((t2 - t1) / (k2 - k1), 1e-7)
        return dict(value=cells / per, unit=UNIT, cores=cores, kind="reference",
                    sample=f"oracle/_ref/{exe} (base-omp source, gcc -fopenmp -O3, {cores} threads, {n}^2) step caps {k1} and {k2}: {t1:.3f}s / {t2:.3f}s")
    exe, cells = {"2d_o1": ("ref_2d_o1_1024", 1024 * 1024), "1d_o1": ("ref_1d_o1_65536", 65536),
                  "1d_o2": ("ref_1d_o2_slice_65536", 65536 * 4), "1d_o2_64k": ("ref_1d_o2_slice_65536", 65536 * 4)}[workload]
    if not O.ref_available(exe):
        return None
    # the copies must fit in host memory (base_shll_2d.c at 1024^2: 49 arrays x 4 MiB): use at most a quarter of what is available
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        cores = max(1, min(cores, int(0.25 * avail / (49 * 4 * cells + (1 << 20)))))
    except Exception:
        pass
    k1, k2 = 2, 6
    t1, t2 = _replica_times(exe, cores, k1), _replica_times(exe, cores, k2)
    per = max((t2 - t1) / (k2 - k1), 1e-7)
    if per * 12 < budget_s:
        k2 = int(max(8, min(200000, budget_s * 0.6 / per))); k1 = max(2, k2 // 4)
        t1, t2 = _replica_times(exe, cores, k1), _replica_times(exe, cores, k2)
        per = max((t2 - t1) / (k2 - k1), 1e-7)
    return dict(value=cores * cells / per, unit=UNIT, cores=cores, kind="reference",
                sample=f"{cores} side-by-side copies of oracle/_ref/{exe} (single-threaded reference source, gcc -O3; {cells} cells each), "
                       f"step caps {k1} and {k2}: mean {t1:.3f}s / {t2:.3f}s; rates added up")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    t0 = time.time()
    base = cpu_reference_rate(args.workload, budget_s=30.0)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/{w['ref']} was not built (needs /root/reference at build time)"}))
        return 0
    # "all the host threads it can use": the reference's OpenMP build where the scheme has one, else one single-threaded copy
    # of the program per core.  The line's value is the better of that and the single program.
    single = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    best = single
    try:
        allc = cpu_all_cores_rate(args.workload, budget_s=25.0)
    except Exception as ex:
        allc = None
        single["all_cores_error"] = str(ex)[:200]
    if allc is not None and allc["value"] > best["value"]:
        best = dict(allc, single_program=single)
    cells_per_step = WORKLOADS[args.workload]["nx"] * (WORKLOADS[args.workload]["ny"] if w["prog"] in ("base_shll_2d", "2nd_order_base_shll") else 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cells_per_step / best["value"] * 1e3, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"],
                   "note": "reference C code on the host cores; ms_per_step = this workload's cells per GPU / value; each run is a step-capped "
                           "sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def run_group_arm(args):
    """`python bench.py --gpus N` WITHOUT torchrun: the same workload through the single-process multi-GPU front end of the C
    ABI (shll_group_*: one slab per GPU behind one handle, one host thread per slab), the shape a C drop-in of the reference
    program has.  Same JSON keys; timing = CUDA events on every slab's stream, the slowest slab counts."""
    import torch
    from dataclasses import replace
    from shll_sve_cfd_b200 import capi, programs
    w = WORKLOADS[args.workload]
    n = args.gpus
    base = programs.PROGRAMS[w["prog"]]
    nx_per = args.nx or w["nx"]
    ny = (args.ny or w["ny"]) if base.dims == 2 else 1
    nx_global = nx_per * n if w["scaling"] == "weak" else nx_per
    pb = base.resized(nx_global, ny) if base.dims == 2 else base.resized(nx_global)
    if base.dims == 2:
        pb = replace(pb, lx=float(nx_global) / float(ny), ly=1.0)
    _, _, _, dtdx, dtdy = programs.time_constants(pb)
    K, W = args.steps, args.warmup
    total_cells = nx_global * ny
    host_in = torch.empty((pb.ncomp, total_cells), dtype=torch.float32, pin_memory=True)
    host_out = torch.empty((pb.ncomp, total_cells), dtype=torch.float32, pin_memory=True)
    host_in.numpy()[...] = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    out = {}
    for mode_name in ([args.mode] if args.no_other_mode else [args.mode, "strict" if args.mode == "fast" else "fast"]):
        mode = capi.MODE_STRICT if mode_name == "strict" else capi.MODE_FAST
        g = capi.Group(pb.dims, pb.nx, pb.ny, ngpus=n, order=pb.order, bc=pb.bc, limiter=pb.limiter, tform=pb.tform, mode=mode,
                       alpha=pb.alpha, dt_on_dx=float(dtdx), dt_on_dy=float(dtdy))
        g.upload_u(host_in.numpy())
        g.run(W)
        l0 = g.launches
        sampler = ClockSampler(0)
        if mode_name == args.mode:
            sampler.start(); sampler.wait_first()
        t0 = time.perf_counter()
        ms = g.run_timed(K)
        t1 = time.perf_counter()
        r = {"ms": ms, "launches": g.launches - l0, "kernel": g.variant(0)}
        if mode_name == args.mode:
            sampler.stop()
            r["clocks"] = sampler.window(t0, t1)
            if not args.no_e2e:
                t0 = time.perf_counter()
                g.upload_u(host_in.numpy()); g.run(K); g.download_u(host_out.numpy())
                r["e2e_s"] = time.perf_counter() - t0
        g.close()
        out[mode_name] = r
    r = out[args.mode]
    peak, peak_src = measured_peak_gbs()
    nloc = total_cells // n
    achieved = w["bpc"] * nloc / (r["ms"] * 1e-3 / K) / 1e9
    line = {
        "metric": METRIC, "value": total_cells * K / (r["ms"] * 1e-3), "unit": UNIT, "n_gpus": n, "steps": K, "warmup": W,
        "ms_per_step": r["ms"] / K, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": w["desc"], "grid_global": [nx_global, ny], "kernel": r["kernel"],
                   "arith_mode": args.mode, "parallelism": f"slab{n}, one process (shll_group_*), one host thread per GPU",
                   "l2": "inputs larger than L2 (no flush needed)" if pb.ncomp * nloc * 4 > L2_BYTES else "state fits in L2"},
        "gpu_launches": r["launches"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": w["bpc"] * nloc, "kernel": r["kernel"], "per": "GPU"},
        "clocks": r.get("clocks"),
    }
    if "e2e_s" in r:
        nbytes = pb.ncomp * total_cells * 4
        line["e2e"] = {"value": total_cells * K / r["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": nbytes / K, "d2h_bytes_per_step": nbytes / K,
                       "seconds": r["e2e_s"], "note": "one shll_group_upload_u (pinned host) + K steps + one shll_group_download_u"}
    others = [m for m in out if m != args.mode]
    if others:
        o = out[others[0]]
        line["other_mode"] = {"arith_mode": others[0], "kernel": o["kernel"], "value": total_cells * K / (o["ms"] * 1e-3), "unit": UNIT,
                              "ms_per_step": o["ms"] / K}
    print(json.dumps(line))
    return 0


FAST_ARITH_NOTE = ("fast: FP32 state, FP32 + explicit FMA arithmetic in face-flux form; tolerance vs the reference C path on every primitive field, "
                   "|x - ref| <= tol*(1+|ref|): 3e-5 on the parity configs and on full-size windows (tests/test_gpu_fast_parity.py), full-length runs per "
                   "tests/conftest.py FAST_TOL_LONG (3e-5 at 1024^2 x 820 steps 1st order, measured 3.7e-6; 1e-4 at 256^2 x 1639 steps 2nd order, measured 3.05e-5 = the reference's own FMA-contraction sensitivity); bitwise independent of the "
                   "GPU count, tile and chunk geometry")
STRICT_ARITH_NOTE = "strict: bit-exact vs the reference C path (FP32 + the reference's two double-promoted expressions per cell)"
MIN_REGION_S = 0.2     # every reported rate comes from >= this much device time (repetitions of the K-step region, median)
EXTRA_WORKLOADS = {"2d_o2": 100, "1d_o2": 100, "1d_o2_64k": 104858}   # the other BASELINE.json configs: steps per timed region


def traffic_entry(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(key)
        if isinstance(e, dict):
            return e.get("bytes"), e.get("source")
        return (e, t.get("_source")) if e is not None else (None, None)
    except Exception:
        return None, None


class Harness:
    """One process per GPU: barriers, max over ranks, problem set-up for a workload."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, xs):
        """element-wise max over ranks of a list of floats"""
        if self.world == 1:
            return list(xs)
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def problem(self, workload, nx=0, ny=0, world=None):
        from dataclasses import replace
        from shll_sve_cfd_b200 import programs
        world = self.world if world is None else world
        w = WORKLOADS[workload]
        base = programs.PROGRAMS[w["prog"]]
        nx_per = nx or w["nx"]
        ny = (ny or w["ny"]) if base.dims == 2 else 1
        nx_global = nx_per * world if w["scaling"] == "weak" else nx_per
        pb = base.resized(nx_global, ny) if base.dims == 2 else base.resized(nx_global)
        if base.dims == 2:
            # keep DX == DY (so DT_ON_DX == DT_ON_DY == 0.125 as in every reference run): the domain is [0, nx/ny] x [0, 1]
            pb = replace(pb, lx=float(nx_global) / float(ny), ly=1.0)
        return pb

    def solver(self, pb, mode):
        from shll_sve_cfd_b200 import slabs
        return slabs.SlabSolver(pb, mode, self.dist if self.world > 1 else None, self.rank, self.world, self.local_rank,
                                gather_device=self.torch.device("cuda", self.local_rank) if self.world > 1 else None)


def timed_reps(h, s, K, sampler=None):
    """Repeat the K-step timed region (CUDA events on the library's stream, barrier + synchronize on both sides of every
    repetition, max over ranks) until >= MIN_REGION_S of device time has been measured; returns the per-repetition times."""
    h.barrier()
    first = h.max_over_ranks([s.run_timed(K)])[0]
    reps = int(min(2000, max(3, np.ceil(1.1 * MIN_REGION_S * 1e3 / max(first, 1e-4)))))
    h.barrier()
    t0 = time.perf_counter()
    ms = []
    for _ in range(reps):
        ms.append(s.run_timed(K))
        if h.world > 1:
            h.dist.barrier()
    h.barrier()
    t1 = time.perf_counter()
    ms = h.max_over_ranks(ms)
    clocks = sampler.window(t0, t1) if sampler is not None else None
    return ms, clocks, reps


def parity_vs_1gpu(h, ss, pb, mode, steps=6):
    """N-GPU correctness signal inside the bench: the slabs of the benched, full-size problem after `steps` steps from the
    initial condition, bit for bit against ONE GPU (rank 0's) marching the whole global domain -- compared through one
    blake2b digest per slab.  The update has no reductions, so the bits may not depend on the decomposition."""
    import hashlib
    from shll_sve_cfd_b200 import programs
    u_loc = ss.initial_state()
    ss.upload(u_loc)
    ss.solver.run(steps)
    out = ss.solver.download_u()
    digest = hashlib.blake2b(np.ascontiguousarray(out).tobytes(), digest_size=16).hexdigest()
    parts = [None] * h.world
    h.dist.all_gather_object(parts, (ss.slab.i0, ss.slab.nx_local, digest))
    ok = None
    if h.rank == 0:
        row = pb.ny if pb.dims == 2 else 1
        u0 = np.empty((pb.ncomp, pb.nx, row), np.float32)   # slab by slab: the temporaries of a 16384^2 initial condition stay small
        for i0, nl, _ in parts:
            u0[:, i0:i0 + nl] = programs.cons_from_prim(pb, programs.initial_primitives(pb, i0=i0, nx_local=nl, nx_global=pb.nx)).reshape(pb.ncomp, nl, row)
        with programs.make_solver(pb, mode, device=h.local_rank) as s1:
            s1.upload_u(u0.reshape(pb.ncomp, -1))
            s1.run(steps)
            full = s1.download_u(u0.reshape(pb.ncomp, -1)).reshape(pb.ncomp, pb.nx, row)
        ok = True
        for i0, nl, dg in parts:
            mine = hashlib.blake2b(np.ascontiguousarray(full[:, i0:i0 + nl]).tobytes(), digest_size=16).hexdigest()
            ok = ok and (mine == dg)
    h.barrier()
    return ok


def measure(h, workload, mode_name, K, W, want_e2e, want_parity, sampler_gpu=None, nx=0, ny=0):
    """Device-resident rate (median repetition), optional end-to-end rate through the C ABI with pinned host buffers, optional
    N-GPU bitwise parity, for one workload in one arithmetic mode."""
    torch = h.torch
    from shll_sve_cfd_b200 import capi
    w = WORKLOADS[workload]
    pb = h.problem(workload, nx, ny)
    mode = capi.MODE_STRICT if mode_name == "strict" else capi.MODE_FAST
    ss = h.solver(pb, mode)
    s = ss.solver
    ny_ = pb.ny if pb.dims == 2 else 1
    nloc = ss.slab.nx_local * ny_
    total_cells = pb.nx * ny_
    ncomp = pb.ncomp
    host_in = torch.empty((ncomp, nloc), dtype=torch.float32, pin_memory=True)   # torch is plumbing: pinned allocation, barriers
    host_in.numpy()[...] = ss.initial_state()
    u_in = host_in.numpy()
    sampler = None
    if sampler_gpu is not None and h.rank == 0:
        sampler = ClockSampler(sampler_gpu)
        sampler.start()
    ss.upload(u_in)
    s.run(W)
    s.sync()
    if sampler is not None:
        sampler.wait_first()
    l0 = s.launches
    hw0 = s.halo_wait_stats() if h.world > 1 else None
    ms, clocks, reps = timed_reps(h, s, K, sampler)
    launches = (s.launches - l0) // (reps + 1)
    halo = None
    if h.world > 1:   # attribution of the scaling loss: time the edge warps spent spinning on a neighbour's flag, per rank
        hw1 = s.halo_wait_stats()
        mine = {"rank": h.rank, "lower_us_per_step": (hw1["lower_s"] - hw0["lower_s"]) * 1e6 / (K * (reps + 1)),
                "upper_us_per_step": (hw1["upper_s"] - hw0["upper_s"]) * 1e6 / (K * (reps + 1)), "waits": hw1["waits"] - hw0["waits"]}
        allr = [None] * h.world
        h.dist.all_gather_object(allr, mine)
        halo = {"halo_steps_per_exchange": getattr(ss, "halo_steps", 1),
                "edge_warp_wait_us_per_step_max_over_ranks": max(max(r["lower_us_per_step"], r["upper_us_per_step"]) for r in allr),
                "edge_warp_wait_us_per_step_per_rank": [[round(r["lower_us_per_step"], 3), round(r["upper_us_per_step"], 3)] for r in allr],
                "note": "summed over the edge warps of a step (2D: one per column tile and side), not wall time: interior warps never wait"}
    if sampler is not None:
        sampler.stop()
    med = float(np.median(ms))
    peak, peak_src = measured_peak_gbs()
    bytes_per_launch = w["bpc"] * nloc
    r = dict(pb=pb, nloc=nloc, total_cells=total_cells, ncomp=ncomp, ms=med, ms_min=float(min(ms)), ms_max=float(max(ms)), reps=reps,
             region_s=float(sum(ms)) * 1e-3, launches=launches, variant=s.variant, clocks=clocks, grid_per_gpu=[ss.slab.nx_local, ny_],
             value=total_cells * K / (med * 1e-3), achieved=bytes_per_launch * K / (med * 1e-3) / 1e9, peak=peak, peak_src=peak_src,
             bytes_per_launch=bytes_per_launch, halo=halo)
    if want_e2e:
        # end to end through the C ABI with host buffers: upload + K steps + download inside the timed region
        host_out = torch.empty((ncomp, nloc), dtype=torch.float32, pin_memory=True)
        u_out = host_out.numpy()
        h.barrier()
        t0 = time.perf_counter()
        ss.upload(u_in)
        t_up = time.perf_counter()
        s.run(K)
        s.sync()
        t_run = time.perf_counter()
        s.download_u(u_out)
        t_dn = time.perf_counter()
        h.barrier()
        t_all = time.perf_counter() - t0
        t_e2e, up, run, dn = h.max_over_ranks([t_all, t_up - t0, t_run - t_up, t_dn - t_run])
        nbytes = ncomp * nloc * 4
        r["e2e"] = {"value": total_cells * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes / K, "d2h_bytes_per_step": nbytes / K,
                    "seconds": t_e2e, "breakdown_s_max_over_ranks": {"upload": up, "steps": run, "download": dn},
                    "pcie_gbs_per_gpu": {"h2d": nbytes / up / 1e9, "d2h": nbytes / dn / 1e9},
                    "note": "one shll_upload_u (pinned host) + K steps + one shll_download_u per run; bytes amortised per step"}
        r["checksum"] = float(u_out[0].astype(np.float64).sum())
    if want_parity and h.world > 1:
        try:
            r["parity_vs_1gpu"] = parity_vs_1gpu(h, ss, pb, mode)
        except Exception as ex:   # a failed check is reported, never hidden
            r["parity_vs_1gpu"] = f"check failed to run: {str(ex)[:160]}"
    ss.close()
    return r


def workload_block(h, name, r, K):
    w = WORKLOADS[name]
    per_launch_note = {}
    if r["launches"] < K:   # persistent kernel: one launch for all K steps
        per_launch_note = {"launches_per_region": r["launches"], "us_per_step": r["ms"] * 1e3 / K}
    tr, tr_src = traffic_entry(f"{name}:fast")
    return {"description": w["desc"], "scaling": w["scaling"], "grid_global": [r["pb"].nx, r["pb"].ny if r["pb"].dims == 2 else 1],
            "grid_per_gpu": r["grid_per_gpu"], "kernel": r["variant"], "arith_mode": "fast", "steps": K, "reps": r["reps"],
            "timed_region_s": r["region_s"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms"] / K,
            "ms_per_step_min_max": [r["ms_min"] / K, r["ms_max"] / K],
            "roofline": {"bound": "hbm", "achieved": r["achieved"], "peak": r["peak"], "unit": "GB/s", "frac": r["achieved"] / r["peak"],
                         "traffic": tr, "traffic_source": tr_src, "algorithmic_bytes_per_step": r["bytes_per_launch"],
                         "steps_per_launch": max(1, K // max(1, r["launches"])) if r["launches"] > 1 else None},
            "clocks": r["clocks"], "parity_vs_1gpu": r.get("parity_vs_1gpu"), "halo_exchange": r.get("halo"), **per_launch_note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps; 0 = the workload's natural run length (3277 steps = t 0.1 for 2d_o1)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="2d_o1", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default=os.environ.get("SHLL_BENCH_MODE", "fast"), choices=["strict", "fast"],
                    help="arithmetic of the headline numbers; the other mode is measured too and reported under 'other_mode'")
    ap.add_argument("--no-other-mode", action="store_true")
    ap.add_argument("--nx", type=int, default=0, help="override per-GPU (weak) / total (strong) nx")
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` block (the other BASELINE.json configs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-GPU vs 1-GPU bitwise check (N > 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.steps <= 0:
        args.steps = {"2d_o1": 3277, "2d_o2": 400, "1d_o2": 400, "1d_o1": 400, "1d_o2_64k": 104858}[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the product path has no CPU fallback", file=sys.stderr)
        return 2
    if world == 1 and args.gpus > 1:   # no torchrun: one process drives all the GPUs through shll_group_*
        return run_group_arm(args)
    h = Harness(args)
    rank = h.rank
    w = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup

    # ---- the headline workload: device-resident rate, end to end, N-GPU parity
    r = measure(h, args.workload, args.mode, K, W, want_e2e=not args.no_e2e, want_parity=not args.no_parity,
                sampler_gpu=h.local_rank, nx=args.nx, ny=args.ny)
    # ---- the other arithmetic mode, device-resident timing only (same workload, same K)
    other = None
    if not args.no_other_mode:
        omode = "strict" if args.mode == "fast" else "fast"
        o = measure(h, args.workload, omode, K, W, want_e2e=False, want_parity=False, nx=args.nx, ny=args.ny)
        other = {"arith_mode": omode, "kernel": o["variant"], "value": o["value"], "unit": UNIT, "ms_per_step": o["ms"] / K,
                 "reps": o["reps"], "roofline_frac": o["achieved"] / o["peak"]}
    # ---- the other BASELINE.json configs, FAST, each over >= MIN_REGION_S of device time
    blocks = {}
    custom = bool(args.nx or args.ny)
    if not args.no_workloads and not custom:
        for name, kw in EXTRA_WORKLOADS.items():
            if name == args.workload or (name == "1d_o2_64k" and world > 1):
                continue
            try:
                rw = measure(h, name, "fast", kw, 10, want_e2e=False, want_parity=not args.no_parity, sampler_gpu=h.local_rank)
                blocks[name] = workload_block(h, name, rw, kw)
            except Exception as ex:
                blocks[name] = {"error": str(ex)[:300]}

    if rank == 0:
        tr, tr_src = traffic_entry(f"{args.workload}:{args.mode}")
        state_bytes = r["ncomp"] * r["nloc"] * 4
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": r["ms"] / K, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "grid_global": [r["pb"].nx, r["pb"].ny if r["pb"].dims == 2 else 1],
                       "grid_per_gpu": r["grid_per_gpu"], "kernel": r["variant"],
                       "arith_mode": FAST_ARITH_NOTE if args.mode == "fast" else STRICT_ARITH_NOTE,
                       "parallelism": f"slab{world}" if world > 1 else "single",
                       "timing": f"the K-step region repeated {r['reps']} times ({r['region_s']:.3f} s of device time, CUDA events, max over ranks per "
                                 f"repetition); value = median repetition; min/max ms per step {r['ms_min'] / K:.5f}/{r['ms_max'] / K:.5f}",
                       "l2": "inputs larger than L2 (no flush needed)" if state_bytes > L2_BYTES else "state fits in L2 (resident between steps by design)"},
            "gpu_launches": r["launches"],
            "reps": r["reps"], "timed_region_s": r["region_s"],
            "e2e": r.get("e2e"),
            "other_mode": other,
            "roofline": {"bound": "hbm", "achieved": r["achieved"], "peak": r["peak"], "unit": "GB/s", "frac": r["achieved"] / r["peak"],
                         "traffic": tr, "traffic_source": tr_src, "peak_source": r["peak_src"],
                         "algorithmic_bytes_per_launch": r["bytes_per_launch"] * (K // max(1, r["launches"])),
                         "steps_per_launch": K // max(1, r["launches"]), "kernel": r["variant"],
                         **({"note": "two time steps per launch (U^(n+1) never leaves the registers): the DRAM traffic of a launch is one read + one "
                                     "write of the state, i.e. half the algorithmic 32 B per cell-update, so frac (algorithmic bytes / measured copy "
                                     "peak) can exceed 1; against the traffic actually moved the launch runs at traffic / launch time"}
                            if r["variant"].endswith("_x2") else {})},
            "clocks": r["clocks"],
            "parity_vs_1gpu": r.get("parity_vs_1gpu"),
            "halo_exchange": r.get("halo"),
            "workloads": blocks,
        }
        if "checksum" in r:
            line["config"]["e2e_checksum_rho"] = r["checksum"]
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_rate(args.workload)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None
                if cb:
                    ac = cpu_all_cores_rate(args.workload)
                    if ac:
                        line["cpu_baseline"]["all_cores"] = ac
            except Exception as ex:  # the baseline is reported context, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        h.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
