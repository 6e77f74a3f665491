#!/usr/bin/env python
"""bench.py -- MLUPS of the D3Q27 f+g (fp64) lattice update on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl ours|reference]

A "step" is one coarse time step of the single-level periodic box (BASELINE.json config 3,
the Taylor-Green deck at SIZE^3 cells per GPU): ghost fill, pull-stream of f and g, macrodata,
q-correction, equilibrium, BGK relax.  N > 1 is launched by torchrun, one rank per GPU, z-slabs,
weak scaling (SIZE^3 per GPU, domain SIZE x SIZE x N*SIZE), ghost planes over NCCL/NVLink.
Rank 0 prints ONE JSON line (contract in the task prompt / DESIGN.md "Measurement").

 value      whole-job MLUPS, state resident in HBM, CUDA events, max over ranks
 e2e        the same metric through the host-buffer call (mbl_step_host): every step uploads f,g from
            pinned host memory and downloads the result
 roofline   dominant kernel (collide pass): 864 algorithmic bytes per cell update / its device time,
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
 cpu_baseline  the unmodified reference (oracle/_ref, OpenMP) on the host cores, bounded sample

--impl reference times the reference's own CPU implementation (oracle/_ref) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_CELL = 864  # (27 f + 27 g) x 8 B x (read + write): SURVEY.md section 8(d)

TG_DECK = """
max_step = 1000000
geometry.prob_lo = -1.0 -1.0 -1.0
geometry.prob_hi =  1.0  1.0  1.0
geometry.is_periodic = 1 1 1
amr.n_cell = {nx} {ny} {nz}
amr.max_level = 0
amr.max_grid_size = {mgs}
amr.plot_int = -1
amr.chk_int = -1
lbm.bc_lo = 0 0 0
lbm.bc_hi = 0 0 0
lbm.dx_outer = 1.0
lbm.dt_outer = 1.0
lbm.nu = 0.1733333333333333
lbm.save_streaming = 0
lbm.ic_type = "taylorgreen"
ic_taylorgreen.rho0 = 1.0
ic_taylorgreen.v0 = 0.1
eb2.geom_type = "all_regular"
amrex.the_arena_is_managed = 0
"""


# Wall / EB workload (BASELINE config 4 at benchmark size, single level): channel inlet,
# zeroth-order outflow, no-slip walls in y, periodic z, EB cylinder along z -- every boundary operator of the path
CHANNEL_DECK = """
max_step = 1000000
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = {nx}.0 {ny}.0 {nz}.0
geometry.is_periodic = 0 0 1
amr.n_cell = {nx} {ny} {nz}
amr.max_level = 0
amr.max_grid_size = {mgs}
amr.plot_int = -1
amr.chk_int = -1
lbm.bc_lo = 2 1 0
lbm.bc_hi = 5 1 0
lbm.dx_outer = 1.0
lbm.dt_outer = 1.0
lbm.nu = 0.0050
lbm.save_streaming = 0
lbm.velocity_bc_type = "channel"
velocity_bc_channel.initial_density = 1.0
velocity_bc_channel.Mach_ref = 0.01
velocity_bc_channel.initial_temperature = 0.03
lbm.ic_type = "constant"
ic_constant.density = 1.0
ic_constant.initial_temperature = 0.03
ic_constant.mach_components = 0.0 0.0 0.0
eb2.geom_type = "cylinder"
eb2.cylinder_radius = {rad}.0
eb2.cylinder_center = {cx}.0 {cy}.0 0.0
eb2.cylinder_has_fluid_inside = 0
eb2.cylinder_height = -1.0
eb2.cylinder_direction = 2
amrex.the_arena_is_managed = 0
"""


def workload_deck(args, world: int):
    """-> (deck text, description, cells per rank are nx*ny*nz/world)"""
    n = args.size
    if args.workload == "channel":
        nx, ny, nz = 2 * n, n // 2, (n // 2) * (world if args.scaling == "weak" else 1)
        text = CHANNEL_DECK.format(nx=nx, ny=ny, nz=nz, mgs=nx, rad=max(2, ny // 8), cx=nx // 4, cy=ny // 2)
        desc = (f"channel {nx}x{ny}x{nz} with an EB cylinder: channel-profile velocity inlet, outflow, no-slip walls, "
                f"periodic z, halfway bounce-back (BASELINE config 4 geometry at benchmark size, single level)")
    else:
        nz = n * world if args.scaling == "weak" else n
        text = TG_DECK.format(nx=n, ny=n, nz=nz, mgs=n)
        per = f"{n}^3 per GPU" if args.scaling == "weak" else f"{n}^3 in total"
        desc = (f"periodic Taylor-Green box {per}, single level, D3Q27 f+g fp64 (BASELINE config 3; domain {n}x{n}x{nz})")
    return text, desc


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_mlups(size: int, steps: int, threads: int | None = None, exe: str | None = None):
    """Run the unmodified reference (oracle/_ref, OpenMP build) on the TG deck at size^3 and return
    (MLUPS, seconds per step, threads).  Per-step time = LBM::evolve() inclusive time of the
    reference's own TinyProfiler table / steps."""
    from oracle import oracle as O
    if exe is None:
        exe = O.REF_OMP if os.path.exists(O.REF_OMP) else O.REF_SERIAL
    if not os.path.exists(exe):
        return None
    threads = threads or os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="mbl_ref_")
    deck = os.path.join(work, "tg.inp")
    with open(deck, "w") as fh:
        fh.write(TG_DECK.format(nx=size, ny=size, nz=size, mgs=max(32, size // 4)))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    t0 = time.time()
    res = subprocess.run([exe, deck, f"max_step={steps}"], cwd=work, env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if res.returncode != 0:
        return None
    evolve = None
    for ln in res.stdout.splitlines():
        m = re.match(r"\s*LBM::evolve\(\)\s+\d+\s+([\d.eE+-]+)\s+([\d.eE+-]+)\s+([\d.eE+-]+)", ln)
        if m and evolve is None and "Incl" not in ln:
            evolve = float(m.group(3))
    # the inclusive table comes second; take the largest figure reported for LBM::evolve()
    vals = [float(x) for ln in res.stdout.splitlines() if ln.strip().startswith("LBM::evolve()")
            for x in re.findall(r"[\d.]+(?:[eE][+-]?\d+)?", ln.split("LBM::evolve()")[1])[1:4]]
    if vals:
        evolve = max(vals)
    sec = (evolve if evolve else wall) / steps
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    return size ** 3 / sec / 1e6, sec, threads, os.path.basename(exe)


def host_memory_bytes():
    try:
        return next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
    except Exception:
        return 32 << 30


def reference_sample_size(steps_total: int, cores: int, budget_s: float = 170.0) -> int:
    """largest box of the reference CPU sample that fits the host memory (the reference keeps ~190 words per cell,
    SURVEY section 8d) and lets `steps_total` steps finish in the time budget at ~0.22 MLUPS per core"""
    mem = host_memory_bytes()
    est = 0.22e6 * max(cores, 1)
    for n in (512, 384, 256, 192, 128, 96, 64):
        if 190 * 8 * n ** 3 * 1.25 < mem and n ** 3 * steps_total / est <= budget_s:
            return n
    return 64


def run_reference_arm(args):
    """The reference's own CPU implementation (unmodified sources, oracle/_ref OpenMP build) on this box's host
    cores.  What runs is said in the line: W warm-up steps as one run, then the K timed steps as three runs of
    about K/3 steps each (median and spread over the runs); each run is the TG deck at `sample` ^3 cells."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, W = max(args.steps, 3), max(args.warmup, 1)
    size = args.ref_size or reference_sample_size(K + W, cores)
    t0 = time.time()
    if reference_mlups(size, W) is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref executable missing or failed"}))
        return
    parts = [K // 3 + (1 if i < K % 3 else 0) for i in range(3)]
    vals = [reference_mlups(size, n) for n in parts if n > 0]
    if any(v is None for v in vals):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref executable failed"}))
        return
    ml = sorted(v[0] for v in vals)
    mlups = statistics.median(ml)
    sec = statistics.median(v[1] for v in vals)
    timed_s = sum(v[1] * n for v, n in zip(vals, parts))
    sample = (f"TG deck at {size}^3 cells (not the {args.size}^3 of the GPU arm: ~190 words per cell and "
              f"~{size ** 3 / (mlups * 1e6):.1f} s per step on {vals[0][2]} cores), {W} warm-up steps, then {K} timed steps as "
              f"{len(vals)} runs of {'/'.join(str(n) for n in parts)} steps; per-step time = LBM::evolve() inclusive "
              f"(TinyProfiler) / steps; median of the runs, spread {100 * (ml[-1] - ml[0]) / mlups:.0f} % ({vals[0][3]})")
    line = {
        "impl": "reference", "metric": "MLUPS (D3Q27 f+g, fp64)", "value": mlups, "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"periodic Taylor-Green box, single level, D3Q27 f+g fp64 (BASELINE config 3), reference CPU "
                               f"sample at {size}^3 cells per step (the GPU arm runs {args.size}^3 per GPU)",
                   "sample_size": size, "gpu_arm_size": args.size},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": vals[0][2], "kind": "reference", "sample": sample,
                         "runs_mlups": ml, "spread_pct": 100 * (ml[-1] - ml[0]) / mlups},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_executed": sum(parts), "warmup_executed": W, "timed_region_s": timed_s, "wall_s": time.time() - t0,
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512, help="cells per side per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-size", type=int, default=None,
                    help="box of the reference CPU sample (default: the largest that fits memory and a ~3 min run)")
    ap.add_argument("--cpu-size", type=int, default=128, help="box of the cpu_baseline sample of the GPU arm (10-30 s)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--workload", default="tg", choices=["tg", "channel"],
                    help="tg: periodic Taylor-Green box (BASELINE config 3, the headline); channel: inlet / outflow / "
                         "no-slip walls / periodic z / EB cylinder at benchmark size")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: SIZE^3 per GPU; strong: SIZE^3 in total, cut into z-slabs")
    ap.add_argument("--ref-gpu-size", type=int, default=384, help="box of the reference-CUDA-build sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variant", type=int, default=None, help="step implementation (default: the library's, 9 = tile carry step marching through z-chunks); 0 two kernels, "
                         "5 tile carry step, 7 with plane pairs; 1-4 and 8 only with MBL_EXPERIMENTS=1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from marbles_b200.inputs import parse_deck
    from marbles_b200.lbm import LBM
    from marbles_b200.parallel import HaloComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keeps NCCL's version banner out of stdout (one JSON line)
        # NCCL's own stream at high priority: the halo send/recv kernels run beside the interior collide
        # kernel instead of queueing behind its pending CTAs
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    n = args.size
    deck_text, workload_desc = workload_deck(args, world)
    deck = parse_deck(text=deck_text)
    comm = HaloComm(rank, world, True, dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream
    lbm = LBM(deck, device=local, rank=rank, world=world, comm=comm, cuda_stream=stream, variant=args.variant)
    lbm.init_data()
    cells_total = lbm.ncells * world
    cells_rank, lbm_shape, halo_lean = lbm.ncells, "x".join(str(v) for v in lbm.n_local), bool(getattr(lbm, "halo_lean", False))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        lbm.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    import ctypes as C
    lbm.lib.mbl_set_timing(lbm.ctx, 1)
    launches0 = lbm.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if world == 1:
        lbm.step(args.steps)
    else:
        for _ in range(args.steps):
            lbm.step(1)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    kms = (C.c_double * 3)()
    nrec = C.c_int()
    lbm.lib.mbl_get_timing(lbm.ctx, kms, C.byref(nrec))
    lbm.lib.mbl_set_timing(lbm.ctx, 0)
    launches = lbm.launches - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if world > 1:
        t = torch.tensor([float(lbm.ncells)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        cells_total = int(t.item())
    ms_per_step = ms / args.steps
    value = cells_total / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (collide pass) -----------------------------
    peak, peak_src = measured_peak_gbs()
    collide_ms = kms[2] / max(nrec.value, 1)
    kname = {0: "k_collide_lean (pull + collide)", 1: "k_fused (q-correction + collide jobs)", 2: "k_fused (collide jobs)",
             3: "k_fused_plain (q-correction + collide jobs)", 4: "k_collide_carry", 5: "k_collide_tile", 6: "k_collide_lean", 7: "k_collide_tile_pair",
             8: "k_march (pull + q-correction + collide, one kernel)",
             9: "k_collide_tile_march", 10: "k_collide_tile_march (pipelined)"}[lbm.variant]
    vname = {0: "two kernels (k_qcorr, k_collide_lean)", 1: "one persistent TMA-pipelined kernel per step",
             2: "persistent TMA kernel, two launches (q-correction, collide)",
             3: "one persistent kernel per step, plain loads",
             4: "carry step: k_qcorr_combine (row sums -> q-corrections) + k_collide_carry (collide, emits the next "
                "step's moment row sums)",
             5: "carry step without marching: k_qcorr_combine + k_collide_tile",
             6: "two kernels, collide with g staged through shared memory (k_qcorr, k_collide_lean)",
             7: "tile carry step with plane pairs: k_qcorr_combine_pair + k_collide_tile_pair",
             8: "march step: one kernel per step, z-marching CTAs, q-corrections recomputed on a one-cell halo",
             9: "tile carry step marching through z-chunks: k_qcorr_combine_march + k_collide_tile_march (z sums completed "
                "on chip, QCorr of most cells finished by the collide kernel)",
             10: "z-march tile carry step with plane k+1 pulled into shared memory while plane k is collided"}[lbm.variant]
    achieved = BYTES_PER_CELL * lbm.ncells / (collide_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tkey = {0: "k_collide_lean", 6: "k_collide_lean", 5: "k_collide_tile", 4: "k_collide_carry", 7: "k_collide_tile_pair", 8: "k_march", 9: "k_collide_tile_march", 10: "k_collide_tile_march_pipe"}.get(lbm.variant)
            traffic = json.load(fh).get("dram_bytes_per_launch_512", {}).get(tkey) if n == 512 else None
            if args.workload != "tg" or world > 1 and args.scaling == "strong":
                traffic = None
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel at 512^3, recorded in "
                           "profiles/traffic.json from the capture under profiles/ (a run under ncu is never a bench "
                           "value; regenerate when the kernel changes)") if traffic else None,
        "peak_source": peak_src,
        "bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": lbm.ncells,
        "kernel_ms": {"ghost_fill": kms[0] / max(nrec.value, 1), "qcorr": kms[1] / max(nrec.value, 1),
                      "collide": collide_ms},
        "step_achieved": BYTES_PER_CELL * lbm.ncells / (ms_per_step * 1e-3) / 1e9,
        "step_frac": BYTES_PER_CELL * lbm.ncells / (ms_per_step * 1e-3) / 1e9 / peak,
    }

    # ---- end to end through the host-buffer call -------------------------------------
    e2e = None
    if not args.no_e2e and args.workload == "tg" and args.scaling == "weak":
        from marbles_b200.lbm import _dptr
        from marbles_b200._lib import check
        # host FABs of f and g: 2 x 27 x 8 B per cell per rank, pinned.  Keep them within half of the free host
        # memory shared by the ranks of this node: if the full box does not fit, the e2e leg runs the same
        # deck with fewer planes per rank (the metric is per cell and PCIe-bound, so it does not depend on nz)
        nz_e = n
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            fit = int(0.5 * avail / local_world // (2 * 27 * 8 * n * n))
            nz_e = max(16, min(n, fit))
        except Exception:
            pass
        if nz_e != n:
            lbm.close()
            deck_e = parse_deck(text=TG_DECK.format(nx=n, ny=n, nz=nz_e * world, mgs=n))
            lbm = LBM(deck_e, device=local, rank=rank, world=world, comm=comm, cuda_stream=stream, variant=args.variant)
            lbm.init_data()
            lbm.step(1)
        nx, ny, nz = lbm.n_local
        shape = (27, nz, ny, nx)
        try:
            fh_t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            gh_t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            pinned = True
        except Exception:
            fh_t, gh_t, pinned = torch.empty(shape, dtype=torch.float64), torch.empty(shape, dtype=torch.float64), False
        fh, gh = fh_t.numpy(), gh_t.numpy()
        check(lbm.lib.mbl_download(lbm.ctx, 0, 0, _dptr(fh), 0))
        check(lbm.lib.mbl_download(lbm.ctx, 0, 1, _dptr(gh), 0))
        def host_step():
            # N > 1: LBM.step_host uploads the boundary planes first, exchanges them, and pipelines the rest
            lbm.step_host(fh, gh, 1, ng=0)
        host_step()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nbytes = 2 * fh.nbytes
        e2e = {"value": lbm.ncells * world * args.e2e_steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
               "steps": args.e2e_steps, "pinned": pinned, "ms_per_step": dt / args.e2e_steps * 1e3,
               "box_per_gpu": [nx, ny, nz],
               "path": "mbl_step_host: z-chunked uploads/downloads overlapped with the kernels" if world == 1 else
                       "mbl_step_host_begin (boundary planes), halo exchange, mbl_step_host_finish (pipelined) per rank"}
        del fh, gh, fh_t, gh_t

    lbm.close()

    # ---- CPU baseline: the unmodified reference on the host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = reference_mlups(args.cpu_size, args.cpu_steps)
        if r is not None:
            cpu = {"value": r[0], "unit": "MLUPS", "cores": r[2], "kind": "reference",
                   "sample": f"bounded sample: TG deck at {args.cpu_size}^3 cells, {args.cpu_steps} steps, "
                             f"LBM::evolve() inclusive time, {r[3]} (the --impl reference arm runs the longer sample)"}

    # ---- the reference's own GPU path (unmodified sources, nvcc build of oracle/refbuild) on this GPU ------------
    ref_gpu = None
    cuda_exe = os.path.join(ROOT, "oracle", "_ref", "marbles3d.cuda.ex")
    if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(cuda_exe):
        torch.cuda.empty_cache()
        r = reference_mlups(args.ref_gpu_size, 10, threads=1, exe=cuda_exe)
        if r is not None:
            ref_gpu = {"value": r[0], "unit": "MLUPS", "kind": "reference CUDA build (AMReX ParallelFor kernels, sm_100a)",
                       "sample": f"TG deck {args.ref_gpu_size}^3 (its ~190 words per cell do not fit 512^3 in 180 GB), "
                                 f"10 steps, LBM::evolve() inclusive time, {r[3]}"}

    if rank == 0:
        line = {
            "metric": "MLUPS (D3Q27 f+g, fp64)", "value": value, "unit": "MLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc,
                       "decomposition": f"{world} z-slab(s) of {lbm_shape}" + (", lean z-halo" if halo_lean else ""),
                       "l2": f"state ({54 * 8 * cells_rank / 1e9:.1f} GB per GPU) is far larger than the 126 MB L2",
                       "variant": vname},
            "roofline": roofline, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
