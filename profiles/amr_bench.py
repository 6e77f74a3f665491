"""AMR-exact mode at size (BASELINE config 4: channel flow past an EB cylinder, 2 levels), NOT the headline bench.

One coarse step = LBM::time_step(0): fillpatch(0); fillpatch(1); 2 x (physbc(1); stream(1); collide(1)); stream(0);
average_down_to(0, 1); collide(0) -- the reference-granular operator sequence of marbles_b200/amr.py over
marbles_b200/csrc/patch.cu.  Reported: cell updates per second (coarse cells + 2 x fine cells per coarse step), CUDA
events on the stream the kernels run on.  With --reference the unmodified reference (oracle/_ref/marbles3d.omp.ex,
all host cores) runs the same deck with a tagging box over the same region for a few steps (its own grid generator
chooses the fine boxes).

    python profiles/amr_bench.py [--nx 512 --ny 128 --nz 64 --mgs 64 --steps 10 --warmup 3] [--reference]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DECK = """
max_step = {max_step}
geometry.prob_lo = 0.0 0.0 0.0
geometry.prob_hi = {nx}.0 {ny}.0 {nz}.0
geometry.is_periodic = 0 0 1
amr.n_cell = {nx} {ny} {nz}
amr.max_level = 1
amr.max_grid_size = {mgs}
amr.blocking_factor = 8
amr.n_error_buf = 0
amr.regrid_int = 1000000
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
eb2.cylinder_center = {cx}.0 {cy}.0 {cz}.0
eb2.cylinder_has_fluid_inside = 0
eb2.cylinder_height = {height}.0
eb2.cylinder_direction = 2
tagging.refinement_indicators = box
tagging.box.in_box_lo = {rx0}.0 {ry0}.0 -1.0
tagging.box.in_box_hi = {rx1}.0 {ry1}.0 {rz1}.0
amrex.the_arena_is_managed = 0
amrex.fpe_trap_invalid = 0
amrex.fpe_trap_zero = 0
amrex.fpe_trap_overflow = 0
"""


def tiles(lo, hi, mgs):
    """chop the box [lo, hi] (inclusive) into tiles of at most mgs cells per side, x fastest"""
    cuts = []
    for d in range(3):
        n = hi[d] - lo[d] + 1
        k = (n + mgs - 1) // mgs
        edges = [lo[d] + (n * i) // k for i in range(k + 1)]
        cuts.append([(edges[i], edges[i + 1] - 1) for i in range(k)])
    return [((x[0], y[0], z[0]), (x[1], y[1], z[1])) for z in cuts[2] for y in cuts[1] for x in cuts[0]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--ny", type=int, default=128)
    ap.add_argument("--nz", type=int, default=64)
    ap.add_argument("--mgs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="the three reference-granular calls instead of mbl_advance's fused pass")
    ap.add_argument("--ref-steps", type=int, default=3)
    args = ap.parse_args()
    nx, ny, nz, mgs = args.nx, args.ny, args.nz, args.mgs
    # refined region in coarse cells: around the cylinder and its wake, the whole (periodic) span
    rx0, rx1, ry0, ry1 = nx // 8, nx // 2, ny // 4, 3 * ny // 4
    deck = dict(nx=nx, ny=ny, nz=nz, mgs=mgs, rad=max(2, ny // 8), cx=nx // 4, cy=ny // 2, cz=nz // 2,
                height=(3 * nz) // 4, rx0=rx0, rx1=rx1, ry0=ry0, ry1=ry1, rz1=nz + 1)
    coarse = tiles((0, 0, 0), (nx - 1, ny - 1, nz - 1), mgs)
    fine = tiles((2 * rx0, 2 * ry0, 0), (2 * rx1 - 1, 2 * ry1 - 1, 2 * nz - 1), mgs)
    nc = sum((h[0] - l[0] + 1) * (h[1] - l[1] + 1) * (h[2] - l[2] + 1) for l, h in coarse)
    nf = sum((h[0] - l[0] + 1) * (h[1] - l[1] + 1) * (h[2] - l[2] + 1) for l, h in fine)
    updates = nc + 2 * nf
    out = {"workload": f"channel {nx}x{ny}x{nz} + EB cylinder, 2 levels: {len(coarse)} coarse boxes ({nc} cells), "
                       f"{len(fine)} fine boxes ({nf} cells), max_grid_size {mgs}",
           "cell_updates_per_coarse_step": updates}

    if args.reference:
        from oracle import oracle as O  # the reference executable's launcher (oracle/_ref)
        import re
        work = tempfile.mkdtemp(prefix="amr_bench_ref_")
        path = os.path.join(work, "case.inp")
        with open(path, "w") as fh:
            fh.write(DECK.format(max_step=args.ref_steps, **deck))
        t0 = time.time()
        stdout = O.run_reference(path, work, [], omp=True)
        wall = time.time() - t0
        shutil.rmtree(work)
        # LBM::evolve() inclusive time of the reference's TinyProfiler table (initialisation excluded)
        vals = [float(x) for ln in stdout.splitlines() if ln.strip().startswith("LBM::evolve()")
                for x in re.findall(r"[\d.]+(?:[eE][+-]?\d+)?", ln.split("LBM::evolve()")[1])[1:4]]
        grids = [ln.strip() for ln in stdout.splitlines() if "grids" in ln and "Level" in ln]
        per_step = (max(vals) if vals else wall) / args.ref_steps
        out["reference"] = {"s_per_coarse_step": per_step, "MLUPS": updates / per_step / 1e6, "cores": os.cpu_count(),
                            "steps": args.ref_steps, "wall_s": wall, "grids": grids[:4],
                            "how": "LBM::evolve() inclusive time (TinyProfiler) / steps, marbles3d.omp.ex"}
        print(json.dumps(out))
        return

    if args.unfused:
        os.environ["MBL_AMR_FUSED"] = "0"
    import torch
    from marbles_b200.amr import AmrLBM
    from marbles_b200.inputs import parse_deck
    # under torchrun: the boxes of both levels spread over the ranks (contiguous runs of the box lists), NCCL exchange
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    exchange = None
    if world > 1:
        import torch.distributed as dist
        from marbles_b200.amr_comm import TorchExchange
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
        exchange = TorchExchange(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    amr = AmrLBM(parse_deck(text=DECK.format(max_step=1000000, **deck)), [coarse, fine], device=local, cuda_stream=stream,
                 rank=rank, world=world, exchange=exchange)
    amr.init_data()
    barrier()
    setup_s = time.time() - t0
    for _ in range(max(args.warmup, 1)):
        amr.step(1)
    barrier()
    l0 = amr.launches
    if exchange is not None:
        exchange.calls, exchange.host_s, exchange.doubles = 0, 0.0, 0
        exchange.timed = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    e0.record()
    amr.step(args.steps)
    e1.record()
    host_ms = (time.perf_counter() - tw) * 1e3 / args.steps  # host time to ISSUE a coarse step (launch-bound if ~ ms)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if exchange is not None:
        out["exchange_rank0"] = {"calls_per_coarse_step": exchange.calls / args.steps,
                                 "host_ms_per_coarse_step": exchange.host_s * 1e3 / args.steps,
                                 "MB_sent_per_coarse_step": exchange.doubles * 8 / 1e6 / args.steps,
                                 "gpu_ms_per_coarse_step": exchange.gpu_ms() / args.steps}
    out["host_issue_ms_per_coarse_step"] = host_ms
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # sanity: the state is finite and the inlet drives a flow
    import numpy as np
    for lev in range(2):
        f = amr.dense(lev, "f")
        assert np.isfinite(f[~np.isnan(f)]).all()
    out["n_gpus"] = world
    if rank != 0:
        amr.close()
        return
    out.update({"ms_per_coarse_step": ms, "MLUPS": updates / ms / 1e3, "launches_per_coarse_step": (amr.launches - l0) / args.steps,
                "setup_s": setup_s, "steps": args.steps,
                "advance": "un-fused (mbl_stream, mbl_average_down, mbl_collide)" if args.unfused else
                           "mbl_advance (stream + collide fused on the finest level)"})
    print(json.dumps(out))
    amr.close()


if __name__ == "__main__":
    main()
