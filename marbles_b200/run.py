"""`python -m marbles_b200.run deck.inp [key=value ...]` -- the reference's `main` / `LBM::evolve` loop
(Source/main.cpp:10-25, Source/LBM.cpp:155-195, 398-448) for single-level decks, with the lattice update on the
GPU: same input deck format, same plotfile / checkpoint names, formats and cadence (amr.plot_int, amr.chk_int,
amr.restart, max_step, stop_time), so a deck that runs under the reference executable runs here unchanged.

This is the thin caller either side of the hot path, not a re-implementation of the
reference's AMR driver: amr.max_level > 0 is refused."""
from __future__ import annotations

import os
import sys

from .inputs import parse_deck
from .lbm import LBM, MarblesError
from .plotfile import write_lbm_plotfile


def _get(deck: dict, key: str, default, cast=str):
    v = deck.get(key, None)
    if v is None:
        return default
    if isinstance(v, (list, tuple)):
        v = v[0]
    return cast(str(v).strip('"'))


def evolve(lbm: LBM, out_dir: str = ".", log=print) -> list[str]:
    """LBM::init_data's output + LBM::evolve.  Returns the paths written."""
    deck = lbm.inp.deck
    max_step = _get(deck, "max_step", 2 ** 31 - 1, int)
    stop_time = _get(deck, "stop_time", float("inf"), float)
    plot_file, plot_int = _get(deck, "amr.plot_file", "plt"), _get(deck, "amr.plot_int", -1, int)
    chk_file, chk_int = _get(deck, "amr.chk_file", "chk"), _get(deck, "amr.chk_int", -1, int)
    restart = _get(deck, "amr.restart", "")
    digits = _get(deck, "amr.file_name_digits", 5, int)  # Source/LBM.cpp:215
    if _get(deck, "amr.max_level", 0, int) > 0:
        raise MarblesError("marbles_b200.run drives single-level decks (amr.max_level = 0)")
    if stop_time < float("inf") and abs(stop_time - round(stop_time)) > 1.0e-9:
        # the reference shortens the LAST lattice step to dt = stop_time - t (compute_dt, Source/LBM.cpp:1067-1071),
        # which changes the relaxation of that step; the fused step runs whole steps only
        raise MarblesError("stop_time must be a whole number of lattice time steps (dt = 1)")
    written = []
    # lbm.compute_forces: one line of EB forces per step (open_forces_file / output_forces_file,
    # Source/LBM.cpp:1925-1969: width 24, 16 significant digits); forces every step means stepping one at a time
    forces_path = None
    if _get(deck, "lbm.compute_forces", 0, int):
        forces_path = os.path.join(out_dir, _get(deck, "lbm.forces_file", "forces.txt"))

    def forces_line():
        if forces_path is None:
            return
        f3 = lbm.compute_eb_forces()
        with open(forces_path, "a") as fh:
            fh.write("".join("%24s" % ("%.16g" % v) for v in (lbm.time, f3[0], f3[1], f3[2])) + "\n")

    def plot():
        written.append(write_lbm_plotfile(lbm, out_dir, plot_file, digits=digits))
        log(f"Writing plot file {written[-1]} at time {lbm.time}")

    def chk():
        written.append(lbm.write_checkpoint_file(out_dir, chk_file, digits=digits))
        log(f"Writing checkpoint file {written[-1]} at time {lbm.time}")

    if restart:
        lbm.read_checkpoint_file(restart if os.path.isabs(restart) else os.path.join(out_dir, restart))
        log(f"Restarting from checkpoint file {restart}")
        if forces_path is not None and not os.path.exists(forces_path):
            open(forces_path, "w").write("".join("%24s" % h for h in ("time", "fx", "fy", "fz")) + "\n")
    else:
        lbm.init_data()
        if chk_int > 0:
            chk()
        if forces_path is not None:
            open(forces_path, "w").write("".join("%24s" % h for h in ("time", "fx", "fy", "fz")) + "\n")
            forces_line()
    if plot_int > 0:
        lbm.f_to_macrodata()  # the state as initialised / read: macrodata without a step
        plot()
    last_plot = 0
    while lbm.isteps < max_step and lbm.time < stop_time:
        # run to the next step at which something is written
        nxt = max_step
        for every in (plot_int, chk_int):
            if every > 0:
                nxt = min(nxt, (lbm.isteps // every + 1) * every)
        if stop_time < float("inf"):
            nxt = min(nxt, lbm.isteps + max(1, int((stop_time - lbm.time) / lbm.dt + 1e-6)))
        n = nxt - lbm.isteps
        want_plot = plot_int > 0 and nxt % plot_int == 0
        # the closing plotfile (below) is written whenever the run ends off the plot cadence, by max_step or by
        # stop_time: its macrodata must belong to the last step (the reference recomputes macrodata every step)
        ends_run = nxt >= max_step or lbm.time + n * lbm.dt >= stop_time - 1.0e-6 * lbm.dt
        want_macro = want_plot or (plot_int > 0 and ends_run) or nxt >= max_step
        if forces_path is None:
            lbm.step(n, want_macrodata=want_macro)
        else:
            for s in range(n):
                lbm.step(1, want_macrodata=(s == n - 1) and want_macro)
                forces_line()
        if want_plot:
            last_plot = lbm.isteps
            plot()
        if chk_int > 0 and lbm.isteps % chk_int == 0:
            chk()
        if lbm.time >= stop_time - 1.0e-6 * lbm.dt:
            break
    if plot_int > 0 and lbm.isteps > last_plot:
        plot()
    return written


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print(__doc__)
        return 2
    deck = parse_deck(argv[0], overrides=argv[1:])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # one process per GPU (torchrun): z-slabs, ghost planes over NCCL; every rank writes its part of a plotfile
        import torch
        import torch.distributed as dist

        from .inputs import lbm_inputs
        from .parallel import HaloComm
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        comm = HaloComm(rank, world, bool(lbm_inputs(deck).periodic[2]), dev)
        lbm = LBM(deck, device=local, rank=rank, world=world, comm=comm)
        log = print if rank == 0 else (lambda s: None)
    else:
        lbm, log = LBM(deck), print
    try:
        evolve(lbm, log=log)
    finally:
        lbm.close()
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
