"""N-rank check (run under torchrun, one rank per GPU) of DISTRIBUTED multi-box levels on real NCCL transport: the
golden hierarchies (static, regridded, with levels that appear and vanish) with their boxes spread over the ranks
against the one-rank run of the same case on rank 0: every FAB bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import amr_regrid_actions, load_amr_golden  # noqa: E402
from marbles_b200.amr import AmrLBM, merge_dense  # noqa: E402
from marbles_b200.amr_comm import TorchExchange  # noqa: E402
from marbles_b200.inputs import parse_deck  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream().cuda_stream
ok = True
for case in ("amr2_tg", "amr3_chcyl", "amr2_sod_regrid", "amr2_tg_appear", "amr2_sod_bc"):
    z, deck_text, steps, boxes, is_fluid = load_amr_golden(case)
    existing = [b for b in boxes if b]
    deck = parse_deck(text=deck_text)
    nsteps = min(steps[-1], 8)

    def drive(amr):
        current = {lev: boxes[lev] for lev in range(1, len(boxes))}
        for done in range(nsteps):
            for lev, what, nb in amr_regrid_actions(z, done + 1, current):
                if what == "make":
                    amr.make_level_from_coarse(lev, nb, is_fluid[lev])
                elif what == "remake":
                    amr.regrid_level(lev, nb, is_fluid[lev])
                else:
                    amr.clear_level(lev)
            amr.step(1, want_macrodata=done + 1 == nsteps)
        amr.sync()
        return [{w: amr.dense(lev, w) for w in ("f", "g", "macro")} for lev in range(amr.finest + 1)]

    amr = AmrLBM(deck, existing, is_fluid[:len(existing)], device=local, cuda_stream=stream, rank=rank, world=world,
                 exchange=TorchExchange(dev), owners=lambda lev, bxs: [(i + lev) % world for i in range(len(bxs))])
    amr.init_data()
    mine = drive(amr)
    launches = amr.launches
    amr.close()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        one = AmrLBM(deck, existing, is_fluid[:len(existing)], device=local, cuda_stream=stream)
        one.init_data()
        ref = drive(one)
        one.close()
        same = all(np.array_equal(merge_dense([p[lev][w] for p in parts]), ref[lev][w], equal_nan=True)
                   for lev in range(len(ref)) for w in ("f", "g", "macro"))
        ok = ok and same
        print(f"{case}: {world} ranks, {nsteps} coarse steps, {len(ref)} levels at the end, {launches} launches on rank 0, "
              f"distributed == one rank: {same}", flush=True)
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
