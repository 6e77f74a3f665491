"""2-rank check (torchrun) of the plotfile of a DISTRIBUTED hierarchy: amr2_chcyl with its boxes spread over the ranks,
4 coarse steps, plotfile.write_amr_plotfile as a collective call (every rank writes its FABs, rank 0 the headers), read
back on rank 0 with the oracle's reader and compared with the reference's plotfile of that step (the golden)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import load_amr_golden  # noqa: E402
from marbles_b200 import plotfile as P  # noqa: E402
from marbles_b200.amr import AmrLBM  # noqa: E402
from marbles_b200.amr_comm import TorchExchange  # noqa: E402
from marbles_b200.inputs import parse_deck  # noqa: E402
from parity import compare, scales  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
z, deck_text, steps, boxes, is_fluid = load_amr_golden("amr2_chcyl")
amr = AmrLBM(parse_deck(text=deck_text), boxes, is_fluid, device=local, cuda_stream=torch.cuda.current_stream().cuda_stream,
             rank=rank, world=world, exchange=TorchExchange(dev), owners=lambda lev, bxs: [(i + lev) % world for i in range(len(bxs))])
amr.init_data()
s = steps[-1]
amr.step(s, want_macrodata=True)
amr.compute_derived()
forces = torch.tensor(amr.compute_eb_forces(), device=dev)
dist.all_reduce(forces)
tmp = [tempfile.mkdtemp(prefix="amr_n2plot_") if rank == 0 else None]
dist.broadcast_object_list(tmp, 0)
path = P.write_amr_plotfile(amr, tmp[0])
dist.barrier()
ok = True
if rank == 0:
    from oracle import oracle as O
    nan0 = lambda d: {k: np.where(np.isnan(v), 0.0, v) for k, v in d.items()}
    for lev in range(amr.finest + 1):
        pf = O.read_plotfile(path, lev)
        pre = f"s{s}_l{lev}_"
        ref = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        sc = scales(nan0(ref), amr.inp.R, amr.inp.gamma, 2 ** lev / amr.inp.dx[0])
        worst, key = compare(nan0({k: pf[k] for k in ref}), nan0(ref), sc, s * 2 ** lev)
        files = sorted(f for f in os.listdir(os.path.join(path, f"Level_{lev}")) if f.startswith("Cell_D"))
        print(f"level {lev}: {len(ref)} components, worst {worst:.2e} ({key}), data files {files}", flush=True)
    print("eb forces summed over ranks:", forces.cpu().numpy(), flush=True)
# the checkpoint of the distributed hierarchy (collective), loaded by ONE rank into an undistributed object
chk = amr.write_checkpoint_file(tmp[0])
mine = [{w: amr.dense(lev, w) for w in ("f", "g")} for lev in range(amr.finest + 1)]
parts = [None] * world
dist.all_gather_object(parts, mine)
if rank == 0:
    from marbles_b200.amr import merge_dense
    one = AmrLBM.from_checkpoint(parse_deck(text=deck_text), chk, is_fluid, device=local)
    # (solid cells differ by design: the restart zeroes them, fill_f_inside_eb, where the running state holds the -1 sentinel)
    same = True
    for lev in range(one.finest + 1):
        fluid = np.asarray(is_fluid[lev]) == 1
        for w in ("f", "g"):
            a, b = merge_dense([p[lev][w] for p in parts]), one.dense(lev, w)
            same = same and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[:, fluid], b[:, fluid], equal_nan=True)
    files = sorted(os.listdir(os.path.join(chk, "Level_1")))
    print(f"checkpoint {os.path.basename(chk)}: {files}; restored on one rank == distributed state: {same}", flush=True)
    ok = ok and same
    one.close()
dist.barrier()
dist.destroy_process_group()
amr.close()
sys.exit(0 if ok else 1)
