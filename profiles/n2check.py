"""2-rank check of the overlapped slab step against the plain slab step (run under torchrun)"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from marbles_b200.inputs import parse_deck
from marbles_b200.lbm import LBM
from marbles_b200.parallel import HaloComm
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64
deck = parse_deck(text=bench.TG_DECK.format(nx=n, ny=n, nz=n * world, mgs=n))
res = []
for ov in ("1", "0"):
    os.environ["MBL_OVERLAP"] = ov
    comm = HaloComm(rank, world, True, dev)
    lbm = LBM(deck, device=local, rank=rank, world=world, comm=comm, variant=0)
    lbm.init_data()
    lbm.step(7)
    lbm.step(1)
    lbm.step(3, want_macrodata=True)
    torch.cuda.synchronize()
    res.append((lbm.get_f(), lbm.get_g(), lbm.can_overlap()))
    lbm.close()
same = np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
print(f"rank {rank}: overlap used {res[0][2]}/{res[1][2]}, overlapped == plain: {same}, sum f {res[0][0].sum():.12e}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if same else 1)
