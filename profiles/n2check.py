"""2-rank check (run under torchrun) of the overlapped slab step against the plain (blocking) slab step, and of the
lean z-halo against the full one, on real NCCL transport: periodic Taylor-Green box and the channel + cylinder deck
(inlet / outflow / no-slip walls / periodic z: ghost fill inside the split step).  Two-kernel step: bit-identical;
default step (z-march carry, whose split chunks the planes differently): within 1e-11 of scale."""
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
decks = {"tg": bench.TG_DECK.format(nx=n, ny=n, nz=n * world, mgs=n),
         "channel": bench.CHANNEL_DECK.format(nx=2 * n, ny=n // 2, nz=(n // 2) * world, mgs=2 * n, rad=4, cx=n // 2, cy=n // 4)}
ok = True
for name, text, variant in [(n_, t_, v_) for v_ in (0, None) for n_, t_ in decks.items()]:
    deck = parse_deck(text=text)
    res = []
    for ov, lean in (("1", "1"), ("0", "0")):
        os.environ["MBL_OVERLAP"], os.environ["MBL_HALO_LEAN"] = ov, lean
        comm = HaloComm(rank, world, True, dev)
        lbm = LBM(deck, device=local, rank=rank, world=world, comm=comm, variant=variant)
        lbm.init_data()
        lbm.step(7)
        lbm.step(1)
        lbm.step(3, want_macrodata=True)
        torch.cuda.synchronize()
        res.append((lbm.get_f(), lbm.get_g(), lbm.can_overlap(), lbm.halo_lean))
        lbm.close()
    if variant == 0:
        same = np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    else:
        same = all(np.abs(res[0][i] - res[1][i]).max() <= 1e-11 * np.abs(res[1][i]).max() for i in (0, 1))
    ok = ok and same
    print(f"rank {rank} {name} variant {variant}: overlap {res[0][2]}/{res[1][2]} with MBL_OVERLAP 1/0, lean halo {res[0][3]}/{res[1][3]}, "
          f"overlapped+lean == blocking+full: {same}, sum f {res[0][0].sum():.12e}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
