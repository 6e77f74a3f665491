"""Host logic of DISTRIBUTED multi-box levels on CPU (no device): the split of the FillBoundary copy-tag list into local
copies and per-peer messages (mbl_fill_boundary_plan = what mbl_level_define_boxes_on does), checked against a brute
force over cells, and the pairing of the messages over gloo with world_size 2: what a rank says it sends is what its
peer says it receives, and an exchange of messages of exactly those sizes completes."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def tiles(n, m):
    cuts = [[(a, min(a + m[d] - 1, n[d] - 1)) for a in range(0, n[d], m[d])] for d in range(3)]
    return [((x[0], y[0], z[0]), (x[1], y[1], z[1])) for z in cuts[2] for y in cuts[1] for x in cuts[0]]


def plan(boxes, owner, n, periodic, ng, rank, world):
    from marbles_b200 import _lib
    lib = _lib.load()
    nb = len(boxes)
    lo = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[0]])
    hi = (C.c_int * (3 * nb))(*[v for b in boxes for v in b[1]])
    own = (C.c_int * nb)(*owner)
    dlo, dhi, per = (C.c_int * 3)(0, 0, 0), (C.c_int * 3)(*[v - 1 for v in n]), (C.c_int * 3)(*periodic)
    send, recv, loc = (C.c_int64 * world)(), (C.c_int64 * world)(), C.c_int64()
    _lib.check(lib.mbl_fill_boundary_plan(nb, lo, hi, own, dlo, dhi, per, ng, rank, world, send, recv, C.byref(loc)))
    return list(send), list(recv), int(loc.value)


def brute(boxes, owner, n, periodic, ng, world):
    """cells[a][b] = ghost cells of boxes on rank b that lie on valid cells (periodic images included) of boxes on rank a"""
    cover = np.full(n[::-1], -1)
    for ib, (lo, hi) in enumerate(boxes):
        cover[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] = ib
    cells = np.zeros((world, world), dtype=np.int64)
    for ib, (lo, hi) in enumerate(boxes):
        idx = [np.arange(lo[d] - ng, hi[d] + ng + 1) for d in (2, 1, 0)]
        ok = [np.ones(a.shape, bool) if periodic[d] else (a >= 0) & (a < n[d]) for a, d in zip(idx, (2, 1, 0))]
        w = [a % n[d] if periodic[d] else np.clip(a, 0, n[d] - 1) for a, d in zip(idx, (2, 1, 0))]
        K, J, I = np.meshgrid(*w, indexing="ij")
        m = ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]
        m[ng:-ng, ng:-ng, ng:-ng] = False  # the box's own valid cells
        src = cover[K, J, I]
        m &= src >= 0
        for s in np.unique(src[m]):
            cells[owner[s], owner[ib]] += int((src[m] == s).sum())
    return cells


@pytest.mark.parametrize("n,m,periodic,world,ng", [((16, 16, 16), (8, 8, 8), (1, 1, 1), 2, 3), ((24, 16, 8), (8, 8, 4), (0, 0, 1), 3, 3),
                                                  ((32, 8, 8), (8, 8, 8), (0, 1, 1), 4, 1), ((12, 12, 4), (4, 6, 4), (1, 0, 1), 2, 3)])
def test_fill_boundary_plan_matches_brute_force(n, m, periodic, world, ng):
    from marbles_b200.amr import default_owners
    boxes = tiles(n, m)
    for owner in (default_owners(boxes, world), [i % world for i in range(len(boxes))]):
        ref = brute(boxes, owner, list(n), periodic, ng, world)
        for r in range(world):
            send, recv, loc = plan(boxes, owner, n, periodic, ng, r, world)
            assert loc == ref[r, r]
            for p in range(world):
                if p != r:
                    assert send[p] == ref[r, p] and recv[p] == ref[p, r], (r, p)


def test_default_owners_balance_cells():
    from marbles_b200.amr import default_owners
    boxes = tiles((64, 32, 16), (16, 16, 8))
    for world in (1, 2, 3, 4, 8):
        own = default_owners(boxes, world)
        assert set(own) == set(range(world))
        cells = np.bincount(own, weights=[np.prod([h[d] - l[d] + 1 for d in range(3)]) for l, h in boxes])
        assert cells.max() <= 1.5 * cells.min()
    # compact regions: fewer cells cross ranks than with runs of the x-fastest box list
    n, world = (64, 32, 16), 4
    runs = [min(world - 1, i * world // len(boxes)) for i in range(len(boxes))]
    cross = lambda own: sum(sum(plan(boxes, own, n, (1, 1, 1), 3, r, world)[0]) for r in range(world))
    assert cross(default_owners(boxes, world)) < cross(runs)


def _worker(rank, world, initfile, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from marbles_b200.amr import default_owners
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    n, periodic, ng = (24, 16, 8), (0, 0, 1), 3
    boxes = tiles(n, (8, 8, 4))
    owner = default_owners(boxes, world)
    send, recv, _ = plan(boxes, owner, n, periodic, ng, rank, world)
    # what I send to a peer is what the peer receives from me
    mine = torch.tensor([send, recv], dtype=torch.int64)
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    ok = all(int(allp[a][0][b]) == int(allp[b][1][a]) for a in range(world) for b in range(world) if a != b)
    # messages of exactly those sizes (27 components of f and of g per cell), posted as the exchange functions do
    ops, bufs = [], []
    for p in range(world):
        if p == rank:
            continue
        if send[p]:
            t = torch.full((send[p] * 54,), float(rank), dtype=torch.float64)
            bufs.append(t)
            ops.append(dist.P2POp(dist.isend, t, p))
        if recv[p]:
            t = torch.empty(recv[p] * 54, dtype=torch.float64)
            bufs.append((p, t))
            ops.append(dist.P2POp(dist.irecv, t, p))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ok = ok and all(bool((b[1] == float(b[0])).all()) for b in bufs if isinstance(b, tuple))
    with open(os.path.join(outdir, f"ok{rank}"), "w") as fh:
        fh.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_messages_pair_up_over_gloo():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, initfile, d), nprocs=world, join=True)
        assert all(open(os.path.join(d, f"ok{r}")).read() == "1" for r in range(world))
