"""z-slab decomposition across ranks: one process per GPU, `torch.distributed` for the plumbing.

The level domain is cut into contiguous z-slabs (`lbm.slab_bounds`).  The path has
no data-path collective: the only exchange is the pairwise ghost-plane swap with
the two z-neighbours (the cross-rank part of the reference's FabArray::FillBoundary,
AMReX_FabArrayCommI.H:8-253), done here with batched isend/irecv -- NCCL over
NVLink on GPUs, gloo in the CPU tests.  Each rank sends its GZ outermost valid
planes of f and g (all 27 components; mbl_halo_pack) and receives the neighbour's
into its ghost planes (mbl_halo_unpack).  eb_forces uses one 3-double all-reduce
(Source/LBM.cpp:1040).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check


def neighbours(rank: int, world: int, periodic_z: bool):
    """(lower, upper) z-neighbour ranks, None at a non-periodic end."""
    lower = rank - 1 if rank > 0 else (world - 1 if periodic_z else None)
    upper = rank + 1 if rank < world - 1 else (0 if periodic_z else None)
    return lower, upper


def exchange_buffers(send_lo, send_hi, recv_lo, recv_hi, lower, upper, group=None):
    """Swap halo buffers with the z-neighbours.  Posting order matters when lower == upper (two ranks,
    periodic): sends go (upper, lower), receives (lower, upper), so the first message a peer sends --
    its top planes -- lands in my low ghost planes."""
    ops = []
    if upper is not None:
        ops.append(dist.P2POp(dist.isend, send_hi, upper, group))
    if lower is not None:
        ops.append(dist.P2POp(dist.isend, send_lo, lower, group))
    if lower is not None:
        ops.append(dist.P2POp(dist.irecv, recv_lo, lower, group))
    if upper is not None:
        ops.append(dist.P2POp(dist.irecv, recv_hi, upper, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class HaloComm:
    """Ghost-plane exchange for `LBM` objects on CUDA devices (NCCL)."""

    def __init__(self, rank: int, world: int, periodic_z: bool, device: torch.device, group=None):
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.lower, self.upper = neighbours(rank, world, periodic_z)
        self.buf = None
        self.bytes_per_exchange = 0

    def _buffers(self, n: int):
        if self.buf is None or self.buf[0].numel() != n:
            self.buf = [torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(4)]
            self.bytes_per_exchange = 8 * n * ((self.lower is not None) + (self.upper is not None))
        return self.buf

    def exchange(self, lbm):
        n = int(lbm.lib.mbl_halo_doubles(lbm.ctx, lbm.lev))
        send_lo, send_hi, recv_lo, recv_hi = self._buffers(n)
        if self.lower is not None:
            check(lbm.lib.mbl_halo_pack(lbm.ctx, lbm.lev, 0, C.c_void_p(send_lo.data_ptr())))
        if self.upper is not None:
            check(lbm.lib.mbl_halo_pack(lbm.ctx, lbm.lev, 1, C.c_void_p(send_hi.data_ptr())))
        exchange_buffers(send_lo, send_hi, recv_lo, recv_hi, self.lower, self.upper, self.group)
        if self.lower is not None:
            check(lbm.lib.mbl_halo_unpack(lbm.ctx, lbm.lev, 0, C.c_void_p(recv_lo.data_ptr())))
        if self.upper is not None:
            check(lbm.lib.mbl_halo_unpack(lbm.ctx, lbm.lev, 1, C.c_void_p(recv_hi.data_ptr())))

    def allreduce_sum(self, a):
        t = torch.as_tensor(a, dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()
