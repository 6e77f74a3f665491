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

    def exchange_macro(self, lbm):
        """the neighbours' adjacent planes of velocity and QCorr into the macrodata ghost planes (compute_derived)"""
        n = int(lbm.lib.mbl_macro_halo_doubles(lbm.ctx, lbm.lev))
        if getattr(self, "mbuf", None) is None or self.mbuf[0].numel() != n:
            self.mbuf = [torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(4)]
        send_lo, send_hi, recv_lo, recv_hi = self.mbuf
        if self.lower is not None:
            check(lbm.lib.mbl_macro_halo(lbm.ctx, lbm.lev, 0, C.c_void_p(send_lo.data_ptr()), 1))
        if self.upper is not None:
            check(lbm.lib.mbl_macro_halo(lbm.ctx, lbm.lev, 1, C.c_void_p(send_hi.data_ptr()), 1))
        exchange_buffers(send_lo, send_hi, recv_lo, recv_hi, self.lower, self.upper, self.group)
        if self.lower is not None:
            check(lbm.lib.mbl_macro_halo(lbm.ctx, lbm.lev, 0, C.c_void_p(recv_lo.data_ptr()), 0))
        if self.upper is not None:
            check(lbm.lib.mbl_macro_halo(lbm.ctx, lbm.lev, 1, C.c_void_p(recv_hi.data_ptr()), 0))

    # ---- overlapped exchange (LBM._step_overlapped) --------------------------------------------
    def _side_stream(self, lbm):
        if getattr(self, "xstream", None) is None:
            # high priority: its small kernels must not queue behind the interior collide's pending CTAs
            self.xstream = torch.cuda.Stream(device=self.device, priority=-1)
            self.ev_boundary = torch.cuda.Event()
            self.ev_exchanged = None
            self.compute = torch.cuda.ExternalStream(lbm.cuda_stream, device=self.device) if lbm.cuda_stream \
                else torch.cuda.default_stream(self.device)
        return self.xstream

    def exchange_next(self, lbm):
        """Asynchronous: once part 0 of the step has written the boundary planes of the buffers being written,
        pack them, swap them with the z-neighbours and unpack them into those buffers' ghost planes -- all on
        the communicator's own stream, so part 1 (interior planes) runs at the same time."""
        xs = self._side_stream(lbm)
        n = int(lbm.lib.mbl_halo_doubles(lbm.ctx, lbm.lev))
        send_lo, send_hi, recv_lo, recv_hi = self._buffers(n)
        self.ev_boundary.record(self.compute)
        xs.wait_event(self.ev_boundary)
        check(lbm.lib.mbl_set_stream(lbm.ctx, C.c_void_p(xs.cuda_stream)))
        try:
            with torch.cuda.stream(xs):
                if self.lower is not None:
                    check(lbm.lib.mbl_halo_pack_next(lbm.ctx, lbm.lev, 0, C.c_void_p(send_lo.data_ptr())))
                if self.upper is not None:
                    check(lbm.lib.mbl_halo_pack_next(lbm.ctx, lbm.lev, 1, C.c_void_p(send_hi.data_ptr())))
                exchange_buffers(send_lo, send_hi, recv_lo, recv_hi, self.lower, self.upper, self.group)
                if self.lower is not None:
                    check(lbm.lib.mbl_halo_unpack_next(lbm.ctx, lbm.lev, 0, C.c_void_p(recv_lo.data_ptr())))
                if self.upper is not None:
                    check(lbm.lib.mbl_halo_unpack_next(lbm.ctx, lbm.lev, 1, C.c_void_p(recv_hi.data_ptr())))
                self.ev_exchanged = torch.cuda.Event()
                self.ev_exchanged.record(xs)
        finally:
            check(lbm.lib.mbl_set_stream(lbm.ctx, C.c_void_p(self.compute.cuda_stream)))

    def wait_exchange(self, lbm):
        """the compute stream waits (on the device) for the last exchange_next"""
        self._side_stream(lbm)
        if self.ev_exchanged is not None:
            self.compute.wait_event(self.ev_exchanged)
            self.ev_exchanged = None

    def allreduce_sum(self, a):
        t = torch.as_tensor(a, dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class LocalSlabs:
    """All z-slabs of one domain held by ONE process on one device: the same pack / unpack / step_local
    sequence as the multi-rank run, with device-to-device copies instead of NCCL.  Used by the GPU parity
    tests to check the slab scheme (one exchange of GZ planes per step, q-correction of the first ghost
    plane recomputed locally) against the single-box run without needing two GPUs."""

    def __init__(self, make_lbm, world: int, periodic_z: bool, device: torch.device):
        self.world, self.periodic_z, self.device = world, periodic_z, device
        self.slabs = [make_lbm(rank, world) for rank in range(world)]
        self.bufs = None

    def exchange(self):
        n = int(self.slabs[0].lib.mbl_halo_doubles(self.slabs[0].ctx, 0))
        if self.bufs is None:
            self.bufs = [[torch.empty(int(s.lib.mbl_halo_doubles(s.ctx, 0)), dtype=torch.float64, device=self.device)
                          for _ in range(2)] for s in self.slabs]
        for s, (lo, hi) in zip(self.slabs, self.bufs):
            check(s.lib.mbl_halo_pack(s.ctx, 0, 0, C.c_void_p(lo.data_ptr())))
            check(s.lib.mbl_halo_pack(s.ctx, 0, 1, C.c_void_p(hi.data_ptr())))
            s.sync()
        for r, s in enumerate(self.slabs):
            lower, upper = neighbours(r, self.world, self.periodic_z)
            if lower is not None:
                check(s.lib.mbl_halo_unpack(s.ctx, 0, 0, C.c_void_p(self.bufs[lower][1].data_ptr())))
            if upper is not None:
                check(s.lib.mbl_halo_unpack(s.ctx, 0, 1, C.c_void_p(self.bufs[upper][0].data_ptr())))
            s.sync()
        del n

    def step(self, nsteps: int = 1, want_macrodata: bool = False):
        for it in range(nsteps):
            self.exchange()
            for s in self.slabs:
                check(s.lib.mbl_step_local(s.ctx, 0, s.time, int(want_macrodata and it == nsteps - 1)))
                s.time += s.dt
                s.isteps += 1
                s.sync()

    def step_overlapped(self, nsteps: int = 1):
        """the split step of LBM._step_overlapped, in its order: part 0 of every slab, the exchange of the
        freshly written boundary planes into the written buffers' ghost planes, part 1"""
        self.exchange()
        for _ in range(nsteps):
            for s in self.slabs:
                check(s.lib.mbl_step_split(s.ctx, 0, 0))
            if self.bufs is None:
                self.exchange()
            for s, (lo, hi) in zip(self.slabs, self.bufs):
                check(s.lib.mbl_halo_pack_next(s.ctx, 0, 0, C.c_void_p(lo.data_ptr())))
                check(s.lib.mbl_halo_pack_next(s.ctx, 0, 1, C.c_void_p(hi.data_ptr())))
                s.sync()
            for r, s in enumerate(self.slabs):
                lower, upper = neighbours(r, self.world, self.periodic_z)
                if lower is not None:
                    check(s.lib.mbl_halo_unpack_next(s.ctx, 0, 0, C.c_void_p(self.bufs[lower][1].data_ptr())))
                if upper is not None:
                    check(s.lib.mbl_halo_unpack_next(s.ctx, 0, 1, C.c_void_p(self.bufs[upper][0].data_ptr())))
                s.sync()
            for s in self.slabs:
                check(s.lib.mbl_step_split(s.ctx, 0, 1))
                s.time += s.dt
                s.isteps += 1
                s.sync()

    def compute_derived(self):
        """compute_derived of every slab with the neighbours' macrodata planes exchanged first"""
        n = int(self.slabs[0].lib.mbl_macro_halo_doubles(self.slabs[0].ctx, 0))
        bufs = [[torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(2)] for _ in self.slabs]
        for s, (lo, hi) in zip(self.slabs, bufs):
            check(s.lib.mbl_macro_halo(s.ctx, 0, 0, C.c_void_p(lo.data_ptr()), 1))
            check(s.lib.mbl_macro_halo(s.ctx, 0, 1, C.c_void_p(hi.data_ptr()), 1))
            s.sync()
        for r, s in enumerate(self.slabs):
            lower, upper = neighbours(r, self.world, self.periodic_z)
            if lower is not None:
                check(s.lib.mbl_macro_halo(s.ctx, 0, 0, C.c_void_p(bufs[lower][1].data_ptr()), 0))
            if upper is not None:
                check(s.lib.mbl_macro_halo(s.ctx, 0, 1, C.c_void_p(bufs[upper][0].data_ptr()), 0))
            check(s.lib.mbl_compute_derived_slab(s.ctx, 0, int(lower is not None), int(upper is not None)))
            s.sync()

    def step_host(self, fabs, ng: int = 0):
        """LBM.step_host of every slab in the multi-rank order: boundary planes up, exchange, pipelined rest.
        fabs: one (f, g) pair of host FAB arrays per slab, updated in place."""
        from .lbm import _dptr
        for s, (f, g) in zip(self.slabs, fabs):
            check(s.lib.mbl_step_host_begin(s.ctx, 0, _dptr(f), _dptr(g), ng))
        self.exchange()
        for s, (f, g) in zip(self.slabs, fabs):
            check(s.lib.mbl_step_host_finish(s.ctx, 0, _dptr(f), _dptr(g), ng))
            s.time += s.dt
            s.isteps += 1

    def gather(self, getter):
        """concatenate a per-slab FAB getter (e.g. LBM.get_f) along z"""
        import numpy as np
        return np.concatenate([getter(s) for s in self.slabs], axis=1)

    def close(self):
        for s in self.slabs:
            s.close()
