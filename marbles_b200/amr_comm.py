"""Exchanges of the per-peer device messages of distributed multi-box levels (mbl_set_exchange).

The library packs the cells that cross ranks into one message per peer and operator and unpacks what arrives; moving
the messages is the caller's part, as FabArray::FillBoundary / ParallelCopy leave it to MPI in the reference
(AMReX_FabArrayCommI.H:8-253).  Two movers:

* `TorchExchange`  -- one process per GPU, `torch.distributed` (NCCL) batched isend / irecv on the library's buffers
                      (wrapped as tensors through `__cuda_array_interface__`, no copy);
* `ThreadExchange` -- several ranks as threads of ONE process on one device (tests): a rendezvous per pair of ranks
                      and a device-to-device copy.

An exchange object is called as `exchange(rank, parts, stream)` with parts = [(peer, send_ptr, n_send, recv_ptr,
n_recv)] (counts in doubles); every rank makes the same sequence of calls.
"""
from __future__ import annotations

import queue


class _DevArray:
    """n doubles of device memory at ptr, for torch.as_tensor"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def dev_tensor(ptr: int, n: int, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, n), device=device)


class TorchExchange:
    """pairwise NCCL send / recv, ordered on the current stream (the one the library was given: mbl_set_stream)"""

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.calls, self.host_s, self.doubles = 0, 0.0, 0  # statistics: exchanges, host time inside them, doubles sent
        self.timed, self._events = False, []               # timed: CUDA events around every exchange (gpu_ms())

    def __call__(self, rank: int, parts, stream: int):
        import time

        import torch
        import torch.distributed as dist
        t0 = time.perf_counter()
        lib_stream = torch.cuda.ExternalStream(stream, device=self.device) if stream else torch.cuda.default_stream(self.device)
        keep, ops = [], []
        with torch.cuda.stream(lib_stream):
            if self.timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            for peer, sp, ns, rp, nr in parts:
                if ns:
                    t = dev_tensor(sp, ns, self.device)
                    keep.append(t)
                    ops.append(dist.P2POp(dist.isend, t, peer, self.group))
                if nr:
                    t = dev_tensor(rp, nr, self.device)
                    keep.append(t)
                    ops.append(dist.P2POp(dist.irecv, t, peer, self.group))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()  # stream-ordered for NCCL: the unpack kernels queue behind the receives
            if self.timed:
                e1.record()
                self._events.append((e0, e1))
        self.calls += 1
        self.host_s += time.perf_counter() - t0
        self.doubles += sum(p[2] for p in parts)


    def gpu_ms(self) -> float:
        """device time between the start and the end of every timed exchange (waiting for the peer included)"""
        import torch
        torch.cuda.synchronize(self.device)
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        self._events = []
        return ms


class ThreadExchange:
    """ranks = threads of one process on one device"""

    def __init__(self, world: int, device=None, timeout: float = 120.0):
        self.device, self.timeout = device, timeout
        pairs = [(a, b) for a in range(world) for b in range(world) if a != b]
        self.msg = {p: queue.Queue() for p in pairs}
        self.ack = {p: queue.Queue() for p in pairs}

    def __call__(self, rank: int, parts, stream: int):
        import torch
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        torch.cuda.synchronize(dev)  # my messages are packed
        for peer, sp, ns, rp, nr in parts:
            if ns:
                self.msg[(rank, peer)].put((sp, ns))
        for peer, sp, ns, rp, nr in parts:
            if nr:
                src, n = self.msg[(peer, rank)].get(timeout=self.timeout)
                if n != nr:
                    raise RuntimeError(f"rank {rank}: rank {peer} sends {n} doubles, {nr} expected")
                dev_tensor(rp, nr, dev).copy_(dev_tensor(src, n, dev))
        torch.cuda.synchronize(dev)
        for peer, sp, ns, rp, nr in parts:
            if nr:
                self.ack[(peer, rank)].put(1)
        for peer, sp, ns, rp, nr in parts:
            if ns:
                self.ack[(rank, peer)].get(timeout=self.timeout)  # the peer has read my message: the buffer is free
