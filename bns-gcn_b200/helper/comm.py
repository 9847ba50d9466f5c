"""Rank-to-rank plumbing behind the reference's ``torch.distributed`` call sites.

The reference calls ``dist.get_rank()/isend/recv/all_reduce`` directly (helper/utils.py:150-213,
helper/feature_buffer.py:101-153, helper/reducer.py:28-49).  Here the same operations go through one
small object so that a rank can be either

* ``DistComm``   one process per GPU over ``torch.distributed`` (NCCL on CUDA tensors; gloo when the
                 host-side setup logic is exercised on CPU tensors), or
* ``ThreadComm`` one of P ranks living as threads of ONE process on ONE GPU (tests, smoke, single-GPU
                 emulation of a P-partition run): messages are device tensors handed over with a CUDA event, or
* ``SoloComm``   a world of one.

Only data movement lives here; every arithmetic kernel is in libbnsgcn.so.
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, List, Optional, Sequence

import torch


class SoloComm:
    rank, size = 0, 1
    kind = "solo"

    def alltoall(self, send: Sequence[Optional[torch.Tensor]], recv: Sequence[Optional[torch.Tensor]], tag: int = 0):
        return None

    def all_reduce_sum(self, t: torch.Tensor):
        return None

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        return [b]

    def barrier(self):
        return None


class ThreadFabric:
    def __init__(self, size: int):
        self.size = size
        self._box: Dict[tuple, "queue.Queue"] = {}
        self._lock = threading.Lock()
        self._bar = threading.Barrier(size)
        self._slots: List = [None] * size
        self.shared: Dict = {}
        self.failed = False          # set when any rank raised: blocked peers give up instead of hanging

    def box(self, src: int, dst: int, tag: int) -> "queue.Queue":
        with self._lock:
            return self._box.setdefault((src, dst, tag), queue.Queue())

    def comm(self, rank: int) -> "ThreadComm":
        return ThreadComm(self, rank)


class ThreadComm:
    kind = "thread"

    def __init__(self, fabric: ThreadFabric, rank: int):
        self.fabric, self.rank, self.size = fabric, rank, fabric.size

    def _put(self, t: torch.Tensor, dst: int, tag: int):
        ev = None
        t = t.detach().clone()      # the sender may reuse its buffer as soon as this returns (NCCL semantics)
        if t.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.device))
        self.fabric.box(self.rank, dst, tag).put((t, ev))

    def _get(self, out: torch.Tensor, src: int, tag: int):
        box, waited = self.fabric.box(src, self.rank, tag), 0.0
        while True:
            try:
                t, ev = box.get(timeout=0.5)
                break
            except queue.Empty:
                waited += 0.5
                if self.fabric.failed or waited > 120:
                    raise RuntimeError(f"rank {self.rank}: no message from rank {src} (tag {tag}); "
                                       f"{'a peer failed' if self.fabric.failed else 'timeout'}")
        if ev is not None:
            torch.cuda.current_stream(out.device).wait_event(ev)
        out.copy_(t.view_as(out))
        if out.is_cuda:                       # keep the sender's tensor alive until the copy ran
            t.record_stream(torch.cuda.current_stream(out.device))

    def post_event(self, dst: int, tag: int, ev):
        """Hand a recorded CUDA event to rank ``dst`` (in-process ranks only; see Buffer p2p transport)."""
        self.fabric.box(self.rank, dst, tag).put((None, ev))

    def take_event(self, src: int, tag: int):
        box, waited = self.fabric.box(src, self.rank, tag), 0.0
        while True:
            try:
                return box.get(timeout=0.5)[1]
            except queue.Empty:
                waited += 0.5
                if self.fabric.failed or waited > 120:
                    raise RuntimeError(f"rank {self.rank}: no event from rank {src} (tag {tag})")

    def alltoall(self, send, recv, tag: int = 0):
        """``recv[j] <- send_j_on_rank_j[self.rank]`` for all peers j (entries for self / None are skipped)."""
        for i in range(1, self.size):
            right = (self.rank + i) % self.size
            if send[right] is not None:
                self._put(send[right], right, tag)
        for i in range(1, self.size):
            left = (self.rank - i + self.size) % self.size
            if recv[left] is not None:
                self._get(recv[left], left, tag)

    def all_reduce_sum(self, t: torch.Tensor):
        f = self.fabric
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        f._slots[self.rank] = t
        f._bar.wait()
        tot = f._slots[0].clone()
        for r in range(1, self.size):        # fixed order: all ranks obtain identical bits
            tot += f._slots[r]
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        f._bar.wait()
        t.copy_(tot)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        f._bar.wait()

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        f = self.fabric
        f._slots[self.rank] = b
        f._bar.wait()
        out = list(f._slots)
        f._bar.wait()
        return out

    def barrier(self):
        self.fabric._bar.wait()


class DistComm:
    """``torch.distributed`` endpoints (one process per GPU).  ``alltoall`` is a single
    ``all_to_all_single``-style grouped send/recv on NCCL, or the reference's ring of isend/irecv on gloo
    (gloo has no all-to-all; helper/feature_buffer.py:111-121 hand-rolls the same ring)."""
    kind = "dist"

    def __init__(self):
        import os
        import torch.distributed as dist
        self._d = dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.backend = dist.get_backend()
        # BNS_COMM=abi: the data-path collectives go through libbnsgcn.so's own communicator (bns_ctx_create /
        # bns_allreduce_sum_f32 / bns_alltoallv_bytes: what a non-torch host would bind);
        # torch.distributed then only bootstraps (hands out the unique id) and serves the setup-time object exchanges
        self._ctx = None
        if self.backend == "nccl" and self.size > 1 and os.environ.get("BNS_COMM", "torch") == "abi":
            self._init_abi()

    def _init_abi(self):
        import ctypes
        from .._lib import COMM_ID_BYTES, check, lib
        ident = ctypes.create_string_buffer(COMM_ID_BYTES)
        if self.rank == 0:
            check(lib.bns_comm_unique_id(ident), "bns_comm_unique_id")
        box = [ident.raw]
        self._d.broadcast_object_list(box, src=0)
        ctx_ = ctypes.c_void_p()
        check(lib.bns_ctx_create(ctypes.byref(ctx_), self.rank, self.size, box[0]), "bns_ctx_create")
        self._ctx = ctx_

    def _abi_alltoall(self, send, recv):
        import ctypes
        from .._lib import check, lib
        P = self.size
        sp, sb = (ctypes.c_void_p * P)(), (ctypes.c_int64 * P)()
        rp, rb = (ctypes.c_void_p * P)(), (ctypes.c_int64 * P)()
        keep = []
        for j in range(P):
            if j == self.rank:
                continue
            if send[j] is not None and send[j].numel():
                t = send[j].contiguous()
                keep.append(t)
                sp[j], sb[j] = t.data_ptr(), t.numel() * t.element_size()
            if recv[j] is not None and recv[j].numel():
                if not recv[j].is_contiguous():
                    raise RuntimeError("alltoall: receive buffers must be contiguous")
                rp[j], rb[j] = recv[j].data_ptr(), recv[j].numel() * recv[j].element_size()
        dev = next(t.device for t in list(send) + list(recv) if t is not None)
        with torch.cuda.device(dev):
            check(lib.bns_alltoallv_bytes(self._ctx, sp, sb, rp, rb, torch.cuda.current_stream(dev).cuda_stream),
                  "bns_alltoallv_bytes")

    def alltoall(self, send, recv, tag: int = 0):
        d = self._d
        if self._ctx is not None and any(t is not None and t.is_cuda for t in list(send) + list(recv)):
            return self._abi_alltoall(send, recv)
        if self.backend == "nccl":
            ops = []
            for i in range(1, self.size):
                right, left = (self.rank + i) % self.size, (self.rank - i + self.size) % self.size
                if send[right] is not None and send[right].numel():
                    ops.append(d.P2POp(d.isend, send[right], right))
                if recv[left] is not None and recv[left].numel():
                    ops.append(d.P2POp(d.irecv, recv[left], left))
            if ops:
                for w in d.batch_isend_irecv(ops):     # one ncclGroupStart/End on the current stream
                    w.wait()
            return
        reqs = []
        for i in range(1, self.size):
            right, left = (self.rank + i) % self.size, (self.rank - i + self.size) % self.size
            if recv[left] is not None:
                reqs.append(d.irecv(recv[left], src=left, tag=tag))
            if send[right] is not None:
                reqs.append(d.isend(send[right].contiguous(), dst=right, tag=tag))
        for r in reqs:
            r.wait()

    def all_reduce_sum(self, t: torch.Tensor):
        if self._ctx is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
            from .._lib import check, lib
            with torch.cuda.device(t.device):
                check(lib.bns_allreduce_sum_f32(self._ctx, t.data_ptr(), t.numel(),
                                                torch.cuda.current_stream(t.device).cuda_stream), "bns_allreduce_sum_f32")
            return
        self._d.all_reduce(t, op=self._d.ReduceOp.SUM)

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        out = [None] * self.size
        self._d.all_gather_object(out, b)
        return out

    def barrier(self):
        self._d.barrier()


def run_threads(n_ranks: int, fn, *args, device: Optional[str] = None):
    """Run ``fn(comm, rank, *args)`` as ``n_ranks`` in-process ranks.  On CUDA every rank gets its own main
    stream (sharing the legacy default stream would serialise the ranks and can deadlock flag waits)."""
    from . import context as ctx
    if n_ranks == 1:
        ctx.set_comm(SoloComm())
        try:
            return [fn(ctx.comm(), 0, *args)]
        finally:
            ctx.reset()
    fabric = ThreadFabric(n_ranks)
    out: List = [None] * n_ranks
    err: List = [None] * n_ranks

    def work(r):
        try:
            c = fabric.comm(r)
            ctx.set_comm(c)
            # autograd normally runs every CUDA backward of the process on ONE device thread; ranks that wait
            # for each other's messages inside backward would starve each other there -> run it on this thread
            with torch.autograd.set_multithreading_enabled(False):
                if device is not None and str(device).startswith("cuda"):
                    torch.cuda.set_device(device)
                    with torch.cuda.stream(torch.cuda.Stream(device)):
                        out[r] = fn(c, r, *args)
                        torch.cuda.current_stream().synchronize()
                else:
                    out[r] = fn(c, r, *args)
        except BaseException as e:      # noqa: BLE001
            err[r] = e
            fabric.failed = True
            fabric._bar.abort()
        finally:
            ctx.reset()

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(n_ranks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    real = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    for e in err:
        if e is not None:
            raise e
    return out
