"""``Buffer``: the sampled boundary-feature exchange (reference: helper/feature_buffer.py).

Same surface -- ``init_buffer(num_in, ratio, f_send_shape, f_recv_shape, layer_size, use_pp, backend)``,
``set_selected(selected)``, ``update(layer, feat) -> [n_U, F]`` whose gradient does the reverse exchange and the
``/ratio`` scatter-add (``__grad_hook``, :169-174) -- but nothing leaves the device:

* ``backend='nccl'`` (staged): pack kernel (K3, ``bns_gather_div_f32``) -> one grouped NCCL send/recv on a side
  stream straight into the tail rows of the concat buffer -> scatter-add kernel (K5) in backward.  Replaces the
  pinned-host gloo ring of :101-129.
* ``backend='p2p'``: the pack kernel of the sender stores ``H[selected]/ratio`` directly into the receiver's concat
  buffer through a peer-mapped pointer (NVLink 5 / NVSwitch) and raises a flag there (``bns_p2p_put_rows_f32`` /
  ``bns_p2p_wait_flag``): K3 + C1 fused, no staging copy, no NCCL launch.

The exchange runs on ``self._comm_stream``; with ``update(..., overlap=True)`` the caller's stream does not wait
for it -- the aggregation op waits on ``h_u._bns_ready`` right before it touches the halo rows, so the transfer
hides behind the inner-edge SpMM.  ``Comm(s)`` is measured with CUDA events on that stream.
"""
from __future__ import annotations

import ctypes
import struct
from typing import List, Optional

import torch

from .. import ops
from .._lib import MAX_PEERS, P2P_HANDLE_BYTES, EpochMaps, PutAll, check, lib
from . import context as ctx
from .timer.timer import comm_timer


class _DevArray:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


class Buffer(object):

    def __init__(self):
        super(Buffer, self).__init__()
        self._num_in = None
        self._selected: List[Optional[torch.Tensor]] = []
        self._n_layers = 0
        self._layer_size = []
        self._ratio = []
        self._recv_shape, self._send_shape = [], []
        self._backend = None
        self._pl, self._pr = [], []
        self._comm_stream = None
        self._send_buf, self._b_recv = None, None
        self._p2p = None
        self._seq = {}
        # CUDA-graph mode (train.GraphedEpoch): no timing events; flag sequence numbers = seq_base + *seq_dev
        self.graph_mode = False
        self.seq_dev = None
        self.seq_base = 0
        self._maps = None

    # helper/feature_buffer.py:23-33
    def __init_pl_pr(self):
        self._pl, self._pr = [], []
        tot = self._num_in
        for j, s in enumerate(self._recv_shape):
            if j == self._rank:
                self._pl.append(None)
                self._pr.append(None)
            else:
                self._pl.append(tot)
                tot += s
                self._pr.append(tot)
        self._n_u = tot

    def init_buffer(self, num_in, ratio, f_send_shape, f_recv_shape, layer_size, use_pp=False, backend='nccl',
                    device=None):
        if use_pp is False:
            raise NotImplementedError            # helper/feature_buffer.py:36-37
        c = ctx.comm()
        # captured now: backward runs on autograd's device thread, where thread-local lookups would miss
        self._comm, self._timer = c, comm_timer._get()
        self._rank, self._size = c.rank, c.size
        self._num_in = num_in
        self._n_layers = len(layer_size)
        self._layer_size = layer_size
        self._recv_shape = [int(s) for s in f_recv_shape]
        self._send_shape = [int(s) for s in f_send_shape]
        self._ratio = ratio
        if backend in ('nccl', 'staged'):
            backend = 'nccl'
        elif backend != 'p2p':
            raise NotImplementedError(f"backend {backend!r}: this build moves boundary rows GPU-to-GPU "
                                      "('nccl' or 'p2p'); the reference's host-staged gloo/mpi paths are what it replaces")
        self._backend = backend
        self.__init_pl_pr()
        if self._size == 1:
            return
        self._device = torch.device(device if device is not None else torch.cuda.current_device())
        if self._device.type != 'cuda':
            raise RuntimeError("Buffer needs a CUDA device: there is no CPU exchange path")
        width = self._layer_size[1]              # the reference sizes every slab with layer_size[1] (:54-55)
        self._width = width
        self._comm_stream = torch.cuda.Stream(self._device)
        self._send_begin, tot = [], 0
        for j in range(self._size):
            self._send_begin.append(tot)
            tot += 0 if j == self._rank else self._send_shape[j]
        self._send_total = tot
        if backend == 'nccl':
            self._send_buf = [None if j == self._rank else
                              torch.zeros(self._send_shape[j], width, device=self._device) for j in range(self._size)]
            self._b_recv = [None if j == self._rank else
                            torch.zeros(self._send_shape[j], width, device=self._device) for j in range(self._size)]
        else:
            self.__init_p2p(c, width)

    # ---- p2p slabs -------------------------------------------------------------------------------
    def __init_p2p(self, c, width):
        n_comm_layers = max(self._n_layers - 1, 1)
        fwd_rows, bwd_rows = self._n_u, max(self._send_total, 1)
        row_bytes = width * 4
        if self._size - 1 > MAX_PEERS:
            raise RuntimeError(f"p2p transport: at most {MAX_PEERS + 1} partitions (BNS_MAX_PEERS)")
        self._fwd_off = [l * (fwd_rows + bwd_rows) * row_bytes for l in range(n_comm_layers)]
        self._bwd_off = [o + fwd_rows * row_bytes for o in self._fwd_off]
        # after the feature regions: the received id lists of the epoch (data_transfer NODE), int64 [sum of recv sizes]
        self._ids_off = (n_comm_layers * (fwd_rows + bwd_rows) * row_bytes + 255) // 256 * 256
        self._hop_begin, tot = [], 0
        for j in range(self._size):
            self._hop_begin.append(tot)
            tot += 0 if j == self._rank else self._recv_shape[j]
        self._recv_total = tot
        slab_bytes = self._ids_off + max(tot, 1) * 8
        self._n_comm_layers = n_comm_layers
        n_flags = (n_comm_layers * 2 + 1) * self._size          # (layer, direction, source) + (ids, source)
        h = ctypes.c_void_p()
        with torch.cuda.device(self._device):
            check(lib.bns_p2p_create(ctypes.byref(h), self._rank, self._size, slab_bytes, n_flags), "bns_p2p_create")
        self._p2p = h
        slab, flags, nbytes = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        check(lib.bns_p2p_local(h, ctypes.byref(slab), ctypes.byref(flags), ctypes.byref(nbytes)), "bns_p2p_local")
        self._slab_ptr = slab.value
        # publish: where peers must write inside MY slab (row offsets of their segment) + how to map my memory
        my = {"fwd_off": self._fwd_off, "bwd_off": self._bwd_off, "pl": self._pl, "send_begin": self._send_begin,
              "ids_off": self._ids_off, "hop_begin": self._hop_begin, "slab_bytes": nbytes.value}
        if c.kind == "thread":
            my["ptrs"] = (slab.value, flags.value)
        else:
            hb = ctypes.create_string_buffer(2 * P2P_HANDLE_BYTES)
            check(lib.bns_p2p_export(h, hb), "bns_p2p_export")
            my["handle"] = hb.raw
        import pickle
        table = [pickle.loads(b) for b in c.all_gather_bytes(pickle.dumps(my))]
        self._peer_layout = table
        for j in range(self._size):
            if j == self._rank:
                continue
            if c.kind == "thread":
                check(lib.bns_p2p_set_peer(h, j, table[j]["ptrs"][0], table[j]["ptrs"][1], table[j]["slab_bytes"]),
                      "bns_p2p_set_peer")
            else:
                with torch.cuda.device(self._device):
                    check(lib.bns_p2p_import(h, j, table[j]["handle"], table[j]["slab_bytes"]), "bns_p2p_import")
        c.barrier()
        self._peers = [j for j in range(self._size) if j != self._rank]               # ascending: segment order
        self._ring_out = [(self._rank + i) % self._size for i in range(1, self._size)]
        self._ring_in = [(self._rank - i + self._size) % self._size for i in range(1, self._size)]

    # ---- per-epoch ids and maps (p2p transport) ------------------------------------------------------
    def set_maps(self, maps: torch.Tensor, n_halo: int, pos):
        """``maps``: ONE int32 allocation ``[n_halo + (P-1) * n_in]`` = the slot map of the partition graph followed by
        the inverse map of every peer (ascending); ``pos``: train.get_pos()."""
        self._maps, self._n_halo, self._pos = maps, n_halo, pos
        self._inv = {}
        for s_, j in enumerate(self._peers):
            b = n_halo + s_ * self._num_in
            self._inv[j] = maps[b:b + self._num_in]

    def uses_p2p_ids(self) -> bool:
        return self._p2p is not None and self._maps is not None

    def exchange_ids(self, sel_cat: torch.Tensor):
        """data_transfer(selected, ..., tag=NODE) (helper/utils.py:187-213, train.py:389) over peer memory: one kernel
        stores every peer's sampled id list into that peer's slab and raises its flag, one kernel waits for the lists of
        all peers.  Returns ``(one_hops_cat, [per-peer views])`` -- views of this rank's slab."""
        cs = self._comm_stream
        main = torch.cuda.current_stream(self._device)
        n = len(self._peers)
        begin = (ctypes.c_int64 * (n + 1))()
        peers = (ctypes.c_int32 * max(n, 1))()
        roff = (ctypes.c_uint64 * max(n, 1))()
        tot = 0
        for s_, j in enumerate(self._peers):
            begin[s_] = tot
            tot += self._send_shape[j]
            peers[s_] = j
            lay = self._peer_layout[j]
            roff[s_] = lay["ids_off"] + lay["hop_begin"][self._rank] * 8
        begin[n] = tot
        flag_base = self._n_comm_layers * 2 * self._size
        seq, seq_dev = self._seq_args("ids", False)
        start = torch.cuda.Event()
        start.record(main)
        cs.wait_event(start)
        with torch.cuda.stream(cs):
            check(lib.bns_p2p_put_ids_i64(self._p2p, n, begin, peers, roff, sel_cat.data_ptr() if tot else None,
                                          flag_base + self._rank, self._size + 15, seq, seq_dev, cs.cuda_stream),
                  "bns_p2p_put_ids_i64")
            for j in self._peers:
                self._post_put_event(j, 1999, cs)
            for j in self._peers:
                self._await_put_event(j, 1999, cs)
            idx = (ctypes.c_int32 * max(n, 1))(*[flag_base + j for j in self._peers])
            check(lib.bns_p2p_wait_all(self._p2p, n, idx, seq, seq_dev, cs.cuda_stream), "bns_p2p_wait_all")
            done = torch.cuda.Event()
            done.record(cs)
        if not self.graph_mode:
            sel_cat.record_stream(cs)
        main.wait_event(done)
        # a private copy: the slab region is rewritten by the peers at the start of THEIR next epoch, and the lists stay
        # visible to the caller (train.TrainState.one_hops) after this epoch has ended
        cat = torch.as_tensor(_DevArray(self._slab_ptr + self._ids_off, (max(self._recv_total, 1),), "<i8"),
                              device=self._device)[:self._recv_total].clone()
        views = [None] * self._size
        for j in self._peers:
            views[j] = cat[self._hop_begin[j]:self._hop_begin[j] + self._recv_shape[j]]
        return cat, views

    def update_maps(self, sel_cat: torch.Tensor, hops_cat: torch.Tensor, slot: torch.Tensor):
        """construct_graph (train.py:256-281) for all peers + the inverse maps of the gradient scatter: one memset, one
        kernel (``bns_epoch_maps_update``)."""
        m = EpochMaps()
        n = len(self._peers)
        m.n_seg = n
        a = b = 0
        for s_, j in enumerate(self._peers):
            m.sel_begin[s_], m.hop_begin[s_] = a, b
            a += self._send_shape[j]
            b += self._recv_shape[j]
            m.pos[s_] = self._pos[j].data_ptr()
            m.inv[s_] = self._inv[j].data_ptr()
        m.sel_begin[n], m.hop_begin[n] = a, b
        m.selected_cat = sel_cat.data_ptr() if a else None
        m.one_hops_cat = hops_cat.data_ptr() if b else None
        m.slot = slot.data_ptr()
        m.n_in = self._num_in
        with torch.cuda.device(self._device):
            check(lib.bns_epoch_maps_update(ctypes.byref(m), self._maps.data_ptr(), self._maps.numel() * 4,
                                            torch.cuda.current_stream(self._device).cuda_stream), "bns_epoch_maps_update")

    # Ranks that are THREADS of one process share a CUDA context, where a kernel spinning on a flag can block the
    # very launch that would set it (lazy module loading and stream->hardware-queue aliasing both synchronise the
    # context).  There, the producer hands over an event recorded after its put and the consumer's stream waits on
    # it BEFORE the flag-wait kernel is launched, so that kernel finds the flag already set and never spins.  With
    # one process per GPU (the deployment shape) these two calls do nothing and the flag is the only signal.
    def _post_put_event(self, peer, tag, stream):
        if self._comm.kind == "thread":
            ev = torch.cuda.Event()
            ev.record(stream)
            self._comm.post_event(peer, tag, ev)

    def _await_put_event(self, peer, tag, stream):
        if self._comm.kind == "thread":
            stream.wait_event(self._comm.take_event(peer, tag))

    def _timer_ctx(self, name, stream):
        import contextlib
        return contextlib.nullcontext() if self.graph_mode else self._timer.timer(name, stream=stream)

    def _seq_args(self, layer, backward):
        """(immediate, device pointer) of this exchange's flag value."""
        if self.graph_mode:
            return self.seq_base, self.seq_dev.data_ptr()
        key = (layer, 1 if backward else 0)
        self._seq[key] = self._seq.get(key, 0) + 1
        return self._seq[key], None

    def _flag(self, layer, backward, src):
        return ((layer - 1) * 2 + (1 if backward else 0)) * self._size + src

    def _slab_view(self, byte_off, rows):
        return torch.as_tensor(_DevArray(self._slab_ptr + byte_off, (rows, self._width)), device=self._device)

    def input_slot(self, layer, rows, width):
        """The rows ``[0, n_in)`` of layer ``layer``'s concat buffer when they can be written in place (peer-mapped
        transport, full slab width), else None.  ``update(layer, feat)`` recognises a ``feat`` that already lives there
        and skips its copy (K4)."""
        if self._size == 1 or self._p2p is None or rows != self._num_in or width != getattr(self, "_width", -1):
            return None
        if not (1 <= layer <= self._n_comm_layers):
            return None
        return self._slab_view(self._fwd_off[layer - 1], self._num_in)

    def set_selected(self, selected, selected_cat=None):
        """``selected_cat``: the same lists concatenated in ascending peer order (what the sampler produced); built here
        when the caller injects per-peer lists."""
        self._selected = selected
        if selected_cat is None and self._p2p is not None:
            parts = [selected[j] for j in self._peers if selected[j] is not None and selected[j].numel()]
            selected_cat = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=self._device)
        self._selected_cat = selected_cat

    # ---- forward ---------------------------------------------------------------------------------
    def update(self, layer, feat, overlap=False):
        """``[feat ; recv_0 ; recv_1 ...]`` with the boundary rows of the peers (helper/feature_buffer.py:93-99)."""
        if self._size == 1:
            return feat
        res = _BoundaryExchange.apply(feat, self, layer, overlap)
        if overlap:
            res._bns_ready = self._last_ready
        res._bns_exchange = (self, layer)          # lets a fused consumer start the gradient return trip early
        return res

    def _forward(self, layer, feat, overlap):
        F = feat.shape[1]
        if F > self._width:
            raise RuntimeError(f"layer width {F} > slab width {self._width} (the reference sizes slabs with layer_size[1])")
        feat = feat.contiguous()
        main, cs = torch.cuda.current_stream(self._device), self._comm_stream
        ready = torch.cuda.Event()
        if self._backend == 'nccl':
            h_u = torch.empty(self._n_u, F, device=self._device)
        else:
            h_u = self._slab_view(self._fwd_off[layer - 1], self._n_u)[:, :F] if F == self._width else None
            if h_u is None:
                raise RuntimeError("p2p transport needs equal hidden widths")
        if feat.data_ptr() != h_u.data_ptr():                          # (written in place by the producer: input_slot)
            ops.copy_rows(feat, h_u, self._num_in)                     # K4: the only copy of the concat
        start = torch.cuda.Event()
        start.record(main)
        cs.wait_event(start)
        with torch.cuda.stream(cs):
            with self._timer_ctx(f'forward_{layer}', cs):
                if self._backend == 'nccl':
                    send = [None] * self._size
                    recv = [None] * self._size
                    for j in range(self._size):
                        if j == self._rank:
                            continue
                        send[j] = self._send_buf[j][:, :F] if F == self._width else \
                            torch.empty(self._send_shape[j], F, device=self._device)
                        ops.gather_div(feat, self._selected[j], self._ratio[j], out=send[j])      # K3
                        recv[j] = h_u[self._pl[j]:self._pr[j]]
                    self._comm.alltoall(send, recv, tag=16 + layer)                               # C1/C2
                else:
                    seq, seq_dev = self._seq_args(layer, False)
                    segs = PutAll()
                    segs.n_seg = len(self._peers)
                    tot = 0
                    for s_, j in enumerate(self._peers):
                        lay = self._peer_layout[j]
                        segs.row_begin[s_] = tot
                        tot += self._send_shape[j]
                        segs.peer[s_] = j
                        segs.remote_off[s_] = lay["fwd_off"][layer - 1] + lay["pl"][self._rank] * self._width * 4
                        segs.div[s_] = float(self._ratio[j]) if self._send_shape[j] else 1.0
                    segs.row_begin[segs.n_seg] = tot
                    sel_cat = self._selected_cat
                    check(lib.bns_p2p_put_all_f32(self._p2p, ctypes.byref(segs), self._width, feat.data_ptr(), feat.stride(0),
                                                  F, sel_cat.data_ptr() if tot else None,
                                                  self._flag(layer, False, self._rank), self._size + (layer - 1) * 2, seq,
                                                  seq_dev, cs.cuda_stream), "bns_p2p_put_all_f32")
                    for j in self._peers:
                        self._post_put_event(j, 2000 + 2 * layer, cs)
                    for j in self._peers:
                        self._await_put_event(j, 2000 + 2 * layer, cs)
                    idx = (ctypes.c_int32 * len(self._peers))(*[self._flag(layer, False, j) for j in self._peers])
                    check(lib.bns_p2p_wait_all(self._p2p, len(self._peers), idx, seq, seq_dev, cs.cuda_stream),
                          "bns_p2p_wait_all")
            ready.record(cs)
        if not self.graph_mode:
            feat.record_stream(cs)
            h_u.record_stream(cs)
        if not overlap:
            main.wait_event(ready)
        self._last_ready = ready
        return h_u

    # ---- backward (the grad hook) -------------------------------------------------------------------
    def begin_backward(self, layer, grad):
        """Start the gradient return trip of ``layer`` as soon as the HALO rows ``grad[n_in:]`` are final: the rows go to
        their owners on the comm stream while the caller still computes ``grad[:n_in]`` (fused.SageConvFn.backward calls
        this between its two transposed aggregations).  ``_backward`` then only waits and scatter-adds."""
        if self._size == 1:
            return
        self._begun = (layer, grad.data_ptr(), self._exchange_backward(layer, grad))

    def _exchange_backward(self, layer, grad):
        """Enqueue send + receive of the halo gradient rows; returns ``(done event, recv list or None)``."""
        F = grad.shape[1]
        main, cs = torch.cuda.current_stream(self._device), self._comm_stream
        start, done = torch.cuda.Event(), torch.cuda.Event()
        start.record(main)
        cs.wait_event(start)
        recv = None
        with torch.cuda.stream(cs):
            with self._timer_ctx(f'backward_{layer}', cs):
                if self._backend == 'nccl':
                    send = [None if j == self._rank else grad[self._pl[j]:self._pr[j]] for j in range(self._size)]
                    recv = [None if j == self._rank else
                            (self._b_recv[j][:, :F] if F == self._width else
                             torch.empty(self._send_shape[j], F, device=self._device)) for j in range(self._size)]
                    self._comm.alltoall(send, recv, tag=64 + layer)
                else:
                    seq, seq_dev = self._seq_args(layer, True)
                    segs = PutAll()
                    segs.n_seg = len(self._peers)
                    tot = 0
                    for s_, j in enumerate(self._peers):
                        lay = self._peer_layout[j]
                        segs.row_begin[s_] = tot
                        tot += self._recv_shape[j]
                        segs.peer[s_] = j
                        segs.remote_off[s_] = lay["bwd_off"][layer - 1] + lay["send_begin"][self._rank] * self._width * 4
                        segs.src_begin[s_] = self._pl[j]
                        segs.div[s_] = 1.0
                    segs.row_begin[segs.n_seg] = tot
                    check(lib.bns_p2p_put_all_f32(self._p2p, ctypes.byref(segs), self._width, grad.data_ptr(), grad.stride(0),
                                                  F, None, self._flag(layer, True, self._rank),
                                                  self._size + (layer - 1) * 2 + 1, seq, seq_dev, cs.cuda_stream),
                          "bns_p2p_put_all_f32")
                    for j in self._peers:
                        self._post_put_event(j, 2001 + 2 * layer, cs)
                    for j in self._peers:
                        self._await_put_event(j, 2001 + 2 * layer, cs)
                    idx = (ctypes.c_int32 * len(self._peers))(*[self._flag(layer, True, j) for j in self._peers])
                    check(lib.bns_p2p_wait_all(self._p2p, len(self._peers), idx, seq, seq_dev, cs.cuda_stream),
                          "bns_p2p_wait_all")
                    recv = [None] * self._size
                    bwd = self._slab_view(self._bwd_off[layer - 1], max(self._send_total, 1))
                    for j in self._peers:
                        recv[j] = bwd[self._send_begin[j]:self._send_begin[j] + self._send_shape[j], :F]
            done.record(cs)
        if not self.graph_mode:
            grad.record_stream(cs)
        return done, recv

    def _backward(self, layer, grad):
        F = grad.shape[1]
        if not grad.is_contiguous():
            grad = grad.contiguous()
        trace = getattr(self, "trace", None)          # tests only: {name: tensor} of the gradients around the exchange
        if trace is not None:
            trace[f"grad_u{layer}"] = grad.detach().clone()
        main = torch.cuda.current_stream(self._device)
        begun, self._begun = getattr(self, "_begun", None), None
        if begun is not None and begun[0] == layer and begun[1] == grad.data_ptr():
            done, recv = begun[2]                     # the producer already sent the halo rows (begin_backward)
        else:
            done, recv = self._exchange_backward(layer, grad)
        main.wait_event(done)
        inner = grad[:self._num_in]
        if self._backend == 'nccl' or self._maps is None:
            for i in range(1, self._size):           # the reference's order: idx = left, i = 1 .. P-1 (:111-129)
                left = (self._rank - i + self._size) % self._size
                ops.scatter_add_div(inner, self._selected[left], recv[left], self._ratio[left])      # K5
        else:
            # the same P-1 scatter-adds, in the same order, as ONE race-free launch over the inverse maps
            order = [j for j in self._ring_in if self._send_shape[j] > 0]
            if order:
                n = len(order)
                inv = (ctypes.c_void_p * n)(*[self._inv[j].data_ptr() for j in order])
                base = self._slab_ptr + self._bwd_off[layer - 1]
                rcv = (ctypes.c_void_p * n)(*[base + self._send_begin[j] * self._width * 4 for j in order])
                div = (ctypes.c_float * n)(*[float(self._ratio[j]) for j in order])
                with torch.cuda.device(self._device):
                    check(lib.bns_scatter_rows_all_f32(inner.data_ptr(), inner.stride(0), self._num_in, F, n, inv, rcv,
                                                       self._width, div, main.cuda_stream), "bns_scatter_rows_all_f32")
        if trace is not None:
            trace[f"grad_h{layer}"] = inner.detach().clone()
        return inner

    def __del__(self):
        h, self._p2p = getattr(self, "_p2p", None), None
        if h:
            try:
                lib.bns_p2p_destroy(h)
            except Exception:
                pass


class _BoundaryExchange(torch.autograd.Function):

    @staticmethod
    def forward(ctx_, feat, buf: Buffer, layer: int, overlap: bool):
        ctx_.buf, ctx_.layer = buf, layer
        return buf._forward(layer, feat, overlap)

    @staticmethod
    def backward(ctx_, grad):
        return ctx_.buf._backward(ctx_.layer, grad), None, None, None
