"""Peer-memory exchange of a gallery-sharded evaluation (csrc/peer.cu).

The ranks of one box write their thresholds, rank counts and top-k candidates straight into each other's exchange buffers with
NVLink stores from the library's own kernels and synchronise through flags in those buffers: one kernel per exchange instead of
an NCCL all-reduce and an all-gather.  torch.distributed is used once, to hand the buffers' CUDA IPC handles round and to agree
that every rank could open all of them; sharded.py falls back to the collectives when that fails (or when the ranks are not
all CUDA ranks of one node).
"""
import ctypes
import os
import socket

import torch
import torch.distributed as dist

from . import _lib

_cache = {}        # (group id, device index, Q, k) -> PeerExchange or None (None: set-up failed, use the collectives)


class PeerExchange(object):
    """The exchange buffers of one (process group, query count, top-k width): allocate -> IPC handles round -> open.

    Collective: every rank of the group must construct it at the same point, and afterwards call thresholds() / results()
    in the same order (the sequence number of an exchange is its position in that order)."""

    def __init__(self, n_queries, k, group=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.q, self.k = int(n_queries), int(k)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seq = 0
        self.own = None
        self.opened = []
        lib = _lib.load()
        nbytes = lib.witw_peer_exchange_bytes(self.q, self.k, self.world)
        ok, handle = True, bytes(64)
        why = ""
        with torch.cuda.device(self.device):
            if nbytes == 0:
                ok, why = False, "unsupported size (Q %d, k %d, world %d)" % (self.q, self.k, self.world)
            else:
                buf, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                if lib.witw_peer_alloc(nbytes, ctypes.byref(buf), h) == 0:
                    self.own, handle = buf.value, h.raw
                else:
                    ok, why = False, _lib.last_error()
            # everybody learns everybody's handle, host and success in one object all-gather
            info = [None] * self.world
            dist.all_gather_object(info, (ok, handle, socket.gethostname(), os.getpid()), group=group)
            ok = all(i[0] for i in info) and len(set(i[2] for i in info)) == 1
            ptrs = []
            if ok:
                for r, (_, h, _, _) in enumerate(info):
                    if r == self.rank:
                        ptrs.append(self.own)
                        continue
                    p = ctypes.c_void_p()
                    if lib.witw_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                        ok, why = False, _lib.last_error()
                        break
                    self.opened.append(p.value)
                    ptrs.append(p.value)
            on_dev = dist.get_backend(group) == "nccl"
            agree = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device if on_dev else "cpu")
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)      # also the barrier between the memsets and the first stores
            torch.cuda.synchronize()
            self.ok = bool(int(agree.item()))
            if not self.ok:
                self.why = why or "another rank could not set its buffers up"
                self.close()
                return
            self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def close(self):
        lib = _lib.load()
        for p in self.opened:
            lib.witw_peer_close(ctypes.c_void_p(p))
        self.opened = []
        if self.own:
            lib.witw_peer_free(ctypes.c_void_p(self.own))
            self.own = None

    def thresholds(self, d_local, true_idx, g_offset, g_local):
        """d_local [Q] fp32: query q's distance to item true_idx[q] where this rank owns it -> the complete d_true [Q]."""
        self.seq += 1
        out = torch.empty(self.q, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("witw_peer_thresholds", d_local.contiguous().data_ptr(), true_idx.contiguous().data_ptr(), int(g_offset), int(g_local),
                      self.q, self.k, self.ptrs.data_ptr(), self.world, self.rank, self.seq, out.data_ptr(), _stream())
        return out

    def results(self, counts, td, ti, flagged):
        """counts [Q] int32, td / ti [Q,k] or None, flagged [1] int32 or None -> (summed counts int64 [Q], merged top-k
        distances, merged top-k indices, summed flags [1] int32 on the device)."""
        self.seq += 1
        total = torch.empty(self.q, dtype=torch.int64, device=self.device)
        n_flag = torch.empty(1, dtype=torch.int32, device=self.device)
        md = mi = None
        if self.k:
            md = torch.empty((self.q, self.k), dtype=torch.float32, device=self.device)
            mi = torch.empty((self.q, self.k), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("witw_peer_results", counts.contiguous().data_ptr(), 0 if td is None else td.contiguous().data_ptr(),
                      0 if ti is None else ti.contiguous().data_ptr(), 0 if flagged is None else flagged.data_ptr(), self.q, self.k,
                      self.ptrs.data_ptr(), self.own, self.world, self.rank, self.seq, total.data_ptr(), n_flag.data_ptr(),
                      0 if md is None else md.data_ptr(), 0 if mi is None else mi.data_ptr(), _stream())
        return total, md, mi, n_flag


def _stream():
    return torch.cuda.current_stream().cuda_stream


def get(n_queries, k, group=None, device=None):
    """The cached PeerExchange of (group, device, Q, k), or None when the peer path is not available.  Collective on first use."""
    if not (dist.is_available() and dist.is_initialized()) or not torch.cuda.is_available():
        return None
    if dist.get_world_size(group) < 2 or dist.get_world_size(group) > 16:
        return None
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (id(group) if group is not None else 0, dev.index, int(n_queries), int(k))
    if key not in _cache:
        px = PeerExchange(n_queries, k, group=group, device=dev)
        _cache[key] = px if px.ok else None
        if not px.ok and dist.get_rank(group) == 0:
            import warnings
            warnings.warn("witw_b200: peer-memory exchange unavailable (%s); using NCCL collectives" % px.why)
    return _cache[key]


def shutdown():
    """Close every exchange buffer (before destroying the process group).  Every rank must have finished its evaluations: a peer
    that still waits on this rank's flags would time out."""
    if torch.cuda.is_available() and any(px is not None for px in _cache.values()):
        torch.cuda.synchronize()
    for px in _cache.values():
        if px is not None:
            px.close()
    _cache.clear()
