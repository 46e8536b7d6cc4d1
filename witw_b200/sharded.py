"""Gallery sharding across the GPUs of one box.

Gallery items are independent, so rank p of P owns a contiguous slice of the gallery and
every rank holds all queries.  One sweep needs three tiny exchanges over NCCL / NVLink:

  1. true-match distances: the owner of gallery item true_idx[q] computes d_true[q] in fp32;
     all-reduce(sum) of a [Q] vector that is zero everywhere else (a NaN distance -- zero-norm
     features -- stays NaN through the sum, as it must)
  2. rank counts: all-reduce(sum) of the local #{g : d[g,q] <= d_true[q]}  -> the exact
     reference ranks (cvig_fov.py:552) for the whole gallery
  3. top-k: all-gather of each shard's [Q,k] (distance, global index) candidates, then a k-way merge

The local compute is pluggable so the exchange logic is testable on CPU (gloo) with the
oracle standing in for the kernels; the default is the CUDA path of ops.py.
"""
import torch
import torch.distributed as dist

from . import _lib, ops


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced partition: the first n_items % world_size ranks own one extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CudaLocal(object):
    """Local shard compute on the B200 kernels.

    event_sink: optional list; every tensor-core sweep appends the (start, end) torch.cuda.Event pair recorded around
    the sweep kernel alone (bench.py's roofline timer)."""

    def __init__(self, path="auto", event_sink=None):
        self.path = path
        self.event_sink = event_sink
        self._key = None
        self._prepared = None
        self._dmat = None

    def _operands(self, ov_local, su):
        """Prepared (GalleryIndex, QueryBatch) of this shard, shared by true_distances() and sweep() of one evaluation."""
        key = (ov_local.data_ptr(), tuple(ov_local.shape), su.data_ptr(), tuple(su.shape), ov_local._version, su._version)
        if self._key != key:
            self._prepared = (ops.GalleryIndex(ov_local, su.shape[3]), ops.QueryBatch(su))
            self._key = key
        return self._prepared

    def _tc(self, ov_local, su):
        g, q, ch, w, sw = ops._feature_dims("sweep", ov_local, su)
        return ops._pick_path(self.path, g, q, ch, w, sw) == "tc"

    def true_distances(self, ov_local, su, local_idx):
        """Exact fp32 distance of query i to local gallery item local_idx[i]."""
        if self._tc(ov_local, su):
            gallery, queries = self._operands(ov_local, su)
            pq = torch.arange(su.shape[0], dtype=torch.int64, device=su.device)
            return ops.pair_distances_prepared(gallery, queries, local_idx.to(torch.int64).contiguous(), pq)[0]
        g = ov_local.shape[0]
        if int(local_idx.numel()) and (int(local_idx.max()) >= g or int(local_idx.min()) < 0):
            raise IndexError("true_distances: index outside the shard")
        # fp32 path: the threshold is read out of the same distance matrix the sweep will count over, so that the match
        # compares equal to itself bit for bit (cvig_fov.py:552 takes both from one tensor); sweep() reuses the matrix
        _, dmat = ops.match(ov_local, su, path="fp32")
        self._dmat = (self._dmat_key(ov_local, su), dmat)
        return dmat[local_idx.to(torch.int64), torch.arange(su.shape[0], device=su.device)]

    @staticmethod
    def _dmat_key(ov_local, su):
        return (ov_local.data_ptr(), tuple(ov_local.shape), su.data_ptr(), tuple(su.shape), ov_local._version, su._version)

    def sweep(self, ov_local, su, d_true, true_idx, g_offset, topk):
        if self._tc(ov_local, su):
            # exact finish inside the shard: fp32 decisions wherever fp16 cannot settle a rank decision and fp32 re-ranking of
            # the shard's top-k candidates, so what is exchanged are already the reference's counts and distances
            gallery, queries = self._operands(ov_local, su)
            gallery.g_offset = int(g_offset)
            self._key = None   # one evaluation per preparation: the caller may overwrite the buffers in place
            ev = None
            if self.event_sink is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                self.event_sink.append(ev)
            # a true match outside this shard gets a local index outside [0, G): it is never counted by index here
            res = ops.evaluate_ranks_prepared(gallery, queries, true_idx=true_idx - g_offset, topk=topk, d_true=d_true, events=ev)
            if topk:
                return res
            return res, None, None
        cached = self._dmat
        if cached is not None and cached[0] == self._dmat_key(ov_local, su):
            dmat = cached[1]
        else:
            _, dmat = ops.match(ov_local, su, path="fp32")
        self._dmat = None
        # the owner's d_true[q] is an element of dmat, so its match counts; NaN compares false as in the reference
        counts = (dmat <= d_true.unsqueeze(0)).sum(dim=0).to(torch.int64)
        if topk:
            td, ti = ops.topk_from_distances(dmat, topk, g_offset=g_offset)
            return counts, td, ti
        return counts, None, None

    def launch(self, ov_local, su, d_true, true_idx, g_offset, topk):
        """sweep() in two halves: everything is enqueued and a handle returned; handle.provisional() = (counts, topk_dist,
        topk_idx, flagged [1] int32 on the device), handle.finish() re-does the queries the fp32 finish flagged (in place)."""
        if not self._tc(ov_local, su):
            return _Done(*self.sweep(ov_local, su, d_true, true_idx, g_offset, topk))
        gallery, queries = self._operands(ov_local, su)
        gallery.g_offset = int(g_offset)
        self._key = None
        ev = None
        if self.event_sink is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.event_sink.append(ev)
        return _Pending(ops.RankEvaluation(gallery, queries, true_idx=true_idx - g_offset, topk=topk, d_true=d_true, events=ev))

    def merge(self, cand_d, cand_i, topk):
        p, q, k = cand_d.shape
        out_d = torch.empty((q, k), dtype=torch.float32, device=cand_d.device)
        out_i = torch.empty((q, k), dtype=torch.int32, device=cand_d.device)
        _lib.call("witw_topk_merge", cand_d.contiguous().data_ptr(), cand_i.contiguous().data_ptr(), p, q, k,
                  out_d.data_ptr(), out_i.data_ptr(), ops._stream())
        return out_d, out_i


class _Done(object):
    """Handle of a local evaluation that is already final."""

    def __init__(self, counts, td, ti):
        self.counts, self.td, self.ti = counts, td, ti

    def provisional(self):
        return self.counts, self.td, self.ti, None

    def finish(self):
        return False


class _Pending(object):
    """Handle of an enqueued ops.RankEvaluation."""

    def __init__(self, evaluation):
        self.ev = evaluation

    def provisional(self):
        nq = self.ev.queries.Q
        flagged = self.ev.defer.n_flagged if self.ev.defer is not None else None
        return self.ev.counts[:nq], self.ev.td, self.ev.ti, flagged

    def finish(self):
        self.ev.result()            # re-does the flagged queries in place (counts / td / ti are views of the evaluation's buffers)
        return self.ev.flagged > 0


# The exchange over peer memory (witw_b200/peer.py, csrc/peer.cu): NVLink stores from the library's own kernels + flags instead of
# NCCL collectives.  Used when the group's ranks are CUDA ranks of one node and every rank could open every buffer; the NCCL
# path below is the fallback (and the gloo path of the CPU tests).
PEER_EXCHANGE = True

# One all-gather of [counts | top-k distances | top-k indices] per rank instead of an all-reduce and two all-gathers: two
# collective launches fewer per sweep (bench.py's `exchange` key times both forms over NCCL).  packed=False keeps the
# separate collectives.
PACKED_EXCHANGE = True


def _exchange_packed(counts, td, ti, world, group, flagged=None):
    """counts [Q], td fp32 [Q,k], ti int32 [Q,k] of this shard (td / ti None without a top-k), flagged [1] int32 or None
    -> (summed counts int64 [Q], [P,Q,k] distances, [P,Q,k] indices, total flagged as a 0-dim tensor)."""
    q = counts.shape[0]
    k = 0 if td is None else td.shape[1]
    mine = torch.zeros((q + 1, 1 + 2 * k), dtype=torch.int32, device=counts.device)
    mine[:q, 0] = counts
    if flagged is not None:
        mine[q, 0] = flagged[0]
    if k:
        mine[:q, 1:1 + k] = td.contiguous().view(torch.int32)
        mine[:q, 1 + k:] = ti
    allp = torch.empty((world,) + tuple(mine.shape), dtype=torch.int32, device=counts.device)
    dist.all_gather(list(allp.unbind(0)), mine, group=group)
    total = allp[:, :q, 0].to(torch.int64).sum(dim=0)
    n_flag = allp[:, q, 0].sum()
    if not k:
        return total, None, None, n_flag
    return total, allp[:, :q, 1:1 + k].contiguous().view(torch.float32), allp[:, :q, 1 + k:].contiguous(), n_flag


class ShardedEvaluation(object):
    """evaluate_ranks_sharded in two halves (cf. ops.RankEvaluation): the constructor enqueues the thresholds' all-reduce,
    the local sweep with its fp32 finish and the exchange of the results; result() learns from the exchanged data whether any
    rank's finish flagged a query -- in that (rare) case every rank re-does its flagged queries in fp32 and the exchange is
    repeated -- and returns ranks [, topk_dist, topk_idx], identical on every rank.  All ranks must construct and finish their
    evaluations in the same order (the collectives are matched by order)."""

    def __init__(self, ov_local, surface_embed, g_offset, n_gallery_total, true_idx=None, topk=0, group=None, local=None, packed=None):
        self.local = local or CudaLocal()
        self.group, self.topk = group, int(topk)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.packed = PACKED_EXCHANGE if packed is None else packed
        dev = surface_embed.device
        q = surface_embed.shape[0]
        g_local = ov_local.shape[0]
        self.peer = None
        if self.world > 1 and PEER_EXCHANGE and packed is None and surface_embed.is_cuda and q > 0:
            from . import peer
            self.peer = peer.get(q, self.topk, group=group, device=dev)
        if true_idx is None:
            if q > n_gallery_total:
                raise IndexError("evaluate_ranks_sharded: %d queries but only %d gallery items and no true_idx" % (q, n_gallery_total))
            t_idx = torch.arange(q, dtype=torch.int64, device=dev)
        else:
            t_idx = true_idx.to(dev, torch.int64)
            if q and (int(t_idx.max()) >= n_gallery_total or int(t_idx.min()) < 0):
                raise IndexError("evaluate_ranks_sharded: true index outside the gallery")
        # (1) true-match distances from their owners.  No host round trip: every rank evaluates all Q pairs with the index
        # clamped into its slice and keeps the ones it owns; the others contribute zero to the sum
        d_true = torch.zeros(q, dtype=torch.float32, device=dev)
        if g_local > 0 and q > 0:
            mine = (t_idx >= g_offset) & (t_idx < g_offset + g_local)
            d = self.local.true_distances(ov_local, surface_embed, (t_idx - g_offset).clamp(0, g_local - 1))
            d_true = torch.where(mine, d.to(torch.float32), d_true)
        if self.peer is not None:
            d_true = self.peer.thresholds(d_true, t_idx, g_offset, g_local)     # the owners' values, stored into every rank's buffer
        elif self.world > 1:
            dist.all_reduce(d_true, op=dist.ReduceOp.SUM, group=group)
        # (2) local sweep
        if hasattr(self.local, "launch"):
            self.handle = self.local.launch(ov_local, surface_embed, d_true, t_idx, g_offset, topk)
        else:
            self.handle = _Done(*self.local.sweep(ov_local, surface_embed, d_true, t_idx, g_offset, topk))
        # (3) exchange of the (provisional) results
        self._n_host = None
        self._exchange()

    def _exchange(self):
        counts, td, ti, flagged = self.handle.provisional()
        self._flag_event = None
        if self.world == 1:
            self.out = (counts.to(torch.int64), td, ti)
            self._local_flag = flagged is not None
            return
        q = counts.shape[0]
        if self.peer is not None:
            total, md, mi, n_flag = self.peer.results(counts.to(torch.int32), td if self.topk else None, ti if self.topk else None, flagged)
            self.out = (total, md, mi)
            self._n_host = torch.empty((), dtype=torch.int32, pin_memory=True)
            self._n_host.copy_(n_flag[0], non_blocking=True)
            self._flag_event = torch.cuda.Event()
            self._flag_event.record()
            return
        if self.packed and q > 0:
            total, all_d, all_i, n_flag = _exchange_packed(counts, td, ti, self.world, self.group, flagged)
        else:
            total = counts.to(torch.int64).clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
            n_flag = (flagged if flagged is not None else torch.zeros(1, dtype=torch.int32, device=counts.device)).clone()
            dist.all_reduce(n_flag, op=dist.ReduceOp.SUM, group=self.group)
            n_flag = n_flag.sum()
            all_d = all_i = None
            if self.topk:
                all_d = torch.empty((self.world,) + tuple(td.shape), dtype=td.dtype, device=td.device)
                all_i = torch.empty((self.world,) + tuple(ti.shape), dtype=ti.dtype, device=ti.device)
                dist.all_gather(list(all_d.unbind(0)), td.contiguous(), group=self.group)
                dist.all_gather(list(all_i.unbind(0)), ti.contiguous(), group=self.group)
        if self.topk:
            md, mi = self.local.merge(all_d, all_i, self.topk)
            self.out = (total, md, mi)
        else:
            self.out = (total, None, None)
        if n_flag.is_cuda:
            self._n_host = torch.empty((), dtype=n_flag.dtype, pin_memory=True)
            self._n_host.copy_(n_flag, non_blocking=True)
            self._flag_event = torch.cuda.Event()
            self._flag_event.record()
        else:
            self._n_host = n_flag

    def result(self):
        if self.world == 1:
            if self.handle.finish():
                counts, td, ti, _ = self.handle.provisional()
                self.out = (counts.to(torch.int64), td, ti)
        else:
            if self._flag_event is not None:
                self._flag_event.synchronize()
            if int(self._n_host) & (1 << 30):
                raise RuntimeError("evaluate_ranks_sharded: the peer-memory exchange timed out waiting for another rank")
            if int(self._n_host) > 0:          # somebody's finish flagged a query: every rank re-does its own, then all exchange again
                self.handle.finish()
                self.handle = _Done(*self.handle.provisional()[:3])
                self._exchange()
                if self._flag_event is not None:
                    self._flag_event.synchronize()
            else:
                self.handle.finish()           # no-op: collects the evaluation's bookkeeping
        ranks, td, ti = self.out
        return (ranks, td, ti) if self.topk else ranks


def evaluate_ranks_sharded(ov_local, surface_embed, g_offset, n_gallery_total, true_idx=None, topk=0, group=None, local=None,
                           packed=None):
    """Sharded evaluate_ranks: every rank passes its gallery slice [g_offset, g_offset+G_local) and all queries.

    Returns ranks int64 [Q] (identical on every rank), plus merged (topk_dist, topk_idx) when topk > 0.
    Works without an initialised process group (world size 1).  packed: see PACKED_EXCHANGE.  ``ShardedEvaluation(...)`` with the
    same arguments is the same evaluation with result() as a separate, later step.
    """
    return ShardedEvaluation(ov_local, surface_embed, g_offset, n_gallery_total, true_idx=true_idx, topk=topk, group=group,
                             local=local, packed=packed).result()
