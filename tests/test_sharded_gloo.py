"""World-size-2 gloo test of the gallery-sharding exchange (CPU): the oracle stands in for the
kernels as the pluggable local compute, so what is tested is the host logic of
witw_b200/sharded.py -- ownership of true matches, the count all-reduce and the top-k gather/merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import witw_oracle as O


class OracleLocal(object):
    def true_distances(self, ov_local, su_owned, local_idx):
        out = torch.empty(su_owned.shape[0])
        for n in range(su_owned.shape[0]):
            _, d = O.match(ov_local[local_idx[n]: local_idx[n] + 1], su_owned[n: n + 1])
            out[n] = d[0, 0]
        return out

    def sweep(self, ov_local, su, d_true, true_idx, g_offset, topk):
        _, d = O.match(ov_local, su)
        counts = (d <= d_true.unsqueeze(0)).sum(0).to(torch.int64)
        if not topk:
            return counts, None, None
        td, ti = torch.topk(d.t(), min(topk, d.shape[0]), dim=1, largest=False, sorted=True)
        return counts, td.contiguous(), (ti + g_offset).to(torch.int32).contiguous()

    def merge(self, cand_d, cand_i, topk):
        p, q, k = cand_d.shape
        d = cand_d.permute(1, 0, 2).reshape(q, p * k)
        i = cand_i.permute(1, 0, 2).reshape(q, p * k)
        td, pos = torch.topk(d, k, dim=1, largest=False, sorted=True)
        return td, torch.gather(i, 1, pos)


class FlaggingLocal(OracleLocal):
    """A local compute whose first answer is wrong on one rank and says so (flagged > 0), as a CUDA shard does when its fp32
    finish has to re-do queries: the exchange must be repeated with the corrected results on every rank."""

    def __init__(self, bad_rank):
        self.bad_rank = bad_rank

    def launch(self, ov_local, su, d_true, true_idx, g_offset, topk):
        counts, td, ti = self.sweep(ov_local, su, d_true, true_idx, g_offset, topk)
        local = self

        class Handle(object):
            def __init__(self):
                self.bad = dist.get_rank() == local.bad_rank
                self.finished = False

            def provisional(self):
                if self.bad and not self.finished:
                    return counts + 5, td + 1.0, ti, torch.ones(1, dtype=torch.int32)
                return counts, td, ti, torch.zeros(1, dtype=torch.int32)

            def finish(self):
                self.finished = True
                return self.bad

        return Handle()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, packed=False, flagging=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from witw_b200.sharded import evaluate_ranks_sharded, shard_bounds

    ov, su, _ = O.synth_features(23, 17, fov=90, noise=10.0, seed=99)
    true_idx = torch.arange(17).flip(0)  # a permutation, so owners differ from the trivial layout
    lo, hi = shard_bounds(23, world, rank)
    local = FlaggingLocal(bad_rank=1) if flagging else OracleLocal()
    ranks, td, ti = evaluate_ranks_sharded(ov[lo:hi], su, lo, 23, true_idx=true_idx, topk=4, local=local, packed=packed)
    if rank == 0:
        np.savez(out, ranks=ranks.numpy(), td=td.numpy(), ti=ti.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("packed,flagging", [(False, False), (True, False), (True, True), (False, True)])
def test_two_shards_equal_one(tmp_path, packed, flagging):
    """packed: counts and top-k candidates travel in one all-gather instead of an all-reduce and two all-gathers.
    flagging: one rank's first results are provisional (its finish flagged queries); the exchange is repeated."""
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, packed, flagging), nprocs=2, join=True)
    got = np.load(out)
    ov, su, _ = O.synth_features(23, 17, fov=90, noise=10.0, seed=99)
    true_idx = torch.arange(17).flip(0)
    _, d = O.match(ov, su)
    want = (d <= d[true_idx, torch.arange(17)].unsqueeze(0)).sum(0).numpy()
    assert np.array_equal(got["ranks"], want)
    assert len(set(want.tolist())) > 1  # non-degenerate
    td, ti = torch.topk(d.t(), 4, dim=1, largest=False, sorted=True)
    assert np.array_equal(got["ti"], ti.numpy().astype(np.int32))
    assert np.allclose(got["td"], td.numpy())


def test_single_process_path():
    from witw_b200.sharded import evaluate_ranks_sharded

    ov, su, _ = O.synth_features(12, 12, fov=360, noise=20.0, seed=3)
    ranks = evaluate_ranks_sharded(ov, su, 0, 12, local=OracleLocal())
    assert np.array_equal(ranks.numpy(), O.rank_loop(ov, su))
