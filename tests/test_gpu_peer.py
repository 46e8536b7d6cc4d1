"""The peer-memory exchange of the sharded evaluation (witw_b200/peer.py, csrc/peer.cu) on real devices: two processes, each
with its own gallery shard, exchange thresholds, rank counts and top-k through each other's CUDA IPC buffers and must return what
one process returns for the whole gallery.  On a box with one GPU both ranks share it (IPC works between processes of one device;
the waiting kernels of the two processes take turns through the device's time slicing)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import witw_oracle as O
from parity_helpers import check_exact_results

pytestmark = pytest.mark.gpu

G, Q, K = 331, 150, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _data():
    true_idx = torch.arange(Q).flip(0) * 2          # spread over both shards, owners differ from the trivial layout
    ov, su_all, _ = O.synth_features(G, G, fov=180, noise=10.0, seed=41)      # query i is planted on item i
    return ov, su_all[true_idx].contiguous(), true_idx


def _worker(rank, world, port, out, rounds):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    n_dev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % n_dev)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # rendezvous only: the data path is the library's own kernels
    from witw_b200 import peer, sharded
    from witw_b200.sharded import CudaLocal, ShardedEvaluation, shard_bounds

    ov, su, true_idx = _data()
    lo, hi = shard_bounds(G, world, rank)
    res = None
    pending = None
    for _ in range(rounds):                                            # several evaluations in flight: sequence numbers, both parities
        cur = ShardedEvaluation(ov[lo:hi].to(dev), su.to(dev), lo, G, true_idx=true_idx.to(dev), topk=K, local=CudaLocal(path="tc"))
        assert cur.peer is not None, "the peer-memory path was not taken"
        if pending is not None:
            res = pending.result()
        pending = cur
    res = pending.result()
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out, ranks=res[0].cpu().numpy(), td=res[1].cpu().numpy(), ti=res[2].cpu().numpy())
    dist.barrier()
    peer.shutdown()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_exchange_through_peer_memory(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, 3), nprocs=2, join=True)
    got = np.load(out)
    ov, su, true_idx = _data()
    _, ref = O.match(ov, su)
    want = check_exact_results(ref, got["ranks"], torch.from_numpy(got["td"]), torch.from_numpy(got["ti"]), K, true_rows=true_idx)
    assert len(set(want.tolist())) > 5
