"""world_size-2 gloo test of the N>1 path's host logic: scans shard by index, only the integer
counters are combined, and the combined counters equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metrics as om
from semanticlidarunc_b200 import dist as sdist

C, NB, N_SCANS = 20, 15, 7


def _scan(i):
    g = torch.Generator().manual_seed(100 + i)
    pred = torch.randint(0, C, (8, 64), generator=g)
    lab = torch.randint(0, C, (8, 64), generator=g)
    conf = torch.rand((8, 64), generator=g)
    return pred, lab, conf


def _counts(indices):
    cm = torch.zeros((C, C), dtype=torch.int64)
    bins = torch.zeros((3, NB), dtype=torch.int64)
    for i in indices:
        pred, lab, conf = _scan(i)
        cm += om.confusion_counts(pred, lab, C)
        valid = lab != 0
        n, nc, _ = om.ece_bin_counts(conf[valid].numpy(), (pred[valid] == lab[valid]).numpy(), NB)
        fx = torch.round(conf[valid].double() * 2.0 ** 32).to(torch.int64)
        idx = torch.from_numpy(om.ece_bin_index(conf[valid].numpy(), om.ece_edges(NB)))
        bins[0] += torch.from_numpy(n)
        bins[1] += torch.from_numpy(nc)
        bins[2].index_add_(0, idx, fx)
    return cm, bins


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sdist.shard_indices(N_SCANS)
    assert list(mine) == list(range(rank, N_SCANS, world))
    cm, bins = _counts(mine)
    sdist.allreduce_counts(cm, bins)
    # the sharded training loss combines exactly one number: the valid-pixel count (losses/evidential.py::_count_reducer)
    from semanticlidarunc_b200.losses.evidential import _count_reducer
    red = _count_reducer(True)
    count = torch.tensor([float(100 + rank)], dtype=torch.float64)
    red(count)
    assert float(count) == 201.0 and _count_reducer(None) is None
    # score histograms (AUROC / AURC state) combine the same way
    hist = torch.full((2, 16), rank + 1, dtype=torch.int64)
    sdist.allreduce_counts(hist)
    assert int(hist.sum()) == 3 * 32
    # without a GPU there are no NVLink mailboxes: the exchanges fall back to the process group, and the non-destructive
    # reduction (a packed copy) leaves the live accumulators alone
    assert sdist.peer_counter() is None and sdist.count_transport() == "nccl"
    live = torch.full((3,), rank + 1, dtype=torch.int64)
    total, _ = sdist.reduced_counts(live)
    assert total.tolist() == [3, 3, 3] and live.tolist() == [rank + 1] * 3
    q.put((rank, cm.numpy(), bins.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(120)
def test_two_rank_counts_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    cm_ref, bins_ref = _counts(range(N_SCANS))
    for _, cm, bins in got:
        assert np.array_equal(cm, cm_ref.numpy()) and np.array_equal(bins, bins_ref.numpy())


def test_shard_indices_cover_everything_once():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in sdist.shard_indices(4071, r, world))
        assert seen == list(range(4071))
    with pytest.raises(ValueError):
        sdist.shard_indices(10, 3, 2)


def test_allreduce_is_noop_without_process_group():
    cm = torch.ones((C, C), dtype=torch.int64)
    sdist.allreduce_counts(cm, None)
    assert int(cm.sum()) == C * C
    from semanticlidarunc_b200.losses.evidential import _count_reducer
    assert _count_reducer(True) is None and _count_reducer(None) is None        # no process group: local mean
