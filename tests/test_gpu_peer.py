"""The NVLink peer-memory exchanges (csrc/slu_peer.cu) on ONE GPU: a group of one rank is the degenerate case of the same
kernels (publish into the own mailbox, wait for it, sum one word), so the round-end single-GPU run exercises the mailbox
life cycle, the packed step / parity logic over many steps, CUDA-graph replays and the output bounds.  The two-GPU test
(tests/test_gpu_dist_loss.py) compares the real exchange with NCCL."""
import ctypes as C

import pytest
import torch

from semanticlidarunc_b200 import _lib, ops
from semanticlidarunc_b200.dist import PeerCounter

pytestmark = pytest.mark.gpu


@pytest.fixture()
def solo(cuda):
    own, handle = C.c_void_p(), (C.c_uint8 * 64)()
    _lib.check(_lib.lib().slu_peer_mailbox_create(C.byref(own), handle), "slu_peer_mailbox_create")
    assert own.value and any(handle)                    # a device pointer and a non-trivial IPC handle
    p = PeerCounter([own.value], 0, 1)
    yield p
    torch.cuda.synchronize()
    assert p.timeouts() == 0
    _lib.check(_lib.lib().slu_peer_mailbox_destroy(own), "slu_peer_mailbox_destroy")


def test_count_exchange_equals_count_valid_over_many_steps(cuda, solo):
    g = torch.Generator().manual_seed(11)
    cnt = torch.zeros(1, dtype=torch.float64, device=cuda)
    for k in range(25):                                 # odd and even steps: both parity halves of the mailbox
        n = 1 + 9973 * k
        tgt = torch.randint(0, 6, (n,), generator=g).to(cuda)
        keep = (torch.rand(n, generator=g) < 0.7).to(cuda) if k % 4 == 3 else None
        ign = () if keep is not None else ((0,), (0, 5), ())[k % 3]
        ops.count_valid_exchange(tgt, cnt, solo, ignore=ign, keep_mask=keep)
        ref = torch.zeros(1, dtype=torch.float64, device=cuda)
        ops.count_valid(tgt, ref, ignore=ign, keep_mask=keep)
        expect = int(keep.sum()) if keep is not None else int((~torch.isin(tgt, torch.tensor(list(ign) or [-1], device=cuda))).sum())
        assert float(cnt) == float(ref) == float(expect), (k, float(cnt), float(ref), expect)


def test_count_exchange_replays_in_a_cuda_graph(cuda, solo):
    tgt = torch.randint(0, 4, (4, 64, 2048), generator=torch.Generator().manual_seed(12)).to(cuda)
    cnt = torch.zeros(1, dtype=torch.float64, device=cuda)
    ops.count_valid_exchange(tgt, cnt, solo, ignore=(0,))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.count_valid_exchange(tgt, cnt, solo, ignore=(0,))
    expect = float((tgt != 0).sum())
    for _ in range(7):                                  # the step number advances in device memory: every replay is a new step
        cnt.fill_(-1.0)
        graph.replay()
        assert float(cnt) == expect


def test_vector_allreduce_of_one_rank_is_the_identity_and_stays_in_bounds(cuda, solo):
    g = torch.Generator().manual_seed(13)
    for n_a, n_b in ((400, 45), (1, 0), (512, 0), (300, 212), (7, 3)):
        a = torch.randint(-2**60, 2**60, (n_a,), generator=g).to(cuda)
        b = torch.randint(-2**60, 2**60, (n_b,), generator=g).to(cuda) if n_b else None
        big = torch.full((n_a + n_b + 64,), 0x5A5A5A5A, dtype=torch.int64, device=cuda)
        out = big[32:32 + n_a + n_b]
        ops.peer_allreduce_i64(a, b, solo, out=out)
        assert torch.equal(out, torch.cat([a, b]) if n_b else a)
        assert bool((big[:32] == 0x5A5A5A5A).all()) and bool((big[32 + n_a + n_b:] == 0x5A5A5A5A).all())
    with pytest.raises(ValueError):
        ops.peer_allreduce_i64(torch.zeros(513, dtype=torch.int64, device=cuda), None, solo)


def test_bad_arguments_are_rejected(cuda, solo):
    lib = _lib.lib()
    cnt = torch.zeros(1, dtype=torch.float64, device=cuda)
    tgt = torch.zeros(8, dtype=torch.int64, device=cuda)
    boxes = solo.boxes_array
    assert lib.slu_count_valid_exchange(_lib.ptr(tgt), None, 8, None, 0, boxes, 1, 1, 2.0, _lib.ptr(cnt), _lib.stream_ptr()) < 0     # rank outside the world
    assert lib.slu_count_valid_exchange(_lib.ptr(tgt), None, 8, None, 0, boxes, 0, 17, 2.0, _lib.ptr(cnt), _lib.stream_ptr()) < 0    # world too large
    assert lib.slu_count_valid_exchange(_lib.ptr(tgt), None, 0, None, 0, boxes, 0, 1, 2.0, _lib.ptr(cnt), _lib.stream_ptr()) < 0     # no pixels
    assert lib.slu_peer_mailbox_open(None, None) < 0
