"""Host-side logic of the multi-GPU paths, run with gloo at world size 2 on CPU (no kernels involved):
the sharded OSCR exchange (all-gather of known records with remote demotion + skip padding, ONE all-reduce
of the counts) against the oracle, and the sharded head's denominator / packed-gradient all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)


def _oscr_worker(rank, port, out_dir):
    _init(rank, port)
    try:
        from openset_imagenet_b200 import dist as D
        from oracle import oracle as O
        n, C = 4000, 23
        gt, s = O.make_oscr_case(n * WORLD, C, seed=17)
        s[::4] = np.round(s[::4], 2)                     # ties within and across shards
        sl = slice(rank * n, (rank + 1) * n)
        rec_unk, rec_known = O.oscr_records(gt[sl], s[sl], -1)       # stand-in for stage A of this shard
        padded = np.full(n, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        padded[:len(rec_known)] = rec_known
        rec_b = D.gather_known_records(torch.from_numpy(padded.view(np.int64)), rank, WORLD)
        assert rec_b.numel() == WORLD * n
        v = rec_b.numpy().view(np.uint64).reshape(WORLD, n)
        for q in range(WORLD):                           # remote records demoted, padding and local records untouched
            typ = v[q] & np.uint64(3)
            if q == rank:
                np.testing.assert_array_equal(v[q], padded)
            else:
                assert set(np.unique(typ).tolist()) <= {0, 3}
        taus, cc, fc = O.oscr_curve_from_records(rec_unk, v.reshape(-1))           # stand-in for osb_oscr_curve
        cc_t, fc_t = D.merge_counts(torch.from_numpy(cc), torch.from_numpy(fc))
        o_taus, o_cc, o_fc, _, _ = O.oscr_counts(gt, s, -1)
        np.testing.assert_array_equal(taus[:-1], o_taus)                           # same thresholds on every rank
        np.testing.assert_array_equal(cc_t.numpy()[:-1], o_cc)
        np.testing.assert_array_equal(fc_t.numpy()[:-1], o_fc)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


class _StubHead:
    """What ShardedHead touches of a FusedOpenSetHead."""

    def __init__(self, variant, C, K, bias):
        self.variant_name = variant
        self.weight = torch.nn.Parameter(torch.zeros(C, K))
        self.bias = torch.nn.Parameter(torch.zeros(C)) if bias else None
        self.den_hook = None


def _head_worker(rank, port, out_dir):
    _init(rank, port)
    try:
        from openset_imagenet_b200 import dist as D
        C, K = 5, 8
        for variant, bias in (("entropic", False), ("garbage", True)):
            head = _StubHead(variant, C, K, bias)
            sh = D.ShardedHead(head)
            assert head.den_hook is not None
            den = head.den_hook(torch.tensor([10.0 + rank], dtype=torch.float64))   # per-shard denominators 10, 11
            assert float(den) == 21.0
            g = torch.Generator().manual_seed(100 + rank)
            head.weight.grad = torch.randn(C, K, generator=g)
            if bias:
                head.bias.grad = torch.randn(C, generator=g)
            local_loss = torch.tensor(0.25 * (rank + 1))
            want_w = sum(torch.randn(C, K, generator=torch.Generator().manual_seed(100 + r)) for r in range(WORLD))
            for async_op in (False, True):
                w0 = head.weight.grad.clone()
                work, unpack = sh.allreduce_grads(local_loss, async_op=async_op)
                loss = unpack()
                assert abs(float(loss) - 0.75) < 1e-6                                 # sum of the shard losses
                if not async_op:
                    torch.testing.assert_close(head.weight.grad, want_w)
                head.weight.grad = w0 if not async_op else head.weight.grad
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _confidence_worker(rank, port, out_dir):
    """the confidence of a row-sharded validation set = one all-reduce of 2 float64 sums + counts."""
    _init(rank, port)
    try:
        from openset_imagenet_b200 import dist as D
        from openset_imagenet_b200.metrics import finish_confidence
        from oracle import oracle as O
        n, C = 3000, 19
        gt, s = O.make_oscr_case(n * WORLD, C, seed=5)
        sl = slice(rank * n, (rank + 1) * n)
        # stand-in for osb_confidence on this shard: the oracle's float64 sums and exact counts
        _, kn_c, _, ng_c, kn_sum, ng_sum = O.confidence(s[sl], gt[sl], 1.0 / C, -1, None)
        sums = torch.tensor([kn_sum, ng_sum], dtype=torch.float64)
        counts = torch.tensor([kn_c, ng_c, 0], dtype=torch.int64)
        g_sums, g_counts = D.allreduce_confidence(sums, counts)
        got = finish_confidence(g_sums, g_counts, C)
        want = O.confidence(s, gt, 1.0 / C, -1, None)
        assert (got[1], got[3]) == (want[1], want[3])                      # counts exact
        assert abs(got[0] - want[0]) < 1e-12 and abs(got[2] - want[2]) < 1e-12
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_oscr_worker, _head_worker, _confidence_worker])
def test_world_size_2_gloo(worker, tmp_path):
    port = _free_port()
    mp.spawn(worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(WORLD))


def test_single_process_paths_need_no_process_group():
    from openset_imagenet_b200 import dist as D
    assert D.world() == (0, 1)
    a, b = torch.arange(4), torch.arange(4) * 2
    c, f = D.merge_counts(a, b)
    assert torch.equal(c, a) and torch.equal(f, b)
    rec = torch.tensor([5, 6, -1], dtype=torch.int64)
    assert D.gather_known_records(rec, 0, 1) is rec
