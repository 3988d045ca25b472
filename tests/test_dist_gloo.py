"""Host-side logic of the multi-GPU path on CPU: two ``gloo`` ranks exchange exactly what the NCCL ranks exchange
(per-image class-probability sums, min/max, dominant-class histogram, per-rank top-k candidate keys) and must
reproduce the single-process result.  The kernels themselves are covered by the ``-m gpu`` tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from helpers import class_weights_ref
from mulactseg_b200 import dist as mdist, selection


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_img, nseg, c, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)                      # every rank draws the WHOLE pool, then keeps its shard
        prob_sum = torch.rand((n_img, c), generator=g, dtype=torch.float64) * 1000
        scores = torch.rand((n_img, nseg), generator=g)
        scores[scores < 0.05] = 0.0
        dominant = torch.randint(0, c, (n_img, nseg), generator=g)
        in_pool = torch.rand((n_img, nseg), generator=g) < 0.8
        lo, hi = mdist.shard_range(n_img, rank, world)

        # (1) class weights from gathered per-image sums == single process (mean of per-batch means, batches of 4)
        gathered = mdist.all_gather_rows(prob_sum[lo:hi].contiguous())
        known = mdist.all_gather_rows(prob_sum[lo:hi].contiguous(), None, mdist.shard_sizes(n_img, world))
        assert torch.equal(gathered, known)                      # known shard sizes: same rows, no size exchange
        worst = mdist.all_reduce_min(torch.tensor([5 - 6 * rank], dtype=torch.int64))
        assert int(worst) == 5 - 6 * (world - 1)
        assert torch.equal(gathered, prob_sum)                   # pool order restored: the class weights follow
        assert torch.equal(class_weights_ref(gathered, 1000, 4, 6.0), class_weights_ref(prob_sum, 1000, 4, 6.0))

        # (2) min over non-zero / max
        shard = scores[lo:hi]
        nz = shard[shard != 0]
        local = torch.tensor([nz.min() if nz.numel() else float("inf"), shard.max() if shard.numel() else float("-inf")])
        mm = mdist.all_reduce_minmax(local)
        assert mm[0] == scores[scores != 0].min() and mm[1] == scores.max()

        # (3) dominant-class histogram
        hist = mdist.all_reduce_sum(torch.bincount(dominant[lo:hi].reshape(-1), minlength=c))
        assert torch.equal(hist, torch.bincount(dominant.reshape(-1), minlength=c))

        # (4) top-k merge: per-rank k best keys -> all_gather -> k best of the union == global k best
        def keys_of(sc, pool, first_img):
            bits = sc.numpy().view(np.uint32).astype(np.uint64) | np.uint64(0x80000000)      # scores >= 0: ordered bits
            tie = (np.arange(first_img, first_img + sc.shape[0], dtype=np.uint64)[:, None] * np.uint64(nseg)
                   + np.arange(nseg, dtype=np.uint64)[None, :])
            key = (bits << np.uint64(32)) | tie
            key[~pool.numpy()] = 0
            return key.reshape(-1)

        mine = np.sort(keys_of(scores[lo:hi], in_pool[lo:hi], lo))[::-1][:k].copy()
        buf = torch.zeros(k, dtype=torch.int64)
        buf[: len(mine)] = torch.from_numpy(mine.view(np.int64))
        count = torch.tensor([int((mine != 0).sum())], dtype=torch.int32)
        merged = mdist.gather_candidates(buf, count, k).numpy().view(np.uint64)
        # the fast path's message: candidates + count in the last slot, one all_gather_into_tensor, counts cleared after
        msg = torch.cat([buf, count.to(torch.int64)])
        boxes = mdist.gather_messages(msg).clone()
        assert boxes.shape == (world, k + 1) and int(boxes[rank, k]) == int(count)
        boxes[:, k] = 0
        np.testing.assert_array_equal(np.sort(boxes[:, :k].reshape(-1).numpy().view(np.uint64)), np.sort(merged))
        best = np.sort(merged)[::-1][:k]
        best = best[best != 0]                                   # k may exceed the number of pool regions
        want = np.sort(keys_of(scores, in_pool, 0))[::-1][:k]
        want = want[want != 0]
        np.testing.assert_array_equal(best, want)
        np.save(os.path.join(out_dir, f"best{rank}.npy"), best)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("n_img,k", [(7, 40), (8, 300), (3, 5)])
def test_two_rank_exchange_matches_single_process(tmp_path, n_img, k):
    world, nseg, c = 2, 24, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_img, nseg, c, k, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "best0.npy"), np.load(tmp_path / "best1.npy")
    np.testing.assert_array_equal(a, b)            # every rank ends with the same ranking


def test_shard_ranges_cover_the_pool():
    for n in (1, 7, 8, 2975, 10582):
        for world in (1, 2, 3, 8):
            ranges = [mdist.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = mdist.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def test_cumulative_cut_and_key_decoding():
    costs = np.array([1, 2, 1, 3, 1, 1])
    assert selection.cumulative_cut(costs, 3) == 3          # stops AFTER the pick that exceeds the budget (strict >)
    assert selection.cumulative_cut(costs, 100) == 6
    im_idx = [["b.png", "lb.png", "sb.png"], ["a.png", "la.png", "sa.png"]]
    rank = selection.image_ranks(im_idx)
    assert rank.tolist() == [1, 0]
    nseg = 4
    score = np.array([0.5, 0.25], dtype=np.float32)
    bits = score.view(np.uint32).astype(np.uint64) | np.uint64(0x80000000)
    keys = (bits << np.uint64(32)) | np.array([rank[0] * nseg + 3, rank[1] * nseg + 1], dtype=np.uint64)
    out = selection.decode_keys(keys, nseg, im_idx, rank)
    assert out == [(0.5, "b.png,lb.png,sb.png", 3), (0.25, "a.png,la.png,sa.png", 1)]
