"""N > 1 host logic on CPU: world_size 2 over gloo. Each rank owns a contiguous shard of the gallery columns, builds
its local thresholds / counts / bit columns with the ORACLE (no kernel calls on a GPU-less box), and the package's
collectives (mmx.sharded) must merge them into exactly the unsharded structures."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_structures(scores, gt_sets, a, b, thr=None):
    """Oracle restatement of one shard's KR passes (MODE_THRESHOLD then MODE_COUNT) on columns [a, b)."""
    nq = scores.shape[0]
    if thr is None:
        t = torch.full((nq,), float("-inf"))
        for q in range(nq):
            for c in gt_sets[q]:
                if a <= c < b:
                    t[q] = max(t[q], scores[q, c])
        return t
    above = scores[:, a:b] > thr[:, None]
    cnt = above.sum(1).to(torch.int32)
    w = (b - a + 31) // 32
    padded = np.zeros((nq, w * 32), dtype=np.uint8)
    padded[:, : b - a] = above.numpy()
    bits = torch.from_numpy(np.packbits(padded, axis=1, bitorder="little").view(np.int32).copy())
    return cnt, bits


def _worker(rank, world, port, n_cols, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmx
    from mmx.ops import ProtocolRanks
    rs = np.random.RandomState(0)
    nq = 50
    scores = torch.from_numpy(rs.randn(nq, n_cols).astype(np.float32))
    gt_sets = [sorted(rs.choice(n_cols, size=rs.randint(1, 4), replace=False).tolist()) for _ in range(nq)]
    bounds = mmx.sharded.shard_bounds(n_cols, world)
    a, b = bounds[rank]
    thr = mmx.sharded.merge_thresholds(_local_structures(scores, gt_sets, a, b))
    cnt, bits = _local_structures(scores, gt_sets, a, b, thr)
    merged = mmx.sharded.merge_ranks(ProtocolRanks(thr=thr, cnt_above=cnt, bits=bits, n_cols=b - a), bounds)
    # unsharded truth
    thr_full = _local_structures(scores, gt_sets, 0, n_cols)
    cnt_full, bits_full = _local_structures(scores, gt_sets, 0, n_cols, thr_full)
    ok = (torch.equal(thr, thr_full) and torch.equal(merged.cnt_above, cnt_full)
          and torch.equal(merged.bits, bits_full) and merged.n_cols == n_cols)
    # the ranks' blocks of bootstrap iterations tile [0, B) (each rank keeps only those rows of the numpy stream)
    ok = ok and mmx.sharded.stats_block(1001) == ((0, 501), (501, 1001))[rank] and mmx.sharded.stats_block(1) == ((0, 1), (1, 1))[rank]
    Path(out_dir, f"ok{rank}").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [64, 1000, 1030, 77])
def test_sharded_merge_world2_gloo(tmp_path, n_cols):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cols, str(tmp_path)), nprocs=world, join=True)
    assert [Path(tmp_path, f"ok{r}").read_text() for r in range(world)] == ["1", "1"]


def _worker_lists(rank, world, port, out_dir):
    """LIST mode over gloo: the per-shard truncated above-GT lists travel in one all-gather, and the exact recount of
    the pairs rank 0 could not settle is served by every rank over its own columns (sharded.unlisted_round)."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmx
    from mmx.ops import ProtocolRanks
    rs = np.random.RandomState(1)
    nq, n_cols, cap, B = 40, 333, 6, 9
    scores = torch.from_numpy(rs.randn(nq, n_cols).astype(np.float32))
    gt_sets = [[int(rs.randint(n_cols))] for _ in range(nq)]
    mult = torch.from_numpy(rs.poisson(1.0, size=(B, n_cols)).astype(np.uint8))
    bounds = mmx.sharded.shard_bounds(n_cols, world)
    a, b = bounds[rank]
    thr = mmx.sharded.merge_thresholds(_local_structures(scores, gt_sets, a, b))
    above = scores[:, a:b] > thr[:, None]
    list_i = torch.full((1, nq, cap), -7, dtype=torch.int32)
    list_n = torch.zeros((1, nq), dtype=torch.int32)
    for q in range(nq):
        cols = (torch.nonzero(above[q]).flatten() + a).to(torch.int32)
        list_n[0, q] = len(cols) + (1 if len(cols) > cap else 0)          # the cursor may overshoot the capacity
        list_i[0, q, : min(len(cols), cap)] = cols[:cap]
    local = ProtocolRanks(thr=thr, cnt_above=above.sum(1).to(torch.int32), bits=None, n_cols=b - a, list_i=list_i,
                          list_n=list_n)
    merged = mmx.sharded.merge_ranks(local, bounds)
    full_above = scores > thr[:, None]
    ok = torch.equal(merged.cnt_above, full_above.sum(1).to(torch.int32)) and merged.list_i.shape == (world, nq, cap)
    for q in range(nq):
        got = [int(c) for r in range(world) for c in merged.list_i[r, q, : min(int(merged.list_n[r, q]), cap)]]
        ok = ok and all(bool(full_above[q, c]) for c in got) and len(set(got)) == len(got)
        if int(merged.cnt_above[q]) <= cap:
            ok = ok and sorted(got) == torch.nonzero(full_above[q]).flatten().tolist()

    def local_resolve(pairs, mult_rows):        # this shard's share: sum_{col in [a, b)} m[b, col] * above[q, col]
        return (mult_rows[pairs[:, 0].long()][:, a:b].long() * above[pairs[:, 1].long()].long()).sum(1)
    if rank == 0:
        pairs = torch.tensor([[0, 3], [2, 17], [1, 3], [2, 39]], dtype=torch.int32)
        rows = mult[[4, 0, 8]].contiguous()                              # pairs[:, 0] indexes these rows
        c = mmx.sharded.unlisted_round(local_resolve, pairs, rows)
        want = (rows[pairs[:, 0].long()].long() * full_above[pairs[:, 1].long()].long()).sum(1)
        ok = ok and torch.equal(c, want)
        ok = ok and mmx.sharded.unlisted_round(local_resolve, None, None) is None      # a K3 call with nothing unsettled
    else:
        ok = ok and mmx.sharded.unlisted_round(local_resolve, None, None) is not None
        ok = ok and mmx.sharded.unlisted_round(local_resolve, None, None) is None
    Path(out_dir, f"ok{rank}").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_list_mode_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker_lists, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [Path(tmp_path, f"ok{r}").read_text() for r in range(world)] == ["1", "1"]


def test_shard_bounds():
    sys.path.insert(0, str(ROOT))
    import mmx
    for n, w in ((5000, 8), (5000, 1), (100, 8), (1030, 4), (31, 2)):
        b = mmx.sharded.shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
        assert max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1
    assert all(x[0] % 32 == 0 for x in mmx.sharded.shard_bounds(5000, 8, align=32) if x[0] < 5000)


def test_place_bit_columns_any_alignment():
    """Bit columns of shards that start in the middle of a 32-column word land on their global positions."""
    sys.path.insert(0, str(ROOT))
    import mmx
    rs = np.random.RandomState(4)
    for n_cols, cuts in ((100, [0, 33, 64, 65, 100]), (5000, [0, 625, 1250, 1875, 2500, 3125, 3750, 4375, 5000]), (70, [0, 70]),
                         (64, [0, 1, 2, 64]), (40, [0, 0, 40])):
        above = rs.rand(7, n_cols) < 0.3
        bounds = list(zip(cuts[:-1], cuts[1:]))
        wmax = max((e - b + 31) // 32 for b, e in bounds)
        parts = []
        for b, e in bounds:
            padded = np.zeros((7, max(wmax, 1) * 32), dtype=np.uint8)
            padded[:, : e - b] = above[:, b:e]
            parts.append(torch.from_numpy(np.packbits(padded, axis=1, bitorder="little").view(np.int32).copy()))
        got = mmx.sharded.place_bit_columns(parts, bounds)
        full = np.zeros((7, (n_cols + 31) // 32 * 32), dtype=np.uint8)
        full[:, :n_cols] = above
        want = torch.from_numpy(np.packbits(full, axis=1, bitorder="little").view(np.int32).copy())
        assert got.dtype == torch.int32 and torch.equal(got, want)
