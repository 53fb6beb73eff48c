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
    Path(out_dir, f"ok{rank}").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [64, 1000, 1030, 77])
def test_sharded_merge_world2_gloo(tmp_path, n_cols):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cols, str(tmp_path)), nprocs=world, join=True)
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
