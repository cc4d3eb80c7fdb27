"""world_size-2 test (gloo, CPU) of the N>1 path: round-robin pair sharding and the fixed-size result
gather used by bench.py / a multi-process deployment.  The per-pair compute itself needs no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superslam_b200 import sharding

K = 16
N_PAIRS = 5


def _fake_result(pair):
    rng = np.random.default_rng(pair)
    n0, n1 = int(rng.integers(1, K + 1)), int(rng.integers(1, K + 1))
    m = np.full(K, -1, np.int32)
    m[:n0] = rng.integers(-1, n1, n0)
    s = np.zeros(K, np.float32)
    s[:n0] = rng.random(n0).astype(np.float32)
    hd = (m >= 0).astype(np.int32)
    return n0, n1, m, s, hd


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_pairs(N_PAIRS, world, rank)
    counts, ms, ss, hs = [], [], [], []
    for p in mine:
        n0, n1, m, s, hd = _fake_result(p)
        counts += [n0, n1]
        ms.append(m), ss.append(s), hs.append(hd)
    slots = (N_PAIRS + world - 1) // world
    rec = sharding.pack_records(mine, counts, ms, ss, hs, K, slots)
    allrec = sharding.gather_records(rec, dist)
    out = sharding.unpack_records(allrec, K)
    ok = sorted(out) == list(range(N_PAIRS))
    for p, r in out.items():
        n0, n1, m, s, hd = _fake_result(p)
        ok &= (r["n_left"], r["n_right"]) == (n0, n1) and np.array_equal(r["matches0"], m)
        ok &= np.array_equal(r["mscores0"], s) and np.array_equal(r["has_depth"], hd)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)            # max-over-ranks style reduction used for timing in bench.py
    q.put((rank, bool(ok), int(t.item())))
    dist.destroy_process_group()


def test_round_robin_sharding():
    assert sharding.shard_pairs(5, 2, 0) == [0, 2, 4] and sharding.shard_pairs(5, 2, 1) == [1, 3]
    assert sharding.shard_pairs(3, 8, 5) == [] and sharding.shard_pairs(8, 8, 7) == [7]
    assert sorted(sum((sharding.shard_pairs(13, 4, r) for r in range(4)), [])) == list(range(13))


def test_two_rank_gather_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res) and all(total == N_PAIRS for _, _, total in res)


def test_block_packing_equals_per_pair_packing():
    rng = np.random.default_rng(3)
    P = 6
    counts = rng.integers(0, K + 1, 2 * P).astype(np.int32)
    m = rng.integers(-1, K, (P, K)).astype(np.int32)
    s = rng.random((P, K)).astype(np.float32)
    hd = rng.integers(0, 2, (P, K)).astype(np.uint8)
    a = sharding.pack_records([3 + 4 * i for i in range(P)], counts.tolist(), m, s, hd, K, P)
    b = sharding.pack_records_block(3, 4, counts, m, s, hd, K)
    assert np.array_equal(a, b)
