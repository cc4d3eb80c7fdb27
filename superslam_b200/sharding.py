"""Data-parallel sharding of stereo pairs over ranks and the fixed-size result gather.

Pairs are independent units: pair i is processed by rank i mod world (round-robin keeps a stream's
latency balanced); no collective is needed on the data path.  The only exchange is the result gather:
every rank packs its pairs into fixed-size padded records

    [pair_index, n_left, n_right, matches0[K], mscores0[K] (bit-cast), has_depth[K]]   (int32)

and one all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) delivers them to all ranks.
"""
from __future__ import annotations

import numpy as np


def shard_pairs(n_pairs: int, world: int, rank: int):
    """Indices of the pairs rank `rank` owns (round-robin)."""
    return list(range(rank, n_pairs, world))


def record_len(K: int) -> int:
    return 3 + 3 * K


def pack_records(pair_indices, counts, matches0, mscores0, has_depth, K: int, slots: int) -> np.ndarray:
    """Pack this rank's results into `slots` records (unused slots have pair_index -1)."""
    rec = np.full((slots, record_len(K)), -1, np.int32)
    for s, p in enumerate(pair_indices):
        rec[s, 0] = p
        rec[s, 1], rec[s, 2] = counts[2 * s], counts[2 * s + 1]
        rec[s, 3:3 + K] = matches0[s]
        rec[s, 3 + K:3 + 2 * K] = np.asarray(mscores0[s], np.float32).view(np.int32)
        rec[s, 3 + 2 * K:] = has_depth[s]
    return rec


def pack_records_block(first_pair: int, stride: int, counts, matches0, mscores0, has_depth, K: int) -> np.ndarray:
    """pack_records for a whole step at once (no per-pair Python loop): slot s holds pair `first_pair + stride * s`
    (rank r of a world of G owns pairs r, r + G, ...: first_pair = r + G * pairs_done, stride = G)."""
    P = len(matches0)
    rec = np.empty((P, record_len(K)), np.int32)
    rec[:, 0] = first_pair + stride * np.arange(P, dtype=np.int32)
    rec[:, 1:3] = np.asarray(counts, np.int32).reshape(P, 2)
    rec[:, 3:3 + K] = matches0
    rec[:, 3 + K:3 + 2 * K] = np.ascontiguousarray(mscores0, np.float32).view(np.int32)
    rec[:, 3 + 2 * K:] = has_depth
    return rec


def unpack_records(all_records: np.ndarray, K: int):
    """{pair_index: dict(n_left, n_right, matches0, mscores0, has_depth)} from gathered records."""
    out = {}
    for r in all_records.reshape(-1, record_len(K)):
        if r[0] < 0:
            continue
        out[int(r[0])] = dict(n_left=int(r[1]), n_right=int(r[2]), matches0=r[3:3 + K].copy(),
                              mscores0=r[3 + K:3 + 2 * K].copy().view(np.float32), has_depth=r[3 + 2 * K:].copy())
    return out


def gather_records(local: np.ndarray, dist, device=None) -> np.ndarray:
    """all_gather of the per-rank record blocks; returns [world, slots, record_len]."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return torch.stack(outs).cpu().numpy()
