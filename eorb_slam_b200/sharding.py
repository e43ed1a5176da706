"""Host-side plumbing of the multi-GPU paths: one process per GPU.

  * frame batches / event windows: contiguous split by unit, no collective (`unit_range`);
  * Hamming search over a row-sharded database: every rank searches its shard (`ORBmatcher.search_device` ->
    eorb_best2[nq]), ONE all-gather of those 16-byte records, then the (dist, global index) merge
    (`ORBmatcher.merge_device` on the GPU; `merge_best2_host` is the same ordering in numpy for host-side
    consumers and for the gloo tests).

`torch.distributed` is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

BEST2_DTYPE = np.dtype([("key1", "<u8"), ("key2", "<u8")])
NONE_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def unit_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous [begin, end) of rank's units; the first n%world ranks get one extra unit"""
    base, rem = divmod(n_units, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def pack_best2(best_dist, best_idx, second_dist, second_idx) -> np.ndarray:
    """(dist << 32 | global idx) keys; idx < 0 -> none"""
    out = np.empty(len(best_dist), BEST2_DTYPE)
    k1 = (np.asarray(best_dist, np.uint64) << np.uint64(32)) | np.asarray(best_idx, np.int64).astype(np.uint64)
    k2 = (np.asarray(second_dist, np.uint64) << np.uint64(32)) | np.asarray(second_idx, np.int64).astype(np.uint64)
    out["key1"] = np.where(np.asarray(best_idx) >= 0, k1, NONE_KEY)
    out["key2"] = np.where(np.asarray(second_idx) >= 0, k2, NONE_KEY)
    return out


def merge_best2_host(parts: np.ndarray) -> np.ndarray:
    """parts: (nshards, nq) BEST2 records -> (nq,) two smallest keys of the union (same min/max network as
    merge_best2_kernel; ties on distance resolve to the lowest global index because the index is in the key)"""
    parts = np.asarray(parts).view(BEST2_DTYPE).reshape(parts.shape[0], -1)
    k1 = np.full(parts.shape[1], NONE_KEY, np.uint64); k2 = k1.copy()
    for p in parts:
        hi = np.maximum(k1, p["key1"]); lo2 = np.minimum(k2, p["key2"])
        k1 = np.minimum(k1, p["key1"]); k2 = np.minimum(hi, lo2)
    out = np.empty(parts.shape[1], BEST2_DTYPE)
    out["key1"] = k1; out["key2"] = k2
    return out


def finalize_matches(merged: np.ndarray, th: int, ratio: float) -> np.ndarray:
    """threshold + ratio test of ORBmatcher (ORBmatcher.cc:380-382) on merged keys -> MATCH_DTYPE"""
    from .synth import MATCH_DTYPE
    out = np.zeros(len(merged), MATCH_DTYPE)
    none1 = merged["key1"] == NONE_KEY; none2 = merged["key2"] == NONE_KEY
    out["best_dist"] = np.where(none1, 256, (merged["key1"] >> np.uint64(32)).astype(np.int64))
    out["best_idx"] = np.where(none1, -1, (merged["key1"] & np.uint64(0xFFFFFFFF)).astype(np.int64))
    out["second_dist"] = np.where(none2, 256, (merged["key2"] >> np.uint64(32)).astype(np.int64))
    ok = (~none1) & (out["best_dist"] <= th) & (out["best_dist"].astype(np.float32) < np.float32(ratio) * out["second_dist"].astype(np.float32))
    out["accepted"] = ok.astype(np.int32)
    return out


def all_gather_best2(part, world: int):
    """part: torch uint8 tensor of nq*16 bytes (device for NCCL, CPU for gloo) -> gathered (world*nq*16)"""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * part.numel(), dtype=torch.uint8, device=part.device)
    if dist.get_backend() == "gloo":
        chunks = list(out.chunk(world))
        dist.all_gather(chunks, part)
    else:
        dist.all_gather_into_tensor(out, part)
    return out
