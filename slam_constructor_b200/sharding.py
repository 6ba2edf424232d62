"""Host-side statement of how a candidate set is sharded over ranks and how per-rank results merge.

The library does both on the device (score.cu: slice_of / k_finalize, one 32-byte-per-rank NCCL
all-gather); this module states the same protocol for launchers, bench.py and the world_size-2 gloo
tests: contiguous index ranges per rank, winner = highest score, lowest candidate index on ties, and
only if it beats the initial pose's score (the reference's sequential strict-'<' accept loop,
src/core/scan_matchers/pose_enumeration_scan_matcher.h:48-65)."""
import math


def slice_of(total, rank, nranks):
    """contiguous [begin, end) of `total` units owned by `rank` (same integer arithmetic as score.cu)"""
    return total * rank // nranks, total * (rank + 1) // nranks


def grid_slice(nx, ny, nt, rank, nranks):
    """the brute-force grid is split by whole (theta, y) rows: candidate index range of `rank`"""
    r0, r1 = slice_of(nt * ny, rank, nranks)
    return r0 * nx, r1 * nx


def local_best(scores, first_index):
    """(score, index) of one rank's slice: max score, lowest index on ties, NaN never wins"""
    best, idx = -math.inf, None
    for k, s in enumerate(scores):
        if s == s and s > best:
            best, idx = s, first_index + k
    return (best, idx if idx is not None else 2 ** 63 - 1)


def merge_best(per_rank, init_score):
    """merge the all-gathered (score, index) pairs; returns (best_score, best_index or -1)"""
    best, idx = -math.inf, 2 ** 63 - 1
    for s, i in per_rank:
        if s > best or (s == best and i < idx):
            best, idx = s, i
    if not (init_score < best) or idx == 2 ** 63 - 1:
        return init_score, -1
    return best, idx
