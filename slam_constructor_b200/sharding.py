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


# ---- K5: a batch of M matches (score_windows) ----------------------------------------------------------------------
def window_chunks(M, nranks):
    """(sharded, chunk): batches of at least 4 matches per rank are split into equal chunks of ceil(M / nranks)
    (the last ranks may get a short or empty one) and all-gathered; smaller batches are scored by every rank"""
    if nranks > 1 and M >= 4 * nranks:
        return True, (M + nranks - 1) // nranks
    return False, M


# ---- K6: particles (particles.cu) ------------------------------------------------------------------------------------
def particle_range(n, rank, nranks):
    """rank r owns the particles [lo, hi): chunks of ceil(n / nranks), so owner(i) = i // chunk"""
    chunk = (n + nranks - 1) // nranks
    lo = min(n, chunk * rank)
    return lo, min(n, lo + chunk), chunk


def resample_plan(src, rank, nranks):
    """What slamgpu_particles_resample does on `rank` for the draw src[i] = the particle that particle i becomes a
    copy of.  Returns (sends, recvs, local) where
      sends = [(i, peer)]: this rank ships its map src[i] to `peer` (the owner of i), in increasing i;
      recvs = [(i, peer)]: this rank receives the map of src[i] from `peer` into a staging buffer;
      local = {i: ("keep", i) | ("move", src[i]) | ("copy", src[i]) | ("staged", i)} for the particles it owns.
    A map is moved (no copy) to the first local particle that draws it when its own slot does not keep it; every
    other use is a copy into a map nobody kept."""
    n = len(src)
    lo, hi, chunk = particle_range(n, rank, nranks)
    owner = lambda i: i // chunk
    sends, recvs = [], []
    for i in range(n):
        q, r = owner(src[i]), owner(i)
        if q == r:
            continue
        if rank == q:
            sends.append((i, r))
        elif rank == r:
            recvs.append((i, q))
    local, taken = {}, set()
    for i in range(lo, hi):
        if src[i] == i:
            local[i] = ("keep", i); taken.add(i)
    for i in range(lo, hi):
        if i in local or not lo <= src[i] < hi:
            continue
        if src[i] not in taken:
            local[i] = ("move", src[i]); taken.add(src[i])
    for i in range(lo, hi):
        if i not in local:
            local[i] = ("copy", src[i]) if lo <= src[i] < hi else ("staged", i)
    return sends, recvs, local
