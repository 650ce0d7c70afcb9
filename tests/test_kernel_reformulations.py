"""CPU models of two re-formulations the ring feature kernel (a-loam_b200/csrc/features.cu) relies on.  They restate the ARGUMENT,
not the kernel: the GPU parity tests compare the kernel itself with the oracle bit for bit.

1. Greedy picks (scanRegistration.cpp:282-398): the six segments of a ring are walked in parallel assuming no marks spill in from
   the previous segment, then re-walked in parallel fixed-point rounds.  Claim: the result equals the sequential sweep in which
   every segment receives the exact spill of its predecessor.
2. Per-ring VoxelGrid (scanRegistration.cpp:401-407): the kernel sorts EVERY in-range position by its absolute voxel coordinates
   before the picks are known and drops the picked positions afterwards.  Claim: the order equals PCL's sort by the voxel index
   relative to the bounding box of the surviving points (stable by position).
"""
import numpy as np
import pytest


def walk_segment(curv, reach_b, reach_f, sp, ep, spill_in):
    """One segment: picks in the reference's order with `spill_in` (set of positions already marked).  Returns
    (less_sharp picks, flat picks, set of marked positions beyond ep)."""
    picked = set(p for p in spill_in if sp <= p <= ep)
    spill_out = set()

    def mark(ind):
        for p in range(ind - reach_b[ind], ind + reach_f[ind] + 1):
            if sp <= p <= ep:
                picked.add(p)
            elif p > ep:
                spill_out.add(p)

    less = []
    order = sorted(range(sp, ep + 1), key=lambda i: (curv[i], i))
    for i in reversed(order):                       # largest curvature first, ties -> larger index
        if i in picked or not curv[i] > 0.1:
            continue
        if len(less) >= 20:
            break
        less.append(i)
        mark(i)
    flat = []
    for i in order:                                 # smallest first, ties -> smaller index
        if i in picked or not curv[i] < 0.1:
            continue
        flat.append(i)
        if len(flat) >= 4:
            break
        mark(i)
    return less, flat, spill_out


def bounds(n, w):
    s_loc, e_loc = 5, n - 6
    span = e_loc - s_loc
    return s_loc + span * w // 6, s_loc + span * (w + 1) // 6 - 1


def sequential(curv, rb, rf):
    n = len(curv)
    out, spill = [], set()
    for w in range(6):
        sp, ep = bounds(n, w)
        less, flat, spill = walk_segment(curv, rb, rf, sp, ep, spill)
        out.append((less, flat))
    return out


def fixed_point(curv, rb, rf):
    n = len(curv)
    res, spill_out, assumed = [None] * 6, [set()] * 6, [set()] * 6
    for w in range(6):                              # speculative pass: nothing spills in
        sp, ep = bounds(n, w)
        less, flat, so = walk_segment(curv, rb, rf, sp, ep, set())
        res[w], spill_out[w] = (less, flat), so
    rounds = 0
    while True:
        todo = []
        for w in range(1, 6):                       # all decisions of a round read the state of the previous round
            S, A = spill_out[w - 1], assumed[w]
            if S == A:
                continue
            own = set(res[w][0]) | set(res[w][1])
            hit = bool(A - S) or bool((S - A) & own)
            todo.append((w, S, hit))
        if not any(h for _, _, h in todo):
            break
        rounds += 1
        new = {}
        for w, S, hit in todo:
            if hit:
                sp, ep = bounds(n, w)
                new[w] = walk_segment(curv, rb, rf, sp, ep, S)
        for w, S, hit in todo:
            assumed[w] = S
            if hit:
                less, flat, so = new[w]
                res[w], spill_out[w] = (less, flat), so
        assert rounds <= 5
    return res, rounds


@pytest.mark.parametrize("seed", range(40))
def test_parallel_rewalk_rounds_equal_the_sequential_sweep(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(60, 400))
    # small segments and a coarse curvature alphabet: many ties, many picks next to the segment boundaries
    curv = rng.choice(np.array([0.0, 0.05, 0.1, 0.2, 0.5, 1.0, 3.0], np.float32), size=n)
    rb = rng.integers(0, 6, size=n)
    rf = rng.integers(0, 6, size=n)
    want = sequential(curv, rb, rf)
    got, rounds = fixed_point(curv, rb, rf)
    assert got == want


def test_the_rewalk_model_exercises_conflicts():
    hits = 0
    for seed in range(40):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(60, 400))
        curv = rng.choice(np.array([0.0, 0.05, 0.1, 0.2, 0.5, 1.0, 3.0], np.float32), size=n)
        rb = rng.integers(0, 6, size=n); rf = rng.integers(0, 6, size=n)
        hits += fixed_point(curv, rb, rf)[1] > 0
    assert hits >= 10          # the equality above is not vacuous


@pytest.mark.parametrize("seed", range(10))
def test_absolute_voxel_order_equals_pcl_relative_index_order(seed):
    rng = np.random.default_rng(100 + seed)
    n = 1500
    pts = (rng.normal(size=(n, 3)) * np.array([30.0, 30.0, 2.0])).astype(np.float32)
    keep = rng.random(n) > 0.1                      # "label <= 0": the picks are removed AFTER the sort in the kernel
    inv = np.float32(1.0) / np.float32(0.2)
    fl = np.floor(pts * inv).astype(np.int64)       # floorf(x * inverse_leaf)
    # PCL: bounding box of the surviving points -> relative index -> sort by (index, position)
    surv = np.nonzero(keep)[0]
    mn, mx = pts[surv].min(0), pts[surv].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    div_b = np.floor(mx * inv).astype(np.int64) - min_b + 1
    rel = (fl[surv] - min_b)
    idx = rel[:, 0] + rel[:, 1] * div_b[0] + rel[:, 2] * div_b[0] * div_b[1]
    pcl_order = surv[np.lexsort((surv, idx))]
    # kernel: all positions by (z, y, x) absolute voxel coordinates then position, picked positions dropped afterwards
    key = ((fl[:, 2] + 65536) << 34) | ((fl[:, 1] + 65536) << 17) | (fl[:, 0] + 65536)
    all_order = np.lexsort((np.arange(n), key))
    mine = all_order[keep[all_order]]
    assert np.array_equal(mine, pcl_order)
    # and the voxel boundaries coincide
    assert np.array_equal(np.diff(key[mine]) != 0, np.diff(idx[np.argsort(surv)[np.searchsorted(np.sort(surv), pcl_order)]]) != 0)
