// ray_pull.h -- the "pull" form of the occupancy ray cast: instead of walking every beam and adding 1 to each cell it
// crosses (one atomic per touch), every CELL of a touched patch asks how many beams cross it.
//
// Reference: Map::computeRay src/sdm/map.cpp:198-227 (integer Bresenham, both end cells excluded) as called by
// PFSlam2D::updateParticleMaps src/pf_slam2d.cpp:463-504 / Slam2D::updateMaps src/slam2d.cpp:247-321, and
// FrequencyOccupancyMap::setFree src/sdm/frequency_occupancy_map.cpp:65-74.
//
// Exactness.  All beams of a 2-D scan start in the same cell O (the sensor origin; Options::truncated_ray == 0) and are
// planar.  For a beam with end offset (ex, ey) from O let n = max(|ex|, |ey|) be its major length and d = min(..) its minor
// length (|ex| == |ey| counts as x-major, like computeRay's `if 2 err >= n` on both axes).  After i steps (1 <= i <= n - 1)
// the walk sits at major offset i and minor offset k(i) = floor((2 i d + n) / (2 n))  (closed form of the error accumulator,
// ray_core.h SegWalk).  Hence the beam crosses the cell at major offset a, minor offset b  iff
//        1 <= a <= n - 1   and   k(a) == b      <=>      2 b n <= 2 a d + n < 2 (b + 1) n,
// i.e. iff its slope d / n lies in [(2 b - 1) / (2 a), (2 b + 1) / (2 a)) and it is long enough.  Beams are split into 8
// classes (major axis, sign along the major axis, sign along the minor axis; a zero minor extent has sign +) and sorted by
// slope inside each class, so the beams that can cross a RECTANGLE of cells form one contiguous range of a class list.
// A warp owns a 32 x 32 patch: for an x-major class every lane takes one COLUMN (major offset a) and all lanes run over the
// same range of beams; beam i lands in row k_i(a) of the lane's column -- computed directly, the division being an exact
// multiplication by a per-beam reciprocal -- and the lane adds 1 to its own counter of that row: no atomics, no conflicts
// (y-major classes: lane = row).  Counter additions commute (ray_core.h), so `visited += count` replaces `count` atomics; cells
// that need the ordered replay (hit cells of this scan, distance-map obstacles) look up the runs of the class lists that cross
// them and replay their touches in beam order.
#pragma once

#include "ray_core.h"

namespace lama_b200 {

// ---- beam classes and the slope order -------------------------------------------------------------------------------------
//   cls = [y-major : 4][negative direction along the major axis : 2][negative direction along the minor axis : 1]
struct PullBeam {
    uint32_t n, d;   // major / minor length
    int cls;
};
LAMA_HD PullBeam pull_classify(int ex, int ey)
{
    const uint32_t ax = (uint32_t)(ex < 0 ? -ex : ex), ay = (uint32_t)(ey < 0 ? -ey : ey);
    const bool xmajor = ax >= ay;
    PullBeam b;
    b.n   = xmajor ? ax : ay;
    b.d   = xmajor ? ay : ax;
    b.cls = (xmajor ? 0 : 4) | ((xmajor ? ex < 0 : ey < 0) ? 2 : 0) | ((xmajor ? ey < 0 : ex < 0) ? 1 : 0);
    return b;
}
// entry of a class list (cell offsets inside a directory window are < 2^12, so n, d < 4096)
struct PullEntry {
    uint32_t nd;      // n | d << 16
    uint32_t beam;
    uint64_t magic;   // floor(x / (2 n)) == (x * magic) >> 38  for every x < 2^25
};
LAMA_HD uint32_t pull_pack(uint32_t n, uint32_t d) { return n | (d << 16); }
// Division by the invariant 2 n as multiplication (Granlund & Montgomery): with D = 2 n <= 2^13 and magic = floor(2^38 / D) + 1,
// x * magic / 2^38 = x / D + x e / (D 2^38) with 0 < e <= D, and x e < 2^25 2^13 = 2^38 keeps the excess below 1 / D: the floor is
// exact for all x < 2^25 (x = 2 a d + n <= 2 * 4095 * 4095 + 4095 < 2^25; the product stays below 2^62).  Checked exhaustively on
// the host for every n (tests/emu).
LAMA_HD uint64_t pull_magic(uint32_t n) { return n ? ((1ull << 38) / (2ull * n) + 1ull) : 0ull; }
// k(a): minor offset of the beam (n, d) after a steps
LAMA_HD uint32_t pull_minor_at(uint32_t nd, uint64_t magic, uint32_t a)
{
    const uint32_t n = nd & 0xFFFFu, d = nd >> 16;
    return (uint32_t)(((uint64_t)(2u * a * d + n) * magic) >> 38);
}
// Sort key: class, then slope, then beam.  Two different slopes d1/n1 != d2/n2 with n < 2^13 differ by more than 2^-26, so
// floor(d 2^39 / n) orders them strictly; equal slopes get equal fixed-point values (and are then ordered by beam).
LAMA_HD uint64_t pull_sort_key(int cls, uint32_t n, uint32_t d, uint32_t beam)
{
    const uint64_t slope = n ? (((uint64_t)d << 39) / n) : 0u;   // <= 2^39
    return ((uint64_t)cls << 58) | (slope << 16) | (beam & 0xFFFFu);
}
LAMA_HD uint32_t pull_key_beam(uint64_t key) { return (uint32_t)key & 0xFFFFu; }
LAMA_HD int pull_key_class(uint64_t key) { return (int)(key >> 58); }

// does a beam (n, d) reach past / below the run of cell (a, b)?
LAMA_HD bool pull_below_upper(uint32_t nd, uint32_t a, uint32_t b)   // slope < (2 b + 1) / (2 a)
{
    const uint32_t n = nd & 0xFFFFu, d = nd >> 16;
    return 2u * a * d + n < 2u * (b + 1u) * n;
}
LAMA_HD bool pull_at_least_lower(uint32_t nd, uint32_t a, uint32_t b)   // slope >= (2 b - 1) / (2 a)
{
    const uint32_t n = nd & 0xFFFFu, d = nd >> 16;
    return 2u * a * d + n >= 2u * b * n;
}
// first entry of list[lo, hi) whose slope is >= the lower bound of cell (a, b)
LAMA_HD int pull_lower_bound(const PullEntry* list, int lo, int hi, uint32_t a, uint32_t b)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pull_at_least_lower(list[mid].nd, a, b)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
// ---- the count tile of a patch, one class at a time -----------------------------------------------------------------------
// The beams of class `cls` that can cross the rectangle [a_lo, a_hi] x [b_lo, b_hi] (major x minor offsets, all >= 0, a_lo >= 1):
// every such beam has a slope in [(2 b_lo - 1) / (2 a_hi), (2 b_hi + 1) / (2 a_lo)).
LAMA_HD void pull_class_range(const PullEntry* list, const int* prefix, int cls, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, int& lo, int& hi)
{
    lo = pull_lower_bound(list, prefix[cls], prefix[cls + 1], a_hi, b_lo);
    hi = pull_lower_bound(list, lo, prefix[cls + 1], a_lo, b_hi + 1u);
}
// Where one beam of a class with minor sign `minor_negative` lands on the lane's grid line: the lane sits at major offset a (>= 1)
// and its line covers the signed minor offsets t0 .. t0 + 31; returns the position 0 .. 31 on the line, or -1 when the beam does
// not touch it (too short, or its cell lies outside the patch).
LAMA_HD int pull_land(const PullEntry& e, uint32_t a, bool minor_negative, int t0)
{
    if ((e.nd & 0xFFFFu) <= a) return -1;
    const int k = (int)pull_minor_at(e.nd, e.magic, a);
    const int pos = (minor_negative ? -k : k) - t0;
    return (unsigned)pos < (unsigned)kPatchLen ? pos : -1;
}
// The four classes of one major axis on a patch whose lanes sit at the signed major offsets m0 .. m0 + 31 and whose lines cover the
// signed minor offsets t0 .. t0 + 31: calls visit(cls, major_negative, minor_negative, lo, hi) for every class with a non-empty
// beam range.  (Lanes on the other side of the origin along the major axis, and the lane at major offset 0, sit a class out.)
template <typename Visit>
LAMA_HD void pull_patch_classes(const PullEntry* list, const int* prefix, int base, int m0, int t0, Visit&& visit)
{
    const int m1 = m0 + kPatchLen - 1, t1 = t0 + kPatchLen - 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1   // one copy of the visitor's beam loop in the kernel, not four
#endif
    for (int mneg = 0; mneg < 2; ++mneg) {
        // major offsets a >= 1 of the lanes on this side of the origin
        int a_lo, a_hi;
        if (!mneg) { a_lo = m0 > 1 ? m0 : 1; a_hi = m1; }
        else { a_lo = m1 < -1 ? -m1 : 1; a_hi = -m0; }
        if (a_hi < a_lo) continue;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int tneg = 0; tneg < 2; ++tneg) {
            int b_lo, b_hi;   // |minor offsets| of the lines on this side (the axis line t = 0 belongs to both)
            if (!tneg) { b_lo = t0 > 0 ? t0 : 0; b_hi = t1; }
            else { b_lo = t1 < 0 ? -t1 : 0; b_hi = -t0; }
            if (b_hi < b_lo || b_lo > a_hi) continue;   // beyond the diagonal no beam of this major axis passes
            const int cls = base | (mneg ? 2 : 0) | (tneg ? 1 : 0);
            int lo, hi;
            pull_class_range(list, prefix, cls, (uint32_t)a_lo, (uint32_t)a_hi, (uint32_t)b_lo, (uint32_t)b_hi, lo, hi);
            if (hi > lo) visit(cls, mneg != 0, tneg != 0, lo, hi);
        }
    }
}

// ---- the runs of ONE cell (ordered replay of candidate cells) -----------------------------------------------------------------
// Four runs [lo, hi) of the sorted list (x-major / y-major beams, positive / negative minor side; empty when not applicable)
// hold the beams whose slope passes through cell (cx, cy) (offsets from O); a beam of run r really crosses the cell iff its
// major length exceeds a[r], and it does so at step a[r] of its walk.  (Fixed slots: the loops over them unroll into registers.)
struct PullRuns {
    int lo[4], hi[4];
    uint32_t a[4];
};
LAMA_HD void pull_set_run(PullRuns& r, int slot, bool on, const PullEntry* list, const int* prefix, int cls, uint32_t a, uint32_t b)
{
    r.lo[slot] = r.hi[slot] = 0;
    r.a[slot]  = a;
    if (!on) return;
    r.lo[slot] = pull_lower_bound(list, prefix[cls], prefix[cls + 1], a, b);
    r.hi[slot] = pull_lower_bound(list, r.lo[slot], prefix[cls + 1], a, b + 1u);
}
LAMA_HD PullRuns pull_cell_runs(const PullEntry* list, const int* prefix, int cx, int cy)
{
    PullRuns r;
    const uint32_t ax = (uint32_t)(cx < 0 ? -cx : cx), ay = (uint32_t)(cy < 0 ? -cy : cy);
    const bool xm = ax >= 1 && ay <= ax, ym = ay >= 1 && ax <= ay;   // x-major / y-major beams can pass here
    const int cmx = cx < 0 ? 2 : 0, cmy = 4 | (cy < 0 ? 2 : 0);
    pull_set_run(r, 0, xm && cy >= 0, list, prefix, cmx, ax, ay);
    pull_set_run(r, 1, xm && cy <= 0, list, prefix, cmx | 1, ax, ay);
    pull_set_run(r, 2, ym && cx >= 0, list, prefix, cmy, ay, ax);
    pull_set_run(r, 3, ym && cx <= 0, list, prefix, cmy | 1, ay, ax);
    return r;
}

// ---- which patches can a beam touch? ----------------------------------------------------------------------------------------
// Calls mark(px, py) (window-relative patch coordinates) for a superset of the patches that hold interior cells of the beam
// from O = (ox, oy) to O + (ex, ey): per 32-cell block along the major axis the minor coordinate is monotone and moves by at
// most 31 cells, so the patches of the block's first and last step cover it.
template <typename Mark>
LAMA_HD void pull_mark_beam(uint32_t ox, uint32_t oy, int ex, int ey, Mark&& mark)
{
    const PullBeam pb = pull_classify(ex, ey);
    if (pb.n < 2) return;
    const bool ymajor = (pb.cls & 4) != 0;
    const int sM = (pb.cls & 2) ? -1 : 1, sm = (pb.cls & 1) ? -1 : 1;
    const int oM = (int)(ymajor ? oy : ox), om = (int)(ymajor ? ox : oy);
    const int n = (int)pb.n, d = (int)pb.d;
    int i = 1;
    while (i <= n - 1) {
        const int M0 = oM + sM * i, blk = M0 >> kPatchLog2;
        const int Mend = sM > 0 ? ((blk << kPatchLog2) + kPatchLen - 1 < oM + (n - 1) ? (blk << kPatchLog2) + kPatchLen - 1 : oM + (n - 1))
                                : ((blk << kPatchLog2) > oM - (n - 1) ? (blk << kPatchLog2) : oM - (n - 1));
        const int iend = sM > 0 ? Mend - oM : oM - Mend;
        const int k0 = (int)((2u * (uint32_t)i * (uint32_t)d + (uint32_t)n) / (2u * (uint32_t)n));
        const int k1 = (int)((2u * (uint32_t)iend * (uint32_t)d + (uint32_t)n) / (2u * (uint32_t)n));
        const int p0 = (om + sm * k0) >> kPatchLog2, p1 = (om + sm * k1) >> kPatchLog2;
        if (ymajor) {
            mark(p0, blk);
            if (p1 != p0) mark(p1, blk);
        } else {
            mark(blk, p0);
            if (p1 != p0) mark(blk, p1);
        }
        i = iend + 1;
    }
}

// ---- hit records ----------------------------------------------------------------------------------------------------------------
// [cell index in the patch : 10 (bits 16..25)][beam : 16]; k_ray_setup groups them by patch (each task names its range), in no
// particular order inside a patch: the replay picks a cell's hits out of the patch's few records by their cell index.
LAMA_HD uint32_t pull_hit_record(uint32_t cell, uint32_t beam) { return (cell << 16) | (beam & 0xFFFFu); }
LAMA_HD uint32_t pull_hit_cell(uint32_t rec) { return rec >> 16; }

// ---- ordered replay of one candidate cell ---------------------------------------------------------------------------------------
// Touches = the crossing beams of the runs (misses) and those of the patch's hit records hits[h0, h1) that name cell `ci`; a beam
// touches a cell at most once.  They are visited in beam order by repeated minimum selection (the lists are short), starting from the cell's
// counters BEFORE the scan; returns the counters after the scan and the new obstacle-mirror bit, and emits the obstacle events
// exactly like replay_cell (ray_core.h).
//   setFree  frequency_occupancy_map.cpp:65-74   setOccupied :81-91   addObstacle / removeObstacle dynamic_distance_map.cpp:212-242
struct PullTouch {
    uint32_t beam, pos;
    bool hit, valid;
};
// the hit records of one cell, picked out of its patch's records once (a cell rarely holds more than a few hits; beyond kMax the
// replay falls back to scanning the patch's records every time)
struct PullCellHits {
    static constexpr int kMax = 4;
    uint32_t beam[kMax];
    int n;   // > kMax: too many for the registers
};
LAMA_HD PullCellHits pull_cell_hits(const uint32_t* hits, int h0, int h1, uint32_t ci)
{
    PullCellHits c;
    c.n = 0;
    for (int k = 0; k < PullCellHits::kMax; ++k) c.beam[k] = 0xFFFFFFFFu;
    for (int i = h0; i < h1; ++i) {
        if (pull_hit_cell(hits[i]) != ci) continue;
        const uint32_t b = hits[i] & 0xFFFFu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < PullCellHits::kMax; ++k)
            if (c.n == k) c.beam[k] = b;
        ++c.n;
    }
    return c;
}
LAMA_HD PullTouch pull_next_touch(const PullEntry* list, const PullRuns& runs, const PullCellHits& ch, const uint32_t* hits, int h0, int h1, uint32_t ci, int after)
{
    PullTouch t{0xFFFFFFFFu, 0u, false, false};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 4; ++r)
        for (int i = runs.lo[r]; i < runs.hi[r]; ++i) {
            if ((list[i].nd & 0xFFFFu) <= runs.a[r]) continue;   // too short to reach the cell
            const uint32_t b = list[i].beam;
            if ((int)b > after && b < t.beam) {
                t.beam = b; t.pos = runs.a[r]; t.hit = false; t.valid = true;
            }
        }
    if (ch.n <= PullCellHits::kMax) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < PullCellHits::kMax; ++k) {
            const uint32_t b = ch.beam[k];
            if (k < ch.n && (int)b > after && b < t.beam) {
                t.beam = b; t.pos = 0u; t.hit = true; t.valid = true;
            }
        }
        return t;
    }
    for (int i = h0; i < h1; ++i) {
        if (pull_hit_cell(hits[i]) != ci) continue;
        const uint32_t b = hits[i] & 0xFFFFu;
        if ((int)b > after && b < t.beam) {
            t.beam = b; t.pos = 0u; t.hit = true; t.valid = true;
        }
    }
    return t;
}
template <typename Emit>
LAMA_HD uint32_t pull_replay_cell(const PullEntry* list, const PullRuns& runs, const uint32_t* hits, int h0, int h1, uint32_t ci, uint32_t word,
                                  bool& obstacle, Emit&& emit)
{
    uint32_t occupied = occ_occupied(word), visited = occ_visited(word);
    const PullCellHits ch = pull_cell_hits(hits, h0, h1, ci);
    int after = -1;
    for (;;) {
        const PullTouch t = pull_next_touch(list, runs, ch, hits, h0, h1, ci, after);
        if (!t.valid) break;
        after = (int)t.beam;
        const uint32_t seq = (t.beam << 15) | (t.pos & 0x7FFFu);   // == log_seq(log_record(., beam, pos, hit))
        if (t.hit) {
            const bool was_occupied = occ_is_occupied(occupied, visited);
            occupied = (occupied + 1) & 0xFFFFu;
            visited  = (visited + 1) & 0xFFFFu;
            if (!was_occupied && occ_is_occupied(occupied, visited) && !obstacle) {
                obstacle = true;
                emit(true, seq);
            }
        } else {
            const bool was_free = occ_is_free(occupied, visited);
            visited = (visited + 1) & 0xFFFFu;
            if (!was_free && occ_is_free(occupied, visited) && obstacle) {
                obstacle = false;
                emit(false, seq);
            }
        }
    }
    return (visited << 16) | occupied;
}
// log-odds cells (ProbabilisticOccupancyMap, probabilistic_occupancy_map.cpp:82-107)
template <typename Emit>
LAMA_HD float pull_replay_cell_prob(const PullEntry* list, const PullRuns& runs, const uint32_t* hits, int h0, int h1, uint32_t ci, float prob,
                                    bool& obstacle, const ProbParams& pp, Emit&& emit)
{
    const PullCellHits ch = pull_cell_hits(hits, h0, h1, ci);
    int after = -1;
    for (;;) {
        const PullTouch t = pull_next_touch(list, runs, ch, hits, h0, h1, ci, after);
        if (!t.valid) break;
        after = (int)t.beam;
        const uint32_t seq = (t.beam << 15) | (t.pos & 0x7FFFu);
        if (t.hit) {
            const bool was_occupied = (double)prob > pp.thresh;
            prob = prob_hit(prob, pp);
            if (!was_occupied && (double)prob > pp.thresh && !obstacle) {
                obstacle = true;
                emit(true, seq);
            }
        } else {
            const bool was_free = (double)prob < pp.thresh;
            prob = prob_miss(prob, pp);
            if (!was_free && (double)prob < pp.thresh && obstacle) {
                obstacle = false;
                emit(false, seq);
            }
        }
    }
    return prob;
}

}  // namespace lama_b200
