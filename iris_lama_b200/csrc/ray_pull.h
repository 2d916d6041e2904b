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
// the walk sits at major offset i and minor offset k = floor((2 i d + n) / (2 n))  (closed form of the error accumulator,
// ray_core.h SegWalk).  Hence the beam crosses the cell at major offset a, minor offset b  iff
//        1 <= a <= n - 1   and   2 b n <= 2 a d + n < 2 (b + 1) n,
// i.e. iff its slope d / n lies in [(2 b - 1) / (2 a), (2 b + 1) / (2 a)) and it is long enough.  Beams are split into 8
// classes (major axis, sign along the major axis, sign along the minor axis; a zero minor extent has sign +) and sorted by
// slope inside each class; the beams that cross a cell then form ONE contiguous run of a class list (two runs when the cell
// lies on an axis through O), and the runs of the cells of one grid line abut: walking a line of the patch in the direction
// of growing b consumes the class list front to back ("chain"), one comparison per crossing beam.
// Counter additions commute (ray_core.h), so `visited += count` replaces `count` atomics; cells that need the ordered replay
// (hit cells of this scan, distance-map obstacles) enumerate their runs and replay their touches in beam order.
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
// entry of a class list: n | d << 16 (cell offsets inside a directory window are < 2^13)
LAMA_HD uint32_t pull_pack(uint32_t n, uint32_t d) { return n | (d << 16); }
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
LAMA_HD int pull_lower_bound(const uint32_t* list, int lo, int hi, uint32_t a, uint32_t b)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pull_at_least_lower(list[mid], a, b)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
// One chain step: `ptr` stands on the first entry whose slope is >= the lower bound of cell (a, b); consumes the run of the
// cell and returns how many of its beams are long enough to reach it.  Afterwards `ptr` stands at the lower bound of (a, b+1).
LAMA_HD uint32_t pull_step(const uint32_t* list, int& ptr, int end, uint32_t a, uint32_t b)
{
    uint32_t count = 0;
    while (ptr < end) {
        const uint32_t nd = list[ptr];
        if (!pull_below_upper(nd, a, b)) break;
        count += (nd & 0xFFFFu) > a ? 1u : 0u;
        ++ptr;
    }
    return count;
}

// ---- one lane's share of an axis pass over a 32 x 32 patch ------------------------------------------------------------------
// The X pass counts the crossings of x-major beams: a lane owns a COLUMN (major offset a_signed = cx) and walks the 32 rows
// of the patch (minor offsets t0 .. t0 + 31 = cy); the Y pass swaps the roles (lane = row, lines = columns) for the y-major
// beams.  `prefix[c] .. prefix[c + 1]` delimit class c in `list`; `base` is 0 (X pass) or 4 (Y pass).
// `out(line, count)` is called for the lines this lane has anything to say about (count may be 0); other lines are untouched.
template <typename Out>
LAMA_HD void pull_lane_pass(const uint32_t* list, const int* prefix, int a_signed, int t0, int base, Out&& out)
{
    const uint32_t a = (uint32_t)(a_signed < 0 ? -a_signed : a_signed);
    if (a == 0) return;   // the walk never visits major offset 0 (the origin's own line)
    const int t1 = t0 + kPatchLen - 1;
    // smallest |t| of the patch: beyond the diagonal (b > a) no beam of this major axis passes
    const uint32_t bmin = (t0 <= 0 && t1 >= 0) ? 0u : (uint32_t)(t0 > 0 ? t0 : -t1);
    if (bmin > a) return;
    const int cm = base | (a_signed < 0 ? 2 : 0);
    uint32_t c0neg = 0;
    if (t0 <= 0) {   // lines on the negative minor side: class cm | 1, walked away from the axis (b = -t grows)
        const int lo = prefix[cm | 1], hi = prefix[(cm | 1) + 1];
        const int tstart = t1 < -1 ? t1 : -1;
        uint32_t b = (uint32_t)(-tstart);
        int ptr;
        if (t1 >= 0) {   // the patch holds the axis line t = 0: its cell also collects the negative-side beams with b = 0
            ptr   = lo;
            c0neg = pull_step(list, ptr, hi, a, 0u);
        } else {
            ptr = pull_lower_bound(list, lo, hi, a, b);
        }
        for (int t = tstart; t >= t0 && b <= a; --t, ++b) out(t - t0, pull_step(list, ptr, hi, a, b));
    }
    if (t1 >= 0) {
        const int lo = prefix[cm], hi = prefix[cm + 1];
        const int tstart = t0 > 0 ? t0 : 0;
        uint32_t b = (uint32_t)tstart;
        int ptr = b == 0 ? lo : pull_lower_bound(list, lo, hi, a, b);
        for (int t = tstart; t <= t1 && b <= a; ++t, ++b) {
            uint32_t c = pull_step(list, ptr, hi, a, b);
            if (t == 0) c += c0neg;
            out(t - t0, c);
        }
    }
}

// The same pass as ONE loop per lane (what the kernel runs): every iteration either consumes a beam of the class list or closes a
// cell and moves to the next line, so the lanes of a warp -- whose columns hold different numbers of beams per cell -- do not wait
// for each other at every line, only at the end of the pass.  `emit(b, count)` closes cell b.
template <typename EmitCell>
LAMA_HD void pull_lane_chain(const uint32_t* list, int ptr, int end, uint32_t a, uint32_t b, uint32_t b_last, EmitCell&& emit)
{
    const uint32_t a2 = 2u * a;
    uint32_t bb = 2u * (b + 1u), count = 0;
    uint32_t n = 1u, d = 0xFFFFu;   // past the end of the list: a beam steeper than every cell bound
    if (ptr < end) { n = list[ptr] & 0xFFFFu; d = list[ptr] >> 16; }
    for (;;) {
        if (a2 * d + n < bb * n) {   // the beam's slope is below the upper bound of cell b: it belongs to this cell's run
            count += n > a ? 1u : 0u;
            ++ptr;
            n = 1u; d = 0xFFFFu;
            if (ptr < end) { n = list[ptr] & 0xFFFFu; d = list[ptr] >> 16; }
        } else {
            emit(b, count);
            if (b == b_last) break;
            count = 0;
            ++b;
            bb += 2u;
        }
    }
}
template <typename Out>
LAMA_HD void pull_lane_pass_flat(const uint32_t* list, const int* prefix, int a_signed, int t0, int base, Out&& out)
{
    const uint32_t a = (uint32_t)(a_signed < 0 ? -a_signed : a_signed);
    if (a == 0) return;
    const int t1 = t0 + kPatchLen - 1;
    const uint32_t bmin = (t0 <= 0 && t1 >= 0) ? 0u : (uint32_t)(t0 > 0 ? t0 : -t1);
    if (bmin > a) return;
    const int cm = base | (a_signed < 0 ? 2 : 0);
    uint32_t c0neg = 0;
    if (t0 <= 0) {   // negative minor side (class cm | 1), away from the axis; with the axis line in the patch the chain starts at b = 0
        const int lo = prefix[cm | 1], hi = prefix[(cm | 1) + 1];
        const uint32_t b0 = t1 >= 0 ? 0u : (uint32_t)(-t1);
        uint32_t b1 = (uint32_t)(-t0);
        if (b1 > a) b1 = a;
        const int ptr = b0 == 0 ? lo : pull_lower_bound(list, lo, hi, a, b0);
        pull_lane_chain(list, ptr, hi, a, b0, b1, [&](uint32_t b, uint32_t c) {
            if (b == 0) c0neg = c;   // kept for the axis cell, which the positive side emits
            else out(-(int)b - t0, c);
        });
    }
    if (t1 >= 0) {
        const int lo = prefix[cm], hi = prefix[cm + 1];
        const uint32_t b0 = t0 > 0 ? (uint32_t)t0 : 0u;
        uint32_t b1 = (uint32_t)t1;
        if (b1 > a) b1 = a;
        const int ptr = b0 == 0 ? lo : pull_lower_bound(list, lo, hi, a, b0);
        pull_lane_chain(list, ptr, hi, a, b0, b1, [&](uint32_t b, uint32_t c) { out((int)b - t0, b == 0 ? c + c0neg : c); });
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
LAMA_HD void pull_set_run(PullRuns& r, int slot, bool on, const uint32_t* list, const int* prefix, int cls, uint32_t a, uint32_t b)
{
    r.lo[slot] = r.hi[slot] = 0;
    r.a[slot]  = a;
    if (!on) return;
    r.lo[slot] = pull_lower_bound(list, prefix[cls], prefix[cls + 1], a, b);
    r.hi[slot] = pull_lower_bound(list, r.lo[slot], prefix[cls + 1], a, b + 1u);
}
LAMA_HD PullRuns pull_cell_runs(const uint32_t* list, const int* prefix, int cx, int cy)
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
LAMA_HD PullTouch pull_next_touch(const uint32_t* list, const uint16_t* beam_of, const PullRuns& runs, const uint32_t* hits, int h0, int h1, uint32_t ci, int after)
{
    PullTouch t{0xFFFFFFFFu, 0u, false, false};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 4; ++r)
        for (int i = runs.lo[r]; i < runs.hi[r]; ++i) {
            if ((list[i] & 0xFFFFu) <= runs.a[r]) continue;   // too short to reach the cell
            const uint32_t b = beam_of[i];
            if ((int)b > after && b < t.beam) {
                t.beam = b; t.pos = runs.a[r]; t.hit = false; t.valid = true;
            }
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
LAMA_HD uint32_t pull_replay_cell(const uint32_t* list, const uint16_t* beam_of, const PullRuns& runs, const uint32_t* hits, int h0, int h1, uint32_t ci, uint32_t word,
                                  bool& obstacle, Emit&& emit)
{
    uint32_t occupied = occ_occupied(word), visited = occ_visited(word);
    int after = -1;
    for (;;) {
        const PullTouch t = pull_next_touch(list, beam_of, runs, hits, h0, h1, ci, after);
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
LAMA_HD float pull_replay_cell_prob(const uint32_t* list, const uint16_t* beam_of, const PullRuns& runs, const uint32_t* hits, int h0, int h1, uint32_t ci, float prob,
                                    bool& obstacle, const ProbParams& pp, Emit&& emit)
{
    int after = -1;
    for (;;) {
        const PullTouch t = pull_next_touch(list, beam_of, runs, hits, h0, h1, ci, after);
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
