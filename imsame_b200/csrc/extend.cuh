// extend.cuh -- ungapped X-drop-to-zero extension of one seed hit on 2-bit
// packed sequences.  extend_hit is the plain base-by-base form (what the tests and the
// table builders are checked against); the scan kernel uses the table-driven window form
// further down.  Restates src/alignmentFunctions.c:276-359 in units of
// POINT (score 12 = 48/4, +-1 per base):
//   forward from (p, e+1) while score > 0 inside both reads; idents++ on match;
//   `if (high_right <= score) end = cur`  (:324);
//   backward from (p-13, e-12) starting at score = high_right while
//   high_left stays at its initial 12 (:303,339); `if (high_left <= score) start = cur` (:347);
//   t_len = end - start (:359);  returns n = 2*idents - t_len  (raw score = 4n, :373).
// The seed length is a parameter `K` of every function here (default: the reference's FIXED_K = 12,
// src/structs.h:15); other values are the reference "recompiled with another FIXED_K" (SURVEY 8(f) rank 4).
// p = database index AFTER the seed's last base (llpos.pos), e = query index of
// the seed's last base (curr_pos).  [xs, xend) and [ys, yend) are the reads.
#pragma once
#include "common.cuh"

namespace imsame {

IMS_HD int extend_hit(const uint32_t *dpk, const uint32_t *qpk, uint32_t p, uint32_t e, uint32_t xs,
                      uint32_t xend, uint32_t ys, uint32_t yend, int K = imsame::K) {
    int fmax = (int)(xend - p);
    {
        const int fq = (int)(yend - (e + 1));
        fmax = fq < fmax ? fq : fmax;
    }
    int sc = K, hr = K, idn = K, fe = -1;
    for (int t = 0; t < fmax && sc > 0; t += 32) {
        const uint32_t mm = mismatch32(fetch32(dpk, (uint64_t)p + t), fetch32(qpk, (uint64_t)e + 1 + t));
        const int lim = (fmax - t) < 32 ? (fmax - t) : 32;
        for (int u = 0; u < lim && sc > 0; u++) {
            const int m = (mm >> u) & 1;
            sc += 1 - 2 * m;
            idn += 1 - m;
            if (hr <= sc) { fe = t + u; hr = sc; }
        }
    }
    int bmax = (int)(p - K - xs);
    {
        const int bq = (int)e - (K - 1) - (int)ys; /* -1 for the cross-read "phantom" word */
        bmax = bq < bmax ? bq : bmax;
    }
    sc = hr;
    int hl = K, bsteps = 0;
    for (int t = 0; t < bmax && sc > 0; t += 32) {
        const int64_t sd = (int64_t)p - (K + 1) - t - 31, sq = (int64_t)e - K - t - 31;
        const uint64_t a = sd >= 0 ? fetch32(dpk, (uint64_t)sd) : (fetch32(dpk, 0) << (2 * (int)(-sd)));
        const uint64_t b = sq >= 0 ? fetch32(qpk, (uint64_t)sq) : (fetch32(qpk, 0) << (2 * (int)(-sq)));
        const uint32_t mm = mismatch32(a, b);
        const int lim = (bmax - t) < 32 ? (bmax - t) : 32;
        for (int u = 0; u < lim && sc > 0; u++) {
            const int m = (mm >> (31 - u)) & 1;
            sc += 1 - 2 * m;
            idn += 1 - m;
            if (hl <= sc) { bsteps = t + u + 1; hl = sc; }
        }
    }
    return 2 * idn - (fe + K + bsteps);
}

}  // namespace imsame

// ---------------------------------------------------------------------------------------
namespace imsame {

IMS_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

}  // namespace imsame

// ---------------------------------------------------------------------------------------
// Window-at-a-time form for the scan kernel: the walk of one hit is a small state machine
// that consumes one 32-base window (forward or backward) per call, four table steps with
// no data-dependent branch, so that the 32 hits of a warp stay converged: every lane
// executes the same window body until all walks of the warp are finished.
//
// Everything is 32-bit: a window is two 16-base halves cut out of three packed words with
// funnel shifts, the 2-bit differences are reduced to one mismatch bit per base per half
// (8 shift/logic pairs) and the walk table carries the score change of its 8 steps, so no
// 64-bit arithmetic and no population count is left in the loop (the first version spent
// 240 instructions per window, two thirds of them on 64-bit address/shift/compress work).
namespace imsame {

// Walk table of the scan kernel: one 32-bit entry per (entry score row, 8 mismatch bits).
// The walk state is ONE word  run = (score << 15) | steps_done  and the running maximum of
// the reference's `if (high <= score) { pos = cur; high = score; }` (:324,347) is ONE word
// best = ((high + 1) << 15) | (1-based step of its last occurrence): later steps have larger step
// numbers, so "replace when >=" is a plain integer maximum of such keys.  Per 8 steps:
//     best = max(best, run + X(entry));   run += Y(entry) - (8 << 15);
//   X = ((max prefix score relative to the entry score + 1) << 15) + step of its last occurrence
//   Y = ((score change + 8) << 15) + steps executed
// Both come out of the entry with one rotate + two masks (fields: steps bits 0-3, step of the
// maximum bits 8-11, change + 8 bits 15-19, maximum + 1 bits 23-26).  The first version kept X and Y
// as a 64-bit pair: 32 lanes reading random 8-byte entries cost 6.4 shared-memory wavefronts per
// lookup against 3.5 for 4-byte entries, and the scan kernel runs at 92 % of the L1 data pipe.
// When the score reaches 0 inside the 8 steps the walk ends (:318,340): the entry then sets the
// score to exactly 0 and counts only the executed steps, and row 0 is absorbing (X never wins,
// Y changes nothing), so a finished walk needs no predicate.  Row 9 = "9 or more".
// Matches are not counted: every step is +-1, so matches = (steps + score_end - score_start) / 2.
constexpr int EXT_ROWS3 = 10;
constexpr int EXT_LUT3_SIZE = EXT_ROWS3 * 256;  // uint32 entries
constexpr int EXT_SC_SHIFT = 15;                // reads of up to 32767 bases
constexpr uint32_t EXT_POS_MASK = (1u << EXT_SC_SHIFT) - 1u;
constexpr uint32_t EXT_MAX_READ = EXT_POS_MASK;
constexpr int EXT_BEST_BIAS = 1 << EXT_SC_SHIFT;  // `best` words carry high + 1
constexpr uint32_t EXT_X_MASK = 0x0007800Fu, EXT_Y_MASK = 0x000F800Fu;

inline void build_ext_lut3(uint32_t *lut /* EXT_LUT3_SIZE */) {
    for (int row = 0; row < EXT_ROWS3; row++)
        for (uint32_t m = 0; m < 256; m++) {
            int rmax1 = 0, amax = 0, delta8 = 8, steps = 0;  // row 0: absorbing
            if (row > 0) {
                int run = 0, rmax = -100;
                bool term = false;
                for (int t = 0; t < 8; t++) {
                    run += ((m >> t) & 1) ? -1 : 1;
                    steps = t + 1;
                    if (run >= rmax) { rmax = run; amax = t + 1; }
                    if (row < 9 && row + run <= 0) { term = true; break; }
                }
                rmax1 = rmax + 1;
                delta8 = (term ? -row : run) + 8;
            }
            lut[row * 256 + m] = (uint32_t)steps | ((uint32_t)amax << 8) | ((uint32_t)delta8 << 15) | ((uint32_t)rmax1 << 23);
        }
}

IMS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

// even bits of x -> low 16 bits
IMS_HD uint32_t compress_even16(uint32_t x) {
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    return (x | (x >> 8)) & 0x0000FFFFu;
}

// 32 bases starting at base i as two 16-base halves.  The three packed words they span are fetched
// as two aligned 64-bit loads (4 words) instead of three 32-bit ones: the scan kernel is bound by the
// L1 data pipe, where every scattered load instruction costs one wavefront per lane whatever its width.
IMS_HD void window_halves(const uint32_t *pk, uint32_t i, uint32_t &lo, uint32_t &hi) {
    const uint32_t base = (i >> 4) & ~1u;
    const unsigned s = (i & 31u) * 2u;  // bit offset inside the 128 bits loaded
#if defined(__CUDA_ARCH__)
    const uint2 q0 = *reinterpret_cast<const uint2 *>(pk + base);
    const uint2 q1 = *reinterpret_cast<const uint2 *>(pk + base + 2);
    const uint32_t w0 = q0.x, w1 = q0.y, w2 = q1.x, w3 = q1.y;
#else
    const uint32_t w0 = pk[base], w1 = pk[base + 1], w2 = pk[base + 2], w3 = pk[base + 3];
#endif
    const bool up = s >= 32u;
    const uint32_t a = up ? w1 : w0, b = up ? w2 : w1, c = up ? w3 : w2;
    lo = funnel_r(a, b, s & 31u);
    hi = funnel_r(b, c, s & 31u);
}

// mismatch bits (bit t <=> bases differ) of the 32-base windows starting at base i of a and j of b
IMS_HD uint32_t window_mismatch(const uint32_t *apk, uint32_t i, const uint32_t *bpk, uint32_t j) {
    uint32_t a0, a1, b0, b1;
    window_halves(apk, i, a0, a1);
    window_halves(bpk, j, b0, b1);
    uint32_t d0 = a0 ^ b0, d1 = a1 ^ b1;
    d0 = (d0 | (d0 >> 1)) & 0x55555555u;
    d1 = (d1 | (d1 >> 1)) & 0x55555555u;
    return compress_even16(d0) | (compress_even16(d1) << 16);
}

struct ExtState {
    int phase;       // 0 forward, 1 backward, 2 done
    int t;           // steps already taken in this phase (multiple of 32)
    int run;         // (score << 15) | steps done in this phase          (units of POINT)
    int best;        // (high << 15) | 1-based step that last reached it  (:324-327, :347-351)
    int pos_f;       // forward result: 1-based step of high_right (fe + 1)
    int idn2;        // 2 * matches walked so far (both phases, without the seed)
    int fmax, bmax;  // steps available inside both reads
};

IMS_HD void ext_init(ExtState &s, uint32_t p, uint32_t e, uint32_t xs, uint32_t xend, uint32_t ys, uint32_t yend,
                     int K = imsame::K) {
    const int fd = (int)(xend - p), fq = (int)(yend - (e + 1));
    const int bd = (int)(p - K - xs), bq = (int)e - (K - 1) - (int)ys;  // -1 for the phantom word
    s.fmax = fq < fd ? fq : fd;
    s.bmax = bq < bd ? bq : bd;
    s.t = 0;
    s.run = K << EXT_SC_SHIFT;   // score = 12 (48 / POINT), :300
    s.best = (K << EXT_SC_SHIFT) + EXT_BEST_BIAS;  // high_right = 12, no step yet
    s.pos_f = 0;
    s.idn2 = 0;
    s.phase = s.fmax > 0 ? 0 : (s.bmax > 0 ? 1 : 2);
}

// one table step (8 bases: bits c .. c+7 of the window's mismatch mask) of a walk
IMS_HD void ext_step(const uint32_t *lut, uint32_t mm, int c, int &run, int &best) {
    uint32_t row = (uint32_t)run >> EXT_SC_SHIFT;
    row = row < 9u ? row : 9u;
    const uint32_t e = lut[row * 256u + ((mm >> c) & 0xFFu)];
    const int key = run + (int)(funnel_r(e, e, 8) & EXT_X_MASK);
    best = key > best ? key : best;
    run += (int)(e & EXT_Y_MASK) - (8 << EXT_SC_SHIFT);
}

// The same table with both addends of a step spelled out (8 bytes per entry): the scan kernel's copy in shared
// memory.  A step is then 7 instructions instead of 11 -- shift, clamp, one byte permute that forms the index
// from the mask byte and the row, one 64-bit shared load, one fused add-maximum, one add -- at the price of
// twice the table bytes per lookup; the kernel is bound by instruction issue (round 2 profile), not by the
// shared-memory pipe.
// Shared-memory banks: mismatch bits are 1 three times out of four, so the low bits of a mask byte are mostly
// 1111 and a warp's 32 lookups crowded into a few banks (9.6 wavefronts per lookup measured).  The ExtXY table
// is therefore indexed by the FOLDED byte  b ^ (b >> 4)  (a bijection on 0..255; the XOR of two biased bits is
// nearly fair), applied to all four bytes of a mask at once (ext_prep).
struct ExtXY {
    int x;   // ((maximum prefix score + 1) << 15) + step of its last occurrence
    int yd;  // (score change << 15) + steps executed
};
IMS_HD uint32_t ext_fold8(uint32_t b) { return (b ^ (b >> 4)) & 0xFFu; }
IMS_HD uint32_t ext_prep(const ExtXY *, uint32_t mm) { return mm ^ ((mm >> 4) & 0x0F0F0F0Fu); }
IMS_HD uint32_t ext_prep(const uint32_t *, uint32_t mm) { return mm; }
IMS_HD ExtXY ext_xy_of(uint32_t e) {
    ExtXY v;
    v.x = (int)(funnel_r(e, e, 8) & EXT_X_MASK);
    v.yd = (int)(e & EXT_Y_MASK) - (8 << EXT_SC_SHIFT);
    return v;
}
IMS_HD void ext_step(const ExtXY *lut, uint32_t mm, int c, int &run, int &best) {
    uint32_t row = (uint32_t)run >> EXT_SC_SHIFT;
    row = row < 9u ? row : 9u;
#if defined(__CUDA_ARCH__)
    const uint32_t idx = __byte_perm(mm, row, 0x5540u | (uint32_t)(c >> 3));  // row << 8 | mask byte (row < 256)
    const int2 e = *reinterpret_cast<const int2 *>(lut + idx);
    best = __viaddmax_s32(run, e.x, best);
    run += e.y;
#else
    const ExtXY e = lut[row * 256u + ((mm >> c) & 0xFFu)];
    const int key = run + e.x;
    best = key > best ? key : best;
    run += e.yd;
#endif
}

// first table step of a walk that starts with a score of 9 or more (seed lengths >= 9: the forward walk starts at
// K, the backward one at high_right >= K): the row is known, two instructions less
IMS_HD void ext_step_row9(const ExtXY *lut, uint32_t mm, int &run, int &best) {
#if defined(__CUDA_ARCH__)
    const int2 e = *reinterpret_cast<const int2 *>(lut + ((mm & 0xFFu) | (9u << 8)));
    best = __viaddmax_s32(run, e.x, best);
    run += e.y;
#else
    ext_step(lut, mm, 0, run, best);
#endif
}
IMS_HD void ext_step_row9(const uint32_t *lut, uint32_t mm, int &run, int &best) { ext_step(lut, mm, 0, run, best); }

template <class LUT>
IMS_HD void ext_window(ExtState &s, const LUT *lut, const uint32_t *dpk, const uint32_t *qpk, uint32_t p,
                       uint32_t e, int K = imsame::K) {
    const bool bwd = s.phase == 1;
    const int maxs = bwd ? s.bmax : s.fmax;
    const int rem = maxs - s.t;
    // first base of the window in both sequences
    const int sd = bwd ? (int)p - (K + 1) - s.t - 31 : (int)p + s.t;
    const int sq = bwd ? (int)e - K - s.t - 31 : (int)e + 1 + s.t;
    uint32_t mm;
    if (sd >= 0 && sq >= 0) {
        mm = window_mismatch(dpk, (uint32_t)sd, qpk, (uint32_t)sq);
    } else {
        // a backward window that starts before base 0 of an array (first read only)
        const uint64_t a = sd >= 0 ? fetch32(dpk, (uint64_t)sd) : (fetch32(dpk, 0) << (2 * (-sd)));
        const uint64_t b = sq >= 0 ? fetch32(qpk, (uint64_t)sq) : (fetch32(qpk, 0) << (2 * (-sq)));
        mm = mismatch32(a, b);
    }
    if (bwd) mm = brev32(mm);                    // step u <-> bit u in both directions
    if (rem < 32) mm |= 0xFFFFFFFFu << rem;      // steps past the read end: mismatches (change neither maximum nor matches)
    int run = s.run, best = s.best;
    const int sc0 = run >> EXT_SC_SHIFT;
    mm = ext_prep(lut, mm);
#pragma unroll
    for (int c = 0; c < 32; c += 8) ext_step(lut, mm, c, run, best);
    const int sc1 = run >> EXT_SC_SHIFT;
    // +-1 per step: 2 * matches = steps + score change (also true for the padded steps)
    s.idn2 += ((run - s.run) & (int)EXT_POS_MASK) + sc1 - sc0;
    const bool phase_over = sc1 == 0 || rem <= 32;
    if (!phase_over) {
        s.t += 32; s.run = run; s.best = best;
    } else if (!bwd) {
        s.pos_f = best & (int)EXT_POS_MASK;
        s.phase = s.bmax > 0 ? 1 : 2;
        s.t = 0;
        s.run = (best & ~(int)EXT_POS_MASK) - EXT_BEST_BIAS;  // backward restarts from high_right (:339), step count 0
        s.best = (K << EXT_SC_SHIFT) + EXT_BEST_BIAS;         // high_left = 12 (:303)
    } else {
        s.best = best;
        s.phase = 2;
    }
}

// Mismatch bits of the first forward and the first backward window of a hit (step u <-> bit u in
// both); steps past the read ends are mismatches.  A phase with no room at all (fmax or bmax <= 0)
// is an all-mismatch window: walking it changes no match count, and no maximum either in the forward
// phase (the first step already scores below high_right = the start score); the backward phase starts
// ABOVE high_left = 12 (:303,339), so there a walk without room is discarded instead (ext_first2).
IMS_HD void ext_first_masks(const ExtState &s, const uint32_t *dpk, const uint32_t *qpk, uint32_t p, uint32_t e,
                            uint32_t &mf, uint32_t &mb, int K = imsame::K) {
    mf = window_mismatch(dpk, p, qpk, e + 1);
    const int sd = (int)p - (K + 1) - 31, sq = (int)e - K - 31;
    if (sd >= 0 && sq >= 0) {
        mb = window_mismatch(dpk, (uint32_t)sd, qpk, (uint32_t)sq);
    } else {  // the window starts before base 0 of an array (first read only)
        const uint64_t a = sd >= 0 ? fetch32(dpk, (uint64_t)sd) : (fetch32(dpk, 0) << (2 * (-sd)));
        const uint64_t b = sq >= 0 ? fetch32(qpk, (uint64_t)sq) : (fetch32(qpk, 0) << (2 * (-sq)));
        mb = mismatch32(a, b);
    }
    mb = brev32(mb);
    const int fm = s.fmax < 0 ? 0 : s.fmax, bm = s.bmax < 0 ? 0 : s.bmax;
    if (fm < 32) mf |= 0xFFFFFFFFu << fm;
    if (bm < 32) mb |= 0xFFFFFFFFu << bm;
}

// ---- the two halves of a seed hit in bit-plane form -------------------------------------------
// The first window each way of a hit (word ending at query base e, database index p after the word) is
// formed from two independent halves: what K1 stores per query word (qtable.cuh: QEntry) and what the scan
// computes once per database position, shared by all hits of that position.  Each half holds its 32 bases
// after / before the word as bit planes (common.cuh) and the steps left inside its read.
struct HitHalf {
    uint32_t f_lo, f_hi, b_lo, b_hi;
    int froom, broom;  // forward: end - (index after the word); backward: bases of the read before the word
};
IMS_HD HitHalf query_half(const uint32_t *qpk, uint32_t e, uint32_t ys, uint32_t yend, int K = imsame::K) {
    HitHalf h;
    planes_fwd(qpk, (uint64_t)e + 1, h.f_lo, h.f_hi);
    planes_bwd(qpk, (int64_t)e - K, h.b_lo, h.b_hi);
    h.froom = (int)(yend - (e + 1));
    h.broom = (int)e - (K - 1) - (int)ys;  // -1 for the cross-read "phantom" word
    return h;
}
IMS_HD HitHalf db_half(const uint32_t *dpk, uint32_t p, uint32_t xs, uint32_t xend, int K = imsame::K) {
    HitHalf h;
    planes_fwd(dpk, p, h.f_lo, h.f_hi);
    planes_bwd(dpk, (int64_t)p - (K + 1), h.b_lo, h.b_hi);
    h.froom = (int)(xend - p);
    h.broom = (int)(p - (uint32_t)K - xs);
    return h;
}
// v << n, 0 for n >= 32 (PTX shl.b32 clamps; the C++ operator is undefined there)
IMS_HD uint32_t shl_clamp(uint32_t v, uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
#else
    return n >= 32u ? 0u : v << n;
#endif
}
// = ext_init (fmax, bmax) + ext_first_masks, without touching either sequence
IMS_HD void hit_first_masks(const HitHalf &d, const HitHalf &q, ExtState &s, uint32_t &mf, uint32_t &mb) {
    s.fmax = q.froom < d.froom ? q.froom : d.froom;
    s.bmax = q.broom < d.broom ? q.broom : d.broom;
    s.pos_f = 0;
    mf = ((d.f_lo ^ q.f_lo) | (d.f_hi ^ q.f_hi)) | shl_clamp(0xFFFFFFFFu, (uint32_t)(s.fmax < 0 ? 0 : s.fmax));
    mb = ((d.b_lo ^ q.b_lo) | (d.b_hi ^ q.b_hi)) | shl_clamp(0xFFFFFFFFu, (uint32_t)(s.bmax < 0 ? 0 : s.bmax));
}

// First forward and first backward window of TWO independent hits, the two dependent chains of
// table lookups interleaved (the scan kernel is bound by the latency of those chains, not by
// instruction issue).  On return each state is done (phase 2) or parked-ready: phase 0 / 1 with
// t = 32, to be continued by ext_window.
template <class LUT>
IMS_HD void ext_first2(ExtState &sa, ExtState &sb, const LUT *lut, uint32_t mfa, uint32_t mba, uint32_t mfb,
                       uint32_t mbb, int K = imsame::K) {
    const int k0 = K << EXT_SC_SHIFT, kb = k0 + EXT_BEST_BIAS;
    int ra = k0, ba = kb, rb = k0, bb = kb;
    mfa = ext_prep(lut, mfa); mba = ext_prep(lut, mba); mfb = ext_prep(lut, mfb); mbb = ext_prep(lut, mbb);
    const bool top9 = K >= 9;  // both walks start in the table's "9 or more" row
    if (top9) { ext_step_row9(lut, mfa, ra, ba); ext_step_row9(lut, mfb, rb, bb); }
    else { ext_step(lut, mfa, 0, ra, ba); ext_step(lut, mfb, 0, rb, bb); }
#pragma unroll
    for (int c = 8; c < 32; c += 8) {
        ext_step(lut, mfa, c, ra, ba);
        ext_step(lut, mfb, c, rb, bb);
    }
    const int sca = ra >> EXT_SC_SHIFT, scb = rb >> EXT_SC_SHIFT;
    const bool fa_over = sca == 0 || sa.fmax <= 32, fb_over = scb == 0 || sb.fmax <= 32;
    // backward restarts from high_right (:339) with high_left = 12 (:303)
    int ra2 = (ba & ~(int)EXT_POS_MASK) - EXT_BEST_BIAS, ba2 = kb, rb2 = (bb & ~(int)EXT_POS_MASK) - EXT_BEST_BIAS, bb2 = kb;
    const int hra = ra2 >> EXT_SC_SHIFT, hrb = rb2 >> EXT_SC_SHIFT;
    if (top9) { ext_step_row9(lut, mba, ra2, ba2); ext_step_row9(lut, mbb, rb2, bb2); }
    else { ext_step(lut, mba, 0, ra2, ba2); ext_step(lut, mbb, 0, rb2, bb2); }
#pragma unroll
    for (int c = 8; c < 32; c += 8) {
        ext_step(lut, mba, c, ra2, ba2);
        ext_step(lut, mbb, c, rb2, bb2);
    }
    if (sa.bmax <= 0) { ra2 = hra << EXT_SC_SHIFT; ba2 = kb; }  // no room: no backward step at all
    if (sb.bmax <= 0) { rb2 = hrb << EXT_SC_SHIFT; bb2 = kb; }
    const int sca2 = ra2 >> EXT_SC_SHIFT, scb2 = rb2 >> EXT_SC_SHIFT;
    // +-1 per step: 2 * matches = steps + score change
    sa.idn2 = (ra & (int)EXT_POS_MASK) + sca - K;
    sb.idn2 = (rb & (int)EXT_POS_MASK) + scb - K;
    if (fa_over) {
        sa.pos_f = ba & (int)EXT_POS_MASK;
        sa.idn2 += (ra2 & (int)EXT_POS_MASK) + sca2 - hra;
        sa.run = ra2; sa.best = ba2; sa.t = 32;
        sa.phase = (sca2 == 0 || sa.bmax <= 32) ? 2 : 1;
    } else {
        sa.run = ra; sa.best = ba; sa.t = 32; sa.phase = 0;
    }
    if (fb_over) {
        sb.pos_f = bb & (int)EXT_POS_MASK;
        sb.idn2 += (rb2 & (int)EXT_POS_MASK) + scb2 - hrb;
        sb.run = rb2; sb.best = bb2; sb.t = 32;
        sb.phase = (scb2 == 0 || sb.bmax <= 32) ? 2 : 1;
    } else {
        sb.run = rb; sb.best = bb; sb.t = 32; sb.phase = 0;
    }
}

// n = 2 * idents - t_len, t_len = fe + K + be + 1 (:359) with fe = pos_f - 1, be = pos_b - 1
// (after the last window `best` is the backward maximum, or still "no step" when there was no backward walk)
IMS_HD int ext_result(const ExtState &s, int K = imsame::K) {
    const int pos_b = s.best & (int)EXT_POS_MASK;
    return 2 * K + s.idn2 - (s.pos_f + K + pos_b - 1);
}

}  // namespace imsame
