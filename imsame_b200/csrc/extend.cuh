// extend.cuh -- ungapped X-drop-to-zero extension of one seed hit on 2-bit
// packed sequences.  Restates src/alignmentFunctions.c:276-359 in units of
// POINT (score 12 = 48/4, +-1 per base):
//   forward from (p, e+1) while score > 0 inside both reads; idents++ on match;
//   `if (high_right <= score) end = cur`  (:324);
//   backward from (p-13, e-12) starting at score = high_right while
//   high_left stays at its initial 12 (:303,339); `if (high_left <= score) start = cur` (:347);
//   t_len = end - start (:359);  returns n = 2*idents - t_len  (raw score = 4n, :373).
// p = database index AFTER the seed's last base (llpos.pos), e = query index of
// the seed's last base (curr_pos).  [xs, xend) and [ys, yend) are the reads.
#pragma once
#include "common.cuh"

namespace imsame {

IMS_HD int extend_hit(const uint32_t *dpk, const uint32_t *qpk, uint32_t p, uint32_t e, uint32_t xs,
                      uint32_t xend, uint32_t ys, uint32_t yend) {
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
// Table-driven form of the same walk: 8 bases per step.  lut[(row-1)*256 + m] describes
// what the loop of src/alignmentFunctions.c:318-333 does on 8 consecutive bases with
// mismatch byte m (bit t = base t differs) when it enters with score `row` (units of
// POINT; row 9 stands for "9 or more": the walk cannot drop to 0 within 8 steps):
//   bits 0-3  matches counted before the walk stops (idents += ...)
//   bit  4    the score reached 0 inside these 8 steps (loop ends)
//   bits 5-8  1 + max prefix score relative to the entry score (over the executed steps)
//   bits 9-12 step (1..8) of the LAST occurrence of that maximum
// The reference's update `if (high <= score) { pos = cur; high = score; }` leaves
// pos = last step whose score equals the maximum, provided that maximum is >= high.
namespace imsame {

constexpr int EXT_LUT_ROWS = 9;
constexpr int EXT_LUT_SIZE = EXT_LUT_ROWS * 256;

IMS_HD uint16_t ext_lut_entry(int row, uint32_t m) {
    int sc = row >= EXT_LUT_ROWS ? 1000 : row, run = 0, rmax = -100, amax = 0, matches = 0, term = 0;
    for (int t = 0; t < 8; t++) {
        if ((m >> t) & 1) run -= 1; else { run += 1; matches++; }
        if (run >= rmax) { rmax = run; amax = t + 1; }
        if (sc + run <= 0) { term = 1; break; }
    }
    return (uint16_t)(matches | (term << 4) | ((rmax + 1) << 5) | (amax << 9));
}

inline void build_ext_lut(uint16_t *lut) {
    for (int row = 1; row <= EXT_LUT_ROWS; row++)
        for (uint32_t m = 0; m < 256; m++) lut[(row - 1) * 256 + m] = ext_lut_entry(row, m);
}

IMS_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
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

// up to `lim` (1..32) steps over mismatch mask mm (bit t = step t); returns true when the walk ended
IMS_HD bool ext_walk32(const uint16_t *lut, uint32_t mm, int lim, int base, int &sc, int &hi, int &last,
                       int &idn) {
    for (int c = 0; c < lim; c += 8) {
        uint32_t m = (mm >> c) & 0xFFu;
        const int l8 = lim - c;
        if (l8 < 8) m |= (0xFFu << l8) & 0xFFu;  // steps past the read end: mismatches (never counted)
        const uint32_t ent = lut[((sc < EXT_LUT_ROWS ? sc : EXT_LUT_ROWS) - 1) * 256 + m];
        idn += (int)(ent & 15u);
        const int cand = sc + (int)((ent >> 5) & 15u) - 1;
        if (cand >= hi) { hi = cand; last = base + c + (int)((ent >> 9) & 15u) - 1; }
        sc += 8 - 2 * popc32(m);
        if (ent & 16u) return true;
    }
    return false;
}

IMS_HD int extend_hit_lut(const uint16_t *lut, const uint32_t *dpk, const uint32_t *qpk, uint32_t p, uint32_t e,
                          uint32_t xs, uint32_t xend, uint32_t ys, uint32_t yend) {
    int fmax = (int)(xend - p);
    {
        const int fq = (int)(yend - (e + 1));
        fmax = fq < fmax ? fq : fmax;
    }
    int sc = K, hr = K, idn = K, fe = -1;
    for (int t = 0; t < fmax; t += 32) {
        const uint32_t mm = mismatch32(fetch32(dpk, (uint64_t)p + t), fetch32(qpk, (uint64_t)e + 1 + t));
        if (ext_walk32(lut, mm, (fmax - t) < 32 ? (fmax - t) : 32, t, sc, hr, fe, idn)) break;
    }
    int bmax = (int)(p - K - xs);
    {
        const int bq = (int)e - (K - 1) - (int)ys;
        bmax = bq < bmax ? bq : bmax;
    }
    sc = hr;
    int hl = K, be = -1;
    for (int t = 0; t < bmax; t += 32) {
        const int64_t sd = (int64_t)p - (K + 1) - t - 31, sq = (int64_t)e - K - t - 31;
        const uint64_t a = sd >= 0 ? fetch32(dpk, (uint64_t)sd) : (fetch32(dpk, 0) << (2 * (int)(-sd)));
        const uint64_t b = sq >= 0 ? fetch32(qpk, (uint64_t)sq) : (fetch32(qpk, 0) << (2 * (int)(-sq)));
        const uint32_t mm = brev32(mismatch32(a, b));  // step u <-> bit u
        if (ext_walk32(lut, mm, (bmax - t) < 32 ? (bmax - t) : 32, t, sc, hl, be, idn)) break;
    }
    return 2 * idn - (fe + K + be + 1);
}

}  // namespace imsame

// ---------------------------------------------------------------------------------------
// Window-at-a-time form for the scan kernel: the walk of one hit is a small state machine
// that consumes one 32-base window (forward or backward) per call, four table steps with
// no data-dependent branch, so that the 32 hits of a warp stay converged: every lane
// executes the same window body until all walks of the warp are finished.
namespace imsame {

struct ExtState {
    int phase;  // 0 forward, 1 backward, 2 done
    int t;      // steps already taken in this phase
    int sc;     // current score (units of POINT)
    int hr, fe; // forward: high_right, last step that reached it (-1: none)   (:324-327)
    int hl, be; // backward: high_left, last step that reached it (-1: none)   (:347-351)
    int idn;    // idents
    int fmax, bmax;  // steps available inside both reads
};

IMS_HD void ext_init(ExtState &s, uint32_t p, uint32_t e, uint32_t xs, uint32_t xend, uint32_t ys, uint32_t yend) {
    const int fd = (int)(xend - p), fq = (int)(yend - (e + 1));
    const int bd = (int)(p - K - xs), bq = (int)e - (K - 1) - (int)ys;  // -1 for the phantom word
    s.fmax = fq < fd ? fq : fd;
    s.bmax = bq < bd ? bq : bd;
    s.t = 0;
    s.sc = K; s.hr = K; s.fe = -1; s.hl = K; s.be = -1; s.idn = K;
    s.phase = s.fmax > 0 ? 0 : (s.bmax > 0 ? 1 : 2);
}

IMS_HD void ext_window(ExtState &s, const uint16_t *lut, const uint32_t *dpk, const uint32_t *qpk, uint32_t p,
                       uint32_t e) {
    const bool bwd = s.phase == 1;
    const int maxs = bwd ? s.bmax : s.fmax;
    const int rem = maxs - s.t;
    const int64_t sd = bwd ? (int64_t)p - (K + 1) - s.t - 31 : (int64_t)p + s.t;
    const int64_t sq = bwd ? (int64_t)e - K - s.t - 31 : (int64_t)e + 1 + s.t;
    const uint64_t a = sd >= 0 ? fetch32(dpk, (uint64_t)sd) : (fetch32(dpk, 0) << (2 * (int)(-sd)));
    const uint64_t b = sq >= 0 ? fetch32(qpk, (uint64_t)sq) : (fetch32(qpk, 0) << (2 * (int)(-sq)));
    uint32_t mm = mismatch32(a, b);
    if (bwd) mm = brev32(mm);                    // step u <-> bit u in both directions
    if (rem < 32) mm |= 0xFFFFFFFFu << rem;      // steps past the read end: mismatches (never counted)
    int sc = s.sc, hi = bwd ? s.hl : s.hr, last = bwd ? s.be : s.fe, idn = s.idn;
    bool term = false;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
        const uint32_t m = (mm >> c) & 0xFFu;
        int row = sc < EXT_LUT_ROWS ? sc : EXT_LUT_ROWS;
        row = row < 1 ? 1 : row;
        const uint32_t ent = lut[(row - 1) * 256 + m];
        if (!term) {
            idn += (int)(ent & 15u);
            const int cand = sc + (int)((ent >> 5) & 15u) - 1;
            if (cand >= hi) { hi = cand; last = s.t + c + (int)((ent >> 9) & 15u) - 1; }
            sc += 8 - 2 * popc32(m);
            term = (ent & 16u) != 0;
        }
    }
    s.idn = idn;
    const bool phase_over = term || rem <= 32;
    if (!bwd) {
        s.hr = hi; s.fe = last;
        if (phase_over) { s.phase = s.bmax > 0 ? 1 : 2; s.t = 0; s.sc = hi; }  // backward restarts from high_right (:339)
        else { s.t += 32; s.sc = sc; }
    } else {
        s.hl = hi; s.be = last;
        if (phase_over) s.phase = 2;
        else { s.t += 32; s.sc = sc; }
    }
}

IMS_HD int ext_result(const ExtState &s) { return 2 * s.idn - (s.fe + K + s.be + 1); }

}  // namespace imsame
