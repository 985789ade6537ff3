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
