// nwp_core.cuh -- "packed word" form of the NW wavefront step (nw_core.cuh) for
// short reads: score, tie-break priority and the forward-carried traceback
// statistics of a cell live in ONE 32-bit word, so that the reference's
// three-way choice with its tie rules (src/alignmentFunctions.c:457-472) is a
// single 3-input integer maximum and no per-cell multiply is left.
//
//   bits  0..7   identities on the cell's traceback path          (<= 255)
//   bits  8..17  alignment columns on the path                    (<= 1023)
//   bits 18..19  priority of the candidate: D = 2, L (row maximum) = 1,
//                R (column maximum) = 0; zero in stored cells
//   bits 20..31  score (signed, two's complement over the whole word)
//
// The reference picks D when D >= L && D >= R, else R when R > L, else L
// (:457-472): on equal scores D beats L beats R, which is exactly the order of
// the priority field, and because the three priorities differ the statistics
// bits below them never decide a comparison.
//
// Lagged affine gaps without multiplies: the reference recomputes
//   L = mf.score + iGap + (j - (mf.y+1)) * eGap         (:443-447)
//   R = mc.score + iGap + (i - (mc.x+1)) * eGap         (:449-453)
// from the stored maximum and its position.  Here the *candidate word* is kept
// instead: when the row maximum is (re)installed at column j from T[i-1][j-2]
// the L candidate is that cell's word + (iGap + eGap, 2 columns, priority L),
// and every further column adds STEP = (eGap, 1 column).  The same for the
// column maxima, row by row.  Only the comparisons of :434 and :476 need the
// raw score of the maximum, kept as a second word with the low 20 bits cleared
// (row maximum, `<=` test) or set (column maximum, `>` test) so that a plain
// 32-bit compare against a full cell word decides on the scores alone.
//
// Cells are stored biased by B = (iGap + eGap, 2 columns, priority 1): a stored
// word *is* the L candidate it would install, the R candidate is one subtract
// away, and the bias is folded into the two per-cell constants (diagonal
// statistics, match score) that come from a 4-cells-per-entry table indexed by
// the mismatch bits of the step.
//
// Valid (pw_eligible) while every field stays inside its bits: both reads short
// enough for 8-bit identities / 10-bit lengths, non-positive gap scores, and
// scores incl. the transient gap candidates inside 12 bits.  Everything else
// runs through nw_core.cuh.  Tested bit for bit against the oracle on the CPU
// (tests/emul/nwp_emul.cpp) and on the GPU (tests/test_gpu_parity.py).
//
// Wide reads (query reads of 257..321 bases, e.g. 2 x 300 sequencing: NW classes
// 9 and 10, 18 / 20 columns per lane).  Score (12 bits), priority (2) and length
// (10) still fit, the identities can reach 320 and would need a 33rd bit.  They
// do not get one: the statistics never decide a comparison (see above), so the
// low 18 bits are simply the integer V = 256 * length + identities carried along
// the chosen path, exact as long as V < 2^18, and the only question is how V
// splits at the end.  identities < 512 leaves two readings, (V >> 8, V & 255)
// and (V >> 8) - 1, (V & 255) + 256).  pw_split_stats() rules one out from the
// geometry of the path that ends in the best cell (identities are diagonal
// steps: at most min(bi, bj), at most the columns, and a path with d diagonal
// steps into (bi, bj) has at most bi + bj - d columns), which settles every
// pair with fewer than 256 identities whose second reading would need more
// diagonal steps than the matrix has -- all random pairs.  What stays open
// (near-complete overlaps: 256 or more identities) is run once more with the
// length unit set to 0 (pw_consts(..., lenu = 0)): the same path, the low bits
// now the identities alone; the length follows from V.
#pragma once
#include "nw_core.cuh"

namespace imsame {

constexpr int PW_LEN1 = 1 << 8;
constexpr int PW_PR1 = 1 << 18;
constexpr int PW_SC1 = 1 << 20;
constexpr int PW_LOW = PW_SC1 - 1;   // id | len | prio
constexpr int PW_SMASK = ~PW_LOW;    // score field
constexpr int PW_PRMASK = 3 * PW_PR1;
constexpr int PW_LANES = 16;         // lanes per pair (half a warp)
constexpr int PW_MAX_S = 20;         // columns per lane (17..20: wide reads, see above)
constexpr int PW_MAX_Y = PW_LANES * PW_MAX_S + 1;  // 321: Y1 <= 320 columns in one pass
constexpr int PW_NARROW_Y1 = 255;    // up to here min(X1, Y1) <= 255 whatever the database read: 8-bit identities
constexpr int PW_MAX_X = 512;
constexpr uint32_t PW_VMASK = (1u << 18) - 1u;  // V = 256 * length + identities

// per-run constants (functions of igap / egap only)
struct PwK {
    int B;        // bias of stored words
    int step;     // one more gap column: (egap, 1 column)
    int negz;     // "-inf" row maximum for the lane that owns column 1 (low bits clear)
    int negb;     // its T[i][-1]: below negz so that the j = 1 test of :434 fails
    int negl;     // "-inf" L candidate at j = 1 (:446)
    int ds_mis;   // diagonal constant on mismatch: D0 = T'[i-1][j-1] + ds  (ds_match = ds_mis + 1)
    int sb_mis;   // after the maximum: T'[i][j] = (m & ~prio) + sb            (sb_match = sb_mis + 8 * SC1)
    int one;      // 1, opaque to the compiler on the device (pw_row: diagonal add as IMAD)
};

// lenu: what one alignment column adds to the statistics bits (PW_LEN1; 0 = carry the identities alone, the
// second run of a wide pair whose statistics did not split).  bias: score offset of the run (pw_bias): every
// word that carries a score is derived from a stored cell by adding constants, and scores are only ever
// compared with each other, so adding `bias` to the stored cells (through B) moves the 12-bit window from
// [-2048, 2047] to [-2048 - bias, 2047 - bias] -- scores reach 4 (min + 1) upwards but, with the gap
// candidates, 4 (min + 1) + |igap| + |egap| (max + 3) downwards.
IMS_HD PwK pw_consts(int igap, int egap, int one = 1, int lenu = PW_LEN1, int bias = 0) {
    PwK k;
    k.one = one;
    const int b0 = (igap + egap) * PW_SC1 + 2 * lenu + PW_PR1;  // a stored word is the L candidate it would install
    k.B = b0 + bias * PW_SC1;
    k.step = egap * PW_SC1 + lenu;
    k.negz = -2046 * PW_SC1;
    k.negb = -2047 * PW_SC1;
    k.negl = -1900 * PW_SC1;
    k.ds_mis = lenu + 2 * PW_PR1 - b0;    // candidates keep the offset of the cell they come from
    k.sb_mis = -NW_POINT * PW_SC1 + b0;
    return k;
}

// how far below zero a score can fall for reads of up to xmax x ymax bases: |T| <= 4 (min(i,j) + 1); gap
// candidates fall at most |igap| + |egap| (mx + 3) below that
IMS_HD long pw_depth(uint32_t xmax, uint32_t ymax, int igap, int egap) {
    const int X1 = (int)xmax - 1, Y1 = (int)ymax - 1;
    const int mn = X1 < Y1 ? X1 : Y1, mx = X1 < Y1 ? Y1 : X1;
    return 4L * (mn + 1) + (long)(-igap) + (long)(-egap) * (mx + 3) + 24;
}
constexpr long PW_DEPTH0 = 1890;  // depth that fits without a bias: stays above negl (and its one STEP)
constexpr long PW_TOP = 2040;     // 4 (min + 1) + bias stays below this

// Score offset of a run whose longest reads have xmax / ymax bases (lengths beyond what packed words take are
// clamped: in a mixed run the generic kernel has those pairs).  0 for everything that fits without one (reads
// of 250 bases with the default gap scores: depth 1533).
IMS_HD int pw_bias(uint32_t xmax, uint32_t ymax, int igap, int egap) {
    if (igap > 0 || egap > 0) return 0;
    if (xmax > (uint32_t)PW_MAX_X) xmax = PW_MAX_X;
    if (ymax > (uint32_t)PW_MAX_Y) ymax = PW_MAX_Y;
    if (xmax < 2 || ymax < 2) return 0;
    const long d = pw_depth(xmax, ymax, igap, egap);
    if (d <= PW_DEPTH0) return 0;
    const int mn = (int)(xmax < ymax ? xmax : ymax) - 1;
    // no offset holds the longest pairs (very negative gap scores): none, as many short pairs as possible stay packed
    if (4L * (mn + 1) + (d - PW_DEPTH0) > PW_TOP) return 0;
    return (int)(d - PW_DEPTH0);
}

// Can every pair with xlen <= xmax, ylen <= ymax run in packed words (in a run with this score offset)?
IMS_HD bool pw_eligible(uint32_t xmax, uint32_t ymax, int igap, int egap, int bias) {
    if (igap > 0 || egap > 0) return false;             // the row-1 R candidate must lose to D (see pw_lane_init)
    if (xmax < 2 || ymax < 2) return true;              // no cells at all
    if (ymax > (uint32_t)PW_MAX_Y || xmax > (uint32_t)PW_MAX_X) return false;
    const int X1 = (int)xmax - 1, Y1 = (int)ymax - 1;
    const int mn = X1 < Y1 ? X1 : Y1;
    // identities <= diagonal steps <= min(X1, Y1): 8 bits up to Y1 = 255 (the classes 1..8 kernels split V
    // as it is), one carry into the length field beyond (classes 9, 10: pw_split_stats)
    if (mn > 511) return false;
    if (X1 + Y1 + 3 > 1023) return false;                // columns on a path <= X1 + Y1, + 2 of bias, + 1 carry
    if (pw_depth(xmax, ymax, igap, egap) - bias > PW_DEPTH0) return false;
    if (4L * (mn + 1) + bias > PW_TOP) return false;
    if (-egap > 100) return false;
    return true;
}

// (both bounds are monotone in both lengths and Y1 <= 255 implies min(X1, Y1) <= 255: a run whose longest reads
// pass consists of pairs that pass, each in the kernel of its own class)
// the same test on ONE pair's own lengths (mixed runs: NwArgs.mixed)
IMS_HD bool pw_pair_eligible(uint32_t xlen, uint32_t ylen, int igap, int egap, int bias) {
    return xlen >= 2 && ylen >= 2 && pw_eligible(xlen, ylen, igap, egap, bias);
}

struct PwLink {
    int a;    // T'[i][j0-1]
    int b;    // T'[i][j0-2]
    int mfz;  // score of the row maximum after cell (i, j0-1), low bits clear
    int lw;   // its L candidate for column j0
};

template <int S>
struct PwRow {
    int h[S + 1];  // T'[.][j0-1+k], k = 0..S (k = 0: left halo column)
    int g;         // T'[.][j0-2]
};

template <int S>
struct PwLane {
    PwRow<S> r0, r1;  // row histories; at step t, r[t&1] is T[i-1] and r[~t&1] is T[i-2]
    int mck[S];       // score of the column maximum of column j0-1+k, low bits set
    int rw[S];        // its R candidate for the next row
    int bw, bz, bi, bj;  // best border cell: word, score-only word, row, column
};

// the two per-cell constants of four consecutive cells
struct PwE4 {
    int ds[4], sb[4];
};
IMS_HD PwE4 pw_e4(const PwK &k, uint32_t mis4 /* bit 2c: cell c mismatches */) {
    PwE4 e;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const bool mis = (mis4 >> (2 * c)) & 1u;
        e.ds[c] = k.ds_mis + (mis ? 0 : 1);
        e.sb[c] = k.sb_mis + (mis ? 0 : 2 * NW_POINT * PW_SC1);
    }
    return e;
}

IMS_HD int pw_row0(const PwK &k, uint32_t x0, uint32_t y) { return ((y == x0) ? NW_POINT : -NW_POINT) * PW_SC1 + k.B; }

// Row 0 (:404-413) and the column maxima mc[j] = (T[0][j], x = 0).  The R candidate of a
// column maximum installed from row x is first needed on row x + 2... row 0 is the exception
// (needed on row 2, the reference has no R on row 1, :449): rw starts one STEP early, so on
// row 1 it reads T[0][j-1] + igap, which can never beat D = T[0][j-1] (igap <= 0, priority
// R < D), and on row 2 it is exact.
// ypk_halo: codes of Y[j0-2], Y[j0-1], Y[j0] .. at bits 0,2,4,...; t_first = step of row 1.
template <int S>
IMS_HD void pw_lane_init(PwLane<S> &L, const PwK &k, uint32_t x0, uint64_t ypk_halo, bool first_lane, int t_first) {
    const bool odd = (t_first & 1) != 0;
#pragma unroll
    for (int c = 0; c <= S; c++) {
        const uint32_t y = (uint32_t)(ypk_halo >> (2 * (c + 1))) & 3u;
        const int row0 = pw_row0(k, x0, y);
        // the other history is "row -1": the first column-maximum test (:476, guard i>1) must fail
        L.r0.h[c] = odd ? k.negb : row0;
        L.r1.h[c] = odd ? row0 : k.negb;
        if (c < S) {
            L.mck[c] = row0 | PW_LOW;
            L.rw[c] = row0 - PW_PR1 - k.step;
        }
    }
    // column 0's maximum is never updated (:476, guard j>1)
    if (first_lane) L.mck[0] = 0x7FFFFFFF;
    const int g0 = pw_row0(k, x0, (uint32_t)ypk_halo & 3u);
    L.r0.g = odd ? k.negb : g0;
    L.r1.g = odd ? g0 : k.negb;
    L.bw = L.bz = (int)0x80000000;
    L.bi = L.bj = 0;
}

// what the lane owning column 1 receives instead of a neighbour's link (cf. nw_first_link)
IMS_HD PwLink pw_first_link(const PwK &k, uint32_t xi, uint32_t y0) {
    PwLink l;
    l.a = pw_row0(k, xi, y0);  // T[i][0] = +-4 (:426)
    l.b = k.negb;
    l.mfz = k.negz;
    l.lw = k.negl;
    return l;
}

// One row of the lane's strip.  P1 = T'[i-1] (read only), P2 = T'[i-2] on entry and T'[i]
// on exit.  ew(g) returns the constants of cells 4g .. 4g+3.  cl / owns_last as in nw_row.
// CL >= 0: the slot of the last column is a compile-time constant (every query read of the launch has
// the same length) and is read straight out of its register; CL = -1: run-time slot, a switch per row.
template <int S, class EW, int CL = -1>
IMS_HD void pw_row(PwLane<S> &L, const PwRow<S> &P1, PwRow<S> &P2, const PwLink &in, PwLink &out, int i, int j0,
                   const EW &ew, const PwK &k, int X1, int Y1, int cl, bool owns_last) {
    int mfz = in.mfz, lw = in.lw;
    int nt = in.a;   // becomes slot c of the new row: T'[i][j0-1+c]
    int t2 = in.b;   // T'[i][j-2]
    int r2 = P1.g;   // T'[i-1][j-2]
    PwE4 e;
#pragma unroll
    for (int c = 0; c < S; c++) {
        if ((c & 3) == 0) e = ew(c >> 2);
        const int o2 = P2.h[c];  // T'[i-2][j-1], consumed below; its slot takes T'[i][j-1]
        P2.h[c] = nt;
        // row maximum: tests row i, copies row i-1 (:434-438)
        const bool up = mfz <= t2;
        mfz = up ? (r2 & PW_SMASK) : mfz;
        lw = up ? r2 : lw;
        const int d = P1.h[c];
#if defined(__CUDA_ARCH__)
        // diagonal candidate as a multiply-add with a run-time 1 (FMA pipe): ptxas then keeps the 3-input
        // maximum as ONE VIMNMX3 instead of fusing the add into a VIADDMNMX + VIMNMX pair (two ALU-pipe
        // instructions; the ALU pipe is what limits this kernel)
        int d0;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d0) : "r"(d), "r"(k.one), "r"(e.ds[c & 3]));
#else
        const int d0 = d + e.ds[c & 3];
#endif
        const int m = max3i(d0, lw, L.rw[c]);
        lw += k.step;
        // column maximum of column j-1 absorbs T[i-2][j-1], strictly greater only (:476-480)
        const bool uc = o2 > L.mck[c];
        L.mck[c] = uc ? (o2 | PW_LOW) : L.mck[c];
        L.rw[c] = (uc ? (o2 - PW_PR1) : L.rw[c]) + k.step;
        t2 = nt;
        r2 = d;
        nt = (m & ~PW_PRMASK) + e.sb[c & 3];
    }
    P2.h[S] = nt;
    P2.g = in.b;
    out.a = nt;
    out.b = t2;
    out.mfz = mfz;
    out.lw = lw;
    // best border cell (:481-484, >= : last in row-major order wins): the last column of every row but
    // the last one here; the last row is scanned once per pair from the row history (pw_last_row)
    if (owns_last && i != X1) {
        int lt = P2.h[S];
        if (CL >= 0) lt = P2.h[(CL >= 0 && CL < S ? CL : 0) + 1];
        else switch (cl) {
#define IMS_CASE(C) case C: if (C < S) lt = P2.h[(C < S ? C : 0) + 1]; break;
            IMS_CASE(0) IMS_CASE(1) IMS_CASE(2) IMS_CASE(3) IMS_CASE(4) IMS_CASE(5) IMS_CASE(6) IMS_CASE(7)
            IMS_CASE(8) IMS_CASE(9) IMS_CASE(10) IMS_CASE(11) IMS_CASE(12) IMS_CASE(13) IMS_CASE(14)
            IMS_CASE(15) IMS_CASE(16) IMS_CASE(17) IMS_CASE(18)
#undef IMS_CASE
            default: break;
        }
        if (lt >= L.bz) { L.bw = lt; L.bz = lt & PW_SMASK; L.bi = i; L.bj = Y1; }
    }
}

// After the lane's last step its row history still holds row X1 (in the history that was written last:
// r1 when that step was even, r0 when odd).  All lanes scan their part of the last row at once, after
// the step loop, instead of one lane per step inside it.
template <int S>
IMS_HD void pw_last_row(PwLane<S> &L, const PwRow<S> &row, int j0, int X1, int Y1) {
#pragma unroll
    for (int c = 0; c < S; c++)
        if (j0 + c <= Y1 && row.h[c + 1] >= L.bz) {
            L.bw = row.h[c + 1]; L.bz = row.h[c + 1] & PW_SMASK; L.bi = X1; L.bj = j0 + c;
        }
}

// stored word -> score / statistics
IMS_HD int pw_score(const PwK &k, int w) { return (w - k.B) >> 20; }
IMS_HD uint32_t pw_len(const PwK &k, int w) { return ((uint32_t)(w - k.B) >> 8) & 1023u; }
IMS_HD uint32_t pw_ids(const PwK &k, int w) { return (uint32_t)(w - k.B) & 255u; }
IMS_HD uint32_t pw_stats(const PwK &k, int w) { return (uint32_t)(w - k.B) & PW_VMASK; }  // V (or, lenu = 0, the identities)

// Wide reads: V = 256 * length + identities of the path into the best cell (bi, bj), identities < 512.
// Returns false with the one split the path's geometry admits, true when both readings are possible (the
// pair is then run again with lenu = 0).
IMS_HD bool pw_split_stats(uint32_t V, int bi, int bj, uint32_t *len, uint32_t *id) {
    const uint32_t f = V >> 8, id0 = V & 255u, id1 = id0 + 256u;
    const uint32_t m = (uint32_t)(bi < bj ? bi : bj), span = (uint32_t)(bi + bj);
    const bool can0 = id0 <= m && id0 <= f && f + id0 <= span;
    const bool can1 = f >= 1 && id1 <= m && id1 <= f - 1 && (f - 1) + id1 <= span;
    if (can1 && !can0) { *len = f - 1; *id = id1; return false; }
    *len = f; *id = id0;
    return can0 && can1;
}

}  // namespace imsame
