// nw_core.cuh -- one wavefront step of the reference's full-matrix NW with
// lagged affine gaps (src/alignmentFunctions.c:389-489) plus forward-carried
// traceback statistics (length/identities of :493-560 and :254-258), written as
// a host/device function so the exact lane logic is unit-tested on the CPU
// (tests/emul/nw_emul.cpp) and executed unchanged by the CUDA kernel (nw.cuh).
//
// Geometry: X = database read (rows i), Y = query read (columns j).  A lane owns
// S consecutive columns j0..j0+S-1 of one row per step; lane l works on row
// i = t - l + 1 at step t (anti-diagonal wavefront over lane strips).  Cell
// (i,j) needs T[i-1][j-1], T[i][j-2], T[i-1][j-2], T[i-2][j-1] and the running
// maxima mf (per row) and mc[j-1] (per column), so everything a lane needs from
// its left neighbour was produced one step earlier (NwLink).
//
// Per-cell recurrence (names as in the reference):
//   if (j>1 && mf.score <= T[i][j-2]) mf = (T[i-1][j-2], y=j-2)              :434-438
//   D = T[i-1][j-1];  L = mf.score + iGap + (j-(mf.y+1))*eGap   (j>1)         :441-447
//   R = mc[j-1].score + iGap + (i-(mc[j-1].x+1))*eGap            (i>1)         :449-453
//   D>=L && D>=R -> D ; else R>L -> R ; else L ; then + (X[i]==Y[j] ? 4 : -4)  :440,457-472
//   if (i>1 && j>1 && T[i-2][j-1] > mc[j-1].score) mc[j-1] = (T[i-2][j-1], x=i-2)   :476-480
// The match score is added to all three candidates, so the choice is made on
// the values without it.  Statistics word p = (length << 16) | identities of the
// cell's own traceback path: diagonal -> +1 column, +1 identity on match; jump to
// (px,py) -> + max(i-px, j-py) columns (src/alignmentFunctions.c:514-543).
#pragma once
#include "common.cuh"

namespace imsame {

constexpr int NW_NEG = -(1 << 28);  // "-inf" that survives a handful of additions
constexpr int NW_POINT = 4;         // POINT, src/structs.h:13

struct NwLink {
    int a, ap;          // T[i][j0-1] and its statistics
    int b, bp;          // T[i][j0-2] and its statistics
    int mfs, mfy, mfp;  // row maximum after cell (i, j0-1): score, column, statistics
};

struct NwBest {
    int s, i, j, p;  // score, row, column, statistics of the best border cell
};

// "last in row-major order among equal scores" (src/alignmentFunctions.c:483 uses >=)
IMS_HD bool best_better(const NwBest &c, const NwBest &b) {
    if (c.s != b.s) return c.s > b.s;
    if (c.i != b.i) return c.i > b.i;
    return c.j > b.j;
}

template <int S>
struct NwLane {
    int h1[S + 1], p1[S + 1];  // T[i-1][j0-1+k] and statistics, k = 0..S (k = 0: left halo column)
    int h2[S], p2[S];          // T[i-2][j0-1+k]
    int g1, gp1;               // T[i-1][j0-2]
    int mcs[S], mcx[S], mcp[S];  // column maximum of column j0-1+k: score, row, statistics
    NwBest best;
};

// Row 0 (src/alignmentFunctions.c:404-413): T[0][j] = +-4, mc[j] = (T[0][j], x=0).
// ypk_halo: 2-bit codes of Y[j0-2], Y[j0-1], Y[j0] .. Y[j0+S-1] at bits 0,2,4,...
// first_lane: the lane that owns column 1 (its halo is column 0, which has no
// left neighbour and whose column maximum is never updated, guard j>1 at :476).
template <int S>
IMS_HD void nw_lane_init(NwLane<S> &L, uint32_t x0, uint64_t ypk_halo, bool first_lane) {
#pragma unroll
    for (int k = 0; k <= S; k++) {
        const uint32_t y = (uint32_t)(ypk_halo >> (2 * (k + 1))) & 3u;
        L.h1[k] = (y == x0) ? NW_POINT : -NW_POINT;
        L.p1[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < S; k++) {
        L.h2[k] = NW_NEG;
        L.p2[k] = 0;
        L.mcs[k] = L.h1[k];
        L.mcx[k] = 0;
        L.mcp[k] = 0;
    }
    const uint32_t yg = (uint32_t)ypk_halo & 3u;
    L.g1 = first_lane ? NW_NEG : ((yg == x0) ? NW_POINT : -NW_POINT);
    L.gp1 = 0;
    L.best.s = NW_NEG * 2;
    L.best.i = L.best.j = L.best.p = 0;
}

// What the lane owning column 1 receives instead of a neighbour's link: column 0
// is T[i][0] = +-4 (:426); the row maximum starts as "-inf" so that L is -inf at
// j = 1 (:446) and the j = 2 test (:434, always true in the reference because
// mf.score == T[i][0]) always fires and installs (T[i-1][0], y = 0).
IMS_HD NwLink nw_first_link(uint32_t xi, uint32_t y0) {
    NwLink k;
    k.a = (xi == y0) ? NW_POINT : -NW_POINT;
    k.ap = 0;
    k.b = NW_NEG - 1;
    k.bp = 0;
    k.mfs = NW_NEG;
    k.mfy = 0;
    k.mfp = 0;
    return k;
}

// One row of the lane's strip.  mm: bit 2c set <=> X[i] != Y[j0+c].
// X1 = xlen-1, Y1 = ylen-1 (last row / column).  Columns j > Y1 are padding:
// they only ever feed cells further right, never a real one.
// TB: additionally store one back-pointer code per cell for the winners-only traceback
// (src/alignmentFunctions.c:459-471 xfrom/yfrom): 0 = diagonal, 0x8000|x = jump to the
// column maximum (x, j-1), 0x4000|y = jump to the row maximum (i-1, y).
constexpr uint16_t TB_DIAG = 0, TB_COL = 0x8000, TB_ROW = 0x4000, TB_MASK = 0x3FFF;

template <int S, bool TB = false>
IMS_HD void nw_row(NwLane<S> &L, const NwLink &in, NwLink &out, int i, int j0, uint32_t mm, int igap,
                   int egap, int X1, int Y1, bool first_lane, uint16_t *tbrow = nullptr) {
    const int rb = (i == 1) ? NW_NEG : igap;  // no R candidate on row 1 (:449)
    int cur[S], curp[S];
    int mfs = in.mfs, mfy = in.mfy, mfp = in.mfp;
#pragma unroll
    for (int c = 0; c < S; c++) {
        const int j = j0 + c;
        // row maximum: tests row i, copies row i-1 (:434-438)
        const int t2 = (c == 0) ? in.b : (c == 1) ? in.a : cur[c >= 2 ? c - 2 : 0];
        const int r2 = (c == 0) ? L.g1 : L.h1[c >= 1 ? c - 1 : 0];
        const int rp2 = (c == 0) ? L.gp1 : L.p1[c >= 1 ? c - 1 : 0];
        const bool up = mfs <= t2;
        mfs = up ? r2 : mfs;
        mfy = up ? j - 2 : mfy;
        mfp = up ? rp2 : mfp;
        const int mis = (int)((mm >> (2 * c)) & 1u);
        const int d = L.h1[c];
        const int dl = j - 1 - mfy;
        const int l = mfs + dl * egap + igap;
        const int dr = i - 1 - L.mcx[c];
        const int r = L.mcs[c] + dr * egap + rb;
        int t, p;
        if (d >= l && d >= r) {
            t = d;
            p = L.p1[c] + 65536 + (1 - mis);
            if (TB) tbrow[c] = TB_DIAG;
        } else if (r > l) {
            t = r;
            p = L.mcp[c] + ((dr + 1) << 16);
            if (TB) tbrow[c] = (uint16_t)(TB_COL | L.mcx[c]);
        } else {
            t = l;
            p = mfp + ((dl + 1) << 16);
            if (TB) tbrow[c] = (uint16_t)(TB_ROW | mfy);
        }
        cur[c] = t + (mis ? -NW_POINT : NW_POINT);
        curp[c] = p;
        // column maximum of column j-1 absorbs T[i-2][j-1], strictly greater only (:476-480)
        const bool uc = L.h2[c] > L.mcs[c];
        L.mcs[c] = uc ? L.h2[c] : L.mcs[c];
        L.mcx[c] = uc ? i - 2 : L.mcx[c];
        L.mcp[c] = uc ? L.p2[c] : L.mcp[c];
    }
    // best border cell (:481-484): last column on every row, every column on the last row
    if (i == X1) {
#pragma unroll
        for (int c = 0; c < S; c++) {
            if (j0 + c <= Y1 && cur[c] >= L.best.s) {
                L.best.s = cur[c]; L.best.i = i; L.best.j = j0 + c; L.best.p = curp[c];
            }
        }
    } else if (Y1 >= j0 && Y1 < j0 + S) {
        int cs = cur[0], cp = curp[0];
#pragma unroll
        for (int c = 1; c < S; c++)
            if (Y1 - j0 == c) { cs = cur[c]; cp = curp[c]; }
        if (cs >= L.best.s) { L.best.s = cs; L.best.i = i; L.best.j = Y1; L.best.p = cp; }
    }
    // hand-over to the right neighbour
    out.a = cur[S - 1];
    out.ap = curp[S - 1];
    out.b = (S >= 2) ? cur[S >= 2 ? S - 2 : 0] : in.a;
    out.bp = (S >= 2) ? curp[S >= 2 ? S - 2 : 0] : in.ap;
    out.mfs = mfs;
    out.mfy = mfy;
    out.mfp = mfp;
    // shift the row history
#pragma unroll
    for (int k = 0; k < S; k++) {
        L.h2[k] = L.h1[k];
        L.p2[k] = L.p1[k];
    }
    if (first_lane) L.h2[0] = NW_NEG;  // column 0's maximum is never updated (:476, j>1)
    L.h1[0] = in.a;
    L.p1[0] = in.ap;
#pragma unroll
    for (int k = 1; k <= S; k++) {
        L.h1[k] = cur[k - 1];
        L.p1[k] = curp[k - 1];
    }
    L.g1 = in.b;
    L.gp1 = in.bp;
}

}  // namespace imsame
