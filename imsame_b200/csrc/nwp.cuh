// nwp.cuh -- K3p: the NW/filter/selection kernel of nw.cuh for short reads, in packed words
// (nwp_core.cuh).  Half a warp per candidate pair: 16 lanes x S columns cover a query read of
// up to 257 bases in one pass, two pairs per warp run in lockstep.  One 3-input maximum, two
// compares and four selects per cell on the ALU pipe, the rest adds/logic ops that the FMA
// pipe co-issues; the two per-cell constants (diagonal statistics, match score) come from a
// shared-memory table indexed by the step's mismatch bits, four cells per 2 x 128-bit load.
// Same inputs, outputs, pruning and epilogue as nw_kernel<S, false>; used whenever
// pw_eligible() holds for the run (capi.cu), bit-identical results otherwise impossible.
// S = 18 and 20 ("wide": query reads of 257..321 bases, nwp_core.cuh): 168 registers, blocks of 128
// threads, three per SM; the table holds the constants of both length units side by side and a pair whose
// statistics word does not split is run a second time by the same half warp (both halves of the warp
// switch together, so that the mode stays warp-uniform; the other half idles unless it has such a pair too).
#pragma once
#include "nw.cuh"
#include "nwp_core.cuh"

namespace imsame {

constexpr int NWP_TBL = 0x56;  // table index = mismatch bits of 4 cells at bits 0,2,4,6
IMS_HD constexpr bool nwp_wide(int S) { return S > 16; }
IMS_HD constexpr int nwp_threads(int S) { return nwp_wide(S) ? 128 : 256; }
#ifndef NWP_WIDE_MIN_BLOCKS
#define NWP_WIDE_MIN_BLOCKS 3  // x 128 threads: 168 registers per thread
#endif
IMS_HD constexpr int nwp_min_blocks(int S) { return nwp_wide(S) ? NWP_WIDE_MIN_BLOCKS : 2; }

#if defined(__CUDACC__)

// Table layout: row idx = 8 copies of the same 16-byte entry (128 bytes = all 32 banks), one table
// for the diagonal constants and one for the match scores.  A 128-bit shared load is served in
// four phases of 8 consecutive lanes; lane l reads copy l & 7, so the 8 lanes of a phase always
// hit 8 different bank groups whatever their indices are: no bank conflicts (the first version,
// one 32-byte entry per index, ran at 96 % of the shared-memory pipe with 70 % conflict wavefronts).
// SH = log2 of the row size in bytes: 7, or 8 in the wide kernels, whose rows hold the entry for the normal
// length unit (8 copies) followed by the entry for length unit 0 (8 copies); `copy` then carries the mode bit.
template <int SH>
struct PwDevEW {
    const int *tbl;
    uint32_t mm;    // mismatch bits of the step at even positions (cells 0..15)
    uint32_t mm2;   // cells 16.. (wide kernels)
    uint32_t copy;  // (lane & 7) * 16 (| mode << 7): this lane's copy inside a row, in bytes
    __device__ __forceinline__ PwE4 operator()(int g) const {
        // row << SH | copy: one shift + one 3-input logic op
        const uint32_t w = g < 4 ? mm : mm2;
        const int sh = 8 * (g & 3) - SH;
        const uint32_t at = ((sh < 0 ? (w << (sh < 0 ? -sh : 0)) : (w >> (sh < 0 ? 0 : sh))) & (0x55u << SH)) | copy;
        const int4 d = *reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(tbl) + at);
        const int4 s = *reinterpret_cast<const int4 *>(reinterpret_cast<const char *>(tbl) + (NWP_TBL << SH) + at);
        PwE4 e;
        e.ds[0] = d.x; e.ds[1] = d.y; e.ds[2] = d.z; e.ds[3] = d.w;
        e.sb[0] = s.x; e.sb[1] = s.y; e.sb[2] = s.z; e.sb[3] = s.w;
        return e;
    }
};

// CL: slot of the last query column inside a lane's strip when every query read has the same length
// (a.q.fixed_len; CL = (ylen - 2) % S), -1 otherwise
template <int S, int CL>
__global__ void __launch_bounds__(nwp_threads(S), nwp_min_blocks(S)) nwp_kernel(NwArgs a) {
    constexpr bool WIDE = nwp_wide(S);
    constexpr int THREADS = nwp_threads(S), WARPS = THREADS / 32;
    constexpr int SH = WIDE ? 8 : 7;              // log2 of the table row size in bytes
    constexpr int ROW = (1 << SH) / 4;            // ... in ints
    __shared__ __align__(256) int tbl[2 * NWP_TBL * ROW];
    __shared__ uint8_t sx_all[WARPS * 2][PW_MAX_X];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hl = lane & 15, half = lane >> 4;
    const unsigned hmask = 0xFFFFu << (16 * half);
    uint8_t *sx = sx_all[warp * 2 + half];
    const uint32_t r_begin = a.range[0], r_end = a.range[1];
    const PwK k = pw_consts(a.igap, a.egap, a.one, PW_LEN1, a.pw_bias);
    const PwK k_id = pw_consts(a.igap, a.egap, a.one, 0, a.pw_bias);  // wide reads, second run: identities alone
    for (int e = threadIdx.x; e < NWP_TBL * (ROW / 4); e += THREADS) {
        // entry e: table index e / (ROW / 4), then (wide) 8 copies for k followed by 8 copies for k_id
        const bool second = WIDE && ((e >> 3) & 1);
        const PwE4 v = pw_e4(second ? k_id : k, (uint32_t)(e / (ROW / 4)));
#pragma unroll
        for (int c = 0; c < 4; c++) { tbl[e * 4 + c] = v.ds[c]; tbl[NWP_TBL * ROW + e * 4 + c] = v.sb[c]; }
    }
    __syncthreads();
    unsigned long long my_cells = 0, my_pairs = 0;
    // wide reads: `again` (warp-uniform) = this round re-runs the pairs whose statistics did not split, `mine` =
    // this half's pair is one of them, v_first = its V of the first run
    bool again = false, mine = false;
    uint32_t v_first = 0;
    struct { uint32_t idx, xs, ys, xlen, ylen, rd; unsigned long long key; } kept = {0, 0, 0, 0, 0, 0, 0};

    for (;;) {
        // each half warp takes the next pair that still has to be aligned
        bool have = false;
        uint32_t idx = 0, xs = 0, ys = 0, xlen = 0, ylen = 0, rd = 0;
        unsigned long long key = 0;
        if (WIDE && again) {
            have = mine;
            idx = kept.idx; xs = kept.xs; ys = kept.ys; xlen = kept.xlen; ylen = kept.ylen; rd = kept.rd; key = kept.key;
        } else for (;;) {
            if (hl == 0) idx = atomicAdd(a.work, 1u);
            idx = __shfl_sync(hmask, idx, 16 * half) + r_begin;
            if (idx >= r_end) break;
            const PairRec pr = a.pairs[idx];
            ys = read_start(a.q, pr.r);
            ylen = (a.q.fixed_len ? a.q.fixed_len : a.q.start[pr.r + 1] - ys);
            if (a.check_class && nw_class_of(ylen) != a.s_class) continue;
            xs = read_start(a.db, pr.s);
            xlen = (a.db.fixed_len ? a.db.fixed_len : a.db.start[pr.s + 1] - xs);
            if (a.mixed && !pw_pair_eligible(xlen, ylen, a.igap, a.egap, a.pw_bias)) continue;  // the generic kernel's pair
            // an earlier hit of this read is accepted?  best[] is lowered by other warps meanwhile: one lane of
            // the half reads it, so that all 16 lanes take the same branch
            int pruned_h = 0;
            if (a.best && hl == 0) pruned_h = pr.key >= a.best[pr.r];
            const bool pruned = __shfl_sync(hmask, pruned_h, 16 * half) != 0;
            if (pruned || xlen < 2 || ylen < 2 || xlen > (uint32_t)PW_MAX_X || ylen > (uint32_t)(PW_LANES * S + 1)) {
                if (hl == 0) {
                    PairRes z; z.score = pruned ? 0 : NW_NEG * 2; z.bx = z.by = 0; z.stats = 0;
                    a.res[idx] = z;
                    if (!pruned) { my_pairs++; }
                }
                continue;
            }
            rd = pr.r; key = pr.key; have = true;
            break;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, have)) break;
        const PwK &kk = (WIDE && again) ? k_id : k;  // warp-uniform
        const int X1 = have ? (int)xlen - 1 : 0, Y1 = have ? (int)ylen - 1 : 0;
        if (have && !(WIDE && again))
            for (uint32_t i = hl; i < xlen; i += 16) sx[i] = (uint8_t)base_at(a.db.pk, (uint64_t)xs + i);
        __syncwarp();
        const int nl = (Y1 + S - 1) / S;
        const int j0 = hl * S + 1;
        uint32_t x0 = 0, y0 = 0, ycols = 0, ycols2 = 0;
        PwLane<S> L;
        {
            uint64_t halo = 0;
            if (have) {
                const int64_t g = (int64_t)ys + j0 - 2;
                halo = g >= 0 ? fetch32(a.q.pk, (uint64_t)g) : (fetch32(a.q.pk, 0) << 2);
                x0 = sx[0];
                y0 = base_at(a.q.pk, ys);
            }
            ycols = (uint32_t)(halo >> 4);  // codes of Y[j0 .. j0+15]
            if (WIDE) ycols2 = (uint32_t)(halo >> 36);  // Y[j0+16 .. j0+29]
            pw_lane_init<S>(L, kk, x0, halo, hl == 0, hl);
        }
        const int cl = Y1 > 0 ? (Y1 - 1) % S : 0;
        const bool owns_last = (Y1 >= j0) && (Y1 < j0 + S);
        int steps = have ? X1 + nl - 1 : 0;
        {
            const int o = __shfl_xor_sync(0xffffffffu, steps, 16);
            steps = steps > o ? steps : o;
        }
        const uint32_t my_copy = (uint32_t)(lane & 7) * 16u | ((WIDE && again) ? 128u : 0u);
        PwLink out;
        out.a = out.b = out.mfz = out.lw = 0;
        // lane l of a half works on row t - l + 1; the two row histories swap roles with the step parity
#define IMS_PW_STEP(T_, PREV1, PREV2)                                                              \
    {                                                                                              \
        const int i = (T_) - hl + 1;                                                               \
        PwLink in;                                                                                 \
        in.a = __shfl_up_sync(0xffffffffu, out.a, 1, 16);                                          \
        in.b = __shfl_up_sync(0xffffffffu, out.b, 1, 16);                                          \
        in.mfz = __shfl_up_sync(0xffffffffu, out.mfz, 1, 16);                                      \
        in.lw = __shfl_up_sync(0xffffffffu, out.lw, 1, 16);                                        \
        const bool act = have && (hl < nl) && (i >= 1) && (i <= X1);                               \
        if (act) {                                                                                 \
            const uint32_t xi = sx[i];                                                             \
            if (hl == 0) in = pw_first_link(kk, xi, y0);                                           \
            const uint32_t d_ = ycols ^ (xi * 0x55555555u);                                        \
            PwDevEW<SH> ew;                                                                        \
            ew.tbl = tbl;                                                                          \
            ew.copy = my_copy;                                                                     \
            ew.mm = (d_ | (d_ >> 1));                                                              \
            ew.mm2 = 0;                                                                            \
            if (WIDE) {                                                                            \
                const uint32_t d2_ = ycols2 ^ (xi * 0x55555555u);                                  \
                ew.mm2 = (d2_ | (d2_ >> 1));                                                       \
            }                                                                                      \
            pw_row<S, PwDevEW<SH>, CL>(L, PREV1, PREV2, in, out, i, j0, ew, kk, X1, Y1, cl, owns_last);  \
        }                                                                                          \
    }
        // always an even number of steps (a surplus step has no active lane): the register roles of the
        // two row histories are then the same at every loop head and no row is ever copied
        for (int t = 0; t < steps; t += 2) {
            IMS_PW_STEP(t, L.r0, L.r1)
            IMS_PW_STEP(t + 1, L.r1, L.r0)
        }
#undef IMS_PW_STEP
        // the last row: lane hl wrote it at step X1 - 1 + hl
        if (have && hl < nl) {
            if ((X1 - 1 + hl) & 1) pw_last_row<S>(L, L.r0, j0, X1, Y1);
            else pw_last_row<S>(L, L.r1, j0, X1, Y1);
        }
        // reduction of the best border cell over the half warp ("last in row-major order" on ties)
        int bz = (have && hl < nl) ? L.bz : (int)0x80000000, bw = L.bw, bi = L.bi, bj = L.bj;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const int cz = __shfl_xor_sync(0xffffffffu, bz, o);
            const int cw = __shfl_xor_sync(0xffffffffu, bw, o);
            const int ci = __shfl_xor_sync(0xffffffffu, bi, o);
            const int cj = __shfl_xor_sync(0xffffffffu, bj, o);
            const bool better = cz > bz || (cz == bz && (ci > bi || (ci == bi && cj > bj)));
            if (better) { bz = cz; bw = cw; bi = ci; bj = cj; }
        }
        // length and identities of the best cell's path
        uint32_t len = 0, id = 0;
        bool done = true;
        if (WIDE) {  // (all 16 lanes hold the reduced cell and decide together)
            if (!again) {
                v_first = pw_stats(k, bw);
                mine = have && pw_split_stats(v_first, bi, bj, &len, &id);  // true: both splits possible
                done = !mine;
                kept.idx = idx; kept.xs = xs; kept.ys = ys; kept.xlen = xlen; kept.ylen = ylen; kept.rd = rd; kept.key = key;
            } else {
                id = pw_stats(k_id, bw);
                len = (v_first - id) >> 8;
                mine = false;
            }
        }
        if (hl == 0 && have && done) {
            if (!WIDE) { len = pw_len(k, bw); id = pw_ids(k, bw); }
            // src/alignmentFunctions.c:163 through the host-built exact tables
            const bool ok = len > 0 && len >= a.lmin[ylen] && id >= a.imin[len];
            PairRes z;
            z.score = pw_score(kk, bw); z.bx = (uint32_t)bi; z.by = (uint32_t)bj;
            z.stats = (len << 16) | id | (ok ? 0x80000000u : 0u);
            a.res[idx] = z;
            if (ok && a.best) atomicMin(&a.best[rd], key);
            my_cells += (unsigned long long)X1 * (unsigned long long)Y1;
            my_pairs++;
        }
        if (WIDE) again = !again && __any_sync(0xffffffffu, mine);
        __syncwarp();
    }
    if (hl == 0 && my_pairs) {
        atomicAdd(a.cells, my_cells);
        atomicAdd(a.cells + 2, my_pairs);  // counters[6]: pairs run through NW
    }
}

#endif  // __CUDACC__

}  // namespace imsame
