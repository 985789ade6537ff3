// nw.cuh -- K3: one warp per candidate (read, db_seq) pair.  Full (xlen-1)(ylen-1)
// evaluation of the reference NW (src/alignmentFunctions.c:389-489) as an
// anti-diagonal wavefront over lane strips (nw_core.cuh), with the traceback's
// length/identities carried forward, the identity/coverage filter
// (src/alignmentFunctions.c:163) and the per-read first-accepted selection
// (:172,189 -> atomicMin of the scan-order key) fused into the epilogue.
// Integer ALU work only: no tensor cores (irregular max/select recurrence).
#pragma once
#include "nw_core.cuh"
#include "nwp_core.cuh"  // pw_pair_eligible: which pairs the packed-word kernel takes in a mixed run
#include "traceback.cuh"

namespace imsame {

constexpr int NW_WARPS = 8;  // warps per block
constexpr int NW_THREADS = NW_WARPS * 32;
constexpr int NW_XBUF = 3008;  // bytes of X codes per warp in shared memory

struct PairRec {
    uint32_t r, s;  // query read, database read (local to the segment)
    uint64_t key;   // smallest scan-order key among the pair's e-value-passing hits
};
struct PairRes {
    int32_t score;
    uint32_t bx, by;
    uint32_t stats;  // (length << 16) | identities; bit 31 = accepted by the filter
};

struct NwArgs {
    SeqMap db, q;
    const PairRec *pairs;
    PairRes *res;
    uint32_t *work;           // atomic work-queue head (relative to range[0])
    int igap, egap;
    const uint16_t *lmin, *imin;
    unsigned long long *best;  // per read scan-order key (atomicMin); may be null (nw_batch)
    unsigned long long *cells;  // [0] cells, [2] pairs evaluated
    NwLink *carry;             // 2 * MAX_READ links per warp of the grid, or null
    int s_class;               // check_class: pairs with nw_class_of(ylen) != s_class are skipped
    // Work range: pairs[range[0] .. range[1]) (device-resident offsets written by the binning
    // kernels).  The scan orders candidates into (NW class, k-mer-end band) bins and the bins of
    // a class are launched in ascending band order: the reference stops at a read's first
    // accepted hit (src/alignmentFunctions.c:172,189), so once an early band has accepted a
    // read, all its later candidates are pruned by the key comparison below instead of aligned.
    const uint32_t *range;
    int check_class;  // 1: unsorted explicit pairs (nw_batch / traceback): skip other classes here
    int one;          // 1, as a run-time value (nwp_core.cuh: pw_row)
    // 1: the work range is shared between the two kernels (a run whose longest reads do not fit packed words):
    // nwp_kernel takes the pairs that do (pw_eligible on the pair's own lengths), nw_kernel the others; each kernel
    // walks the whole range with its own work-queue head and skips the other's pairs without touching them
    int mixed;
    int pw_bias;      // score offset of the packed words of this run (nwp_core.cuh: pw_bias)
    // TB = true only (K4, winners-only traceback): back-pointer codes per cell
    uint16_t *tb;              // codes of pair idx start at tb + tb_off[idx]
    const uint64_t *tb_off;
};

// NW class of a query read = the bin its candidates are sorted into and the kernel instantiation that aligns them.
//   1..8   Y1 = ylen - 1 <= 255 columns: c = ceil(Y1 / 32); generic kernel c columns per lane, packed 2c per lane
//   9, 10  Y1 <= 288 / 320: packed-word kernel with 18 / 20 columns per lane (nwp_core.cuh, "wide reads"); pairs
//          it cannot take run in the generic kernel in two balanced passes of 5 columns per lane
//   longer reads: P = ceil(Y1 / 256) passes of the generic kernel, columns per lane chosen so that the passes are
//          balanced (class = ceil(Y1 / 32P) = 6..8).  With 8 columns per lane whatever the length, a 300-base
//          query read ran one full pass and one pass on 6 lanes of 32: 55 % of the lane-steps did work, now 85 %
constexpr int NW_CLASSES = 10;
IMS_HD int nw_class_of(uint32_t ylen) {
    const int y1 = (int)ylen - 1;
    if (y1 <= 255) { const int c = (y1 + 31) / 32; return c < 1 ? 1 : c; }
    if (y1 <= 288) return 9;
    if (y1 <= 320) return 10;
    const int passes = (y1 + 255) / 256;
    return (y1 + 32 * passes - 1) / (32 * passes);
}
// columns per lane of the generic kernel that aligns class c
IMS_HD int nw_class_cols(int c) { return c <= 8 ? c : 5; }

#if defined(__CUDACC__)

__device__ __forceinline__ NwLink shfl_up_link(const NwLink &o) {
    NwLink in;
    in.a = __shfl_up_sync(0xffffffffu, o.a, 1);
    in.ap = __shfl_up_sync(0xffffffffu, o.ap, 1);
    in.b = __shfl_up_sync(0xffffffffu, o.b, 1);
    in.bp = __shfl_up_sync(0xffffffffu, o.bp, 1);
    in.mfs = __shfl_up_sync(0xffffffffu, o.mfs, 1);
    in.mfy = __shfl_up_sync(0xffffffffu, o.mfy, 1);
    in.mfp = __shfl_up_sync(0xffffffffu, o.mfp, 1);
    return in;
}

template <int S, bool TB>
__global__ void __launch_bounds__(NW_THREADS) nw_kernel(NwArgs a) {
    __shared__ uint8_t sx_all[NW_WARPS][NW_XBUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *sx = sx_all[warp];
    const uint32_t r_begin = a.range[0], r_end = a.range[1];
    unsigned long long my_cells = 0, my_pairs = 0;
    NwLink *carry0 = a.carry ? a.carry + (size_t)(blockIdx.x * NW_WARPS + warp) * 2 * MAX_READ : nullptr;

    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(a.work, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0) + r_begin;
        if (idx >= r_end) break;
        const PairRec pr = a.pairs[idx];
        const uint32_t ys = read_start(a.q, pr.r);
        const uint32_t ylen = (a.q.fixed_len ? a.q.fixed_len : a.q.start[pr.r + 1] - ys);
        if (a.check_class && nw_class_of(ylen) != a.s_class) continue;
        if (a.mixed) {  // the packed-word kernel's pair?
            const uint32_t xs_ = read_start(a.db, pr.s);
            const uint32_t xlen_ = (a.db.fixed_len ? a.db.fixed_len : a.db.start[pr.s + 1] - xs_);
            if (pw_pair_eligible(xlen_, ylen, a.igap, a.egap, a.pw_bias)) continue;
        }
        // an earlier hit of this read is already accepted?  best[] is lowered by other warps meanwhile: one lane
        // reads it, so that the whole warp takes the same branch
        int pruned_w = 0;
        if (a.best && lane == 0) pruned_w = pr.key >= a.best[pr.r];
        if (__shfl_sync(0xffffffffu, pruned_w, 0)) {
            if (lane == 0) { PairRes z; z.score = 0; z.bx = z.by = 0; z.stats = 0; a.res[idx] = z; }
            continue;
        }
        const uint32_t xs = read_start(a.db, pr.s);
        const uint32_t xlen = (a.db.fixed_len ? a.db.fixed_len : a.db.start[pr.s + 1] - xs);
        if (xlen > (uint32_t)MAX_READ || ylen > (uint32_t)MAX_READ) {
            // the reference aborts here (src/alignmentFunctions.c:155); the host reports IMSAME_EREADSIZE
            if (lane == 0) { PairRes z; z.score = 0; z.bx = z.by = 0; z.stats = 0; a.res[idx] = z; }
            continue;
        }
        const int X1 = (int)xlen - 1, Y1 = (int)ylen - 1;
        uint16_t *tb_pair = nullptr;
        uint32_t tb_str = 0;
        if (TB) {
            tb_pair = a.tb + a.tb_off[idx];
            tb_str = tb_stride(ylen, S);
        }
        __syncwarp();
        for (uint32_t i = lane; i < xlen; i += 32) sx[i] = (uint8_t)base_at(a.db.pk, (uint64_t)xs + i);
        __syncwarp();
        const uint32_t x0 = sx[0];
        const uint32_t y0 = base_at(a.q.pk, ys);

        NwBest wbest;
        wbest.s = NW_NEG * 2; wbest.i = wbest.j = wbest.p = 0;
        int pass = 0;
        for (int jb = 0; jb < Y1; jb += 32 * S, pass++) {
            int nl = (Y1 - jb + S - 1) / S;
            nl = nl > 32 ? 32 : nl;
            const bool more = jb + 32 * S < Y1;
            const int j0 = jb + lane * S + 1;
            const bool first = (jb == 0) && (lane == 0);
            // codes of Y[j0-2 .. j0+S-1]
            uint64_t halo;
            {
                const int64_t g = (int64_t)ys + j0 - 2;
                halo = g >= 0 ? fetch32(a.q.pk, (uint64_t)g) : (fetch32(a.q.pk, 0) << 2);
            }
            const uint32_t ycols = (uint32_t)(halo >> 4) & ((1u << (2 * S)) - 1u);
            NwLane<S> L;
            nw_lane_init<S>(L, x0, halo, first, lane);
            const int cl = (Y1 - 1 - jb) % S;  // slot of the last column inside its strip (warp-uniform)
            const bool owns_last = (Y1 >= j0) && (Y1 < j0 + S);
            NwLink out;
            out.a = out.ap = out.b = out.bp = out.mfs = out.mfy = out.mfp = 0;
            const NwLink *cin = carry0 ? carry0 + (size_t)((pass + 1) & 1) * MAX_READ : nullptr;
            NwLink *cout = carry0 ? carry0 + (size_t)(pass & 1) * MAX_READ : nullptr;
            const int steps = X1 + nl - 1;
            // lane l works on row t - l + 1; the two row histories swap roles with the step parity
#define IMS_NW_STEP(T_, PREV1, PREV2)                                                                   \
    {                                                                                                   \
        const int i = (T_) - lane + 1;                                                                  \
        NwLink in = shfl_up_link(out);                                                                  \
        const bool act = (lane < nl) && (i >= 1) && (i <= X1);                                          \
        if (act) {                                                                                      \
            const uint32_t xi = sx[i];                                                                  \
            if (lane == 0) in = (jb == 0) ? nw_first_link(xi, y0) : cin[i];                             \
            const uint32_t d_ = ycols ^ (xi * 0x55555555u);                                             \
            const uint32_t mm = (d_ | (d_ >> 1)) & 0x55555555u;                                         \
            nw_row<S, TB>(L, PREV1, PREV2, in, out, i, j0, mm, a.igap, a.egap, X1, Y1, cl, owns_last,   \
                          first, TB ? tb_pair + (size_t)(i - 1) * tb_str + (j0 - 1) : nullptr);         \
            if (more && lane == 31) cout[i] = out;                                                      \
        }                                                                                               \
    }
            for (int t = 0; t < steps; t += 2) {
                IMS_NW_STEP(t, L.r0, L.r1)
                if (t + 1 < steps) IMS_NW_STEP(t + 1, L.r1, L.r0)
            }
#undef IMS_NW_STEP
            // warp reduction of the best border cell ("last in row-major order" on ties)
            NwBest b = L.best;
            if (lane >= nl) b.s = NW_NEG * 2;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                NwBest c;
                c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
                c.i = __shfl_xor_sync(0xffffffffu, b.i, o);
                c.j = __shfl_xor_sync(0xffffffffu, b.j, o);
                c.p = __shfl_xor_sync(0xffffffffu, b.p, o);
                if (best_better(c, b)) b = c;
            }
            if (best_better(b, wbest)) wbest = b;
            __syncwarp();
        }
        if (lane == 0) {
            const uint32_t len = nw_stat_len(wbest.p), id = nw_stat_ids(wbest.p);
            // src/alignmentFunctions.c:163 through the host-built exact tables
            const bool ok = (X1 >= 1 && Y1 >= 1) && len > 0 && len >= a.lmin[ylen] && id >= a.imin[len];
            PairRes z;
            z.score = wbest.s; z.bx = (uint32_t)wbest.i; z.by = (uint32_t)wbest.j;
            z.stats = (len << 16) | id | (ok ? 0x80000000u : 0u);
            a.res[idx] = z;
            if (ok && a.best) atomicMin(&a.best[pr.r], (unsigned long long)pr.key);
            my_cells += (unsigned long long)X1 * (unsigned long long)Y1;
            my_pairs++;
        }
    }
    if (lane == 0 && my_pairs) {
        atomicAdd(a.cells, my_cells);
        atomicAdd(a.cells + 2, my_pairs);  // counters[6]: pairs run through NW
    }
}

// K4b: one thread per winner walks the stored back-pointers (traceback.cuh)
static __global__ void tb_walk_kernel(const PairRes *res, const uint16_t *tb, const uint64_t *tb_off,
                               const uint32_t *strides, uint32_t n_pairs, uint32_t *ops,
                               const uint64_t *ops_off, uint32_t *n_ops, uint32_t *end_xy) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        uint32_t ex = 0, ey = 0;
        n_ops[i] = tb_walk(tb + tb_off[i], strides[i], res[i].bx, res[i].by, ops + ops_off[i], &ex, &ey);
        end_xy[2 * i] = ex;
        end_xy[2 * i + 1] = ey;
    }
}

// K4c: the walks leave their run-length ops in per-pair slots of xlen + ylen + 2 words of which a handful are
// used; one warp per pair moves them to their place in the compact list (offsets = exclusive scan of n_ops), so
// that only the ops themselves cross PCIe (cfg2: 12 MB instead of 1 GB)
static __global__ void compact_ops_kernel(const uint32_t *__restrict__ ops, const uint64_t *__restrict__ ops_off,
                                          const uint32_t *__restrict__ n_ops, const uint32_t *__restrict__ dst_off,
                                          uint32_t n_pairs, uint32_t *__restrict__ dst) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n_pairs; i += n_warps) {
        const uint32_t n = n_ops[i];
        const uint32_t *src = ops + ops_off[i];
        uint32_t *d = dst + dst_off[i];
        for (uint32_t k = lane; k < n; k += 32) d[k] = src[k];
    }
}

#endif  // __CUDACC__

}  // namespace imsame
