// qtable.cuh -- packing and K1: the query word table in HBM.
//
// The reference indexes the DATABASE (12-D pointer table of linked lists,
// src/IMSAME.c:232-281) and streams the query (src/alignmentFunctions.c:91-203).
// Here the mirror image is built: a CSR table over the QUERY words
//     off[4^k + 1], qpos[n_words]   (k = 12 unless imsame_gpu_set_kmer says otherwise)
//       (qpos = index of the word's last base = curr_pos)
// holding exactly the words the reference's scan would look up, including its
// cross-read "phantom" word: every read that is not the first of its pthread
// chunk starts its word stream on the LAST base of the previous read and never
// uses its own last base (src/alignmentFunctions.c:93-105,189-199).
//
// A table entry (QEntry, 24 bytes) carries everything the database scan needs to extend a seed hit
// on this word: the 32 query bases after and the 32 before it as bit planes (common.cuh) and the room
// left inside the read on both sides.  The scan then reads the entries of a bucket as ONE contiguous
// stream instead of chasing word position -> read bounds -> two query windows through three dependent
// random gathers per hit (round 1: 5.3x the algorithmic DRAM traffic, load latency the top stall).
#pragma once
#include "common.cuh"
#include "extend.cuh"

namespace imsame {

#if defined(__CUDACC__)

// ASCII (A/C/G/T) -> 2 bits per base; one thread per output word (16 bases, 128-bit load)
__global__ void pack_kernel(const uint8_t *__restrict__ in, uint64_t n_bases, uint32_t *__restrict__ out) {
    const uint64_t n_words = (n_bases + 15) / 16;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        if (w * 16 + 16 <= n_bases) {
            const uint4 q = reinterpret_cast<const uint4 *>(in)[w];
            const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // per byte: (c >> 1) & 3, gathered to 8 bits per 32-bit lane
                const uint32_t t = (x[k] >> 1) & 0x03030303u;
                const uint32_t g = (t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xFFu;
                v |= g << (8 * k);
            }
        } else {
            for (uint64_t i = w * 16; i < n_bases; i++) v |= (uint32_t)((in[i] >> 1) & 3u) << (2 * (i - w * 16));
        }
        out[w] = v;
    }
}

// blk[b] = read containing base 64*b ; one thread per read
__global__ void blk_kernel(const uint32_t *__restrict__ start, uint32_t n, uint32_t *__restrict__ blk) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t s = start[r], e = start[r + 1];
        for (uint32_t b = (s + 63) >> 6; ((uint64_t)b << 6) < e; b++) blk[b] = r;
    }
}

struct QEntry {
    uint32_t f_lo, f_hi;  // query bases e+1 .. e+32 (forward walk, src/alignmentFunctions.c:318-333)
    uint32_t b_lo, b_hi;  // query bases e-k, e-k-1, .. (backward walk, :342-357), bit u = base e-k-u
    uint32_t e;           // index of the word's last base (curr_pos)
    uint32_t lim;         // fq | bq1 << 16: steps left inside the read, forward (yend - e - 1) and
                          // backward + 1 (e - k + 2 - ys; 0 for the phantom word); ylen = fq + bq1 + k - 1
};
static_assert(sizeof(QEntry) == 24, "QEntry is read as three 8-byte words");
// Array of structs on purpose.  Storing the three 8-byte words in three planes makes every warp load a coalesced
// 256 bytes, but a bucket (~14 entries) then lies in three 114-byte pieces instead of one 343-byte piece, and DRAM
// is read in 128-byte lines: measured 487 GB instead of 340 GB of DRAM reads per 0.5 Gbase segment and 547 instead
// of 505 ms per cfg2 step (profiles/r02_ncu_scan_v12_full_summary.txt vs ..._v11_...).  The three loads of a lane
// hit the same sectors, so the first one brings them into L1 (hit rate 58 % of the ideal 67 %).

struct QTableArgs {
    SeqMap q;
    uint32_t per, n_threads;  // chunking of src/IMSAME.c:414,433
    uint32_t *cnt;            // pass 0: histogram ; pass 1: bucket cursors
    QEntry *qtab;
    int k;                    // seed length (the reference: FIXED_K = 12, src/structs.h:15)
    // Which of the words go into this table (capi.cu: "early words first").  The reference walks a read's words
    // left to right and stops at the first accepted hit (src/alignmentFunctions.c:172,189), so once the words that
    // end before position e_split of their read have been scanned and aligned, the later words of every read that
    // has an accepted hit by then can never matter: the second table simply does not contain them.
    //   0: every word   1: words with e_rel < e_split   2: words with e_rel >= e_split of reads without a key
    // e_rel = e - ys + 1, the k-mer end inside the read as the scan-order key counts it (common.cuh: make_key)
    int part;
    uint32_t e_split;
    const unsigned long long *best;  // part 2: per read scan-order key after the early bands (KEY_NONE = none)
};

// does a query word end at base e, and which word?  
__device__ __forceinline__ bool query_word_at(const QTableArgs &a, uint32_t e, uint32_t &code, uint32_t &ys,
                                              uint32_t &yend) {
    const uint32_t r = find_read(a.q, e);
    ys = read_start(a.q, r);
    yend = a.q.fixed_len ? ys + a.q.fixed_len : a.q.start[r + 1];
    const uint32_t lo = (ys == 0 || is_chunk_first(r, a.per, a.n_threads)) ? ys : ys - 1;
    const uint32_t hi = (r == a.q.n - 1) ? a.q.total - 1 : yend - 2;
    const uint32_t k1 = (uint32_t)a.k - 1u;
    if (e < lo + k1 || e > hi || yend < 2 + ys) return false;
    if (a.part) {
        const uint32_t e_rel = e - ys + 1;
        if (a.part == 1 ? e_rel >= a.e_split : (e_rel < a.e_split || a.best[r] != KEY_NONE)) return false;
    }
    code = fetch16(a.q.pk, (uint64_t)e - k1) & kmask_of(a.k);
    return true;
}

template <int PASS>
__global__ void qtable_kernel(QTableArgs a) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < a.q.total; e += gridDim.x * blockDim.x) {
        uint32_t code, ys, yend;
        if (!query_word_at(a, e, code, ys, yend)) continue;
        const uint32_t slot = atomicAdd(&a.cnt[code], 1u);
        if (PASS == 1) {
            const HitHalf h = query_half(a.q.pk, e, ys, yend, a.k);
            uint2 *dst = reinterpret_cast<uint2 *>(a.qtab + slot);
            dst[0] = make_uint2(h.f_lo, h.f_hi);
            dst[1] = make_uint2(h.b_lo, h.b_hi);
            dst[2] = make_uint2(e, (uint32_t)h.froom | ((uint32_t)(h.broom + 1) << 16));  // reads <= 32767 bases
        }
    }
}

// ---- exclusive scan of the 4^12 counters (three small kernels) ----
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        wsum[lane] = s;
    }
    __syncthreads();
    const uint32_t base = warp ? wsum[warp - 1] : 0;
    if (total) *total = wsum[31];
    __syncthreads();
    return base + x - v;
}

// phase 0: per-tile sums ; phase 2: write exclusive offsets (tile offset + local scan)
template <int PHASE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                  uint32_t *__restrict__ tile_sums,
                                                                  uint32_t *__restrict__ out) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;  // 4^16 counters: 64-bit
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
    if (PHASE == 0) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
    } else {
        ex += tile_sums[blockIdx.x];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) out[base + k] = ex;
            ex += v[k];
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = ex;
    }
}

// phase 1: exclusive scan of the tile sums in one block, SCAN_TILE sums at a time with a running carry
// (4^12 counters = 4096 tile sums = one round; 4^15 = 64 rounds)
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint32_t *tile_sums, uint32_t n_tiles) {
    uint32_t carry = 0;
    for (uint32_t chunk = 0; chunk < n_tiles; chunk += SCAN_TILE) {
        const uint32_t base = chunk + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (base + k < n_tiles) ? tile_sums[base + k] : 0u;
            s += v[k];
        }
        uint32_t total;
        uint32_t ex = carry + block_exclusive_scan(s, &total);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n_tiles) tile_sums[base + k] = ex;
            ex += v[k];
        }
        carry += total;
    }
}

#endif  // __CUDACC__

}  // namespace imsame
