// scan.cuh -- K2: database scan.  Streams the 2-bit database, looks every
// database word up in the query word table (qtable.cuh), runs the ungapped
// extension + integer e-value test of src/alignmentFunctions.c:276-387 on
// every (database word, query word) hit, and records the candidates that pass
// (src/alignmentFunctions.c:139) in a pair table keyed by (read, db_seq) with
// an atomicMin of the scan-order key.  K2b compacts that table into the work
// queue of the NW kernel with warp-aggregated atomics.
//
// Which database words exist (src/IMSAME.c:215-283): words never span two
// reads (word_size = 0 at :283) nor a dropped non-ACGT character (:229-231);
// pos = index AFTER the word's last base (:247,265).
#pragma once
#include "common.cuh"
#include "extend.cuh"
#include "nw.cuh"
#include "qtable.cuh"

namespace imsame {

constexpr uint64_t HASH_EMPTY = ~0ull;
constexpr int SCAN_WARPS = 8;
constexpr int SCAN_THREADS_K2 = SCAN_WARPS * 32;

struct ScanArgs {
    SeqMap db, q;
    const uint32_t *off;   // query table: bucket offsets (4^12 + 1)
    const QEntry *qtab;    // query table: one entry per query word, bucket by bucket (qtable.cuh)
    const uint32_t *brk;   // database word breaks (segment-local), ascending
    uint32_t n_brk;
    const uint16_t *nmin;  // e-value threshold table by ylen
    const uint32_t *lut;   // extension walk table (extend.cuh: build_ext_lut3), EXT_LUT3_SIZE words
    uint64_t seg_pos_base; // global index of the segment's first base
    unsigned long long *hkeys, *hvals;  // pair table (open addressing)
    uint32_t hmask;
    const unsigned long long *best;  // current per-read best key (prunes later segments)
    unsigned long long *counters;    // [0] db words, [1] hits, [2] e-value passes, [3] anomalies
    int *overflow;
    int k;  // seed length when the kernel is not specialised on it (scan_kernel<0>)
    int count_words;  // 0: counters[0] has the database words already (second pass of a two-pass run)
};

IMS_HD uint64_t pair_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void pair_insert(const ScanArgs &a, uint32_t r, uint32_t s, uint64_t key) {
    const unsigned long long pk = ((unsigned long long)r << 32) | s;
    uint32_t h = (uint32_t)pair_hash(pk) & a.hmask;
    for (uint32_t probe = 0; probe <= a.hmask; probe++) {
        unsigned long long cur = a.hkeys[h];
        if (cur == HASH_EMPTY) cur = atomicCAS(&a.hkeys[h], HASH_EMPTY, pk);
        if (cur == HASH_EMPTY || cur == pk) {
            atomicMin(&a.hvals[h], (unsigned long long)key);
            return;
        }
        h = (h + 1) & a.hmask;
        if (probe > 4096) break;
    }
    *a.overflow = 1;
}

// is there a word break inside the word ending at base q (bases q-k+1 .. q)?
__device__ __forceinline__ bool word_broken(const ScanArgs &a, uint32_t q, int k) {
    if (a.n_brk == 0) return false;
    // first break >= q-k+2
    uint32_t lo = 0, hi = a.n_brk;
    const uint32_t want = q - (uint32_t)(k - 2);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.brk[mid] < want) lo = mid + 1; else hi = mid;
    }
    return lo < a.n_brk && a.brk[lo] <= q;
}

// pos / fixed_len without a division: q = mulhi(pos, floor(2^32 / L)) is the quotient or one less
// (pos < 2^32), one conditional increment makes it exact.  inv = 0: ragged reads, block lookup.
__device__ __forceinline__ uint32_t find_read_inv(const SeqMap &m, uint32_t inv, uint32_t pos) {
    if (!inv) return find_read(m, pos);
    uint32_t q = __umulhi(pos, inv);
    if (pos - q * m.fixed_len >= m.fixed_len) q++;
    return q;
}
__device__ __forceinline__ uint32_t inv_of(uint32_t fixed_len) {
    return fixed_len >= 2 ? (uint32_t)((1ull << 32) / fixed_len) : 0u;
}

// A walk that is still alive after its first window in either direction (a true overlap: up to a whole read
// both ways; or one of the ~25 % of random hits that survive 32 bases) is parked in a per-warp queue.  The
// queue is worked off 32 walks at a time, ONE further window per walk and round; what is still alive goes
// back into the queue.  Every lane of a round has a window to do, however different the walks' lengths are
// (round 1 ran each batch of 32 to completion: half of the lanes idle behind the batch's longest walk).
constexpr int SCAN_QCAP = 64;

struct ParkedWalk {
    uint32_t p, e, ylen;
    ExtState st;
};
// seven words per walk, stored field by field (7 arrays of SCAN_QCAP words): consecutive lanes touch
// consecutive banks
constexpr int PARK_FIELDS = 7;
__device__ __forceinline__ void park_store(uint32_t (*q)[SCAN_QCAP], int slot, const ParkedWalk &w) {
    q[0][slot] = w.p; q[1][slot] = w.e;
    q[2][slot] = (uint32_t)w.st.run; q[3][slot] = (uint32_t)w.st.best;
    q[4][slot] = (uint32_t)w.st.pos_f | ((uint32_t)w.st.idn2 << 15);                 // <= 32767 | <= 65534
    q[5][slot] = (uint32_t)(w.st.fmax < 0 ? 0 : w.st.fmax) | ((uint32_t)(w.st.bmax < 0 ? 0 : w.st.bmax) << 16);
    q[6][slot] = w.ylen | ((uint32_t)w.st.phase << 15) | ((uint32_t)w.st.t << 11);  // t: multiple of 32, < 32768
}
__device__ __forceinline__ ParkedWalk park_load(const uint32_t (*q)[SCAN_QCAP], int slot) {
    ParkedWalk w;
    w.p = q[0][slot]; w.e = q[1][slot];
    w.st.run = (int)q[2][slot]; w.st.best = (int)q[3][slot];
    const uint32_t pi = q[4][slot], fb = q[5][slot], yt = q[6][slot];
    w.st.pos_f = (int)(pi & 0x7FFFu); w.st.idn2 = (int)(pi >> 15);
    w.st.fmax = (int)(fb & 0xFFFFu); w.st.bmax = (int)(fb >> 16);
    w.ylen = yt & 0x7FFFu; w.st.phase = (int)((yt >> 15) & 1u); w.st.t = (int)((yt >> 16) << 5);
    return w;
}

// the e-value test (src/alignmentFunctions.c:139) on a finished walk, through the integer table nmin[ylen]
// (host/thresholds.c); the query read itself is only looked up for the ~0.3 % of the hits that pass.
// thr_fixed: nmin of the one read length when all query reads have the same, else -1.
__device__ __forceinline__ void finish_hit(const ScanArgs &a, uint32_t q_inv, int thr_fixed, uint32_t p, uint32_t e,
                                           uint32_t ylen, const ExtState &st, unsigned long long &c_pass,
                                           unsigned long long &c_anom, int k) {
    const int n = ext_result(st, k);
    if (n >= 0 && n < (thr_fixed >= 0 ? thr_fixed : (int)a.nmin[ylen])) return;
    if (n < 0) c_anom++;  // the reference's unsigned wrap (:373) would make this pass; unreachable
    const uint32_t r = find_read_inv(a.q, q_inv, e);
    const uint32_t ys = read_start(a.q, r);
    c_pass++;
    const uint64_t key = make_key(e - ys + 1, a.seg_pos_base + p);
    if (key < a.best[r]) {
        const uint32_t s = find_read(a.db, p - 1);
        pair_insert(a, r, s, key);
    }
}

// one round: the last (up to) 32 parked walks advance by one window each; returns the new queue length
__device__ __forceinline__ int drain_round(const ScanArgs &a, uint32_t q_inv, int thr_fixed, const ExtXY *s_lut,
                                           uint32_t (*q)[SCAN_QCAP], int n_parked, int lane,
                                           unsigned long long &c_pass, unsigned long long &c_anom, int k) {
    const int take = n_parked < 32 ? n_parked : 32, first = n_parked - take;
    ParkedWalk w;
    w.st.phase = 2;
    w.p = w.e = w.ylen = 0;
    if (lane < take) w = park_load(q, first + lane);
    __syncwarp();  // every slot has been read before any is overwritten
    if (lane < take) {
        ext_window(w.st, s_lut, a.db.pk, a.q.pk, w.p, w.e, k);
        if (w.st.phase == 2) finish_hit(a, q_inv, thr_fixed, w.p, w.e, w.ylen, w.st, c_pass, c_anom, k);
    }
    const bool again = lane < take && w.st.phase < 2;
    const unsigned m = __ballot_sync(0xffffffffu, again);
    if (again) park_store(q, first + __popc(m & ((1u << lane) - 1u)), w);
    __syncwarp();
    return first + __popc(m);
}

// KT: the seed length as a compile-time constant (12 = the reference's FIXED_K, the only value it has),
// or 0 = read it from the arguments (imsame_gpu_set_kmer)
//
// A warp takes 32 consecutive database positions.  Per position with a non-empty bucket: the bucket and the
// database windows around the word as bit planes (common.cuh), kept in shared memory -- they are the same
// for every hit of the position.  The ~14 hits per position are then spread evenly over the lanes, two hits
// per lane per iteration: the position a hit belongs to comes from a 64-bit mask of the bucket starts inside
// the iteration's 64 hit slots (two warp-wide OR reductions + population counts; round 1 searched a prefix
// array with five dependent shared-memory loads per hit).  A hit reads its 24-byte table entry (consecutive
// lanes read consecutive entries: a streaming access), forms both first-window mismatch masks with four logic
// operations and walks them through the max-plus table (extend.cuh: ext_first2).  Walks that are still alive
// after their first window are parked (above).
template <int KT>
__global__ void __launch_bounds__(SCAN_THREADS_K2) scan_kernel(ScanArgs a) {
    const int k = KT ? KT : a.k;
    __shared__ uint32_t s_park[SCAN_WARPS][PARK_FIELDS][SCAN_QCAP];
    // per non-empty position of the warp's tile, in position order
    __shared__ uint32_t s_excl[SCAN_WARPS][32];  // first hit slot of the position's bucket
    __shared__ uint32_t s_b0[SCAN_WARPS][32];    // first table entry of the bucket
    __shared__ uint32_t s_p[SCAN_WARPS][32];     // llpos.pos: index after the word (src/IMSAME.c:247,265)
    __shared__ uint2 s_wf[SCAN_WARPS][32];       // database forward window (lo, hi planes)
    __shared__ uint2 s_wb[SCAN_WARPS][32];       // database backward window
    __shared__ int2 s_room[SCAN_WARPS][32];      // steps left inside the database read: forward, backward
    __shared__ __align__(8) ExtXY s_lut[EXT_LUT3_SIZE];
    for (int i = threadIdx.x; i < EXT_LUT3_SIZE; i += SCAN_THREADS_K2)
        s_lut[(i & ~0xFF) | (int)ext_fold8((uint32_t)i & 0xFFu)] = ext_xy_of(a.lut[i]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    const uint32_t q_inv = inv_of(a.q.fixed_len), db_inv = inv_of(a.db.fixed_len);
    const int thr_fixed = a.q.fixed_len ? (int)a.nmin[a.q.fixed_len] : -1;
    const uint32_t n_tiles = (a.db.total + 31) / 32;
    const uint32_t gw = blockIdx.x * SCAN_WARPS + warp, nw = gridDim.x * SCAN_WARPS;
    unsigned long long c_words = 0, c_hits = 0, c_pass = 0, c_anom = 0;
    int n_parked = 0;  // warp-uniform

    for (uint32_t tile = gw; tile < n_tiles; tile += nw) {
        // pair table too small: the host grows it and reruns.  One lane reads the flag (other warps set it
        // concurrently): the decision must be the same in every lane of the warp
        if (__shfl_sync(0xffffffffu, lane == 0 ? *(volatile int *)a.overflow : 0, 0)) break;
        const uint32_t q = tile * 32 + lane;  // index of the word's last base
        uint32_t cnt = 0, b0 = 0, xs = 0, xe = 0;
        if (q < a.db.total) {
            const uint32_t s = find_read_inv(a.db, db_inv, q);
            xs = read_start(a.db, s);
            xe = a.db.fixed_len ? xs + a.db.fixed_len : a.db.start[s + 1];
            if (q >= xs + (uint32_t)(k - 1) && !word_broken(a, q, k)) {
                const uint32_t code = fetch16(a.db.pk, (uint64_t)q - (uint32_t)(k - 1)) & kmask_of(k);
                b0 = a.off[code];
                cnt = a.off[(size_t)code + 1] - b0;
                c_words++;
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned ne = __ballot_sync(0xffffffffu, cnt != 0);
        if (cnt) {
            const int slot = __popc(ne & lt_mask);
            const HitHalf h = db_half(a.db.pk, q + 1, xs, xe, k);
            s_excl[warp][slot] = incl - cnt;
            s_b0[warp][slot] = b0;
            s_p[warp][slot] = q + 1;
            s_wf[warp][slot] = make_uint2(h.f_lo, h.f_hi);
            s_wb[warp][slot] = make_uint2(h.b_lo, h.b_hi);
            s_room[warp][slot] = make_int2(h.froom, h.broom);
        }
        __syncwarp();
        // lane i keeps the first hit slot of the i-th non-empty position
        const uint32_t my_start = lane < __popc(ne) ? s_excl[warp][lane] : 0xFFFFFFFFu;
        int n_before = 0;  // non-empty positions whose bucket starts before h0 (warp-uniform)
        for (uint32_t h0 = 0; h0 < total; h0 += 64) {
            // bit j of (m_lo, m_hi): a bucket starts at hit slot h0 + j
            const uint32_t d = my_start - h0;
            const uint32_t m_lo = __reduce_or_sync(0xffffffffu, d < 32u ? 1u << d : 0u);
            const uint32_t m_hi = __reduce_or_sync(0xffffffffu, (d - 32u) < 32u ? 1u << (d - 32u) : 0u);
            const int c_lo = __popc(m_lo);
            // position of a hit slot = (number of bucket starts at or before it) - 1; slots past the last hit
            // fall into the last bucket and are clamped to the last hit
            const int oa = n_before + __popc(m_lo & le_mask) - 1;
            const int ob = n_before + c_lo + __popc(m_hi & le_mask) - 1;
            n_before += c_lo + __popc(m_hi);
            const uint32_t ha = h0 + lane, hb = h0 + 32 + lane;
            const bool la = ha < total, lb = hb < total;
            const uint32_t sa_h = la ? ha : total - 1, sb_h = lb ? hb : total - 1;
            const uint2 *qa = reinterpret_cast<const uint2 *>(a.qtab + (s_b0[warp][oa] + (sa_h - s_excl[warp][oa])));
            const uint2 *qb = reinterpret_cast<const uint2 *>(a.qtab + (s_b0[warp][ob] + (sb_h - s_excl[warp][ob])));
            // streamed once per database position that hits the bucket: no reuse worth a cache line
            const uint2 fa = __ldcs(qa), ba = __ldcs(qa + 1), ma = __ldcs(qa + 2);
            const uint2 fb = __ldcs(qb), bb = __ldcs(qb + 1), mb2 = __ldcs(qb + 2);
            const uint2 dfa = s_wf[warp][oa], dba = s_wb[warp][oa], dfb = s_wf[warp][ob], dbb = s_wb[warp][ob];
            const int2 rma = s_room[warp][oa], rmb = s_room[warp][ob];
            const uint32_t ea = ma.x, eb = mb2.x;
            const uint32_t pa = s_p[warp][oa], pb = s_p[warp][ob];
            ExtState sta, stb;
            // mismatch bits of the first window each way (step u <-> bit u); steps past a read end are
            // mismatches (extend.cuh: hit_first_masks = ext_init + ext_first_masks on the two halves)
            uint32_t mfa, mba, mfb, mbb;
            {
                HitHalf da, qa_h, dbh, qb_h;
                da.f_lo = dfa.x; da.f_hi = dfa.y; da.b_lo = dba.x; da.b_hi = dba.y; da.froom = rma.x; da.broom = rma.y;
                dbh.f_lo = dfb.x; dbh.f_hi = dfb.y; dbh.b_lo = dbb.x; dbh.b_hi = dbb.y; dbh.froom = rmb.x; dbh.broom = rmb.y;
                qa_h.f_lo = fa.x; qa_h.f_hi = fa.y; qa_h.b_lo = ba.x; qa_h.b_hi = ba.y;
                qa_h.froom = (int)(ma.y & 0xFFFFu); qa_h.broom = (int)(ma.y >> 16) - 1;
                qb_h.f_lo = fb.x; qb_h.f_hi = fb.y; qb_h.b_lo = bb.x; qb_h.b_hi = bb.y;
                qb_h.froom = (int)(mb2.y & 0xFFFFu); qb_h.broom = (int)(mb2.y >> 16) - 1;
                hit_first_masks(da, qa_h, sta, mfa, mba);
                hit_first_masks(dbh, qb_h, stb, mfb, mbb);
            }
            ext_first2(sta, stb, s_lut, mfa, mba, mfb, mbb, k);
            c_hits += (la ? 1 : 0) + (lb ? 1 : 0);
            const uint32_t ylen_a = (ma.y & 0xFFFFu) + (ma.y >> 16) + (uint32_t)(k - 1);
            const uint32_t ylen_b = (mb2.y & 0xFFFFu) + (mb2.y >> 16) + (uint32_t)(k - 1);
            if (la && sta.phase == 2) finish_hit(a, q_inv, thr_fixed, pa, ea, ylen_a, sta, c_pass, c_anom, k);
            if (lb && stb.phase == 2) finish_hit(a, q_inv, thr_fixed, pb, eb, ylen_b, stb, c_pass, c_anom, k);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const bool unfinished = half ? (lb && stb.phase < 2) : (la && sta.phase < 2);
                const unsigned park = __ballot_sync(0xffffffffu, unfinished);
                if (park) {
                    if (unfinished) {
                        ParkedWalk w;
                        w.p = half ? pb : pa; w.e = half ? eb : ea; w.ylen = half ? ylen_b : ylen_a; w.st = half ? stb : sta;
                        park_store(s_park[warp], n_parked + __popc(park & lt_mask), w);
                    }
                    n_parked += __popc(park);
                    __syncwarp();
                    while (n_parked >= 32)
                        n_parked = drain_round(a, q_inv, thr_fixed, s_lut, s_park[warp], n_parked, lane, c_pass, c_anom, k);
                }
            }
        }
        __syncwarp();  // the next tile overwrites the per-position shared arrays
    }
    while (n_parked > 0)
        n_parked = drain_round(a, q_inv, thr_fixed, s_lut, s_park[warp], n_parked, lane, c_pass, c_anom, k);
    // counters: warp-reduce then one atomic per warp
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c_words += __shfl_xor_sync(0xffffffffu, c_words, o);
        c_hits += __shfl_xor_sync(0xffffffffu, c_hits, o);
        c_pass += __shfl_xor_sync(0xffffffffu, c_pass, o);
        c_anom += __shfl_xor_sync(0xffffffffu, c_anom, o);
    }
    if (lane == 0) {
        if (a.count_words) atomicAdd(&a.counters[0], c_words);
        atomicAdd(&a.counters[1], c_hits);
        atomicAdd(&a.counters[2], c_pass);
        if (c_anom) atomicAdd(&a.counters[3], c_anom);
    }
}

// K2b: pair table -> work queue ordered by (NW class, k-mer-end band).  Three small kernels:
// histogram of the occupied slots per bin, exclusive scan of the NBINS counters, scatter with
// block-aggregated atomics (one global atomic per bin per block); the scatter also resets the table.
constexpr int NW_BANDS = 32;
constexpr int NW_NBINS = (NW_CLASSES + 1) * NW_BANDS;  // classes 1..NW_CLASSES (nw_class_of)
constexpr int BIN_THREADS = 256;
constexpr int BIN_ITEMS = 16;  // slots per thread

struct BinArgs {
    unsigned long long *hkeys, *hvals;
    uint32_t n_slots;
    SeqMap q;
    uint32_t band_width;  // k-mer-end positions per band
    uint32_t *bin_count;  // NW_NBINS
    uint32_t *bin_off;    // NW_NBINS + 1 (exclusive offsets), then NW_NBINS cursors
    PairRec *pairs;
};

__device__ __forceinline__ int pair_bin(const BinArgs &a, unsigned long long pk, unsigned long long key) {
    const uint32_t r = (uint32_t)(pk >> 32);
    const uint32_t ylen = a.q.fixed_len ? a.q.fixed_len : a.q.start[r + 1] - a.q.start[r];
    uint32_t band = key_erel(key) / a.band_width;
    band = band < (uint32_t)NW_BANDS ? band : (uint32_t)NW_BANDS - 1;
    return nw_class_of(ylen) * NW_BANDS + (int)band;
}

template <int PASS>
__global__ void __launch_bounds__(BIN_THREADS) bin_kernel(BinArgs a) {
    __shared__ uint32_t s_cnt[NW_NBINS];
    __shared__ uint32_t s_base[NW_NBINS];
    for (int i = threadIdx.x; i < NW_NBINS; i += BIN_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const uint32_t chunk = BIN_THREADS * BIN_ITEMS;
    for (uint32_t base = blockIdx.x * chunk; base < a.n_slots; base += gridDim.x * chunk) {
        unsigned long long k[BIN_ITEMS], v[BIN_ITEMS];
        int bin[BIN_ITEMS];
        uint32_t loc[BIN_ITEMS];
#pragma unroll
        for (int t = 0; t < BIN_ITEMS; t++) {
            const uint32_t i = base + t * BIN_THREADS + threadIdx.x;
            k[t] = i < a.n_slots ? a.hkeys[i] : HASH_EMPTY;
            bin[t] = -1;
            if (k[t] != HASH_EMPTY) {
                v[t] = a.hvals[i];
                bin[t] = pair_bin(a, k[t], v[t]);
                loc[t] = atomicAdd(&s_cnt[bin[t]], 1u);
                if (PASS == 1) { a.hkeys[i] = HASH_EMPTY; a.hvals[i] = ~0ull; }
            }
        }
        __syncthreads();
        if (PASS == 0) continue;  // counts are flushed once at the end
        for (int i = threadIdx.x; i < NW_NBINS; i += BIN_THREADS) {
            s_base[i] = s_cnt[i] ? atomicAdd(&a.bin_off[NW_NBINS + 1 + i], s_cnt[i]) : 0u;
            s_cnt[i] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < BIN_ITEMS; t++)
            if (bin[t] >= 0) {
                PairRec pr;
                pr.r = (uint32_t)(k[t] >> 32);
                pr.s = (uint32_t)k[t];
                pr.key = v[t];
                a.pairs[s_base[bin[t]] + loc[t]] = pr;
            }
        __syncthreads();
    }
    if (PASS == 0)
        for (int i = threadIdx.x; i < NW_NBINS; i += BIN_THREADS)
            if (s_cnt[i]) atomicAdd(&a.bin_count[i], s_cnt[i]);
}

// exclusive scan of the bin counters; initialises the scatter cursors and n_pairs, and writes the
// work range of every (class, launch) : bands are merged into fewer launches when there are too
// few candidates to fill the GPU per band (min_per_launch), later launches of a class stay empty.
__global__ void bin_offsets_kernel(const uint32_t *bin_count, uint32_t *bin_off, uint32_t *launch_range,
                                   uint32_t *n_pairs, uint32_t min_per_launch, uint32_t base) {
    if (threadIdx.x == 0) {
        uint32_t acc = base;  // this segment's candidates follow those of the earlier segments
        for (int i = 0; i < NW_NBINS; i++) {
            bin_off[i] = acc;
            bin_off[NW_NBINS + 1 + i] = acc;
            acc += bin_count[i];
        }
        bin_off[NW_NBINS] = acc;
        *n_pairs = acc - base;
        uint32_t launches = (acc - base) / (min_per_launch ? min_per_launch : 1u);
        launches = launches < 1u ? 1u : (launches > (uint32_t)NW_BANDS ? (uint32_t)NW_BANDS : launches);
        const uint32_t merge = ((uint32_t)NW_BANDS + launches - 1) / launches;  // bands per launch
        for (int c = 0; c <= NW_CLASSES; c++)
            for (uint32_t b = 0; b < (uint32_t)NW_BANDS; b++) {
                const uint32_t lo = b * merge, hi = (b + 1) * merge;
                const uint32_t b0 = lo < (uint32_t)NW_BANDS ? lo : (uint32_t)NW_BANDS;
                const uint32_t b1 = hi < (uint32_t)NW_BANDS ? hi : (uint32_t)NW_BANDS;
                launch_range[2 * (c * NW_BANDS + b)] = bin_off[c * NW_BANDS + b0];
                launch_range[2 * (c * NW_BANDS + b) + 1] = bin_off[c * NW_BANDS + b1];
            }
    }
}

// after NW: the pair whose key equals the read's final key owns the record
__global__ void select_kernel(const PairRec *pairs, const PairRes *res, uint32_t n,
                              const unsigned long long *best, unsigned long long *payload,
                              unsigned long long *pkey, uint64_t seg_seq_base, unsigned long long *pairs_total) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(pairs_total, (unsigned long long)n);
    unsigned needed = 0;  // diagnostic: candidates not later in scan order than their read's final hit
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const PairRec pr = pairs[i];
        const PairRes z = res[i];
        needed += pr.key <= best[pr.r] ? 1u : 0u;
        if ((z.stats & 0x80000000u) && pr.key == best[pr.r]) {
            payload[pr.r] = ((unsigned long long)(seg_seq_base + pr.s) << 32) | (z.stats & 0x7FFFFFFFu);
            pkey[pr.r] = pr.key;  // the key this payload belongs to (run_end drops superseded payloads)
        }
    }
    needed = __reduce_add_sync(0xffffffffu, needed);
    if ((threadIdx.x & 31) == 0 && needed) atomicAdd(pairs_total + 2, (unsigned long long)needed);  // counters[7]
}

// "Read size reached for gapped alignment." (src/alignmentFunctions.c:155): the reference stops when an
// e-value-passing hit with a read of more than MAX_READ_SIZE bases is reached BEFORE its query read has been
// accepted, i.e. when such a candidate is earlier in scan order than the read's final hit.
__global__ void readsize_kernel(const PairRec *pairs, uint32_t n, SeqMap db, SeqMap q,
                                const unsigned long long *best, unsigned long long *flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const PairRec pr = pairs[i];
        const uint32_t ylen = q.fixed_len ? q.fixed_len : q.start[pr.r + 1] - q.start[pr.r];
        const uint32_t xlen = db.fixed_len ? db.fixed_len : db.start[pr.s + 1] - db.start[pr.s];
        if ((xlen > (uint32_t)MAX_READ || ylen > (uint32_t)MAX_READ) && pr.key < best[pr.r]) *flag = 1ull;
    }
}

// multi-shard: keep the payload only where this shard owns the reduced key
__global__ void mask_payload_kernel(const unsigned long long *reduced, const unsigned long long *local,
                                    unsigned long long *payload, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (reduced[i] != local[i] || reduced[i] == KEY_NONE) payload[i] = 0ull;
}

__global__ void add_counters_kernel(unsigned long long *dst, const unsigned long long *src, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] += src[threadIdx.x];
}

__global__ void fill_u64_kernel(unsigned long long *p, unsigned long long v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

#endif  // __CUDACC__

}  // namespace imsame
