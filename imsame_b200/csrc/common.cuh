// common.cuh -- shared device/host helpers of the IMSAME hot path (sm_100a).
//
// Data layout in HBM:
//   * sequences: 2 bits per base (code = (ascii >> 1) & 3: A0 C1 T2 G3), 16 bases
//     per little-endian uint32 word, base i at bits 2*(i&15) of word i>>4; arrays
//     are padded with 16 zero bytes so 3-word window fetches never fault;
//   * read offsets: uint32 start[n+1] local to the array (a database segment
//     holds < 2^32 bases; global coordinates are added from the segment base);
//   * read lookup: blk[p >> 6] = read containing base 64*(p>>6), then a short
//     forward walk; or pure arithmetic when every read has the same length.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IMS_HD __host__ __device__ __forceinline__
#define IMS_D __device__ __forceinline__
#else
#define IMS_HD inline
#define IMS_D inline
#endif

namespace imsame {

constexpr int K = 12;                    // FIXED_K, src/structs.h:15
constexpr uint32_t KMASK = (1u << (2 * K)) - 1;
constexpr uint32_t NCODES = 1u << (2 * K);
constexpr int K_MIN = 4, K_MAX = 16;     // run-time seed lengths (imsame_gpu_set_kmer): a word is one 32-bit code
IMS_HD uint32_t kmask_of(int k) { return k >= 16 ? 0xFFFFFFFFu : (1u << (2 * k)) - 1u; }
IMS_HD uint64_t ncodes_of(int k) { return 1ull << (2 * k); }
constexpr int MAX_READ = 3000;           // MAX_READ_SIZE, src/structs.h:19
constexpr uint64_t KEY_NONE = 0x7FFFFFFFFFFFFFFFull;
constexpr int KEY_POS_BITS = 40;
constexpr uint64_t KEY_POS_MASK = (1ull << KEY_POS_BITS) - 1;

// winner order of the reference scan (src/alignmentFunctions.c:91-203): query
// k-mer end ascending, then database position DESCENDING (head-inserted lists,
// src/IMSAME.c:263-267).  Smaller key = found earlier by the reference.
IMS_HD uint64_t make_key(uint32_t e_rel, uint64_t db_pos_global) {
    return ((uint64_t)e_rel << KEY_POS_BITS) | (KEY_POS_MASK - db_pos_global);
}
IMS_HD uint32_t key_erel(uint64_t key) { return (uint32_t)(key >> KEY_POS_BITS); }
IMS_HD uint64_t key_dbpos(uint64_t key) { return KEY_POS_MASK - (key & KEY_POS_MASK); }

// 32 bases starting at base index i (i need not be aligned): base i+t at bits 2t
IMS_HD uint64_t fetch32(const uint32_t *pk, uint64_t i) {
    const uint64_t w = i >> 4;
    const unsigned sh = (unsigned)(i & 15) * 2;
    const uint64_t lo = (uint64_t)pk[w] | ((uint64_t)pk[w + 1] << 32);
    const uint64_t hi = pk[w + 2];
    return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}
// 16 bases starting at base index i
IMS_HD uint32_t fetch16(const uint32_t *pk, uint64_t i) {
    const uint64_t w = i >> 4;
    const unsigned sh = (unsigned)(i & 15) * 2;
    const uint64_t lo = (uint64_t)pk[w] | ((uint64_t)pk[w + 1] << 32);
    return (uint32_t)(lo >> sh);
}
IMS_HD uint32_t base_at(const uint32_t *pk, uint64_t i) { return (pk[i >> 4] >> ((i & 15) * 2)) & 3u; }

// mismatch mask of two 32-base windows: bit t set <=> bases differ
IMS_HD uint32_t mismatch32(uint64_t a, uint64_t b) {
    uint64_t d = a ^ b;
    d = (d | (d >> 1)) & 0x5555555555555555ull;
    // compress even bits to the low 32
    d = (d | (d >> 1)) & 0x3333333333333333ull;
    d = (d | (d >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    d = (d | (d >> 4)) & 0x00FF00FF00FF00FFull;
    d = (d | (d >> 8)) & 0x0000FFFF0000FFFFull;
    d = (d | (d >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)d;
}

// even bits of a 64-bit word -> low 32 bits
IMS_HD uint32_t even_bits64(uint64_t d) {
    d &= 0x5555555555555555ull;
    d = (d | (d >> 1)) & 0x3333333333333333ull;
    d = (d | (d >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    d = (d | (d >> 4)) & 0x00FF00FF00FF00FFull;
    d = (d | (d >> 8)) & 0x0000FFFF0000FFFFull;
    d = (d | (d >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)d;
}
IMS_HD uint32_t rev_bits32(uint32_t v) {
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

// Bit planes of a 32-base window: lo bit t = low code bit of base t, hi bit t = high code bit.  Two windows
// in this form differ at base t  <=>  bit t of (lo ^ lo') | (hi ^ hi')  -- two 3-input logic ops for 32 bases,
// against ~20 shift/mask operations to compare and compress two 2-bit-interleaved words.
//   forward window:  base t = seq[first + t]
//   backward window: base u = seq[last - u]   (walk order of src/alignmentFunctions.c:342-357)
// Bases before index 0 read as code 0 (callers mask them out through their step limits).
IMS_HD void planes_fwd(const uint32_t *pk, uint64_t first, uint32_t &lo, uint32_t &hi) {
    const uint64_t w = fetch32(pk, first);
    lo = even_bits64(w);
    hi = even_bits64(w >> 1);
}
IMS_HD void planes_bwd(const uint32_t *pk, int64_t last, uint32_t &lo, uint32_t &hi) {
    const int64_t first = last - 31;
    uint64_t w;
    if (first >= 0) w = fetch32(pk, (uint64_t)first);
    else w = first <= -32 ? 0ull : (fetch32(pk, 0) << (2 * (int)(-first)));
    lo = rev_bits32(even_bits64(w));
    hi = rev_bits32(even_bits64(w >> 1));
}

struct SeqMap {
    const uint32_t *pk;     // packed bases
    const uint32_t *start;  // n + 1 offsets
    const uint32_t *blk;    // read containing base 64*b (unused when fixed_len != 0)
    uint32_t n;             // reads
    uint32_t total;         // bases
    uint32_t fixed_len;     // != 0: every read has this length
};

IMS_HD uint32_t find_read(const SeqMap &m, uint32_t pos) {
    if (m.fixed_len) return pos / m.fixed_len;
    uint32_t r = m.blk[pos >> 6];
    while (m.start[r + 1] <= pos) r++;
    return r;
}
IMS_HD uint32_t read_start(const SeqMap &m, uint32_t r) { return m.fixed_len ? r * m.fixed_len : m.start[r]; }

// first read of a pthread chunk (src/IMSAME.c:414,433): only those start their
// k-mer stream on their own first base; every other read starts one base early
// (src/alignmentFunctions.c:93-105).
IMS_HD bool is_chunk_first(uint32_t r, uint32_t per, uint32_t n_threads) {
    if (per == 0) return r == 0;
    return (r % per == 0) && (r / per < n_threads);
}

}  // namespace imsame
