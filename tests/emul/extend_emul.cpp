// CPU check of the packed-sequence helpers and the ungapped extension used by the
// scan kernel (common.cuh / extend.cuh) against the oracle's byte-wise extension.
//   usage: extend_emul <seed> [k] [max_len]   (k = seed length, default 12 = the reference's FIXED_K; max_len = longest
//          read beyond the seed, default 200 -- thousands of bases exercise walks of many windows and the 15-bit step field)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#include "../../imsame_b200/csrc/extend.cuh"
extern "C" {
#include "../../oracle/imsame_oracle.h"
}
using namespace imsame;
static uint64_t rng_state;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static std::vector<uint32_t> pack(const std::vector<unsigned char> &s) {
    std::vector<uint32_t> pk(s.size() / 16 + 20, 0);
    for (size_t i = 0; i < s.size(); i++) pk[i >> 4] |= (uint32_t)((s[i] >> 1) & 3) << ((i & 15) * 2);
    return pk;
}

int main(int argc, char **argv) {
    rng_state = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    const int k = argc > 2 ? atoi(argv[2]) : K;
    const uint32_t kmask = kmask_of(k);
    const int max_len = argc > 3 ? atoi(argv[3]) : 200;
    const char B[4] = {'A', 'C', 'G', 'T'};
    // database: reads of varying length cut from a small genome (so that many words repeat)
    std::vector<unsigned char> genome(max_len > 600 ? (size_t)max_len * 5 : 3000);
    for (auto &c : genome) c = B[rnd() & 3];
    std::vector<unsigned char> D, Q;
    std::vector<uint64_t> ds, qs;
    for (int r = 0; r < 120; r++) {
        ds.push_back(D.size());
        int len = k + rnd() % max_len, at = rnd() % (genome.size() - len);
        for (int i = 0; i < len; i++) D.push_back((rnd() % 100 < 2) ? B[rnd() & 3] : genome[at + i]);
    }
    for (int r = 0; r < 60; r++) {
        qs.push_back(Q.size());
        int len = k - 1 + rnd() % max_len, at = rnd() % (genome.size() - len);
        for (int i = 0; i < len; i++) Q.push_back((rnd() % 100 < 5) ? B[rnd() & 3] : genome[at + i]);
    }
    ds.push_back(D.size()); qs.push_back(Q.size());
    orc_seqs db = {D.data(), ds.data(), D.size(), ds.size() - 1, nullptr, 0};
    orc_seqs q = {Q.data(), qs.data(), Q.size(), qs.size() - 1, nullptr, 0};
    auto dpk = pack(D), qpk = pack(Q);
    // fetch / base_at sanity
    for (size_t i = 0; i + 40 < D.size(); i += 7) {
        uint64_t w = fetch32(dpk.data(), i);
        for (int t = 0; t < 32; t++) if (((w >> (2 * t)) & 3) != ((D[i + t] >> 1) & 3u)) { printf("fetch32 mismatch\n"); return 1; }
        if ((fetch16(dpk.data(), i) & 3) != base_at(dpk.data(), i)) { printf("fetch16 mismatch\n"); return 1; }
    }
    // all word hits (including words that start one base before a query read: the phantom)
    std::unordered_multimap<uint32_t, uint32_t> qwords;  // code -> e
    for (size_t r = 0; r + 1 < qs.size(); r++) {
        int64_t lo = r == 0 ? (int64_t)qs[r] : (int64_t)qs[r] - 1;
        for (int64_t e = lo + k - 1; e < (int64_t)qs[r + 1]; e++) qwords.emplace(fetch16(qpk.data(), e - (k - 1)) & kmask, (uint32_t)e);
    }
    std::vector<uint32_t> lut2(EXT_LUT3_SIZE);
    build_ext_lut3(lut2.data());
    std::vector<ExtXY> lutxy(EXT_LUT3_SIZE);  // the scan kernel's shared-memory form of the same table
    for (int i = 0; i < EXT_LUT3_SIZE; i++) lutxy[(i & ~0xFF) | (int)ext_fold8((uint32_t)i & 0xFFu)] = ext_xy_of(lut2[i]);
    // window_mismatch (32-bit halves) against the 64-bit form
    for (size_t i = 0; i + 40 < D.size() && i + 40 < Q.size(); i += 3) {
        size_t j = (i * 7 + 5) % (Q.size() - 40);
        if (window_mismatch(dpk.data(), (uint32_t)i, qpk.data(), (uint32_t)j) != mismatch32(fetch32(dpk.data(), i), fetch32(qpk.data(), j))) { printf("window_mismatch differs\n"); return 1; }
    }
    long hits = 0, bad = 0;
    for (size_t s = 0; s + 1 < ds.size(); s++)
        for (uint64_t x = ds[s] + k - 1; x < ds[s + 1]; x++) {
            uint32_t code = fetch16(dpk.data(), x - (k - 1)) & kmask;
            auto range = qwords.equal_range(code);
            for (auto it = range.first; it != range.second; ++it) {
                uint32_t e = it->second, p = (uint32_t)x + 1;
                // read containing e
                size_t r = 0;
                while (qs[r + 1] <= e) r++;
                int64_t want = orc_extend_k(&db, &q, p, (uint64_t)e + 1, r, s, k);
                int got = extend_hit(dpk.data(), qpk.data(), p, e, (uint32_t)ds[s], (uint32_t)ds[s + 1], (uint32_t)qs[r], (uint32_t)qs[r + 1], k);
                ExtState st;
                ext_init(st, p, e, (uint32_t)ds[s], (uint32_t)ds[s + 1], (uint32_t)qs[r], (uint32_t)qs[r + 1], k);
                while (st.phase < 2) ext_window(st, lut2.data(), dpk.data(), qpk.data(), p, e, k);
                if (want != ext_result(st, k)) { if (bad++ < 10) printf("WINDOW MISMATCH s=%zu r=%zu p=%u e=%u want=%ld got=%d\n", s, r, p, e, (long)want, ext_result(st, k)); }
                {   // the scan kernel's form: first windows of two hits at once, the rest window by window
                    static ExtState prev; static uint32_t prev_p = 0, prev_e = 0; static long prev_want = 0; static bool have_prev = false;
                    ExtState cur;
                    ext_init(cur, p, e, (uint32_t)ds[s], (uint32_t)ds[s + 1], (uint32_t)qs[r], (uint32_t)qs[r + 1], k);
                    if (have_prev) {
                        uint32_t mfa, mba, mfb, mbb;
                        ExtState A = prev, Bst = cur;
                        ext_first_masks(A, dpk.data(), qpk.data(), prev_p, prev_e, mfa, mba, k);
                        ext_first_masks(Bst, dpk.data(), qpk.data(), p, e, mfb, mbb, k);
                        {   // the same masks and step limits from the two bit-plane halves (query table entry + database position)
                            ExtState C2;
                            uint32_t mf2, mb2;
                            const HitHalf dh = db_half(dpk.data(), p, (uint32_t)ds[s], (uint32_t)ds[s + 1], k);
                            const HitHalf qh = query_half(qpk.data(), e, (uint32_t)qs[r], (uint32_t)qs[r + 1], k);
                            hit_first_masks(dh, qh, C2, mf2, mb2);
                            const uint32_t ylen = (uint32_t)qh.froom + (uint32_t)(qh.broom + 1) + (uint32_t)(k - 1);
                            if (mf2 != mfb || mb2 != mbb || C2.fmax != cur.fmax || C2.bmax != cur.bmax || ylen != qs[r + 1] - qs[r] ||
                                qh.froom < 0 || qh.broom + 1 < 0) {
                                if (bad++ < 10) printf("PLANES MISMATCH p=%u e=%u mf %08x/%08x mb %08x/%08x\n", p, e, mf2, mfb, mb2, mbb);
                            }
                        }
                        ext_first2(A, Bst, lutxy.data(), mfa, mba, mfb, mbb, k);
                        while (A.phase < 2) ext_window(A, lutxy.data(), dpk.data(), qpk.data(), prev_p, prev_e, k);
                        while (Bst.phase < 2) ext_window(Bst, lutxy.data(), dpk.data(), qpk.data(), p, e, k);
                        if (prev_want != ext_result(A, k)) { if (bad++ < 10) printf("FIRST2(A) MISMATCH p=%u e=%u want=%ld got=%d\n", prev_p, prev_e, prev_want, ext_result(A, k)); }
                        if (want != ext_result(Bst, k)) { if (bad++ < 10) printf("FIRST2(B) MISMATCH p=%u e=%u want=%ld got=%d\n", p, e, (long)want, ext_result(Bst, k)); }
                    }
                    prev = cur; prev_p = p; prev_e = e; prev_want = (long)want; have_prev = true;
                }
                hits++;
                if (want != got) { if (bad++ < 10) printf("MISMATCH s=%zu r=%zu p=%u e=%u want=%ld got=%d\n", s, r, p, e, (long)want, got); }
            }
        }
    printf("%ld hits, %ld mismatches\n", hits, bad);
    return bad != 0 || hits < 1000;
}
