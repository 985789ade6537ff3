// tools/prefilter_bound.cpp -- CPU experiment for DESIGN.md 9.2: how many random 12-mer hits can be rejected EXACTLY
// (never a passing hit: "wrong" must print 0) before the table walks of the scan kernel, from the two 32-base
// mismatch masks alone (bound 0) or after the forward walk (bound 1)?  g++ -O2 -std=c++17 tools/prefilter_bound.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <unordered_map>
#include "../imsame_b200/csrc/extend.cuh"
using namespace imsame;
static uint64_t st = 12345;
static uint64_t rnd() { uint64_t z = (st += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
int main() {
    const int L = 250, ND = 40000, NQ = 4000, NMIN = 66;
    std::vector<uint32_t> dpk(ND * L / 16 + 20, 0), qpk(NQ * L / 16 + 20, 0);
    for (auto &w : dpk) w = (uint32_t)rnd();
    for (auto &w : qpk) w = (uint32_t)rnd();
    std::unordered_multimap<uint32_t, uint32_t> qw;
    for (int r = 0; r < NQ; r++) for (int e = r * L + 11; e < (r + 1) * L - 1; e++) qw.emplace(fetch16(qpk.data(), e - 11) & KMASK, e);
    std::vector<uint32_t> lut(EXT_LUT3_SIZE); build_ext_lut3(lut.data());
    long hits = 0, pass = 0, rej0 = 0, rej1 = 0, wrong = 0, fwd_in_window = 0, both_in = 0;
    for (int s = 0; s < ND; s++) for (int x = s * L + 11; x < (s + 1) * L; x++) {
        uint32_t code = fetch16(dpk.data(), x - 11) & KMASK;
        auto rg = qw.equal_range(code);
        for (auto it = rg.first; it != rg.second; ++it) {
            uint32_t e = it->second, p = x + 1; int r = e / L;
            ExtState a; ext_init(a, p, e, s * L, (s + 1) * L, r * L, (r + 1) * L);
            uint32_t mf, mb; ext_first_masks(a, dpk.data(), qpk.data(), p, e, mf, mb);
            int n = extend_hit(dpk.data(), qpk.data(), p, e, s * L, (s + 1) * L, r * L, (r + 1) * L);
            hits++; if (n >= NMIN) pass++;
            // bound 0: popcounts only
            int Pf = __builtin_popcount(mf), Pb = __builtin_popcount(mb), Zf = 32 - Pf, Zb = 32 - Pb;
            bool f_in = K + 32 - 2 * Pf <= 0;           // forward ends inside its window
            bool b_in0 = K + Zf + 32 - 2 * Pb <= 0;     // backward (start <= K + Zf) ends inside its window
            if (f_in && b_in0 && 2 * (K + Zf + Zb) - (K - 1) < NMIN) { rej0++; if (n >= NMIN) wrong++; }
            // bound 1: after the exact forward walk (4 table steps): hr, matches, fe known
            int run = K << EXT_SC_SHIFT, best = (K << EXT_SC_SHIFT) + EXT_BEST_BIAS;
            for (int c = 0; c < 32; c += 8) ext_step(lut.data(), (mf >> c) & 0xFF, run, best);
            int sc = run >> EXT_SC_SHIFT; bool fover = sc == 0 || a.fmax <= 32;
            if (fover) {
                fwd_in_window++;
                int hr = ((best & ~(int)EXT_POS_MASK) - EXT_BEST_BIAS) >> EXT_SC_SHIFT;
                int pos_f = best & (int)EXT_POS_MASK;               // fe + 1
                int idn2 = (run & (int)EXT_POS_MASK) + sc - K;      // 2 * matches forward
                bool b_in = hr + 32 - 2 * Pb <= 0;
                if (b_in) both_in++;
                // n = 2K + idn2 + 2 Mb - (pos_f + K + pos_b - 1), Mb <= Zb, pos_b >= 0... (pos_b = 0: no backward max: then -(..-1))
                int nmax = 2 * K + idn2 + 2 * Zb - (pos_f + K - 1);
                if (b_in && nmax < NMIN) { rej1++; if (n >= NMIN) wrong++; }
            }
        }
    }
    printf("hits %ld pass %ld (%.3f%%)  reject0 %.1f%%  fwd-in-window %.1f%%  both-in %.1f%%  reject1 %.1f%%  wrong %ld\n", hits, pass, 100.0 * pass / hits,
           100.0 * rej0 / hits, 100.0 * fwd_in_window / hits, 100.0 * both_in / hits, 100.0 * rej1 / hits, wrong);
}
