// CPU emulation of the winners-only traceback path: wavefront with back-pointer
// codes (nw_core.cuh, TB = true) -> tb_walk (traceback.cuh) -> host renderer
// (host/render.c), compared with the oracle's table-based traceback + text.
//   usage: tb_emul <n_cases> <seed> [long_len]   (long_len: every 7th / 5th case draws its X / Y length below it, default 700 / 900)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../imsame_b200/csrc/nw_core.cuh"
#include "../../imsame_b200/csrc/traceback.cuh"
extern "C" {
#include "../../imsame_b200/host/imsame_host.h"
#include "../../oracle/imsame_oracle.h"
}
using namespace imsame;
static uint64_t rng_state;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline uint32_t code(unsigned char c) { return (c >> 1) & 3; }

template <int S>
static NwBest run_pair_tb(const unsigned char *X, int xlen, const unsigned char *Y, int ylen, int igap, int egap,
                          std::vector<uint16_t> &tb, uint32_t &stride) {
    const int X1 = xlen - 1, Y1 = ylen - 1;
    stride = tb_stride(ylen, S);
    tb.assign((size_t)X1 * stride + 64, 0xFFFF);
    std::vector<NwLink> carry(xlen + 1), carry_next(xlen + 1);
    NwBest best; best.s = NW_NEG * 2; best.i = best.j = best.p = 0;
    for (int jb = 0; jb < Y1; jb += 32 * S) {
        int nl = (Y1 - jb + S - 1) / S; if (nl > 32) nl = 32;
        NwLane<S> lanes[32]; NwLink outs[32];
        for (int l = 0; l < nl; l++) {
            int j0 = jb + l * S + 1;
            uint64_t halo = 0;
            for (int k = 0; k < S + 2; k++) { int j = j0 - 2 + k; uint64_t c = (j >= 0 && j < ylen) ? code(Y[j]) : 0; halo |= c << (2 * k); }
            nw_lane_init<S>(lanes[l], code(X[0]), halo, jb == 0 && l == 0, l);
        }
        for (int t = 0; t <= X1 + nl - 2; t++)
            for (int l = nl - 1; l >= 0; l--) {
                int i = t - l + 1;
                if (i < 1 || i > X1) continue;
                int j0 = jb + l * S + 1;
                NwLink in = (l == 0) ? (jb == 0 ? nw_first_link(code(X[i]), code(Y[0])) : carry[i]) : outs[l - 1];
                uint32_t mm = 0;
                for (int c = 0; c < S; c++) { int j = j0 + c; uint32_t y = j < ylen ? code(Y[j]) : 0; if (y != code(X[i])) mm |= 1u << (2 * c); }
                NwLink out;
                { const int cl = (Y1 - 1 - jb) % S; const bool owns = Y1 >= j0 && Y1 < j0 + S;
                  uint16_t *tbr = tb.data() + (size_t)(i - 1) * stride + (j0 - 1);
                  if (t & 1) nw_row<S, true>(lanes[l], lanes[l].r1, lanes[l].r0, in, out, i, j0, mm, igap, egap, X1, Y1, cl, owns, jb == 0 && l == 0, tbr);
                  else nw_row<S, true>(lanes[l], lanes[l].r0, lanes[l].r1, in, out, i, j0, mm, igap, egap, X1, Y1, cl, owns, jb == 0 && l == 0, tbr); }
                outs[l] = out;
                if (l == 31) carry_next[i] = out;
            }
        for (int l = 0; l < nl; l++) if (best_better(lanes[l].best, best)) best = lanes[l].best;
        carry.swap(carry_next);
    }
    return best;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 200; rng_state = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    const int long_x = argc > 3 ? atoi(argv[3]) : 700, long_y = argc > 3 ? atoi(argv[3]) : 900;
    const char B[4] = {'A', 'C', 'G', 'T'};
    int bad = 0;
    std::vector<char> want(200000), got(200000);
    for (int it = 0; it < n; it++) {
        int xlen = 12 + rnd() % (it % 7 == 0 ? long_x : 300), ylen = 11 + rnd() % (it % 5 == 0 ? long_y : 300);
        if (it % 11 == 0) { xlen = 250; ylen = 250; }
        if (it % 13 == 0) xlen = 2 + rnd() % 6;
        if (it % 17 == 0) ylen = 2 + rnd() % 6;
        std::vector<unsigned char> X(xlen), Y(ylen);
        for (auto &c : X) c = B[rnd() & 3];
        int mode = rnd() % 4;
        if (mode == 0) for (auto &c : Y) c = B[rnd() & 3];
        else {
            int off = (int)(rnd() % (xlen)) - xlen / 3; double pe = (mode == 1) ? 0.03 : (mode == 2 ? 0.15 : 0.30);
            int src = off;
            for (int j = 0; j < ylen; j++) {
                double u = (rnd() >> 11) * (1.0 / 9007199254740992.0);
                if (u < pe / 6) src += 1 + rnd() % 4;
                if (u > 1 - pe / 6) { Y[j] = B[rnd() & 3]; continue; }
                unsigned char c = (src >= 0 && src < xlen) ? X[src] : B[rnd() & 3];
                if (u > 0.5 && u < 0.5 + pe) c = B[rnd() & 3];
                Y[j] = c; src++;
            }
        }
        int igap = -(int)(rnd() % 8), egap = -(int)(rnd() % 4);
        if (it % 3 == 0) { igap = -5; egap = -2; }
        int32_t os; uint32_t obx, oby, olen, oid;
        orc_nw_traceback(X.data(), xlen, Y.data(), ylen, igap, egap, &os, &obx, &oby, &olen, &oid, want.data(), want.size());
        std::vector<uint16_t> tb; uint32_t stride; NwBest b;
        switch (it % 4) {
            case 0: b = run_pair_tb<8>(X.data(), xlen, Y.data(), ylen, igap, egap, tb, stride); break;
            case 1: b = run_pair_tb<5>(X.data(), xlen, Y.data(), ylen, igap, egap, tb, stride); break;
            case 2: b = run_pair_tb<1>(X.data(), xlen, Y.data(), ylen, igap, egap, tb, stride); break;
            default: b = run_pair_tb<3>(X.data(), xlen, Y.data(), ylen, igap, egap, tb, stride); break;
        }
        std::vector<uint32_t> ops(xlen + ylen + 4);
        uint32_t ex, ey;
        uint32_t nops = tb_walk(tb.data(), stride, (uint32_t)b.i, (uint32_t)b.j, ops.data(), &ex, &ey);
        uint64_t tl = imsame_render_alignment(got.data(), X.data(), xlen, Y.data(), ylen, (uint32_t)b.i, (uint32_t)b.j, ops.data(), nops);
        bool ok = b.s == os && (uint32_t)b.i == obx && (uint32_t)b.j == oby && nw_stat_len(b.p) == olen &&
                  nw_stat_ids(b.p) == oid && tl == strlen(want.data()) && memcmp(got.data(), want.data(), tl) == 0;
        if (!ok) { bad++; if (bad < 5) printf("MISMATCH it=%d xlen=%d ylen=%d\n--- want\n%s--- got\n%s", it, xlen, ylen, want.data(), got.data()); }
    }
    printf("%d cases, %d mismatches\n", n, bad);
    return bad != 0;
}
