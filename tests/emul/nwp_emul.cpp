// CPU emulation of the 16-lane packed-word wavefront (nwp_core.cuh, kernel nwp.cuh): lanes
// stepped in lockstep, neighbour hand-over = value of the previous step.  Checks
// imsame::pw_row / pw_lane_init against the oracle's forward NW, bit for bit, on every pair
// that pw_eligible admits (and checks that the eligibility bound is what keeps it exact by
// also running pairs at the edge of the admitted range).  Wide reads (query reads of 257..321 bases,
// 18 / 20 columns per lane): the statistics word V = 256 * length + identities is split by
// pw_split_stats and, where both splits are possible, the pair is run again with length unit 0 --
// the kernel's epilogue, step for step.
//   usage: nwp_emul <n_cases> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../imsame_b200/csrc/nwp_core.cuh"
extern "C" {
#include "../../oracle/imsame_oracle.h"
}
using namespace imsame;

static uint64_t rng_state;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline uint32_t code(unsigned char c) { return (c >> 1) & 3; }

struct Res { int s, i, j; uint32_t len, id; };

struct HostEW {
    const PwK *k; uint64_t mm;
    PwE4 operator()(int g) const { return pw_e4(*k, (mm >> (8 * g)) & 0xFFu); }
};

static int n_again = 0, n_wide = 0, n_biased = 0;

struct Raw { bool have; int bw, bi, bj; };

template <int S>
static Raw run_raw(const unsigned char *X, int xlen, const unsigned char *Y, int ylen, const PwK &k) {
    const int X1 = xlen - 1, Y1 = ylen - 1;
    Raw raw; raw.have = false; raw.bw = raw.bi = raw.bj = 0;
    if (X1 < 1 || Y1 < 1) return raw;
    const int nl = (Y1 + S - 1) / S;
    if (nl > PW_LANES) { fprintf(stderr, "bad S\n"); exit(2); }
    PwLane<S> lanes[PW_LANES]; PwLink outs[PW_LANES];
    for (int l = 0; l < nl; l++) {
        int j0 = l * S + 1;
        uint64_t halo = 0;
        for (int c = 0; c < S + 2; c++) { int j = j0 - 2 + c; uint64_t cd = (j >= 0 && j < ylen) ? code(Y[j]) : (rnd() & 3); halo |= cd << (2 * c); }
        pw_lane_init<S>(lanes[l], k, code(X[0]), halo, l == 0, l);
    }
    for (int t = 0; t <= X1 + nl - 2; t++) {
        for (int l = nl - 1; l >= 0; l--) {
            int i = t - l + 1;
            if (i < 1 || i > X1) continue;
            int j0 = l * S + 1;
            PwLink in = (l == 0) ? pw_first_link(k, code(X[i]), code(Y[0])) : outs[l - 1];
            uint64_t mm = 0;
            for (int c = 0; c < S; c++) { int j = j0 + c; uint32_t y = j < ylen ? code(Y[j]) : (uint32_t)(rnd() & 3); if (y != code(X[i])) mm |= 1ull << (2 * c); }
            HostEW ew{&k, mm};
            PwLink out;
            const int cl = (Y1 - 1) % S; const bool owns = Y1 >= j0 && Y1 < j0 + S;
            if (t & 1) pw_row<S>(lanes[l], lanes[l].r1, lanes[l].r0, in, out, i, j0, ew, k, X1, Y1, cl, owns);
            else pw_row<S>(lanes[l], lanes[l].r0, lanes[l].r1, in, out, i, j0, ew, k, X1, Y1, cl, owns);
            outs[l] = out;
        }
    }
    for (int l = 0; l < nl; l++) {
        if ((X1 - 1 + l) & 1) pw_last_row<S>(lanes[l], lanes[l].r0, l * S + 1, X1, Y1);
        else pw_last_row<S>(lanes[l], lanes[l].r1, l * S + 1, X1, Y1);
    }
    bool have = false; int bz = 0, bw = 0, bi = 0, bj = 0;
    for (int l = 0; l < nl; l++) {
        const PwLane<S> &L = lanes[l];
        if (L.bw == (int)0x80000000) continue;
        bool better = !have || L.bz > bz || (L.bz == bz && (L.bi > bi || (L.bi == bi && L.bj > bj)));
        if (better) { have = true; bz = L.bz; bw = L.bw; bi = L.bi; bj = L.bj; }
    }
    raw.have = have; raw.bw = bw; raw.bi = bi; raw.bj = bj;
    return raw;
}

// the epilogue of nwp_kernel<S, CL>
template <int S>
static Res run_pair(const unsigned char *X, int xlen, const unsigned char *Y, int ylen, int igap, int egap, int bias) {
    const PwK k = pw_consts(igap, egap, 1, PW_LEN1, bias);
    Res best; best.s = NW_NEG * 2; best.i = best.j = 0; best.len = best.id = 0;
    if (xlen < 2 || ylen < 2) return best;
    const Raw r = run_raw<S>(X, xlen, Y, ylen, k);
    if (!r.have) return best;
    best.s = pw_score(k, r.bw); best.i = r.bi; best.j = r.bj;
    if (S <= 16) { best.len = pw_len(k, r.bw); best.id = pw_ids(k, r.bw); return best; }
    n_wide++;
    const uint32_t v = pw_stats(k, r.bw);
    if (!pw_split_stats(v, r.bi, r.bj, &best.len, &best.id)) return best;
    n_again++;
    const PwK k_id = pw_consts(igap, egap, 1, 0, bias);
    const Raw r2 = run_raw<S>(X, xlen, Y, ylen, k_id);
    if (r2.bi != r.bi || r2.bj != r.bj || pw_score(k_id, r2.bw) != best.s) { fprintf(stderr, "second run took another path\n"); exit(3); }
    best.id = pw_stats(k_id, r2.bw);
    best.len = (v - best.id) >> 8;
    return best;
}

template <int S>
static Res dispatch(int s, const unsigned char *X, int xlen, const unsigned char *Y, int ylen, int igap, int egap, int bias) {
    if (s == S) return run_pair<S>(X, xlen, Y, ylen, igap, egap, bias);
    if constexpr (S < PW_MAX_S) return dispatch<S + 1>(s, X, xlen, Y, ylen, igap, egap, bias);
    fprintf(stderr, "bad S\n"); exit(2);
}

// A run whose longest reads are eligible launches no other NW kernel (capi.cu: use_packed), and a mixed run gives the
// packed-word kernel every pair that is eligible under the RUN's score offset: check that eligibility with a run's
// offset really holds for every pair inside the run's box of lengths, and that the class of a query read gives its
// packed-word kernel enough columns (nw.cuh is not included here: 16 lanes x 2 * class columns, classes as below).
static int check_eligibility_boxes() {
    int bad = 0;
    const int gaps[][2] = {{-5, -2}, {0, 0}, {-8, -1}, {-7, -3}, {-1, 0}, {-12, -4}, {-40, -3}, {-3, -1}};
    for (auto &g : gaps) {
        for (int trial = 0; trial < 40; trial++) {
            uint32_t xmax = 2 + rnd() % 560, ymax = 2 + rnd() % 340;
            if (trial == 0) { xmax = 512; ymax = 321; }
            if (trial == 1) { xmax = 250; ymax = 250; }
            if (trial == 2) { xmax = 3000; ymax = 3000; }
            if (trial == 3) { xmax = 300; ymax = 300; }
            const int bias = pw_bias(xmax, ymax, g[0], g[1]);
            if (xmax == 250 && ymax == 250 && g[0] == -5 && bias != 0) { printf("cfg2 must not need a score offset\n"); bad++; }
            if (pw_eligible(xmax, ymax, g[0], g[1], bias)) {  // "packed" run: every pair of the box must be eligible
                for (uint32_t x = 2; x <= xmax; x++)
                    for (uint32_t y = 2; y <= ymax; y++)
                        if (!pw_pair_eligible(x, y, g[0], g[1], bias)) { if (bad++ < 5) printf("box %u x %u gaps %d,%d: pair %u x %u not eligible\n", xmax, ymax, g[0], g[1], x, y); }
            }
            // whatever the run: an eligible pair keeps every field inside its bits (the bounds pw_eligible states)
            for (int k = 0; k < 2000; k++) {
                const uint32_t x = 2 + rnd() % (xmax < 600 ? xmax : 600), y = 2 + rnd() % (ymax < 340 ? ymax : 340);
                if (!pw_pair_eligible(x, y, g[0], g[1], bias)) continue;
                const int X1 = (int)x - 1, Y1 = (int)y - 1, mn = X1 < Y1 ? X1 : Y1;
                const long depth = pw_depth(x, y, g[0], g[1]);
                const bool ok = y <= (uint32_t)PW_MAX_Y && x <= (uint32_t)PW_MAX_X && depth - bias <= 1890 && 4L * (mn + 1) + bias <= 2040 &&
                                X1 + Y1 + 3 <= 1023 && (Y1 > PW_NARROW_Y1 || mn <= 255) && mn <= 511;
                if (!ok && bad++ < 5) printf("pair %u x %u gaps %d,%d bias %d: field bounds violated\n", x, y, g[0], g[1], bias);
            }
        }
    }
    return bad;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 200; rng_state = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    if (check_eligibility_boxes()) { printf("eligibility self-check failed\n"); return 1; }
    const char B[4] = {'A', 'C', 'G', 'T'};
    int bad = 0, skipped = 0;
    for (int it = 0; it < n; it++) {
        int xlen = 2 + rnd() % (it % 7 == 0 ? 510 : 255), ylen = 2 + rnd() % 256;
        const bool wide = (it % 3 == 1);
        if (wide) { ylen = 257 + rnd() % 65; xlen = (it % 4 == 0) ? 2 + rnd() % 511 : 230 + rnd() % 100; }
        if (it % 11 == 0) { xlen = 250; ylen = 250; }
        if (it % 19 == 0) { xlen = 256; ylen = 257; }
        if (it % 23 == 0) { xlen = 512; ylen = 256; }
        if (it % 41 == 0) { xlen = 300; ylen = 300; }
        if (it % 43 == 0) { xlen = 309; ylen = 309; }
        if (it % 47 == 0) { xlen = 260 + rnd() % 50; ylen = 321; }
        if (it % 53 == 0) { xlen = 321; ylen = 321; }
        if (it % 59 == 0) { xlen = 512; ylen = 321; }
        if (it % 13 == 0) { xlen = 2 + rnd() % 6; }
        if (it % 17 == 0) { ylen = 2 + rnd() % 6; }
        std::vector<unsigned char> X(xlen), Y(ylen);
        int alpha = (it % 29 == 0) ? 1 : 3;  // low-complexity reads: many equal scores (tie rules)
        for (auto &c : X) c = B[rnd() & alpha];
        int mode = rnd() % 4;
        if (mode == 0) for (auto &c : Y) c = B[rnd() & alpha];
        else {
            int off = (int)(rnd() % (xlen)) - xlen / 3; double pe = (mode == 1) ? 0.03 : (mode == 2 ? 0.15 : 0.30);
            if (it % 5 == 0) off = 0;
            if (wide && it % 2 == 0) { off = (int)(rnd() % 40) - 20; pe = (rnd() % 3 == 0) ? 0.0 : 0.02; }  // near-complete overlaps: 256 and more identities
            int src = off;
            for (int j = 0; j < ylen; j++) {
                double u = (rnd() >> 11) * (1.0 / 9007199254740992.0);
                if (u < pe / 6) { src += 1 + rnd() % 4; }
                if (u > 1 - pe / 6) { Y[j] = B[rnd() & 3]; continue; }
                unsigned char c = (src >= 0 && src < xlen) ? X[src] : B[rnd() & alpha];
                if (u > 0.5 && u < 0.5 + pe) c = B[rnd() & 3];
                Y[j] = c; src++;
            }
        }
        int igap = -(int)(rnd() % 8), egap = -(int)(rnd() % 4);
        if (it % 3 == 0 || (wide && it % 2 == 0)) { igap = -5; egap = -2; }
        if (it % 31 == 0) { igap = 0; egap = 0; }
        if (it % 37 == 0) { igap = -(int)(rnd() % 40); egap = -(int)(rnd() % 7); }
        // the score offset is a property of the RUN (its longest reads): the pair's own, that of a run with longer
        // reads, or that of a run with reads beyond what packed words take (clamped to 512 x 321)
        int bias = pw_bias(xlen, ylen, igap, egap);
        if (it % 4 == 1) bias = pw_bias(xlen + rnd() % 200, ylen + rnd() % 60, igap, egap);
        if (it % 4 == 2) bias = pw_bias(3000, 3000, igap, egap);
        if (bias) n_biased++;
        if (!pw_eligible(xlen, ylen, igap, egap, bias)) { skipped++; continue; }
        int32_t os; uint32_t obx, oby, olen, oid;
        orc_nw_forward(X.data(), xlen, Y.data(), ylen, igap, egap, &os, &obx, &oby, &olen, &oid);
        const int smin = (ylen - 1 + PW_LANES - 1) / PW_LANES;
        int s = smin < 1 ? 1 : smin;
        if (ylen - 1 > PW_NARROW_Y1) s = ylen - 1 <= 288 ? (it % 5 == 2 && smin <= 17 ? 17 + (int)(rnd() % 4) : 18) : (it % 5 == 2 && smin <= 19 ? 19 : 20);  // classes 9 and 10 (sometimes another width)
        else if (it % 2 == 0) { s = 2 * ((ylen - 1 + 31) / 32); if (s < 2) s = 2; }  // what the kernel uses: 2 * class
        else if (s < 16 && (it % 4 == 1)) s += rnd() % (17 - s);
        Res b = dispatch<1>(s, X.data(), xlen, Y.data(), ylen, igap, egap, bias);
        bool ok;
        if (xlen < 2 || ylen < 2) ok = true;
        else ok = b.s == os && (uint32_t)b.i == obx && (uint32_t)b.j == oby && b.len == olen && b.id == oid;
        if (!ok) {
            bad++;
            if (bad < 10) printf("MISMATCH it=%d xlen=%d ylen=%d S=%d gaps=%d,%d: emul (%d,%d,%d,%u,%u) oracle (%d,%u,%u,%u,%u)\n", it, xlen, ylen, s, igap, egap,
                                 b.s, b.i, b.j, b.len, b.id, os, obx, oby, olen, oid);
        }
    }
    printf("%d cases, %d skipped (not eligible), %d mismatches; %d wide pairs, %d of them run twice; %d cases with a score offset\n", n, skipped, bad, n_wide, n_again, n_biased);
    return bad != 0;
}
