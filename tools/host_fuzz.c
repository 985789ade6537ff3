/*
 * host_fuzz.c -- seeded fuzz of the host-side C around the GPU hot path, meant to be built with
 * -fsanitize=address,undefined (tests/test_host_cpu.py::test_host_fuzz_under_sanitizers does that):
 *
 *   1. imsame_fasta_parse_mem on random byte soups (record starts anywhere, CRLF, lower case, IUPAC letters,
 *      headers without a newline, text before the first header, empty input) with the piece size forced down
 *      to a few bytes (IMSAME_TEST_FASTA_PIECE), so that every input is cut into many pieces parsed by
 *      different threads, against a char-at-a-time model of the reference loader written here
 *      (src/IMSAME.c:193-285 database, :323-347 query: a record starts at '>', its header runs to '\n', only
 *      A/C/G/T are stored upper-cased, a dropped character other than '\n' restarts the seed word);
 *   2. imsame_revcomp_mem (threaded, cut into the same tiny pieces) on the same soups against a serial
 *      record-by-record model (the form tests/test_host_cpu.py::test_revcomp_fuzz_against_the_reference_tool pinned
 *      to the reference tool), and the parse of its output;
 *   3. imsame_render_alignment on random paths with the buffer the header promises (6 (xlen + ylen) + 256).
 *
 * usage: host_fuzz [iterations] [seed]          prints one summary line, exit status 1 on a mismatch
 */
#include "../imsame_b200/host/imsame_host.h"
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state;
static uint64_t rnd(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}
static uint64_t below(uint64_t n) { return n ? rnd() % n : 0; }

/* the loader, one character at a time */
typedef struct {
    unsigned char *seq;
    uint64_t *start, *brk;
    uint64_t pos, n, nb;
} model;

static void model_parse(const unsigned char *b, size_t n, int is_db, model *m) {
    m->seq = (unsigned char *)malloc(n + 1);
    m->start = (uint64_t *)malloc((n + 2) * sizeof(uint64_t));
    m->brk = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    m->pos = m->n = m->nb = 0;
    size_t i = 0;
    while (i < n) {
        if (b[i] != '>') { i++; continue; }
        const uint64_t rec = m->pos;
        m->start[m->n++] = m->pos;
        while (i < n && b[i] != '\n') i++;
        int word_reset = 0;
        for (i++; i < n && b[i] != '>'; i++) {
            unsigned char c = b[i];
            if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 32);
            if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
                if (word_reset && is_db && m->pos > rec) m->brk[m->nb++] = m->pos;
                word_reset = 0;
                m->seq[m->pos++] = c;
            } else if (c != '\n') {
                word_reset = 1;
            }
        }
    }
    m->start[m->n] = m->pos;
}

static void model_free(model *m) { free(m->seq); free(m->start); free(m->brk); }

/* revComp, one record after the other (src/reverseComplement.c:47-112; this serial form is what
 * tests/test_host_cpu.py::test_revcomp_fuzz_against_the_reference_tool pinned to the reference tool) */
static size_t model_revcomp(const unsigned char *b, size_t n, unsigned char **out) {
    size_t nrec = 0, cap = 0;
    for (size_t i = 0; i < n; i++) nrec += b[i] == '>';
    size_t *off = (size_t *)malloc((nrec + 1) * sizeof(size_t));
    nrec = 0;
    for (size_t i = 0; i < n; i++)
        if (b[i] == '>') off[nrec++] = i;
    for (size_t r = 0; r < nrec; r++) cap += n - off[r] + 2;
    unsigned char *dst = (unsigned char *)malloc(cap + 1);
    size_t w = 0;
    for (size_t r = nrec; r-- > 0;) {
        size_t i = off[r], h = i;
        while (h < n && b[h] != '\n') h++;
        if (h < n) h++;
        size_t end = h;
        while (end < n && b[end] != '>') end++;
        memcpy(dst + w, b + i, h - i);
        w += h - i;
        for (size_t k = end; k-- > h;) {
            unsigned char c = b[k];
            if (!((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) continue;
            switch (c) {
                case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break; case 'U': c = 'A'; break;
                case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break; case 'u': c = 'a'; break;
                default: break;
            }
            dst[w++] = c;
        }
        dst[w++] = '\n';
    }
    free(off);
    *out = dst;
    return w;
}

static size_t soup(unsigned char *dst, size_t cap) {
    static const char *tok[] = {"A", "C", "G", "T", "a", "c", "g", "t", "N", "n", "\n", "\r\n", ">", ">h\n", " ", "-", "U", "u",
                                "R", "\t", "ACGTACGTACGTACGTACGTTTGGCCAA", "\n>x y\n", "\n\n", ">\n", "\n>", "ACGTTGCAAC\n",
                                "GGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTT\n", ">>", "\n>r\nAC"};
    const size_t nt = sizeof tok / sizeof tok[0];
    size_t n = 0, want = below(3) == 0 ? below(40) : below(cap - 80);
    if (below(2)) { memcpy(dst, ">first rec\n", 11); n = 11; }
    while (n < want) {
        const char *t = tok[below(nt)];
        const size_t l = strlen(t);
        memcpy(dst + n, t, l);
        n += l;
    }
    return n;
}

static int same_parse(const imsame_fasta *f, const model *m, int is_db) {
    if (f->total_len != m->pos || f->n_seqs != m->n) return 0;
    if (m->pos && memcmp(f->sequences, m->seq, m->pos)) return 0;
    if (memcmp(f->start_pos, m->start, (m->n + 1) * sizeof(uint64_t))) return 0;
    if (is_db) {
        if (f->n_breaks != m->nb) return 0;
        if (m->nb && memcmp(f->break_pos, m->brk, m->nb * sizeof(uint64_t))) return 0;
    }
    return 1;
}

static int check_parse(const unsigned char *b, size_t n, uint64_t it, const char *what) {
    for (int is_db = 0; is_db < 2; is_db++) {
        model m;
        model_parse(b, n, is_db, &m);
        imsame_fasta f;
        if (imsame_fasta_parse_mem(b, n, is_db, &f)) { fprintf(stderr, "parse failed\n"); return 1; }
        const int ok = same_parse(&f, &m, is_db);
        if (!ok)
            fprintf(stderr, "host_fuzz: %s mismatch at iteration %llu (is_db %d, %zu bytes): reads %llu / %llu, bases %llu / %llu, breaks %llu / %llu\n",
                    what, (unsigned long long)it, is_db, n, (unsigned long long)f.n_seqs, (unsigned long long)m.n,
                    (unsigned long long)f.total_len, (unsigned long long)m.pos, (unsigned long long)f.n_breaks, (unsigned long long)m.nb);
        imsame_fasta_free(&f);
        model_free(&m);
        if (!ok) return 1;
    }
    return 0;
}

/* a random path from (bx, by), the best cell on the last row or column, back to the first row or column, as the
 * run-length ops of csrc/traceback.cuh: 1 = c diagonal steps, 2 = c columns (x -= c, y -= 1), 3 = c rows */
static int check_render(void) {
    const uint32_t xlen = 1 + (uint32_t)below(400), ylen = 1 + (uint32_t)below(400);
    unsigned char *X = (unsigned char *)malloc(xlen), *Y = (unsigned char *)malloc(ylen);
    for (uint32_t i = 0; i < xlen; i++) X[i] = (unsigned char)"ACGT"[below(4)];
    for (uint32_t i = 0; i < ylen; i++) Y[i] = (unsigned char)"ACGT"[below(4)];
    uint32_t bx = xlen - 1, by = ylen - 1;
    if (below(2)) bx = (uint32_t)below(xlen); else by = (uint32_t)below(ylen);
    uint32_t *ops = (uint32_t *)malloc((size_t)(xlen + ylen + 2) * sizeof(uint32_t));
    uint64_t n_ops = 0, ncol = 0;
    uint32_t x = bx, y = by;
    while (x > 0 && y > 0) {
        const uint64_t k = below(10);
        uint32_t type, c;
        if (k < 7) { type = 1; c = 1 + (uint32_t)below(x < y ? x : y); if (below(3)) c = 1; x -= c; y -= c; }
        else if (k < 9) { type = 2; c = 1 + (uint32_t)below(x); x -= c; y -= 1; }
        else { type = 3; c = 1 + (uint32_t)below(y); y -= c; x -= 1; }
        if (n_ops && type == 1 && (ops[n_ops - 1] >> 28) == 1) ops[n_ops - 1] += c;
        else ops[n_ops++] = (type << 28) | c;
        ncol += c;
    }
    const size_t cap = 6 * ((size_t)xlen + ylen) + 256; /* imsame_host.h */
    char *dst = (char *)malloc(cap);
    const uint64_t w = imsame_render_alignment(dst, X, xlen, Y, ylen, bx, by, ops, n_ops);
    int bad = w >= cap || dst[w] != 0 || strlen(dst) != w;
    /* every block is three lines; the text ends with one empty line */
    uint64_t lines = 0;
    for (uint64_t i = 0; i < w; i++) lines += dst[i] == '\n';
    bad |= (lines % 3) != 1;
    for (uint64_t i = 0; i < w; i++) bad |= strchr("ACGT-* \n", dst[i]) == NULL;
    /* columns: left overhang + path + what is left of the longer tail, three lines per started block of 60 */
    {
        const uint64_t lead = x > y ? x : y, Lx = lead + ncol + (xlen - 1 - bx), Ly = lead + ncol + (ylen - 1 - by);
        const uint64_t shorter = Lx < Ly ? Lx : Ly;
        bad |= lines != 3 * ((shorter + 59) / 60) + 1;
    }
    if (bad) fprintf(stderr, "host_fuzz: render check failed (xlen %u ylen %u bx %u by %u, %llu ops, %llu bytes of %zu)\n", xlen, ylen, bx, by,
                     (unsigned long long)n_ops, (unsigned long long)w, cap);
    free(dst); free(ops); free(X); free(Y);
    return bad;
}

int main(int argc, char **av) {
    const uint64_t iters = argc > 1 ? strtoull(av[1], NULL, 10) : 2000;
    rng_state = 0x9E3779B97F4A7C15ull ^ (argc > 2 ? strtoull(av[2], NULL, 10) : 1);
    const size_t cap = 6000;
    unsigned char *b = (unsigned char *)malloc(cap + 128);
    uint64_t parsed = 0, rendered = 0;
    for (uint64_t it = 0; it < iters; it++) {
        /* pieces of 1..64 bytes: an input of a few kB is cut at (nearly) every "\n>" */
        char piece[32];
        snprintf(piece, sizeof piece, "%llu", (unsigned long long)(1 + below(64)));
        setenv("IMSAME_TEST_FASTA_PIECE", piece, 1);
        const size_t n = soup(b, cap);
        /* an exact-size copy: reads past the end of the image are the sanitizer's to find */
        unsigned char *img = (unsigned char *)malloc(n ? n : 1);
        memcpy(img, b, n);
        if (check_parse(img, n, it, "parse")) return 1;
        unsigned char *rc = NULL;
        size_t rl = 0;
        if (imsame_revcomp_mem(img, n, &rc, &rl)) { fprintf(stderr, "revcomp failed\n"); return 1; }
        {
            unsigned char *want = NULL;
            const size_t wl = model_revcomp(img, n, &want);
            if (wl != rl || (rl && memcmp(want, rc, rl))) {
                fprintf(stderr, "host_fuzz: revcomp mismatch at iteration %llu (%zu bytes in, %zu / %zu bytes out)\n", (unsigned long long)it, n, rl, wl);
                return 1;
            }
            free(want);
        }
        unsigned char *rimg = (unsigned char *)malloc(rl ? rl : 1);
        memcpy(rimg, rc, rl);
        free(rc);
        if (check_parse(rimg, rl, it, "parse of the reverse complement")) return 1;
        free(rimg);
        free(img);
        parsed += 4;
        if (check_render()) return 1;
        rendered++;
    }
    free(b);
    printf("host_fuzz: %llu parses (many pieces each) and %llu reverse complements equal to the serial models, %llu rendered paths, 0 mismatches\n",
           (unsigned long long)parsed, (unsigned long long)rendered, (unsigned long long)rendered);
    return 0;
}
