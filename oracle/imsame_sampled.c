/*
 * oracle/imsame_sampled.c -- TEST INFRASTRUCTURE ONLY (see imsame_oracle.h).
 *
 * Per-read checker for inputs far too large for the reference's database index
 * (24 B per database base, src/structs.h:26-30): the first accepted hit of a
 * SAMPLE of query reads against the whole database (or one shard of it), with
 * no database index at all.  The k-mers of the sampled reads are hashed
 * (the mirror image of src/IMSAME.c:232-281), the database is streamed once
 * on all host threads, every seed hit runs orc_extend_k + orc_evalue exactly
 * as the reference's list walk does (src/alignmentFunctions.c:126-139), and
 * then every sampled read replays ITS e-value-passing hits in the reference's
 * scan order -- k-mer end ascending (src/alignmentFunctions.c:91-203), database
 * position descending (head-inserted lists, src/IMSAME.c:263-267) -- through
 * orc_nw_forward and the filter (:163) until the first one is accepted
 * (:172,189).  That is the reference's sequential early exit restricted to the
 * sampled reads; reads are independent of each other in the reference (the only
 * coupling, the cross-read "phantom" word, depends on the neighbouring read's
 * bases, not on its result).
 *
 * Shards: `db` may be a contiguous range of database reads; db_pos_base /
 * db_seq_base make positions and read indices global, p->db_total_len_global is
 * the e-value's database length (src/alignmentFunctions.c:384).  The first
 * accepted hit over the whole database is the shard result with the smallest
 * (qpos_end, -db_pos).
 */
#define _GNU_SOURCE
#include "imsame_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* host threads of the checker (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline unsigned sbase2(unsigned char c) { /* src/IMSAME.c:55-59 */
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;
}

typedef struct { uint32_t si; uint64_t e; } sk_entry;              /* sample index, k-mer end (query coords) */
typedef struct { uint32_t si; uint32_t s; uint64_t e, pos; } sk_cand; /* e-value-passing hit (pos, s local) */

typedef struct { sk_cand *v; uint64_t n, cap; } cand_vec;
static int cand_push(cand_vec *c, sk_cand x) {
    if (c->n == c->cap) {
        uint64_t nc = c->cap ? c->cap * 2 : 4096;
        sk_cand *nv = (sk_cand *)realloc(c->v, nc * sizeof(sk_cand));
        if (!nv) return -1;
        c->v = nv; c->cap = nc;
    }
    c->v[c->n++] = x;
    return 0;
}

/* scan order: sample, then k-mer end ascending, then database position descending */
static int cand_cmp(const void *a, const void *b) {
    const sk_cand *x = (const sk_cand *)a, *y = (const sk_cand *)b;
    if (x->si != y->si) return x->si < y->si ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    if (x->pos != y->pos) return x->pos > y->pos ? -1 : 1;
    return 0;
}

static int sk_accept(uint32_t length, uint32_t identities, uint64_t ylen, const orc_params *p) { /* :163 */
    return ((long double)length / ylen) >= p->min_coverage && ((long double)identities / length) >= p->min_identity;
}

int orc_align_sampled(const orc_seqs *db, const orc_seqs *q, const orc_params *p, const uint64_t *reads,
                      uint64_t n_reads, uint64_t db_pos_base, uint64_t db_seq_base, orc_best *best,
                      orc_stats *st) {
    const int k = p->k;
    if (k < 4 || k > 13) return -9;
    if (n_reads >= 0xFFFFFFFFull) return -9;
    const uint64_t n_codes = 1ull << (2 * k), mask = n_codes - 1;
    const uint64_t T = p->n_threads ? p->n_threads : 1, per = q->n_seqs / T;
    const uint64_t db_total = p->db_total_len_global ? p->db_total_len_global : db->total_len;
    for (uint64_t i = 0; i < n_reads; i++) memset(&best[i], 0, sizeof(orc_best));

    /* ---- table of the sampled reads' words (incl. the phantom word, src/alignmentFunctions.c:93-105) ---- */
    uint32_t *off = (uint32_t *)calloc(n_codes + 2, sizeof(uint32_t));
    uint64_t *bits = (uint64_t *)calloc(n_codes / 64 + 1, sizeof(uint64_t));
    if (!off || !bits) return -2;
    sk_entry *ent = NULL;
    for (int pass = 0; pass < 2; pass++) {
        for (uint64_t i = 0; i < n_reads; i++) {
            const uint64_t r = reads[i];
            if (r >= q->n_seqs) return -3;
            int first = 0;
            if (per == 0) first = (r == 0);
            else if (r % per == 0 && r / per < T) first = 1;
            const int64_t lo = first ? (int64_t)q->start[r] : (int64_t)q->start[r] - 1;
            const int64_t hi = r < q->n_seqs - 1 ? (int64_t)q->start[r + 1] - 2 : (int64_t)q->total_len - 1;
            for (int64_t e = lo + k - 1; e <= hi; e++) {
                uint64_t code = 0;
                for (int64_t x = e - k + 1; x <= e; x++) code = ((code << 2) | sbase2(q->seq[x])) & mask;
                if (pass == 0) { off[code + 1]++; bits[code >> 6] |= 1ull << (code & 63); }
                else { sk_entry en; en.si = (uint32_t)i; en.e = (uint64_t)e; ent[off[code]++] = en; }
            }
        }
        if (pass == 0) {
            for (uint64_t c = 0; c < n_codes; c++) off[c + 1] += off[c];
            ent = (sk_entry *)malloc(((uint64_t)off[n_codes] + 1) * sizeof(sk_entry));
            if (!ent) return -2;
        } else {
            for (uint64_t c = n_codes; c > 0; c--) off[c] = off[c - 1];
            off[0] = 0;
        }
    }

    /* ---- stream the database: words inside one read, never across a break (src/IMSAME.c:229-231,283) ---- */
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    cand_vec *cv = (cand_vec *)calloc((size_t)nt, sizeof(cand_vec));
    uint64_t hits = 0, pass_cnt = 0;
    int fail = 0;
#pragma omp parallel num_threads(nt) reduction(+ : hits, pass_cnt)
    {
        int me = 0;
#ifdef _OPENMP
        me = omp_get_thread_num();
#endif
        cand_vec *mine = &cv[me];
#pragma omp for schedule(dynamic, 4096)
        for (uint64_t s = 0; s < db->n_seqs; s++) {
            uint64_t word = 0, code = 0;
            uint64_t bi = 0;
            if (db->n_brk) { /* first break > start[s] (a break at the read start is a no-op) */
                uint64_t lo = 0, hi = db->n_brk;
                while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (db->brk[mid] <= db->start[s]) lo = mid + 1; else hi = mid; }
                bi = lo;
            }
            for (uint64_t x = db->start[s]; x < db->start[s + 1]; x++) {
                if (bi < db->n_brk && db->brk[bi] == x) { word = 0; bi++; }
                code = ((code << 2) | sbase2(db->seq[x])) & mask;
                if (word < (uint64_t)k) word++;
                if (word < (uint64_t)k) continue;
                if (!((bits[code >> 6] >> (code & 63)) & 1)) continue;
                const uint64_t pos = x + 1; /* index after the word's last base (src/IMSAME.c:247,265) */
                for (uint32_t h = off[code]; h < off[code + 1]; h++) {
                    const uint64_t r = reads[ent[h].si], e = ent[h].e;
                    hits++;
                    const int64_t n = orc_extend_k(db, q, pos, e + 1, r, s, k);
                    const uint64_t ylen = q->start[r + 1] - q->start[r];
                    if (!(orc_evalue(n, ylen, db_total) < p->min_e_value)) continue;
                    pass_cnt++;
                    sk_cand c; c.si = ent[h].si; c.s = (uint32_t)s; c.e = e; c.pos = pos;
                    if (cand_push(mine, c)) {
#pragma omp atomic write
                        fail = 1;
                    }
                }
            }
        }
    }
    free(off); free(bits); free(ent);
    if (fail) { for (int t = 0; t < nt; t++) free(cv[t].v); free(cv); return -2; }

    /* ---- per sampled read: replay in scan order until the first accepted alignment ---- */
    uint64_t total = 0;
    for (int t = 0; t < nt; t++) total += cv[t].n;
    sk_cand *all = (sk_cand *)malloc((total + 1) * sizeof(sk_cand));
    uint64_t *first_of = (uint64_t *)calloc(n_reads + 1, sizeof(uint64_t));
    if (!all || !first_of) return -2;
    {
        uint64_t at = 0;
        for (int t = 0; t < nt; t++) { if (cv[t].n) memcpy(all + at, cv[t].v, cv[t].n * sizeof(sk_cand)); at += cv[t].n; free(cv[t].v); }
        free(cv);
    }
    qsort(all, total, sizeof(sk_cand), cand_cmp);
    for (uint64_t i = 0; i < total; i++) first_of[all[i].si + 1]++;
    for (uint64_t i = 0; i < n_reads; i++) first_of[i + 1] += first_of[i];
    uint64_t nw_calls = 0, accepted = 0;
    int too_long = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nw_calls, accepted)
    for (uint64_t i = 0; i < n_reads; i++) {
        const uint64_t r = reads[i], ylen = q->start[r + 1] - q->start[r];
        /* NW depends on the pair only: remember the database reads that already failed the filter */
        uint32_t *failed = NULL;
        uint64_t nf = 0, capf = 0;
        for (uint64_t h = first_of[i]; h < first_of[i + 1]; h++) {
            const sk_cand *c = &all[h];
            uint64_t f;
            for (f = 0; f < nf; f++) if (failed[f] == c->s) break;
            if (f < nf) continue;
            const uint64_t xlen = db->start[c->s + 1] - db->start[c->s];
            if (xlen > ORC_MAX_READ || ylen > ORC_MAX_READ) { /* src/alignmentFunctions.c:155 */
#pragma omp atomic write
                too_long = 1;
                break;
            }
            int32_t sc; uint32_t bx, by, len, id;
            nw_calls++;
            int ok = 0;
            if (orc_nw_forward(db->seq + db->start[c->s], xlen, q->seq + q->start[r], ylen, p->igap, p->egap, &sc, &bx,
                               &by, &len, &id) == 0)
                ok = len > 0 && sk_accept(len, id, ylen, p);
            if (ok) {
                orc_best *b = &best[i];
                b->db_seq = db_seq_base + c->s; b->qpos_end = c->e; b->db_pos = db_pos_base + c->pos;
                b->length = len; b->identities = id; b->score = sc; b->bx = bx; b->by = by; b->accepted = 1;
                accepted++;
                break;
            }
            if (nf == capf) { capf = capf ? capf * 2 : 64; failed = (uint32_t *)realloc(failed, capf * sizeof(uint32_t)); }
            failed[nf++] = c->s;
        }
        free(failed);
    }
    free(all); free(first_of);
    if (st) { st->hits = hits; st->evalue_pass = pass_cnt; st->nw_calls = nw_calls; st->accepted = accepted; }
    return too_long ? -5 : 0;
}
