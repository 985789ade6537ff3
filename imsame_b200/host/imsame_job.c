/* imsame_job.c -- see imsame_job.h */
#define _GNU_SOURCE
#include "imsame_job.h"

#include <inttypes.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* one database shard on one GPU */
typedef struct {
    int device;
    imsame_ctx *ctx; /* borrowed (cached) or NULL: create + destroy */
    const imsame_seqinfo *query;
    imsame_seqinfo db;
    imsame_params params;
    imsame_best *best;
    imsame_stats stats;
    int kmer;
    int rc;
    char err[256];
} shard_job;

static void *shard_main(void *arg) {
    shard_job *j = (shard_job *)arg;
    imsame_ctx *ctx = j->ctx;
    if (!ctx) {
        j->rc = imsame_gpu_create(&ctx, j->device);
        if (j->rc) return NULL;
    }
    j->rc = imsame_gpu_set_kmer(ctx, j->kmer);
    if (!j->rc) j->rc = imsame_gpu_align(ctx, &j->db, j->query, &j->params, j->best, &j->stats);
    if (j->rc) snprintf(j->err, sizeof j->err, "%s", imsame_gpu_last_cuda_error(ctx));
    if (!j->ctx) imsame_gpu_destroy(ctx);
    return NULL;
}

int imsame_run_job(const imsame_fasta *qf, const imsame_fasta *dbf, const imsame_job_opts *o, FILE *fout,
                   imsame_ctx **ctx_cache, uint64_t *accepted_out, char *err, size_t errlen) {
    const imsame_fasta q = *qf, db = *dbf;
    double tp = now_s();
    uint64_t accepted = 0;
    if (err && errlen) err[0] = 0;
    *accepted_out = 0;
    if (!(o->n_threads > 0 && q.n_seqs > 0 && db.n_seqs > 0)) return IMSAME_OK;
    imsame_best *best = (imsame_best *)calloc(q.n_seqs ? q.n_seqs : 1, sizeof(imsame_best));
    if (!best) return IMSAME_ENOMEM;
    imsame_seqinfo qv, dv;
    imsame_fasta_view(&q, &qv);
    imsame_fasta_view(&db, &dv);
    imsame_params p;
    memset(&p, 0, sizeof p);
    p.min_e_value = o->minevalue;
    p.min_coverage = o->mincoverage;
    p.min_identity = o->minidentity;
    p.igap = o->igap;
    p.egap = o->egap;
    p.n_threads = o->n_threads;

    int ng = o->gpus < 1 ? 1 : o->gpus;
    if ((uint64_t)ng > db.n_seqs) ng = (int)db.n_seqs;
    imsame_ctx *ctx = NULL; /* single-GPU jobs: one context for alignment and traceback, kept if the caller caches */
    int rc = IMSAME_OK;
    if (ng == 1) {
        if (ctx_cache && *ctx_cache) ctx = *ctx_cache;
        else if ((rc = imsame_gpu_create(&ctx, o->device))) { free(best); return rc; }
        if (ctx_cache) *ctx_cache = ctx;
    }
    shard_job *jobs = (shard_job *)calloc((size_t)ng, sizeof(shard_job));
    pthread_t *th = (pthread_t *)calloc((size_t)ng, sizeof(pthread_t));
    for (int g = 0; g < ng; g++) {
        /* contiguous read ranges; global coordinates keep keys and the e-value exact */
        uint64_t r0 = db.n_seqs * (uint64_t)g / (uint64_t)ng, r1 = db.n_seqs * (uint64_t)(g + 1) / (uint64_t)ng;
        uint64_t b0 = db.start_pos[r0], b1 = db.start_pos[r1];
        shard_job *j = &jobs[g];
        j->device = o->device + g;
        j->ctx = ng == 1 ? ctx : NULL;
        j->query = &qv;
        j->db.sequences = db.sequences + b0;
        j->db.total_len = b1 - b0;
        j->db.n_seqs = r1 - r0;
        uint64_t *st = (uint64_t *)malloc((r1 - r0 + 1) * sizeof(uint64_t));
        for (uint64_t r = r0; r <= r1; r++) st[r - r0] = db.start_pos[r] - b0;
        j->db.start_pos = st;
        uint64_t nb = 0, *bk = (uint64_t *)malloc((db.n_breaks + 1) * sizeof(uint64_t));
        for (uint64_t k = 0; k < db.n_breaks; k++)
            if (db.break_pos[k] >= b0 && db.break_pos[k] < b1) bk[nb++] = db.break_pos[k] - b0;
        j->db.break_pos = bk;
        j->db.n_breaks = nb;
        j->params = p;
        j->params.db_total_len_global = db.total_len;
        j->params.db_pos_base = b0;
        j->params.db_seq_base = r0;
        j->kmer = o->kmer ? o->kmer : 12; /* FIXED_K, src/structs.h:15 */
        j->best = ng == 1 ? best : (imsame_best *)calloc(q.n_seqs, sizeof(imsame_best));
        if (ng == 1) shard_main(j);
        else if (pthread_create(&th[g], NULL, shard_main, j)) { j->rc = IMSAME_ECUDA; snprintf(j->err, sizeof j->err, "pthread_create"); }
    }
    for (int g = 0; g < ng; g++) {
        if (ng > 1 && !(jobs[g].rc == IMSAME_ECUDA && !strcmp(jobs[g].err, "pthread_create"))) pthread_join(th[g], NULL);
        if (jobs[g].rc && !rc) {
            rc = jobs[g].rc;
            if (err && errlen) snprintf(err, errlen, "device %d: %s", jobs[g].device, jobs[g].err);
        }
    }
    if (!rc && ng > 1) {
        /* first accepted hit in the reference's scan order: k-mer end ascending, db position descending */
        for (uint64_t r = 0; r < q.n_seqs; r++)
            for (int g = 0; g < ng; g++) {
                const imsame_best *c = &jobs[g].best[r];
                if (!c->accepted) continue;
                if (!best[r].accepted || c->qpos_end < best[r].qpos_end ||
                    (c->qpos_end == best[r].qpos_end && c->db_pos > best[r].db_pos))
                    best[r] = *c;
            }
    }
    for (int g = 0; g < ng; g++) {
        if (ng > 1) free(jobs[g].best);
        free((void *)jobs[g].db.start_pos);
        free((void *)jobs[g].db.break_pos);
    }
    free(jobs);
    free(th);
    if (rc) { free(best); if (ctx && !ctx_cache) imsame_gpu_destroy(ctx); return rc; }
    for (uint64_t r = 0; r < q.n_seqs; r++) accepted += best[r].accepted;
    if (o->trace) { fprintf(stderr, "[imsame] align (all shards) %.3f s\n", now_s() - tp); tp = now_s(); }

    if (fout != NULL && accepted > 0) {
        if (!ctx && (rc = imsame_gpu_create(&ctx, o->device))) { free(best); return rc; }
        uint64_t *ops_off = (uint64_t *)malloc((q.n_seqs + 1) * sizeof(uint64_t));
        uint32_t *cell = (uint32_t *)malloc(4 * q.n_seqs * sizeof(uint32_t)), *ops = NULL;
        rc = imsame_gpu_traceback(ctx, &dv, &qv, &p, best, ops_off, &ops, cell);
        if (rc) {
            if (err && errlen) snprintf(err, errlen, "%s", imsame_gpu_last_cuda_error(ctx));
        } else {
            if (o->trace) { fprintf(stderr, "[imsame] traceback (GPU) %.3f s\n", now_s() - tp); tp = now_s(); }
            /* records are rendered by all host threads, each into its own buffer for a contiguous range of
               reads, and written in ascending read order (the reference's order with -n_threads 1) */
            int nt = 1;
#ifdef _OPENMP
            nt = omp_get_max_threads();
#endif
            if (nt > 64) nt = 64;
            if ((uint64_t)nt > q.n_seqs) nt = (int)q.n_seqs;
            char *bufs[64];
            size_t lens[64];
            int failed = 0;
            memset(bufs, 0, sizeof bufs);
            memset(lens, 0, sizeof lens);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
            for (int t = 0; t < nt; t++) {
                const uint64_t r0 = q.n_seqs * (uint64_t)t / (uint64_t)nt, r1 = q.n_seqs * (uint64_t)(t + 1) / (uint64_t)nt;
                const size_t rec_max = 6 * (2 * (size_t)IMSAME_MAX_READ_SIZE) + 768;
                size_t cap = 1 << 20, len = 0;
                char *buf = (char *)malloc(cap);
                for (uint64_t r = r0; r < r1 && buf; r++) {
                    if (!best[r].accepted) continue;
                    if (cap - len < rec_max) {
                        cap = cap * 2 + rec_max;
                        char *nb2 = (char *)realloc(buf, cap);
                        if (!nb2) { free(buf); buf = NULL; break; }
                        buf = nb2;
                    }
                    const uint64_t s = best[r].db_seq;
                    const uint32_t xlen = (uint32_t)(db.start_pos[s + 1] - db.start_pos[s]);
                    const uint32_t ylen = (uint32_t)(q.start_pos[r + 1] - q.start_pos[r]);
                    len += (size_t)imsame_format_header(buf + len, r, s, best[r].length, best[r].identities, ylen);
                    len += (size_t)imsame_render_alignment(buf + len, db.sequences + db.start_pos[s], xlen,
                                                           q.sequences + q.start_pos[r], ylen, cell[4 * r], cell[4 * r + 1],
                                                           ops + ops_off[r], ops_off[r + 1] - ops_off[r]);
                }
                if (!buf) {
#pragma omp atomic write
                    failed = 1;
                }
                bufs[t] = buf;
                lens[t] = len;
            }
            for (int t = 0; t < nt; t++) {
                if (!failed && bufs[t]) fwrite(bufs[t], 1, lens[t], fout);
                free(bufs[t]);
            }
            if (failed) rc = IMSAME_ENOMEM;
            if (o->trace) { fprintf(stderr, "[imsame] render + write %.3f s\n", now_s() - tp); tp = now_s(); }
        }
        imsame_gpu_free(ops);
        free(ops_off);
        free(cell);
    }
    if (ctx && !(ctx_cache && *ctx_cache == ctx)) imsame_gpu_destroy(ctx);
    free(best);
    *accepted_out = accepted;
    return rc;
}
