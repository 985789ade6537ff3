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

/* contexts of a multi-GPU job come up side by side (0.5 - 1 s each) */
typedef struct {
    int device, kmer, rc;
    imsame_ctx *ctx;
} ctx_job;

static void *ctx_main(void *arg) {
    ctx_job *j = (ctx_job *)arg;
    j->rc = imsame_gpu_create(&j->ctx, j->device);
    if (!j->rc) j->rc = imsame_gpu_set_kmer(j->ctx, j->kmer);
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
    const int kmer = o->kmer ? o->kmer : 12; /* FIXED_K, src/structs.h:15 */

    int ng = o->gpus < 1 ? 1 : o->gpus;
    if ((uint64_t)ng > db.n_seqs) ng = (int)db.n_seqs;
    imsame_ctx *ctx = NULL; /* device `o->device`: alignment (shard 0) and traceback; kept if the caller caches */
    int rc = IMSAME_OK;
    if (ctx_cache && *ctx_cache) ctx = *ctx_cache;
    if (ng == 1) {
        if (!ctx && (rc = imsame_gpu_create(&ctx, o->device))) { free(best); return rc; }
        if (ctx_cache) *ctx_cache = ctx;
        rc = imsame_gpu_set_kmer(ctx, kmer);
        if (!rc) {
            if (o->q_sample && o->db_sample) rc = imsame_gpu_align_samples(ctx, o->db_sample, o->q_sample, &p, best, NULL);
            else rc = imsame_gpu_align(ctx, &dv, &qv, &p, best, NULL);
        }
        if (rc && err && errlen) snprintf(err, errlen, "%s", imsame_gpu_last_cuda_error(ctx));
    } else {
        /* the database sharded by contiguous read ranges over ng GPUs; the per-read first accepted hit is
           reduced with NCCL inside the library (imsame_gpu_align_sharded), replacing src/IMSAME.c:430-467 */
        ctx_job *jobs = (ctx_job *)calloc((size_t)ng, sizeof(ctx_job));
        pthread_t *th = (pthread_t *)calloc((size_t)ng, sizeof(pthread_t));
        imsame_ctx **ctxs = (imsame_ctx **)calloc((size_t)ng, sizeof(imsame_ctx *));
        if (!jobs || !th || !ctxs) {
            free(jobs); free(th); free(ctxs); free(best);
            return IMSAME_ENOMEM;
        }
        for (int g = 0; g < ng; g++) {
            jobs[g].device = o->device + g;
            jobs[g].kmer = kmer;
            if (g == 0 && ctx) { jobs[g].ctx = ctx; jobs[g].rc = imsame_gpu_set_kmer(ctx, kmer); continue; }
            if (pthread_create(&th[g], NULL, ctx_main, &jobs[g])) { ctx_main(&jobs[g]); th[g] = 0; }
        }
        for (int g = 0; g < ng; g++) {
            if (th[g]) pthread_join(th[g], NULL);
            ctxs[g] = jobs[g].ctx;
            if (jobs[g].rc && !rc) {
                rc = jobs[g].rc;
                if (err && errlen) snprintf(err, errlen, "device %d: %s", jobs[g].device, imsame_gpu_strerror(rc));
            }
        }
        if (o->trace) { fprintf(stderr, "[imsame] %d contexts %.3f s\n", ng, now_s() - tp); tp = now_s(); }
        if (!rc) {
            rc = imsame_gpu_align_sharded(ctxs, ng, &dv, &qv, &p, best, NULL);
            if (rc && err && errlen) snprintf(err, errlen, "%s", imsame_gpu_last_cuda_error(ctxs[0]));
        }
        for (int g = 1; g < ng; g++) imsame_gpu_destroy(ctxs[g]);
        ctx = ctxs[0];
        if (ctx_cache && ctx) *ctx_cache = ctx;
        free(jobs);
        free(th);
        free(ctxs);
    }
    if (rc) { free(best); if (ctx && !(ctx_cache && *ctx_cache == ctx)) imsame_gpu_destroy(ctx); return rc; }
    for (uint64_t r = 0; r < q.n_seqs; r++) accepted += best[r].accepted;
    if (o->trace) { fprintf(stderr, "[imsame] align (all shards) %.3f s\n", now_s() - tp); tp = now_s(); }

    if (fout != NULL && accepted > 0) {
        if (!ctx && (rc = imsame_gpu_create(&ctx, o->device))) { free(best); return rc; }
        uint64_t *ops_off = (uint64_t *)malloc((q.n_seqs + 1) * sizeof(uint64_t));
        uint32_t *cell = (uint32_t *)malloc(4 * q.n_seqs * sizeof(uint32_t)), *ops = NULL;
        rc = (ops_off && cell) ? imsame_gpu_traceback(ctx, &dv, &qv, &p, best, ops_off, &ops, cell) : IMSAME_ENOMEM;
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
