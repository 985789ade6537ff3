/*
 * IMSAME -- drop-in command line for the reference binary (src/IMSAME.c:34-578):
 * same flags and defaults (init_args, :520-578), same `.align` record format
 * (src/alignmentFunctions.c:167-168), same [INFO] summary lines on stdout
 * (src/IMSAME.c:63,102,106,295,317,407,416,470-473), same error text / exit
 * status (terror, src/commonFunctions.c:10-13; --help exits 1).
 *
 * The host side stays in C: FASTA ingest, thresholds, reporting.  Index build,
 * scan, extension, NW, filter and first-hit selection (src/IMSAME.c:232-281 and
 * :409-467) run on the GPU through include/imsame_gpu.h; the alignment text of
 * the accepted reads comes from a device traceback rendered by host/render.c.
 * Records are written in ascending read order, which is the reference's own
 * order with -n_threads 1 and a legal interleaving of its threads otherwise.
 *
 * Opt-in additions that default to reference behaviour:
 *   -gpus N     shard the database by contiguous read ranges over N GPUs (default 1)
 *   -device D   first CUDA device to use (default 0)
 */
#define _GNU_SOURCE
#include <inttypes.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "imsame_host.h"

static void terror(const char *s) { /* src/commonFunctions.c:10-13: message on STDOUT, exit(-1) */
    printf("ERR**** %s ****\n", s);
    exit(-1);
}

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

typedef struct {
    const char *query, *db, *out;
    uint64_t n_threads;
    long double minevalue, mincoverage, minidentity;
    int igap, egap;
    int gpus, device;
} cli_args;

static void usage_and_exit(void) { /* src/IMSAME.c:526-538 */
    fprintf(stdout, "USAGE:\n");
    fprintf(stdout, "           IMSAME -query [query] -db [database]\n");
    fprintf(stdout, "OPTIONAL:\n");
    fprintf(stdout, "           -n_threads  [Integer:   0<n_threads] (default 4)\n");
    fprintf(stdout, "           -evalue     [Double:    0<=pval<1] (default: 1 * 10^-20)\n");
    fprintf(stdout, "           -coverage   [Double:    0<coverage<=1 (default: 0.5)\n");
    fprintf(stdout, "           -identity   [Double:    0<identity<=1 (default: 0.5)\n");
    fprintf(stdout, "           -igap       [Integer:   (default: 5)\n");
    fprintf(stdout, "           -egap       [Integer:   (default: 2)\n");
    fprintf(stdout, "           -out        [File path]\n");
    fprintf(stdout, "           --verbose   Turns verbose on\n");
    fprintf(stdout, "           --help      Shows help for program usage\n");
    exit(1);
}

static void parse_args(int argc, char **av, cli_args *a) {
    /* defaults: src/IMSAME.c:44-49 */
    a->query = a->db = a->out = NULL;
    a->n_threads = 4;
    a->minevalue = 1 / powl(10, 20);
    a->mincoverage = 0.5;
    a->minidentity = 0.5;
    a->igap = -5;
    a->egap = -2;
    a->gpus = 1;
    a->device = 0;
    for (int i = 0; i < argc; i++) { /* the reference also scans av[0] and never skips values */
        const char *nxt = (i + 1 < argc) ? av[i + 1] : NULL;
        if (strcmp(av[i], "--help") == 0) usage_and_exit();
        if (strcmp(av[i], "-query") == 0 && nxt) a->query = nxt;
        if (strcmp(av[i], "-db") == 0 && nxt) a->db = nxt;
        if (strcmp(av[i], "-out") == 0 && nxt) a->out = nxt;
        if (strcmp(av[i], "-evalue") == 0 && nxt) {
            a->minevalue = (long double)atof(nxt); /* double first, then widened (:553) */
            if (a->minevalue < 0) terror("Min-e-value must be larger than zero");
        }
        if (strcmp(av[i], "-coverage") == 0 && nxt) {
            a->mincoverage = (long double)atof(nxt);
            if (a->mincoverage <= 0) terror("Min-coverage must be larger than zero");
        }
        if (strcmp(av[i], "-identity") == 0 && nxt) {
            a->minidentity = (long double)atof(nxt);
            if (a->minidentity <= 0) terror("Min-identity must be larger than zero");
        }
        if (strcmp(av[i], "-igap") == 0 && nxt) a->igap = -(atoi(nxt));
        if (strcmp(av[i], "-egap") == 0 && nxt) a->egap = -(atoi(nxt));
        if (strcmp(av[i], "-n_threads") == 0 && nxt) a->n_threads = (uint64_t)atoi(nxt);
        if (strcmp(av[i], "-gpus") == 0 && nxt) a->gpus = atoi(nxt);
        if (strcmp(av[i], "-device") == 0 && nxt) a->device = atoi(nxt);
    }
}

static void gpu_fail(imsame_ctx *ctx, int rc) {
    char msg[512];
    if (rc == IMSAME_EREADSIZE) terror("Read size reached for gapped alignment."); /* :155 */
    snprintf(msg, sizeof msg, "GPU hot path failed: %s%s%s", imsame_gpu_strerror(rc),
             ctx && imsame_gpu_last_cuda_error(ctx)[0] ? " / " : "", ctx ? imsame_gpu_last_cuda_error(ctx) : "");
    terror(msg);
}

/* one database shard on one GPU */
typedef struct {
    int device;
    const imsame_seqinfo *query;
    imsame_seqinfo db;
    imsame_params params;
    imsame_best *best;
    imsame_stats stats;
    int rc;
    char err[256];
} shard_job;

static void *shard_main(void *arg) {
    shard_job *j = (shard_job *)arg;
    imsame_ctx *ctx = NULL;
    j->rc = imsame_gpu_create(&ctx, j->device);
    if (j->rc) return NULL;
    j->rc = imsame_gpu_align(ctx, &j->db, j->query, &j->params, j->best, &j->stats);
    if (j->rc) snprintf(j->err, sizeof j->err, "%s", imsame_gpu_last_cuda_error(ctx));
    imsame_gpu_destroy(ctx);
    return NULL;
}

int main(int argc, char **av) {
    cli_args a;
    parse_args(argc, av, &a);
    FILE *fq = a.query ? fopen(a.query, "rt") : NULL;
    FILE *fd = a.db ? fopen(a.db, "rt") : NULL;
    FILE *fout = a.out ? fopen(a.out, "wt") : NULL; /* a failed -out open silently disables output (:549-550) */
    if (fq == NULL || fd == NULL) terror("A query and database is required");
    fclose(fq);
    fclose(fd);

    double t0 = now_s();
    fprintf(stdout, "[INFO] Init. quick table\n");
    if (a.gpus < 1) a.gpus = 1;
    fprintf(stdout, "[INFO] Initialization took %e seconds \n", now_s() - t0);

    fprintf(stdout, "[INFO] Loading database\n");
    t0 = now_s();
    imsame_fasta db, q;
    if (imsame_fasta_load(a.db, 1, &db)) terror("Could not allocate memory for database vector");
    fprintf(stdout, "[INFO] Database loaded and of length %" PRIu64 ". Hash table building took %e seconds\n",
            db.total_len, now_s() - t0);
    t0 = now_s();
    fprintf(stdout, "[INFO] Loading query.\n");
    if (imsame_fasta_load(a.query, 0, &q)) terror("Could not allocate memory for query vector");
    fprintf(stdout, "[INFO] Query loaded and of length %" PRIu64 ". Took %e seconds\n", q.total_len, now_s() - t0);

    t0 = now_s();
    fprintf(stdout, "[INFO] Computing alignments.\n");
    /* src/IMSAME.c:414,430-452 + src/alignmentFunctions.c:88: every thread announces its range */
    if (a.n_threads > 0) {
        uint64_t per = (uint64_t)floorl((long double)q.n_seqs / (long double)a.n_threads);
        for (uint64_t t = 0; t < a.n_threads; t++)
            fprintf(stdout, "Going from %" PRIu64 " to %" PRIu64 "\n", t * per,
                    t == a.n_threads - 1 ? q.n_seqs : (t + 1) * per);
    }
    fflush(stdout);

    const int trace = getenv("IMSAME_TRACE") != NULL; /* phase wall times on stderr (not part of the reference's output) */
    double tp = now_s();
    uint64_t accepted = 0;
    imsame_best *best = (imsame_best *)calloc(q.n_seqs ? q.n_seqs : 1, sizeof(imsame_best));
    if (!best) terror("Could not allocate arguments for hash table");
    imsame_seqinfo qv, dv;
    imsame_fasta_view(&q, &qv);
    imsame_fasta_view(&db, &dv);
    imsame_params p;
    memset(&p, 0, sizeof p);
    p.min_e_value = a.minevalue;
    p.min_coverage = a.mincoverage;
    p.min_identity = a.minidentity;
    p.igap = a.igap;
    p.egap = a.egap;
    p.n_threads = a.n_threads;

    if (a.n_threads > 0 && q.n_seqs > 0 && db.n_seqs > 0) {
        int ng = a.gpus;
        if ((uint64_t)ng > db.n_seqs) ng = (int)db.n_seqs;
        shard_job *jobs = (shard_job *)calloc((size_t)ng, sizeof(shard_job));
        pthread_t *th = (pthread_t *)calloc((size_t)ng, sizeof(pthread_t));
        for (int g = 0; g < ng; g++) {
            /* contiguous read ranges; global coordinates keep keys and the e-value exact */
            uint64_t r0 = db.n_seqs * (uint64_t)g / (uint64_t)ng, r1 = db.n_seqs * (uint64_t)(g + 1) / (uint64_t)ng;
            uint64_t b0 = db.start_pos[r0], b1 = db.start_pos[r1];
            shard_job *j = &jobs[g];
            j->device = a.device + g;
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
            j->best = ng == 1 ? best : (imsame_best *)calloc(q.n_seqs, sizeof(imsame_best));
            if (ng == 1) shard_main(j);
            else if (pthread_create(&th[g], NULL, shard_main, j)) terror("Could not launch");
        }
        for (int g = 0; g < ng; g++) {
            if (ng > 1) pthread_join(th[g], NULL);
            if (jobs[g].rc) {
                if (jobs[g].rc == IMSAME_EREADSIZE) terror("Read size reached for gapped alignment.");
                char msg[512];
                snprintf(msg, sizeof msg, "GPU hot path failed on device %d: %s %s", jobs[g].device,
                         imsame_gpu_strerror(jobs[g].rc), jobs[g].err);
                terror(msg);
            }
        }
        if (ng > 1) {
            /* first accepted hit in the reference's scan order: k-mer end ascending, db position descending */
            for (uint64_t r = 0; r < q.n_seqs; r++)
                for (int g = 0; g < ng; g++) {
                    const imsame_best *c = &jobs[g].best[r];
                    if (!c->accepted) continue;
                    if (!best[r].accepted || c->qpos_end < best[r].qpos_end ||
                        (c->qpos_end == best[r].qpos_end && c->db_pos > best[r].db_pos))
                        best[r] = *c;
                }
            for (int g = 0; g < ng; g++) free(jobs[g].best);
        }
        for (int g = 0; g < ng; g++) { free((void *)jobs[g].db.start_pos); free((void *)jobs[g].db.break_pos); }
        free(jobs);
        free(th);
        for (uint64_t r = 0; r < q.n_seqs; r++) accepted += best[r].accepted;
        if (trace) { fprintf(stderr, "[imsame] align (all shards) %.3f s\n", now_s() - tp); tp = now_s(); }

        if (fout != NULL && accepted > 0) {
            imsame_ctx *ctx = NULL;
            int rc = imsame_gpu_create(&ctx, a.device);
            if (rc) gpu_fail(NULL, rc);
            uint64_t *ops_off = (uint64_t *)malloc((q.n_seqs + 1) * sizeof(uint64_t));
            uint32_t *cell = (uint32_t *)malloc(4 * q.n_seqs * sizeof(uint32_t)), *ops = NULL;
            rc = imsame_gpu_traceback(ctx, &dv, &qv, &p, best, ops_off, &ops, cell);
            if (rc) gpu_fail(ctx, rc);
            if (trace) { fprintf(stderr, "[imsame] traceback (GPU) %.3f s\n", now_s() - tp); tp = now_s(); }
            char *text = (char *)malloc(6 * (2 * (size_t)IMSAME_MAX_READ_SIZE) + 512), hdr[256];
            for (uint64_t r = 0; r < q.n_seqs; r++) {
                if (!best[r].accepted) continue;
                const uint64_t s = best[r].db_seq;
                const uint32_t xlen = (uint32_t)(db.start_pos[s + 1] - db.start_pos[s]);
                const uint32_t ylen = (uint32_t)(q.start_pos[r + 1] - q.start_pos[r]);
                int hl = imsame_format_header(hdr, r, s, best[r].length, best[r].identities, ylen);
                fwrite(hdr, 1, (size_t)hl, fout);
                uint64_t tl = imsame_render_alignment(text, db.sequences + db.start_pos[s], xlen,
                                                      q.sequences + q.start_pos[r], ylen, cell[4 * r], cell[4 * r + 1],
                                                      ops + ops_off[r], ops_off[r + 1] - ops_off[r]);
                fwrite(text, 1, (size_t)tl, fout);
            }
            if (trace) { fprintf(stderr, "[imsame] render + write %.3f s\n", now_s() - tp); tp = now_s(); }
            free(text);
            imsame_gpu_free(ops);
            free(ops_off);
            free(cell);
            imsame_gpu_destroy(ctx);
        }
    }

    fprintf(stdout, "[INFO] Alignments computed in %e seconds.\n", now_s() - t0);
    fprintf(stdout,
            "[INFO] %" PRIu64 " reads (%" PRIu64 ") from the query were found in the database (%" PRIu64
            ") at a minimum e-value of %Le and minimum coverage of %d%%.\n",
            accepted, q.n_seqs, db.n_seqs, (long double)a.minevalue, (int)(100 * a.mincoverage));
    fprintf(stdout, "[INFO] The Jaccard-index is: %Le\n",
            (long double)accepted / ((db.n_seqs + q.n_seqs) - accepted));
    fprintf(stdout, "[INFO] Deallocating heap memory.\n");
    if (fout != NULL) fclose(fout);
    free(best);
    imsame_fasta_free(&db);
    imsame_fasta_free(&q);
    return 0;
}
