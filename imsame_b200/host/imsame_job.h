/*
 * imsame_job.h -- one IMSAME comparison (query sample vs database sample) on the GPU(s), shared by
 * the drop-in command line (imsame_main.c: one job per process, like the reference) and by the
 * in-process all-vs-all driver (imsame_allvsall_main.c: many jobs, one CUDA context).
 * Covers src/IMSAME.c:409-467 (fan-out, join, accepted count) and the record output of
 * src/alignmentFunctions.c:163-173.
 */
#ifndef IMSAME_JOB_H
#define IMSAME_JOB_H
#include "imsame_host.h"

typedef struct imsame_job_opts {
    uint64_t n_threads;
    long double minevalue, mincoverage, minidentity;
    int igap, egap; /* negated, as stored by the reference (src/IMSAME.c:565,568) */
    int gpus, device;
    int trace; /* phase wall times on stderr */
    int kmer;  /* seed length, 0 = the reference's FIXED_K (12) */
    /* optional (gpus == 1): the two read sets already resident on the device (imsame_gpu_sample_*); the host
       copies q / db are still what the records are rendered from */
    imsame_sample *q_sample;
    const imsame_sample *db_sample;
} imsame_job_opts;

/* Aligns every read of q against db and writes the records of the accepted reads to fout (may be
 * NULL: only count).  *ctx_cache (may be NULL) keeps the context of `device` alive between jobs when
 * gpus == 1.  Returns 0 or an IMSAME_E* code; err receives the CUDA detail. */
int imsame_run_job(const imsame_fasta *q, const imsame_fasta *db, const imsame_job_opts *o, FILE *fout,
                   imsame_ctx **ctx_cache, uint64_t *accepted, char *err, size_t errlen);

#endif
