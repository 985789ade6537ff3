/*
 * synth.c -- deterministic synthetic metagenome generator.
 *
 * Genome pool: G genomes x Lg bases, iid uniform, genome g seeded by seed^g so a
 * larger pool extends a smaller one.  Database read: uniform genome, uniform
 * start, forward strand, Illumina-like errors.  Query read: with probability
 * 1/2 drawn like a database read plus divergence d, else iid random ("absent
 * species").  Every read has its own splitmix64 stream keyed by (seed, stream,
 * index), so any contiguous range of reads (a database shard) can be produced
 * independently and in parallel.
 *
 * Output is the reference's in-memory form (src/structs.h:40-45 SeqInfo): one
 * ASCII byte per base, reads concatenated without separators.
 */
#include "imsame_host.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t mix_key(uint64_t seed, uint64_t stream, uint64_t index) {
    uint64_t s = seed ^ (stream * 0xD6E8FEB86659FD93ull) ^ (index * 0xA24BAED4963EE407ull);
    splitmix64(&s);
    return splitmix64(&s);
}

struct imsame_synth_pool {
    uint64_t seed;
    uint32_t n_genomes;
    uint64_t genome_len;
    uint64_t words_per_genome; /* 32 bases per 64-bit word */
    uint64_t *bits;
};

/* threads of the generator loops (launchers such as torchrun pin OMP_NUM_THREADS=1 in the environment) */
void imsame_synth_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

imsame_synth_pool *imsame_synth_pool_create(uint64_t seed, uint32_t n_genomes, uint64_t genome_len) {
    imsame_synth_pool *p = (imsame_synth_pool *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->seed = seed;
    p->n_genomes = n_genomes;
    p->genome_len = genome_len;
    p->words_per_genome = (genome_len + 31) / 32;
    p->bits = (uint64_t *)malloc(p->words_per_genome * n_genomes * sizeof(uint64_t));
    if (!p->bits) { free(p); return NULL; }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t g = 0; g < (int64_t)n_genomes; g++) {
        uint64_t s = mix_key(seed ^ (uint64_t)g, 0x67656e6f6d65ull, 0);
        uint64_t *w = p->bits + (uint64_t)g * p->words_per_genome;
        for (uint64_t i = 0; i < p->words_per_genome; i++) w[i] = splitmix64(&s);
    }
    return p;
}

void imsame_synth_pool_destroy(imsame_synth_pool *p) {
    if (!p) return;
    free(p->bits);
    free(p);
}

static inline unsigned pool_base(const imsame_synth_pool *p, uint32_t g, uint64_t i) {
    return (unsigned)(p->bits[(uint64_t)g * p->words_per_genome + (i >> 5)] >> ((i & 31) * 2)) & 3u;
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

/* one read of length L sampled from the pool with the given per-base rates */
static void sample_read(const imsame_synth_pool *p, uint64_t *rng, uint32_t n_genomes_used, uint32_t L,
                        double sub, double ins, double del, unsigned char *out) {
    uint32_t g = (uint32_t)(splitmix64(rng) % n_genomes_used);
    uint64_t span = p->genome_len > 2ull * L ? p->genome_len - 2ull * L : 1;
    uint64_t at = splitmix64(rng) % span;
    const double inv = 1.0 / 18446744073709551616.0;
    uint32_t n = 0;
    while (n < L) {
        double u = (double)splitmix64(rng) * inv;
        if (u < del) { at++; continue; }
        if (u < del + ins) { out[n++] = (unsigned char)ACGT[splitmix64(rng) & 3]; continue; }
        unsigned b = at < p->genome_len ? pool_base(p, g, at) : (unsigned)(splitmix64(rng) & 3);
        at++;
        if (u < del + ins + sub) b = (b + 1 + (unsigned)(splitmix64(rng) % 3)) & 3u;
        out[n++] = (unsigned char)ACGT[b];
    }
}

void imsame_synth_db_reads(const imsame_synth_pool *p, uint64_t seed, uint64_t first, uint64_t count,
                           uint32_t L, unsigned char *out) {
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        uint64_t rng = mix_key(seed, 0x6462ull, first + (uint64_t)i);
        sample_read(p, &rng, p->n_genomes, L, 0.005, 0.00005, 0.00005, out + (uint64_t)i * L);
    }
}

void imsame_synth_query_reads(const imsame_synth_pool *p, uint64_t seed, uint64_t first, uint64_t count,
                              uint32_t L, double divergence, uint32_t n_genomes_used, unsigned char *out) {
    if (n_genomes_used == 0 || n_genomes_used > p->n_genomes) n_genomes_used = p->n_genomes;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        uint64_t rng = mix_key(seed, 0x7175657279ull, first + (uint64_t)i);
        unsigned char *o = out + (uint64_t)i * L;
        if (splitmix64(&rng) & 1) {
            sample_read(p, &rng, n_genomes_used, L, 0.005 + divergence, 0.00005 + divergence / 15.0,
                        0.00005 + divergence / 15.0, o);
        } else {
            for (uint32_t n = 0; n < L; n++) o[n] = (unsigned char)ACGT[splitmix64(&rng) & 3];
        }
    }
}

int imsame_synth_write_fasta(const char *path, const unsigned char *seq, uint64_t n_reads, uint32_t L,
                             char prefix) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    char *line = (char *)malloc((size_t)L + 64);
    for (uint64_t i = 0; i < n_reads; i++) {
        int h = sprintf(line, ">%c%llu\n", prefix, (unsigned long long)i);
        memcpy(line + h, seq + i * L, L);
        line[h + L] = '\n';
        fwrite(line, 1, (size_t)h + L + 1, f);
    }
    free(line);
    fclose(f);
    return 0;
}
