/*
 * fasta.c -- FASTA ingest producing the reference's in-memory form.
 *
 * Behaviour follows src/IMSAME.c:196-289 (database) and :320-371 (query):
 *   - a record starts at '>' ; the header is skipped up to the first '\n';
 *   - text before the first '>' is ignored;
 *   - characters are upper-cased; only A/C/G/T are stored, everything else is
 *     dropped (multi-line records therefore concatenate);
 *   - start_pos[r] = number of stored bases before record r;
 *   - for the database, a dropped character other than '\n' resets the seed
 *     word (src/IMSAME.c:229-231): the index of the next stored base is
 *     recorded in break_pos[] when it is not already a read start.
 * The file is read in one piece and parsed with a class table instead of the
 * reference's char-at-a-time buffered reader (src/commonFunctions.c:15-23).
 */
#include "imsame_host.h"
#include <stdlib.h>
#include <string.h>

enum { C_BASE = 0, C_NL = 1, C_GT = 2, C_OTHER = 3 };

int imsame_fasta_load(const char *path, int is_db, imsame_fasta *out) {
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(path, "rb");
    if (!f) return IMSAME_EARG;
    if (fseeko(f, 0, SEEK_END)) { fclose(f); return IMSAME_EARG; }
    off_t flen = ftello(f);
    fseeko(f, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc((size_t)flen + 1);
    if (!buf) { fclose(f); return IMSAME_ENOMEM; }
    size_t got = fread(buf, 1, (size_t)flen, f);
    fclose(f);
    if (got != (size_t)flen) { free(buf); return IMSAME_EARG; }

    unsigned char cls[256], up[256];
    for (int c = 0; c < 256; c++) { cls[c] = C_OTHER; up[c] = (unsigned char)c; }
    cls['\n'] = C_NL;
    cls['>'] = C_GT;
    const char *b = "ACGTacgt";
    for (int i = 0; i < 8; i++) { cls[(unsigned char)b[i]] = C_BASE; up[(unsigned char)b[i]] = (unsigned char)b[i & 3]; }

    /* the stored sequence is never longer than the file: parse in place into a second buffer */
    unsigned char *seq = (unsigned char *)malloc((size_t)flen + 64);
    uint64_t cap_s = 1 << 16, cap_b = 64;
    uint64_t *start = (uint64_t *)malloc(cap_s * sizeof(uint64_t));
    uint64_t *brk = (uint64_t *)malloc(cap_b * sizeof(uint64_t));
    if (!seq || !start || !brk) { free(buf); free(seq); free(start); free(brk); return IMSAME_ENOMEM; }
    uint64_t pos = 0, n = 0, nb = 0;
    size_t i = 0, end = (size_t)flen;
    while (i < end && buf[i] != '>') i++;
    while (i < end) {
        /* buf[i] == '>' */
        if (n + 2 > cap_s) {
            cap_s *= 2;
            start = (uint64_t *)realloc(start, cap_s * sizeof(uint64_t));
            if (!start) { free(buf); free(seq); free(brk); return IMSAME_ENOMEM; }
        }
        start[n++] = pos;
        unsigned char *nl = (unsigned char *)memchr(buf + i, '\n', end - i);
        if (!nl) break;
        i = (size_t)(nl - buf) + 1;
        int pending = 0;
        while (i < end) {
            unsigned char c = buf[i];
            unsigned k = cls[c];
            if (k == C_BASE) {
                if (pending) {
                    if (is_db && pos > start[n - 1]) {
                        if (nb + 1 > cap_b) {
                            cap_b *= 2;
                            brk = (uint64_t *)realloc(brk, cap_b * sizeof(uint64_t));
                            if (!brk) { free(buf); free(seq); free(start); return IMSAME_ENOMEM; }
                        }
                        brk[nb++] = pos;
                    }
                    pending = 0;
                }
                seq[pos++] = up[c];
            } else if (k == C_GT) {
                break;
            } else if (k == C_OTHER) {
                pending = 1;
            }
            i++;
        }
    }
    start[n] = pos;
    free(buf);
    out->sequences = seq;
    out->start_pos = start;
    out->break_pos = brk;
    out->total_len = pos;
    out->n_seqs = n;
    out->n_breaks = nb;
    return IMSAME_OK;
}

void imsame_fasta_free(imsame_fasta *f) {
    free(f->sequences);
    free(f->start_pos);
    free(f->break_pos);
    memset(f, 0, sizeof(*f));
}

void imsame_fasta_view(const imsame_fasta *f, imsame_seqinfo *v) {
    v->sequences = f->sequences;
    v->start_pos = f->start_pos;
    v->total_len = f->total_len;
    v->n_seqs = f->n_seqs;
    v->break_pos = f->break_pos;
    v->n_breaks = f->n_breaks;
}
