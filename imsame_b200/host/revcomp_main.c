/*
 * revComp -- drop-in for the reference's reverse-complement tool
 * (src/reverseComplement.c:21-118), used by all_vs_all_metagenomes_IMSAME.sh:47.
 *
 * Behaviour kept: records are emitted in REVERSE file order (:56); a record
 * starts at every '>' byte (:47-52); the header line is copied verbatim (:59-62);
 * only letters are kept from the body (:65-70) and written reverse-complemented on
 * ONE line (:71-112) with A<->T, C<->G, U->A (case preserved), every other letter
 * unchanged.  Not kept: the 1 000 000-record stack array and the fixed 2 GB
 * buffer (:17-19,25,34) -- the file is read in one piece instead.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void terror(const char *s) { /* src/commonFunctions.c:10-13 */
    printf("ERR**** %s ****\n", s);
    exit(-1);
}

int main(int ac, char **av) {
    if (ac != 3) terror("USE: reverseComplement seqFile.IN reverseComplementarySeq.OUT");
    FILE *fi = fopen(av[1], "rb");
    if (!fi) terror("opening IN sequence FASTA file");
    FILE *fo = fopen(av[2], "wb");
    if (!fo) terror("opening OUT sequence Words file");
    fseeko(fi, 0, SEEK_END);
    off_t n = ftello(fi);
    fseeko(fi, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc((size_t)n + 1);
    if (!buf) terror("memory for Seq");
    if (fread(buf, 1, (size_t)n, fi) != (size_t)n) terror("Empty file");
    fclose(fi);
    unsigned char comp[256];
    for (int c = 0; c < 256; c++) comp[c] = (unsigned char)c;
    comp['A'] = 'T'; comp['C'] = 'G'; comp['G'] = 'C'; comp['T'] = 'A'; comp['U'] = 'A';
    comp['a'] = 't'; comp['c'] = 'g'; comp['g'] = 'c'; comp['t'] = 'a'; comp['u'] = 'a';
    size_t cap = 1024, nrec = 0;
    off_t *off = (off_t *)malloc(cap * sizeof(off_t));
    for (off_t i = 0; i < n; i++)
        if (buf[i] == '>') {
            if (nrec == cap) { cap *= 2; off = (off_t *)realloc(off, cap * sizeof(off_t)); }
            off[nrec++] = i;
        }
    unsigned char *seq = (unsigned char *)malloc((size_t)n + 2);
    for (size_t r = nrec; r-- > 0;) {
        /* header = up to the first newline (fgets, :59), even if it holds another '>';
           body = up to the next '>' after the header (:65) */
        off_t i = off[r], h = i;
        while (h < n && buf[h] != '\n') h++;
        if (h < n) h++;
        off_t end = h;
        while (end < n && buf[end] != '>') end++;
        fwrite(buf + i, 1, (size_t)(h - i), fo);
        size_t len = 0;
        for (off_t k = end; k-- > h;)
            if (isupper(buf[k]) || islower(buf[k])) seq[len++] = comp[buf[k]];
        seq[len] = '\n';
        fwrite(seq, 1, len + 1, fo);
    }
    fclose(fo);
    free(buf); free(off); free(seq);
    return 0;
}
