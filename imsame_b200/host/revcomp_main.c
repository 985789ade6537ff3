/*
 * revComp -- drop-in for the reference's reverse-complement tool
 * (src/reverseComplement.c:21-118), used by all_vs_all_metagenomes_IMSAME.sh:47.
 *
 * Behaviour kept: records are emitted in REVERSE file order (:56); a record
 * starts at every '>' byte (:47-52); the header line is copied verbatim (:59-62);
 * only letters are kept from the body (:65-70) and written reverse-complemented on
 * ONE line (:71-112) with A<->T, C<->G, U->A (case preserved), every other letter
 * unchanged.  Not kept: the 1 000 000-record stack array and the fixed 2 GB
 * buffer (:17-19,25,34) -- the file is mapped in one piece instead.  The transformation itself is
 * imsame_revcomp_mem (host/fasta.c), shared with the in-process all-vs-all driver.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "imsame_host.h"

static void terror(const char *s) { /* src/commonFunctions.c:10-13 */
    printf("ERR**** %s ****\n", s);
    exit(-1);
}

int main(int ac, char **av) {
    if (ac != 3) terror("USE: reverseComplement seqFile.IN reverseComplementarySeq.OUT");
    imsame_file_image img;
    if (imsame_file_map(av[1], &img)) terror("opening IN sequence FASTA file");
    FILE *fo = fopen(av[2], "wb");
    if (!fo) terror("opening OUT sequence Words file");
    unsigned char *out = NULL;
    size_t out_len = 0;
    if (imsame_revcomp_mem(img.data, img.len, &out, &out_len)) terror("memory for Seq"); /* host/fasta.c */
    if (fwrite(out, 1, out_len, fo) != out_len) terror("writing OUT sequence file");
    fclose(fo);
    free(out);
    imsame_file_unmap(&img);
    return 0;
}
