#!/bin/sh
# Compile the UNMODIFIED reference from its own sources, where they lie, into
# oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
# Flags are those of the reference's src/Makefile:3 (-w instead of -Wall).
# No reference source is copied into this repository.
set -e
REF=${IMSAME_REFERENCE_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")" && pwd)
if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$HERE/_ref"
gcc -O3 -D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE -w -DVERBOSE \
    "$REF/alignmentFunctions.c" "$REF/commonFunctions.c" "$REF/IMSAME.c" -lpthread -lm -o "$HERE/_ref/IMSAME"
gcc -O3 -D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE -w \
    "$REF/commonFunctions.c" "$REF/reverseComplement.c" -o "$HERE/_ref/revComp"
# the reference's own workflow script, unmodified, next to the reference binaries it calls through its BINDIR
# (bin/all_vs_all_metagenomes_IMSAME.sh:10): used by tools/allvsall_bench.py --reference (cfg4 baseline)
cp "$REF/../bin/all_vs_all_metagenomes_IMSAME.sh" "$HERE/_ref/all_vs_all_metagenomes_IMSAME.sh"
chmod +x "$HERE/_ref/all_vs_all_metagenomes_IMSAME.sh"
echo "build_ref: built $HERE/_ref/IMSAME and revComp"
