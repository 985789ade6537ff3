// traceback.cuh -- K4: winners-only traceback (src/alignmentFunctions.c:493-546).
// The NW kernel is re-run on the accepted (read, db_seq) pairs only (at most one
// per query read) with TB = true, storing one 16-bit back-pointer code per cell
// (nw_core.cuh); one thread per pair then walks from the best border cell back
// to row 0 / column 0 and emits the path as run-length ops.  The host renderer
// (host/render.c) turns ops + sequences into the reference's alignment text.
#pragma once
#include "nw_core.cuh"

namespace imsame {

// op word = type << 28 | count, in traceback order (from the best cell backwards)
constexpr uint32_t OP_DIAG = 1u << 28;  // count diagonal steps: X[x] over Y[y]
constexpr uint32_t OP_COL = 2u << 28;   // jump up a column: count bases of X over '-', then y -= 1
constexpr uint32_t OP_ROW = 3u << 28;   // jump along a row: '-' over count bases of Y, then x -= 1
constexpr uint32_t OP_COUNT_MASK = (1u << 28) - 1;

// tb: codes of rows 1..X1 (row i at (i-1)*stride), columns 1..Y1 (column j at j-1)
IMS_HD uint32_t tb_walk(const uint16_t *tb, uint32_t stride, uint32_t bx, uint32_t by, uint32_t *ops,
                        uint32_t *ex, uint32_t *ey) {
    uint32_t x = bx, y = by, n = 0, run = 0;
    while (x > 0 && y > 0) {
        const uint16_t code = tb[(uint64_t)(x - 1) * stride + (y - 1)];
        if (code == TB_DIAG) {
            run++;
            x--;
            y--;
            continue;
        }
        if (run) { ops[n++] = OP_DIAG | run; run = 0; }
        if (code & TB_COL) {
            const uint32_t ux = code & TB_MASK;
            ops[n++] = OP_COL | (x - ux);
            x = ux;
            y -= 1;
        } else {
            const uint32_t uy = code & TB_MASK;
            ops[n++] = OP_ROW | (y - uy);
            y = uy;
            x -= 1;
        }
    }
    if (run) ops[n++] = OP_DIAG | run;
    *ex = x;
    *ey = y;
    return n;
}

IMS_HD uint32_t tb_stride(uint32_t ylen, int s_class) {
    const uint32_t block = 32u * (uint32_t)s_class;
    const uint32_t y1 = ylen > 1 ? ylen - 1 : 1;
    return (y1 + block - 1) / block * block;
}

}  // namespace imsame
