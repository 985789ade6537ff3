// nwp_launch.h -- host entry points of the packed-word NW kernels (nwp.cuh), compiled in nwp_launch.cu
#pragma once
#include <cuda_runtime.h>
#include "nw.cuh"
#include "nwp_core.cuh"

namespace imsame {

// resident blocks per SM of the class-c kernel (16 lanes x 2c columns per pair; c = 1..NW_CLASSES), < 0 on a CUDA error
int nwp_blocks_per_sm(int c);
// launch the class-c kernel; when every query read of the run has the same length (a.q.fixed_len) the
// variant with the last column's register slot compiled in is chosen; ymax = longest query read of the run / batch
// (the wide classes take 17 / 19 columns per lane instead of 18 / 20 when that is enough)
void nwp_launch(int c, int grid, cudaStream_t stream, const NwArgs &a, uint32_t ymax);

}  // namespace imsame
