// nwp_launch.cu -- instantiations of nwp_kernel<S, CL> (nwp.cuh): S = 2 * class columns per lane, CL = register
// slot of the last query column (or -1: decided at run time).  A translation unit of its own: these ~150 kernels
// are most of the library's compile time.
#include "nwp_launch.h"
#include "nwp.cuh"

namespace imsame {

namespace {

template <int C, int CL>
void launch_variant(int grid, cudaStream_t stream, const NwArgs &a, int cl) {
    if (cl == CL) {
        nwp_kernel<2 * C, CL><<<grid, NWP_THREADS, 0, stream>>>(a);
        return;
    }
    if constexpr (CL + 1 < 2 * C) launch_variant<C, CL + 1>(grid, stream, a, cl);
    else nwp_kernel<2 * C, -1><<<grid, NWP_THREADS, 0, stream>>>(a);
}

template <int C>
void launch_class(int grid, cudaStream_t stream, const NwArgs &a) {
    int cl = -1;
    if (a.q.fixed_len >= 2 && !a.check_class && nw_class_of(a.q.fixed_len) == C) cl = (int)((a.q.fixed_len - 2) % (2 * C));
    if (cl >= 0) launch_variant<C, 0>(grid, stream, a, cl);
    else nwp_kernel<2 * C, -1><<<grid, NWP_THREADS, 0, stream>>>(a);
}

template <int C>
int blocks_per_sm() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nwp_kernel<2 * C, -1>, NWP_THREADS, 0) != cudaSuccess) return -1;
    return per_sm;
}

}  // namespace

int nwp_blocks_per_sm(int c) {
    switch (c) {
        case 1: return blocks_per_sm<1>();
        case 2: return blocks_per_sm<2>();
        case 3: return blocks_per_sm<3>();
        case 4: return blocks_per_sm<4>();
        case 5: return blocks_per_sm<5>();
        case 6: return blocks_per_sm<6>();
        case 7: return blocks_per_sm<7>();
        default: return blocks_per_sm<8>();
    }
}

void nwp_launch(int c, int grid, cudaStream_t stream, const NwArgs &a) {
    switch (c) {
        case 1: launch_class<1>(grid, stream, a); break;
        case 2: launch_class<2>(grid, stream, a); break;
        case 3: launch_class<3>(grid, stream, a); break;
        case 4: launch_class<4>(grid, stream, a); break;
        case 5: launch_class<5>(grid, stream, a); break;
        case 6: launch_class<6>(grid, stream, a); break;
        case 7: launch_class<7>(grid, stream, a); break;
        default: launch_class<8>(grid, stream, a); break;
    }
}

}  // namespace imsame
