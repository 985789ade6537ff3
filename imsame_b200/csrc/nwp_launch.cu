// nwp_launch.cu -- instantiations of nwp_kernel<S, CL> (nwp.cuh): S columns per lane (2 * class for the classes
// 1..8; 18 and 20 for the wide classes 9 and 10, or 17 and 19 when the run's longest query read fits them: 300-base
// reads fill 299 of 304 columns instead of 299 of 320), CL = register slot of the last query column (or -1: decided
// at run time).  A translation unit of its own: these ~160 kernels are most of the library's compile time.
#include "nwp_launch.h"
#include "nwp.cuh"

namespace imsame {

namespace {

constexpr int cols_of_class(int c) { return 2 * c; }  // 2, 4, .. 16 and 18, 20 (wide)

template <int S, int CL>
void launch_variant(int grid, cudaStream_t stream, const NwArgs &a, int cl) {
    if (cl == CL) {
        nwp_kernel<S, CL><<<grid, nwp_threads(S), 0, stream>>>(a);
        return;
    }
    if constexpr (CL + 1 < S) launch_variant<S, CL + 1>(grid, stream, a, cl);
    else nwp_kernel<S, -1><<<grid, nwp_threads(S), 0, stream>>>(a);
}

template <int C, int S = cols_of_class(C)>
void launch_class(int grid, cudaStream_t stream, const NwArgs &a) {
    int cl = -1;
    if (a.q.fixed_len >= 2 && !a.check_class && nw_class_of(a.q.fixed_len) == C) cl = (int)((a.q.fixed_len - 2) % S);
    if (cl >= 0) launch_variant<S, 0>(grid, stream, a, cl);
    else nwp_kernel<S, -1><<<grid, nwp_threads(S), 0, stream>>>(a);
}

template <int C>
int blocks_per_sm() {
    constexpr int S = cols_of_class(C);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nwp_kernel<S, -1>, nwp_threads(S), 0) != cudaSuccess) return -1;
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
        case 8: return blocks_per_sm<8>();
        case 9: return blocks_per_sm<9>();
        default: return blocks_per_sm<10>();
    }
}

void nwp_launch(int c, int grid, cudaStream_t stream, const NwArgs &a, uint32_t ymax) {
    // the wide classes with one column per lane less when no query read of the run needs it
    if (c == 9 && ymax <= (uint32_t)(PW_LANES * 17 + 1)) { launch_class<9, 17>(grid, stream, a); return; }
    if (c == 10 && ymax <= (uint32_t)(PW_LANES * 19 + 1)) { launch_class<10, 19>(grid, stream, a); return; }
    switch (c) {
        case 1: launch_class<1>(grid, stream, a); break;
        case 2: launch_class<2>(grid, stream, a); break;
        case 3: launch_class<3>(grid, stream, a); break;
        case 4: launch_class<4>(grid, stream, a); break;
        case 5: launch_class<5>(grid, stream, a); break;
        case 6: launch_class<6>(grid, stream, a); break;
        case 7: launch_class<7>(grid, stream, a); break;
        case 8: launch_class<8>(grid, stream, a); break;
        case 9: launch_class<9>(grid, stream, a); break;
        default: launch_class<10>(grid, stream, a); break;
    }
}

}  // namespace imsame
