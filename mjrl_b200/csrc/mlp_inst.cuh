// Explicit instantiation helper: one translation unit per hidden width so the build parallelises.
#pragma once
#include "mlp_kernel.cuh"

namespace mjb {

template <int H, int MT, int MODE, int ACT>
static cudaError_t launch_one(const MlpArgs& args, int grid, cudaStream_t stream) {
    const size_t smem = mlp_smem_bytes(H, MT, MODE, args.L.YR);
    auto kern = mlp_kernel<H, MT, MODE, ACT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, smem, stream>>>(args);
    return cudaGetLastError();
}

template <int H, int MT, int MODE, int ACT>
static int occupancy_one(int YR) {
    const size_t smem = mlp_smem_bytes(H, MT, MODE, YR);
    auto kern = mlp_kernel<H, MT, MODE, ACT>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem) != cudaSuccess) return 0;
    return nb;
}

#define MJB_DEFINE_MLP(H, MT)                                                                          \
    cudaError_t launch_mlp_h##H(int mode, const MlpArgs& args, int grid, cudaStream_t stream) {        \
        switch (mode) {                                                                                \
            case MODE_EVAL: return launch_one<H, MT, MODE_EVAL, 0>(args, grid, stream);                \
            case MODE_VPG: return launch_one<H, MT, MODE_VPG, 0>(args, grid, stream);                  \
            case MODE_FVP: return launch_one<H, MT, MODE_FVP, 0>(args, grid, stream);                  \
            case MODE_VF: return launch_one<H, MT, MODE_VF, 1>(args, grid, stream);                    \
        }                                                                                              \
        return cudaErrorInvalidValue;                                                                  \
    }                                                                                                  \
    int occupancy_mlp_h##H(int mode, int YR) {                                                         \
        switch (mode) {                                                                                \
            case MODE_EVAL: return occupancy_one<H, MT, MODE_EVAL, 0>(YR);                             \
            case MODE_VPG: return occupancy_one<H, MT, MODE_VPG, 0>(YR);                               \
            case MODE_FVP: return occupancy_one<H, MT, MODE_FVP, 0>(YR);                               \
            case MODE_VF: return occupancy_one<H, MT, MODE_VF, 1>(YR);                                 \
        }                                                                                              \
        return 0;                                                                                      \
    }

}  // namespace mjb
