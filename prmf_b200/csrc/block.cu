// block.cu -- translation unit of the persistent step kernel (block.cuh): its ten instantiations (k = 1..10) compile
// here, in parallel with the rest of the library.
#include "block.cuh"

using namespace prmf;

namespace {
template <int K>
cudaError_t launch_t(const BlockParams& prm, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(block_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    BlockParams copy = prm;
    void* args[] = {(void*)&copy};
    // cooperative: the CTAs wait for each other, so co-residency must be guaranteed by the driver
    return cudaLaunchCooperativeKernel((const void*)block_kernel<K>, dim3((unsigned)grid), dim3(kBlkThreads), args, smem, stream);
}
}  // namespace

cudaError_t prmf_launch_block_kernel(int k, const BlockParams& prm, int grid, size_t smem, cudaStream_t stream) {
    switch (k) {
        case 1: return launch_t<1>(prm, grid, smem, stream);
        case 2: return launch_t<2>(prm, grid, smem, stream);
        case 3: return launch_t<3>(prm, grid, smem, stream);
        case 4: return launch_t<4>(prm, grid, smem, stream);
        case 5: return launch_t<5>(prm, grid, smem, stream);
        case 6: return launch_t<6>(prm, grid, smem, stream);
        case 7: return launch_t<7>(prm, grid, smem, stream);
        case 8: return launch_t<8>(prm, grid, smem, stream);
        case 9: return launch_t<9>(prm, grid, smem, stream);
        case 10: return launch_t<10>(prm, grid, smem, stream);
        default: return cudaErrorInvalidValue;
    }
}

#ifdef PRMF_BLOCK_TIMING
extern "C" int prmf_debug_block_stamps(unsigned long long* out, int count, int reset) {
    if (reset) {
        void* sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_blk_dbg) != cudaSuccess) return -1;
        return cudaMemset(sym, 0, sizeof(unsigned long long) * 160 * kBlkDbgHalves * 8) == cudaSuccess ? 0 : -1;
    }
    return cudaMemcpyFromSymbol(out, g_blk_dbg, sizeof(unsigned long long) * count) == cudaSuccess ? 0 : -1;
}
#endif
