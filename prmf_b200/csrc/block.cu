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

namespace {
__global__ void __launch_bounds__(64) small_allreduce_kernel(double* buf, int count, PeerSmall ps) {
    const int t = threadIdx.x;
    if (t >= count) return;
    const size_t par = (size_t)(ps.seq & 1ull) * kMaxPeers;
    const double v = buf[t];
    for (int r = 0; r < ps.nranks; ++r) ll_store(ps.slot[r] + (par + ps.rank) * kSmallAllreduceMax + t, v, ps.seq);
    buf[t] = ll_gather(ps.slot[ps.rank], par, kSmallAllreduceMax, ps.nranks, (size_t)t, ps.seq, ps.err, ps.timeout_ns);
}
}  // namespace

cudaError_t prmf_launch_small_allreduce(double* buf, int count, const PeerSmall& ps, cudaStream_t stream) {
    small_allreduce_kernel<<<1, 64, 0, stream>>>(buf, count, ps);
    return cudaGetLastError();
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
