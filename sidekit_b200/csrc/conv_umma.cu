// Host launcher for the tcgen05 shift-GEMM convolution (kernel in conv_umma.cuh).
#include "sidekit_b200.h"
#include "conv_umma.cuh"
#include "layers.cuh"

namespace skb {

int conv_pick_ncta(int cout) { return cout <= 32 ? 32 : (cout <= 64 ? 64 : 128); }
int conv_tile_m(int n_cta) { return n_cta == 32 ? 512 : 256; }

template <int N_CTA, int MT, bool BF16>
static int launch_one(const ConvParams& p, cudaStream_t st) {
    using Cfg = ConvCfg<N_CTA, MT>;
    static bool configured = false;
    const size_t smem = Cfg::smem_bytes(p.rows_pad);
    if (smem > 227 * 1024) {
        set_last_error(__FILE__, __LINE__, "conv slab does not fit in shared memory");
        return SKB_ERR_ARG;
    }
    if (!configured) {
        SKB_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<N_CTA, MT, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024));
        configured = true;
    }
    const int n_pix = p.p_end - p.G;
    dim3 grid((n_pix + Cfg::kTileM - 1) / Cfg::kTileM, p.cout / N_CTA);
    conv_umma_kernel<N_CTA, MT, BF16><<<grid, kConvThreads, smem, st>>>(p);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

int launch_conv_umma(const ConvParams& p, int n_cta, bool bf16, cudaStream_t st) {
    if (p.cin % kConvKC != 0 || p.cout % n_cta != 0 || p.taps < 1 || p.taps > 10 || p.rows_pad % 8 != 0) {
        set_last_error(__FILE__, __LINE__, "conv_umma: unsupported shape");
        return SKB_ERR_ARG;
    }
    switch (n_cta) {
        case 32: return bf16 ? launch_one<32, 4, true>(p, st) : launch_one<32, 4, false>(p, st);
        case 64: return bf16 ? launch_one<64, 2, true>(p, st) : launch_one<64, 2, false>(p, st);
        case 128: return bf16 ? launch_one<128, 2, true>(p, st) : launch_one<128, 2, false>(p, st);
    }
    set_last_error(__FILE__, __LINE__, "conv_umma: unsupported N tile");
    return SKB_ERR_ARG;
}

}  // namespace skb
