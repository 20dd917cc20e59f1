// Shift-GEMM convolution on tcgen05 (sm_100a).
//
// Activations live in HBM in a "chunk-plane, padded-linear" layout:
//     act[chunk j = c/8][pixel p][8 channels]   (16 bytes per (j, p), 16-bit elements)
// where pixels enumerate every line of every utterance back to back with ONE zero pad column per
// line (w == W) and ONE zero pad line between utterances:  p = G + row * Wp + w,  Wp = W + 1.
// In that layout a 3x3 / stride-1 / pad-1 convolution is nine GEMMs whose A operands are the SAME
// pixel slab shifted by (r-1)*Wp + (s-1) rows, and a contiguous run of pixels of one chunk plane is
// exactly one column of 8x16-byte UMMA core matrices (K-major, no swizzle).  So a CTA
//   * pulls its slab (128*MT output pixels + halo) with one 1-D bulk copy per chunk plane
//     (cp.async.bulk, no tensor map, every input byte read ~once),
//   * streams the folded-BN weights, pre-packed on the host into per-(tap, k-chunk) UMMA images,
//   * issues tcgen05.mma M=128 x N x K=16 from one thread, descriptors differing per tap only in
//     their start address, accumulating MT tiles in TMEM,
//   * and four epilogue warps read TMEM with tcgen05.ld, add the bias, apply ReLU / the pad mask,
//     optionally subsample (stride-2 convs are computed at stride 1 and every other pixel kept),
//     optionally reduce per-(utterance, channel) sums for the squeeze-excitation layer, and store
//     16-byte channel groups, fully coalesced across the warp.
// Replaces cuDNN conv + BN + ReLU kernels behind sidekit/nnet/res_net.py:309-320 and the Conv1d
// GEMMs of sidekit/nnet/xvector.py:467-483.
#pragma once
#include "common.cuh"

namespace skb {

struct ConvParams {
    const uint16_t* in;     // input planes
    long long in_plane;     // pixels per input chunk plane
    const uint16_t* w;      // packed weights: [n_split][kc][tap] images of [4][N_CTA][8]
    const float* bias;      // [cout] folded BN bias (fp32)
    uint16_t* out;          // output planes
    long long out_plane;
    int cin, cout, taps;
    int Wp, W;              // input/level geometry
    int G;                  // first computed pixel
    int p_end;              // one past the last computed pixel
    int halo;               // slab rows before the tile's first pixel (Wp + 1 for 3x3, 0 for 1x1 / causal taps)
    int rows_pad;           // slab rows per chunk plane (multiple of 8)
    int tap_shift[10];      // pixel shift of each tap: (r-1)*Wp + (s-1) for 3x3; k*dilation for the TDNN
    int act;                // 0 none, 1 ReLU, 2 LeakyReLU(0.2)
    const int* row_b;       // [n_rows] utterance of each line, -1 for pad lines
    const int* row_h;       // [n_rows] line index inside its utterance, -1 for pad lines
    int subsample;          // 1: keep even (h, w) only and write into the next level's geometry
    int out_G, out_Wp;
    const int* out_utt_row0;  // [B] first line of each utterance at the output level
    unsigned long long* se_sums;   // [B][cout] fixed-point (2^24) channel sums, or nullptr
};

constexpr int kConvKC = 32;            // input channels per A/B stage (two K=16 MMAs)
constexpr int kConvAStages = 2;
constexpr int kConvThreads = 7 * 32;   // warps: 0 A-producer, 1 B-producer, 2 MMA, 3..6 epilogue

template <int N_CTA, int MT>
struct ConvCfg {
    static constexpr int kTmemCols = N_CTA * MT;   // 128 or 256 (power of two)
    static constexpr int kBStageBytes = N_CTA * kConvKC * 2;
    static constexpr int kBStages = (N_CTA <= 32) ? 8 : ((N_CTA <= 64) ? 6 : 4);
    static constexpr int kTileM = 128 * MT;
    static size_t smem_bytes(int rows_pad) {
        return 1024 + (size_t)kConvAStages * rows_pad * (kConvKC / 8) * 16 + (size_t)kBStages * kBStageBytes;
    }
};

template <int N_CTA, int MT, bool BF16>
__global__ void __launch_bounds__(kConvThreads) conv_umma_kernel(const ConvParams p) {
    using Cfg = ConvCfg<N_CTA, MT>;
    extern __shared__ __align__(128) uint8_t smem[];
    // control block (first 1024 bytes)
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);           // [2]
    uint64_t* a_empty = a_full + kConvAStages;                      // [2]
    uint64_t* b_full = a_empty + kConvAStages;                      // [kBStages]
    uint64_t* b_empty = b_full + Cfg::kBStages;                     // [kBStages]
    uint64_t* acc_full = b_empty + Cfg::kBStages;                   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    uint8_t* a_smem = smem + 1024;
    const uint32_t a_stage_bytes = (uint32_t)p.rows_pad * (kConvKC / 8) * 16;
    uint8_t* b_smem = a_smem + kConvAStages * a_stage_bytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int nsplit = blockIdx.y;
    const int p0 = p.G + tile * Cfg::kTileM;
    const int n_kc = p.cin / kConvKC;
    const int n_it = n_kc * p.taps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kConvAStages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < Cfg::kBStages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- A producer: activation slabs
        if (lane == 0) {
            const long long q0 = (long long)p0 - p.halo;    // first slab pixel (>= 0 thanks to the guard G)
            for (int kc = 0; kc < n_kc; ++kc) {
                const int s = kc % kConvAStages;
                const int u = kc / kConvAStages;
                mbar_wait(&a_empty[s], (u & 1) ^ 1);
                mbar_arrive_expect_tx(&a_full[s], a_stage_bytes);
                const uint32_t plane_bytes = (uint32_t)p.rows_pad * 16;
#pragma unroll
                for (int j = 0; j < kConvKC / 8; ++j) {
                    const uint16_t* src = p.in + ((size_t)(kc * (kConvKC / 8) + j) * p.in_plane + q0) * 8;
                    bulk_g2s(a_smem + s * a_stage_bytes + j * plane_bytes, src, plane_bytes, &a_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- B producer: packed weights
        if (lane == 0) {
            const uint16_t* wbase = p.w + (size_t)nsplit * n_it * (Cfg::kBStageBytes / 2);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % Cfg::kBStages;
                const int u = it / Cfg::kBStages;
                mbar_wait(&b_empty[s], (u & 1) ^ 1);
                mbar_arrive_expect_tx(&b_full[s], Cfg::kBStageBytes);
                bulk_g2s(b_smem + s * Cfg::kBStageBytes, wbase + (size_t)it * (Cfg::kBStageBytes / 2),
                         Cfg::kBStageBytes, &b_full[s]);
            }
        }
    } else if (warp == 2) {
        // ---------------------------------------------------------------- MMA issuer (single thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, N_CTA, BF16);
            const uint32_t a_lbo = (uint32_t)p.rows_pad * 16;   // next 8-channel plane of the slab
            const uint32_t b_lbo = N_CTA * 16;
            int it = 0;
            for (int kc = 0; kc < n_kc; ++kc) {
                const int as = kc % kConvAStages;
                mbar_wait(&a_full[as], (kc / kConvAStages) & 1);
                const uint32_t a_base = smem_u32(a_smem + as * a_stage_bytes);
                for (int tap = 0; tap < p.taps; ++tap, ++it) {
                    const int bs = it % Cfg::kBStages;
                    mbar_wait(&b_full[bs], (it / Cfg::kBStages) & 1);
                    tc_fence_after();
                    const int shift = p.tap_shift[tap];
                    const uint32_t b_base = smem_u32(b_smem + bs * Cfg::kBStageBytes);
#pragma unroll
                    for (int ks = 0; ks < kConvKC / 16; ++ks) {
                        const uint64_t bdesc = umma_desc_kmajor_noswz(b_base + ks * 2 * b_lbo, b_lbo, 128);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint32_t a_addr = a_base + ks * 2 * a_lbo + (uint32_t)(p.halo + shift + mt * 128) * 16;
                            const uint64_t adesc = umma_desc_kmajor_noswz(a_addr, a_lbo, 128);
                            umma_f16(tmem_base + mt * N_CTA, adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&b_empty[bs]);
                }
                umma_commit(&a_empty[as]);
            }
            umma_commit(acc_full);
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps (3..6)
        const int q = warp & 3;   // TMEM lane quadrant this warp may read
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int n_base = nsplit * N_CTA;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const int pix = p0 + mt * 128 + q * 32 + lane;
            const bool in_range = pix < p.p_end;
            const int rel = pix - p.G;
            const int row = rel / p.Wp;
            const int w = rel - row * p.Wp;
            int b = -1, h = -1;
            if (in_range) {
                b = __ldg(p.row_b + row);
                h = __ldg(p.row_h + row);
            }
            const bool valid = in_range && (w < p.W) && (h >= 0);
            bool do_store = in_range;
            long long opix = pix;
            if (p.subsample) {
                do_store = valid && !(h & 1) && !(w & 1);
                if (do_store) opix = (long long)p.out_G + (long long)(__ldg(p.out_utt_row0 + b) + (h >> 1)) * p.out_Wp + (w >> 1);
            }
#pragma unroll 1
            for (int c0 = 0; c0 < N_CTA; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mt * N_CTA + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float t = v[i] + __ldg(p.bias + n_base + c0 + i);
                    if (p.act == 1) t = fmaxf(t, 0.f);
                    else if (p.act == 2) t = t > 0.f ? t : 0.2f * t;
                    v[i] = valid ? t : 0.f;
                }
                if (do_store) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack2<BF16>(v[j * 8 + 0], v[j * 8 + 1]);
                        o.y = pack2<BF16>(v[j * 8 + 2], v[j * 8 + 3]);
                        o.z = pack2<BF16>(v[j * 8 + 4], v[j * 8 + 5]);
                        o.w = pack2<BF16>(v[j * 8 + 6], v[j * 8 + 7]);
                        uint16_t* dst = p.out + ((size_t)((n_base + c0) / 8 + j) * p.out_plane + opix) * 8;
                        *reinterpret_cast<uint4*>(dst) = o;
                    }
                }
                if (p.se_sums != nullptr) {
                    // per-(utterance, channel) sums of the fp32 values for the SE squeeze, accumulated in
                    // 2^-24 fixed point: integer addition is associative, so the result is bit-identical
                    // whatever the tile / warp / atomic order (and however the batch is packed).  A warp's
                    // 32 pixels almost always belong to one utterance; the loop covers the boundaries.
                    unsigned vmask = __ballot_sync(0xffffffffu, valid);
                    while (vmask) {
                        const int leader = __ffs(vmask) - 1;
                        const int b0 = __shfl_sync(0xffffffffu, b, leader);
                        const bool mine = valid && (b == b0);
                        const unsigned mm = __ballot_sync(0xffffffffu, mine);
                        long long t[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) t[i] = mine ? __float2ll_rn(v[i] * 16777216.f) : 0ll;
                        // transposing butterfly: 31 exchanges, lane l ends with the sum of channel c0 + l
#pragma unroll
                        for (int s = 16; s >= 1; s >>= 1) {
                            const bool upper = (lane & s) != 0;
#pragma unroll
                            for (int i = 0; i < s; ++i) {
                                const long long send = upper ? t[i] : t[i + s];
                                const long long keep = upper ? t[i + s] : t[i];
                                t[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                            }
                        }
                        atomicAdd(p.se_sums + (size_t)b0 * p.cout + n_base + c0 + lane, (unsigned long long)t[0]);
                        vmask &= ~mm;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace skb
