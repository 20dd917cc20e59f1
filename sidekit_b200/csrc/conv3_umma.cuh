// 3x3 / stride-1 convolution for the NARROW layers (Cout = 32 or 64) with the three horizontal taps fused into the
// MMA's N dimension.
//
// Why: tcgen05.mma M=128, K=16 reads its A operand (4 KB) from shared memory for every instruction, so an instruction
// costs max(N/2, 32 + N/4) cycles and the nine N=32 (N=64) MMAs of a 3x3 tap loop run at 36 % (67 %) of the tensor
// peak (tools/umma_bench.cu, profiles/r01b_ncu_conv_L1_conv1.txt: shared-memory operand pipe 78 % busy).  Here one MMA
// per VERTICAL tap r computes, from ONE read of the slab shifted by (r-1)*Wp,
//     D_s[q][co] += sum_ci in[q + (r-1)*Wp][ci] * W[co][ci][r][s]        for s = 0, 1, 2   (N = 3 * Cout = 96 / 192)
// and the epilogue adds the three partial sums of neighbouring pixels:  out[p] = D_0[p-1] + D_1[p] + D_2[p+1].
// In TMEM a pixel is a lane, so the +-1 shifts are warp shuffles; the lanes at the edge of a warp's 32-pixel group get
// their neighbour's value through a small shared-memory exchange (one named barrier per work item), and consecutive
// items overlap by two pixels (rows 0 and 128*MT-1 of an item are computed but not stored).  Every output is
// a1 + (x0 + x2) with the same association whatever its lane, so results stay independent of the packing.
//
// Same roles and pipelines as conv_umma_kernel (conv_umma.cuh); weights are always resident (<= 72 KB).
#pragma once
#include "conv_umma.cuh"

namespace skb {

template <int N_CTA, int MT>
struct Conv3Cfg {
    static_assert(N_CTA * MT == 64, "fused-tap conv: (32, 2) or (64, 1)");
    static constexpr int kN3 = 3 * N_CTA;              // MMA N
    static constexpr int kAccCols = kN3 * MT;          // 192
    static constexpr int kTmemCols = 512;              // 2 x 192 rounded up to a power of two
    static constexpr int kTileM = 128 * MT;
    static constexpr int kTileOut = kTileM - 2;        // stored pixels per item
    static constexpr int kBImgBytes = kN3 * kConvKC * 2;   // one (k-chunk, vertical tap) weight image
    static constexpr int kMmaWarps = MT >= 2 ? 2 : 1;
    static constexpr int kGroups = MT * 4;             // 32-pixel groups per item
    static constexpr int kXchgBytes = 2 * kGroups * 2 * N_CTA * 4;   // [parity][group][D0 of lane 31 | D2 of lane 0][channel]
    static constexpr int kOnesBytes = 2 * 128 * 16;
    static size_t fixed_bytes(int n_kc, int scale_bytes) {
        return kConvCtrlBytes + kOnesBytes + (size_t)kN3 * 32 + kXchgBytes + scale_bytes + (size_t)n_kc * 3 * kBImgBytes;
    }
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N_CTA, int MT, bool BF16>
__global__ void __launch_bounds__(kConvThreads, 1) conv3_umma_kernel(const ConvParams p) {
    using Cfg = Conv3Cfg<N_CTA, MT>;
    constexpr int N3 = Cfg::kN3;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);           // [kConvMaxAStages]
    uint64_t* a_empty = a_full + kConvMaxAStages;
    uint64_t* b_full = a_empty + kConvMaxAStages;                   // [1]
    uint64_t* acc_full = b_full + 1;                                // [2]
    uint64_t* acc_empty = acc_full + 2;                             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* ones_smem = smem + kConvCtrlBytes;
    uint8_t* biasimg_smem = ones_smem + Cfg::kOnesBytes;            // [2 planes][N3 rows][8 halves]
    float* xchg = reinterpret_cast<float*>(biasimg_smem + (size_t)N3 * 32);
    float* scale_s = xchg + Cfg::kXchgBytes / 4;                    // [n_utt][N_CTA] SE scales (when p.scale_smem_bytes > 0)
    uint8_t* b_smem = reinterpret_cast<uint8_t*>(scale_s) + p.scale_smem_bytes;
    const int n_kc = p.cin / kConvKC;
    uint8_t* a_smem = b_smem + (size_t)n_kc * 3 * Cfg::kBImgBytes;
    const uint32_t a_stage_bytes = (uint32_t)p.rows_pad * (kConvKC / 8) * 16;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = p.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kConvMaxAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], Cfg::kMmaWarps); }
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], Cfg::kMmaWarps); mbar_init(&acc_empty[i], 8); }
        fence_barrier_init();
    }
    {
        // "ones" A operand and the bias B operand (see conv_umma_kernel); only the centre tap's columns carry the bias
        const uint32_t one2 = pack2<BF16>(1.f, 1.f);
        for (int i = threadIdx.x; i < 2 * 128; i += blockDim.x)
            reinterpret_cast<uint4*>(ones_smem)[i] = make_uint4(i < 128 ? one2 : 0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < 2 * N3; i += blockDim.x) {
            const int plane = i / N3, n = i % N3;
            uint32_t w0 = 0u;
            if (plane == 0 && n / N_CTA == 1) {
                const float b = p.bias[n - N_CTA];
                const float2 hi2 = unpack2<BF16>(pack2<BF16>(b, 0.f));
                w0 = pack2<BF16>(b, b - hi2.x);
            }
            reinterpret_cast<uint4*>(biasimg_smem)[i] = make_uint4(w0, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // the SE scale table is small (n_utt x Cout floats): keep it in shared memory so that the epilogue's per-pixel
        // scale rows are shared-memory reads instead of dependent L2 loads
        if (p.scale_smem_bytes > 0)
            for (int i = threadIdx.x; i < p.scale_smem_bytes / 4; i += blockDim.x) scale_s[i] = p.se_scale[i];
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- A producer: activation slabs (halo = one line)
        if (lane == 0) {
            const uint32_t plane_bytes = (uint32_t)p.rows_pad * 16;
            int s = 0;
            uint32_t ph = 1;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const long long q0 = (long long)p.G - 1 + (long long)item * Cfg::kTileOut - p.halo;   // >= 0: G >= Wp + 1
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&a_empty[s], ph);
                    mbar_arrive_expect_tx(&a_full[s], a_stage_bytes);
#pragma unroll
                    for (int j = 0; j < kConvKC / 8; ++j) {
                        const uint16_t* src = p.in + ((size_t)(kc * (kConvKC / 8) + j) * p.in_plane + q0) * 8;
                        bulk_g2s(a_smem + (size_t)s * a_stage_bytes + j * plane_bytes, src, plane_bytes, &a_full[s]);
                    }
                    if (++s == p.a_stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- weights: resident for the life of the CTA
        if (lane == 0) {
            const int n_img = n_kc * 3;
            mbar_arrive_expect_tx(b_full, (uint32_t)n_img * Cfg::kBImgBytes);
            for (int it = 0; it < n_img; ++it)
                bulk_g2s(b_smem + (size_t)it * Cfg::kBImgBytes, p.w3 + (size_t)it * (Cfg::kBImgBytes / 2), Cfg::kBImgBytes, b_full);
        }
    } else if (warp == 2 || warp == 3) {
        // ---------------------------------------------------------------- MMA issuers
        if (warp - 2 < Cfg::kMmaWarps) {
            constexpr int MTW = MT / Cfg::kMmaWarps;
            const int mtw0 = (warp - 2) * MTW;
            const uint32_t idesc = umma_idesc_f16(128, N3, BF16);
            const uint32_t a_lbo = (uint32_t)p.rows_pad * 16;
            const uint32_t b_lbo = N3 * 16;
            const uint64_t desc_hi_a = (static_cast<uint64_t>((a_lbo >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                       (static_cast<uint64_t>(1) << 46);
            const uint64_t desc_hi_b = (static_cast<uint64_t>((b_lbo >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                       (static_cast<uint64_t>(1) << 46);
            int as = 0;
            uint32_t a_ph = 0, n_done = 0;
            mbar_wait(b_full, 0);
            tc_fence_after();
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
                const int buf = (int)(n_done & 1);
                mbar_wait(&acc_empty[buf], ((n_done >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * Cfg::kAccCols;
                if (elect_one()) {
                    const uint64_t ones_desc = (static_cast<uint64_t>(2048 >> 4) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                               (static_cast<uint64_t>(1) << 46) | ((smem_u32(ones_smem) >> 4) & 0x3FFF);
                    const uint64_t bias_desc = desc_hi_b | ((smem_u32(biasimg_smem) >> 4) & 0x3FFF);
#pragma unroll
                    for (int mt = mtw0; mt < mtw0 + MTW; ++mt) umma_f16(d_tmem + mt * N3, ones_desc, bias_desc, idesc, 0u);
                }
                __syncwarp();
                for (int kc = 0; kc < n_kc; ++kc) {
                    mbar_wait(&a_full[as], a_ph);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(a_smem + (size_t)as * a_stage_bytes);
                    const uint32_t b_base = smem_u32(b_smem) + (uint32_t)(kc * 3) * Cfg::kBImgBytes;
                    if (elect_one()) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint32_t a_tap = a_base + (uint32_t)(r * p.halo) * 16;      // slab row 0 = tile row 0 - Wp
                            const uint32_t b_tap = b_base + r * Cfg::kBImgBytes;
#pragma unroll
                            for (int ks = 0; ks < kConvKC / 16; ++ks) {
                                const uint64_t bdesc = desc_hi_b | (((b_tap + ks * 2 * b_lbo) >> 4) & 0x3FFF);
#pragma unroll
                                for (int mt = mtw0; mt < mtw0 + MTW; ++mt) {
                                    const uint64_t adesc = desc_hi_a | (((a_tap + ks * 2 * a_lbo + mt * 2048) >> 4) & 0x3FFF);
                                    umma_f16(d_tmem + mt * N3, adesc, bdesc, idesc, 1u);
                                }
                            }
                        }
                        umma_commit(&a_empty[as]);
                    }
                    __syncwarp();
                    if (++as == p.a_stages) { as = 0; a_ph ^= 1; }
                }
                if (elect_one()) umma_commit(&acc_full[buf]);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue warps (4..11): 32 pixels x 32 channels each
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        const int h = (warp - 4) >> 2;
        const int mt = MT == 2 ? h : 0;
        const int ch0 = MT == 2 ? 0 : h * 32;          // first of this warp's 32 channels
        const int g = mt * 4 + q;                      // 32-pixel group of the item
        uint32_t n_done = 0;
        const size_t plane8 = (size_t)p.out_plane * 8;
        const size_t rplane8 = (size_t)p.res_plane * 8;
        const float slope = p.act_slope;
        const bool fused = p.se_scale != nullptr;
        // Per-pixel metadata and the residual come from HBM / L2 (~1 us away) and depend on each other (pixel ->
        // utterance -> residual / scale row).  They are fetched ONE ITEM AHEAD: without that every item paid the whole
        // dependent chain (the first version of this kernel: 10 k cycles per item against 0.7 k of MMAs).
        const int row = g * 32 + lane;                 // row of the item; rows 0 and kTileM-1 only feed their neighbours
        const bool row_ok = row >= 1 && row <= Cfg::kTileM - 2;
        auto item_pix = [&](int item) { return (long long)p.G - 1 + (long long)item * Cfg::kTileOut + row; };
        auto fetch_meta = [&](int item, int& bidx_o, long long& opix_o) {
            const long long px = item_pix(item);
            const bool in_range = item < n_items && row_ok && px < p.p_end;
            bidx_o = in_range ? __ldg(p.pix_b + (px - p.G)) : -1;
            opix_o = in_range ? px : -1;
            if (p.pix_sub != nullptr) opix_o = in_range ? (long long)__ldg(p.pix_sub + (px - p.G)) : -1;
        };
        auto fetch_res = [&](int item, int bidx_i, uint4 (&r)[4]) {
            const long long px = item_pix(item);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = make_uint4(0u, 0u, 0u, 0u);
                if (fused && bidx_i >= 0) r[k] = *reinterpret_cast<const uint4*>(p.res + (size_t)px * 8 + (size_t)((ch0 >> 3) + k) * rplane8);
            }
        };
        // pipeline: metadata two items ahead, residual one item ahead
        int bidx_n, bidx_nn;
        long long opix_n, opix_nn;
        uint4 rv_n[4];
        fetch_meta(blockIdx.x, bidx_n, opix_n);
        fetch_meta(blockIdx.x + gridDim.x, bidx_nn, opix_nn);
        fetch_res(blockIdx.x, bidx_n, rv_n);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
            const int buf = (int)(n_done & 1);
            const int bidx = bidx_n;
            const long long opix = opix_n;
            uint4 rv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rv[k] = rv_n[k];
            uint16_t* optr = opix >= 0 ? p.out + (size_t)opix * 8 : nullptr;
            const bool valid = bidx >= 0;
            bidx_n = bidx_nn; opix_n = opix_nn;
            fetch_res(item + gridDim.x, bidx_n, rv_n);                       // next item's residual: its metadata is already here
            fetch_meta(item + 2 * gridDim.x, bidx_nn, opix_nn);              // metadata of the item after that
            float* xq = xchg + (size_t)(n_done & 1) * (Cfg::kGroups * 2 * N_CTA);
            mbar_wait(&acc_full[buf], (n_done >> 1) & 1);
            tc_fence_after();
            const uint32_t cbase = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::kAccCols + mt * N3;
            // phase A: the horizontal neighbours' partial sums
            float t[32];
            {
                uint32_t a0[2][16], a2[2][16];
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
                    tmem_ld16_nowait(cbase + ch0 + grp * 16, a0[grp]);
                    tmem_ld16_nowait(cbase + 2 * N_CTA + ch0 + grp * 16, a2[grp]);
                }
                tmem_ld_wait();
                // exports first, each as ONE divergent region of eight 16-byte stores (a branch per value cost ~3 k cycles)
                if (lane == 31) {                                  // D0 of my last pixel -> next group
                    float4* dst = reinterpret_cast<float4*>(xq + (g * 2 + 0) * N_CTA + ch0);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        dst[k] = make_float4(__uint_as_float(a0[k >> 2][(k & 3) * 4]), __uint_as_float(a0[k >> 2][(k & 3) * 4 + 1]),
                                             __uint_as_float(a0[k >> 2][(k & 3) * 4 + 2]), __uint_as_float(a0[k >> 2][(k & 3) * 4 + 3]));
                }
                if (lane == 0) {                                   // D2 of my first pixel -> previous group
                    float4* dst = reinterpret_cast<float4*>(xq + (g * 2 + 1) * N_CTA + ch0);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        dst[k] = make_float4(__uint_as_float(a2[k >> 2][(k & 3) * 4]), __uint_as_float(a2[k >> 2][(k & 3) * 4 + 1]),
                                             __uint_as_float(a2[k >> 2][(k & 3) * 4 + 2]), __uint_as_float(a2[k >> 2][(k & 3) * 4 + 3]));
                }
                const bool first = lane == 0, last = lane == 31;
#pragma unroll
                for (int grp = 0; grp < 2; ++grp)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float s0 = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[grp][i]), 1);
                        const float s2 = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[grp][i]), 1);
                        t[grp * 16 + i] = first ? s2 : (last ? s0 : s0 + s2);
                    }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");       // all eight epilogue warps: exchange buffer complete
            if (lane == 0 && g > 0) {
                const float4* src = reinterpret_cast<const float4*>(xq + ((g - 1) * 2 + 0) * N_CTA + ch0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 x = src[k];
                    t[4 * k] = x.x + t[4 * k]; t[4 * k + 1] = x.y + t[4 * k + 1]; t[4 * k + 2] = x.z + t[4 * k + 2]; t[4 * k + 3] = x.w + t[4 * k + 3];
                }
            }
            if (lane == 31 && g < Cfg::kGroups - 1) {
                const float4* src = reinterpret_cast<const float4*>(xq + ((g + 1) * 2 + 1) * N_CTA + ch0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 x = src[k];
                    t[4 * k] = t[4 * k] + x.x; t[4 * k + 1] = t[4 * k + 1] + x.y; t[4 * k + 2] = t[4 * k + 2] + x.z; t[4 * k + 3] = t[4 * k + 3] + x.w;
                }
            }
            // phase B: centre column + neighbours, SE tail, activation, store
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                float v[16];
                tmem_ld16(cbase + N_CTA + ch0 + grp * 16, v);
                if (grp == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += t[grp * 16 + i];
                if (fused && valid) {
                    float4 sc4[4];
                    if (p.scale_smem_bytes > 0) {
                        const float4* sp = reinterpret_cast<const float4*>(scale_s + (size_t)bidx * N_CTA + ch0 + grp * 16);
#pragma unroll
                        for (int k = 0; k < 4; ++k) sc4[k] = sp[k];
                    } else {
                        const float4* sp = reinterpret_cast<const float4*>(p.se_scale + (size_t)bidx * p.cout + ch0 + grp * 16);
#pragma unroll
                        for (int k = 0; k < 4; ++k) sc4[k] = __ldg(sp + k);
                    }
                    const uint32_t rw[8] = {rv[2 * grp].x, rv[2 * grp].y, rv[2 * grp].z, rv[2 * grp].w,
                                            rv[2 * grp + 1].x, rv[2 * grp + 1].y, rv[2 * grp + 1].z, rv[2 * grp + 1].w};
                    const float scv[16] = {sc4[0].x, sc4[0].y, sc4[0].z, sc4[0].w, sc4[1].x, sc4[1].y, sc4[1].z, sc4[1].w,
                                           sc4[2].x, sc4[2].y, sc4[2].z, sc4[2].w, sc4[3].x, sc4[3].y, sc4[3].z, sc4[3].w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 r2 = unpack2<BF16>(rw[i]);
                        v[2 * i] = fmaf(v[2 * i], scv[2 * i], r2.x);
                        v[2 * i + 1] = fmaf(v[2 * i + 1], scv[2 * i + 1], r2.y);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * slope);
                if (optr != nullptr) {
                    uint16_t* dst = optr + (size_t)((ch0 + grp * 16) >> 3) * plane8;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 o;
                        o.x = valid ? pack2<BF16>(v[j * 8 + 0], v[j * 8 + 1]) : 0u;
                        o.y = valid ? pack2<BF16>(v[j * 8 + 2], v[j * 8 + 3]) : 0u;
                        o.z = valid ? pack2<BF16>(v[j * 8 + 4], v[j * 8 + 5]) : 0u;
                        o.w = valid ? pack2<BF16>(v[j * 8 + 6], v[j * 8 + 7]) : 0u;
                        *reinterpret_cast<uint4*>(dst + j * plane8) = o;
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
