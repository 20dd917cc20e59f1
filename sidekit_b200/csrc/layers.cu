// Memory-bound and small kernels around the tcgen05 convolutions: stem conv, squeeze-excitation,
// residual merge, statistics / attentive pooling, the embedding head, and an fp32 SIMT GEMM for the
// small dense layers.  All operate on the chunk-plane activation layout described in conv_umma.cuh.
#include "sidekit_b200.h"
#include "common.cuh"
#include "layers.cuh"

namespace skb {

// ----------------------------------------------------------------------------- stem: 3x3 conv 1->32 + BN + ReLU
// sidekit/nnet/res_net.py:549 (relu(bn1(conv1(x)))) on the (B,1,T,F) view of the features.
// One thread per level-1 pixel; the normalised features are frame-major so a warp reads contiguous
// mel bins.  fp32 math on CUDA cores (0.1 % of the trunk's FLOPs), 16-bit chunk-plane output.
template <bool BF16>
__global__ void stem_kernel(const float* __restrict__ feats, const long long* __restrict__ feat_off,
                            const int* __restrict__ n_frames, const float* __restrict__ w /*[32][9]*/,
                            const float* __restrict__ bias /*[32]*/, uint16_t* __restrict__ out, long long out_plane,
                            int G, int p_end, int Wp, int W, const int* __restrict__ row_b,
                            const int* __restrict__ row_h) {
    __shared__ float sw[32 * 9 + 32];
    for (int i = threadIdx.x; i < 32 * 9 + 32; i += blockDim.x) sw[i] = i < 288 ? w[i] : bias[i - 288];
    __syncthreads();
    const int pix = G + blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p_end) return;
    const int rel = pix - G;
    const int row = rel / Wp, f = rel - row * Wp;
    const int b = row_b[row], t = row_h[row];
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    const bool valid = (b >= 0) && (f < W);
    if (valid) {
        const int T = n_frames[b];
        const float* fb = feats + (size_t)feat_off[b] * W;
        float x[9];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt)
#pragma unroll
            for (int df = 0; df < 3; ++df) {
                const int tt = t + dt - 1, ff = f + df - 1;
                x[dt * 3 + df] = (tt >= 0 && tt < T && ff >= 0 && ff < W) ? __ldg(fb + (size_t)tt * W + ff) : 0.f;
            }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float a = sw[288 + c];
#pragma unroll
            for (int k = 0; k < 9; ++k) a = fmaf(sw[c * 9 + k], x[k], a);
            acc[c] = fmaxf(a, 0.f);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack2<BF16>(acc[j * 8 + 0], acc[j * 8 + 1]);
        o.y = pack2<BF16>(acc[j * 8 + 2], acc[j * 8 + 3]);
        o.z = pack2<BF16>(acc[j * 8 + 4], acc[j * 8 + 5]);
        o.w = pack2<BF16>(acc[j * 8 + 6], acc[j * 8 + 7]);
        *reinterpret_cast<uint4*>(out + ((size_t)j * out_plane + pix) * 8) = o;
    }
}

int launch_stem(bool bf16, const float* feats, const long long* feat_off, const int* n_frames, const float* w,
                const float* bias, uint16_t* out, long long out_plane, int G, int p_end, int Wp, int W,
                const int* row_b, const int* row_h, cudaStream_t st) {
    const int n = p_end - G;
    const int threads = 128;
    const int blocks = (n + threads - 1) / threads;
    if (bf16)
        stem_kernel<true><<<blocks, threads, 0, st>>>(feats, feat_off, n_frames, w, bias, out, out_plane, G, p_end, Wp, W, row_b, row_h);
    else
        stem_kernel<false><<<blocks, threads, 0, st>>>(feats, feat_off, n_frames, w, bias, out, out_plane, G, p_end, Wp, W, row_b, row_h);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- SE squeeze, computed BEFORE conv2 runs
// The SE layer needs mean_{h,w}(bn2(conv2(y1))) per (utterance, channel) (sidekit/nnet/res_net.py:272-281, :316-317).
// conv2 is linear, so that mean follows from sums of its INPUT y1:
//     sum_{h,w} y2[c] = N*b2[c] + sum_{tap,ci} W2[c][ci][tap] * S_tap[ci],
//     S_(dr,ds)[ci] = Total[ci] - (excluded border row) - (excluded border column) + (their corner)
// (a tap shifted by (dr, ds) sees every pixel except one border row / column; the zero padding contributes nothing).
// Knowing the scales up front lets conv2's epilogue apply scale + residual + ReLU directly: y2 is never stored and
// the separate "scale, add, ReLU" pass disappears.
//
// plane_sum_kernel: Total[b][ci] over the valid pixels, in 2^-24 fixed point.  16-bit activations times 2^24 are
// exact integers and integer addition is associative, so the sums -- and through them every embedding -- are
// bit-identical whatever the packing, grid shape or atomic order.
template <bool BF16>
__global__ void __launch_bounds__(256) plane_sum_kernel(const uint16_t* __restrict__ act, long long plane, int G, int p_end,
                                                        const int* __restrict__ pix_b, int C, unsigned long long* __restrict__ sums) {
    constexpr int PX = 8;                                   // pixels per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.y;
    const int base = G + (blockIdx.x * 8 + warp) * 32 * PX;
    long long t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = 0ll;
    int b_acc = -1;
    bool mixed = false;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int pix = base + k * 32 + lane;
        int b = -1;
        if (pix < p_end) b = __ldg(pix_b + (pix - G));
        if (b >= 0) {
            if (b_acc >= 0 && b != b_acc) {               // utterance boundary inside this lane's pixels (rare): flush
                for (int e = 0; e < 8; ++e) {
                    atomicAdd(sums + (size_t)b_acc * C + j * 8 + e, (unsigned long long)t[e]);
                    t[e] = 0ll;
                }
                mixed = true;
            }
            b_acc = b;
            const uint4 a = *reinterpret_cast<const uint4*>(act + ((size_t)j * plane + pix) * 8);
            const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 v = unpack2<BF16>(u[e]);
                t[2 * e] += __float2ll_rn(v.x * 16777216.f);
                t[2 * e + 1] += __float2ll_rn(v.y * 16777216.f);
            }
        }
    }
    const unsigned has = __ballot_sync(0xffffffffu, b_acc >= 0);
    if (has == 0u) return;
    const int b0 = __shfl_sync(0xffffffffu, b_acc, __ffs(has) - 1);
    if (__all_sync(0xffffffffu, (b_acc < 0 || b_acc == b0) && !mixed)) {
        // exact warp sum of 64-bit values with the 32-bit hardware reduction: 20-bit low limb + signed high limb
        long long mine = 0ll;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int slo = __reduce_add_sync(0xffffffffu, (int)(t[e] & 0xFFFFFll));
            const int shi = __reduce_add_sync(0xffffffffu, (int)(t[e] >> 20));
            if (lane == e) mine = ((long long)shi << 20) + (long long)slo;
        }
        if (lane < 8) atomicAdd(sums + (size_t)b0 * C + j * 8 + lane, (unsigned long long)mine);
    } else if (b_acc >= 0) {
        for (int e = 0; e < 8; ++e) atomicAdd(sums + (size_t)b_acc * C + j * 8 + e, (unsigned long long)t[e]);
    }
}

int launch_plane_sum(bool bf16, const uint16_t* act, long long plane, int G, int p_end, const int* pix_b, int C,
                     unsigned long long* sums, cudaStream_t st) {
    const int n = p_end - G;
    dim3 grid((n + 2047) / 2048, C / 8);
    if (bf16)
        plane_sum_kernel<true><<<grid, 256, 0, st>>>(act, plane, G, p_end, pix_b, C, sums);
    else
        plane_sum_kernel<false><<<grid, 256, 0, st>>>(act, plane, G, p_end, pix_b, C, sums);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// se_border_kernel: sums of the first / last row and column of y1 and its four corner pixels, per (utterance, channel).
// One CTA per (utterance, 8-channel chunk); threads stride over the border pixels with 16-byte loads; the block
// reduction runs in a fixed order, so the result is deterministic.  brd layout: [B][8 kinds][C].
template <bool BF16>
__global__ void __launch_bounds__(128) se_border_kernel(const uint16_t* __restrict__ y1, long long plane, int G, int Wp, int W,
                                                        const int* __restrict__ utt_row0, const int* __restrict__ utt_count, int C,
                                                        float* __restrict__ brd) {
    __shared__ float part[128][33];
    const int b = blockIdx.x, j = blockIdx.y;
    const int H = utt_count[b] / W;
    const uint16_t* base = y1 + ((size_t)j * plane + G + (size_t)utt_row0[b] * Wp) * 8;     // pixel (0, 0) of chunk j
    float acc[32];                                  // [row0 | rowL | col0 | colL][8 channels]
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    auto add8 = [&](const uint16_t* p, int o) {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 v = unpack2<BF16>(u[e]);
            acc[o + 2 * e] += v.x;
            acc[o + 2 * e + 1] += v.y;
        }
    };
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        add8(base + (size_t)w * 8, 0);
        add8(base + ((size_t)(H - 1) * Wp + w) * 8, 8);
    }
    for (int hh = threadIdx.x; hh < H; hh += blockDim.x) {
        add8(base + (size_t)hh * Wp * 8, 16);
        add8(base + ((size_t)hh * Wp + W - 1) * 8, 24);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) part[threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 32) {                         // thread i sums quantity i over the 128 partials, in order
        double a = 0.0;
        for (int t = 0; t < 128; ++t) a += (double)part[t][threadIdx.x];
        const int kind = threadIdx.x >> 3, e = threadIdx.x & 7;
        brd[((size_t)b * 8 + kind) * C + j * 8 + e] = (float)a;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 64) {    // corners: (0,0) (0,W-1) (H-1,0) (H-1,W-1)
        const int k = (threadIdx.x - 32) >> 3, e = threadIdx.x & 7;
        const int hh = (k >> 1) ? H - 1 : 0, ww = (k & 1) ? W - 1 : 0;
        const uint16_t u = base[((size_t)hh * Wp + ww) * 8 + e];
        brd[((size_t)b * 8 + 4 + k) * C + j * 8 + e] = unpack2<BF16>((uint32_t)u).x;
    }
}

// se_mean_partial_kernel: the mean of conv2's output through the (16-bit-rounded, BN-folded) conv2 weights, as a small
// GEMM  partial[ks][b][co] = sum_{i in K-slice ks} W2t[i][co] * S[b][i],  i = ci * 9 + tap,  S = the nine shifted sums.
// Grid = (K / 64 slices, utterance groups of 16): enough CTAs to cover the L2 latency of the weight stream; every
// weight row is read once per utterance group.  Fixed summation order everywhere -> deterministic.
constexpr int kSeRows = 64, kSeUtt = 16;
__global__ void __launch_bounds__(256) se_mean_partial_kernel(const unsigned long long* __restrict__ sums, const float* __restrict__ brd,
                                                              int B, int Cin, int Cout, const float* __restrict__ w2t,
                                                              float* __restrict__ partial) {
    __shared__ float sS[kSeUtt][kSeRows];
    const int K = 9 * Cin;
    const int k0 = blockIdx.x * kSeRows, b0 = blockIdx.y * kSeUtt;
    const int nr = min(kSeRows, K - k0), nu = min(kSeUtt, B - b0);
    for (int idx = threadIdx.x; idx < kSeUtt * kSeRows; idx += blockDim.x) {
        const int u = idx / kSeRows, r = idx - u * kSeRows;
        float v = 0.f;
        if (u < nu && r < nr) {
            const int i = k0 + r, c = i / 9, tap = i - c * 9;
            const int dr = tap / 3 - 1, ds = tap % 3 - 1;
            const int b = b0 + u;
            const float* bb = brd + (size_t)b * 8 * Cin + c;
            v = (float)((double)(long long)sums[(size_t)b * Cin + c] * (1.0 / 16777216.0));
            if (dr < 0) v -= bb[Cin];                 // shifted up: the last row is never read
            if (dr > 0) v -= bb[0];
            if (ds < 0) v -= bb[3 * Cin];
            if (ds > 0) v -= bb[2 * Cin];
            if (dr < 0 && ds < 0) v += bb[7 * Cin];
            if (dr < 0 && ds > 0) v += bb[6 * Cin];
            if (dr > 0 && ds < 0) v += bb[5 * Cin];
            if (dr > 0 && ds > 0) v += bb[4 * Cin];
        }
        sS[u][r] = v;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
        float acc[kSeUtt];
#pragma unroll
        for (int u = 0; u < kSeUtt; ++u) acc[u] = 0.f;
        const float* wp = w2t + (size_t)k0 * Cout + co;
#pragma unroll 8
        for (int r = 0; r < nr; ++r) {
            const float w = __ldg(wp + (size_t)r * Cout);
#pragma unroll
            for (int u = 0; u < kSeUtt; ++u) acc[u] = fmaf(w, sS[u][r], acc[u]);
        }
        for (int u = 0; u < nu; ++u) partial[((size_t)blockIdx.x * B + b0 + u) * Cout + co] = acc[u];
    }
}

// se_fc_kernel: mean = b2 + (1/N) * sum of the K-slice partials (in slice order); scale = sigmoid(W2 relu(W1 mean))
// (sidekit/nnet/res_net.py:272-281).  One CTA per utterance; re-zeroes the fixed-point channel sums for the next block.
__global__ void __launch_bounds__(256) se_fc_kernel(unsigned long long* __restrict__ sums, const float* __restrict__ partial, int n_slices,
                                                    const int* __restrict__ utt_count, int B, int Cin, int Cout,
                                                    const float* __restrict__ b2, const float* __restrict__ fc1 /*[Cout/16][Cout]*/,
                                                    const float* __restrict__ fc2 /*[Cout][Cout/16]*/, float* __restrict__ scale) {
    __shared__ float mean[256];
    __shared__ float hid[16];
    const int b = blockIdx.x;
    const float inv_n = 1.f / (float)utt_count[b];
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) sums[(size_t)b * Cin + c] = 0ull;
    for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < n_slices; ++k) a += partial[((size_t)k * B + b) * Cout + co];
        mean[co] = fmaf(a, inv_n, b2[co]);
    }
    __syncthreads();
    const int R = Cout / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int jj = warp; jj < R; jj += blockDim.x >> 5) {
        float a = 0.f;
        for (int c = lane; c < Cout; c += 32) a = fmaf(fc1[jj * Cout + c], mean[c], a);
        a = warp_sum(a);
        if (lane == 0) hid[jj] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
        float a = 0.f;
        for (int jj = 0; jj < R; ++jj) a = fmaf(fc2[c * R + jj], hid[jj], a);
        scale[(size_t)b * Cout + c] = 1.f / (1.f + __expf(-a));
    }
}

int launch_se_scale(bool bf16, unsigned long long* sums, const uint16_t* y1, long long plane, int G, int Wp, int W,
                    const int* utt_row0, const int* utt_count, int B, int Cin, int Cout, const float* w2t, const float* b2,
                    const float* fc1, const float* fc2, float* brd_ws, float* scale, cudaStream_t st) {
    // brd_ws: [B][8][Cin] border sums, followed by [n_slices][B][Cout] partial means
    dim3 grid(B, Cin / 8);
    if (bf16)
        se_border_kernel<true><<<grid, 128, 0, st>>>(y1, plane, G, Wp, W, utt_row0, utt_count, Cin, brd_ws);
    else
        se_border_kernel<false><<<grid, 128, 0, st>>>(y1, plane, G, Wp, W, utt_row0, utt_count, Cin, brd_ws);
    const int n_slices = (9 * Cin + kSeRows - 1) / kSeRows;
    float* partial = brd_ws + (size_t)B * 8 * Cin;
    dim3 g2(n_slices, (B + kSeUtt - 1) / kSeUtt);
    se_mean_partial_kernel<<<g2, Cout < 256 ? (Cout < 32 ? 32 : Cout) : 256, 0, st>>>(sums, brd_ws, B, Cin, Cout, w2t, partial);
    se_fc_kernel<<<B, 256, 0, st>>>(sums, partial, n_slices, utt_count, B, Cin, Cout, b2, fc1, fc2, scale);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- planes -> dense fp32 frames
// X[frame][c * W + f] = act[c/8][G + (row0_b + t) * Wp + f][c%8]: the (B, C*F, T) view of
// sidekit/nnet/pooling.py:160-163, stored frame-major.  Thread per (frame, f, chunk).
template <bool BF16>
__global__ void gather_frames_kernel(const uint16_t* __restrict__ act, long long plane, int C, int W, int Wp, int G,
                                     const int* __restrict__ frame_row, int n_frames, float* __restrict__ X) {
    const int chunks = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_frames * W * chunks;
    if (idx >= total) return;
    const int f = (int)(idx % W);
    const int j = (int)((idx / W) % chunks);
    const int fr = (int)(idx / ((long long)W * chunks));
    const long long pix = (long long)G + (long long)frame_row[fr] * Wp + f;
    const uint4 a = *reinterpret_cast<const uint4*>(act + ((size_t)j * plane + pix) * 8);
    const uint32_t u[4] = {a.x, a.y, a.z, a.w};
    float* dst = X + (size_t)fr * C * W;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 v = unpack2<BF16>(u[e]);
        dst[(j * 8 + 2 * e) * W + f] = v.x;
        dst[(j * 8 + 2 * e + 1) * W + f] = v.y;
    }
}

int launch_gather_frames(bool bf16, const uint16_t* act, long long plane, int C, int W, int Wp, int G,
                         const int* frame_row, int n_frames, float* X, cudaStream_t st) {
    const long long total = (long long)n_frames * W * (C / 8);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    if (blocks == 0) return SKB_OK;
    if (bf16)
        gather_frames_kernel<true><<<blocks, threads, 0, st>>>(act, plane, C, W, Wp, G, frame_row, n_frames, X);
    else
        gather_frames_kernel<false><<<blocks, threads, 0, st>>>(act, plane, C, W, Wp, G, frame_row, n_frames, X);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- mean / unbiased std over time
// MeanStdPooling (sidekit/nnet/pooling.py:55-70): out[b] = [mean_t x ; std_t x (ddof=1)].  Optional per-channel
// affine (s, t) applied as mean' = s*mean + t, std' = |s|*std (folds the TDNN's last BatchNorm).
__global__ void meanstd_kernel(const float* __restrict__ X, const long long* __restrict__ frame_off,
                               const int* __restrict__ n_fr, int D, const float* __restrict__ aff_s,
                               const float* __restrict__ aff_t, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int T = n_fr[b];
    const float* x = X + (size_t)frame_off[b] * D + d;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += x[(size_t)t * D];
    const float mean = s / (float)T;
    float v = 0.f;
    for (int t = 0; t < T; ++t) {
        const float dlt = x[(size_t)t * D] - mean;
        v = fmaf(dlt, dlt, v);
    }
    float sd = sqrtf(v / (float)(T - 1));
    float mu = mean;
    if (aff_s) {
        mu = aff_s[d] * mean + aff_t[d];
        sd = fabsf(aff_s[d]) * sd;
    }
    out[(size_t)b * 2 * D + d] = mu;
    out[(size_t)b * 2 * D + D + d] = sd;
}

int launch_meanstd(const float* X, const long long* frame_off, const int* n_fr, int B, int D, const float* aff_s,
                   const float* aff_t, float* out, cudaStream_t st) {
    dim3 grid((D + 127) / 128, B);
    meanstd_kernel<<<grid, 128, 0, st>>>(X, frame_off, n_fr, D, aff_s, aff_t, out);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- attention activation
// h = tanh(BN(relu(h + hb[utt])))  (sidekit/nnet/pooling.py:134-139: Conv1d -> ReLU -> BatchNorm1d -> Tanh);
// hb is the time-constant global-context part of the first Conv1d (+ its bias), one row per utterance.
__global__ void att_act_kernel(float* __restrict__ h, const float* __restrict__ hb, const int* __restrict__ frame_utt,
                               const float* __restrict__ bn_s, const float* __restrict__ bn_t, int n_frames, int A) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_frames * A) return;
    const int fr = (int)(idx / A), a = (int)(idx % A);
    const float v = fmaxf(h[idx] + hb[(size_t)frame_utt[fr] * A + a], 0.f);
    h[idx] = tanhf(fmaf(v, bn_s[a], bn_t[a]));
}

int launch_att_act(float* h, const float* hb, const int* frame_utt, const float* bn_s, const float* bn_t, int n_frames,
                   int A, cudaStream_t st) {
    const long long total = (long long)n_frames * A;
    if (total == 0) return SKB_OK;
    att_act_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(h, hb, frame_utt, bn_s, bn_t, n_frames, A);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- softmax over time + weighted stats
// w = softmax_t(logits); mu = sum x w; rh = sqrt(clamp(sum x^2 w - mu^2, 1e-9))  (pooling.py:167-169).
__global__ void softmax_pool_kernel(const float* __restrict__ X, const float* __restrict__ logit,
                                    const long long* __restrict__ frame_off, const int* __restrict__ n_fr, int D,
                                    float* __restrict__ out) {
    const int b = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const int T = n_fr[b];
    const size_t base = (size_t)frame_off[b] * D + d;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, logit[base + (size_t)t * D]);
    float se = 0.f, sx = 0.f, sxx = 0.f;
    for (int t = 0; t < T; ++t) {
        const float e = __expf(logit[base + (size_t)t * D] - m);
        const float x = X[base + (size_t)t * D];
        se += e;
        sx = fmaf(x, e, sx);
        sxx = fmaf(x * x, e, sxx);
    }
    const float mu = sx / se;
    const float var = sxx / se - mu * mu;
    out[(size_t)b * 2 * D + d] = mu;
    out[(size_t)b * 2 * D + D + d] = sqrtf(fmaxf(var, 1e-9f));
}

int launch_softmax_pool(const float* X, const float* logit, const long long* frame_off, const int* n_fr, int B, int D,
                        float* out, cudaStream_t st) {
    dim3 grid((D + 127) / 128, B);
    softmax_pool_kernel<<<grid, 128, 0, st>>>(X, logit, frame_off, n_fr, D, out);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- embedding head tail
// x -> optional per-feature affine (folded BatchNorm1d) -> l2_norm (loss.py:91-100, no eps)
// -> F.normalize (eps 1e-12) (xvector.py:893-903).  emb_pre = after l2_norm (input of the margin head),
// emb = the returned embedding.
__global__ void head_norm_kernel(const float* __restrict__ x, const float* __restrict__ aff_s, const float* __restrict__ aff_t,
                                 int E, int norm_embedding, float* __restrict__ emb_pre, float* __restrict__ emb) {
    extern __shared__ float hs[];
    __shared__ float red[32];
    const int b = blockIdx.x;
    float ss = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        float v = x[(size_t)b * E + i];
        if (aff_s) v = fmaf(v, aff_s[i], aff_t[i]);
        hs[i] = v;
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float nrm = sqrtf(red[0]);
    // l2_norm divides by the norm (no eps); F.normalize then divides by max(norm', 1e-12)
    const float s1 = norm_embedding ? 1.f / nrm : 1.f;
    const float n2 = norm_embedding ? (nrm * s1) : nrm;       // norm after the first step (== 1 up to rounding)
    const float s2 = 1.f / fmaxf(n2, 1e-12f);
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        const float v = hs[i] * s1;
        emb_pre[(size_t)b * E + i] = v;
        emb[(size_t)b * E + i] = v * s2;
    }
}

int launch_head_norm(const float* x, const float* aff_s, const float* aff_t, int B, int E, int norm_embedding,
                     float* emb_pre, float* emb, cudaStream_t st) {
    head_norm_kernel<<<B, 256, E * sizeof(float), st>>>(x, aff_s, aff_t, E, norm_embedding, emb_pre, emb);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- fp32 SIMT GEMM (NT)
// C[m][n] = alpha * sum_k A[m][k] * B[n][k] + bias[n];  64x64 tile, 16-deep K slices, 4x4 register micro-tiles.
// Used for the small dense layers (attention projections, embedding and margin heads), which are
// < 0.5 % of the network's FLOPs and need fp32 inputs for the softmax logits.
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, const float* __restrict__ bias, int M, int N,
                                                       int K, int lda, int ldb, int ldc, float alpha) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = threadIdx.x >> 2;          // 0..63 row of the tile
    const int lk = (threadIdx.x & 3) * 4;     // 0,4,8,12
    for (int k0 = 0; k0 < K; k0 += 16) {
        {
            const int m = m0 + lr, n = n0 + lr;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + lk + e;
                As[lk + e][lr] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
                Bs[lk + e][lr] = (n < N && k < K) ? B[(size_t)n * ldb + k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[(size_t)m * ldc + n] = alpha * acc[i][j] + (bias ? bias[n] : 0.f);
        }
    }
}

int launch_sgemm_nt(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                    int ldc, float alpha, cudaStream_t st) {
    if (M == 0 || N == 0) return SKB_OK;
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    sgemm_nt_kernel<<<grid, 256, 0, st>>>(A, B, C, bias, M, N, K, lda, ldb, ldc, alpha);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

// ----------------------------------------------------------------------------- fp32 frame-major -> 16-bit chunk planes
// Used by the TDNN path: (frames, C) features -> act[c/8][G + frame][8] (Wp == 1 geometry), zero padded channels.
template <bool BF16>
__global__ void pack_frames_kernel(const float* __restrict__ X, int C_src, int C_dst, int n_rows, const int* __restrict__ row_src,
                                   uint16_t* __restrict__ out, long long plane, int G) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = C_dst >> 3;
    if (idx >= (long long)n_rows * chunks) return;
    const int row = (int)(idx % n_rows), j = (int)(idx / n_rows);
    const int src = row_src[row];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = j * 8 + e;
        v[e] = (src >= 0 && c < C_src) ? X[(size_t)src * C_src + c] : 0.f;
    }
    uint4 o;
    o.x = pack2<BF16>(v[0], v[1]); o.y = pack2<BF16>(v[2], v[3]); o.z = pack2<BF16>(v[4], v[5]); o.w = pack2<BF16>(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + ((size_t)j * plane + G + row) * 8) = o;
}

int launch_pack_frames(bool bf16, const float* X, int C_src, int C_dst, int n_rows, const int* row_src, uint16_t* out,
                       long long plane, int G, cudaStream_t st) {
    const long long total = (long long)n_rows * (C_dst / 8);
    if (total == 0) return SKB_OK;
    const int blocks = (int)((total + 255) / 256);
    if (bf16)
        pack_frames_kernel<true><<<blocks, 256, 0, st>>>(X, C_src, C_dst, n_rows, row_src, out, plane, G);
    else
        pack_frames_kernel<false><<<blocks, 256, 0, st>>>(X, C_src, C_dst, n_rows, row_src, out, plane, G);
    SKB_CUDA_CHECK(cudaGetLastError());
    return SKB_OK;
}

}  // namespace skb
