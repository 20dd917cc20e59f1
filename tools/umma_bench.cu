// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16, SS mode) as a function of N and of the
// shared-memory operand layout (no-swizzle interleaved vs 32/64/128-byte swizzle).  Operand contents are
// irrelevant (timing only).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Isidekit_b200/csrc -Iinclude
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SKB_OK 0
#define SKB_ERR_CUDA (-2)
#include "common.cuh"
namespace skb { void set_last_error(const char*, int, const char*) {} }
using namespace skb;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// mode 0: no swizzle (LBO = rows*16, SBO = 128); 1: SW32 (layout 6), 2: SW64 (4), 3: SW128 (2): SBO = 8 rows * row bytes
__global__ void __launch_bounds__(128) bench(int N, int mode, int iters, int a_shift_rows, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(sm), b = smem_u32(sm + 96 * 1024);
        uint32_t layout = 0, lbo_a = 512 * 16, lbo_b = (uint32_t)N * 16, sbo = 128;
        if (mode == 1) { layout = 6; lbo_a = lbo_b = 16; sbo = 8 * 32; }
        if (mode == 2) { layout = 4; lbo_a = lbo_b = 16; sbo = 8 * 64; }
        if (mode == 3) { layout = 2; lbo_a = lbo_b = 16; sbo = 8 * 128; }
        const uint32_t row_bytes = mode == 0 ? 16 : (mode == 1 ? 32 : (mode == 2 ? 64 : 128));
        const uint32_t idesc = umma_idesc_f16(128, N, false);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t ad = make_desc(a + (uint32_t)(a_shift_rows * (k % 3)) * row_bytes + (mode == 0 ? (k & 3) * 2 * lbo_a : (k & 3) * 32 % row_bytes), lbo_a, sbo, layout);
                const uint64_t bd = make_desc(b + (mode == 0 ? (k & 3) * 2 * lbo_b : (k & 3) * 32 % row_bytes), lbo_b, sbo, layout);
                umma_f16(tm + (k & 1) * 256, ad, bd, idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

// TMEM -> register bandwidth: `nw` warps (warp w reads lane quadrant w % 4) each issue `iters` tcgen05.ld 32x32b.x32
// (32 lanes x 32 columns x 4 B = 4 KB per instruction) back to back.
__global__ void __launch_bounds__(256) ldtm_bench(int iters, long long* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((i * 32) & 511), v);
#pragma unroll
        for (int k = 0; k < 32; ++k) acc += v[k];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x % 32 == 0) out[blockIdx.x * 8 + warp] = (t1 - t0) + (acc == 12345.f ? 1 : 0);
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

int main() {
    {
        long long* dl; cudaMalloc(&dl, 148 * 8 * 8);
        for (int nw : {1, 4, 8}) {
            const int iters = 4000;
            ldtm_bench<<<148, nw * 32>>>(iters, dl);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ldtm nw=%d: %s\n", nw, cudaGetErrorString(e)); return 1; }
            long long h[148 * 8]; cudaMemcpy(h, dl, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0; for (int i = 0; i < 148; ++i) for (int w = 0; w < nw; ++w) mx = h[i * 8 + w] > mx ? h[i * 8 + w] : mx;
            printf("tcgen05.ld 32x32b.x32 + wait, %d warps/SM: %.1f cycles per instruction per warp -> %.1f B/cycle/SM\n", nw,
                   mx / iters, nw * 4096.0 * iters / mx);
        }
    }
    long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[4] = {"noswz", "sw32", "sw64", "sw128"};
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {32, 64, 96, 128, 192, 256})
            for (int shift : {0, 1}) {
                const int iters = 2000;
                bench<<<148, 128, 200 * 1024>>>(N, mode, iters, shift, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s N=%d shift=%d: %s\n", names[mode], N, shift, cudaGetErrorString(e)); return 1; }
                long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
                printf("%-6s N=%3d rowshift=%d : %.1f cycles/MMA (ideal %d)\n", names[mode], N, shift, avg / (iters * 8.0), N / 2);
            }
    return 0;
}
