// Probe of tcgen05.shift.down (sm_100a): which TMEM rows / columns one instruction moves, in which direction, and what it
// costs.  Motivation: a 3x3 convolution whose three horizontal taps share one A slab (N = 3 * Cout) needs the three
// accumulator groups combined across +-1 accumulator ROWS; doing that with shuffles in the epilogue cost more than the
// fusion saved (profiles/r01c_fused_taps.txt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Isidekit_b200/csrc -Iinclude -o tools/shift_probe tools/shift_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace skb;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
    asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

// out[col][lane] = value read back; value written = lane * 256 + col
__global__ void probe(int lane_off, int col_off, int n_shift, int* out, long long* cycles) {
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<64>(&tmem_base);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_base;
    const uint32_t mine = base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 256 + c0 + i;
        tmem_st16(mine + c0, r);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int i = 0; i < n_shift; ++i) tmem_shift_down(base + ((uint32_t)lane_off << 16) + col_off);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) { t1 = clock64(); *cycles = t1 - t0; }
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(mine + c0, v);
        for (int i = 0; i < 16; ++i) out[(c0 + i) * 128 + threadIdx.x] = __float_as_int(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(base);
}

// throughput: n_shift shifts spread over `chunks` 8-column chunks, one commit at the end
__global__ void timing(int n_shift, int chunks, long long* cycles) {
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<64>(&tmem_base);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < n_shift; ++i) tmem_shift_down(tmem_base + (i % chunks) * 8);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        *cycles = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
}

int main() {
    int* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 64 * 128 * sizeof(int));
    cudaMalloc(&d_cyc, sizeof(long long));
    static int h[64 * 128];
    const int cfgs[][3] = {{0, 0, 1}, {0, 0, 2}, {0, 8, 1}, {32, 16, 1}, {64, 4, 1}, {96, 24, 1}};
    for (auto& c : cfgs) {
        probe<<<1, 128>>>(c[0], c[1], c[2], d_out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lane_off=%d col_off=%d n=%d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
        long long cyc;
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        printf("== shift at lane %d col %d, %d time(s), %lld cycles issue->complete\n", c[0], c[1], c[2], cyc);
        // which columns changed, and for a changed column the source lane of every lane
        int first = -1, last = -1;
        for (int col = 0; col < 64; ++col) {
            bool changed = false;
            for (int l = 0; l < 128; ++l) if (h[col * 128 + l] != l * 256 + col) changed = true;
            if (changed) { if (first < 0) first = col; last = col; }
        }
        printf("   changed columns: %d .. %d\n", first, last);
        if (first >= 0) {
            printf("   column %d: lane <- source lane (value col): ", first);
            for (int l = 0; l < 128; ++l) {
                const int v = h[first * 128 + l];
                if (v / 256 != l || v % 256 != first) printf("%d<-%d(c%d) ", l, v / 256, v % 256);
            }
            printf("\n");
        }
    }
    for (int chunks : {1, 4, 8}) for (int n : {8, 64, 256}) {
        timing<<<1, 128>>>(n, chunks, d_cyc);
        cudaDeviceSynchronize();
        long long cyc;
        cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        printf("timing: %3d shifts over %d chunk(s): %6lld cycles  (%.1f per shift)\n", n, chunks, cyc, (double)cyc / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
