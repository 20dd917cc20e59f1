// Micro-benchmark: cost of tensor-map TMA stores (cp.async.bulk.tensor.2d.global.shared::cta, SWIZZLE_128B) of fp32 boxes of
// 32 columns x R rows from shared memory into a 20480 x 20480 matrix, one issuing lane per CTA, as an epilogue of
// score_gemm_kernel would issue them.  Also checks the SWIZZLE_128B shared-memory pattern (16-byte chunk c of row r at
// r * 128 + ((c ^ (r & 7)) * 16)) by reading the matrix back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_store_bench tools/tma_store_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int ROWS, int FL>
__global__ void __launch_bounds__(128, 1) store_kernel(const __grid_constant__ CUtensorMap map, int n_col_blocks, int n_row_blocks, int ops_per_cta) {
    constexpr int in_flight = FL;
    extern __shared__ __align__(1024) uint8_t smem[];
    // fill `in_flight` boxes: element (r, c) of box b holds r * 1000 + c + b (swizzled)
    for (int i = threadIdx.x; i < in_flight * ROWS * 32; i += blockDim.x) {
        const int b = i / (ROWS * 32), r = (i / 32) % ROWS, c = i % 32;
        float* box = reinterpret_cast<float*>(smem + (size_t)b * ROWS * 128);
        box[r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3)] = (float)(r * 1000 + c + b);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long total = (long long)n_col_blocks * n_row_blocks;
        long long t = (long long)blockIdx.x * ops_per_cta;
        for (int i = 0; i < ops_per_cta; ++i, ++t) {
            const long long tt = t % total;
            const int cb = (int)(tt % n_col_blocks), rb = (int)(tt / n_col_blocks);
            const int b = i % in_flight;
            if (i >= in_flight) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(FL - 1) : "memory");   // the box about to be reused has been read
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(cb * 32), "r"(rb * ROWS),
                         "r"(smem_u32(smem + (size_t)b * ROWS * 128))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const int N = 20480;
    float* d;
    cudaMalloc(&d, (size_t)N * N * 4);
    cudaMemset(d, 0, (size_t)N * N * 4);
    PFN_cuTensorMapEncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int rows, int in_flight) {
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)N};
        const cuuint64_t gstride[1] = {(cuuint64_t)N * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
        const int ncb = N / 32, nrb = N / rows;
        const int ops = (int)((long long)ncb * nrb / 148);
        const size_t smem = (size_t)in_flight * rows * 128;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
#define LAUNCH(R, F) if (rows == R && in_flight == F) { cudaFuncSetAttribute(store_kernel<R, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); store_kernel<R, F><<<148, 128, smem>>>(map, ncb, nrb, ops); }
            LAUNCH(32, 1) LAUNCH(32, 2) LAUNCH(32, 4) LAUNCH(32, 8) LAUNCH(128, 1) LAUNCH(128, 2) LAUNCH(128, 4)
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double bytes = 148.0 * ops * rows * 128;
        printf("box 32 x %3d (%2d KB), %d in flight: %7.3f ms  %6.2f TB/s  %6.0f cycles per store at 1.9 GHz  (%s)\n", rows, rows * 128 / 1024, in_flight, ms,
               bytes / ms * 1e-9, ms * 1e-3 * 1.9e9 / ops, cudaGetErrorString(cudaGetLastError()));
    };
    for (int rows : {32, 128})
        for (int fl : {1, 2, 4, 8}) if (!(rows == 128 && fl == 8)) run(rows, fl);
    // pattern check on the last configuration (box 32 x 128, 4 boxes in flight): element (r, c) of the box at (rb, cb)
    float h[4];
    const int rr = 5, cc = 13;
    cudaMemcpy(h, d + (size_t)rr * N + cc, 4, cudaMemcpyDeviceToHost);
    printf("matrix[5][13] = %.0f (expected 5013 + box index 0..3)\n", h[0]);
    cudaMemcpy(h, d + (size_t)(128 + 77) * N + 32 * 3 + 31, 4, cudaMemcpyDeviceToHost);
    printf("matrix[205][127] = %.0f (expected 77031 + box index)\n", h[0]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
