// Micro-benchmark: achievable L2 -> shared-memory bandwidth of 1-D bulk copies (cp.async.bulk, mbarrier completion) from an
// L2-RESIDENT operand, as score_gemm_kernel / conv_umma_kernel stream their B tiles: one producer lane per CTA keeps
// `stages` copies of `stage_bytes` in flight; the CTAs walk the operand's 64 KB tiles from different starting points.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_stream_bench tools/l2_stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* src, size_t src_bytes, int stage_bytes, int stages, int copies_per_cta,
                                                        unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);      // [stages]
    uint8_t* buf = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // start offset: CTAs spread over the operand like the persistent CTAs of the GEMM
        size_t off = ((size_t)blockIdx.x * 9 * 65536) % src_bytes;
        int issued = 0, done = 0;
        unsigned long long acc = 0;
        for (; issued < stages && issued < copies_per_cta; ++issued) {
            mbar_expect_tx(&full[issued], stage_bytes);
            bulk_load(buf + (size_t)issued * stage_bytes, src + off, stage_bytes, &full[issued]);
            off += stage_bytes; if (off + stage_bytes > src_bytes) off = 0;
        }
        while (done < copies_per_cta) {
            const int s = done % stages;
            mbar_wait(&full[s], (done / stages) & 1);
            acc += *reinterpret_cast<volatile unsigned long long*>(buf + (size_t)s * stage_bytes);
            ++done;
            if (issued < copies_per_cta) {
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_load(buf + (size_t)s * stage_bytes, src + off, stage_bytes, &full[s]);
                off += stage_bytes; if (off + stage_bytes > src_bytes) off = 0;
                ++issued;
            }
        }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

int main() {
    const size_t src_bytes = 10ull << 20;                    // 20 000 x 256 fp16: the packed test side
    uint8_t* d; unsigned long long* sink;
    cudaMalloc(&d, src_bytes); cudaMalloc(&sink, 8);
    cudaMemset(d, 1, src_bytes);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grids[] = {148, 74};
    for (int g : grids)
        for (int stage_kb : {8, 16, 32, 64})
            for (int stages : {2, 3, 4, 7, 12}) {
                const int stage_bytes = stage_kb * 1024;
                if ((size_t)stage_bytes * stages + 1024 > 227 * 1024) continue;
                const int copies = (int)((size_t)64 * 166 * 1024 / stage_bytes);     // ~10.6 MB per CTA, as one CTA of the 20k x 20k GEMM
                stream_kernel<<<g, 128, 1024 + (size_t)stage_bytes * stages>>>(d, src_bytes, stage_bytes, stages, copies, sink);
                cudaEventRecord(e0);
                stream_kernel<<<g, 128, 1024 + (size_t)stage_bytes * stages>>>(d, src_bytes, stage_bytes, stages, copies, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)g * copies * stage_bytes;
                printf("grid %3d  stage %2d KB x %2d in flight (%3d KB): %7.3f ms  %6.2f TB/s  %5.1f B/clk/SM at 1.9 GHz\n", g, stage_kb, stages,
                       stage_kb * stages, ms, bytes / ms * 1e-9, bytes / g / (ms * 1e-3) / 1.9e9);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
