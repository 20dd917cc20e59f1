// Micro-benchmark: achievable HBM WRITE bandwidth of a (rows x cols) fp32 score matrix as a function of the store
// pattern -- a linear stream vs tile-shaped writers like the epilogue of score_gemm_kernel (each warp instruction
// stores `seg` contiguous bytes of `32*16/seg` different rows).  Gives the real ceiling of the HBM-write-bound
// scoring kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/wr_bench tools/wr_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// linear: every thread writes float4s, grid-stride
__global__ void linear_kernel(float4* out, size_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

// tiles: the matrix is cut into (TR x TC) tiles; CTA b owns the contiguous run of tiles [b*T/G, (b+1)*T/G) in
// panel-major order (all column tiles of a row panel, then the next panel).  Inside a tile, warp w owns rows
// [w*TR/nw, ...) and writes them as full-row segments: lanes cover (TC*4/16) 16-byte pieces of one row, remaining lanes
// the following rows.
__global__ void tile_kernel(float* out, int rows, int cols, int TR, int TC, int col_major_walk) {
    const int n_ct = cols / TC, n_rt = rows / TR;
    const long long tiles = (long long)n_ct * n_rt;
    const long long t0 = tiles * blockIdx.x / gridDim.x, t1 = tiles * (blockIdx.x + 1) / gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int lanes_per_row = TC / 4 < 32 ? TC / 4 : 32;      // 16-byte pieces of one row covered per instruction
    const int rows_per_instr = 32 / lanes_per_row;
    const int pieces = TC / 4 / lanes_per_row;                // instructions per row group
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long t = t0; t < t1; ++t) {
        int rt, ct;
        if (col_major_walk) { ct = (int)(t / n_rt); rt = (int)(t % n_rt); }
        else { rt = (int)(t / n_ct); ct = (int)(t % n_ct); }
        const int rpw = TR / nw;
        for (int r = 0; r < rpw; r += rows_per_instr) {
            const int row = rt * TR + warp * rpw + r + lane / lanes_per_row;
            float* o = out + (size_t)row * cols + (size_t)ct * TC + (lane % lanes_per_row) * 4;
            for (int pc = 0; pc < pieces; ++pc) *reinterpret_cast<float4*>(o + pc * lanes_per_row * 4) = v;
        }
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    const int rows = 20480, cols = 20480;
    const size_t bytes = (size_t)rows * cols * 4;
    float* d;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms) { printf("%-58s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms * 1e-6); };
    for (int rep = 0; rep < 2; ++rep) {
        for (int g : {148 * 2, 148 * 8}) {
            cudaEventRecord(e0);
            for (int i = 0; i < 5; ++i) linear_kernel<<<g, 1024>>>((float4*)d, bytes / 16);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            char nm[128]; snprintf(nm, sizeof(nm), "linear float4 grid=%d x 1024", g);
            report(nm, time_ms(e0, e1) / 5);
        }
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) cudaMemsetAsync(d, 0, bytes);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        report("cudaMemsetAsync", time_ms(e0, e1) / 5);
        struct Cfg { int TR, TC, threads, ctas_per_sm, cm; };
        const Cfg cfgs[] = {{128, 128, 256, 1, 0}, {128, 128, 512, 1, 0}, {128, 128, 256, 2, 0}, {128, 128, 256, 4, 0},
                            {128, 256, 256, 1, 0}, {128, 256, 512, 1, 0}, {128, 512, 256, 1, 0}, {256, 128, 256, 1, 0},
                            {64, 512, 256, 1, 0},  {32, 1024, 256, 1, 0}, {128, 32, 256, 1, 0}, {128, 64, 256, 1, 0}, {128, 128, 256, 1, 1}, {128, 256, 256, 1, 1}};
        for (const Cfg& c : cfgs) {
            cudaEventRecord(e0);
            for (int i = 0; i < 5; ++i) tile_kernel<<<148 * c.ctas_per_sm, c.threads>>>(d, rows, cols, c.TR, c.TC, c.cm);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            char nm[128];
            snprintf(nm, sizeof(nm), "tile %3dx%-4d threads=%d ctas/sm=%d %s", c.TR, c.TC, c.threads, c.ctas_per_sm, c.cm ? "column-major walk" : "panel-major walk");
            report(nm, time_ms(e0, e1) / 5);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
