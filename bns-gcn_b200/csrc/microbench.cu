// microbench.cu -- independent L2 / HBM bandwidth probes (diagnostics, NOT part of the product library).
//
// "the claim 'at the L2->SM ceiling' rests on the same kernel run on a 51 MB matrix -- a
// self-referential ceiling".  These kernels share no code with spmm_kernel: no CSR, no index loads, no shared memory.
//
//   bnsm_stream_read   every thread streams 16-byte loads over a buffer of `bytes` (L2-resident when it fits the 126 MB
//                      L2 and was touched before; HBM otherwise), `reps` passes inside ONE launch.
//   bnsm_row_gather    every warp gathers pseudo-random rows of `row_bytes` (512 / 1024: the slab rows of the SpMM) from
//                      a table of `n_rows` rows, UNROLL independent 16-byte loads in flight per lane, ids from an
//                      in-register LCG (no index stream).  This is the access pattern of the SpMM stripped of everything
//                      else: its GB/s is the fabric ceiling for "random 512-byte-row gather".
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --shared -Xcompiler -fPIC -o libbnsmicro.so microbench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace {

__device__ __forceinline__ float4 ld_nc_na(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) stream_read_kernel(const float4 *__restrict__ buf, int64_t n4, int reps,
                                                          float *__restrict__ sink) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) v[u] = ld_nc_na(buf + i + u * stride);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
        for (; i < n4; i += stride) {
            const float4 v = ld_nc_na(buf + i);
            acc += (v.x + v.y) + (v.z + v.w);
        }
    }
    if (acc == 123.456f) *sink = acc;        // never true: keeps the loads alive
}

// one warp per "chunk": gathers `per_warp` rows; lane l reads 16 bytes at offset 16*l (+512 for the second half of a
// 1 KB row) of each row -> a fully coalesced 512-byte request per row per instruction, like the SpMM's gathers
template <int UNROLL, int NV>
__global__ void __launch_bounds__(256) row_gather_kernel(const float4 *__restrict__ table, uint32_t n_rows, int row_f4,
                                                         int64_t per_warp, uint32_t seed, float *__restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    uint32_t s = seed ^ (uint32_t)(warp * 2654435761u);
    float4 acc[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t j = 0; j < per_warp; j += UNROLL) {
        float4 v[UNROLL][NV];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = s * 1664525u + 1013904223u;                        // same id on every lane of the warp
            const uint32_t row = (uint32_t)(((uint64_t)(s >> 4) * n_rows) >> 28);
            const float4 *p = table + (int64_t)row * row_f4 + lane;
#pragma unroll
            for (int t = 0; t < NV; ++t) v[u][t] = ld_nc_na(p + 32 * t);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int t = 0; t < NV; ++t) {
                acc[t].x += v[u][t].x; acc[t].y += v[u][t].y; acc[t].z += v[u][t].z; acc[t].w += v[u][t].w;
            }
    }
    float tot = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) tot += (acc[t].x + acc[t].y) + (acc[t].z + acc[t].w);
    if (tot == 123.456f) *sink = tot;
}

__global__ void fill_kernel(float4 *buf, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int sms() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

}  // namespace

// returns GB/s (bytes * reps / time of the launch), or a negative CUDA error code
extern "C" double bnsm_stream_read(int64_t bytes, int reps, int blocks_per_sm, int iters) {
    float4 *buf = nullptr;
    float *sink = nullptr;
    const int64_t n4 = bytes / 16;
    if (cudaMalloc(&buf, n4 * 16) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) return -1.0;
    fill_kernel<<<sms() * 8, 256>>>(buf, n4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)(sms() * blocks_per_sm);
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it) {              // 2 warm-up launches (they also pull the buffer into L2)
        cudaEventRecord(e0);
        stream_read_kernel<8><<<grid, 256>>>(buf, n4, reps, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2.0;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2 && ms < best) best = ms;
    }
    cudaFree(buf);
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -3.0;
    return (double)n4 * 16.0 * reps / (best * 1e-3) / 1e9;
}

// returns GB/s of gathered row bytes
extern "C" double bnsm_row_gather(int64_t n_rows, int row_bytes, int64_t total_rows_gathered, int blocks_per_sm, int unroll,
                                  int iters) {
    float4 *buf = nullptr;
    float *sink = nullptr;
    const int row_f4 = row_bytes / 16;
    const int64_t n4 = n_rows * row_f4;
    if (row_bytes != 512 && row_bytes != 1024) return -4.0;
    if (cudaMalloc(&buf, n4 * 16) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) return -1.0;
    fill_kernel<<<sms() * 8, 256>>>(buf, n4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)(sms() * blocks_per_sm);
    const int64_t warps = (int64_t)grid * 8;
    int64_t per_warp = total_rows_gathered / warps;
    per_warp -= per_warp % 8;
    if (per_warp < 8) per_warp = 8;
    float best = 1e30f;
    for (int it = 0; it < iters + 2; ++it) {
        cudaEventRecord(e0);
        const uint32_t seed = 12345u + it;
        if (row_bytes == 1024) {
            if (unroll >= 4) row_gather_kernel<4, 2><<<grid, 256>>>(buf, (uint32_t)n_rows, row_f4, per_warp, seed, sink);
            else row_gather_kernel<2, 2><<<grid, 256>>>(buf, (uint32_t)n_rows, row_f4, per_warp, seed, sink);
        } else {
            if (unroll >= 8) row_gather_kernel<8, 1><<<grid, 256>>>(buf, (uint32_t)n_rows, row_f4, per_warp, seed, sink);
            else row_gather_kernel<4, 1><<<grid, 256>>>(buf, (uint32_t)n_rows, row_f4, per_warp, seed, sink);
        }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2.0;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2 && ms < best) best = ms;
    }
    cudaFree(buf);
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -3.0;
    const double bytes = (double)per_warp * (double)warps * (double)row_bytes;
    return bytes / (best * 1e-3) / 1e9;
}
