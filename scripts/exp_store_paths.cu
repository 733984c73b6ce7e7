// exp_store_paths.cu -- how fast can 1920-byte observation rows be WRITTEN on a B200, by store path?
// Stand-alone experiment (not part of the library):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/exp_store_paths scripts/exp_store_paths.cu
// Variants (all expand 15 words of bits per row into 480 f32 0/1, like emit_obs_row):
//   0  warp-per-row float4 stores (what k_produce does)
//   1  same with st.global.cs (streaming)
//   2  rows staged in shared memory, one cp.async.bulk.global.shared::cta per chunk of R rows (TMA bulk store)
//   3  cudaMemsetAsync (pure-write ceiling)
//   4  persistent variant of 0 (grid = 148 * k)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t row_word(int64_t row, int w) {
    uint32_t x = (uint32_t)row * 2654435761u + (uint32_t)w * 0x9E3779B9u;
    x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13;
    return x & (x >> 7) & (x << 3);  // sparse bits, like a real observation
}
__device__ __forceinline__ float4 nib4(uint32_t nib) {
    return make_float4((nib & 1u) ? 1.f : 0.f, (nib & 2u) ? 1.f : 0.f, (nib & 4u) ? 1.f : 0.f, (nib & 8u) ? 1.f : 0.f);
}

template <int CS>
__global__ void __launch_bounds__(128) k_plain(float4* out, int64_t n_rows, int rows_per_warp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t total_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t base = warp * rows_per_warp; base < n_rows; base += total_warps * rows_per_warp) {
        for (int r = 0; r < rows_per_warp && base + r < n_rows; ++r) {
            const int64_t row = base + r;
            uint32_t my = lane < 15 ? row_word(row, lane) : 0u;
            float4* dst = out + row * 120;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = lane + 32 * k;
                uint32_t w = __shfl_sync(0xffffffffu, my, (j >> 3) & 15);
                if (j < 120) {
                    float4 v = nib4((w >> ((j & 7) * 4)) & 15u);
                    if (CS) __stcs(dst + j, v); else dst[j] = v;
                }
            }
        }
    }
}

// block-cooperative: a block owns B consecutive rows; its warps write them interleaved (warp w: rows w, w+W, ...)
__global__ void __launch_bounds__(256) k_block(float4* out, int64_t n_rows, int B) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * B;
    for (int r = w; r < B && base + r < n_rows; r += W) {
        const int64_t row = base + r;
        uint32_t my = lane < 15 ? row_word(row, lane) : 0u;
        float4* dst = out + row * 120;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int j = lane + 32 * k;
            uint32_t wd = __shfl_sync(0xffffffffu, my, (j >> 3) & 15);
            if (j < 120) dst[j] = nib4((wd >> ((j & 7) * 4)) & 15u);
        }
    }
}

// like k_block, plus what the env kernels do before writing: the first B threads each read 80 B of state
// (5 coalesced 128-bit plane loads) and stage 15 words in shared memory; GATHER adds a dependent 16-byte
// gather from a 4.8 MB (L2-resident) table
template <int GATHER>
__global__ void __launch_bounds__(256) k_block_rd(float4* out, const uint4* st, const uint4* table, int64_t n_rows, int B) {
    __shared__ uint32_t R[32 * 17];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, W = blockDim.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * B;
    if (tid < B && base + tid < n_rows) {
        const int64_t i = base + tid;
        uint4 p0 = st[i], p1 = st[n_rows + i], p2 = st[2 * n_rows + i], p3 = st[3 * n_rows + i], p4 = st[4 * n_rows + i];
        uint32_t x = p0.x ^ p1.y ^ p2.z ^ p3.w ^ p4.x;
        if (GATHER) { uint4 g = table[(row_word(i, 3) ^ x) % 300000u]; x ^= g.x; }
        for (int k = 0; k < 15; ++k) R[tid * 17 + k] = row_word(i, k) ^ (x & 1u);
    }
    __syncthreads();
    for (int r = w; r < B && base + r < n_rows; r += W) {
        float4* dst = out + (base + r) * 120;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int j = lane + 32 * k;
            if (j < 120) dst[j] = nib4((R[r * 17 + (j >> 3)] >> ((j & 7) * 4)) & 15u);
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp: double-buffered chunk of R rows in smem, bulk-stored by lane 0
template <int R>
__global__ void __launch_bounds__(128) k_bulk(float4* out, int64_t n_rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4* buf = reinterpret_cast<float4*>(smem) + (size_t)wib * 2 * R * 120;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int64_t total_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int b = 0;
    for (int64_t base = warp * R; base < n_rows; base += total_warps * R, b ^= 1) {
        // buffer b was last used two chunks ago: allow at most 1 outstanding bulk group (reads of smem done)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        float4* dstb = buf + (size_t)b * R * 120;
        const int nr = (int)((n_rows - base) < R ? (n_rows - base) : R);
        for (int r = 0; r < nr; ++r) {
            const int64_t row = base + r;
            uint32_t my = lane < 15 ? row_word(row, lane) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = lane + 32 * k;
                uint32_t w = __shfl_sync(0xffffffffu, my, (j >> 3) & 15);
                if (j < 120) dstb[r * 120 + j] = nib4((w >> ((j & 7) * 4)) & 15u);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + base * 120),
                         "r"(smem_u32(dstb)), "r"(nr * 1920) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float time_ms(F f, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

int main(int argc, char** argv) {
    const int64_t n_rows = argc > 1 ? atoll(argv[1]) : 1048576;
    const size_t bytes = (size_t)n_rows * 1920;
    float4* out; CK(cudaMalloc(&out, bytes));
    float4* src; CK(cudaMalloc(&src, bytes));
    CK(cudaMemset(src, 0, bytes));
    const int iters = 20;
    auto report = [&](const char* name, float ms, double mult = 1.0) {
        printf("%-58s %8.4f ms  %8.1f GB/s\n", name, ms, mult * bytes / ms / 1e6);
    };
    report("cudaMemsetAsync", time_ms([&] { CK(cudaMemsetAsync(out, 0, bytes)); }, iters));
    report("cudaMemcpyAsync d2d (read+write bytes)", time_ms([&] { CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToDevice)); }, iters), 2.0);
    for (int rpw : {1, 8, 32}) {
        char nm[128];
        unsigned grid = (unsigned)((n_rows + 4 * rpw - 1) / (4 * rpw));
        snprintf(nm, sizeof nm, "plain float4, %d rows/warp, grid %u", rpw, grid);
        report(nm, time_ms([&] { k_plain<0><<<grid, 128>>>(out, n_rows, rpw); }, iters));
        snprintf(nm, sizeof nm, "st.cs float4, %d rows/warp, grid %u", rpw, grid);
        report(nm, time_ms([&] { k_plain<1><<<grid, 128>>>(out, n_rows, rpw); }, iters));
    }
    for (int rpw : {2, 4, 16}) {
        char nm[128];
        unsigned grid = (unsigned)((n_rows + 4 * rpw - 1) / (4 * rpw));
        snprintf(nm, sizeof nm, "plain float4, %d rows/warp, grid %u", rpw, grid);
        report(nm, time_ms([&] { k_plain<0><<<grid, 128>>>(out, n_rows, rpw); }, iters));
    }
    for (int threads : {128, 256})
        for (int B : {32, 64, 128, 256}) {
            char nm[128];
            unsigned grid = (unsigned)((n_rows + B - 1) / B);
            snprintf(nm, sizeof nm, "block-coop, %d rows/block of %d thr, grid %u", B, threads, grid);
            report(nm, time_ms([&] { k_block<<<grid, threads>>>(out, n_rows, B); }, iters));
        }
    {
        uint4* st = reinterpret_cast<uint4*>(src);  // zeros: 80 B/row planes
        for (int threads : {64, 128, 256})
            for (int B : {8, 16, 32}) {
                char nm[128];
                unsigned grid = (unsigned)((n_rows + B - 1) / B);
                snprintf(nm, sizeof nm, "block-coop + 80 B state read, %d rows/block of %d thr", B, threads);
                report(nm, time_ms([&] { k_block_rd<0><<<grid, threads>>>(out, st, st, n_rows, B); }, iters));
                snprintf(nm, sizeof nm, "block-coop + state read + table gather, %d rows/block of %d thr", B, threads);
                report(nm, time_ms([&] { k_block_rd<1><<<grid, threads>>>(out, st, st + 5 * n_rows, n_rows, B); }, iters));
            }
    }
    for (int k : {4, 8, 16}) {
        char nm[128];
        snprintf(nm, sizeof nm, "plain float4 persistent, 8 rows/warp, grid 148*%d", k);
        report(nm, time_ms([&] { k_plain<0><<<148 * k, 128>>>(out, n_rows, 8); }, iters));
    }
    {
        auto run = [&](auto kern, int R, int bps) {
            size_t sm = (size_t)4 * 2 * R * 1920;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            char nm[128];
            snprintf(nm, sizeof nm, "TMA bulk store, %d rows/chunk, grid 148*%d, %zu KB smem", R, bps, sm / 1024);
            report(nm, time_ms([&] { kern<<<148 * bps, 128, sm>>>(out, n_rows); }, iters));
        };
        run(k_bulk<1>, 1, 8); run(k_bulk<1>, 1, 14);
        run(k_bulk<2>, 2, 7);
        run(k_bulk<4>, 4, 3);
        run(k_bulk<8>, 8, 1);
        run(k_bulk<2>, 2, 4);
        run(k_bulk<4>, 4, 2);
    }
    CK(cudaDeviceSynchronize());
    // correctness spot check of the bulk path vs the plain path
    {
        float4* ref; CK(cudaMalloc(&ref, (size_t)4096 * 1920));
        k_plain<0><<<128, 128>>>(ref, 4096, 8);
        CK(cudaMemset(out, 0xFF, (size_t)4096 * 1920));
        CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2 * 4 * 1920));
        k_bulk<4><<<37, 128, 4 * 2 * 4 * 1920>>>(out, 4096);
        CK(cudaDeviceSynchronize());
        size_t nb = (size_t)4096 * 1920;
        unsigned char* ha = (unsigned char*)malloc(nb); unsigned char* hb = (unsigned char*)malloc(nb);
        CK(cudaMemcpy(ha, ref, nb, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, out, nb, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (size_t i = 0; i < nb; ++i) bad += ha[i] != hb[i];
        printf("bulk vs plain mismatching bytes: %zu\n", bad);
    }
    return 0;
}
