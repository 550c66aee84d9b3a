// K6: CFL update geometry.  Gram matrix G = U·Uᵀ of the n client updates U[n, P] (n ≤ #clients of a cluster,
// P up to 10⁷) in ONE pass over U: each CTA streams a P-chunk of all n rows through shared memory and
// accumulates the n×n partial products in registers/smem (fp32 products, fp64 cross-CTA accumulation).
// The reference computes n² numpy dots over re-flattened dicts in a python double loop
// (FedAvgEnsDataLoader.py:1236-1243) — O(n²·P) memory traffic; this is O(n·P).
#include "common.cuh"
#include "kernels.h"

namespace fdb {

constexpr int kChunk = 512;  // floats of each row staged per iteration

__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ U, int n, long long P, double* __restrict__ part) {
    extern __shared__ float tile[];  // [n][kChunk + 1]
    const int pairs = n * (n + 1) / 2;
    // each thread owns a set of (i<=j) pairs; partial sums in fp32 per chunk, flushed to fp64 accumulators
    double acc_local[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc_local[u] = 0.0;
    for (long long c0 = (long long)blockIdx.x * kChunk; c0 < P; c0 += (long long)gridDim.x * kChunk) {
        const int len = (int)min((long long)kChunk, P - c0);
        for (int e = threadIdx.x; e < n * kChunk; e += blockDim.x) {
            const int r = e / kChunk, k = e - r * kChunk;
            tile[r * (kChunk + 1) + k] = (k < len) ? __ldcs(U + (size_t)r * P + c0 + k) : 0.f;
        }
        __syncthreads();
        // 8 lanes cooperate on one pair: split the chunk in 8 interleaved slices
        const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
        int slot = 0;
        for (int pr = grp; pr < pairs && slot < 8; pr += ngrp, ++slot) {
            int i = 0, rem = pr;  // unrank (i, j) with i <= j
            while (rem >= n - i) { rem -= n - i; ++i; }
            const int j = i + rem;
            const float* a = tile + i * (kChunk + 1);
            const float* b = tile + j * (kChunk + 1);
            float s = 0.f;
            for (int k = sub; k < kChunk; k += 8) s = fmaf(a[k], b[k], s);
            acc_local[slot] += (double)s;
        }
        __syncthreads();
    }
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
    int slot = 0;
    for (int pr = grp; pr < pairs && slot < 8; pr += ngrp, ++slot) {
        double v = acc_local[slot];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == 0) part[(size_t)blockIdx.x * pairs + pr] = v;   // one partial per (CTA, pair): no atomics on hot addresses
    }
}

// second stage: fold the per-CTA partial Gram matrices (grid × pairs doubles) into the symmetric n×n result
__global__ void gram_finish_kernel(const double* __restrict__ part, int grid, int n, double* __restrict__ G) {
    const int pairs = n * (n + 1) / 2;
    for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
        double v = 0.0;
        for (int b = 0; b < grid; ++b) v += part[(size_t)b * pairs + pr];
        int i = 0, rem = pr;
        while (rem >= n - i) { rem -= n - i; ++i; }
        const int j = i + rem;
        G[(size_t)i * n + j] = v;
        G[(size_t)j * n + i] = v;
    }
}

int gram_launch(const float* U, int n, long long P, double* G, cudaStream_t stream) {
    // supports n(n+1)/2 <= 8 * 32 pairs per CTA (n <= 22); larger clusters are tiled by the caller
    if (n * (n + 1) / 2 > 8 * 32) return -5;
    const int pairs = n * (n + 1) / 2;
    const int smem = n * (kChunk + 1) * (int)sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const long long chunks = (P + kChunk - 1) / kChunk;
    const int blocks = (int)max(1LL, min(chunks, 148LL * 4));
    double* part = nullptr;
    if (cudaMallocAsync(&part, sizeof(double) * (size_t)blocks * pairs, stream) != cudaSuccess) return -8;
    gram_kernel<<<blocks, 256, smem, stream>>>(U, n, P, part);
    gram_finish_kernel<<<1, 256, 0, stream>>>(part, blocks, n, G);
    cudaFreeAsync(part, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
