// K6: CFL update geometry.  Gram matrix G = U·Uᵀ of the n client updates U[n, P] (n ≤ #clients of a cluster,
// P up to 10⁷) in ONE pass over U: each CTA streams a P-chunk of all n rows through shared memory and
// accumulates the n×n partial products in registers/smem (fp32 products, fp64 cross-CTA accumulation).
// The reference computes n² numpy dots over re-flattened dicts in a python double loop
// (FedAvgEnsDataLoader.py:1236-1243) — O(n²·P) memory traffic; this is O(n·P).
#include "common.cuh"
#include "kernels.h"

namespace fdb {

constexpr int kChunk = 512;  // floats of each row staged per iteration
constexpr int kRB = 8;       // rows per register block

// Register-blocked Gram: the n rows are grouped in blocks of 8; every warp owns one (I ≤ J) block pair and a
// slice of the staged columns.  Per column a lane reads 8+8 values from shared memory (conflict-free: lanes
// are adjacent columns) and issues 64 FMAs from registers — 4 FMA per LDS, i.e. balanced against the SM's
// 128-lane FMA rate — so the kernel stays on the HBM roofline instead of the shared-memory one.
// The chunk stream is DOUBLE-BUFFERED with cp.async (LDGSTS, 16-byte when rows are 16-byte aligned, 4-byte
// otherwise — model sizes are rarely multiples of 4): chunk i+1 is in flight while chunk i is multiplied.
// fp32 accumulators per lane (≤ a few hundred terms each), fp64 across lanes / CTAs.
FDB_DEVICE void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
FDB_DEVICE void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int RB>
__global__ void __launch_bounds__(384) gram_kernel(const float* __restrict__ U, int n, long long P, int nrb, int nbp, int wpb, int vec,
                                                   double* __restrict__ part) {
    extern __shared__ __align__(16) float tile[];  // [2][nrb*RB][kChunk + 4]; rows >= n stay zero
    constexpr int LD = kChunk + 4;   // +4: a row whose global start is not 16-byte aligned is staged shifted by (r·P mod 4)
    const int rows_pad = nrb * RB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bp = warp % nbp, ws = warp / nbp;
    int I = 0, rem = bp;
    while (rem >= nrb - I) { rem -= nrb - I; ++I; }
    const int J = I + rem;
    for (int b = 0; b < 2; ++b)
        for (int e = threadIdx.x; e < (rows_pad - n) * LD; e += blockDim.x) tile[(b * rows_pad + n) * LD + e] = 0.f;

    auto issue = [&](int b, long long c0) {
        float* dst = tile + (size_t)b * rows_pad * LD;
        const int len = (int)min((long long)kChunk, P - c0);
        if (vec && c0 + kChunk + 4 <= P) {
            // 16-byte LDGSTS from the aligned-down global address; element k of row r lands at r·LD + (r·P mod 4) + k
            constexpr int VPR = kChunk / 4 + 1;
            for (int e = threadIdx.x; e < n * VPR; e += blockDim.x) {
                const int r = e / VPR, v = e - r * VPR;
                const long long g0 = ((long long)r * P + c0) & ~3LL;
                cp_async16(dst + r * LD + 4 * v, U + g0 + 4 * v);
            }
        } else {
            for (int e = threadIdx.x; e < n * kChunk; e += blockDim.x) {
                const int r = e / kChunk, k = e - r * kChunk;
                float* d = dst + r * LD + (int)(((long long)r * P) & 3) + k;
                if (k < len) cp_async4(d, U + (size_t)r * P + c0 + k);
                else *d = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[RB][RB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
        for (int b = 0; b < RB; ++b) acc[a][b] = 0.f;
    const long long stride = (long long)gridDim.x * kChunk;
    long long c0 = (long long)blockIdx.x * kChunk;
    int buf = 0;
    if (c0 < P) issue(0, c0);
    for (; c0 < P; c0 += stride, buf ^= 1) {
        if (c0 + stride < P) {
            issue(buf ^ 1, c0 + stride);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* A = tile + ((size_t)buf * rows_pad + I * RB) * LD;
        const float* B = tile + ((size_t)buf * rows_pad + J * RB) * LD;
        int oa[RB], ob[RB];   // per-row staging shift (0..3)
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            oa[r] = r * LD + (int)(((long long)(I * RB + r) * P) & 3);
            ob[r] = r * LD + (int)(((long long)(J * RB + r) * P) & 3);
        }
        if (I == J) {   // diagonal block (warp-uniform): upper triangle only
            for (int k = lane + 32 * ws; k < kChunk; k += 32 * wpb) {
                float a[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) a[r] = A[oa[r] + k];
#pragma unroll
                for (int x = 0; x < RB; ++x)
#pragma unroll
                    for (int y = x; y < RB; ++y) acc[x][y] = fmaf(a[x], a[y], acc[x][y]);
            }
        } else {
            for (int k = lane + 32 * ws; k < kChunk; k += 32 * wpb) {
                float a[RB], b[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) { a[r] = A[oa[r] + k]; b[r] = B[ob[r] + k]; }
#pragma unroll
                for (int x = 0; x < RB; ++x)
#pragma unroll
                    for (int y = 0; y < RB; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
            }
        }
        __syncthreads();   // everyone is done with `buf` before the next iteration's prefetch overwrites it
    }
    double* out = part + ((size_t)blockIdx.x * (nbp * wpb) + warp) * (kRB * kRB);
#pragma unroll
    for (int x = 0; x < RB; ++x)
#pragma unroll
        for (int y = 0; y < RB; ++y) {
            double v = (double)acc[x][y];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) out[x * kRB + y] = v;
        }
}

// second stage (one CTA): fold the per-(CTA, warp) partial blocks into G, then the cosine normalisation
// S_ij = G_ij / (‖u_i‖‖u_j‖ + eps) and the row norms — the whole K6 post-processing without further launches.
__global__ void __launch_bounds__(1024) gram_finish_kernel(const double* __restrict__ part, int grid, int n, int rb, int nrb, int nbp,
                                                           int wpb, double eps, double* __restrict__ S, double* __restrict__ nrm) {
    __shared__ double Gs[32 * 32];
    const int pairs = n * (n + 1) / 2;
    const int nw = nbp * wpb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pr = warp; pr < pairs; pr += nwarps) {
        int i = 0, rem = pr;
        while (rem >= n - i) { rem -= n - i; ++i; }
        const int j = i + rem;
        const int I = i / rb, J = j / rb;
        int bp = 0;
        for (int q = 0; q < I; ++q) bp += nrb - q;
        bp += J - I;
        const int off = (i % rb) * kRB + (j % rb);
        double v = 0.0;
        for (int e = lane; e < grid * wpb; e += 32) {
            const int b = e / wpb, w = bp + (e % wpb) * nbp;
            v += part[((size_t)b * nw + w) * (kRB * kRB) + off];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) { Gs[i * 32 + j] = v; Gs[j * 32 + i] = v; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        S[e] = Gs[i * 32 + j] / (sqrt(Gs[i * 32 + i]) * sqrt(Gs[j * 32 + j]) + eps);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) nrm[i] = sqrt(Gs[i * 32 + i]);
}

template <int RB>
static void gram_dispatch(int blocks, int threads, int smem, cudaStream_t stream, const float* U, int n, long long P, int nrb, int nbp, int wpb,
                          int vec, double* part) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(gram_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gram_kernel<RB><<<blocks, threads, smem, stream>>>(U, n, P, nrb, nbp, wpb, vec, part);
}

int gram_launch(const float* U, int n, long long P, double eps, double* S, double* nrm, cudaStream_t stream) {
    // n ≤ 32 per call (≤ 4 row blocks of ≤ 8 rows → ≤ 10 block pairs); larger clusters fall back in the caller
    if (n > 32 || n < 1) return -5;
    const int nrb = (n + kRB - 1) / kRB;
    const int rb = (n + nrb - 1) / nrb;          // rows per block: the smallest that covers n (n=10 → 2 blocks of 5)
    const int nbp = nrb * (nrb + 1) / 2;
    const int wpb = nbp == 1 ? 8 : nbp == 3 ? 4 : nbp == 6 ? 2 : 1;   // 8 / 12 / 12 / 10 warps
    const int threads = nbp * wpb * 32;
    const int smem = 2 * nrb * rb * (kChunk + 4) * (int)sizeof(float);
    const int vec = (reinterpret_cast<uintptr_t>(U) & 15) == 0;
    const long long chunks = (P + kChunk - 1) / kChunk;
    const int per_sm = max(1, min(3, (200 * 1024) / smem));
    const int blocks = (int)max(1LL, min(chunks, 148LL * per_sm));
    double* part = nullptr;
    if (cudaMallocAsync(&part, sizeof(double) * (size_t)blocks * nbp * wpb * kRB * kRB, stream) != cudaSuccess) return -8;
    switch (rb) {
        case 1: gram_dispatch<1>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 2: gram_dispatch<2>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 3: gram_dispatch<3>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 4: gram_dispatch<4>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 5: gram_dispatch<5>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 6: gram_dispatch<6>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        case 7: gram_dispatch<7>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
        default: gram_dispatch<8>(blocks, threads, smem, stream, U, n, P, nrb, nbp, wpb, vec, part); break;
    }
    gram_finish_kernel<<<1, 1024, 0, stream>>>(part, blocks, n, rb, nrb, nbp, wpb, eps, S, nrm);
    cudaFreeAsync(part, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
