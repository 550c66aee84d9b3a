// K13: finite-field GEMM for TurboAggregate's BGW / Lagrange-coded secret sharing (reference:
// fedml_api/distributed/turboaggregate/mpc_function.py:4-275 — numpy int64 `np.mod(A.dot(B), p)` which
// silently overflows for p > 2³¹).  C = (A · B) mod p, exact for any modulus p < 2⁶³.
//
// Odd p (every prime field the protocol uses): MONTGOMERY arithmetic with R = 2⁶⁴.  A is converted to Montgomery form
// ã = a·R mod p while it is staged into shared memory (one REDC per element), B stays plain, so REDC(ã·b) = a·b mod p needs
// no conversion back.  Products are accumulated UNREDUCED in 128 bits (IMAD.WIDE chains) for `lazy` = ⌊2⁶⁴/p⌋ terms — the
// largest count that keeps the sum below p·R — and reduced with ONE REDC (two 64×64 multiplies) per `lazy` terms: for the
// 31/32-bit primes of the reference that is one reduction per K-tile, ~2 integer multiply-adds per term.
// Even p falls back to the shift-subtract kernel.  2×2 outputs per thread, tiles staged through shared memory.
#include "common.cuh"
#include "kernels.h"

namespace fdb {

FDB_DEVICE unsigned long long mulmod_u64(unsigned long long a, unsigned long long b, unsigned long long p) {
    const unsigned long long hi = __umul64hi(a, b), lo = a * b;
    if (hi == 0) return lo % p;
    // reduce the 128-bit product: (hi·2⁶⁴ + lo) mod p via 64 shift-subtract steps on hi
    unsigned long long r = hi % p;
    for (int i = 0; i < 64; ++i) {
        const unsigned long long top = r >> 63;
        r <<= 1;
        if (top || r >= p) r -= p;
    }
    r += lo % p;
    if (r >= p || r < lo % p) r -= p;
    return r;
}

constexpr int kT = 16;
__global__ void modp_matmul_kernel(const long long* __restrict__ A, const long long* __restrict__ B, long long* __restrict__ C, int M, int K,
                                   int N, unsigned long long p) {
    __shared__ unsigned long long As[kT][kT + 1], Bs[kT][kT + 1];
    const int row = blockIdx.y * kT + threadIdx.y, col = blockIdx.x * kT + threadIdx.x;
    unsigned long long acc = 0;
    for (int k0 = 0; k0 < K; k0 += kT) {
        long long a = (row < M && k0 + threadIdx.x < K) ? A[(size_t)row * K + k0 + threadIdx.x] : 0;
        long long b = (col < N && k0 + threadIdx.y < K) ? B[(size_t)(k0 + threadIdx.y) * N + col] : 0;
        a %= (long long)p; if (a < 0) a += (long long)p;
        b %= (long long)p; if (b < 0) b += (long long)p;
        As[threadIdx.y][threadIdx.x] = (unsigned long long)a;
        Bs[threadIdx.y][threadIdx.x] = (unsigned long long)b;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kT; ++k) {
            const unsigned long long t = mulmod_u64(As[threadIdx.y][k], Bs[k][threadIdx.x], p);
            acc += t;
            if (acc >= p) acc -= p;
        }
        __syncthreads();
    }
    if (row < M && col < N) C[(size_t)row * N + col] = (long long)acc;
}

// ---- Montgomery path -------------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;
typedef unsigned long long u64;

// REDC: t < p·2⁶⁴  →  t·2⁻⁶⁴ mod p
FDB_DEVICE u64 mont_redc(u128 t, u64 p, u64 pinv_neg) {
    const u64 m = (u64)t * pinv_neg;
    const u128 s = t + (u128)m * p;          // < 2¹²⁸ because both terms are < 2¹²⁷ (p < 2⁶³)
    u64 u = (u64)(s >> 64);
    if (u >= p) u -= p;
    return u;
}

constexpr int kMT = 32;   // 32×32 output tile, 16×16 threads, 2×2 outputs per thread
__global__ void __launch_bounds__(256) modp_matmul_mont_kernel(const long long* __restrict__ A, const long long* __restrict__ B,
                                                               long long* __restrict__ C, int M, int K, int N, u64 p, u64 pinv_neg, u64 r2,
                                                               int lazy) {
    __shared__ u64 As[kMT][kT + 1], Bs[kT][kMT + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
    const int row0 = blockIdx.y * kMT, col0 = blockIdx.x * kMT;
    u64 res[2][2] = {{0, 0}, {0, 0}};
    u128 acc[2][2] = {{0, 0}, {0, 0}};
    int pending = 0;
    for (int k0 = 0; k0 < K; k0 += kT) {
        for (int i = tid; i < kMT * kT; i += 256) {     // A tile [32 rows][16 k] → Montgomery form
            const int r = i / kT, k = i % kT;
            long long a = (row0 + r < M && k0 + k < K) ? A[(size_t)(row0 + r) * K + k0 + k] : 0;
            a %= (long long)p; if (a < 0) a += (long long)p;
            As[r][k] = mont_redc((u128)(u64)a * r2, p, pinv_neg);
        }
        for (int i = tid; i < kT * kMT; i += 256) {     // B tile [16 k][32 cols], plain residues
            const int k = i / kMT, c = i % kMT;
            long long b = (col0 + c < N && k0 + k < K) ? B[(size_t)(k0 + k) * N + col0 + c] : 0;
            b %= (long long)p; if (b < 0) b += (long long)p;
            Bs[k][c] = (u64)b;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kT; ++k) {
            const u64 a0 = As[ty][k], a1 = As[ty + 16][k], b0 = Bs[k][tx], b1 = Bs[k][tx + 16];
            acc[0][0] += (u128)a0 * b0; acc[0][1] += (u128)a0 * b1;
            acc[1][0] += (u128)a1 * b0; acc[1][1] += (u128)a1 * b1;
            if (++pending == lazy) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        u64 v = res[i][j] + mont_redc(acc[i][j], p, pinv_neg);
                        if (v >= p) v -= p;
                        res[i][j] = v; acc[i][j] = 0;
                    }
                pending = 0;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            u64 v = res[i][j];
            if (pending) { v += mont_redc(acc[i][j], p, pinv_neg); if (v >= p) v -= p; }
            const int r = row0 + ty + 16 * i, c = col0 + tx + 16 * j;
            if (r < M && c < N) C[(size_t)r * N + c] = (long long)v;
        }
}

int modp_matmul_launch(const long long* A, const long long* B, long long* C, int M, int K, int N, long long p, cudaStream_t stream) {
    const u64 up = (u64)p;
    if (p > 2 && (up & 1ull)) {
        // -p⁻¹ mod 2⁶⁴ by Newton iteration; R² mod p by 128 doublings
        u64 inv = up;                                   // correct to 3 bits for odd p
        for (int i = 0; i < 6; ++i) inv *= 2ull - up * inv;
        const u64 pinv_neg = 0ull - inv;
        u64 r2 = 1ull % up;
        for (int i = 0; i < 128; ++i) { r2 <<= 1; if (r2 >= up) r2 -= up; }   // p < 2⁶³: no overflow
        u64 lazy = (up <= 1ull) ? 1ull : (~0ull / up);                       // ⌊(2⁶⁴−1)/p⌋ terms keep Σ < p·2⁶⁴
        if (lazy < 1) lazy = 1;
        if (lazy > (1ull << 30)) lazy = 1ull << 30;
        dim3 block(16, 16), grid((N + kMT - 1) / kMT, (M + kMT - 1) / kMT);
        modp_matmul_mont_kernel<<<grid, block, 0, stream>>>(A, B, C, M, K, N, up, pinv_neg, r2, (int)lazy);
        return cudaGetLastError() == cudaSuccess ? 0 : -4;
    }
    dim3 block(kT, kT), grid((N + kT - 1) / kT, (M + kT - 1) / kT);
    modp_matmul_kernel<<<grid, block, 0, stream>>>(A, B, C, M, K, N, up);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
