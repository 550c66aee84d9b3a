// K13: finite-field GEMM for TurboAggregate's BGW / Lagrange-coded secret sharing (reference:
// fedml_api/distributed/turboaggregate/mpc_function.py:4-275 — numpy int64 `np.mod(A.dot(B), p)` which
// silently overflows for p > 2³¹).  C = (A · B) mod p with unsigned 64×64→128-bit products reduced per term,
// so any modulus p < 2⁶³ is exact.  Tiled through shared memory; one thread per output element.
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

int modp_matmul_launch(const long long* A, const long long* B, long long* C, int M, int K, int N, long long p, cudaStream_t stream) {
    dim3 block(kT, kT), grid((N + kT - 1) / kT, (M + kT - 1) / kT);
    modp_matmul_kernel<<<grid, block, 0, stream>>>(A, B, C, M, K, N, (unsigned long long)p);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
