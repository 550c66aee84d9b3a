// K6: CFL update geometry.  Gram matrix G = U·Uᵀ of the n client updates U[n, P] (n ≤ 32 per call, P up to 10⁸) in
// ONE pass over U, followed by the cosine normalisation — the reference computes n² numpy dots over re-flattened
// state_dicts in a python double loop (FedAvgEnsDataLoader.py:1236-1243): O(n²·P) traffic; this is O(n·P).
//
// Design (HBM-bound by construction):
//   * every CTA streams 512-column chunks of all n rows into a 3-stage shared-memory ring with TMA bulk copies
//     (cp.async.bulk, ONE 2 KB copy per row per chunk, completion on an mbarrier with a transaction count): chunks i+1
//     and i+2 are in flight while chunk i is multiplied, and no thread spends issue slots on address generation.  Rows
//     whose global start is not 16-byte aligned (model sizes are rarely multiples of 4) are copied from the
//     aligned-down address and read back with a per-row shift, so the fast path does not depend on P;
//   * the products run on the tensor cores: mma.sync m16n8k8 TF32.  For P < 2¹⁶ the 3×TF32 split (a = hi + lo; hi·hi +
//     hi·lo + lo·hi) gives fp32-level accuracy; for longer rows plain round-to-nearest TF32 is used: its rounding error
//     is unbiased, so the relative error of a Gram entry is ≈ 2⁻¹¹/√P ≤ 2·10⁻⁶ — below the fp32 accumulation error —
//     at a third of the tensor-pipe work.  Because B = Uᵀ, the B fragment of n-tile j IS the A fragment data of rows
//     8j..8j+7 — one set of 2 shared-memory loads per 8 rows feeds both operands.  The row pitch (516 floats ≡ 4 mod 32
//     banks) makes the fragment loads conflict-free.  Only tiles on or above the diagonal are computed;
//   * the skinny shape (n ≤ 32 rows) is why this is mma.sync and not tcgen05: a UMMA tile is at least 64 rows of M,
//     which would waste ≥ half of the tensor core and all of TMEM for no bandwidth gain;
//   * fp32 MMA accumulators are flushed into per-(CTA, warp) fp64 slots every 64 chunks (and at the end); a one-CTA
//     second kernel folds the slots, mirrors the triangle and writes S_ij = G_ij / (‖u_i‖‖u_j‖ + eps) and the norms.
#include "common.cuh"
#include "kernels.h"

namespace fdb {

constexpr int kChunk = 512;        // floats of each row staged per iteration
constexpr int kLD = kChunk + 4;    // row pitch: +4 for the alignment shift; 516 ≡ 4 (mod 32) → conflict-free fragments
constexpr int kGramWarps = 8;
constexpr int kTileDoubles = 16 * 8;   // one m16n8 accumulator tile

constexpr uint32_t kRowBytes = kLD * 4;   // 2064 B = 129 × 16 B: one bulk copy per row per chunk

FDB_DEVICE uint32_t gm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FDB_DEVICE void gm_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gm_smem(bar)), "r"(count));
}
FDB_DEVICE void gm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gm_smem(bar)), "r"(bytes) : "memory");
}
FDB_DEVICE void gm_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    SpinGuard g;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(gm_smem(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (g.expired(2000000000LL)) __trap();   // a protocol bug must fault, never hang the GPU
    }
}
FDB_DEVICE void gm_bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gm_smem(dst)), "l"(src), "r"(bytes), "r"(gm_smem(bar)) : "memory");
}
FDB_DEVICE uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
FDB_DEVICE void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tiles (mt, nt) on/above the diagonal for R8 row groups of 8: nt ≥ 2·mt
template <int R8> struct GramTiles {
    static constexpr int MT = (R8 + 1) / 2;
    static constexpr int count() { int c = 0; for (int mt = 0; mt < MT; ++mt) for (int nt = 2 * mt; nt < R8; ++nt) ++c; return c; }
};

template <int R8> constexpr int gram_stages() { return R8 == 1 ? 4 : 3; }

template <int R8, bool SPLIT3>
__global__ void __launch_bounds__(kGramWarps * 32) gram_kernel(const float* __restrict__ U, int n, long long P, double* __restrict__ part) {
    extern __shared__ __align__(128) float tile[];  // [kStages][ROWS][kLD] then kStages mbarriers; rows >= n stay zero
    constexpr int kStages = gram_stages<R8>();
    constexpr int MT = GramTiles<R8>::MT;
    constexpr int ROWS = R8 * 8;          // only real 8-row groups are staged; fragment rows beyond them are constant zeros
    constexpr int NT = GramTiles<R8>::count();
    uint64_t* full = reinterpret_cast<uint64_t*>(tile + (size_t)kStages * ROWS * kLD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    for (int b = 0; b < kStages; ++b)
        for (int e = threadIdx.x; e < (ROWS - n) * kLD; e += blockDim.x) tile[(b * ROWS + n) * kLD + e] = 0.f;
    if (threadIdx.x == 0) {
        for (int b = 0; b < kStages; ++b) gm_mbar_init(full + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // chunks that can be bulk-copied: 516 floats from the aligned-down start must stay inside U
    const long long stride = (long long)gridDim.x * kChunk;
    auto bulk_ok = [&](long long c0) { return c0 + kChunk + 4 <= P; };
    auto issue = [&](int b, long long c0) {   // warp 0 only
        float* dst = tile + (size_t)b * ROWS * kLD;
        if (lane == 0) gm_mbar_expect_tx(full + b, (uint32_t)n * kRowBytes);
        __syncwarp();
        for (int r = lane; r < n; r += 32) {
            const long long g0 = ((long long)r * P + c0) & ~3LL;
            gm_bulk_copy(dst + r * kLD, U + g0, kRowBytes, full + b);
        }
    };

    // this lane's fragment rows: g + 8j, j < 2·MT, with their staging shift
    int roff[2 * MT];
#pragma unroll
    for (int j = 0; j < 2 * MT; ++j) {
        const int r = g + 8 * j;
        roff[j] = (j < R8) ? r * kLD + (r < n ? (int)(((long long)r * P) & 3) : 0) + t : 0;
    }
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
    double* slot = part + ((size_t)blockIdx.x * kGramWarps + warp) * (NT * kTileDoubles);
    bool first = true;
    auto flush = [&]() {   // fp32 tile registers → this warp's private fp64 slots (plain RMW: the warp owns them)
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            double* tp = slot + i * kTileDoubles;
            const int e0 = g * 8 + 2 * t, e1 = (g + 8) * 8 + 2 * t;
            if (first) { tp[e0] = acc[i][0]; tp[e0 + 1] = acc[i][1]; tp[e1] = acc[i][2]; tp[e1 + 1] = acc[i][3]; }
            else { tp[e0] += acc[i][0]; tp[e0 + 1] += acc[i][1]; tp[e1] += acc[i][2]; tp[e1 + 1] += acc[i][3]; }
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
        }
        first = false;
    };

    auto multiply = [&](const float* T) {
#pragma unroll 2
        for (int k8 = warp; k8 < kChunk / 8; k8 += kGramWarps) {
            uint32_t hi[2 * MT][2], lo[2 * MT][2];
#pragma unroll
            for (int j = 0; j < 2 * MT; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (j < R8) {
                        const float x = T[roff[j] + 8 * k8 + 4 * h];
                        hi[j][h] = to_tf32(x);
                        lo[j][h] = SPLIT3 ? to_tf32(x - __uint_as_float(hi[j][h])) : 0u;
                    } else {
                        hi[j][h] = 0u; lo[j][h] = 0u;   // padding row group of the m16 tile
                    }
                }
            int ti = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t ah[4] = {hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1]};
                const uint32_t al[4] = {lo[2 * mt][0], lo[2 * mt + 1][0], lo[2 * mt][1], lo[2 * mt + 1][1]};
#pragma unroll
                for (int nt = 2 * mt; nt < R8; ++nt, ++ti) {
                    if (SPLIT3) {
                        mma_tf32(acc[ti], al, hi[nt][0], hi[nt][1]);   // small terms first
                        mma_tf32(acc[ti], ah, lo[nt][0], lo[nt][1]);
                    }
                    mma_tf32(acc[ti], ah, hi[nt][0], hi[nt][1]);
                }
            }
        }
    };

    long long c0 = (long long)blockIdx.x * kChunk;
    if (warp == 0)
        for (int b = 0; b < kStages; ++b)
            if (c0 + b * stride < P && bulk_ok(c0 + b * stride)) issue(b, c0 + b * stride);
    int since_flush = 0;
    long long it = 0;
    for (; c0 < P && bulk_ok(c0); c0 += stride, ++it) {
        const int b = (int)(it % kStages);
        gm_mbar_wait(full + b, (uint32_t)((it / kStages) & 1));
        multiply(tile + (size_t)b * ROWS * kLD);
        if (++since_flush == 64) { flush(); since_flush = 0; }
        __syncthreads();   // every warp is done with stage b before it is refilled
        const long long nxt = c0 + kStages * stride;
        if (warp == 0 && nxt < P && bulk_ok(nxt)) issue(b, nxt);
    }
    if (c0 < P) {
        // the (single) tail chunk of the matrix: partial length or within 4 floats of the end — plain loads, zero fill
        float* dst = tile;   // stage 0 is idle: all of this CTA's bulk chunks have been consumed
        const int len = (int)min((long long)kChunk, P - c0);
        for (int e = threadIdx.x; e < n * kChunk; e += blockDim.x) {
            const int r = e / kChunk, k = e - r * kChunk;
            dst[r * kLD + (int)(((long long)r * P) & 3) + k] = (k < len) ? __ldg(U + (size_t)r * P + c0 + k) : 0.f;
        }
        __syncthreads();
        multiply(dst);
    }
    flush();
    // fold the 8 warps' slots into warp 0's slot so the second stage reads one slot per CTA
    __syncthreads();
    double* cta = part + (size_t)blockIdx.x * kGramWarps * (NT * kTileDoubles);
    for (int e = threadIdx.x; e < NT * kTileDoubles; e += blockDim.x) {
        double v = cta[e];
#pragma unroll
        for (int w = 1; w < kGramWarps; ++w) v += cta[(size_t)w * NT * kTileDoubles + e];
        cta[e] = v;
    }
}

// second stage (one CTA): fold the per-(CTA, warp) tiles into G, mirror the triangle, then the cosine normalisation
// S_ij = G_ij / (‖u_i‖‖u_j‖ + eps) and the row norms — the whole K6 post-processing without further launches.
__global__ void __launch_bounds__(1024) gram_finish_kernel(const double* __restrict__ part, int slots, int n, int r8, double eps,
                                                           double* __restrict__ S, double* __restrict__ nrm) {
    __shared__ double Gs[32 * 32];
    const int mtiles = (r8 + 1) / 2;
    int ntile = 0;
    for (int mt = 0; mt < mtiles; ++mt) ntile += r8 - 2 * mt;
    const int pairs = n * (n + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pr = warp; pr < pairs; pr += nwarps) {
        int i = 0, rem = pr;
        while (rem >= n - i) { rem -= n - i; ++i; }
        const int j = i + rem;                      // i ≤ j  →  tile (mt = i/16, nt = j/8) has nt ≥ 2·mt: it was computed
        const int mt = i >> 4, nt = j >> 3;
        int ti = 0;
        for (int q = 0; q < mt; ++q) ti += r8 - 2 * q;
        ti += nt - 2 * mt;
        const int off = ti * kTileDoubles + (i & 15) * 8 + (j & 7);
        double v = 0.0;
        for (int e = lane; e < slots; e += 32) v += part[(size_t)e * kGramWarps * ntile * kTileDoubles + off];
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

constexpr int kGramMaxBlocks = 148 * 4;
// doubles of scratch the caller must provide (per-(CTA, warp) accumulator tiles); upper bound over all n ≤ 32
long long gram_workspace_doubles() { return (long long)kGramMaxBlocks * kGramWarps * GramTiles<4>::count() * kTileDoubles; }

template <int R8, bool SPLIT3>
static void gram_dispatch2(int per_sm_cap, cudaStream_t stream, const float* U, int n, long long P, double* part, int* blocks_out) {
    constexpr int stages = gram_stages<R8>();
    const int smem = stages * R8 * 8 * kLD * (int)sizeof(float) + stages * 8;
    if (smem > 48 * 1024) cudaFuncSetAttribute(gram_kernel<R8, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const long long chunks = (P + kChunk - 1) / kChunk;
    const int per_sm = max(1, min(per_sm_cap, (220 * 1024) / smem));   // R8=1: 3 CTAs/SM (66 KB), 2: 2 (99 KB), 3-4: 1
    const int blocks = (int)max(1LL, min(chunks, 148LL * per_sm));
    *blocks_out = blocks;
    gram_kernel<R8, SPLIT3><<<blocks, kGramWarps * 32, smem, stream>>>(U, n, P, part);
}
template <int R8>
static void gram_dispatch(bool split3, cudaStream_t stream, const float* U, int n, long long P, double* part, int* blocks_out) {
    if (split3) gram_dispatch2<R8, true>(4, stream, U, n, P, part, blocks_out);
    else gram_dispatch2<R8, false>(4, stream, U, n, P, part, blocks_out);
}

int gram_launch(const float* U, int n, long long P, double eps, double* S, double* nrm, double* part, cudaStream_t stream) {
    // n ≤ 32 per call (two m16 tiles × four n8 tiles); larger clusters fall back in the caller
    if (n > 32 || n < 1) return -5;
    if (reinterpret_cast<uintptr_t>(U) & 15) return -5;   // bulk copies need a 16-byte aligned base (torch allocations are)
    const int r8 = (n + 7) / 8;
    const bool split3 = P < 65536;   // see the accuracy note in the header
    int blocks = 1;
    switch (r8) {
        case 1: gram_dispatch<1>(split3, stream, U, n, P, part, &blocks); break;
        case 2: gram_dispatch<2>(split3, stream, U, n, P, part, &blocks); break;
        case 3: gram_dispatch<3>(split3, stream, U, n, P, part, &blocks); break;
        default: gram_dispatch<4>(split3, stream, U, n, P, part, &blocks); break;
    }
    gram_finish_kernel<<<1, 1024, 0, stream>>>(part, blocks, n, r8, eps, S, nrm);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
