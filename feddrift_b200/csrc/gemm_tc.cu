// gemm_tn: D[M,N] = act(A[M,K] · B[N,K]ᵀ + bias[N])   bf16 operands, fp32 accumulation in TMEM.
//
// Hand-written Blackwell GEMM used by TcLinear (fnn-MNIST 784→1568→10, CNN fc 9216→128, LSTM/classifier heads):
//   * PERSISTENT, warp-specialised: grid = min(#tiles, #SMs); every CTA walks tiles tile = blockIdx.x + i·gridDim.x
//     (M-fastest rasterisation so concurrently running CTAs share the same B panel in L2);
//   * warp 0 = TMA producer (cp.async.bulk.tensor.2d, 128B-swizzled 128×64 / BN×64 bf16 boxes) into a 4 / 6 / 8-stage (BN 256 / 128 / 64) smem ring
//     with full/empty mbarriers; warp 1 = single-thread tcgen05.mma issuer (UMMA 128×BN×16, BN ∈ {128, 256});
//     warp 2 owns the TMEM allocation; warps 4-7 = epilogue;
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 × BN fp32 columns): the epilogue of tile i (tcgen05.ld →
//     bias + ReLU + cast in registers → 128-byte row-segment stores) overlaps the MMAs of tile i+1
//     (tmem_full / tmem_empty mbarrier pair per buffer);
//   * split-K (grid.z) for skinny outputs (e.g. 512×128×9216): partial tiles are accumulated with fp32 atomics into a
//     zeroed output and bias/activation are applied by a small second kernel.
//   * operand layouts: each of A and B may be K-major ([rows, K], the "TN" form) or MN-major ([K, rows], rows
//     contiguous).  MN-major tiles are fetched as 64×64 TMA boxes (128-byte rows along M/N, 8-row swizzle atoms along
//     K) and described to the tensor core with the MN-major canonical layout (LBO = 8 KB between 64-wide M/N groups,
//     SBO = 1 KB between 8-row K groups, a_major/b_major bits of the instruction descriptor), so the backward GEMMs
//     dX = dY·W and dW = dYᵀ·X run directly on the row-major tensors autograd hands us — no transpose kernels.
//   * IMPLICIT-GEMM CONVOLUTION on the same mainloop (struct ConvIm below, ops/conv.py): the producer thread issues TMA *im2col*
//     loads (cp.async.bulk.tensor.4d…im2col) from the bf16 NHWC activation tensor — the TMA unit walks the output pixels with the
//     conv stride, applies the filter-tap offset and zero-fills the halo — for the forward, the data gradient (the forward
//     weight pack read MN-major, taps flipped; strided layers on zero-dilated dY) and the weight gradient (reduction over
//     pixels, im2col boxes as the MN-major B operand, cp.reduce.async.bulk adds into the channels_last gradient), optionally
//     GROUPED with one group per stacked (client, model) pair (sim/stacked.py).
// All waits are bounded (trap after 2 s) so a protocol bug faults the context instead of hanging the GPU.
// The reference's equivalent is eager `nn.Linear` + separate bias/ReLU kernels in fp32 on cuBLAS
// (fedml_api/model/fnn/fnn.py:11-15, cv/cnn.py:128-136).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc05.cuh"

namespace fdb {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kGemmThreads = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-7 epilogue
constexpr uint32_t kStageBytesA = BM * BK * 2;
struct GemmBatch { int batch, a_k0, a_kstride, b_k0, b_kstride; };
// Implicit-GEMM convolution modes: one operand is fetched with TMA *im2col* loads straight from the bf16 NHWC tensor.
//   mode 1 (forward / stride-1 data gradient): A[pixel, (tap, c)] — K-major 128-pixel × 64-channel boxes, one per k-block
//           (k-block kb = tap·cchunks + c-chunk); B = packed weights [Cout, R·S·C], K-major for the forward; the data gradient
//           reads THE SAME pack as an MN-major operand (b_mn: 64(cin) × 64(cout) boxes at column tap'·Cin, taps flipped), so
//           no transposed copy of the weights is ever made;
//   GROUPED convolution (one group per stacked (client, model) pair, sim/stacked.py): gb.batch = groups, M / N are per-group
//   sizes; group g reads channel chunk g·Cg + c of the NHWC tensor, weight rows g·N + n, and writes output columns g·N + n
//   (mode 1) or slice g of the 3-D gradient map [groups][Cout][R·S·C] (mode 2, which therefore clips rows ≥ Cout);
//   mode 2 (weight gradient): reduction over output pixels; A = dY [pixels, Cout] (MN-major tiled loads), B[(tap, c), pixel] —
//           MN-major 64-pixel × 64-channel im2col boxes, one per 64-wide column group of the N tile.
struct ConvIm { int mode, S, cchunks, ntaps, Q, PQ, stride, pad_h, pad_w, flip, bcols; };

// FDB_GEMM_DBG bit 3: CTA 0 accumulates SM-clock cycles per pipeline role / wait site (read back with gemm_debug_counters())
__device__ long long g_gemm_dbg[16];

template <int BN> struct GemmCfg {
    static constexpr uint32_t kStageBytesB = BN * BK * 2;
    static constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered accumulator (256 or 512 columns)
    static constexpr uint32_t kStagingBytes = 4 /*epilogue warps*/ * 2 /*double buffer*/ * 4096;   // 32 rows × 128 B per buffer
    // 192 KB of operand ring in every configuration: 4 × 48 KB (BN 256), 6 × 32 KB (BN 128), 8 × 24 KB (BN 64) — the narrow tiles
    // of the convolution modes have short K loops, a deeper ring keeps their TMA (im2col) loads far enough ahead
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr uint32_t kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    // c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, void* __restrict__ D, const float* __restrict__ bias, int M, int N, int K,
               int relu, int out_fp32, int splits, int tma_out, int a_mn, int b_mn, GemmBatch gb, ConvIm ci, int dbg) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * kStageBytesA;
    uint8_t* staging = smem + STAGES * (kStageBytesA + Cfg::kStageBytesB);   // 1024-aligned: [4 warps][2][32 rows × 128 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp index in a uniform register
    const int m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
    // batched mode (gb.batch > 1; both operands MN-major, M % 128 == 0, K % 64 == 0): batch bt multiplies the reduction rows
    // [a_k0 + bt·a_kstride, +K) of A' with [b_k0 + bt·b_kstride, +K) of B' into rows [bt·M, bt·M + M) of D — one launch for
    // the weight gradients dW_p = dG_pᵀ·H_p of every (client, model) pair
    const int tiles_per_batch = m_tiles * n_tiles;
    const int num_tiles = tiles_per_batch * gb.batch;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_per = (kb_total + splits - 1) / splits;
    const int kb_lo = blockIdx.z * kb_per, kb_hi = min(kb_total, kb_lo + kb_per);
    const int nkb = max(kb_hi - kb_lo, 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        if (tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (nkb > 0) {  // ===== TMA producer: the whole warp walks the loop (uniform control flow), one elected lane issues
            uint32_t it = 0;
            const bool prof = (dbg & 8) && blockIdx.x == 0;
            long long c_wait = 0, c_all0 = prof ? clock64() : 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int bt = tile / tiles_per_batch, rem = tile - bt * tiles_per_batch;
                const int m_blk = rem % m_tiles, n_blk = rem / m_tiles;
                const int ak = gb.a_k0 + bt * gb.a_kstride, bk = gb.b_k0 + bt * gb.b_kstride;   // 0 unless batched
                int cw = 0, chh = 0, cn = 0;          // im2col anchor of the tile's first pixel (mode 1)
                if (ci.mode == 1) {
                    const int m0 = m_blk * BM;
                    cn = m0 / ci.PQ;
                    const int rem2 = m0 - cn * ci.PQ, p = rem2 / ci.Q, q = rem2 - p * ci.Q;
                    cw = q * ci.stride - ci.pad_w;
                    chh = p * ci.stride - ci.pad_h;
                }
                int nvalid = BN / 64;                  // mode 2: 64-wide (tap, c-chunk) column groups of this N tile that exist
                if (ci.mode == 2) nvalid = max(0, min(BN / 64, ci.ntaps * ci.cchunks - n_blk * (BN / 64)));
                for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long w0 = prof ? clock64() : 0;
                    if (!(dbg & 16)) mbar_wait(empty_bar + s, ph ^ 1);   // bit 4 (with bit 0): free-running ring, no stage release
                    if (prof) c_wait += clock64() - w0;
                    uint8_t* sa = smem_a + s * kStageBytesA;
                    uint8_t* sb = smem_b + s * Cfg::kStageBytesB;
                    if (dbg & 1) { if (elect_one()) mbar_arrive(full_bar + s); continue; }   // FDB_GEMM_DBG bit 0: pipeline without operand loads (timing study)
                    if (!elect_one()) continue;
                    if (ci.mode == 1) {
                        mbar_expect_tx(full_bar + s, kStageBytesA + Cfg::kStageBytesB);
                        const int tap = kb / ci.cchunks, cc = kb - tap * ci.cchunks;
                        const int r = tap / ci.S, sx = tap - r * ci.S;
                        const int cg = bt * ci.cchunks * 64;        // first channel of this group in the NHWC tensor / first K row
                        tma_load_im2col_4d(&map_a, full_bar + s, sa, cg + cc * 64, cw, chh, cn, (uint16_t)sx, (uint16_t)r);
                        const int wtap = ci.flip ? ci.ntaps - 1 - tap : tap;
                        if (b_mn) {
#pragma unroll
                            for (int h = 0; h < BN / 64; ++h)
                                tma_load_2d(&map_b, full_bar + s, sb + h * 8192, wtap * ci.bcols + n_blk * BN + h * 64, cg + cc * BK);
                        } else {
                            tma_load_2d(&map_b, full_bar + s, sb, (wtap * ci.cchunks + cc) * BK, bt * N + n_blk * BN);
                        }
                        continue;
                    }
                    if (ci.mode == 2) {
                        mbar_expect_tx(full_bar + s, kStageBytesA + (uint32_t)nvalid * 8192u);
#pragma unroll
                        for (int h = 0; h < BM / 64; ++h) tma_load_2d(&map_a, full_bar + s, sa + h * 8192, bt * M + m_blk * BM + h * 64, kb * BK);
                        const int p0 = kb * BK, pn = p0 / ci.PQ, prem = p0 - pn * ci.PQ, pp = prem / ci.Q, pq = prem - pp * ci.Q;
                        for (int h = 0; h < nvalid; ++h) {
                            const int j = n_blk * (BN / 64) + h, tap = j / ci.cchunks, cc = j - tap * ci.cchunks;
                            const int r = tap / ci.S, sx = tap - r * ci.S;
                            tma_load_im2col_4d(&map_b, full_bar + s, sb + h * 8192, bt * ci.cchunks * 64 + cc * 64, pq * ci.stride - ci.pad_w,
                                               pp * ci.stride - ci.pad_h, pn, (uint16_t)sx, (uint16_t)r);
                        }
                        continue;
                    }
                    mbar_expect_tx(full_bar + s, kStageBytesA + Cfg::kStageBytesB);
                    if (a_mn) {   // [K, M] tensor: two 64(M)×64(K) boxes
#pragma unroll
                        for (int h = 0; h < BM / 64; ++h) tma_load_2d(&map_a, full_bar + s, sa + h * 8192, m_blk * BM + h * 64, ak + kb * BK);
                    } else {
                        tma_load_2d(&map_a, full_bar + s, sa, kb * BK, m_blk * BM);
                    }
                    if (b_mn) {
#pragma unroll
                        for (int h = 0; h < BN / 64; ++h) tma_load_2d(&map_b, full_bar + s, sb + h * 8192, n_blk * BN + h * 64, bk + kb * BK);
                    } else {
                        tma_load_2d(&map_b, full_bar + s, sb, kb * BK, n_blk * BN);
                    }
                }
            }
            if (prof && lane == 0) { g_gemm_dbg[0] = c_wait; g_gemm_dbg[1] = clock64() - c_all0; g_gemm_dbg[2] = it; }
        }
    } else if (warp == 1) {
        if (nkb > 0) {  // ===== MMA issuer: uniform loop over the whole warp, one elected lane issues tcgen05.mma / commit
            uint32_t it = 0, tl = 0;
            const bool prof = (dbg & 8) && blockIdx.x == 0;
            long long c_wfull = 0, c_wacc = 0, c_all0 = prof ? clock64() : 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                const long long a0 = prof ? clock64() : 0;
                mbar_wait(tmem_empty + acc, aph ^ 1);   // epilogue has drained this accumulator
                if (prof) c_wacc += clock64() - a0;
                tcgen05_fence_after();
                const uint32_t d_addr = tmem_base + acc * BN;
                const uint32_t idesc = Cfg::kIdesc | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
                const uint64_t desc_a0 = a_mn ? make_smem_desc_mn(smem_u32(smem_a)) : make_smem_desc(smem_u32(smem_a));
                const uint64_t desc_b0 = b_mn ? make_smem_desc_mn(smem_u32(smem_b)) : make_smem_desc(smem_u32(smem_b));
                const uint32_t a_kstep = a_mn ? (2048u >> 4) : ((UMMA_K * 2u) >> 4), b_kstep = b_mn ? (2048u >> 4) : ((UMMA_K * 2u) >> 4);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long f0 = prof ? clock64() : 0;
                    mbar_wait(full_bar + s, ph);
                    if (prof) c_wfull += clock64() - f0;
                    if (!(dbg & 32)) tcgen05_fence_after();
                    // stage-0 descriptors + (stage offset >> 4) in the 14-bit start-address field; per UMMA_K step the start address
                    // moves by 32 B (K-major: 16 elements along the row) or 2048 B (MN-major: 16 K-rows = two 1024-B atoms)
                    const uint64_t da0 = desc_a0 + (uint64_t)((s * kStageBytesA) >> 4);
                    const uint64_t db0 = desc_b0 + (uint64_t)((s * Cfg::kStageBytesB) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t da = da0 + (uint64_t)(k * a_kstep), db = db0 + (uint64_t)(k * b_kstep);
                            if (!(dbg & 2)) umma_f16(d_addr, da, db, idesc, (kb | k) != 0 ? 1u : 0u);   // bit 1: no MMAs
                        }
                        if (!(dbg & 16)) tcgen05_commit(empty_bar + s);  // frees the smem stage once these MMAs retire
                    }
                }
                if (elect_one()) tcgen05_commit(tmem_full + acc);    // accumulator complete → epilogue
            }
            if (prof && lane == 0) { g_gemm_dbg[3] = c_wfull; g_gemm_dbg[4] = c_wacc; g_gemm_dbg[5] = clock64() - c_all0; g_gemm_dbg[6] = tl; }
        }
    } else if (warp >= 4 && nkb > 0) {
        // ===== epilogue: warp (4+q) owns TMEM lanes [32q, 32q+32) == output rows of the tile
        const int q = warp - 4;
        uint32_t tl = 0, chunk_it = 0;
        const bool prof = (dbg & 8) && blockIdx.x == 0 && q == 0;
        long long c_wtm = 0, c_all0 = prof ? clock64() : 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int bt = tile / tiles_per_batch, rem = tile - bt * tiles_per_batch;
            const int m_blk = rem % m_tiles, n_blk = rem / m_tiles;
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            const long long e0 = prof ? clock64() : 0;
            mbar_wait(tmem_full + acc, aph);
            if (prof) c_wtm += clock64() - e0;
            tcgen05_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            if (tma_out) {
                // Coalesced path: the warp's 32-row slab goes TMEM → registers (bias/ReLU/cast) → 128B-swizzled smem
                // staging (conflict-free 16-byte stores) → ONE TMA tile store (or fp32 reduce-add for split-K) per
                // 128-byte column chunk.  Staging is double-buffered per warp; TMA clips the M/N edges.
                const int cols_per_chunk = out_fp32 ? 32 : 64;
                // batched GEMM outputs are stacked along the rows of D; grouped convolutions write column block g·N (mode 1) or
                // slice g of the 3-D gradient map (mode 2)
                const int row0 = (ci.mode ? 0 : bt * M) + m_blk * BM + q * 32;
                const int colg = ci.mode == 1 ? bt * N : 0;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += cols_per_chunk, ++chunk_it) {
                    if (n_blk * BN + c0 >= N) break;   // warp-uniform
                    const int col0 = colg + n_blk * BN + c0;
                    uint8_t* buf = staging + (q * 2 + (chunk_it & 1)) * 4096;
                    if (chunk_it >= 2) tma_store_wait_read<1>();   // (all lanes; only the electing lane owns bulk groups) the store that last read this buffer is done
                    __syncwarp();
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0;
                    const uint32_t sbase = smem_u32(buf) + lane * 128;
                    tmem_ld_32x32(taddr, v);
                    if (out_fp32) {
                        if (splits == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float x = __uint_as_float(v[j]);
                                if (bias) x += (col0 - colg + j < N) ? __ldg(bias + col0 + j) : 0.f;
                                if (relu) x = fmaxf(x, 0.f);
                                v[j] = __float_as_uint(x);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            st_shared_v4(sbase + ((c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    } else {
                        uint32_t w[32];
                        tmem_ld_32x32(taddr + 32, w);
                        uint32_t pk[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float x0 = __uint_as_float(v[j]), x1 = __uint_as_float(v[j + 1]);
                            float y0 = __uint_as_float(w[j]), y1 = __uint_as_float(w[j + 1]);
                            if (bias) {
                                x0 += (col0 + j < N) ? __ldg(bias + col0 + j) : 0.f;
                                x1 += (col0 + j + 1 < N) ? __ldg(bias + col0 + j + 1) : 0.f;
                                y0 += (col0 + 32 + j < N) ? __ldg(bias + col0 + 32 + j) : 0.f;
                                y1 += (col0 + 33 + j < N) ? __ldg(bias + col0 + 33 + j) : 0.f;
                            }
                            if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                            __nv_bfloat162 a = __floats2bfloat162_rn(x0, x1), b = __floats2bfloat162_rn(y0, y1);
                            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&a);
                            pk[16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&b);
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            st_shared_v4(sbase + ((c ^ (lane & 7)) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (!(dbg & 4) && elect_one()) {   // bit 2: no output stores
                        if (ci.mode == 2) tma_reduce_add_3d(&map_d, buf, col0, row0, bt);   // wgrad always accumulates into slice g
                        else if (splits > 1) tma_reduce_add_2d(&map_d, buf, col0, row0);
                        else tma_store_2d(&map_d, buf, col0, row0);
                        tma_store_commit();
                    }
                }
            } else {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int col0 = n_blk * BN + c0;
                if (col0 >= N) break;   // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, v);
                if (row < M) {
                    if (splits > 1) {
                        float* out = reinterpret_cast<float*>(D) + (size_t)row * N + col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) atomicAdd(out + j, __uint_as_float(v[j]));
                    } else if (out_fp32) {
                        float* out = reinterpret_cast<float*>(D) + (size_t)row * N + col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (col0 + j < N) {
                                float x = __uint_as_float(v[j]);
                                if (bias) x += __ldg(bias + col0 + j);
                                if (relu) x = fmaxf(x, 0.f);
                                out[j] = x;
                            }
                        }
                    } else {
                        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(D) + (size_t)row * N + col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (col0 + j < N) {
                                float x = __uint_as_float(v[j]);
                                if (bias) x += __ldg(bias + col0 + j);
                                if (relu) x = fmaxf(x, 0.f);
                                out[j] = __float2bfloat16(x);
                            }
                        }
                    }
                }
            }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + acc);   // 4 epilogue warps → the MMA warp may reuse the buffer
        }
        if (prof && lane == 0) { g_gemm_dbg[7] = c_wtm; g_gemm_dbg[8] = clock64() - c_all0; }
        if (tma_out) tma_store_wait_all();   // smem must outlive the in-flight bulk stores (a no-op for lanes without bulk groups)
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
    }
}

// bias + activation (+ cast) pass for split-K outputs
__global__ void bias_act_kernel(const float* __restrict__ acc, void* __restrict__ D, const float* __restrict__ bias, long long MN, int N,
                                int relu, int out_fp32) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < MN; i += (long long)gridDim.x * blockDim.x) {
        float x = acc[i];
        if (bias) x += bias[i % N];
        if (relu) x = fmaxf(x, 0.f);
        if (out_fp32) reinterpret_cast<float*>(D)[i] = x;
        else reinterpret_cast<__nv_bfloat16*>(D)[i] = __float2bfloat16(x);
    }
}

// ---------------------------------------------------------------- host side: tensor maps via the driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rows, int cols /*K*/, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// exported for conv_igemm.cu: K-major 128B-swizzled map of a bf16 [rows, cols] matrix with (64 × box_rows) boxes
int make_kmajor_sw128_map(void* map_out, const void* base, int rows, int cols, int box_rows) {
    return make_map(reinterpret_cast<CUtensorMap*>(map_out), base, rows, cols, box_rows);
}

// output map: 32-row × 128-byte boxes (32 fp32 or 64 bf16 columns), 128B swizzle — what one epilogue warp stages per chunk.
// Returns 0 and sets *ok = 1 when D qualifies for TMA stores (16-byte aligned base and row pitch).
static int make_out_map(CUtensorMap* map, void* D, int M, int N, int out_fp32, int* ok) {
    *ok = 0;
    memset(map, 0, sizeof(*map));
    const size_t es = out_fp32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(D) & 15) || ((size_t)N * es) % 16 != 0) return 0;
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * es};
    cuuint32_t box[2] = {(cuuint32_t)(out_fp32 ? 32 : 64), 32u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, out_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -2;
    *ok = 1;
    return 0;
}

// MN-major operand X'[K, rows] (rows contiguous): 64×64 boxes, 128-byte rows
static int make_map_mn(CUtensorMap* map, const void* base, int rows, int K) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)rows * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// im2col map over a bf16 NHWC tensor [N, H, W, C]: boxes of `pixels` anchor pixels × 64 channels, 128B swizzle; the bounding
// box of anchors is the set of output positions of an R×S filter with symmetric padding, walked with the conv stride
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn get_encode_im2col() {
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(ptr);
    }
    return fn;
}
static int make_im2col_map(CUtensorMap* map, const void* base, int N, int H, int W, int C, int R, int S, int pad_h, int pad_w, int stride,
                           int pixels) {
    EncodeIm2colFn enc = get_encode_im2col();
    if (!enc) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad_w, -pad_h};
    int upper[2] = {pad_w - (S - 1), pad_h - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64u, (cuuint32_t)pixels,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, void* D, const float* bias, int M, int N, int K, int relu,
                       int out_fp32, int splits, int sms, int a_mn, int b_mn, cudaStream_t stream, GemmBatch gb = GemmBatch{1, 0, 0, 0, 0},
                       ConvIm ci = ConvIm{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, const CUtensorMap* md_conv = nullptr) {
    CUtensorMap md;
    int tma_out = 0;
    if (md_conv) { md = *md_conv; tma_out = 1; }     // convolution modes bring their own output map (grouped columns / 3-D gradient)
    else if (make_out_map(&md, D, M * gb.batch, N, out_fp32, &tma_out) != 0) return -7;
    if (gb.batch > 1 && !tma_out) return -9;
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess) return -3;
        attr_set = true;
    }
    static const int dbg = getenv("FDB_GEMM_DBG") ? atoi(getenv("FDB_GEMM_DBG")) : 0;   // timing-study switches, results are garbage
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * gb.batch;
    dim3 grid(min(tiles, max(1, sms / splits)), 1, splits);
    gemm_tn_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ma, mb, md, D, bias, M, N, K, relu, out_fp32, splits, tma_out, a_mn, b_mn, gb, ci, dbg);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int gemm_debug_counters(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_gemm_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -4;
}

// number of K-splits gemm_launch will use for this problem (callers that want a bf16 output pre-allocate the fp32
// accumulation workspace with their own allocator when this is > 1)
int gemm_split_count(int M, int N, int K) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int bn = (N > 128) ? 256 : 128;
    const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    const int kb_total = (K + BK - 1) / BK;
    int splits = 1;
    if (tiles * 4 <= sms && kb_total >= 16) {   // too few output tiles to fill the machine and a long K
        splits = min(min(sms / tiles, kb_total / 4), 32);
        if (splits < 2) splits = 1;
    }
    return splits;
}

int gemm_launch(const void* A, const void* B, void* D, const float* bias, int M, int N, int K, int a_mn, int b_mn, int relu, int out_fp32,
                cudaStream_t stream, float* splitk_ws) {
    // TMA global strides must be multiples of 16 B: the contiguous extent of each operand must be a multiple of 8 bf16
    if (M <= 0 || N <= 0 || K <= 0) return -5;
    if ((a_mn ? M : K) % 8 != 0 || (b_mn ? N : K) % 8 != 0) return -5;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -6;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int bn = (N > 128) ? 256 : 128;
    CUtensorMap ma, mb;
    if ((a_mn ? make_map_mn(&ma, A, M, K) : make_map(&ma, A, M, K, BM)) != 0) return -7;
    if ((b_mn ? make_map_mn(&mb, B, N, K) : make_map(&mb, B, N, K, bn)) != 0) return -7;
    const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    const int kb_total = (K + BK - 1) / BK;
    (void)tiles; (void)kb_total;
    const int splits = gemm_split_count(M, N, K);
    if (splits > 1) {
        float* acc = nullptr;
        const size_t bytes = (size_t)M * N * sizeof(float);
        bool own = false;
        if (out_fp32) acc = reinterpret_cast<float*>(D);
        else if (splitk_ws) acc = splitk_ws;     // caller's allocator (torch caching allocator: no driver call on the hot path)
        else { if (cudaMallocAsync(&acc, bytes, stream) != cudaSuccess) return -8; own = true; }
        cudaMemsetAsync(acc, 0, bytes, stream);
        int rc = (bn == 256) ? launch_gemm<256>(ma, mb, acc, nullptr, M, N, K, 0, 1, splits, sms, a_mn, b_mn, stream)
                             : launch_gemm<128>(ma, mb, acc, nullptr, M, N, K, 0, 1, splits, sms, a_mn, b_mn, stream);
        if (rc != 0) return rc;
        if (bias || relu || !out_fp32) {
            const long long MN = (long long)M * N;
            bias_act_kernel<<<(int)min((MN + 255) / 256, 148LL * 8), 256, 0, stream>>>(acc, D, bias, MN, N, relu, out_fp32);
        }
        if (own) cudaFreeAsync(acc, stream);
        return cudaGetLastError() == cudaSuccess ? 0 : -4;
    }
    return (bn == 256) ? launch_gemm<256>(ma, mb, D, bias, M, N, K, relu, out_fp32, 1, sms, a_mn, b_mn, stream)
                       : launch_gemm<128>(ma, mb, D, bias, M, N, K, relu, out_fp32, 1, sms, a_mn, b_mn, stream);
}

// Batched dW-style GEMM: D[bt·M + m, n] = Σ_k A'[a_k0 + bt·a_kstride + k, m] · B'[b_k0 + bt·b_kstride + k, n]   (k < K), fp32 out.
// A' is [a_rows_total, M] and B' is [b_rows_total, N] (bf16, row-major = MN-major operands).
int gemm_batched_mn_launch(const void* A, const void* B, float* D, int M, int N, int K, int batch, int a_rows_total, int b_rows_total,
                           int a_k0, int a_kstride, int b_k0, int b_kstride, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return -5;
    if (M % BM != 0 || K % BK != 0 || M % 8 != 0 || N % 8 != 0) return -5;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -6;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int bn = (N > 128) ? 256 : 128;
    CUtensorMap ma, mb;
    if (make_map_mn(&ma, A, M, a_rows_total) != 0) return -7;
    if (make_map_mn(&mb, B, N, b_rows_total) != 0) return -7;
    GemmBatch gb{batch, a_k0, a_kstride, b_k0, b_kstride};
    return (bn == 256) ? launch_gemm<256>(ma, mb, D, nullptr, M, N, K, 0, 1, 1, sms, 1, 1, stream, gb)
                       : launch_gemm<128>(ma, mb, D, nullptr, M, N, K, 0, 1, 1, sms, 1, 1, stream, gb);
}

static int launch_bn(int bn, const CUtensorMap& ma, const CUtensorMap& mb, void* D, const float* bias, int M, int N, int K, int relu,
                     int out_fp32, int splits, int sms, int a_mn, int b_mn, cudaStream_t stream, const ConvIm& ci, int groups,
                     const CUtensorMap& md) {
    const GemmBatch gb{groups, 0, 0, 0, 0};
    if (bn == 256) return launch_gemm<256>(ma, mb, D, bias, M, N, K, relu, out_fp32, splits, sms, a_mn, b_mn, stream, gb, ci, &md);
    if (bn == 128) return launch_gemm<128>(ma, mb, D, bias, M, N, K, relu, out_fp32, splits, sms, a_mn, b_mn, stream, gb, ci, &md);
    return launch_gemm<64>(ma, mb, D, bias, M, N, K, relu, out_fp32, splits, sms, a_mn, b_mn, stream, gb, ci, &md);
}

// 3-D fp32 output map of the weight-gradient GEMM: [groups][Cout][R·S·C] with `gstride` floats between the groups' gradients
// (the flat gradient rows of the stacked pairs); 32-row × 32-column × 1 boxes, rows ≥ Cout are clipped
static int make_wgrad_map(CUtensorMap* map, float* base, int groups, int cout, int rsc, long long gstride) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)rsc, (cuuint64_t)cout, (cuuint64_t)groups};
    cuuint64_t strides[2] = {(cuuint64_t)rsc * 4, (cuuint64_t)(groups > 1 ? gstride : (long long)rsc * cout) * 4};
    cuuint32_t box[3] = {32u, 32u, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// Implicit-GEMM convolution on the GEMM mainloop (TMA im2col producer), optionally GROUPED (groups > 1: one group per stacked
// (client, model) pair; C and Cout are PER-GROUP channel counts, the tensors hold groups·C / groups·Cout channels).
// xb: bf16 NHWC [N, H, W, groups·C] (C % 64 == 0), wq: bf16 weights [groups·Cout, R·S·C] (the channels_last storage of the
// parameters, cast), y: fp32 NHWC [N, P, Q, groups·Cout].  Square filters, symmetric padding.
//   dgrad = 0: forward, y = act(conv(x, w) + bias).
//   dgrad = 1: stride-1 data gradient: xb = dY [N, P, Q, groups·Cout_fwd] (C = Cout_fwd), wq = the SAME forward pack
//              [groups·Cout_fwd, R·S·Cin_fwd], `Cout` = Cin_fwd, pad = R-1-pad_fwd; the pack is read MN-major with flipped taps.
int conv_tma_fwd_launch(const void* xb, const void* wq, float* y, const float* bias, int N, int H, int W, int C, int Cout, int R, int S, int P,
                        int Q, int pad, int stride, int dgrad, int relu, int groups, cudaStream_t stream) {
    if (C % 64 != 0 || Cout % 8 != 0 || R != S || N <= 0 || groups < 1) return -5;
    if (groups > 1 && Cout % 32 != 0) return -5;       // a 32-column output chunk must not straddle two groups
    if ((reinterpret_cast<uintptr_t>(xb) & 15) || (reinterpret_cast<uintptr_t>(wq) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return -6;
    const long long Mll = (long long)N * P * Q;
    if (Mll >= (1LL << 31) - 256 || (long long)groups * C >= (1LL << 31)) return -5;
    const int M = (int)Mll, K = R * S * C;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = (M + BM - 1) / BM;
    // widest N tile that still gives every SM a tile; otherwise the narrowest and split K over the taps
    int bn = Cout <= 64 ? 64 : 128;
    if (Cout >= 256 && (long long)m_tiles * ((Cout + 255) / 256) * groups >= sms) bn = 256;
    const long long tiles = (long long)m_tiles * ((Cout + bn - 1) / bn) * groups;
    const int kb_total = K / BK;
    int splits = 1;
    if (tiles * 2 <= sms && kb_total >= 8) splits = std::max(1, std::min(std::min(sms / (int)tiles, kb_total / 4), 16));
    CUtensorMap ma, mb, md;
    int ok = 0;
    if (make_im2col_map(&ma, xb, N, H, W, groups * C, R, S, pad, pad, stride, BM) != 0) return -7;
    if (dgrad) { if (make_map_mn(&mb, wq, R * S * Cout, groups * C) != 0) return -7; }   // [groups·Cout_fwd rows (K), R·S·Cin_fwd contiguous]
    else if (make_map(&mb, wq, groups * Cout, K, bn) != 0) return -7;
    if (make_out_map(&md, y, M, groups * Cout, 1, &ok) != 0 || !ok) return -7;
    const ConvIm ci{1, S, C / 64, R * S, Q, P * Q, stride, pad, pad, dgrad, Cout};
    if (splits > 1) {
        cudaMemsetAsync(y, 0, (size_t)M * groups * Cout * sizeof(float), stream);
        int rc = launch_bn(bn, ma, mb, y, nullptr, M, Cout, K, 0, 1, splits, sms, 0, dgrad, stream, ci, groups, md);
        if (rc != 0) return rc;
        if (bias || relu) {
            const long long MN = (long long)M * groups * Cout;
            bias_act_kernel<<<(int)std::min<long long>((MN + 255) / 256, 148LL * 8), 256, 0, stream>>>(y, y, bias, MN, groups * Cout, relu, 1);
        }
        return cudaGetLastError() == cudaSuccess ? 0 : -4;
    }
    return launch_bn(bn, ma, mb, y, bias, M, Cout, K, relu, 1, 1, sms, 0, dgrad, stream, ci, groups, md);
}

// Weight gradient: dw[g][Cout, (r, s, c)] (fp32) += Σ_pixels dY[pixel, g·Cout + k] · X[gather(pixel, r, s), g·C + c] — the tiles are
// REDUCE-ADDED (cp.reduce.async.bulk) into the buffer, which is the channels_last storage of the parameters' gradients: the
// flat gradient rows of the federated executors (`gstride` floats between the groups' segments), or a zeroed tensor.
// xb: bf16 NHWC activations [N, H, W, groups·C], dyb: bf16 [N·P·Q, groups·Cout].
int conv_tma_wgrad_launch(const void* xb, const void* dyb, float* dw_ohwi, int N, int H, int W, int C, int Cout, int R, int S, int P, int Q,
                          int pad, int stride, int groups, long long gstride, cudaStream_t stream) {
    if (C % 64 != 0 || Cout % 8 != 0 || R != S || N <= 0 || groups < 1) return -5;
    if ((reinterpret_cast<uintptr_t>(xb) & 15) || (reinterpret_cast<uintptr_t>(dyb) & 15) || (reinterpret_cast<uintptr_t>(dw_ohwi) & 15)) return -6;
    if (groups > 1 && (gstride % 4 != 0 || gstride < (long long)Cout * R * S * C)) return -6;
    const long long Kll = (long long)N * P * Q;
    if (Kll >= (1LL << 31) - 256) return -5;
    const int Kpix = (int)Kll, RSC = R * S * C;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int bn = 128;
    const long long tiles = (long long)((Cout + BM - 1) / BM) * ((RSC + bn - 1) / bn) * groups;
    const int kb_total = (Kpix + BK - 1) / BK;
    int splits = tiles >= sms ? 1 : std::max(1, std::min(std::min(sms / (int)tiles, kb_total / 2), 64));
    CUtensorMap ma, mb, md;
    if (make_map_mn(&ma, dyb, groups * Cout, Kpix) != 0) return -7;
    if (make_im2col_map(&mb, xb, N, H, W, groups * C, R, S, pad, pad, stride, BK) != 0) return -7;
    if (make_wgrad_map(&md, dw_ohwi, groups, Cout, RSC, gstride) != 0) return -7;
    const ConvIm ci{2, S, C / 64, R * S, Q, P * Q, stride, pad, pad, 0, 0};
    return launch_bn(bn, ma, mb, dw_ohwi, nullptr, Cout, RSC, Kpix, 0, 1, splits, sms, 1, 1, stream, ci, groups, md);
}

int gemm_tn_launch(const void* A, const void* B, void* D, const float* bias, int M, int N, int K, int relu, int out_fp32,
                   cudaStream_t stream) {
    return gemm_launch(A, B, D, bias, M, N, K, 0, 0, relu, out_fp32, stream, nullptr);
}

}  // namespace fdb
