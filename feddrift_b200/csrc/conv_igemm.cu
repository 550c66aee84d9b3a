// conv_igemm — implicit-GEMM 2-D convolution on tcgen05 (forward and data-gradient), NHWC activations.
//
// Reference: cuDNN fp32 `nn.Conv2d` (fedml_api/model/cv/cnn.py:110-117, torchvision resnet18 blocks main_fedavg.py:219-223).
// Round 1 used an EXPLICIT im2col (9× activation blow-up through HBM) + GEMM and lost to cuDNN; here the im2col matrix never
// exists in memory: the producer warps gather the K-slice of every output pixel straight from the NHWC activation tensor
// into the shared-memory operand tile.
//
//   D[pixel, k_out] = Σ_{r,s,c} X[n, oy·st − pad + r, ox·st − pad + s, c] · W[k_out][r][s][c]        (mode 0, forward)
//   D[pixel, c_in ] = Σ_{r,s,k} dY[n, (iy + pad − r)/st, (ix + pad − s)/st, k] · W'[c_in][r][s][k]   (mode 1, dgrad: taps whose
//                                                                              source row/col is not an integer are skipped)
//
// Persistent, warp-specialised:
//   warps 0-3  PRODUCERS — one thread per output pixel of the 128-pixel tile: per K-chunk (one filter tap × CK channels) the
//              thread loads its pixel's CK fp32 channels (contiguous in NHWC; zero for padding), converts to bf16 and stores
//              them as 16-byte pieces into the no-swizzle K-major core-matrix layout tcgen05 reads; the same warps copy the
//              weight tile (bf16 [Kout][R][S][C], packed once per step by conv_pack_weights_kernel); a 4-stage mbarrier ring;
//   warp 4     single-thread tcgen05.mma issuer (UMMA 128×BN×16, accumulator double-buffered in TMEM);
//   warps 5-8  epilogue: tcgen05.ld → bias / ReLU → fp32 NHWC rows (each thread owns one pixel: 128-byte row segments).
// fp32 activations in / out (the networks keep fp32 master activations, like the reference), bf16 tensor-core operands,
// fp32 accumulation.  All waits are bounded (trap).
#include <algorithm>

#include "kernels.h"
#include "tc05.cuh"

namespace fdb {

namespace cv {
constexpr int BM = 128, STAGES = 4;
constexpr int kThreads = 9 * 32;
}

FDB_DEVICE uint64_t make_desc_kmajor_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
FDB_DEVICE uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int CK>
__global__ void __launch_bounds__(cv::kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvArgs a) {
    using namespace cv;
    constexpr int KC = CK / 8;                         // 16-byte k-cores per chunk
    constexpr uint32_t A_BYTES = KC * 16 * 128;        // [KC][16 m-cores][8 rows][16 B]
    constexpr uint32_t B_BYTES = KC * (BN / 8) * 128;  // [KC][BN/8 n-cores][8 rows][16 B]
    constexpr uint32_t LBO_A = 16 * 128, LBO_B = (BN / 8) * 128, SBO = 128;
    constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PQ = a.P * a.Q;
    const long long Mtot = (long long)a.N * PQ;
    const int m_tiles = (int)((Mtot + BM - 1) / BM), n_tiles = a.Kout / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int nCk = a.C / CK, nK = a.R * a.S * nCk;

    if (warp == 4 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 128); mbar_init(empty_bar + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 4) {
        // ===================================================== producers: thread m = output pixel m of the tile
        const int m = threadIdx.x;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const long long g = (long long)m_blk * BM + m;
            const bool row_ok = g < Mtot;
            const int n = row_ok ? (int)(g / PQ) : 0, rem = row_ok ? (int)(g % PQ) : 0;
            const int oy = rem / a.Q, ox = rem % a.Q;
            const __nv_bfloat16* wrow = a.wq + (size_t)n_blk * BN * a.R * a.S * a.C;
            for (int kidx = 0; kidx < nK; ++kidx, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int tap = kidx / nCk, c0 = (kidx - tap * nCk) * CK;
                const int r = tap / a.S, ss = tap - r * a.S;
                int iy, ix;
                bool ok = row_ok;
                if (a.mode == 0) {
                    iy = oy * a.stride - a.pad_h + r;
                    ix = ox * a.stride - a.pad_w + ss;
                } else {
                    const int ty = oy + a.pad_h - r, tx = ox + a.pad_w - ss;
                    ok = ok && ty >= 0 && tx >= 0 && (ty % a.stride) == 0 && (tx % a.stride) == 0;
                    iy = ty / a.stride; ix = tx / a.stride;
                }
                ok = ok && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
                // issue the global loads first (they do not depend on the ring), then wait for the stage
                uint32_t pk[CK / 2];
                if (ok) {
                    const float4* src = reinterpret_cast<const float4*>(a.x + (((size_t)n * a.H + iy) * a.W + ix) * a.C + c0);
#pragma unroll
                    for (int j = 0; j < CK / 4; ++j) {
                        const float4 f = __ldg(src + j);
                        pk[2 * j] = pack2_bf16(f.x, f.y);
                        pk[2 * j + 1] = pack2_bf16(f.z, f.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CK / 2; ++j) pk[j] = 0u;
                }
                constexpr int WCH = (BN * KC + 127) / 128;       // 16-byte weight pieces per thread
                uint4 wv[WCH];
#pragma unroll
                for (int q = 0; q < WCH; ++q) {
                    const int j = m + 128 * q;
                    if (j < BN * KC) {
                        const int nn = j / KC, kc = j % KC;
                        wv[q] = __ldg(reinterpret_cast<const uint4*>(wrow + ((size_t)nn * a.R * a.S + tap) * a.C + c0 + kc * 8));
                    }
                }
                mbar_wait(empty_bar + s, ph ^ 1);
                const uint32_t sa = smem_u32(smem_a + s * A_BYTES) + (uint32_t)((m >> 3) * 128 + (m & 7) * 16);
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) st_shared_v4(sa + kc * LBO_A, pk[4 * kc], pk[4 * kc + 1], pk[4 * kc + 2], pk[4 * kc + 3]);
                const uint32_t sb = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
                for (int q = 0; q < WCH; ++q) {
                    const int j = m + 128 * q;
                    if (j < BN * KC) {
                        const int nn = j / KC, kc = j % KC;
                        st_shared_v4(sb + (uint32_t)((kc * (BN / 8) + (nn >> 3)) * 128 + (nn & 7) * 16), wv[q].x, wv[q].y, wv[q].z, wv[q].w);
                    }
                }
                fence_proxy_async_smem();        // generic-proxy stores → visible to the tensor core (async proxy)
                mbar_arrive(full_bar + s);
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            // ===================================================== MMA issuer
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(tmem_empty + acc, aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_addr = tmem_base + acc * BN;
                for (int kidx = 0; kidx < nK; ++kidx, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar + s, ph);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < CK / 16; ++k)
                        umma_f16(d_addr, make_desc_kmajor_nosw(a_addr + k * 2 * LBO_A, LBO_A, SBO),
                                 make_desc_kmajor_nosw(b_addr + k * 2 * LBO_B, LBO_B, SBO), kIdesc, (kidx | k) != 0 ? 1u : 0u);
                    tcgen05_commit(empty_bar + s);
                }
                tcgen05_commit(tmem_full + acc);
            }
        }
    } else {
        // ===================================================== epilogue: warp quarter q owns TMEM lanes [32q, 32q+32) = tile rows
        const int q = warp & 3;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            mbar_wait(tmem_full + acc, aph);
            tcgen05_fence_after();
            const long long g = (long long)m_blk * BM + q * 32 + lane;
            float* out = a.y + (size_t)g * a.Kout + (size_t)n_blk * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, v);
                if (g < Mtot) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]); o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
                        if (a.bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n_blk * BN + c0 + j));
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4*>(out + c0 + j) = o;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + acc);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// fp32 OIHW weights [K][C][R][S] → bf16 tap-major tiles.  mode 0: out[k][r][s][c] (forward); mode 1: out[c][r][s][k] (dgrad)
__global__ void conv_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int K, int C, int R, int S, int mode) {
    const long long total = (long long)K * C * R * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        int inner, s, r, outer;
        if (mode == 0) { inner = (int)(t % C); t /= C; s = (int)(t % S); t /= S; r = (int)(t % R); outer = (int)(t / R);
                         out[i] = __float2bfloat16(w[(((size_t)outer * C + inner) * R + r) * S + s]); }
        else           { inner = (int)(t % K); t /= K; s = (int)(t % S); t /= S; r = (int)(t % R); outer = (int)(t / R);
                         out[i] = __float2bfloat16(w[(((size_t)inner * C + outer) * R + r) * S + s]); }
    }
}

int conv_pack_weights_launch(const float* w, void* out, int K, int C, int R, int S, int mode, cudaStream_t stream) {
    const long long total = (long long)K * C * R * S;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    conv_pack_weights_kernel<<<blocks, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), K, C, R, S, mode);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ======================================================================================================= weight gradient
// dW[k][r][s][c] = Σ_pixels dY[pixel][k] · X[gather(pixel, r, s)][c]   — a GEMM whose REDUCTION runs over the output pixels:
//   D[(tap, c) 128 rows × BN k-columns] += Aᵀ · B with A[j = (tap, c)][pixel] gathered from X (MN-major: 8 consecutive
//   channels of one pixel are one 16-byte piece) and B[k][pixel] = dY (MN-major as well: the NHWC rows are used as they are).
// grid = (row tiles × k tiles × pixel splits); every CTA accumulates its pixel range in TMEM and adds the tile into the fp32
// [K][R][S][C] gradient with coalesced red.global.add (the buffer is zeroed by the caller).  Same warp roles as above.
template <int BN>
__global__ void __launch_bounds__(cv::kThreads, 1) conv_wgrad_kernel(const __grid_constant__ ConvArgs a, int splits) {
    using namespace cv;
    constexpr int PK = 64, KCP = PK / 8;                  // pixels per stage, 8-pixel k-cores
    constexpr uint32_t A_BYTES = 16 * KCP * 128;          // [16 m-cores][8 k-cores][8 pixels][16 B]
    constexpr uint32_t B_BYTES = (BN / 8) * KCP * 128;
    constexpr uint32_t SBO = KCP * 128, LBO = 128;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)(BM >> 4) << 24);                                     // both operands MN-major
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int PQ = a.P * a.Q, RSC = a.R * a.S * a.C;
    const long long Mtot = (long long)a.N * PQ;
    const int m_tiles = (RSC + BM - 1) / BM, n_tiles = a.Kout / BN;
    int t = blockIdx.x;
    const int split = t % splits; t /= splits;
    const int m_blk = t % m_tiles, n_blk = t / m_tiles;
    const int chunks_total = (int)((Mtot + PK - 1) / PK);
    const int per = (chunks_total + splits - 1) / splits;
    const int ch_lo = split * per, ch_hi = min(chunks_total, ch_lo + per);
    const int nch = max(ch_hi - ch_lo, 0);

    if (warp == 4 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 128); mbar_init(empty_bar + s, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 4) {
        // ===================================================== producers: thread = (pixel kp of the 64-pixel chunk, half hm)
        const int tid = threadIdx.x, kp = tid & 63, hm = tid >> 6;
        for (int ci = 0; ci < nch; ++ci) {
            const int s = ci % STAGES;
            const uint32_t ph = (ci / STAGES) & 1;
            const long long g = (long long)(ch_lo + ci) * PK + kp;
            const bool row_ok = g < Mtot;
            const int n = row_ok ? (int)(g / PQ) : 0, rem = row_ok ? (int)(g % PQ) : 0;
            const int oy = rem / a.Q, ox = rem % a.Q;
            // A: 8 pieces of 8 channels: rows j = m_blk·128 + hm·64 + 8·i … (+8)
            uint4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = m_blk * BM + hm * 64 + 8 * i;
                av[i] = make_uint4(0u, 0u, 0u, 0u);
                if (row_ok && j < RSC) {
                    const int tap = j / a.C, c = j - tap * a.C;
                    const int r = tap / a.S, ss = tap - r * a.S;
                    const int iy = oy * a.stride - a.pad_h + r, ix = ox * a.stride - a.pad_w + ss;
                    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
                        const float4* src = reinterpret_cast<const float4*>(a.x + (((size_t)n * a.H + iy) * a.W + ix) * a.C + c);
                        const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
                        av[i] = make_uint4(pack2_bf16(f0.x, f0.y), pack2_bf16(f0.z, f0.w), pack2_bf16(f1.x, f1.y), pack2_bf16(f1.z, f1.w));
                    }
                }
            }
            // B: BN/16 pieces of 8 output channels of dY's row g: channels n_blk·BN + hm·(BN/2) + 8·i
            constexpr int NPB = BN / 16;
            uint4 bv[NPB];
#pragma unroll
            for (int i = 0; i < NPB; ++i) {
                bv[i] = make_uint4(0u, 0u, 0u, 0u);
                if (row_ok) {
                    const float4* src = reinterpret_cast<const float4*>(a.dy + (size_t)g * a.Kout + n_blk * BN + hm * (BN / 2) + 8 * i);
                    const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
                    bv[i] = make_uint4(pack2_bf16(f0.x, f0.y), pack2_bf16(f0.z, f0.w), pack2_bf16(f1.x, f1.y), pack2_bf16(f1.z, f1.w));
                }
            }
            mbar_wait(empty_bar + s, ph ^ 1);
            // MN-major no-swizzle: piece (mn-core i, pixel k) at i·SBO + (k/8)·LBO + (k%8)·16
            const uint32_t koff = (uint32_t)((kp >> 3) * LBO + (kp & 7) * 16);
            const uint32_t sa = smem_u32(smem_a + s * A_BYTES) + koff, sb = smem_u32(smem_b + s * B_BYTES) + koff;
#pragma unroll
            for (int i = 0; i < 8; ++i) st_shared_v4(sa + (uint32_t)(hm * 8 + i) * SBO, av[i].x, av[i].y, av[i].z, av[i].w);
#pragma unroll
            for (int i = 0; i < NPB; ++i) st_shared_v4(sb + (uint32_t)(hm * NPB + i) * SBO, bv[i].x, bv[i].y, bv[i].z, bv[i].w);
            fence_proxy_async_smem();
            mbar_arrive(full_bar + s);
        }
    } else if (warp == 4) {
        if (lane == 0) {
            for (int ci = 0; ci < nch; ++ci) {
                const int s = ci % STAGES;
                const uint32_t ph = (ci / STAGES) & 1;
                mbar_wait(full_bar + s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < PK / 16; ++k)
                    umma_f16(tmem_base, make_desc_kmajor_nosw(a_addr + k * 2 * LBO, LBO, SBO), make_desc_kmajor_nosw(b_addr + k * 2 * LBO, LBO, SBO),
                             kIdesc, (ci | k) != 0 ? 1u : 0u);
                tcgen05_commit(empty_bar + s);
            }
            tcgen05_commit(done_bar);
        }
    } else if (nch > 0) {
        // ===================================================== epilogue: lane = row j of the tile, reduce-add into dW[k][j]
        const int q = warp & 3;
        mbar_wait(done_bar, 0);
        tcgen05_fence_after();
        const int j = m_blk * BM + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (j < RSC) {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(a.dw + (size_t)(n_blk * BN + c0 + i) * RSC + j, __uint_as_float(v[i]));
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

template <int BN>
static int launch_wgrad(const ConvArgs& a, cudaStream_t stream) {
    using namespace cv;
    constexpr size_t smem = (size_t)STAGES * (16 * 8 * 128 + (BN / 8) * 8 * 128) + 256 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int RSC = a.R * a.S * a.C;
    const int tiles = ((RSC + BM - 1) / BM) * (a.Kout / BN);
    const long long Mtot = (long long)a.N * a.P * a.Q;
    const int chunks = (int)((Mtot + 63) / 64);
    int splits = std::max(1, std::min(chunks, (2 * sms + tiles - 1) / tiles));    // ~2 waves of CTAs
    conv_wgrad_kernel<BN><<<tiles * splits, kThreads, smem, stream>>>(a, splits);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// a.x = forward input (NHWC fp32), a.dy = output gradient [N·P·Q, Kout] fp32, a.dw = zeroed fp32 [Kout][R][S][C]
int conv_wgrad_launch(const ConvArgs& a, cudaStream_t stream) {
    if (a.C % 8 != 0 || a.Kout % 32 != 0 || a.N <= 0) return -5;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.dy) & 15)) return -6;
    const int bn = (a.Kout % 256 == 0) ? 256 : (a.Kout % 128 == 0) ? 128 : (a.Kout % 64 == 0) ? 64 : 32;
    if (bn == 256) return launch_wgrad<256>(a, stream);
    if (bn == 128) return launch_wgrad<128>(a, stream);
    if (bn == 64) return launch_wgrad<64>(a, stream);
    return launch_wgrad<32>(a, stream);
}

template <int BN, int CK>
static int launch_conv(const ConvArgs& a, cudaStream_t stream) {
    using namespace cv;
    constexpr int KC = CK / 8;
    constexpr size_t smem = (size_t)STAGES * (KC * 16 * 128 + KC * (BN / 8) * 128) + 256 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_igemm_kernel<BN, CK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long Mtot = (long long)a.N * a.P * a.Q;
    const int tiles = (int)((Mtot + BM - 1) / BM) * (a.Kout / BN);
    conv_igemm_kernel<BN, CK><<<std::min(tiles, sms), kThreads, smem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// picks the largest instantiated (BN | Kout, CK | C) tile; -5 when the shape is not supported (caller falls back)
int conv_igemm_launch(const ConvArgs& a, cudaStream_t stream) {
    if (a.C % 16 != 0 || a.Kout % 32 != 0 || a.N <= 0) return -5;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.y) & 15) || (reinterpret_cast<uintptr_t>(a.wq) & 15)) return -6;
    const int bn = (a.Kout % 256 == 0) ? 256 : (a.Kout % 128 == 0) ? 128 : (a.Kout % 64 == 0) ? 64 : 32;
    const int ck = (a.C % 64 == 0) ? 64 : (a.C % 32 == 0) ? 32 : 16;
#define FDB_CONV_CASE(B, K) if (bn == B && ck == K) return launch_conv<B, K>(a, stream);
    FDB_CONV_CASE(256, 64) FDB_CONV_CASE(256, 32) FDB_CONV_CASE(256, 16)
    FDB_CONV_CASE(128, 64) FDB_CONV_CASE(128, 32) FDB_CONV_CASE(128, 16)
    FDB_CONV_CASE(64, 64) FDB_CONV_CASE(64, 32) FDB_CONV_CASE(64, 16)
    FDB_CONV_CASE(32, 64) FDB_CONV_CASE(32, 32) FDB_CONV_CASE(32, 16)
#undef FDB_CONV_CASE
    return -5;
}

}  // namespace fdb
