// conv_igemm — implicit-GEMM 2-D convolution on tcgen05 with SOFTWARE-GATHER producers (forward, data gradient, weight gradient),
// NHWC activations.  These kernels cover the layers whose channel counts are multiples of 32 but not 64 (the MNIST CNN's 32→64
// convolution); everything with Cin % 64 == 0 runs on the GEMM mainloop with the TMA-im2col producer (gemm_tc.cu conv modes,
// 2–4× faster).  The file also holds the small helpers both paths share (bf16 casts, the [K][RS][C] → [C][RS][K] transpose).
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
// Persistent, warp-specialised (416 threads):
//   warps 0-7  PRODUCERS (two groups of 4 warps working on alternate stages) — one thread per output pixel of the 128-pixel
//              tile: per K-chunk (one filter tap × CK channels) the thread loads its pixel's CK fp32 channels (contiguous in
//              NHWC; zero for padding; per-pixel origins from a shared-memory row table), converts to bf16 and stores them as
//              16-byte pieces into the no-swizzle K-major core-matrix layout tcgen05 reads; the weight tile (bf16
//              [Kout][R][S][C] = the cast channels_last parameter) arrives by TMA (CK = 64) or is copied by the same warps;
//              a 4-stage mbarrier ring;
//   warp 8     tcgen05.mma issuer: warp-uniform loop, one elected lane issues (UMMA 128×BN×16, accumulator double-buffered in TMEM);
//   warps 9-12 epilogue: tcgen05.ld → bias / ReLU → fp32 NHWC rows (each thread owns one pixel: 128-byte row segments).
// fp32 activations in / out (the networks keep fp32 master activations, like the reference), bf16 tensor-core operands,
// fp32 accumulation.  All waits are bounded (trap).
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc05.cuh"

namespace fdb {

namespace cv {
constexpr int BM = 128, STAGES = 4;
constexpr int kThreads = 9 * 32;          // wgrad: 4 producer warps + issuer + 4 epilogue warps
constexpr int kFwdThreads = 13 * 32;      // forward / dgrad: 2 producer groups of 4 warps + issuer + 4 epilogue warps
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

template <int BN, int CK, int MODE>
__global__ void __launch_bounds__(cv::kFwdThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvArgs a, const __grid_constant__ CUtensorMap map_w,
                                                                          int tma_w, int splits) {
    using namespace cv;
    constexpr int KC = CK / 8;                         // 16-byte k-cores per chunk
    // operand tiles, no-swizzle K-major: [k-core][row-core][8 rows][16 B]; the A k-core stride is padded by 16 B so that the
    // 8 k-cores a half-warp writes for one pixel land in different banks
    constexpr uint32_t LBO_A = 16 * 128 + 16, LBO_B = (BN / 8) * 128, SBO = 128;
    constexpr uint32_t A_BYTES = KC * LBO_A, B_BYTES = KC * LBO_B;
    constexpr uint32_t A_STRIDE = (A_BYTES + 127) & ~127u;
    constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem;                               // weight tiles first: the TMA path (128B swizzle) needs 1024-byte alignment
    uint8_t* smem_a = smem + STAGES * B_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_a + STAGES * A_STRIDE);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    int2* rinfo = reinterpret_cast<int2*>(tmem_ptr_smem + 4);      // [128] per-row gather origin of the current tile

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
    const int PQ = a.P * a.Q;
    const long long Mtot = (long long)a.N * PQ;
    const int m_tiles = (int)((Mtot + BM - 1) / BM), n_tiles = a.Kout / BN;
    // split-K (deep layers: few output pixels, long reductions): work item = (tile, split); split i reduces k-chunks
    // [i·kper, (i+1)·kper) and adds its partial tile into the zero-initialised output with red.global.add
    const int num_tiles = m_tiles * n_tiles * splits;
    const int nCk = a.C / CK, nKall = a.R * a.S * nCk;
    const int kper = (nKall + splits - 1) / splits;

    const bool use_tma = (CK == 64) && tma_w != 0;        // weight tile by ONE TMA box per stage (K-major, 128B swizzle)
    if (warp == 8 && lane == 0) {
        if (use_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, use_tma ? 129 : 128); mbar_init(empty_bar + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 8) {
        // ===================================================== producers: two groups of 128 threads fill alternate stages.
        // Loads are COALESCED: the CK/4 float4 of one pixel are read by CK/4 consecutive lanes (a warp reads whole pixels),
        // each lane converts its 4 channels to bf16 and stores the 8-byte half of the pixel's 16-byte k-core piece.
        const int grp = warp >> 2, t = threadIdx.x & 127;
        constexpr int F4R = CK / 4;                       // float4 per pixel row and chunk
        constexpr int NI = F4R;                           // = (128 rows · F4R) / 128 threads
        uint32_t it = 0;
        for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
            const int tile = work / splits, sp = work - tile * splits;
            const int k_lo = sp * kper, k_hi = min(nKall, k_lo + kper);
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            asm volatile("bar.sync 1, 256;" ::: "memory");        // everybody is done with the previous tile's row table
            if (grp == 0) {
                // row table: .x = pixel index of the gather origin (n, y0, x0) in the source tensor (may be negative),
                // .y = y0 in the high / x0 in the low 16 bits (biased by 0x4000); rows past the end get an origin far outside
                const long long g = (long long)m_blk * BM + t;
                int2 ri = make_int2(0, 0);                          // y0 = x0 = -0x4000: every tap is out of range
                if (g < Mtot) {
                    const int n = (int)(g / PQ), rem = (int)(g % PQ);
                    const int oy = rem / a.Q, ox = rem % a.Q;
                    const int y0 = (MODE == 0) ? oy * a.stride - a.pad_h : oy + a.pad_h;
                    const int x0 = (MODE == 0) ? ox * a.stride - a.pad_w : ox + a.pad_w;
                    ri = (MODE == 0) ? make_int2((n * a.H + y0) * a.W + x0, ((y0 + 0x4000) << 16) | (x0 + 0x4000))
                                     : make_int2(n, ((y0 + 0x4000) << 16) | (x0 + 0x4000));
                }
                rinfo[t] = ri;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const __nv_bfloat16* wrow = a.wq + (size_t)n_blk * BN * a.R * a.S * a.C;
            const uint32_t rinfo_s = smem_u32(rinfo);
            const int sh = (a.stride == 2) ? 1 : 0;               // dgrad supports stride 1 and 2 (host-checked)
            for (int kidx = k_lo; kidx < k_hi; ++kidx, ++it) {
                if ((int)(it & 1) != grp) continue;
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int tap = kidx / nCk, c0 = (kidx - tap * nCk) * CK;
                const int r = tap / a.S, ss = tap - r * a.S;
                // (1) row table → registers, (2) addresses + all global loads back to back, (3) convert + store
                int2 rr[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int m = (i * 128 + t) / F4R;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(rr[i].x), "=r"(rr[i].y) : "r"(rinfo_s + (uint32_t)m * 8u));
                }
                float4 f[NI];
                const int tapoff = r * a.W + ss;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int cq = (i * 128 + t) % F4R;
                    const int y0 = (rr[i].y >> 16) - 0x4000, x0 = (rr[i].y & 0xFFFF) - 0x4000;
                    bool ok;
                    int pix;
                    if (MODE == 0) {
                        ok = (unsigned)(y0 + r) < (unsigned)a.H && (unsigned)(x0 + ss) < (unsigned)a.W;
                        pix = rr[i].x + tapoff;
                    } else {
                        const int ty = y0 - r, tx = x0 - ss;
                        ok = ty >= 0 && tx >= 0 && ((ty | tx) & sh) == 0 && (ty >> sh) < a.H && (tx >> sh) < a.W;
                        pix = (rr[i].x * a.H + (ty >> sh)) * a.W + (tx >> sh);
                    }
                    f[i] = ok ? __ldg(reinterpret_cast<const float4*>(a.x + (size_t)(unsigned)pix * a.C + c0) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                constexpr int WCH = (BN * KC + 127) / 128;       // 16-byte weight pieces per thread
                uint4 wv[WCH];
                if (!use_tma) {
#pragma unroll
                    for (int q = 0; q < WCH; ++q) {
                        const int j = t + 128 * q;
                        if (j < BN * KC) {
                            const int nn = j / KC, kc = j % KC;
                            wv[q] = __ldg(reinterpret_cast<const uint4*>(wrow + ((size_t)nn * a.R * a.S + tap) * a.C + c0 + kc * 8));
                        }
                    }
                }
                mbar_wait(empty_bar + s, ph ^ 1);
                if (use_tma && t == 0) {       // the stage is free: one elected thread arms the tx count and fires the weight box
                    mbar_expect_tx(full_bar + s, B_BYTES);
                    tma_load_2d(&map_w, full_bar + s, smem_b + s * B_BYTES, tap * a.C + c0, n_blk * BN);
                }
                const uint32_t sa = smem_u32(smem_a + s * A_STRIDE);
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int q = i * 128 + t, m = q / F4R, cq = q % F4R;
                    const uint32_t addr = sa + (uint32_t)((cq >> 1) * LBO_A + (m >> 3) * 128 + (m & 7) * 16 + (cq & 1) * 8);
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack2_bf16(f[i].x, f[i].y)), "r"(pack2_bf16(f[i].z, f[i].w)) : "memory");
                }
                if (!use_tma) {
                    const uint32_t sb = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
                    for (int q = 0; q < WCH; ++q) {
                        const int j = t + 128 * q;
                        if (j < BN * KC) {
                            const int nn = j / KC, kc = j % KC;
                            st_shared_v4(sb + (uint32_t)(kc * LBO_B + (nn >> 3) * 128 + (nn & 7) * 16), wv[q].x, wv[q].y, wv[q].z, wv[q].w);
                        }
                    }
                }
                fence_proxy_async_smem();        // generic-proxy stores → visible to the tensor core (async proxy)
                mbar_arrive(full_bar + s);
            }
        }
    } else if (warp == 8) {
        {
            // ===================================================== MMA issuer: uniform loop over the warp, one elected lane issues
            // (operands in uniform registers — tc05.cuh::elect_one)
            uint32_t it = 0, tl = 0;
            for (int work = blockIdx.x; work < num_tiles; work += gridDim.x, ++tl) {
                const int sp = work % splits;
                const int k_lo = sp * kper, k_hi = min(nKall, k_lo + kper);
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(tmem_empty + acc, aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_addr = tmem_base + acc * BN;
                for (int kidx = k_lo; kidx < k_hi; ++kidx, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar + s, ph);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + s * A_STRIDE), b_addr = smem_u32(smem_b + s * B_BYTES);
                    const uint64_t da0 = make_desc_kmajor_nosw(a_addr, LBO_A, SBO);
                    const uint64_t db0 = use_tma ? make_smem_desc(b_addr) : make_desc_kmajor_nosw(b_addr, LBO_B, SBO);
                    const uint32_t bstep = use_tma ? (32u >> 4) : ((2 * LBO_B) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < CK / 16; ++k)
                            umma_f16(d_addr, da0 + (uint64_t)((k * 2 * LBO_A) >> 4), db0 + (uint64_t)(k * bstep), kIdesc, (kidx != k_lo || k != 0) ? 1u : 0u);
                        tcgen05_commit(empty_bar + s);
                    }
                }
                if (elect_one()) tcgen05_commit(tmem_full + acc);
            }
        }
    } else {
        // ===================================================== epilogue: warp quarter q owns TMEM lanes [32q, 32q+32) = tile rows
        const int q = warp & 3;
        uint32_t tl = 0;
        for (int work = blockIdx.x; work < num_tiles; work += gridDim.x, ++tl) {
            const int tile = work / splits, sp = work - tile * splits;
            const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            mbar_wait(tmem_full + acc, aph);
            tcgen05_fence_after();
            const long long g = (long long)m_blk * BM + q * 32 + lane;
            float* out = a.y + (size_t)g * a.Kout + (size_t)n_blk * BN;
            const bool add_bias = a.bias != nullptr && sp == 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)c0, v);
                if (g < Mtot) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]); o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
                        if (add_bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n_blk * BN + c0 + j));
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (splits > 1) {
                            atomicAdd(out + c0 + j, o.x); atomicAdd(out + c0 + j + 1, o.y); atomicAdd(out + c0 + j + 2, o.z); atomicAdd(out + c0 + j + 3, o.w);
                        } else {
                            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                            *reinterpret_cast<float4*>(out + c0 + j) = o;
                        }
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
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// fp32 → bf16 cast of a contiguous tensor (the TMA-im2col path consumes bf16 NHWC operands); with `gate` the value is zeroed where
// gate ≤ 0 — the ReLU backward mask fused into the cast of dY.  8 elements per thread: two 16-byte loads, one 16-byte store.
__global__ void conv_cast_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gate, __nv_bfloat16* __restrict__ out, long long n8,
                                      long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (gate) {
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gate) + 2 * i), gb = __ldg(reinterpret_cast<const float4*>(gate) + 2 * i + 1);
            const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = g[j] > 0.f ? v[j] : 0.f;
        }
        uint4 o;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {      // tail (n % 8 elements)
        const long long i = (n & ~7LL) + threadIdx.x;
        float v = x[i];
        if (gate && !(gate[i] > 0.f)) v = 0.f;
        out[i] = __float2bfloat16(v);
    }
}
int conv_cast_bf16_launch(const float* x, const float* gate, void* out, long long n, cudaStream_t stream) {
    const long long n8 = n / 8;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n8 + 255) / 256, 148 * 16));
    conv_cast_bf16_kernel<<<blocks, 256, 0, stream>>>(x, gate, reinterpret_cast<__nv_bfloat16*>(out), n8, n);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// rows variant: x is [rows][n] with `row_stride` floats between rows (the staged parameter rows of the stacked pairs), out is
// contiguous bf16 [rows][n]; n % 8 == 0, 16-byte aligned rows
__global__ void conv_cast_rows_bf16_kernel(const float* __restrict__ x, long long row_stride, __nv_bfloat16* __restrict__ out, int rows, long long n8) {
    const long long total = (long long)rows * n8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n8, j = i - r * n8;
        const float4* src = reinterpret_cast<const float4*>(x + r * row_stride) + 2 * j;
        const float4 a = __ldg(src), b = __ldg(src + 1);
        uint4 o;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(a.x, a.y); o.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(a.z, a.w); o.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(b.x, b.y); o.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(b.z, b.w); o.w = *reinterpret_cast<uint32_t*>(&t);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}
int conv_cast_rows_bf16_launch(const float* x, long long row_stride, void* out, int rows, long long n, cudaStream_t stream) {
    if (n % 8 != 0 || row_stride % 4 != 0) return -5;
    const long long total = (long long)rows * (n / 8);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16));
    conv_cast_rows_bf16_kernel<<<blocks, 256, 0, stream>>>(x, row_stride, reinterpret_cast<__nv_bfloat16*>(out), rows, n / 8);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// bf16 [K][RS][C] → bf16 [C][RS][K] (32×32 tiles through shared memory), for the strided layers' software-gather data gradient
__global__ void conv_pack_t_kernel(const __nv_bfloat16* __restrict__ wq, __nv_bfloat16* __restrict__ out, int K, int C, int RS) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int kt = (K + 31) / 32, ct = (C + 31) / 32;
    for (int blk = blockIdx.x; blk < kt * ct * RS; blk += gridDim.x) {
        const int rs = blk % RS, t = blk / RS, k0 = (t / ct) * 32, c0 = (t % ct) * 32;
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            const int kk = i >> 5, c = i & 31;
            if (k0 + kk < K && c0 + c < C) tile[kk][c] = wq[((size_t)(k0 + kk) * RS + rs) * C + c0 + c];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            const int c = i >> 5, kk = i & 31;
            if (k0 + kk < K && c0 + c < C) out[((size_t)(c0 + c) * RS + rs) * K + k0 + kk] = tile[kk][c];
        }
        __syncthreads();
    }
}
int conv_pack_t_launch(const void* wq, void* out, int K, int C, int RS, cudaStream_t stream) {
    const int tiles = ((K + 31) / 32) * ((C + 31) / 32) * RS;
    conv_pack_t_kernel<<<std::min(tiles, 148 * 16), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(wq), reinterpret_cast<__nv_bfloat16*>(out), K, C, RS);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ======================================================================================================= weight gradient
// dW[k][r][s][c] = Σ_pixels dY[pixel][k] · X[gather(pixel, r, s)][c]   — a GEMM whose REDUCTION runs over the output pixels:
//   D[(tap, c) 128 rows × BN k-columns] += Aᵀ · B with A[j = (tap, c)][pixel] gathered from X (MN-major: 8 consecutive
//   channels of one pixel are one 16-byte piece) and B[k][pixel] = dY (MN-major as well: the NHWC rows are used as they are).
// grid = (row tiles × k tiles × pixel splits); every CTA accumulates its pixel range in TMEM and adds the tile into the fp32
// [K][R][S][C] gradient with coalesced red.global.add (the buffer is zeroed by the caller).  Same warp roles as above.
template <int BN>
__global__ void __launch_bounds__(cv::kFwdThreads, 1) conv_wgrad_kernel(const __grid_constant__ ConvArgs a, int splits) {
    using namespace cv;
    constexpr int PK = 64, KCP = PK / 8;                  // pixels per stage, 8-pixel k-cores
    // MN-major no-swizzle tiles: piece (mn-core i, pixel k) at i·SBO + (k/8)·LBO + (k%8)·16; SBO padded by 16 B (bank spread)
    constexpr uint32_t LBO = 128, SBO = KCP * 128 + 16;
    constexpr uint32_t A_BYTES = (16 * SBO + 127) & ~127u, B_BYTES = ((BN / 8) * SBO + 127) & ~127u;
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
    int2* pinfo = reinterpret_cast<int2*>(tmem_ptr_smem + 4);     // [2 groups][64] per-pixel gather origin of the group's stage

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // uniform warp index
    const int PQ = a.P * a.Q, RSC = a.R * a.S * a.C;
    const long long Mtot = (long long)a.N * PQ;
    const int m_tiles = (RSC + BM - 1) / BM;
    int tt = blockIdx.x;
    const int split = tt % splits; tt /= splits;
    const int m_blk = tt % m_tiles, n_blk = tt / m_tiles;
    const int chunks_total = (int)((Mtot + PK - 1) / PK);
    const int per = (chunks_total + splits - 1) / splits;
    const int ch_lo = split * per, ch_hi = min(chunks_total, ch_lo + per);
    const int nch = max(ch_hi - ch_lo, 0);

    if (warp == 8 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 128); mbar_init(empty_bar + s, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 8) {
        // ===================================================== producers: two groups of 128 threads fill alternate stages;
        // coalesced: 32 consecutive lanes read the 128 channels (32 float4) of one pixel's tile rows, BN/4 lanes one dY row
        const int grp = warp >> 2, t = threadIdx.x & 127;
        int2* pi = pinfo + grp * PK;
        const uint32_t pi_s = smem_u32(pi);
        const int j0 = m_blk * BM;
        for (int ci = grp; ci < nch; ci += 2) {
            const int s = ci % STAGES;
            const uint32_t ph = (ci / STAGES) & 1;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");        // the group is done with the previous pixel table
            if (t < PK) {
                const long long g = (long long)(ch_lo + ci) * PK + t;
                int2 v = make_int2(-1, 0);
                if (g < Mtot) {
                    const int n = (int)(g / PQ), rem = (int)(g % PQ);
                    const int oy = rem / a.Q, ox = rem % a.Q;
                    const int y0 = oy * a.stride - a.pad_h, x0 = ox * a.stride - a.pad_w;
                    v = make_int2(n, ((y0 + 0x4000) << 16) | (x0 + 0x4000));
                }
                pi[t] = v;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            // A: 64 pixels × 32 float4 (128 tile rows j = j0 + 4·cq … +4)
            float4 fa[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int q = i * 128 + t, kp = q >> 5, cq = q & 31;
                int pn, pyx;
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(pn), "=r"(pyx) : "r"(pi_s + (uint32_t)kp * 8u));
                const int j = j0 + 4 * cq;
                const int tap = j / a.C, c = j - tap * a.C;
                const int r = tap / a.S, ss = tap - r * a.S;
                const int iy = (pyx >> 16) - 0x4000 + r, ix = (pyx & 0xFFFF) - 0x4000 + ss;
                const bool ok = pn >= 0 && j < RSC && (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
                fa[i] = ok ? __ldg(reinterpret_cast<const float4*>(a.x + ((size_t)(unsigned)((pn * a.H + iy) * a.W + ix)) * a.C + c))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // B: 64 pixels × BN/4 float4 of dY rows
            constexpr int F4B = BN / 4, NIB = (PK * F4B) / 128;
            float4 fb[NIB];
#pragma unroll
            for (int i = 0; i < NIB; ++i) {
                const int q = i * 128 + t, kp = q / F4B, cq = q % F4B;
                const long long g = (long long)(ch_lo + ci) * PK + kp;
                fb[i] = (g < Mtot) ? __ldg(reinterpret_cast<const float4*>(a.dy + (size_t)g * a.Kout + n_blk * BN) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(empty_bar + s, ph ^ 1);
            const uint32_t sa = smem_u32(smem_a + s * A_BYTES), sb = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int q = i * 128 + t, kp = q >> 5, cq = q & 31;
                const uint32_t addr = sa + (uint32_t)((cq >> 1) * SBO + (kp >> 3) * LBO + (kp & 7) * 16 + (cq & 1) * 8);
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack2_bf16(fa[i].x, fa[i].y)), "r"(pack2_bf16(fa[i].z, fa[i].w)) : "memory");
            }
#pragma unroll
            for (int i = 0; i < NIB; ++i) {
                const int q = i * 128 + t, kp = q / F4B, cq = q % F4B;
                const uint32_t addr = sb + (uint32_t)((cq >> 1) * SBO + (kp >> 3) * LBO + (kp & 7) * 16 + (cq & 1) * 8);
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack2_bf16(fb[i].x, fb[i].y)), "r"(pack2_bf16(fb[i].z, fb[i].w)) : "memory");
            }
            fence_proxy_async_smem();
            mbar_arrive(full_bar + s);
        }
    } else if (warp == 8) {
        {
            for (int ci = 0; ci < nch; ++ci) {
                const int s = ci % STAGES;
                const uint32_t ph = (ci / STAGES) & 1;
                mbar_wait(full_bar + s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
                const uint64_t da0 = make_desc_kmajor_nosw(a_addr, LBO, SBO), db0 = make_desc_kmajor_nosw(b_addr, LBO, SBO);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < PK / 16; ++k)
                        umma_f16(tmem_base, da0 + (uint64_t)((k * 2 * LBO) >> 4), db0 + (uint64_t)((k * 2 * LBO) >> 4), kIdesc, (ci | k) != 0 ? 1u : 0u);
                    tcgen05_commit(empty_bar + s);
                }
            }
            if (elect_one()) tcgen05_commit(done_bar);
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
            if (j < RSC) {   // dW is written in the framework's OIHW layout: [k][c][r][s]
                const int tap = j / a.C, c = j - tap * a.C, RS = a.R * a.S;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float* dst = a.mode == 1 ? a.dw + (size_t)(n_blk * BN + c0 + i) * RSC + j      // [k][r][s][c]: lanes = consecutive j
                                             : a.dw + ((size_t)(n_blk * BN + c0 + i) * a.C + c) * RS + tap;
                    atomicAdd(dst, __uint_as_float(v[i]));
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

template <int BN>
static int launch_wgrad(const ConvArgs& a, cudaStream_t stream) {
    using namespace cv;
    constexpr size_t smem = (size_t)STAGES * ((16 * (8 * 128 + 16) + 127) / 128 * 128 + ((BN / 8) * (8 * 128 + 16) + 127) / 128 * 128) + 256 + 1024 + 128;
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
    conv_wgrad_kernel<BN><<<tiles * splits, kFwdThreads, smem, stream>>>(a, splits);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// a.x = forward input (NHWC fp32), a.dy = output gradient [N·P·Q, Kout] fp32, a.dw = zeroed fp32 [Kout][R][S][C]
int conv_wgrad_launch(const ConvArgs& a, cudaStream_t stream) {
    if (a.C % 8 != 0 || a.Kout % 32 != 0 || a.N <= 0) return -5;
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.dy) & 15)) return -6;
    if ((long long)a.N * a.H * a.W >= (1ll << 30) || a.H >= 0x3000 || a.W >= 0x3000) return -5;
    const int bn = (a.Kout % 128 == 0) ? 128 : (a.Kout % 64 == 0) ? 64 : 32;    // ≤ 128: the dY slice of a stage stays in registers
    if (bn == 128) return launch_wgrad<128>(a, stream);
    if (bn == 64) return launch_wgrad<64>(a, stream);
    return launch_wgrad<32>(a, stream);
}

template <int BN, int CK, int MODE>
static int launch_conv_m(const ConvArgs& a, cudaStream_t stream) {
    using namespace cv;
    constexpr int KC = CK / 8;
    constexpr size_t smem = (size_t)STAGES * (((KC * (16 * 128 + 16)) + 127) / 128 * 128 + KC * (BN / 8) * 128) + 256 + 128 * 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_igemm_kernel<BN, CK, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long Mtot = (long long)a.N * a.P * a.Q;
    const int tiles = (int)((Mtot + BM - 1) / BM) * (a.Kout / BN);
    CUtensorMap map_w;
    memset(&map_w, 0, sizeof(map_w));
    int tma_w = 0;
    if (CK == 64 && make_kmajor_sw128_map(&map_w, a.wq, a.Kout, a.R * a.S * a.C, BN) == 0) tma_w = 1;
    int splits = 1;
    const int nKall = a.R * a.S * (a.C / CK);
    if (!a.relu && tiles * 2 <= sms && nKall >= 8) {          // too few tiles to fill the machine and a long reduction
        splits = std::min(std::min(sms / tiles, nKall / 4), 16);
        if (splits < 2) splits = 1;
        const int kper = (nKall + splits - 1) / splits;
        splits = (nKall + kper - 1) / kper;                   // every split gets a non-empty k range
    }
    if (splits > 1) cudaMemsetAsync(a.y, 0, (size_t)Mtot * a.Kout * sizeof(float), stream);
    conv_igemm_kernel<BN, CK, MODE><<<std::min(tiles * splits, sms), kFwdThreads, smem, stream>>>(a, map_w, tma_w, splits);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

template <int BN, int CK>
static int launch_conv(const ConvArgs& a, cudaStream_t stream) {
    return a.mode == 0 ? launch_conv_m<BN, CK, 0>(a, stream) : launch_conv_m<BN, CK, 1>(a, stream);
}

// picks the largest instantiated (BN | Kout, CK | C) tile; -5 when the shape is not supported (caller falls back)
int conv_igemm_launch(const ConvArgs& a, cudaStream_t stream) {
    if (a.C % 16 != 0 || a.Kout % 32 != 0 || a.N <= 0) return -5;
    if (a.mode == 1 && a.stride != 1 && a.stride != 2) return -5;                               // dgrad: stride 1 or 2
    if ((long long)a.N * a.H * a.W >= (1ll << 30) || a.H >= 0x3000 || a.W >= 0x3000) return -5;  // 32-bit pixel indices, 16-bit coords
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.y) & 15) || (reinterpret_cast<uintptr_t>(a.wq) & 15)) return -6;
    int bn = (a.Kout % 256 == 0) ? 256 : (a.Kout % 128 == 0) ? 128 : (a.Kout % 64 == 0) ? 64 : 32;
    {   // few output pixels (deep layers): narrower column tiles so that every SM gets a tile
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long m_tiles = ((long long)a.N * a.P * a.Q + cv::BM - 1) / cv::BM;
        while (bn > 64 && m_tiles * (a.Kout / bn) < sms) bn >>= 1;
    }
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
