// gemm_tn: D[M,N] = act(A[M,K] · B[N,K]ᵀ + bias[N])   bf16 operands, fp32 accumulation in TMEM.
//
// Hand-written Blackwell GEMM used by TcLinear (fnn-MNIST 784→1568→10, CNN fc 9216→128, LSTM/classifier heads):
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128B-swizzled 128×64 bf16 tiles) into a 4-stage smem ring,
//     full/empty mbarrier pipeline, one elected producer thread;
//   * tcgen05.mma.cta_group::1.kind::f16 (UMMA 128×128×16) issued by ONE thread, accumulator in TMEM
//     (128 lanes × 128 fp32 columns), completion signalled with tcgen05.commit → mbarrier;
//   * epilogue: 4 warps read the accumulator with tcgen05.ld.32x32b.x32, fuse bias + ReLU + (optional) bf16 cast
//     in registers and store 128-byte row segments straight to global memory (no smem round trip, no extra pass).
// The reference's equivalent is eager `nn.Linear` + separate bias/ReLU kernels in fp32 on cuBLAS
// (fedml_api/model/fnn/fnn.py:11-15, cv/cnn.py:128-136).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace fdb {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int kTmemCols = 128;
constexpr int kGemmThreads = 256;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr uint32_t kSmemBytes = STAGES * (kStageBytesA + kStageBytesB) + 1024 /*align slack*/ + 256 /*barriers*/;

// ---------------------------------------------------------------- PTX wrappers
FDB_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FDB_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
FDB_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FDB_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must fault the context (trap), never hang the GPU
FDB_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 2000000000LL) __trap();
    }
}
FDB_DEVICE void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
FDB_DEVICE void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FDB_DEVICE void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FDB_DEVICE void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
FDB_DEVICE void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128B-swizzled operand tile: rows are 128 B, 8-row atoms are 1024 B apart (SBO), version = 1 (sm_100)
FDB_DEVICE uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)(0) << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((1024u) >> 4) << 32;                 // stride byte offset
    d |= (uint64_t)1 << 46;                              // descriptor version
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, void* __restrict__ D,
               const float* __restrict__ bias, int M, int N, int K, int relu, int out_fp32) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * kStageBytesA;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (kStageBytesA + kStageBytesB));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.y, n_blk = blockIdx.x;
    const int num_k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_expect_tx(full_bar + s, kStageBytesA + kStageBytesB);
                tma_load_2d(&map_a, full_bar + s, smem_a + s * kStageBytesA, kb * BK, m_blk * BM);
                tma_load_2d(&map_b, full_bar + s, smem_b + s * kStageBytesB, kb * BK, n_blk * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer (single thread)
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(full_bar + s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + s * kStageBytesA);
                const uint32_t b_addr = smem_u32(smem_b + s * kStageBytesB);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 2);
                    umma_f16(tmem_base, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
                }
                tcgen05_commit(empty_bar + s);  // frees the smem stage once these MMAs retire
            }
            tcgen05_commit(tmem_full_bar);      // accumulator complete → epilogue
        }
    } else if (warp >= 4) {
        // ===== epilogue: warp (4+q) owns TMEM lanes [32q, 32q+32) == output rows
        const int q = warp - 4;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        const int row = m_blk * BM + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int col0 = n_blk * BN + c0;
            if (row < M && col0 < N) {
                if (out_fp32) {
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
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
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

int gemm_tn_launch(const void* A, const void* B, void* D, const float* bias, int M, int N, int K, int relu, int out_fp32,
                   cudaStream_t stream) {
    if (K % 8 != 0 || M <= 0 || N <= 0) return -5;  // TMA global stride must be a multiple of 16 B
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -6;
    CUtensorMap ma, mb;
    if (make_map(&ma, A, M, K, BM) != 0 || make_map(&mb, B, N, K, BN) != 0) return -7;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) return -3;
        attr_set = true;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    gemm_tn_kernel<<<grid, kGemmThreads, kSmemBytes, stream>>>(ma, mb, D, bias, M, N, K, relu, out_fp32);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
