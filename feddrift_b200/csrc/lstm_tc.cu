// lstm_tc — persistent, cluster-resident 2-layer LSTM (hidden 256) forward + BPTT for RNN_OriginalFedAvg
// (reference: fedml_api/model/nlp/rnn.py:18-33 — Embedding(90,8) → 2×LSTM(256, batch_first) → Linear on the last step;
// the reference runs it through cuDNN's per-timestep kernels: ~10 launches per timestep per layer per direction).
//
// ONE launch runs the whole sequence for MANY (client, model) pairs: grid = npairs × 8 CTAs, one thread-block CLUSTER of 8
// CTAs per pair (and per 16-row batch chunk).  CTA j of a cluster owns hidden units [32j, 32j+32) of BOTH layers, i.e. 128
// gate rows (i, f, g, o × 32 units) per layer.
//
// Forward (lstm2_fwd_kernel):
//   * the CTA's weight slices live in TENSOR MEMORY for the whole sequence: A1 = [W_hh1 | W_ih1] (128 × 272, 136 TMEM
//     columns) and A2 = [W_ih2 | W_hh2] (128 × 512, 256 columns), written once with tcgen05.st as packed bf16 pairs; shared
//     memory only holds the activations;
//   * every timestep is one tcgen05.mma chain per layer in the TS form (A from TMEM, B = activations [16 × K] from
//     shared memory in the no-swizzle K-major core-matrix layout, D = 128 gate rows × 16 batch columns fp32 in TMEM):
//     gatesᵀ = W_slice · [x_t | h_{t-1}]ᵀ — the recurrent GEMM runs "swapped" so that the 128-row MMA is full even at batch 16;
//   * layers are WAVEFRONT-pipelined: phase p computes layer 1 at time p and layer 2 at time p-1 (both only need h1_{p-1}),
//     so there is one cluster exchange per timestep instead of two;
//   * the epilogue (tcgen05.ld → bias → σ/tanh → cell update, c kept in registers) produces the CTA's 32 new hidden units
//     for 16 batch rows, stages them as bf16 in the operand layout and BROADCASTS the 1-KB slice into all 8 CTAs' next-step
//     operand buffers with cp.async.bulk shared::cta → shared::cluster (DSMEM); arrival is tracked by an mbarrier
//     transaction count in each destination — no cluster barrier in the time loop;
//   * gate activations, cell states and hidden states are written to a history workspace for BPTT.
//
// Backward (lstm2_bwd_kernel): same ownership, reversed wavefront.  TMEM holds the TRANSPOSED slices (W_hh2ᵀ, W_ih2ᵀ,
// W_hh1ᵀ: 256 hidden rows × 128 own gate rows each, two 128-lane tiles per matrix).  Per phase: sum the 8 partial dh blocks
// that arrived in the inbox (fixed order → deterministic), elementwise LSTM backward for the own units (dc carried in
// registers), dG (bf16) → smem operand + global history, partial dhᵀ[256 × 16] = W_sliceᵀ · dGᵀ on tcgen05, tcgen05.ld and
// a DSMEM bulk REDUCE-SCATTER of the 2-KB blocks to the owners of those hidden units.  Weight gradients are GEMMs over the
// saved histories (dW = dGᵀ · H with the MN-major tcgen05 GEMM of gemm_tc.cu), outside this file.
//
// All waits are bounded (trap after 4 s).  bf16 operands, fp32 accumulation, fp32 cell state / gates / gradients.
#include <cooperative_groups.h>

#include "kernels.h"
#include "tc05.cuh"

namespace cg = cooperative_groups;

namespace fdb {

namespace lstm {
constexpr int H = 256, CL = 8, U = 32, NB = 16, KX = 16;
constexpr int kThreads = 128;      // backward kernel
constexpr int kFwdThreads = 288;   // forward: 2 epilogue groups of 4 warps + 1 issuer warp
// forward TMEM map (columns)
// accumulators: D1 = 2 × 16 columns (split K), D2 = 4 × 16 columns
constexpr uint32_t A1_COL = 0, A1_XCOL = 128, A2_COL = 136, D1_COL = 392, D2_COL = 424, F_TMEM = 512;
// backward TMEM map: three transposed matrices × 2 tiles × 64 columns, then 6 accumulators × 16 columns
constexpr uint32_t BT_HH2 = 0, BT_IH2 = 128, BT_HH1 = 256, BD_REC2 = 384, BD_IN1 = 416, BD_REC1 = 448, B_TMEM = 512;
// instruction descriptor: D fp32, A/B bf16, both K-major, N = 16, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kLBO = (NB / 8) * 128, kSBO = 128;   // no-swizzle K-major: [k_core][n_core][8 rows][16 B]
}  // namespace lstm

// element offset of (batch row b, reduction index k) in a no-swizzle K-major operand tile of 16 rows
FDB_DEVICE int op_off(int b, int k) { return (((k >> 3) * (lstm::NB / 8) + (b >> 3)) << 6) + ((b & 7) << 3) + (k & 7); }

FDB_DEVICE uint64_t make_desc_nosw(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lstm::kLBO >> 4) << 16;
    d |= (uint64_t)(lstm::kSBO >> 4) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    return d;                          // layout type 0 = SWIZZLE_NONE
}
// D[tmem] (+)= A[tmem] · B[smem]   (TS form: the A operand is read from tensor memory)
FDB_DEVICE void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
FDB_DEVICE void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
}
FDB_DEVICE void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
FDB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
FDB_DEVICE void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
FDB_DEVICE uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
// DSMEM bulk copy: this CTA's shared memory → a cluster peer's shared memory; completion = tx bytes on the PEER's mbarrier
FDB_DEVICE void bulk_copy_s2c(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
FDB_DEVICE void mbar_wait_long(uint64_t* bar, uint32_t parity) {
    SpinGuard g;
    while (!mbar_try_wait(bar, parity)) {
        if (g.expired(4000000000LL)) __trap();
    }
}
FDB_DEVICE uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
    return *reinterpret_cast<uint32_t*>(&v);
}
// MUFU.TANH: one instruction, |abs err| ≲ 5e-4 — far below the bf16 rounding of the operands these activations feed; the
// forward saves the activations it used, so BPTT differentiates exactly the function that was evaluated
FDB_DEVICE float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
FDB_DEVICE float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
FDB_DEVICE void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
FDB_DEVICE float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }
FDB_DEVICE float tanh_f(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }

// 64 consecutive fp32 weights of one row → 32 packed bf16 pairs → 32 TMEM columns of this thread's lane
FDB_DEVICE void load_row_chunk_to_tmem(const float* __restrict__ src, uint32_t taddr) {
    uint32_t v[32];
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src) + j);
            v[2 * j] = pack_bf16(f.x, f.y);
            v[2 * j + 1] = pack_bf16(f.z, f.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = pack_bf16(__ldg(src + 2 * j), __ldg(src + 2 * j + 1));
    }
    tmem_st_x32(taddr, v);
}

// ======================================================================================================= forward
// Warp roles (288 threads): warps 0-3 = layer-1 epilogue group, warps 4-7 = layer-2 epilogue group (warp w reads TMEM lanes
// 32·(w mod 4) … = gate (w mod 4) of the CTA's 32 units), warp 8 = tcgen05.mma issuer.  Per phase the issuer waits for the
// inbound hidden-state slices (mbarrier tx count), issues the layer-1 chain (2 interleaved accumulators → commit bar 1) and
// the layer-2 chain (4 interleaved accumulators → commit bar 2): tiny M128×N16×K16 MMAs that accumulate into ONE tile are
// latency-bound (~57 cycles each back to back), independent accumulators overlap that latency and are summed by the
// epilogue.  Each group runs epilogue → cell update → staging → DSMEM broadcast on its own named barrier, so the layer-1
// epilogue overlaps the layer-2 MMAs.
__global__ void __cluster_dims__(lstm::CL, 1, 1) __launch_bounds__(lstm::kFwdThreads, 1)
lstm2_fwd_kernel(const __grid_constant__ LstmArgs a) {
    using namespace lstm;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int pair = blockIdx.x / CL;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), l = tid & 31;   // uniform warp index
    const int T = a.T, E = a.E;

    // ---- shared memory carve-up
    __nv_bfloat16* H1s = reinterpret_cast<__nv_bfloat16*>(smem_raw);            // [2][NB*256]
    __nv_bfloat16* H2s = H1s + 2 * NB * H;                                       // [2][NB*256]
    __nv_bfloat16* Xs = H2s + 2 * NB * H;                                        // [2][NB*16]
    __nv_bfloat16* stage = Xs + 2 * NB * KX;                                     // [2 parities][2 layers][NB*32]
    float* act_s = reinterpret_cast<float*>(stage + 2 * 2 * NB * U);             // [2 layers][4 gates][NB][32]
    uint64_t* hbar = reinterpret_cast<uint64_t*>(act_s + 2 * 4 * NB * U);        // [2]
    uint64_t* mma_bar = hbar + 2;                                                // [2]: layer 1, layer 2
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 2);

    const float* prow = a.params + a.row_off[pair];
    const float* w_ih1 = prow + a.off_wih1;
    const float* w_hh1 = prow + a.off_whh1;
    const float* w_ih2 = prow + a.off_wih2;
    const float* w_hh2 = prow + a.off_whh2;
    const float* emb = prow + a.off_emb;
    const int* tok = a.tokens + (size_t)pair * NB * T;

    if (tid == 0) {
        mbar_init(hbar + 0, 1); mbar_init(hbar + 1, 1); mbar_init(mma_bar + 0, 1); mbar_init(mma_bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(F_TMEM));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // zero the operand buffers (h_{-1} = 0, padded x columns = 0)
    for (int i = tid; i < (2 * NB * H * 2 + 2 * NB * KX) / 2; i += kFwdThreads) reinterpret_cast<uint32_t*>(H1s)[i] = 0u;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    const int grp = warp >> 2;                      // 0: layer-1 group, 1: layer-2 group, 2: issuer warp
    const int w = warp & 3, gt = tid & 127;         // gate index / thread index inside the group
    const uint32_t lane_addr = tmem + ((uint32_t)(w * 32) << 16);
    const int R = w * H + crank * U + l;            // row of the PyTorch [4H, K] weight matrices owned by this thread
    float bias = 0.f;
    if (grp == 0) {
        // ---- layer-1 weights → TMEM: A1 = [W_hh1 | W_ih1]
#pragma unroll 1
        for (int c = 0; c < 4; ++c) load_row_chunk_to_tmem(w_hh1 + (size_t)R * H + 64 * c, lane_addr + A1_COL + 32 * c);
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float lo = (2 * j < E) ? __ldg(w_ih1 + (size_t)R * E + 2 * j) : 0.f;
            const float hi = (2 * j + 1 < E) ? __ldg(w_ih1 + (size_t)R * E + 2 * j + 1) : 0.f;
            v[j] = pack_bf16(lo, hi);
        }
        tmem_st_x8(lane_addr + A1_XCOL, v);
        tmem_st_wait();
        bias = __ldg(prow + a.off_bih1 + R) + __ldg(prow + a.off_bhh1 + R);
        for (int i = gt; i < NB * KX; i += 128) {    // x_0 → Xs[0]
            const int b = i / KX, k = i % KX;
            const float xv = (k < E) ? __ldg(emb + (size_t)tok[b * T + 0] * E + k) : 0.f;
            Xs[op_off(b, k)] = __float2bfloat16(xv);
        }
    } else if (grp == 1) {
        // ---- layer-2 weights → TMEM: A2 = [W_ih2 | W_hh2]
#pragma unroll 1
        for (int c = 0; c < 4; ++c) load_row_chunk_to_tmem(w_ih2 + (size_t)R * H + 64 * c, lane_addr + A2_COL + 32 * c);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) load_row_chunk_to_tmem(w_hh2 + (size_t)R * H + 64 * c, lane_addr + A2_COL + 128 + 32 * c);
        tmem_st_wait();
        bias = __ldg(prow + a.off_bih2 + R) + __ldg(prow + a.off_bhh2 + R);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    cluster.sync();            // every CTA's barriers / buffers are initialised before any peer writes into them
    tcgen05_fence_after();

    // history layout is LAYER-outermost: [2][npairs][T (+1)][16][...]
    const int NP = (int)gridDim.x / CL;
    const bool keep = a.gates != nullptr;            // training: save gates / cell states for BPTT
    const bool keep_h = a.hhist != nullptr;
    const int unit = crank * U + l;
    const bool dbg = a.dbg != nullptr && blockIdx.x == 0;
    const uint32_t bytes_each = NB * U * 2;

    if (grp == 2) {
        // =================================================== MMA issuer: the whole warp walks the phase loop in uniform control flow,
        // one elected lane issues (operands in uniform registers: no R2UR/ELECT serialisation loop around every UTCHMMA)
        {
            long long seg[3] = {0, 0, 0}, tprev = clock64();
            for (int p = 0; p <= T; ++p) {
                const bool doL1 = p < T, doL2 = p >= 1;
                // arm this phase's inbound barrier first (the slices of phase p land in hbar[p&1])
                if (p < T && elect_one()) mbar_expect_tx(hbar + (p & 1), CL * bytes_each * ((doL1 ? 1u : 0u) + (doL2 ? 1u : 0u)));
                if (p >= 1) mbar_wait_long(hbar + ((p - 1) & 1), ((p - 1) >> 1) & 1);   // h1_{p-1} (and h2_{p-2}) from all 8 CTAs
                if (dbg) { const long long tn = clock64(); seg[0] += tn - tprev; tprev = tn; }
                tcgen05_fence_after();
                const uint32_t h1prev = smem_u32(H1s + ((p + 1) & 1) * NB * H);     // h1_{p-1}
                const uint64_t d_h1 = make_desc_nosw(h1prev), d_x = make_desc_nosw(smem_u32(Xs + (p & 1) * NB * KX));
                if (doL1 && elect_one()) {
                    // two accumulators: k-steps alternate; the x step joins accumulator 1
#pragma unroll
                    for (int s = 0; s < 16; ++s)
                        umma_ts_f16(tmem + D1_COL + 16 * (s & 1), tmem + A1_COL + 8 * s, d_h1 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc, s > 1 ? 1u : 0u);
                    umma_ts_f16(tmem + D1_COL + 16, tmem + A1_XCOL, d_x, kIdesc, 1u);
                    tcgen05_commit(mma_bar + 0);
                }
                if (doL2) {
                    const uint32_t h2prev = smem_u32(H2s + (p & 1) * NB * H);       // h2_{p-2}
                    const uint64_t d_h2 = make_desc_nosw(h2prev);
                    if (elect_one()) {
                        // four accumulators: (h1 even, h1 odd, h2 even, h2 odd) k-steps, issued round-robin
#pragma unroll
                        for (int s = 0; s < 16; ++s) {
                            umma_ts_f16(tmem + D2_COL + 16 * (s & 1), tmem + A2_COL + 8 * s, d_h1 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc, s > 1 ? 1u : 0u);
                            umma_ts_f16(tmem + D2_COL + 32 + 16 * (s & 1), tmem + A2_COL + 128 + 8 * s, d_h2 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc,
                                        s > 1 ? 1u : 0u);
                        }
                        tcgen05_commit(mma_bar + 1);
                    }
                }
                if (dbg) { const long long tn = clock64(); seg[1] += tn - tprev; tprev = tn; }
            }
            if (dbg && l == 0) { a.dbg[0] = seg[0]; a.dbg[1] = seg[1]; }
        }
    } else {
        // =================================================== epilogue / cell groups (layer = grp)
        const int layer = grp;
        float* gates_l = a.gates + (size_t)(layer * NP + pair) * T * NB * 4 * H;
        float* cst_l = a.cst + (size_t)(layer * NP + pair) * T * NB * H;
        __nv_bfloat16* hh_l = reinterpret_cast<__nv_bfloat16*>(a.hhist) + (size_t)(layer * NP + pair) * (T + 1) * NB * H;
        float* act = act_s + layer * 4 * NB * U;
        float creg[4] = {0.f, 0.f, 0.f, 0.f};
        const int nacc = layer == 0 ? 2 : 4;
        const uint32_t dcol = layer == 0 ? D1_COL : D2_COL;
        long long seg[4] = {0, 0, 0, 0}, tprev = clock64();
        const bool dbgt = dbg && gt == 0 && layer == 0;
        // layer 1 runs in phases 0 … T-1 (time p), layer 2 in phases 1 … T (time p-1)
        const int p_lo = layer == 0 ? 0 : 1, p_hi = layer == 0 ? T - 1 : T;
        for (int p = p_lo; p <= p_hi; ++p) {
            const int t = p - layer;
            // prefetch next step's token / embedding values early (layer-1 group only): 2 values per thread
            float xv[2] = {0.f, 0.f};
            if (layer == 0 && p + 1 < T) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = gt + 128 * q, b = i / KX, k = i % KX;
                    if (k < E) xv[q] = __ldg(emb + (size_t)__ldg(tok + b * T + p + 1) * E + k);
                }
            }
            mbar_wait_long(mma_bar + layer, (p - p_lo) & 1);
            tcgen05_fence_after();
            if (dbgt) { const long long tn = clock64(); seg[0] += tn - tprev; tprev = tn; }
            // ---- epilogue: this thread = gate row (gate w, unit l); 16 batch columns; sum the split accumulators
            {
                float z[16], z2[16];
                tmem_ld_x16(lane_addr + dcol, z);
                for (int q = 1; q < nacc; ++q) {
                    tmem_ld_x16(lane_addr + dcol + 16 * q, z2);
#pragma unroll
                    for (int b = 0; b < NB; ++b) z[b] += z2[b];
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float x = z[b] + bias;
                    act[(w * NB + b) * U + l] = (w == 2) ? tanh_fast(x) : sigmoid_fast(x);
                }
            }
            tcgen05_fence_before();
            named_bar_sync(1 + layer, 128);
            if (dbgt) { const long long tn = clock64(); seg[1] += tn - tprev; tprev = tn; }
            // ---- cell update for the own 32 units (unit l) × rows b = w + 4 i; stage the bf16 slice in operand layout
            __nv_bfloat16* st = stage + ((p & 1) * 2 + layer) * NB * U;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = w + 4 * i;
                const float gi = act[(0 * NB + b) * U + l], gf = act[(1 * NB + b) * U + l];
                const float gg = act[(2 * NB + b) * U + l], go = act[(3 * NB + b) * U + l];
                creg[i] = gf * creg[i] + gi * gg;
                const float h = go * tanh_fast(creg[i]);
                const size_t row = (size_t)t * NB + b;
                if (keep) {
                    float* g = gates_l + row * 4 * H + unit;
                    g[0] = gi; g[H] = gf; g[2 * H] = gg; g[3 * H] = go;
                    cst_l[row * H + unit] = creg[i];
                }
                const __nv_bfloat16 hb = __float2bfloat16(h);
                if (keep_h) hh_l[((size_t)(t + 1) * NB + b) * H + unit] = hb;
                st[op_off(b, l)] = hb;     // k_core = l/8 local to the slice
                if (layer == 1 && t == T - 1) a.hlast[((size_t)pair * NB + b) * H + unit] = h;
            }
            if (layer == 0 && p + 1 < T) {   // x_{p+1} → Xs[(p+1)&1]
                __nv_bfloat16* xd = Xs + ((p + 1) & 1) * NB * KX;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = gt + 128 * q;
                    xd[op_off(i / KX, i % KX)] = __float2bfloat16(xv[q]);
                }
            }
            fence_proxy_async_smem();   // generic-proxy writes (stage, Xs) → visible to the async proxy (bulk copy, tcgen05)
            named_bar_sync(1 + layer, 128);
            if (dbgt) { const long long tn = clock64(); seg[2] += tn - tprev; tprev = tn; }
            // ---- all-gather: my 1-KB slice → every CTA's next-step operand buffer (incl. my own); lane d serves CTA d.
            //      (the last phase's h2_{T-1} is only needed as hlast, no exchange)
            if (p < T && gt < CL) {
                const uint32_t bar = smem_u32(hbar + (p & 1));
                const uint32_t dst = (layer == 0 ? smem_u32(H1s + (p & 1) * NB * H) : smem_u32(H2s + ((p + 1) & 1) * NB * H)) + crank * bytes_each;
                bulk_copy_s2c(mapa_u32(dst, gt), smem_u32(st), bytes_each, mapa_u32(bar, gt));
            }
            if (dbgt) { const long long tn = clock64(); seg[3] += tn - tprev; tprev = tn; }
        }
        if (dbgt) for (int i = 0; i < 4; ++i) a.dbg[2 + i] = seg[i];
    }
    // ---- teardown: nobody may exit while peers still copy into its shared memory
    tcgen05_fence_before();
    cluster.sync();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(F_TMEM));
}

// ======================================================================================================= backward
// 64 own gate rows (r0 .. r0+63) of column `k` of W → 32 packed pairs → 32 TMEM columns (transposed slice)
FDB_DEVICE void load_col_chunk_to_tmem(const float* __restrict__ W, int ldw, int k, int crank, int r0, uint32_t taddr) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int ra = r0 + 2 * j, rb = ra + 1;
        const int Ra = (ra >> 5) * lstm::H + crank * lstm::U + (ra & 31), Rb = (rb >> 5) * lstm::H + crank * lstm::U + (rb & 31);
        v[j] = pack_bf16(__ldg(W + (size_t)Ra * ldw + k), __ldg(W + (size_t)Rb * ldw + k));
    }
    tmem_st_x32(taddr, v);
}

// Warp roles (288 threads): warps 0-3 = layer-2 group (time p), warps 4-7 = layer-1 group (time p+1), warp 8 = MMA issuer.
// Per phase each group: prefetch its history rows → wait for the inbox → Σ partial dh → LSTM cell backward → dG (bf16) to the
// smem operand + global history → signal the issuer (mbarrier) → wait for its accumulator tiles → tcgen05.ld → staging →
// DSMEM bulk reduce-scatter (one lane per copy).  The issuer runs the layer-2 chains (4 independent accumulators, 32 MMAs)
// and the layer-1 chains (2 accumulators, 16 MMAs) as soon as the respective dG operand is ready, so one layer's
// epilogue / exchange overlaps the other layer's tensor work.
__global__ void __cluster_dims__(lstm::CL, 1, 1) __launch_bounds__(lstm::kFwdThreads, 1)
lstm2_bwd_kernel(const __grid_constant__ LstmArgs a) {
    using namespace lstm;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int pair = blockIdx.x / CL;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), l = tid & 31;   // uniform warp index
    const int T = a.T;

    // ---- shared memory: inbox [2 bufs][8 src][3 kinds][NB][32] fp32, out staging [2 bufs][3 kinds][8 dst][NB][32] fp32,
    //      dG operands [2 layers][NB × 128] bf16
    constexpr int BLK = NB * U;   // 512 floats = 2 KB
    float* inbox = reinterpret_cast<float*>(smem_raw);                  // 2*8*3*BLK
    float* outst = inbox + 2 * CL * 3 * BLK;                            // 2*3*8*BLK
    __nv_bfloat16* dGs = reinterpret_cast<__nv_bfloat16*>(outst + 2 * 3 * CL * BLK);   // [2][NB*128]: [0] layer 1, [1] layer 2
    uint64_t* ibar = reinterpret_cast<uint64_t*>(dGs + 2 * NB * 128);  // [2] inbox arrival (tx bytes)
    uint64_t* gbar = ibar + 2;                                          // [2] dG operand ready: [0] layer 2, [1] layer 1
    uint64_t* mma_bar = gbar + 2;                                       // [2] accumulators ready: [0] layer 2, [1] layer 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 2);

    const float* prow = a.params + a.row_off[pair];
    const float* w_hh1 = prow + a.off_whh1;
    const float* w_ih2 = prow + a.off_wih2;
    const float* w_hh2 = prow + a.off_whh2;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(ibar + i, 1); mbar_init(gbar + i, 1); mbar_init(mma_bar + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(B_TMEM));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const int grp = warp >> 2;                 // 0: layer 2, 1: layer 1, 2: issuer
    const int w = warp & 3, gt = tid & 127;
    const uint32_t lane_addr = tmem + ((uint32_t)(w * 32) << 16);

    // ---- transposed weight slices → TMEM: tile h, lane = hidden column k = 128 h + gt; K = own 128 gate rows.
    //      group 0 loads W_hh2ᵀ and half of W_ih2ᵀ, group 1 the other half and W_hh1ᵀ
    if (grp < 2) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int k = 128 * h + gt;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                if (grp == 0) {
                    load_col_chunk_to_tmem(w_hh2, H, k, crank, 64 * c, lane_addr + BT_HH2 + 64 * h + 32 * c);
                    if (h == 0) load_col_chunk_to_tmem(w_ih2, H, k, crank, 64 * c, lane_addr + BT_IH2 + 64 * h + 32 * c);
                } else {
                    load_col_chunk_to_tmem(w_hh1, H, k, crank, 64 * c, lane_addr + BT_HH1 + 64 * h + 32 * c);
                    if (h == 1) load_col_chunk_to_tmem(w_ih2, H, k, crank, 64 * c, lane_addr + BT_IH2 + 64 * h + 32 * c);
                }
            }
        }
        tmem_st_wait();
    }
    tcgen05_fence_before();
    cluster.sync();
    tcgen05_fence_after();

    const int unit = crank * U + l;
    const int NP = (int)gridDim.x / CL;

    if (grp == 2) {
        // =================================================== MMA issuer (uniform loop, one elected lane issues — see the forward kernel)
        {
            const uint64_t d_g2 = make_desc_nosw(smem_u32(dGs + NB * 128)), d_g1 = make_desc_nosw(smem_u32(dGs));
            for (int p = T - 1, it = 0; p >= 0; --p, ++it) {
                const bool doL1 = p + 1 <= T - 1;
                if (elect_one()) mbar_expect_tx(ibar + (it & 1), (uint32_t)(CL * (doL1 ? 3 : 2) * BLK * 4));   // this phase's inbound partial blocks
                mbar_wait_long(gbar + 0, it & 1);          // dG of layer 2 (time p) is in shared memory
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            umma_ts_f16(tmem + BD_REC2 + 16 * h, tmem + BT_HH2 + 64 * h + 8 * s, d_g2 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc, s > 0 ? 1u : 0u);
                            umma_ts_f16(tmem + BD_IN1 + 16 * h, tmem + BT_IH2 + 64 * h + 8 * s, d_g2 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc, s > 0 ? 1u : 0u);
                        }
                    }
                    tcgen05_commit(mma_bar + 0);
                }
                if (doL1) {
                    mbar_wait_long(gbar + 1, (it - 1) & 1);   // dG of layer 1 (time p+1)
                    tcgen05_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                umma_ts_f16(tmem + BD_REC1 + 16 * h, tmem + BT_HH1 + 64 * h + 8 * s, d_g1 + (uint64_t)((s * 2 * kLBO) >> 4), kIdesc, s > 0 ? 1u : 0u);
                        }
                        tcgen05_commit(mma_bar + 1);
                    }
                }
            }
        }
    } else {
        // =================================================== layer groups
        const int layer = 1 - grp;                 // group 0 ↔ layer index 1 (second layer), group 1 ↔ layer index 0
        const float* gates_l = a.gates + (size_t)(layer * NP + pair) * T * NB * 4 * H;
        const float* cst_l = a.cst + (size_t)(layer * NP + pair) * T * NB * H;
        __nv_bfloat16* dG_l = reinterpret_cast<__nv_bfloat16*>(a.dgates) + (size_t)(layer * NP + pair) * T * NB * 4 * H;
        __nv_bfloat16* dg_s = dGs + layer * NB * 128;
        float dcar[4] = {0.f, 0.f, 0.f, 0.f};
        // group 0 (layer 2) runs in phases it = 0 … T-1 at time t = T-1-it; group 1 (layer 1) in phases it = 1 … T at time t = T-it
        const int it_lo = grp, it_hi = T - 1 + grp;
        for (int it = it_lo; it <= it_hi; ++it) {
            const int t = T - 1 - it + grp;
            const int p = T - 1 - it;                                  // the phase's layer-2 time (−1 in the last phase)
            // ---- prefetch this step's history rows (independent of the inbox) before waiting
            float hg[4][4], hc[4], hcp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = w + 4 * i;
                const size_t row = (size_t)t * NB + b;
                const float* g = gates_l + row * 4 * H + unit;
                hg[i][0] = g[0]; hg[i][1] = g[H]; hg[i][2] = g[2 * H]; hg[i][3] = g[3 * H];
                hc[i] = cst_l[row * H + unit];
                hcp[i] = (t > 0) ? cst_l[((size_t)(t - 1) * NB + b) * H + unit] : 0.f;
            }
            const bool have_in = it > 0;
            const int ibuf = (it + 1) & 1;                             // inbox buffer written during phase it-1
            if (have_in) mbar_wait_long(ibar + ibuf, ((it - 1) >> 1) & 1);
            const float* in = inbox + (size_t)ibuf * CL * 3 * BLK;
            const bool prev_had_L1 = have_in && (p + 2 <= T - 1);      // phase it-1 also ran layer 1 (at time p+2)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = w + 4 * i;
                float dh = 0.f;
                if (grp == 0) {
                    if (a.dh2_all) dh = a.dh2_all[(((size_t)pair * T + t) * NB + b) * H + unit];
                    else if (t == T - 1) dh = a.dh2_last[((size_t)pair * NB + b) * H + unit];
                    if (have_in) {
#pragma unroll
                        for (int sidx = 0; sidx < CL; ++sidx) dh += in[(sidx * 3 + 0) * BLK + b * U + l];
                    }
                } else {
#pragma unroll
                    for (int sidx = 0; sidx < CL; ++sidx) dh += in[(sidx * 3 + 1) * BLK + b * U + l];
                    if (prev_had_L1) {
#pragma unroll
                        for (int sidx = 0; sidx < CL; ++sidx) dh += in[(sidx * 3 + 2) * BLK + b * U + l];
                    }
                }
                const float gi = hg[i][0], gf = hg[i][1], gg = hg[i][2], go = hg[i][3];
                const float tc = tanh_fast(hc[i]);
                const float dcv = dcar[i] + dh * go * (1.f - tc * tc);
                const float dzi = dcv * gg * gi * (1.f - gi), dzf = dcv * hcp[i] * gf * (1.f - gf);
                const float dzg = dcv * gi * (1.f - gg * gg), dzo = dh * tc * go * (1.f - go);
                dcar[i] = dcv * gf;
                __nv_bfloat16* o = dG_l + ((size_t)t * NB + b) * 4 * H + unit;
                const __nv_bfloat16 bi = __float2bfloat16(dzi), bf = __float2bfloat16(dzf), bg = __float2bfloat16(dzg), bo = __float2bfloat16(dzo);
                o[0] = bi; o[H] = bf; o[2 * H] = bg; o[3 * H] = bo;
                dg_s[op_off(b, 0 * 32 + l)] = bi; dg_s[op_off(b, 1 * 32 + l)] = bf; dg_s[op_off(b, 2 * 32 + l)] = bg; dg_s[op_off(b, 3 * 32 + l)] = bo;
            }
            if (p < 0) break;                                          // layer 1 at time 0: its dh_{-1} has no consumer
            fence_proxy_async_smem();
            named_bar_sync(1 + grp, 128);
            if (gt == 0) mbar_arrive(gbar + grp);                      // → issuer: this layer's dG operand is ready
            // ---- my accumulator tiles → staging blocks [b][unit-in-owner] (this warp's lanes = 32 hidden units of owner 4h + w)
            mbar_wait_long(mma_bar + grp, (it - it_lo) & 1);
            tcgen05_fence_after();
            float* ob = outst + (size_t)(it & 1) * 3 * CL * BLK;
            const int k_lo = grp == 0 ? 0 : 2, k_hi = grp == 0 ? 1 : 2;
            for (int kind = k_lo; kind <= k_hi; ++kind) {
                const uint32_t col = (kind == 0) ? BD_REC2 : (kind == 1 ? BD_IN1 : BD_REC1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float z[16];
                    tmem_ld_x16(lane_addr + col + 16 * h, z);
                    float* dst = ob + (size_t)(kind * CL + (4 * h + w)) * BLK;
#pragma unroll
                    for (int b = 0; b < NB; ++b) dst[b * U + l] = z[b];
                }
            }
            fence_proxy_async_smem();
            tcgen05_fence_before();
            named_bar_sync(1 + grp, 128);
            // ---- reduce-scatter: block (kind, owner d) → owner d's inbox slot [src = me][kind]; one lane per copy
            {
                const int obuf = it & 1, ncopy = (k_hi - k_lo + 1) * CL;
                if (gt < ncopy) {
                    const int kind = k_lo + gt / CL, d = gt % CL;
                    const uint32_t dst = smem_u32(inbox + ((size_t)obuf * CL + crank) * 3 * BLK + kind * BLK);
                    bulk_copy_s2c(mapa_u32(dst, d), smem_u32(ob + (size_t)(kind * CL + d) * BLK), BLK * 4, mapa_u32(smem_u32(ibar + obuf), d));
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster.sync();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(B_TMEM));
}

// ======================================================================================================= classifier head
// Linear(256 → V) on the last hidden state + softmax cross-entropy + every gradient of the head, one CTA per 16-row chunk:
// logits, CE (mean over the pair's real rows via `scale`), dlogits, dW_fc, db_fc and dh2_{T-1} (the BPTT kernel's input).
// Replaces fc GEMM + log_softmax + nll + three backward GEMMs + bias reduction (~9 launches) per pair and step.
__global__ void __launch_bounds__(256) lstm_head_kernel(const __grid_constant__ LstmHeadArgs a) {
    constexpr int H = lstm::H, NB = lstm::NB, VP = 96;
    __shared__ float h_s[NB][H];
    __shared__ float dl_s[NB][VP];
    __shared__ float red_s[8];
    const int ch = blockIdx.x, tid = threadIdx.x, w = tid >> 5, l = tid & 31, V = a.V;
    const float* prow = a.params + a.row_off[ch];
    const float* Wfc = prow + a.off_fcw;
    const float* bfc = prow + a.off_fcb;
    const float* hl = a.hlast + (size_t)ch * NB * H;
    for (int i = tid; i < NB * H; i += 256) h_s[i / H][i % H] = hl[i];
    __syncthreads();
    // logits[b][v] = h[b]·W[v] + bias[v]: one warp per output neuron, lanes split K, 16 rows at once
    for (int v = w; v < V; v += 8) {
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = __ldg(Wfc + (size_t)v * H + l + 32 * j);
        const float bv = __ldg(bfc + v);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(wv[j], h_s[b][l + 32 * j], acc);
            acc = warp_sum(acc);
            if (l == 0) dl_s[b][v] = acc + bv;
        }
    }
    __syncthreads();
    // softmax-CE per row; dlogits = (p − onehot)·scale   (scale = 1 / #real rows of the pair; padding rows: label < 0 → 0)
    const float scale = a.scale[ch];
    float lsum = 0.f;
    for (int b = w; b < NB; b += 8) {
        const int y = a.labels[ch * NB + b];
        float mx = -INFINITY;
        for (int v = l; v < V; v += 32) mx = fmaxf(mx, dl_s[b][v]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int v = l; v < V; v += 32) se += __expf(dl_s[b][v] - mx);
        se = warp_sum(se);
        const float inv = 1.f / se;
        for (int v = l; v < V; v += 32) {
            const float z = dl_s[b][v];
            const float pr = __expf(z - mx) * inv;
            if (y >= 0 && v == y) lsum += (mx + __logf(se) - z) * scale;
            dl_s[b][v] = (y >= 0) ? (pr - (v == y ? 1.f : 0.f)) * scale : 0.f;
        }
    }
    lsum = warp_sum(lsum);
    if (l == 0) red_s[w] = lsum;
    __syncthreads();
    if (tid == 0 && a.loss) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red_s[i]; a.loss[ch] = t; }
    // thread k: dW[v][k] = Σ_b dl[b][v] h[b][k],  dh[b][k] = Σ_v dl[b][v] W[v][k]
    {
        const int k = tid;
        float hr[NB], dh[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) { hr[b] = h_s[b][k]; dh[b] = 0.f; }
        float* dW = a.dW + (size_t)ch * V * H;
        for (int v = 0; v < V; ++v) {
            const float wv = __ldg(Wfc + (size_t)v * H + k);
            float acc = 0.f;
#pragma unroll
            for (int b = 0; b < NB; ++b) { const float d = dl_s[b][v]; acc = fmaf(d, hr[b], acc); dh[b] = fmaf(d, wv, dh[b]); }
            dW[(size_t)v * H + k] = acc;
        }
        float* dho = a.dh + (size_t)ch * NB * H;
#pragma unroll
        for (int b = 0; b < NB; ++b) dho[b * H + k] = dh[b];
    }
    if (tid < V) {
        float acc = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) acc += dl_s[b][tid];
        a.db[(size_t)ch * V + tid] = acc;
    }
}

int lstm_head_launch(const LstmHeadArgs& a, int nchunks, cudaStream_t stream) {
    if (a.V > 96 || a.V < 1) return -5;
    lstm_head_kernel<<<nchunks, 256, 0, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ======================================================================================================= small gradients
// Bias, W_ih1 and embedding gradients of every chunk in one pass over the bf16 gate-gradient histories.  Because the layer-1
// input is x_r = emb[tok_r], all three follow from the token-segmented column sums S[v][col] = Σ_{r: tok_r = v} dG1[r][col]:
//   db1[col] = Σ_v S[v][col],   dW_ih1[col][e] = Σ_v S[v][col]·emb[v][e],   d emb[v][e] = Σ_col S[v][col]·W_ih1[col][e]
// (emb / W_ih1 rounded to bf16 exactly as the forward kernel fed them to the tensor core); db2 is a plain column sum of dG2.
// grid = (chunks, 8 column slices of 128); one thread per gate column; S lives in shared memory (no atomics: a thread owns
// its column).  Replaces ~12 eager kernels that converted both histories to fp32 (≈1.3 GB of traffic per local step).
__global__ void __launch_bounds__(128) lstm_small_grads_kernel(const __grid_constant__ LstmSmallArgs a) {
    constexpr int H4 = 4 * lstm::H, NB = lstm::NB, VP = 96, EP = 16;
    extern __shared__ float sg[];
    float* S = sg;                       // [VP][129]
    float* embq = S + VP * 129;          // [VP][EP]
    float* wq = embq + VP * EP;          // [128][EP]
    int* tokr = reinterpret_cast<int*>(wq + 128 * EP);   // [T*NB] token of history row r = t*16 + b
    const int ch = blockIdx.x, slice = blockIdx.y, tid = threadIdx.x, col = slice * 128 + tid;
    const int T = a.T, E = a.E, V = a.V, TB = T * NB, NP = (int)gridDim.x;
    const float* prow = a.params + a.row_off[ch];
    for (int i = tid; i < VP * 129; i += 128) S[i] = 0.f;
    for (int i = tid; i < VP * EP; i += 128) {
        const int v = i / EP, e = i % EP;
        embq[i] = (v < V && e < E) ? __bfloat162float(__float2bfloat16(__ldg(prow + a.off_emb + (size_t)v * E + e))) : 0.f;
    }
    for (int i = tid; i < 128 * EP; i += 128) {
        const int c = i / EP, e = i % EP;
        wq[i] = (e < E) ? __bfloat162float(__float2bfloat16(__ldg(prow + a.off_wih1 + (size_t)(slice * 128 + c) * E + e))) : 0.f;
    }
    const int* tk = a.tokens + (size_t)ch * NB * T;
    for (int r = tid; r < TB; r += 128) tokr[r] = tk[(r % NB) * T + r / NB];
    __syncthreads();
    const __nv_bfloat16* g1 = reinterpret_cast<const __nv_bfloat16*>(a.dgates) + ((size_t)(0 * NP + ch) * TB) * H4 + col;
    const __nv_bfloat16* g2 = reinterpret_cast<const __nv_bfloat16*>(a.dgates) + ((size_t)(1 * NP + ch) * TB) * H4 + col;
    float b2 = 0.f;
    int r = 0;
    for (; r + 4 <= TB; r += 4) {          // 8 independent loads in flight per thread
        float x1[4], x2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { x1[q] = __bfloat162float(g1[(size_t)(r + q) * H4]); x2[q] = __bfloat162float(g2[(size_t)(r + q) * H4]); }
#pragma unroll
        for (int q = 0; q < 4; ++q) { S[tokr[r + q] * 129 + tid] += x1[q]; b2 += x2[q]; }
    }
    for (; r < TB; ++r) { S[tokr[r] * 129 + tid] += __bfloat162float(g1[(size_t)r * H4]); b2 += __bfloat162float(g2[(size_t)r * H4]); }
    // per-column results (this thread's column only: no sync needed yet)
    float b1 = 0.f, dw[EP];
#pragma unroll
    for (int e = 0; e < EP; ++e) dw[e] = 0.f;
    for (int v = 0; v < V; ++v) {
        const float sv = S[v * 129 + tid];
        b1 += sv;
#pragma unroll
        for (int e = 0; e < EP; ++e) dw[e] = fmaf(sv, embq[v * EP + e], dw[e]);
    }
    a.db1[(size_t)ch * H4 + col] = b1;
    a.db2[(size_t)ch * H4 + col] = b2;
    for (int e = 0; e < E; ++e) a.dwih1[((size_t)ch * H4 + col) * E + e] = dw[e];
    __syncthreads();
    // embedding partial of this column slice: demb_part[ch][slice][v][e] = Σ_c S[v][c]·W_ih1[c][e]
    for (int i = tid; i < V * E; i += 128) {
        const int v = i / E, e = i % E;
        float acc = 0.f;
        for (int c = 0; c < 128; ++c) acc = fmaf(S[v * 129 + c], wq[c * EP + e], acc);
        a.demb_part[(((size_t)ch * 8 + slice) * V + v) * E + e] = acc;
    }
}

int lstm_small_grads_launch(const LstmSmallArgs& a, int nchunks, cudaStream_t stream) {
    if (a.V > 96 || a.E > 16) return -5;
    const size_t smem = (size_t)(96 * 129 + 96 * 16 + 128 * 16) * 4 + (size_t)a.T * lstm::NB * 4;
    cudaError_t e = cudaFuncSetAttribute(lstm_small_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -3;
    lstm_small_grads_kernel<<<dim3(nchunks, 8), 128, smem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ======================================================================================================= launchers
static size_t fwd_smem_bytes() {
    using namespace lstm;
    size_t b = (size_t)(2 * NB * H + 2 * NB * H + 2 * NB * KX + 2 * 2 * NB * U) * 2 + (size_t)2 * 4 * NB * U * 4 + 128;
    return b < 120 * 1024 ? 120 * 1024 : b;   // > half an SM's shared memory: exactly one CTA (one 512-column TMEM owner) per SM
}
static size_t bwd_smem_bytes() {
    using namespace lstm;
    return (size_t)(2 * CL * 3 + 2 * 3 * CL) * NB * U * 4 + (size_t)2 * NB * 128 * 2 + 128;
}

int lstm2_fwd_launch(const LstmArgs& a, int npairs, cudaStream_t stream) {
    const size_t smem = fwd_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(lstm2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -3;
    lstm2_fwd_kernel<<<npairs * lstm::CL, lstm::kFwdThreads, smem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int lstm2_bwd_launch(const LstmArgs& a, int npairs, cudaStream_t stream) {
    const size_t smem = bwd_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(lstm2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -3;
    lstm2_bwd_kernel<<<npairs * lstm::CL, lstm::kFwdThreads, smem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
