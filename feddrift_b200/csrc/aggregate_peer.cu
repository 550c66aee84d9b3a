// fedavg_reduce_apply_peer — the multi-GPU per-cluster FedAvg aggregation + broadcast as ONE kernel per rank, with the
// collective done by the kernel itself over NVLink peer memory (no NCCL), PIPELINED chunk by chunk:
//
//   producer warps  (12 of the 16 warps of EVERY CTA) stream the local clients' rows IN PLACE from the client arena (row list `cidx`, no gather copy) and
//                   write un-normalised partial sums part_g[m, chunk] = Σ_{c on g} n[c,m]·θ_c[m, chunk]   (HBM-bound, like K1);
//                   the last CTA to finish a chunk publishes a per-chunk epoch flag to every peer
//                   (fence.sys + st.release.sys);
//   consumer warps  (the other 4 warps of every CTA) run concurrently: as soon as chunk k is flagged by ALL ranks they reduce-scatter + normalise +
//                   all-gather it: rank g owns 1/W of the chunk, pulls it from every peer's partial buffer (128-bit peer
//                   loads, or ONE multimem.ld_reduce = in-switch add when the buffer has an NVLS multicast mapping),
//                   divides by the global weight total and pushes the finished piece into EVERY rank's θ buffer
//                   (peer stores / multimem.st) — the broadcast of the new cluster models is the epilogue of the reduction;
//   so the HBM stream of chunk k+1 overlaps the NVLink traffic of chunk k (round 1 ran the two phases back to back behind
//   a grid barrier + peer barrier).  One final barrier makes θ complete everywhere before the kernel retires.
//
// Wire bytes per rank: (W-1)/W·M·P·4 in + the same out — the reduce-scatter/all-gather minimum; the reference moves
// N pickled state_dicts of ALL M models to and from rank 0 every round.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace fdb {

struct PeerAggParams {
    const float* cp;      // client arena [C_arena, M, P]; row c of the local client list is cidx[c] (or c when cidx == nullptr)
    const int* cidx;      // [C] arena rows of this rank's clients, or nullptr
    const float* n;       // [C, M] local weights
    float* part[8];       // part[r]: rank r's symmetric partial buffer [M, P] (part[rank] is local)
    float* theta[8];      // theta[r]: rank r's symmetric model buffer  [M, theta_stride]
    float* tot_inbox[8];  // tot_inbox[r]: rank r's [world, M] weight-total inbox
    float* mc_part;       // NVLS: multicast alias of the partial buffers (nullptr → peer loads)
    float* mc_theta;      // NVLS: multicast alias of the θ buffers      (nullptr → peer stores)
    unsigned* flags[8];   // flags[r]: rank r's [2 + nchunks, world] epoch words: slot 0 totals, slot 1 final, slot 2+k chunk k
    unsigned* chunk_done; // local [nchunks] monotonic counters (producer CTAs that finished the chunk)
    int n_prod;           // producer CTAs (the rest of the grid consumes)
    int chunk4;           // chunk length in float4
    unsigned launch_idx;  // number of earlier launches (chunk_done holds launch_idx · n_prod before this one)
    unsigned* grid_sync;  // local monotonically increasing grid-barrier counter
    unsigned epoch;       // this launch's epoch (monotonic across launches)
    unsigned grid_base;   // grid_sync value expected before this launch
    int C, M, P, theta_stride, world, rank;
    long long spin_timeout_ns;
    int* error_flag;
};

FDB_DEVICE void grid_barrier(unsigned* counter, unsigned target, long long timeout_ns, int* err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        SpinGuard g;
        while ((int)(*reinterpret_cast<volatile unsigned*>(counter) - target) < 0) {
            if (g.expired(timeout_ns)) { if (err) atomicExch(err, 3); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

FDB_DEVICE void peer_barrier(const PeerAggParams& p, int slot, unsigned epoch) {
    // every rank publishes `epoch` into word [slot, rank] of all peers, then waits for all of its own words
    if (blockIdx.x == 0) {
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + slot * p.world + p.rank, epoch);
    }
    if ((int)threadIdx.x < p.world) {
        const unsigned* f = p.flags[p.rank] + slot * p.world + threadIdx.x;
        SpinGuard g;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if (g.expired(p.spin_timeout_ns)) { if (p.error_flag) atomicExch(p.error_flag, 4); break; }
        }
    }
    __syncthreads();
}

FDB_DEVICE float4 ld_peer_f4(const float* ptr) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
    return v;
}
FDB_DEVICE void st_peer_f4(float* ptr, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NVLS (NVLink SHARP): one instruction reduces the same address on every GPU inside the switch / broadcasts a store
FDB_DEVICE float4 multimem_ld_reduce_f4(const float* mc_ptr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_ptr) : "memory");
    return v;
}
FDB_DEVICE void multimem_st_f4(float* mc_ptr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

constexpr int kAggProd = 384, kAggCons = 128;   // per-CTA warp roles: 12 producer warps (HBM stream) + 4 consumer warps (NVLink)

FDB_DEVICE void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(512) fedavg_reduce_apply_peer_kernel(const __grid_constant__ PeerAggParams p) {
    extern __shared__ float wsm[];  // [C] local weights of the current model, then [M] totals
    const int C = p.C, M = p.M, P = p.P, W = p.world;
    const int P4 = P >> 2;  // host guarantees P % 4 == 0 (rows are padded)
    float* tot_s = wsm + C;
    const int G = (int)gridDim.x;
    const int cpm = (P4 + p.chunk4 - 1) / p.chunk4, nchunks = M * cpm;
    const int tid = threadIdx.x;

    if (tid < kAggProd) {
        // ================================================================ producers (every CTA: the HBM stream uses all SMs)
        if (blockIdx.x == 0 && tid == 0) {
            // local weight totals → every peer's inbox, published with the slot-0 flag (C·M scalars: serial is fine)
            for (int m = 0; m < M; ++m) {
                float tot = 0.f;
                for (int c = 0; c < C; ++c) tot += p.n[c * M + m];
                for (int r = 0; r < W; ++r) st_relaxed_sys_f32(p.tot_inbox[r] + p.rank * M + m, tot);
            }
            __threadfence_system();
            for (int r = 0; r < W; ++r) st_release_sys(p.flags[r] + 0 * W + p.rank, p.epoch);
        }
        float* mine = p.part[p.rank];
        int cur_m = -1;
        const size_t cstride = (size_t)M * P;
        for (int ck = 0; ck < nchunks; ++ck) {
            const int m = ck / cpm, k = ck - m * cpm;
            if (m != cur_m) {
                named_bar(1, kAggProd);
                for (int c = tid; c < C; c += kAggProd) wsm[c] = p.n[c * M + m];
                named_bar(1, kAggProd);
                cur_m = m;
            }
            const int lo = k * p.chunk4, hi = min(P4, lo + p.chunk4);
            const float* base = p.cp + (size_t)m * P;
            // two column groups per iteration × 4 client rows each = 8 independent 16-byte streaming loads in flight per thread
            const int cstep = G * kAggProd;
            for (int i = lo + blockIdx.x * kAggProd + tid; i < hi; i += 2 * cstep) {
                const bool two = i + cstep < hi;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                int c = 0;
                for (; c + 4 <= C; c += 4) {
                    float4 v0[4], v1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int row = p.cidx ? p.cidx[c + u] : c + u;
                        const float4* src = reinterpret_cast<const float4*>(base + (size_t)row * cstride);
                        v0[u] = __ldcs(src + i);
                        v1[u] = two ? __ldcs(src + i + cstep) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float w = wsm[c + u];
                        a0.x = fmaf(v0[u].x, w, a0.x); a0.y = fmaf(v0[u].y, w, a0.y); a0.z = fmaf(v0[u].z, w, a0.z); a0.w = fmaf(v0[u].w, w, a0.w);
                        a1.x = fmaf(v1[u].x, w, a1.x); a1.y = fmaf(v1[u].y, w, a1.y); a1.z = fmaf(v1[u].z, w, a1.z); a1.w = fmaf(v1[u].w, w, a1.w);
                    }
                }
                for (; c < C; ++c) {
                    const int row = p.cidx ? p.cidx[c] : c;
                    const float4* src = reinterpret_cast<const float4*>(base + (size_t)row * cstride);
                    const float4 v0 = __ldcs(src + i);
                    const float4 v1 = two ? __ldcs(src + i + cstep) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float w = wsm[c];
                    a0.x = fmaf(v0.x, w, a0.x); a0.y = fmaf(v0.y, w, a0.y); a0.z = fmaf(v0.z, w, a0.z); a0.w = fmaf(v0.w, w, a0.w);
                    a1.x = fmaf(v1.x, w, a1.x); a1.y = fmaf(v1.y, w, a1.y); a1.z = fmaf(v1.z, w, a1.z); a1.w = fmaf(v1.w, w, a1.w);
                }
                reinterpret_cast<float4*>(mine + (size_t)m * P)[i] = a0;
                if (two) reinterpret_cast<float4*>(mine + (size_t)m * P)[i + cstep] = a1;
            }
            // chunk complete on this CTA; the LAST CTA to finish it publishes the chunk to every peer
            __threadfence();
            named_bar(1, kAggProd);
            if (tid == 0) {
                const unsigned old = atomicAdd(p.chunk_done + ck, 1u);
                if (old == p.launch_idx * (unsigned)G + (unsigned)G - 1u) {
                    __threadfence_system();
                    for (int r = 0; r < W; ++r) st_release_sys(p.flags[r] + (2 + ck) * W + p.rank, p.epoch);
                }
            }
        }
    } else {
        // ================================================================ consumers (4 warps of every CTA: NVLink has few bytes
        //                                                                  in flight per SM, it needs breadth, not SM-exclusive CTAs)
        const int ct = tid - kAggProd;
        if (ct < W) {   // weight totals of every rank
            const unsigned* f = p.flags[p.rank] + 0 * W + ct;
            SpinGuard g;
            while ((int)(ld_acquire_sys(f) - p.epoch) < 0) {
                if (g.expired(p.spin_timeout_ns)) { if (p.error_flag) atomicExch(p.error_flag, 4); break; }
            }
        }
        named_bar(2, kAggCons);
        for (int m = ct; m < M; m += kAggCons) {
            float t = 0.f;
            for (int r = 0; r < W; ++r) t += ld_relaxed_sys_f32(p.tot_inbox[p.rank] + r * M + m);
            tot_s[m] = t;
        }
        named_bar(2, kAggCons);
        const int stride = G * kAggCons;
        for (int ck = 0; ck < nchunks; ++ck) {
            const int m = ck / cpm, k = ck - m * cpm;
            const float tot = tot_s[m];
            if (!(tot > 0.f)) continue;       // unused cluster: leave θ untouched everywhere (uniform)
            if (ct < W) {                     // chunk k of EVERY rank's partial buffer is complete
                const unsigned* f = p.flags[p.rank] + (2 + ck) * W + ct;
                SpinGuard g;
                while ((int)(ld_acquire_sys(f) - p.epoch) < 0) {
                    if (g.expired(p.spin_timeout_ns)) { if (p.error_flag) atomicExch(p.error_flag, 5); break; }
                }
            }
            named_bar(2, kAggCons);
            const float inv = 1.0f / tot;
            const int clo = k * p.chunk4, chi = min(P4, clo + p.chunk4);
            const int per = (chi - clo + W - 1) / W, lo = clo + p.rank * per, hi = min(chi, lo + per);   // my 1/W of the chunk
            if (p.mc_part != nullptr) {
                // NVLS: the switch adds the W partials (multimem.ld_reduce) and replicates the finished piece into every θ.
                // A multimem round trip is several µs: 4 independent reductions in flight per thread keep the links busy
                for (int i0 = lo + blockIdx.x * kAggCons + ct; i0 < hi; i0 += 4 * stride) {
                    float4 acc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * stride < hi) acc[u] = multimem_ld_reduce_f4(p.mc_part + (size_t)m * P + (size_t)(i0 + u * stride) * 4);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * stride < hi) {
                            acc[u].x *= inv; acc[u].y *= inv; acc[u].z *= inv; acc[u].w *= inv;
                            multimem_st_f4(p.mc_theta + (size_t)m * p.theta_stride + (size_t)(i0 + u * stride) * 4, acc[u]);
                        }
                }
            } else {
                for (int i0 = lo + blockIdx.x * kAggCons + ct; i0 < hi; i0 += 2 * stride) {
                    float4 acc[2];
                    float4 v[2][8];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (r < W && i0 + u * stride < hi) v[u][r] = ld_peer_f4(p.part[r] + (size_t)m * P + (size_t)(i0 + u * stride) * 4);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (i0 + u * stride >= hi) continue;
                        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (r < W) { acc[u].x += v[u][r].x; acc[u].y += v[u][r].y; acc[u].z += v[u][r].z; acc[u].w += v[u][r].w; }
                        acc[u].x *= inv; acc[u].y *= inv; acc[u].z *= inv; acc[u].w *= inv;
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (r < W) st_peer_f4(p.theta[r] + (size_t)m * p.theta_stride + (size_t)(i0 + u * stride) * 4, acc[u]);
                    }
                }
            }
        }
    }
    // ---- θ complete on every rank before anyone leaves (and before the next launch may overwrite the partial buffers)
    grid_barrier(p.grid_sync, p.grid_base + gridDim.x, p.spin_timeout_ns, p.error_flag);
    peer_barrier(p, 1, p.epoch);
}

int fedavg_reduce_apply_peer_launch(const float* cp, const int* cidx, const float* n, int C, int M, int P, int theta_stride, int world, int rank,
                                    const long long* part_ptrs, const long long* theta_ptrs, const long long* tot_ptrs,
                                    const long long* flag_ptrs, long long mc_part, long long mc_theta, unsigned* grid_sync, unsigned* chunk_done,
                                    int max_chunks, unsigned launch_idx, unsigned epoch, unsigned grid_base, int grid, long long timeout_ms,
                                    int* error_flag, cudaStream_t stream) {
    if (world < 1 || world > 8 || (P & 3) || grid < 2) return -5;
    PeerAggParams p{};
    p.cp = cp; p.cidx = cidx; p.n = n; p.C = C; p.M = M; p.P = P; p.theta_stride = theta_stride; p.world = world; p.rank = rank;
    for (int r = 0; r < world; ++r) {
        p.part[r] = reinterpret_cast<float*>(part_ptrs[r]);
        p.theta[r] = reinterpret_cast<float*>(theta_ptrs[r]);
        p.tot_inbox[r] = reinterpret_cast<float*>(tot_ptrs[r]);
        p.flags[r] = reinterpret_cast<unsigned*>(flag_ptrs[r]);
    }
    p.mc_part = reinterpret_cast<float*>(mc_part); p.mc_theta = reinterpret_cast<float*>(mc_theta);
    p.grid_sync = grid_sync; p.epoch = epoch; p.grid_base = grid_base;
    p.chunk_done = chunk_done; p.launch_idx = launch_idx;
    // every CTA hosts both roles; chunks: each producer thread streams 6 float4 columns per chunk (a chunk barrier costs a
    // fence + an atomic, chunks must amortise it: ~5 MB of partial sums on 148 SMs), capped by the allocated flag slots
    p.n_prod = grid;
    const int P4 = P >> 2;
    int chunk4 = grid * 384 * 6;
    while ((long long)M * ((P4 + chunk4 - 1) / chunk4) > max_chunks) chunk4 *= 2;
    p.chunk4 = chunk4;
    p.spin_timeout_ns = timeout_ms * 1000000LL; p.error_flag = error_flag;
    const int smem = (C + M + 8) * (int)sizeof(float);
    void* args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)fedavg_reduce_apply_peer_kernel, dim3(grid), dim3(512), args, smem, stream);
    return e == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
