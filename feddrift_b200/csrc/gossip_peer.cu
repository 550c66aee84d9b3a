// gossip_mix_peer — K12 on several GPUs: one decentralized-SGD / PushSum mixing step
//     x_i ← Σ_j W_ij · x_j        (j over the in-neighbours of rank i, W row-stochastic, W_ii included)
// as ONE kernel per rank with the neighbour exchange done by the kernel itself over NVLink peer memory.
//
// Every rank keeps its vector in a symmetric-memory double buffer  x[2][P]; step k reads buffer k&1 of every
// in-neighbour with 128-bit peer loads (all neighbours' loads of a thread are in flight together) and writes the local
// buffer (k+1)&1.  Synchronisation is one monotonically increasing epoch word per (rank, peer):
//   * before reading, a rank waits until every peer has published "step k-1 done" (their buffer k&1 is complete and
//     they no longer read the buffer this rank is about to overwrite);
//   * after the grid finished writing (grid-wide barrier) block 0 publishes "step k done" into every peer's flag row
//     with st.release.sys — the wait is a LOCAL spin on the rank's own flag row.
// The PushSum scalar ω rides along as element P of the vector (host side), so DSGD and PushSum are the same kernel.
// Reference: client_dsgd.py:88-102, client_pushsum.py:104-129 (python loops over neighbour tensors on one process),
// decentralized_worker_manager.py:41-46 (one MPI message per neighbour per step).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace fdb {

struct GossipParams {
    float* x[8];          // x[r]: rank r's symmetric double buffer [2][P]
    unsigned* flags[8];   // flags[r]: rank r's [world] epoch words (word j = last finished step of rank j)
    float w[8];           // this rank's mixing row (0 = no edge)
    unsigned* grid_sync;  // local grid-barrier counter (monotonic)
    unsigned grid_base, epoch;
    int P, world, rank;
    long long spin_timeout_ns;
    int* error_flag;
};

FDB_DEVICE float4 ldp4(const float* ptr) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
    return v;
}

__global__ void __launch_bounds__(512) gossip_mix_peer_kernel(const __grid_constant__ GossipParams p) {
    const int W = p.world;
    const unsigned k = p.epoch;            // 1-based step index
    const int cur = (k - 1) & 1, nxt = k & 1;
    // ---- wait: every peer finished step k-1
    if ((int)threadIdx.x < W && (int)threadIdx.x != p.rank) {
        const unsigned* f = p.flags[p.rank] + threadIdx.x;
        SpinGuard g;
        while ((int)(ld_acquire_sys(f) - (k - 1)) < 0) {
            if (g.expired(p.spin_timeout_ns)) { if (p.error_flag) atomicExch(p.error_flag, 6); break; }
        }
    }
    __syncthreads();
    // ---- mix
    const int P4 = p.P >> 2;
    float4* out = reinterpret_cast<float4*>(p.x[p.rank] + (size_t)nxt * p.P);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P4; i += gridDim.x * blockDim.x) {
        float4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < W && p.w[r] != 0.f) v[r] = ldp4(p.x[r] + (size_t)cur * p.P + 4 * (size_t)i);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < W && p.w[r] != 0.f) {
                acc.x = fmaf(p.w[r], v[r].x, acc.x); acc.y = fmaf(p.w[r], v[r].y, acc.y);
                acc.z = fmaf(p.w[r], v[r].z, acc.z); acc.w = fmaf(p.w[r], v[r].w, acc.w);
            }
        out[i] = acc;
    }
    // ---- grid-wide completion, then publish "step k done" to every peer
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p.grid_sync, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const unsigned target = p.grid_base + gridDim.x;
            SpinGuard g;
            while ((int)(*reinterpret_cast<volatile unsigned*>(p.grid_sync) - target) < 0) {
                if (g.expired(p.spin_timeout_ns)) { if (p.error_flag) atomicExch(p.error_flag, 7); break; }
            }
            __threadfence_system();
        }
        __syncthreads();
        if ((int)threadIdx.x < W) st_release_sys(p.flags[threadIdx.x] + p.rank, k);
    }
}

int gossip_mix_peer_launch(const long long* x_ptrs, const long long* flag_ptrs, const float* w, int P, int world, int rank,
                           unsigned* grid_sync, unsigned grid_base, unsigned epoch, int grid, long long timeout_ms, int* error_flag,
                           cudaStream_t stream) {
    if (world < 1 || world > 8 || (P & 3)) return -5;
    GossipParams p{};
    for (int r = 0; r < world; ++r) {
        p.x[r] = reinterpret_cast<float*>(x_ptrs[r]);
        p.flags[r] = reinterpret_cast<unsigned*>(flag_ptrs[r]);
        p.w[r] = w[r];
    }
    p.grid_sync = grid_sync; p.grid_base = grid_base; p.epoch = epoch; p.P = P; p.world = world; p.rank = rank;
    p.spin_timeout_ns = timeout_ms * 1000000LL; p.error_flag = error_flag;
    void* args[] = {&p};
    // cooperative launch: block 0 waits for the whole grid, so all CTAs must be co-resident
    cudaError_t e = cudaLaunchCooperativeKernel((void*)gossip_mix_peer_kernel, dim3(grid), dim3(512), args, 0, stream);
    return e == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
