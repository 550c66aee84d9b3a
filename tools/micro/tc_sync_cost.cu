// Cost (SM cycles, single thread) of the synchronisation primitives on the per-k-block path of the tcgen05 mainloop:
// clock64, mbarrier.try_wait (already complete), mbarrier.arrive, tcgen05.fence::after_thread_sync, tcgen05.commit (issue cost,
// back to back), and the latency commit → mbarrier phase flip.   nvcc -gencode arch=compute_100a,code=sm_100a -o tc_sync_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// mode 0: thread 0 alone; mode 1: warps 4-7 (all lanes) spin on a barrier that never completes, like the GEMM's epilogue warps
// waiting for an accumulator; mode 2: only lane 0 of warps 4-7 spins; mode 3: all lanes spin with __nanosleep(64) back-off
__global__ void k(long long* out, int N, int mode) {
    __shared__ uint64_t bar[5];
    __shared__ volatile int stop;
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 128 && mode != 0 && (mode != 2 || (threadIdx.x & 31) == 0)) {
        while (!stop) {
            if (try_wait(bar + 4, 0)) break;      // never completes
            if (mode == 3) __nanosleep(64);
        }
    }
    if (threadIdx.x == 0) {
        long long t0, t1;
        // (0) clock64 back to back
        t0 = clock64();
        long long acc = 0;
        for (int i = 0; i < N; ++i) acc += clock64();
        t1 = clock64();
        out[0] = (t1 - t0) / N; out[15] = acc;
        // (1) try_wait on a completed phase (parity 1 of a fresh barrier = "previous phase complete")
        t0 = clock64();
        int okc = 0;
        for (int i = 0; i < N; ++i) okc += try_wait(bar + 0, 1);
        t1 = clock64();
        out[1] = (t1 - t0) / N; out[14] = okc;
        // (2) tcgen05.fence::after_thread_sync
        t0 = clock64();
        for (int i = 0; i < N; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t1 = clock64();
        out[2] = (t1 - t0) / N;
        // (3) mbarrier.arrive (count 1 → flips the phase each time)
        t0 = clock64();
        for (int i = 0; i < N; ++i) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 1)) : "memory");
        t1 = clock64();
        out[3] = (t1 - t0) / N;
        // (4) tcgen05.commit issue cost, back to back, nobody waits
        t0 = clock64();
        for (int i = 0; i < N; ++i)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 2)) : "memory");
        t1 = clock64();
        out[4] = (t1 - t0) / N;
        // drain: wait until all N arrives landed (phase parity after N flips)
        // (5) commit → wait for the flip → commit …  (round-trip latency of one commit)
        // first synchronise on bar[3]
        uint32_t ph = 0;
        t0 = clock64();
        for (int i = 0; i < N; ++i) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 3)) : "memory");
            while (!try_wait(bar + 3, ph)) {}
            ph ^= 1;
        }
        t1 = clock64();
        out[5] = (t1 - t0) / N;
        // (6) the per-k-block sequence of the GEMM's MMA thread without MMAs: try_wait(ok) + fence + commit
        t0 = clock64();
        for (int i = 0; i < N; ++i) {
            okc += try_wait(bar + 0, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 2)) : "memory");
        }
        t1 = clock64();
        out[6] = (t1 - t0) / N; out[13] = okc;
        stop = 1;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_slot));
}
int main() {
    long long* d; cudaMalloc(&d, 16 * sizeof(long long));
    const int N = 2000;
  for (int mode = 0; mode < 4; ++mode) {
    k<<<1, 256>>>(d, N, mode); cudaDeviceSynchronize();
    k<<<1, 256>>>(d, N, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"spinners\": %d, \"err\": \"%s\", \"clock64\": %lld, \"try_wait_ok\": %lld, \"tcgen05_fence_after\": %lld, \"mbar_arrive\": %lld, \"commit_issue\": %lld, "
           "\"commit_roundtrip\": %lld, \"wait_fence_commit\": %lld}\n", mode, cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
    return 0;
}
