// Shared device helpers for the feddrift_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#define FDB_HOST_DEVICE __host__ __device__ __forceinline__
#define FDB_DEVICE __device__ __forceinline__

namespace fdb {

// ---------------------------------------------------------------- counter-based RNG (same as ops/reference.py)
FDB_HOST_DEVICE uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
FDB_HOST_DEVICE uint32_t batch_hash(uint32_t seed, uint32_t rnd, uint32_t client, uint32_t model, uint32_t step) {
    uint32_t h = mix32(seed + 0x9E3779B9u * (rnd + 1u));
    h = mix32(h ^ (client * 0x85EBCA6Bu + 0x165667B1u));
    h = mix32(h ^ (model * 0xC2B2AE35u + 0x27D4EB2Fu));
    h = mix32(h ^ (step * 0x2545F491u + 1u));
    return h;
}
FDB_DEVICE uint32_t hash_choice(uint32_t h, uint32_t n) { return __umulhi(h, n); }

// ---------------------------------------------------------------- warp reductions
FDB_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
FDB_DEVICE double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
FDB_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in every thread
template <typename T>
FDB_DEVICE T block_sum(T v, T* smem /* >= 32 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T r = (lane < nw) ? smem[lane] : T(0);
    r = warp_sum(r);
    return r;
}

// ---------------------------------------------------------------- cross-GPU flag helpers (system scope)
FDB_DEVICE void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
FDB_DEVICE unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
FDB_DEVICE void st_relaxed_sys_f32(float* p, float v) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
FDB_DEVICE float ld_relaxed_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// LL ("low latency") words: payload and epoch tag travel in ONE 8- or 16-byte store, so the receiver needs neither a
// fence nor a separate flag — it polls the word until the tag matches.  (8-byte stores are single-copy atomic; the 16-byte
// variant carries the tag twice and the reader checks both halves, as NCCL's LL lines do.)
FDB_DEVICE void st_ll(uint2* p, float v, unsigned tag) {
    asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
FDB_DEVICE uint2 ld_ll(const uint2* p) {
    uint2 w;
    asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p) : "memory");
    return w;
}
FDB_DEVICE void st_ll2(uint4* p, float a, float b, unsigned tag) {
    asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag),
                 "r"(__float_as_uint(b)), "r"(tag) : "memory");
}
FDB_DEVICE uint4 ld_ll2(const uint4* p) {
    uint4 w;
    asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
FDB_DEVICE long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded spinning without a timer read on the fast path: %globaltimer is a slow, chip-level register — reading it before and
// inside every wait put hundreds of cycles on each hop of the mbarrier / flag handshake chains (the pipeline skeleton of the
// tcgen05 kernels alone cost ~0.4 µs per k-block).  The guard looks at the timer only every 256 failed polls; the first look
// starts the clock.
struct SpinGuard {
    long long t0 = 0;
    unsigned n = 0;
    FDB_DEVICE bool expired(long long timeout_ns) {
        if ((++n & 255u) != 0) return false;
        const long long now = globaltimer_ns();
        if (t0 == 0) { t0 = now; return false; }
        return now - t0 > timeout_ns;
    }
};

}  // namespace fdb
