// K14 (FedGKT distillation loss), K15 (vertical-FL logit sum + BCE gradient), K16 (GroupNorm forward).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace fdb {

// loss = T² · mean_b Σ_k p_t (log(p_t + 1e-7) - log_softmax(s/T));   grad_s = T · (softmax(s/T) · Σp_t - p_t) / B
// (reference: fedgkt/utils.py:75-94 KL_Loss — three eager softmax/log kernels + a reduction per batch).
__global__ void __launch_bounds__(256) kd_kl_kernel(const float* __restrict__ s, const float* __restrict__ t, int B, int K, float T,
                                                    float* __restrict__ loss1, float* __restrict__ grad_s) {
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float lacc = 0.f;
    const float invT = 1.f / T;
    for (int row = blockIdx.x * wpb + wib; row < B; row += gridDim.x * wpb) {
        const float* zs = s + (size_t)row * K;
        const float* zt = t + (size_t)row * K;
        float ms = -INFINITY, mt = -INFINITY;
        for (int k = lane; k < K; k += 32) { ms = fmaxf(ms, zs[k] * invT); mt = fmaxf(mt, zt[k] * invT); }
        ms = warp_max(ms); mt = warp_max(mt);
        float ss = 0.f, st = 0.f;
        for (int k = lane; k < K; k += 32) { ss += expf(zs[k] * invT - ms); st += expf(zt[k] * invT - mt); }
        ss = warp_sum(ss); st = warp_sum(st);
        const float lss = logf(ss);
        float l = 0.f, sum_pt = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float pt = expf(zt[k] * invT - mt) / st + 1e-7f;   // the reference adds 1e-7 to the teacher probs
            const float ls = zs[k] * invT - ms - lss;
            l += pt * (logf(pt) - ls);
            sum_pt += pt;
        }
        l = warp_sum(l); sum_pt = warp_sum(sum_pt);
        if (grad_s)
            for (int k = lane; k < K; k += 32) {
                const float pt = expf(zt[k] * invT - mt) / st + 1e-7f;
                const float ps = expf(zs[k] * invT - ms) / ss;
                grad_s[(size_t)row * K + k] = T * (ps * sum_pt - pt) / (float)B;
            }
        if (lane == 0) lacc += l;
    }
    lacc = block_sum(lacc, red);
    if (threadIdx.x == 0) atomicAdd(loss1, lacc * T * T / (float)B);
}
int kd_kl_launch(const float* s, const float* t, int B, int K, float T, float* loss1, float* grad_s, cudaStream_t stream) {
    cudaMemsetAsync(loss1, 0, sizeof(float), stream);
    kd_kl_kernel<<<max(1, min((B + 7) / 8, 148 * 8)), 256, 0, stream>>>(s, t, B, K, T, loss1, grad_s);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// z_b = Σ_k parts[k, b]; loss = mean BCE-with-logits; grad_b = (σ(z_b) - y_b) / B   (guest_trainer.py:85-104)
__global__ void __launch_bounds__(256) vfl_bce_kernel(const float* __restrict__ parts, const float* __restrict__ y, int Kp, int B,
                                                      float* __restrict__ loss1, float* __restrict__ grad) {
    __shared__ float red[32];
    float lacc = 0.f;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float z = 0.f;
        for (int k = 0; k < Kp; ++k) z += parts[(size_t)k * B + b];
        const float yy = y[b];
        lacc += fmaxf(z, 0.f) - z * yy + log1pf(expf(-fabsf(z)));
        grad[b] = (1.f / (1.f + expf(-z)) - yy) / (float)B;
    }
    lacc = block_sum(lacc, red);
    if (threadIdx.x == 0) atomicAdd(loss1, lacc / (float)B);
}
int vfl_bce_launch(const float* parts, const float* y, int K, int B, float* loss1, float* grad, cudaStream_t stream) {
    cudaMemsetAsync(loss1, 0, sizeof(float), stream);
    vfl_bce_kernel<<<max(1, min((B + 255) / 256, 148 * 4)), 256, 0, stream>>>(parts, y, K, B, loss1, grad);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// GroupNorm forward, NCHW: one CTA per (n, group); two passes over the group's contiguous C/G·HW slab
// (mean/var with fp32 Welford-free two-moment sum, then normalise + affine).  The reference reshapes into
// F.batch_norm (model/cv/group_normalization.py:35-40) — 4 eager kernels + 2 reshapes.
__global__ void __launch_bounds__(256) group_norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                                             const float* __restrict__ b, int C, int HW, int G, float eps,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    __shared__ float red[32];
    const int n = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G;
    const size_t base = ((size_t)n * C + (size_t)g * cpg) * HW;
    const int len = cpg * HW;
    float s = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < len; i += blockDim.x) { const float v = x[base + i]; s += v; s2 = fmaf(v, v, s2); }
    s = block_sum(s, red); s2 = block_sum(s2, red);
    const float mean = s / (float)len;
    const float var = fmaxf(s2 / (float)len - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    if (mean_out && threadIdx.x == 0) { mean_out[blockIdx.x] = mean; rstd_out[blockIdx.x] = rstd; }   // saved for the backward kernel
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int c = g * cpg + i / HW;
        const float gamma = w ? w[c] : 1.f, beta = b ? b[c] : 0.f;
        y[base + i] = (x[base + i] - mean) * rstd * gamma + beta;
    }
}
int group_norm_fwd_launch(const float* x, float* y, const float* w, const float* b, int N, int C, int HW, int G, float eps,
                          cudaStream_t stream, float* mean_out, float* rstd_out) {
    group_norm_fwd_kernel<<<N * G, 256, 0, stream>>>(x, y, w, b, C, HW, G, eps, mean_out, rstd_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// GroupNorm backward, NCHW: one CTA per (n, group).  Pass 1 walks the group's channels: per-channel Σdy and Σdy·x̂ (written
// as [N, C] partials for dγ/dβ — summed over N by the caller, deterministic) and the group sums Σγ·dy, Σγ·dy·x̂; pass 2 writes
// dx = rstd·(γ·dy − mean_g(γ·dy) − x̂·mean_g(γ·dy·x̂)).  (Reference: autograd through the F.batch_norm reshaping trick,
// model/cv/group_normalization.py:35-40 — ~10 eager kernels.)
__global__ void __launch_bounds__(256) group_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w,
                                                             const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                             float* __restrict__ dx, float* __restrict__ dg_part, float* __restrict__ db_part,
                                                             int C, int HW, int G) {
    __shared__ float red[32];
    const int n = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G;
    const size_t base = ((size_t)n * C + (size_t)g * cpg) * HW;
    const float mean = mean_in[blockIdx.x], rstd = rstd_in[blockIdx.x];
    float s1 = 0.f, s2 = 0.f;   // Σ γ·dy, Σ γ·dy·x̂ over the group (identical in every thread after the block sums)
    for (int cc = 0; cc < cpg; ++cc) {
        const int c = g * cpg + cc;
        const float gamma = w ? w[c] : 1.f;
        float a = 0.f, bsum = 0.f;
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
            const float d = dy[base + (size_t)cc * HW + i];
            const float xh = (x[base + (size_t)cc * HW + i] - mean) * rstd;
            a += d; bsum = fmaf(d, xh, bsum);
        }
        a = block_sum(a, red); bsum = block_sum(bsum, red);
        if (threadIdx.x == 0) { db_part[(size_t)n * C + c] = a; dg_part[(size_t)n * C + c] = bsum; }
        s1 = fmaf(gamma, a, s1); s2 = fmaf(gamma, bsum, s2);
    }
    const float inv = 1.f / (float)(cpg * HW);
    const float m1 = s1 * inv, m2 = s2 * inv;
    for (int i = threadIdx.x; i < cpg * HW; i += blockDim.x) {
        const int c = g * cpg + i / HW;
        const float gamma = w ? w[c] : 1.f;
        const float xh = (x[base + i] - mean) * rstd;
        dx[base + i] = rstd * (gamma * dy[base + i] - m1 - xh * m2);
    }
}
int group_norm_bwd_launch(const float* x, const float* dy, const float* w, const float* mean, const float* rstd, float* dx, float* dg_part,
                          float* db_part, int N, int C, int HW, int G, cudaStream_t stream) {
    group_norm_bwd_kernel<<<N * G, 256, 0, stream>>>(x, dy, w, mean, rstd, dx, dg_part, db_part, C, HW, G);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ------------------------------------------------------------------------------------------------ NHWC batch normalisation
// Training-mode BatchNorm2d over a channels_last tensor viewed as X[rows = N·H·W][C] — the normalisation of the pair-stacked
// networks (sim/stacked.py), where C = npairs·channels runs into the ten-thousands while rows shrink to a few hundred (cuDNN's
// NHWC kernels take 200–300 µs per call there).  Two passes per direction, both coalesced along C:
//   stats : grid (C/32, row splits), block 32 channels × 8 row lanes; per-channel Σx, Σx² (forward) or Σdy, Σdy·x̂ (backward)
//           reduced through shared memory and atomically added into a zeroed [2][C] buffer;
//   apply : elementwise y = (x − μ)·rstd·γ + β (also writes μ / rstd and the running statistics), or
//           dx = γ·rstd·(dy − Σdy/R − x̂·Σdy·x̂/R).
__global__ void __launch_bounds__(256) bn_nhwc_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ sums, long long rows, int C,
                                                            int rows_per_split) {
    __shared__ float sh[2][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const long long r0 = (long long)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
    float s0 = 0.f, s1 = 0.f;
    if (c < C) {
        if (dy == nullptr) {
            // sums of the data SHIFTED by the channel's first value (the same shift in every row split): E[(x−K)²] − E[x−K]² does
            // not cancel catastrophically when the variance is small against the mean
            const float K = __ldg(x + c);
            for (long long r = r0 + rl; r < r1; r += 8) { const float v = __ldg(x + r * C + c) - K; s0 += v; s1 = fmaf(v, v, s1); }
        } else {
            const float mu = mean[c], rs = rstd[c];
            for (long long r = r0 + rl; r < r1; r += 8) {
                const float g = __ldg(dy + r * C + c), xh = (__ldg(x + r * C + c) - mu) * rs;
                s0 += g; s1 = fmaf(g, xh, s1);
            }
        }
    }
    sh[0][rl][cl] = s0; sh[1][rl][cl] = s1;
    __syncthreads();
    if (rl < 2 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[rl][i][cl];
        atomicAdd(sums + (size_t)rl * C + c, t);
    }
}
// apply kernels: thread = one channel (coalesced along C), blockIdx.y strides over the rows — no index arithmetic per element
__global__ void __launch_bounds__(256) bn_nhwc_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ sums, const float* __restrict__ w,
                                                                const float* __restrict__ b, float* __restrict__ y, float* __restrict__ mean,
                                                                float* __restrict__ rstd, float* __restrict__ run_mean, float* __restrict__ run_var,
                                                                long long rows, int C, float eps, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float inv = 1.f / (float)rows;
    const float ms = sums[c] * inv;                                  // mean of the shifted data
    const float mu = ms + __ldg(x + c);
    const float var = fmaxf(fmaf(-ms, ms, sums[C + c] * inv), 0.f);
    const float rs = rsqrtf(var + eps);
    const float g = (w ? w[c] : 1.f) * rs, be = fmaf(-mu, g, b ? b[c] : 0.f);      // y = x·g + be
    if (blockIdx.y == 0) {                             // publish the statistics once per channel
        mean[c] = mu; rstd[c] = rs;
        if (run_mean) {
            const float unb = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
            run_mean[c] = fmaf(momentum, mu - run_mean[c], run_mean[c]);
            run_var[c] = fmaf(momentum, unb - run_var[c], run_var[c]);
        }
    }
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) y[r * C + c] = fmaf(__ldg(x + r * C + c), g, be);
}
__global__ void __launch_bounds__(256) bn_nhwc_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ sums,
                                                                const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, long long rows, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float inv = 1.f / (float)rows;
    const float mu = mean[c], rs = rstd[c], sdy = sums[c], sdyx = sums[C + c];
    const float k = (w ? w[c] : 1.f) * rs, a = sdy * inv, bb = sdyx * inv;
    if (blockIdx.y == 0) { if (dw) dw[c] = sdyx; if (db) db[c] = sdy; }
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
        const float xh = (__ldg(x + r * C + c) - mu) * rs;
        dx[r * C + c] = k * (__ldg(dy + r * C + c) - a - xh * bb);
    }
}
static void bn_stats_launch(const float* x, const float* dy, const float* mean, const float* rstd, float* sums, long long rows, int C, cudaStream_t stream) {
    const int cblocks = (C + 31) / 32;
    int splits = (int)std::max<long long>(1, std::min<long long>((2 * 148 + cblocks - 1) / cblocks, (rows + 63) / 64));
    const int rps = (int)((rows + splits - 1) / splits);
    splits = (int)((rows + rps - 1) / rps);
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)C, stream);
    bn_nhwc_stats_kernel<<<dim3(cblocks, splits), 256, 0, stream>>>(x, dy, mean, rstd, sums, rows, C, rps);
}
int bn_nhwc_fwd_launch(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, float* run_mean, float* run_var,
                       float* sums, long long rows, int C, float eps, float momentum, cudaStream_t stream) {
    if (rows <= 0 || C <= 0) return -5;
    bn_stats_launch(x, nullptr, nullptr, nullptr, sums, rows, C, stream);
    const int cb = (C + 255) / 256;
    const dim3 grid(cb, (unsigned)std::max<long long>(1, std::min<long long>(rows, (148 * 8 + cb - 1) / cb)));
    bn_nhwc_fwd_apply_kernel<<<grid, 256, 0, stream>>>(x, sums, w, b, y, mean, rstd, run_mean, run_var, rows, C, eps, momentum);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
int bn_nhwc_bwd_launch(const float* x, const float* dy, const float* w, const float* mean, const float* rstd, float* dx, float* dw, float* db,
                       float* sums, long long rows, int C, cudaStream_t stream) {
    if (rows <= 0 || C <= 0) return -5;
    bn_stats_launch(x, dy, mean, rstd, sums, rows, C, stream);
    const int cb = (C + 255) / 256;
    const dim3 grid(cb, (unsigned)std::max<long long>(1, std::min<long long>(rows, (148 * 8 + cb - 1) / cb)));
    bn_nhwc_bwd_apply_kernel<<<grid, 256, 0, stream>>>(x, dy, sums, w, mean, rstd, dx, dw, db, rows, C);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
