// Evaluation reductions (K4 for arbitrary models, K7): everything the aggregators need from a batch of logits is
// reduced ON DEVICE into a few floats — no per-batch `.item()` host syncs (reference: `_infer`
// FedAvgEnsAggregatorSoftCluster.py:305-326 does two D2H syncs per batch; `_mse`/`_infer_ens`/`_confusion_matrix`
// FedAvgEnsAggregatorAue.py:216-283, FedAvgEnsAggregatorKue.py:234-302 round-trip whole prob tensors through numpy).
#include "common.cuh"
#include "kernels.h"

namespace fdb {

// one warp per row; lanes stride over classes (K up to ~10k: shakespeare 90, stackoverflow 10004)
__global__ void __launch_bounds__(256) eval_logits_kernel(const float* __restrict__ logits, const int* __restrict__ target, int B, int K,
                                                          float* __restrict__ acc3, float* __restrict__ sq_out) {
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float corr = 0.f, loss = 0.f, sq = 0.f;
    for (int row = blockIdx.x * wpb + wib; row < B; row += gridDim.x * wpb) {
        const float* z = logits + (size_t)row * K;
        float mx = -INFINITY; int am = 0x7fffffff;
        for (int k = lane; k < K; k += 32) { const float v = z[k]; if (v > mx) { mx = v; am = k; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }  // first max wins (torch.max semantics)
        }
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += expf(z[k] - mx);
        s = warp_sum(s);
        if (lane == 0) {
            const int y = target[row];
            const float zy = z[y];
            loss += logf(s) + mx - zy;
            corr += (am == y) ? 1.f : 0.f;
            const float py = expf(zy - mx) / s;
            sq += (1.f - py) * (1.f - py);
        }
    }
    corr = block_sum(corr, red); loss = block_sum(loss, red); sq = block_sum(sq, red);
    if (threadIdx.x == 0) {
        if (acc3) {
            atomicAdd(acc3 + 0, corr); atomicAdd(acc3 + 1, loss);
            if (blockIdx.x == 0) atomicAdd(acc3 + 2, (float)B);
        }
        if (sq_out) atomicAdd(sq_out, sq);
    }
}

int eval_logits_launch(const float* logits, const int* target, int B, int K, float* acc3, cudaStream_t stream) {
    const int blocks = max(1, min((B + 7) / 8, 148 * 8));
    eval_logits_kernel<<<blocks, 256, 0, stream>>>(logits, target, B, K, acc3, nullptr);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
int aue_sqerr_launch(const float* logits, const int* target, int B, int K, float* out1, cudaStream_t stream) {
    cudaMemsetAsync(out1, 0, sizeof(float), stream);
    const int blocks = max(1, min((B + 7) / 8, 148 * 8));
    eval_logits_kernel<<<blocks, 256, 0, stream>>>(logits, target, B, K, nullptr, out1);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// weighted hard vote: out[b] = argmax_c Σ_k w[k]·[preds[k,b] == c]   (classes ≤ 1024 via smem tally per thread row)
__global__ void ensemble_vote_kernel(const int* __restrict__ preds, const float* __restrict__ w, int Km, int B, int classes,
                                     int* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int best = 0; float bestv = -1.f;
    for (int c = 0; c < classes; ++c) {
        float v = 0.f;
        for (int k = 0; k < Km; ++k) v += (preds[(size_t)k * B + b] == c) ? w[k] : 0.f;
        if (v > bestv) { bestv = v; best = c; }
    }
    out[b] = best;
}
int ensemble_vote_launch(const int* preds, const float* w, int Kmodels, int B, int classes, int* out, cudaStream_t stream) {
    ensemble_vote_kernel<<<(B + 255) / 256, 256, 0, stream>>>(preds, w, Kmodels, B, classes, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// confusion-matrix histogram: smem-privatised per CTA when classes² fits, one global atomic per non-empty bin
__global__ void confusion_kernel(const int* __restrict__ pred, const int* __restrict__ target, int B, int classes, int* __restrict__ out) {
    extern __shared__ int hist[];
    const int bins = classes * classes;
    const bool priv = bins <= 12288;
    if (priv) { for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0; __syncthreads(); }
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const int idx = target[b] * classes + pred[b];
        if (priv) atomicAdd(hist + idx, 1); else atomicAdd(out + idx, 1);
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < bins; i += blockDim.x) if (hist[i]) atomicAdd(out + i, hist[i]);
    }
}
int confusion_matrix_launch(const int* pred, const int* target, int B, int classes, int* out, cudaStream_t stream) {
    const int bins = classes * classes;
    cudaMemsetAsync(out, 0, bins * sizeof(int), stream);
    const int smem = bins <= 12288 ? bins * (int)sizeof(int) : 0;
    confusion_kernel<<<max(1, min((B + 255) / 256, 148 * 4)), 256, smem, stream>>>(pred, target, B, classes, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
