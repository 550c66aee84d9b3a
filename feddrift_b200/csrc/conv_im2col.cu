// Convolution as GEMM on the tcgen05 kernel (K3b "conv via GEMM"):
//   forward   cols[B·Ho·Wo, Cin·kh·kw] (bf16) = im2col(x)            — this file, fused fp32→bf16 cast
//             y[B·Ho·Wo, Cout]        = gemm_tn(cols, W[Cout, Cin·kh·kw]) + bias (+ReLU)   — gemm_tc.cu (NHWC output)
//   backward  dcols = gemm_tn(dy[B·Ho·Wo, Cout], Wᵀ)   →   dx = col2im(dcols)   — this file (gather form, no atomics)
//             dW    = gemm_tn(dyᵀ, colsᵀ)
// The K order of a cols row is (cin, kh, kw) = `weight.view(Cout, -1)`, so the weights need no re-layout.
// Reference: nn.Conv2d on cuDNN in fp32 (fedml_api/model/cv/cnn.py:110-117, resnet*.py).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace fdb {

struct ConvGeom {
    int B, C, H, W, kh, kw, sh, sw, ph, pw, Ho, Wo;
};

// one thread per (output pixel, cin, kh) → kw contiguous bf16 elements of the cols row
__global__ void __launch_bounds__(256) im2col_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols, ConvGeom g,
                                                          long long sxb, long long sxc, long long sxh, long long sxw) {
    const long long K = (long long)g.C * g.kh * g.kw;
    const long long total = (long long)g.B * g.Ho * g.Wo * g.C * g.kh;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % g.kh);
        long long q = i / g.kh;
        const int c = (int)(q % g.C);
        q /= g.C;                                   // output pixel index (b, ho, wo)
        const int wo = (int)(q % g.Wo);
        const long long q2 = q / g.Wo;
        const int ho = (int)(q2 % g.Ho), b = (int)(q2 / g.Ho);
        const int hi = ho * g.sh - g.ph + r;
        __nv_bfloat16* dst = cols + q * K + ((long long)c * g.kh + r) * g.kw;
        const bool hok = hi >= 0 && hi < g.H;
        const float* src = x + b * sxb + c * sxc + (long long)hi * sxh;
        for (int s = 0; s < g.kw; ++s) {
            const int wi = wo * g.sw - g.pw + s;
            const float v = (hok && wi >= 0 && wi < g.W) ? __ldg(src + (long long)wi * sxw) : 0.f;
            dst[s] = __float2bfloat16(v);
        }
    }
}

// dx[b,c,hi,wi] = Σ_{r,s : (hi+ph-r) % sh == 0, (wi+pw-s) % sw == 0} dcols[(b,ho,wo), (c,r,s)]   (gather: no atomics)
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcols, float* __restrict__ dx, ConvGeom g) {
    const long long K = (long long)g.C * g.kh * g.kw;
    const long long total = (long long)g.B * g.C * g.H * g.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int wi = (int)(i % g.W);
        long long q = i / g.W;
        const int hi = (int)(q % g.H);
        q /= g.H;
        const int c = (int)(q % g.C), b = (int)(q / g.C);
        float acc = 0.f;
        for (int r = 0; r < g.kh; ++r) {
            const int hn = hi + g.ph - r;
            if (hn < 0 || hn % g.sh) continue;
            const int ho = hn / g.sh;
            if (ho >= g.Ho) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int wn = wi + g.pw - s;
                if (wn < 0 || wn % g.sw) continue;
                const int wo = wn / g.sw;
                if (wo >= g.Wo) continue;
                acc += __ldg(dcols + (((long long)b * g.Ho + ho) * g.Wo + wo) * K + ((long long)c * g.kh + r) * g.kw + s);
            }
        }
        dx[i] = acc;   // dx is NCHW contiguous
    }
}

static int grid_for(long long total) { return (int)max(1LL, min((total + 255) / 256, 148LL * 16)); }

int im2col_bf16_launch(const float* x, void* cols, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int Ho, int Wo,
                       long long sxb, long long sxc, long long sxh, long long sxw, cudaStream_t stream) {
    ConvGeom g{B, C, H, W, kh, kw, sh, sw, ph, pw, Ho, Wo};
    const long long total = (long long)B * Ho * Wo * C * kh;
    im2col_bf16_kernel<<<grid_for(total), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(cols), g, sxb, sxc, sxh, sxw);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

int col2im_launch(const float* dcols, float* dx, int B, int C, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int Ho, int Wo,
                  cudaStream_t stream) {
    ConvGeom g{B, C, H, W, kh, kw, sh, sw, ph, pw, Ho, Wo};
    col2im_kernel<<<grid_for((long long)B * C * H * W), 256, 0, stream>>>(dcols, dx, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace fdb
