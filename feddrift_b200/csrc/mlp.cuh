// Register-resident tiny MLP family (LogisticRegression / FeedForwardNN of the drift experiments).
// Parameter layout == state_dict order of feddrift_b200.models.basic (fc1.weight[HID,IN], fc1.bias[HID],
// fc2.weight[OUT,HID], fc2.bias[OUT]  |  linear.weight[OUT,IN], linear.bias[OUT]).
#pragma once
#include "common.cuh"

namespace fdb {

// fast-math intrinsics (ex2.approx / lg2.approx based; ≤ 2 ulp) — the softmax/CE of a 2-class MLP is not
// precision critical and this removes ~150 instructions of range reduction per sample from the critical path
FDB_DEVICE float fexp(float x) { return __expf(x); }
FDB_DEVICE float flog(float x) { return __logf(x); }
FDB_DEVICE float fdiv(float a, float b) { return __fdividef(a, b); }

template <int KIND, int IN, int HID, int OUT>
struct Mlp {
    static constexpr int kKind = KIND, kIn = IN, kHid = HID, kOut = OUT;
    static constexpr int P = (KIND == 0) ? (OUT * IN + OUT) : (HID * IN + HID + OUT * HID + OUT);
    static constexpr int W1 = 0, B1 = HID * IN, W2 = HID * IN + HID, B2 = HID * IN + HID + OUT * HID;  // fnn offsets
    static constexpr int LW = 0, LB = OUT * IN;                                                         // lr offsets

    // logits as fed to cross-entropy (lr applies the sigmoid first — reference quirk, lr.py:10)
    FDB_DEVICE static void forward(const float (&th)[P], const float (&x)[IN], float (&z)[OUT], float (&h)[HID > 0 ? HID : 1]) {
        if constexpr (KIND == 0) {
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                float a = th[LB + o];
#pragma unroll
                for (int i = 0; i < IN; ++i) a = fmaf(th[LW + o * IN + i], x[i], a);
                z[o] = fdiv(1.0f, 1.0f + fexp(-a));
            }
        } else {
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                float a = th[B1 + j];
#pragma unroll
                for (int i = 0; i < IN; ++i) a = fmaf(th[W1 + j * IN + i], x[i], a);
                h[j] = fmaxf(a, 0.0f);
            }
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                float a = th[B2 + o];
#pragma unroll
                for (int j = 0; j < HID; ++j) a = fmaf(th[W2 + o * HID + j], h[j], a);
                z[o] = a;
            }
        }
    }

    // softmax cross-entropy on z; returns loss, fills prob p, argmax (first max wins like torch.max)
    FDB_DEVICE static float softmax_ce(const float (&z)[OUT], int y, float (&p)[OUT], int& amax) {
        float mx = z[0];
        amax = 0;
#pragma unroll
        for (int o = 1; o < OUT; ++o)
            if (z[o] > mx) { mx = z[o]; amax = o; }
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) { p[o] = fexp(z[o] - mx); s += p[o]; }
        const float inv = fdiv(1.0f, s);
        float zy = z[0];
#pragma unroll
        for (int o = 0; o < OUT; ++o) { p[o] *= inv; if (o == y) zy = z[o]; }
        return flog(s) + mx - zy;
    }

    // accumulate d(mean CE)/dθ for one sample into g (scale = 1/batch)
    FDB_DEVICE static void backward_accum(const float (&th)[P], const float (&x)[IN], const float (&z)[OUT],
                                          const float (&h)[HID > 0 ? HID : 1], const float (&p)[OUT], int y, float scale,
                                          float (&g)[P]) {
        float dz[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) dz[o] = (p[o] - (o == y ? 1.0f : 0.0f)) * scale;
        if constexpr (KIND == 0) {
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                const float ds = dz[o] * z[o] * (1.0f - z[o]);
                g[LB + o] += ds;
#pragma unroll
                for (int i = 0; i < IN; ++i) g[LW + o * IN + i] = fmaf(ds, x[i], g[LW + o * IN + i]);
            }
        } else {
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                g[B2 + o] += dz[o];
#pragma unroll
                for (int j = 0; j < HID; ++j) g[W2 + o * HID + j] = fmaf(dz[o], h[j], g[W2 + o * HID + j]);
            }
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                float dh = 0.f;
#pragma unroll
                for (int o = 0; o < OUT; ++o) dh = fmaf(th[W2 + o * HID + j], dz[o], dh);
                dh = (h[j] > 0.f) ? dh : 0.f;
                g[B1 + j] += dh;
#pragma unroll
                for (int i = 0; i < IN; ++i) g[W1 + j * IN + i] = fmaf(dh, x[i], g[W1 + j * IN + i]);
            }
        }
    }
};

// The instantiation table.  (kind, in, hid, out); kind 0 = lr, 1 = fnn.  hidden = 2·in for fnn
// (main_fedavg.py:215).  Extend here to add a shape; dispatch is generated from the same list.
#define FDB_MLP_SHAPES(X) \
    X(1, 3, 6, 2)   /* SEA fnn      */ \
    X(1, 2, 4, 2)   /* sine/circle  */ \
    X(0, 3, 0, 2)   /* SEA lr       */ \
    X(0, 2, 0, 2)   /* sine/circle lr */ \
    X(1, 4, 8, 3)   /* generic small (tests) */ \
    X(1, 5, 10, 2)  /* generic small (tests) */

}  // namespace fdb
