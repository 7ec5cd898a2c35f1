// LSTM recurrence (persistent: one CTA walks all T steps of one clip with W_hh resident in shared memory),
// its BPTT, and the small-batch classifier head: Linear(+ReLU+Dropout) layers, sigmoid+BCE, ArcFace+CE and the
// audio-face fusion loss (CB-focal + alignment + temporal smoothness).
//
// Reference: nn.LSTM(2048,H,1,batch_first) (XceptionLSTMV.py:18-23,67-68), fc_layers/fc_out/sigmoid
// (XceptionLSTMV.py:25-44,69-70), nn.BCELoss (train_audio.py:20,39), ArcFaceHead + CrossEntropyLoss
// (train_visual.py:455-474,532), fusion head + CBFocalLoss (train_au_face.py:423-458,659-674).
// The input projection x_t W_ih^T for all T steps is one tcgen05 GEMM (gemm.cu); this file adds the biases.
#include "common.cuh"

namespace xcp {

XCP_DEVINL float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ============================================================================================ LSTM forward
// grid = B, block = NT threads (NT = min(4H, 1024)).  wt = W_hh^T as bf16 [H][4H].
// gates_out[b,t,:] = (i,f,g,o) post-activation, c_out[b,t,:], h_out[b,t,:] (fp32) (+ bf16 copy of h_{t-1}).
__global__ void __launch_bounds__(1024)
lstm_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                const __nv_bfloat16* __restrict__ wt, float* __restrict__ h_out, float* __restrict__ gates_out,
                float* __restrict__ c_out, float* __restrict__ hn, float* __restrict__ cn, int T, int H, int w_in_smem) {
    extern __shared__ uint8_t lsm[];
    float* s_h = reinterpret_cast<float*>(lsm);           // [H]
    float* s_g = s_h + H;                                 // [4H]
    __nv_bfloat16* s_w = reinterpret_cast<__nv_bfloat16*>(s_g + 4 * H);   // [H][4H] if it fits
    const int b = blockIdx.x, G = 4 * H;
    if (w_in_smem) {
        const uint4* src = reinterpret_cast<const uint4*>(wt);
        uint4* dst = reinterpret_cast<uint4*>(s_w);
        for (int i = threadIdx.x; i < H * G / 8; i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_h[i] = 0.f;
    float c_reg[4] = {0.f, 0.f, 0.f, 0.f};   // thread j (< H, strided) keeps its cell state in registers (H <= 4*blockDim)
    __syncthreads();
    const __nv_bfloat16* W = w_in_smem ? s_w : wt;
    for (int t = 0; t < T; ++t) {
        const float* xp = xproj + ((long long)b * T + t) * G;
        for (int j = threadIdx.x; j < G; j += blockDim.x) {
            float acc = xp[j] + b_ih[j] + b_hh[j];
#pragma unroll 8
            for (int k = 0; k < H; ++k) acc = fmaf(__bfloat162float(W[(long long)k * G + j]), s_h[k], acc);
            s_g[j] = acc;
        }
        __syncthreads();
        int slot = 0;
        for (int j = threadIdx.x; j < H; j += blockDim.x, ++slot) {
            const float ig = sigmoidf_(s_g[j]), fg = sigmoidf_(s_g[H + j]);
            const float gg = tanhf(s_g[2 * H + j]), og = sigmoidf_(s_g[3 * H + j]);
            const float c = fg * c_reg[slot] + ig * gg;
            c_reg[slot] = c;
            const float h = og * tanhf(c);
            const long long o = (long long)b * T + t;
            gates_out[o * G + j] = ig; gates_out[o * G + H + j] = fg; gates_out[o * G + 2 * H + j] = gg; gates_out[o * G + 3 * H + j] = og;
            c_out[o * H + j] = c;
            h_out[o * H + j] = h;
            s_h[j] = h;
            if (t == T - 1) { hn[(long long)b * H + j] = h; cn[(long long)b * H + j] = c; }
        }
        __syncthreads();
    }
}

// ============================================================================================ LSTM backward (BPTT)
// grid = B.  w = W_hh bf16 [4H][H].  Produces pre-activation gate grads dgates (bf16 [B,T,4H]) for the
// dW_ih / dW_hh / dX GEMMs, h_{t-1} as bf16 [B,T,H], and accumulates the bias gradient.
__global__ void __launch_bounds__(1024)
lstm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dhn, const float* __restrict__ dcn,
                const float* __restrict__ gates, const float* __restrict__ cst, const float* __restrict__ hst,
                const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ dgates, __nv_bfloat16* __restrict__ hprev,
                float* __restrict__ dbias_ih, float* __restrict__ dbias_hh, int T, int H) {
    extern __shared__ uint8_t lsm[];
    float* s_dg = reinterpret_cast<float*>(lsm);   // [4H] pre-activation gate grads of the current step
    float* s_dh = s_dg + 4 * H;                    // [H] recurrent dh for step t-1
    float* s_db = s_dh + H;                        // [4H] bias grad accumulator
    const int b = blockIdx.x, G = 4 * H;
    for (int i = threadIdx.x; i < G; i += blockDim.x) s_db[i] = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_dh[i] = dhn ? dhn[(long long)b * H + i] : 0.f;
    float dc_reg[4];
    {
        int slot = 0;
        for (int j = threadIdx.x; j < H; j += blockDim.x, ++slot) dc_reg[slot] = dcn ? dcn[(long long)b * H + j] : 0.f;
    }
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
        const long long o = (long long)b * T + t;
        int slot = 0;
        for (int j = threadIdx.x; j < H; j += blockDim.x, ++slot) {
            const float ig = gates[o * G + j], fg = gates[o * G + H + j], gg = gates[o * G + 2 * H + j], og = gates[o * G + 3 * H + j];
            const float c = cst[o * H + j];
            const float cprev = t > 0 ? cst[(o - 1) * H + j] : 0.f;
            const float tc = tanhf(c);
            const float dh = s_dh[j] + (dout ? dout[o * H + j] : 0.f);
            const float dc = dc_reg[slot] + dh * og * (1.f - tc * tc);
            const float di = dc * gg * ig * (1.f - ig);
            const float df = dc * cprev * fg * (1.f - fg);
            const float dgg = dc * ig * (1.f - gg * gg);
            const float dog = dh * tc * og * (1.f - og);
            dc_reg[slot] = dc * fg;
            s_dg[j] = di; s_dg[H + j] = df; s_dg[2 * H + j] = dgg; s_dg[3 * H + j] = dog;
            hprev[o * H + j] = __float2bfloat16(t > 0 ? hst[(o - 1) * H + j] : 0.f);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < G; j += blockDim.x) {
            const float v = s_dg[j];
            dgates[o * G + j] = __float2bfloat16(v);
            s_db[j] += v;
        }
        // dh_{t-1}[k] = sum_j W_hh[j][k] * dg[j]
        for (int k = threadIdx.x; k < H; k += blockDim.x) {
            float acc = 0.f;
#pragma unroll 8
            for (int j = 0; j < G; ++j) acc = fmaf(__bfloat162float(w[(long long)j * H + k]), s_dg[j], acc);
            s_dh[k] = acc;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        if (dbias_ih) atomicAdd(&dbias_ih[i], s_db[i]);
        if (dbias_hh) atomicAdd(&dbias_hh[i], s_db[i]);
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16(x[i]);
}

// ============================================================================================ small-batch Linear
// out[b][n] = act( sum_k a[b][k] W[n][k] + bias[n] ) * dropmask ; warp per output neuron, B <= 32 rows processed in
// chunks of 8.  act: 0 none, 1 relu.  mask (optional, uint8 [B][N], 1 = keep) scaled by drop_scale = 1/(1-p).
// Weight-bandwidth bound (fp32 masters, 4 MB per 1024x1024 layer): every lane issues its float4 weight loads for
// 8 K-chunks back to back before the FMAs so a warp keeps 4 KB in flight; the activations come from L1.
constexpr int MAXB = 32;
constexpr int LIN_BCH = 8;
__global__ void __launch_bounds__(256)
linear_small_fwd_kernel(const float* __restrict__ a, const float* __restrict__ W, const float* __restrict__ bias,
                        const uint8_t* __restrict__ mask, float drop_scale, int act, float* __restrict__ out, int B, int N, int K) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* wrow = W + (long long)n * K;
    for (int b0 = 0; b0 < B; b0 += LIN_BCH) {
        float acc[LIN_BCH];
#pragma unroll
        for (int b = 0; b < LIN_BCH; ++b) acc[b] = 0.f;
        if ((K & 3) == 0) {
            const int K4 = K >> 2;
            for (int c0 = 0; c0 < K4; c0 += 8 * 32) {
                float4 wv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j * 32 + lane;
                    wv[j] = c < K4 ? __ldg(reinterpret_cast<const float4*>(wrow) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int b = 0; b < LIN_BCH; ++b) {
                    if (b0 + b < B) {
                        const float4* ar = reinterpret_cast<const float4*>(a + (long long)(b0 + b) * K);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = c0 + j * 32 + lane;
                            if (c < K4) {
                                const float4 av = ar[c];
                                acc[b] = fmaf(wv[j].x, av.x, acc[b]); acc[b] = fmaf(wv[j].y, av.y, acc[b]);
                                acc[b] = fmaf(wv[j].z, av.z, acc[b]); acc[b] = fmaf(wv[j].w, av.w, acc[b]);
                            }
                        }
                    }
                }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                const float wv = wrow[k];
#pragma unroll
                for (int b = 0; b < LIN_BCH; ++b)
                    if (b0 + b < B) acc[b] = fmaf(wv, a[(long long)(b0 + b) * K + k], acc[b]);
            }
        }
#pragma unroll
        for (int b = 0; b < LIN_BCH; ++b) {
            if (b0 + b < B) {
                float v = warp_sum(acc[b]);
                if (lane == 0) {
                    v += bias ? bias[n] : 0.f;
                    if (act == 1) v = fmaxf(v, 0.f);
                    if (mask) v = mask[(long long)(b0 + b) * N + n] ? v * drop_scale : 0.f;
                    out[(long long)(b0 + b) * N + n] = v;
                }
            }
        }
    }
}

// Backward of the layer above.  delta_raw = dL/d(out) ; out_act = saved layer output (null for a linear output);
// effective delta = delta_raw * (out_act > 0 ? drop_scale : 0).  dW += delta^T a ; db += sum_b delta ;
// din[b][k] += sum_n delta[b][n] W[n][k]   (din must be zero-initialised).
// A CTA owns LIN_NS output neurons x a 4*KT-wide column chunk (blockIdx.y): thread = (row group, float4 column), the
// row of W and of dW is streamed once with 16-byte accesses while din partials stay in registers; one vector RED
// per (b, float4 column) at the end.  K % 4 == 0.
constexpr int LIN_NS = 16;
template <int KT>
__global__ void __launch_bounds__(256)
linear_small_bwd_kernel(const float* __restrict__ delta_raw, const float* __restrict__ out_act, float drop_scale,
                        const float* __restrict__ a, const float* __restrict__ W, float* __restrict__ dW, float* __restrict__ db,
                        float* __restrict__ din, int B, int N, int K) {
    constexpr int RG = 256 / KT;
    const int kc = threadIdx.x % KT, rg = threadIdx.x / KT;
    const int K4 = K >> 2;
    const int c4 = blockIdx.y * KT + kc;                  // float4 column
    const bool col_ok = c4 < K4;
    const int n0 = blockIdx.x * LIN_NS;
    __shared__ float s_d[LIN_NS][LIN_BCH];
    for (int b0 = 0; b0 < B; b0 += LIN_BCH) {
        __syncthreads();
        if (threadIdx.x < LIN_NS * LIN_BCH) {
            const int nn = threadIdx.x / LIN_BCH, b = threadIdx.x % LIN_BCH;
            float v = 0.f;
            if (n0 + nn < N && b0 + b < B) {
                v = delta_raw[(long long)(b0 + b) * N + n0 + nn];
                if (out_act) v = out_act[(long long)(b0 + b) * N + n0 + nn] > 0.f ? v * drop_scale : 0.f;
            }
            s_d[nn][b] = v;
        }
        __syncthreads();
        if (db != nullptr && blockIdx.y == 0 && threadIdx.x < LIN_NS && n0 + threadIdx.x < N) {
            float t = 0.f;
#pragma unroll
            for (int b = 0; b < LIN_BCH; ++b) t += s_d[threadIdx.x][b];
            db[n0 + threadIdx.x] += t;
        }
        if (!col_ok) continue;
        float4 av[LIN_BCH], acc[LIN_BCH];
#pragma unroll
        for (int b = 0; b < LIN_BCH; ++b) {
            av[b] = (b0 + b < B) ? reinterpret_cast<const float4*>(a + (long long)(b0 + b) * K)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
            acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 4
        for (int nn = rg; nn < LIN_NS; nn += RG) {
            const int n = n0 + nn;
            if (n >= N) break;
            const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (long long)n * K) + c4);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dW) g = reinterpret_cast<const float4*>(dW + (long long)n * K)[c4];
#pragma unroll
            for (int b = 0; b < LIN_BCH; ++b) {
                const float d = s_d[nn][b];
                g.x = fmaf(d, av[b].x, g.x); g.y = fmaf(d, av[b].y, g.y); g.z = fmaf(d, av[b].z, g.z); g.w = fmaf(d, av[b].w, g.w);
                acc[b].x = fmaf(d, wv.x, acc[b].x); acc[b].y = fmaf(d, wv.y, acc[b].y);
                acc[b].z = fmaf(d, wv.z, acc[b].z); acc[b].w = fmaf(d, wv.w, acc[b].w);
            }
            if (dW) reinterpret_cast<float4*>(dW + (long long)n * K)[c4] = g;
        }
        if (din != nullptr) {
#pragma unroll
            for (int b = 0; b < LIN_BCH; ++b) {
                if (b0 + b < B) {
                    float* dp = din + (long long)(b0 + b) * K + c4 * 4;
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(acc[b].x), "f"(acc[b].y),
                                 "f"(acc[b].z), "f"(acc[b].w) : "memory");
                }
            }
        }
    }
}

// p = sigmoid(z), loss = BCE mean (log clamped at -100 like nn.BCELoss), dz = dL/dz = (p - y)/B * gscale.
// smoothing > 0 implements LabelSmoothingBCEWithLogitsLoss (train_au_patch.py:203-211).
__global__ void bce_fwd_bwd_kernel(const float* __restrict__ z, const float* __restrict__ y, float smoothing, float* __restrict__ probs,
                                   float* __restrict__ loss, float* __restrict__ dz, int B) {
    __shared__ float s[32];
    float l = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float zz = z[b];
        const float p = sigmoidf_(zz);
        const float t = y[b] * (1.f - smoothing) + 0.5f * smoothing;
        float lp, l1p;
        if (smoothing > 0.f) {   // with-logits form (numerically exact)
            lp = -(fmaxf(-zz, 0.f) + log1pf(__expf(-fabsf(zz))));
            l1p = -(fmaxf(zz, 0.f) + log1pf(__expf(-fabsf(zz))));
        } else {
            lp = fmaxf(logf(p), -100.f);
            l1p = fmaxf(logf(1.f - p), -100.f);
        }
        l += -(t * lp + (1.f - t) * l1p);
        probs[b] = p;
        if (dz) dz[b] = (p - t) / (float)B;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s[i];
        *loss = tot / (float)B;
    }
}

// nn.BCELoss(reduction="mean") on probabilities (train_audio.py:20,39): loss = -mean(y*max(log p,-100) + (1-y)*max(log(1-p),-100))
// and dp = dL/dp = (p - y) / max(p(1-p), 1e-12) / n  (torch's binary_cross_entropy_backward), in one launch.
__global__ void bce_prob_fwd_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y, float* __restrict__ loss,
                                        float* __restrict__ dp, int n) {
    __shared__ float s[32];
    float l = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float pp = p[i], t = y[i];
        l -= t * fmaxf(logf(pp), -100.f) + (1.f - t) * fmaxf(logf(1.f - pp), -100.f);
        if (dp) dp[i] = (pp - t) / fmaxf((1.f - pp) * pp, 1e-12f) / (float)n;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s[i];
        *loss = tot / (float)n;
    }
}

// ============================================================================================ ArcFace (+CE / CB-focal)
// One warp per sample.  logits[b][c] = s * (c == y ? cos(theta + m) : cos), cos = <x/|x|, w_c/|w_c|> clamped.
// loss_mode 0: CrossEntropy mean ; 1: class-balanced focal (gamma, class weights).  labels < 0 => inference logits only.
__global__ void __launch_bounds__(256)
arcface_loss_kernel(const float* __restrict__ x, const float* __restrict__ w, const long long* __restrict__ labels, float s_,
                    float m_, int loss_mode, const float* __restrict__ class_w, float gamma,
                    const float* __restrict__ dlogits_in, float* __restrict__ logits, float* __restrict__ loss_rows, float* __restrict__ dx, float* __restrict__ dw_partial, int B, int D, float gscale) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* xb = x + (long long)b * D;
    float xx = 0.f, w0w0 = 0.f, w1w1 = 0.f, xw0 = 0.f, xw1 = 0.f;
    for (int k = lane; k < D; k += 32) {
        const float xv = xb[k], a = w[k], c = w[D + k];
        xx = fmaf(xv, xv, xx); w0w0 = fmaf(a, a, w0w0); w1w1 = fmaf(c, c, w1w1);
        xw0 = fmaf(xv, a, xw0); xw1 = fmaf(xv, c, xw1);
    }
    xx = warp_sum(xx); w0w0 = warp_sum(w0w0); w1w1 = warp_sum(w1w1); xw0 = warp_sum(xw0); xw1 = warp_sum(xw1);
    const float nx = fmaxf(sqrtf(xx), 1e-12f), nw0 = fmaxf(sqrtf(w0w0), 1e-12f), nw1 = fmaxf(sqrtf(w1w1), 1e-12f);
    const float cosv[2] = {xw0 / (nx * nw0), xw1 / (nx * nw1)};
    const long long y = labels ? labels[b] : -1;
    float lg[2] = {s_ * cosv[0], s_ * cosv[1]};
    float dt_dcos = 1.f;   // d target_logit / d cos_y (before the s factor)
    if (y >= 0) {
        const float cy = cosv[y];
        const float cl = fminf(fmaxf(cy, -1.f + 1e-7f), 1.f - 1e-7f);
        const float th = acosf(cl);
        lg[y] = s_ * cosf(th + m_);
        const bool inside = (cy >= -1.f + 1e-7f) && (cy <= 1.f - 1e-7f);
        dt_dcos = inside ? sinf(th + m_) / sqrtf(fmaxf(1.f - cl * cl, 1e-30f)) : 0.f;
    }
    if (lane == 0) { logits[b * 2] = lg[0]; logits[b * 2 + 1] = lg[1]; }
    if (y < 0) return;
    // softmax / CE
    const float mx = fmaxf(lg[0], lg[1]);
    const float e0 = __expf(lg[0] - mx), e1 = __expf(lg[1] - mx);
    const float den = e0 + e1;
    const float p[2] = {e0 / den, e1 / den};
    const float ce_plain = -(lg[y] - mx - logf(den));
    float dl[2];   // dL/dlogits for this row (already including the 1/B mean and gscale)
    float row_loss = 0.f;
    if (dlogits_in != nullptr) {   // external criterion: upstream gradient supplied by the caller
        dl[0] = dlogits_in[b * 2]; dl[1] = dlogits_in[b * 2 + 1];
    } else if (loss_mode == 0) {
        row_loss = ce_plain;
        dl[0] = (p[0] - (y == 0 ? 1.f : 0.f));
        dl[1] = (p[1] - (y == 1 ? 1.f : 0.f));
    } else {
        // ce = w_y * ce_plain ; pt = exp(-ce) ; loss = (1-pt)^gamma * ce   (train_au_face.py:455-458)
        const float wy = class_w[y];
        const float ce = wy * ce_plain;
        const float pt = __expf(-ce);
        const float om = 1.f - pt;
        row_loss = powf(om, gamma) * ce;
        // d/dce [ (1-pt)^g * ce ] = g (1-pt)^(g-1) * pt * ce + (1-pt)^g
        const float dloss_dce = gamma * powf(fmaxf(om, 1e-30f), gamma - 1.f) * pt * ce + powf(om, gamma);
        dl[0] = dloss_dce * wy * (p[0] - (y == 0 ? 1.f : 0.f));
        dl[1] = dloss_dce * wy * (p[1] - (y == 1 ? 1.f : 0.f));
    }
    if (dlogits_in == nullptr) {
        const float inv = gscale / (float)B;
        dl[0] *= inv; dl[1] *= inv;
    }
    if (lane == 0 && loss_rows != nullptr) loss_rows[b] = row_loss / (float)B;
    // d cos_c : s * dl_c (* dt_dcos for the label class)
    const float dc0 = s_ * dl[0] * (y == 0 ? dt_dcos : 1.f);
    const float dc1 = s_ * dl[1] * (y == 1 ? dt_dcos : 1.f);
    // d cos_c / dx = (w_c/|w_c| - cos_c x/|x|)/|x| ; d cos_c / dw_c = (x/|x| - cos_c w_c/|w_c|)/|w_c|
    for (int k = lane; k < D; k += 32) {
        const float xv = xb[k] / nx, a = w[k] / nw0, c = w[D + k] / nw1;
        if (dx) dx[(long long)b * D + k] = (dc0 * (a - cosv[0] * xv) + dc1 * (c - cosv[1] * xv)) / nx;
        if (dw_partial) {
            atomicAdd(&dw_partial[k], dc0 * (xv - cosv[0] * a) / nw0);
            atomicAdd(&dw_partial[D + k], dc1 * (xv - cosv[1] * c) / nw1);
        }
    }
}

// p = sigmoid(z) ; dz = dp * p * (1 - p)      (nn.Sigmoid, XceptionLSTMV.py:44,70)
__global__ void sigmoid_kernel(const float* __restrict__ z, float* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.f / (1.f + expf(-z[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ dz, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dz[i] = dp[i] * p[i] * (1.f - p[i]);
}

__global__ void sum_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out, int accumulate) {
    __shared__ float s[32];
    float l = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) l += v[i];
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s[i];
        *out = accumulate ? *out + tot : tot;
    }
}

// ============================================================================================ fusion regularisers
// tokens v, a : [B,T,D].  pooled[b] = [mean_t v | mean_t a] ; loss_reg = la * mse(vpool, apool) + lt * 0.5 *
// (mean (v[t+1]-v[t])^2 + mean (a[t+1]-a[t])^2).  Also writes d loss_reg / d tokens into dv, da (=, not +=).
__global__ void fusion_pool_reg_kernel(const float* __restrict__ v, const float* __restrict__ a, float* __restrict__ pooled,
                                       float* __restrict__ loss_reg, float* __restrict__ dv, float* __restrict__ da, int B, int T,
                                       int D, float la, float lt, float gscale) {
    // one block per sample; threads over D
    const int b = blockIdx.x;
    __shared__ float s[32];
    float local = 0.f;
    const float inv_bd = 1.f / ((float)B * D);
    const float inv_tm = T > 1 ? 1.f / ((float)B * (T - 1) * D) : 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float sv = 0.f, sa = 0.f;
        for (int t = 0; t < T; ++t) {
            sv += v[((long long)b * T + t) * D + d];
            sa += a[((long long)b * T + t) * D + d];
        }
        const float vp = sv / T, ap = sa / T;
        pooled[(long long)b * 2 * D + d] = vp;
        pooled[(long long)b * 2 * D + D + d] = ap;
        const float diff = vp - ap;
        local += la * diff * diff * inv_bd;
        const float gal = la * 2.f * diff * inv_bd / T * gscale;   // d/dv[t] of the align term
        for (int t = 0; t < T; ++t) {
            float gv = gal, ga = -gal;
            if (T > 1) {
                const long long o = ((long long)b * T + t) * D + d;
                float tv = 0.f, ta = 0.f;
                if (t + 1 < T) { const float e = v[o + D] - v[o]; const float f = a[o + D] - a[o]; local += lt * 0.5f * (e * e + f * f) * inv_tm; tv -= e; ta -= f; }
                if (t > 0) { tv += v[o] - v[o - D]; ta += a[o] - a[o - D]; }
                gv += lt * 0.5f * 2.f * tv * inv_tm * gscale;
                ga += lt * 0.5f * 2.f * ta * inv_tm * gscale;
            }
            if (dv) dv[((long long)b * T + t) * D + d] = gv;
            if (da) da[((long long)b * T + t) * D + d] = ga;
        }
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s[i];
        atomicAdd(loss_reg, tot);
    }
}

// dtokens[b,t,d] += dpooled[b, off + d] / T
__global__ void fusion_pool_bwd_kernel(const float* __restrict__ dpooled, float* __restrict__ dv, float* __restrict__ da, int B, int T,
                                       int D) {
    const long long n = (long long)B * T * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int b = (int)(i / ((long long)T * D));
        dv[i] += dpooled[(long long)b * 2 * D + d] / T;
        da[i] += dpooled[(long long)b * 2 * D + D + d] / T;
    }
}

}  // namespace xcp

using namespace xcp;
#define ST ((cudaStream_t)stream)

extern "C" int xcp_cast_f32_bf16(const float* x, void* out, long long n, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    long long g = (n + 255) / 256;
    if (g > 4096) g = 4096;
    if (g < 1) g = 1;
    cast_f32_bf16_kernel<<<(int)g, 256, 0, ST>>>(x, (__nv_bfloat16*)out, n);
    return check_cuda(cudaGetLastError(), "cast launch");
}

namespace xcp {
// lstm_cluster.cu: cluster kernels for H = 256 / 512; -1000 = not applicable here (fall back to the single-CTA kernels)
int lstm_fwd_cluster(const float* xproj, const float* b_ih, const float* b_hh, const void* w_hh_t, float* h_out, float* gates,
                     float* cstate, float* hn, float* cn, int B, int T, int H, cudaStream_t st);
int lstm_bwd_cluster(const float* dout, const float* dhn, const float* dcn, const float* gates, const float* cstate,
                     const float* hstate, const void* w_hh, void* dgates, void* hprev, float* dbias_ih, float* dbias_hh, int B,
                     int T, int H, cudaStream_t st);
}  // namespace xcp

extern "C" int xcp_lstm_fwd(const float* xproj, const float* b_ih, const float* b_hh, const void* w_hh_t, float* h_out,
                            float* gates, float* cstate, float* hn, float* cn, int B, int T, int H, int device, void* stream) {
    XCP_REQUIRE(B > 0 && T > 0 && H > 0 && H % 8 == 0 && H <= 4096, "xcp_lstm_fwd: bad shape B=%d T=%d H=%d", B, T, H);
    XCP_CUDA(cudaSetDevice(device));
    {
        const int rc = xcp::lstm_fwd_cluster(xproj, b_ih, b_hh, w_hh_t, h_out, gates, cstate, hn, cn, B, T, H, ST);
        if (rc != -1000) return rc;
    }
    const int nt = 4 * H < 1024 ? 4 * H : 1024;
    XCP_REQUIRE(H <= 4 * nt, "xcp_lstm_fwd: H too large for the register-resident cell state");
    size_t smem = (size_t)5 * H * sizeof(float);
    const size_t wbytes = (size_t)H * 4 * H * 2;
    const int w_in_smem = (smem + wbytes <= 200 * 1024) ? 1 : 0;
    if (w_in_smem) smem += wbytes;
    XCP_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_fwd_kernel<<<B, nt, smem, ST>>>(xproj, b_ih, b_hh, (const __nv_bfloat16*)w_hh_t, h_out, gates, cstate, hn, cn, T, H,
                                         w_in_smem);
    return check_cuda(cudaGetLastError(), "lstm_fwd launch");
}

extern "C" int xcp_lstm_bwd(const float* dout, const float* dhn, const float* dcn, const float* gates, const float* cstate,
                            const float* hstate, const void* w_hh, void* dgates, void* hprev, float* dbias_ih, float* dbias_hh,
                            int B, int T, int H, int device, void* stream) {
    XCP_REQUIRE(B > 0 && T > 0 && H > 0 && H % 8 == 0, "xcp_lstm_bwd: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    {
        const int rc = xcp::lstm_bwd_cluster(dout, dhn, dcn, gates, cstate, hstate, w_hh, dgates, hprev, dbias_ih, dbias_hh, B, T, H, ST);
        if (rc != -1000) return rc;
    }
    const int nt = 4 * H < 1024 ? 4 * H : 1024;
    const size_t smem = (size_t)9 * H * sizeof(float);
    XCP_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_bwd_kernel<<<B, nt, smem, ST>>>(dout, dhn, dcn, gates, cstate, hstate, (const __nv_bfloat16*)w_hh,
                                         (__nv_bfloat16*)dgates, (__nv_bfloat16*)hprev, dbias_ih, dbias_hh, T, H);
    return check_cuda(cudaGetLastError(), "lstm_bwd launch");
}

extern "C" int xcp_linear_small_fwd(const float* a, const float* W, const float* bias, const void* mask, float drop_scale, int act,
                                    float* out, int B, int N, int K, int device, void* stream) {
    XCP_REQUIRE(B > 0 && B <= MAXB, "xcp_linear_small_fwd: batch %d exceeds %d rows", B, MAXB);
    XCP_CUDA(cudaSetDevice(device));
    linear_small_fwd_kernel<<<(N + 7) / 8, 256, 0, ST>>>(a, W, bias, (const uint8_t*)mask, drop_scale, act, out, B, N, K);
    return check_cuda(cudaGetLastError(), "linear_small_fwd launch");
}

extern "C" int xcp_linear_small_bwd(const float* delta_raw, const float* out_act, float drop_scale, const float* a, const float* W,
                                    float* dW, float* db, float* din, int B, int N, int K, int device, void* stream) {
    XCP_REQUIRE(B > 0 && B <= MAXB, "xcp_linear_small_bwd: batch %d exceeds %d rows", B, MAXB);
    XCP_REQUIRE(K % 4 == 0, "xcp_linear_small_bwd: K must be a multiple of 4 (got %d)", K);
    XCP_CUDA(cudaSetDevice(device));
    const int K4 = K / 4;
    const int nb = (N + LIN_NS - 1) / LIN_NS;
    if (K4 <= 32) linear_small_bwd_kernel<32><<<dim3(nb, 1), 256, 0, ST>>>(delta_raw, out_act, drop_scale, a, W, dW, db, din, B, N, K);
    else if (K4 <= 64) linear_small_bwd_kernel<64><<<dim3(nb, 1), 256, 0, ST>>>(delta_raw, out_act, drop_scale, a, W, dW, db, din, B, N, K);
    else if (K4 <= 128) linear_small_bwd_kernel<128><<<dim3(nb, 1), 256, 0, ST>>>(delta_raw, out_act, drop_scale, a, W, dW, db, din, B, N, K);
    else linear_small_bwd_kernel<256><<<dim3(nb, (K4 + 255) / 256), 256, 0, ST>>>(delta_raw, out_act, drop_scale, a, W, dW, db, din, B, N, K);
    return check_cuda(cudaGetLastError(), "linear_small_bwd launch");
}

extern "C" int xcp_sigmoid_fwd(const float* z, float* p, int n, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    sigmoid_kernel<<<(n + 255) / 256, 256, 0, ST>>>(z, p, n);
    return check_cuda(cudaGetLastError(), "sigmoid launch");
}

extern "C" int xcp_sigmoid_bwd(const float* p, const float* dp, float* dz, int n, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    sigmoid_bwd_kernel<<<(n + 255) / 256, 256, 0, ST>>>(p, dp, dz, n);
    return check_cuda(cudaGetLastError(), "sigmoid_bwd launch");
}

extern "C" int xcp_bce_fwd_bwd(const float* z, const float* y, float smoothing, float* probs, float* loss, float* dz, int B,
                               int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    bce_fwd_bwd_kernel<<<1, 256, 0, ST>>>(z, y, smoothing, probs, loss, dz, B);
    return check_cuda(cudaGetLastError(), "bce launch");
}

extern "C" int xcp_bce_prob_fwd_bwd(const float* p, const float* y, float* loss, float* dp, int n, int device, void* stream) {
    XCP_REQUIRE(n > 0, "xcp_bce_prob_fwd_bwd: empty batch");
    XCP_CUDA(cudaSetDevice(device));
    bce_prob_fwd_bwd_kernel<<<1, 256, 0, ST>>>(p, y, loss, dp, n);
    return check_cuda(cudaGetLastError(), "bce_prob launch");
}

// ArcFace logits (+ CE or CB-focal loss and gradients when labels != null).  loss (scalar) is overwritten,
// dw is accumulated (+=), dx is overwritten.  loss_rows: workspace [B].
extern "C" int xcp_arcface_loss(const float* x, const float* w, const long long* labels, float s, float m, int loss_mode,
                                const float* class_w, float gamma, const float* dlogits_in, float* logits, float* loss,
                                float* loss_rows, float* dx, float* dw, int B, int D, float gscale, int device, void* stream) {
    XCP_REQUIRE(B > 0 && D > 0, "xcp_arcface_loss: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    arcface_loss_kernel<<<(B + 7) / 8, 256, 0, ST>>>(x, w, labels, s, m, loss_mode, class_w, gamma, dlogits_in, logits, loss_rows,
                                                     dx, dw, B, D, gscale);
    XCP_CUDA(cudaGetLastError());
    if (labels != nullptr && loss != nullptr && dlogits_in == nullptr) sum_rows_kernel<<<1, 256, 0, ST>>>(loss_rows, B, loss, 0);
    return check_cuda(cudaGetLastError(), "arcface launch");
}

extern "C" int xcp_fusion_pool_reg(const float* v, const float* a, float* pooled, float* loss_reg, float* dv, float* da, int B,
                                   int T, int D, float lambda_align, float lambda_temp, float gscale, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    XCP_CUDA(cudaMemsetAsync(loss_reg, 0, sizeof(float), ST));
    fusion_pool_reg_kernel<<<B, 256, 0, ST>>>(v, a, pooled, loss_reg, dv, da, B, T, D, lambda_align, lambda_temp, gscale);
    return check_cuda(cudaGetLastError(), "fusion_pool_reg launch");
}

extern "C" int xcp_fusion_pool_bwd(const float* dpooled, float* dv, float* da, int B, int T, int D, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    long long n = (long long)B * T * D;
    long long g = (n + 255) / 256;
    if (g > 2048) g = 2048;
    fusion_pool_bwd_kernel<<<(int)g, 256, 0, ST>>>(dpooled, dv, da, B, T, D);
    return check_cuda(cudaGetLastError(), "fusion_pool_bwd launch");
}
