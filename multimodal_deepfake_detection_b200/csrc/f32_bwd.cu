// fp32 VALIDATION path, second half: the fused-signature forward twins and the whole BACKWARD of the Xception plan in
// plain fp32 arithmetic (double accumulation for every long reduction).
//
// Why it exists ("What's missing" #1): the production backward computes in bf16 on tcgen05, where the best
// achievable per-tensor gradient agreement with the fp32 reference is ~1e-2 -- too coarse to tell a wrong term in the
// BatchNorm / max-pool / residual chain rule from rounding.  These kernels take the SAME arguments, with the same
// fused semantics (producer-BN affine + ReLU prologues, pool/skip/add routing, BN-backward sums coming out of the
// depthwise backward, channel padding), as the production entry points they shadow, on fp32 NHWC activations; the plan
// executor (executor.py) drives both families through one code path (ops.py dispatches on the activation dtype), so the
// sequencing logic that trains is the logic that is checked against the fp32 oracle at <= 1e-4 per tensor
// (tests/test_fp32_plan_gpu.py).  The pointwise GEMMs of this plan run either on the FFMA kernels of f32.cu or, with
// xcp_split3_bf16, on the production tcgen05 kernels fed 3-way bf16 splits of the fp32 operands.
//
// Reference arithmetic: Xception.py:44-47, 89-99, 167-201 and the autograd formulas of the torch ops they call.
#include "common.cuh"

namespace {
using namespace xcp;
#define ST ((cudaStream_t)stream)

inline unsigned grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = 148LL * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float act_of(float v, const float* scale, const float* shift, int c, int relu) {
    if (scale != nullptr) v = fmaf(v, scale[c], shift[c]);
    return relu ? fmaxf(v, 0.f) : v;
}

// ---------------------------------------------------------------------------------------------- forward twins
// out = depthwise3x3(act(x)), act = relu?(scale*x + shift); zero padding applies to act(x).  w9 = [9][C] tap-major.
__global__ void __launch_bounds__(256)
f32_dw_fused_kernel(const float* __restrict__ x, const float* __restrict__ w9, const float* __restrict__ scale,
                    const float* __restrict__ shift, int relu, float* __restrict__ out, int F, int H, int W, int C) {
    const long long total = (long long)F * H * W * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const long long f = p / H;
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hi = h + kh - 1, wi = w + kw - 1;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W)
                    acc = fmaf(act_of(x[((f * H + hi) * W + wi) * C + c], scale, shift, c, relu), w9[(kh * 3 + kw) * C + c], acc);
            }
        out[i] = acc;
    }
}

// out = maxpool3x3s2p1(scale*y + shift) + (scale_s*ys + shift_s); idx = first maximal tap in row-major scan order
__global__ void __launch_bounds__(256)
f32_pool_add_fused_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ ys, const float* __restrict__ scale_s, const float* __restrict__ shift_s,
                          float* __restrict__ out, unsigned char* __restrict__ idx, float* __restrict__ ymax, int F, int H, int W, int C,
                          int Ho, int Wo) {
    const long long total = (long long)F * Ho * Wo * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const long long f = p / Ho;
        float m = -INFINITY, raw = 0.f;
        int best = 4;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int hi = 2 * ho + kh - 1, wi = 2 * wo + kw - 1;
                if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                const float yv = y[((f * H + hi) * W + wi) * C + c];
                const float v = fmaf(yv, scale[c], shift[c]);
                if (v > m) { m = v; best = kh * 3 + kw; raw = yv; }
            }
        float s = ys[i];
        if (scale_s != nullptr) s = fmaf(s, scale_s[c], shift_s[c]);
        out[i] = m + s;
        if (idx != nullptr) idx[i] = (unsigned char)best;
        if (ymax != nullptr) ymax[i] = raw;
    }
}

// out = scale*y + shift + (scale_s ? scale_s*skip + shift_s : skip)
__global__ void __launch_bounds__(256)
f32_bn_add_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ skip, const float* __restrict__ scale_s, const float* __restrict__ shift_s,
                  float* __restrict__ out, long long n, int C) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        float s = skip[i];
        if (scale_s != nullptr) s = fmaf(s, scale_s[c], shift_s[c]);
        out[i] = fmaf(y[i], scale[c], shift[c]) + s;
    }
}

// feat[f,c] = mean over HW of relu(scale*y + shift)
__global__ void __launch_bounds__(256)
f32_bn_relu_gap_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                       float* __restrict__ feat, int F, int HW, int C) {
    const long long total = (long long)F * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        const long long f = i / C;
        double s = 0.0;
        for (int p = 0; p < HW; ++p) s += (double)fmaxf(fmaf(y[(f * HW + p) * C + c], scale[c], shift[c]), 0.f);
        feat[i] = (float)(s / (double)HW);
    }
}

// ---------------------------------------------------------------------------------------------- BN backward
enum { SRC_DIRECT = 0, SRC_RELU = 1, SRC_POOL = 2, SRC_GAP_RELU = 3 };
struct BnSrc {
    int mode;
    const float* G;             // modes 0,1: [F,H,W,C]; mode 2: [F,Ho,Wo,C]
    const unsigned char* idx;   // mode 2
    const float* dfeat;         // mode 3: [F,C]
    const float* scale;         // modes 1,3
    const float* shift;
    int F, H, W, C;
};

// gradient wrt the BN output z = scale*y + shift at flat element i = ((f*H+h)*W+w)*C + c
__device__ __forceinline__ float bn_dz(const BnSrc& s, long long i, float yv) {
    const int c = (int)(i % s.C);
    float dz;
    if (s.mode == SRC_DIRECT || s.mode == SRC_RELU) {
        dz = s.G[i];
    } else if (s.mode == SRC_POOL) {
        long long p = i / s.C;
        const int w = (int)(p % s.W); p /= s.W;
        const int h = (int)(p % s.H);
        const long long f = p / s.H;
        const int Ho = (s.H - 1) / 2 + 1, Wo = (s.W - 1) / 2 + 1;
        dz = 0.f;
        // windows (oh, ow) whose 3x3 footprint 2o-1 .. 2o+1 contains (h, w)
        for (int oh = (h + 1) / 2 - ((h & 1) ? 1 : 0); oh <= (h + 1) / 2; ++oh) {
            if (oh < 0 || oh >= Ho) continue;
            const int kh = h - (2 * oh - 1);
            if (kh < 0 || kh > 2) continue;
            for (int ow = (w + 1) / 2 - ((w & 1) ? 1 : 0); ow <= (w + 1) / 2; ++ow) {
                if (ow < 0 || ow >= Wo) continue;
                const int kw = w - (2 * ow - 1);
                if (kw < 0 || kw > 2) continue;
                const long long o = ((f * Ho + oh) * Wo + ow) * s.C + c;
                if ((int)s.idx[o] == kh * 3 + kw) dz += s.G[o];
            }
        }
    } else {
        const long long f = i / ((long long)s.H * s.W * s.C);
        dz = s.dfeat[f * s.C + c] / (float)(s.H * s.W);
    }
    if (s.mode == SRC_RELU || s.mode == SRC_GAP_RELU)
        if (!(fmaf(yv, s.scale[c], s.shift[c]) > 0.f)) dz = 0.f;
    return dz;
}

// sums[0][c] = sum dz, sums[1][c] = sum dz*y; block = 32 channels x 8 pixel lanes
__global__ void __launch_bounds__(256)
f32_bn_bwd_reduce_kernel(const float* __restrict__ y, const BnSrc s, float* __restrict__ sums) {
    __shared__ double sh1[8][32], sh2[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const long long M = (long long)s.F * s.H * s.W;
    double a1 = 0.0, a2 = 0.0;
    if (c < s.C)
        for (long long p = ry; p < M; p += 8) {
            const long long i = p * s.C + c;
            const float yv = y[i];
            const float dz = bn_dz(s, i, yv);
            a1 += (double)dz; a2 += (double)dz * (double)yv;
        }
    sh1[ry][cx] = a1; sh2[ry][cx] = a2;
    __syncthreads();
    if (ry == 0 && c < s.C) {
        for (int r = 1; r < 8; ++r) { a1 += sh1[r][cx]; a2 += sh2[r][cx]; }
        sums[c] = (float)a1; sums[s.C + c] = (float)a2;
    }
}

// coefficients of dy = A*dz + B*y + Cc and the parameter gradients (accumulated), pad channels -> 0
__global__ void f32_bn_bwd_finalize_kernel(const float* __restrict__ sums, int C, int C_real, double count, const float* __restrict__ gamma,
                                           const float* __restrict__ mean, const float* __restrict__ rstd, int training,
                                           float* __restrict__ coef, float* dgamma, float* dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (c >= C_real) { coef[c] = 0.f; coef[C + c] = 0.f; coef[2 * C + c] = 0.f; return; }
    const double s1 = sums[c], s2 = sums[C + c];
    const double m = mean[c], r = rstd[c], g = gamma[c];
    const double dg = r * (s2 - m * s1);                 // sum dz * xhat
    const double A = g * r;
    double B = 0.0, Cc = 0.0;
    if (training) {
        B = -g * r * r * dg / count;
        Cc = -B * m - A * s1 / count;
    }
    coef[c] = (float)A; coef[C + c] = (float)B; coef[2 * C + c] = (float)Cc;
    if (dgamma != nullptr) dgamma[c] += (float)dg;
    if (dbeta != nullptr) dbeta[c] += (float)s1;
}

__global__ void __launch_bounds__(256)
f32_bn_bwd_apply_kernel(const float* __restrict__ y, const BnSrc s, const float* __restrict__ coef, float* __restrict__ dy,
                        int grid_w, int grid_h) {
    const long long total = (long long)s.F * s.H * s.W * s.C;
    const int C = s.C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        const float yv = y[i];
        const float dz = bn_dz(s, i, yv);
        long long o = i;
        if (grid_w > 0) {
            long long p = i / C;
            const int w = (int)(p % s.W); p /= s.W;
            const int h = (int)(p % s.H);
            const long long f = p / s.H;
            o = ((f * grid_h + h) * grid_w + w) * C + c;
        }
        dy[o] = fmaf(coef[c], dz, fmaf(coef[C + c], yv, coef[2 * C + c]));
    }
}

// ---------------------------------------------------------------------------------------------- depthwise backward
// dz = mask * conv_transpose(dD) [+ add_full] [+ add_half at even pixels]
__global__ void __launch_bounds__(256)
f32_dw_bwd_dz_kernel(const float* __restrict__ dD, const float* __restrict__ xin, const float* __restrict__ w9,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu, float* __restrict__ dz,
                     const float* __restrict__ add_full, const float* __restrict__ add_half, int F, int H, int W, int C) {
    const long long total = (long long)F * H * W * C;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const long long f = p / H;
        float g = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ho = h - kh + 1, wo = w - kw + 1;       // output pixel that read (h, w) through tap (kh, kw)
                if (ho >= 0 && ho < H && wo >= 0 && wo < W) g = fmaf(dD[((f * H + ho) * W + wo) * C + c], w9[(kh * 3 + kw) * C + c], g);
            }
        if (relu) {
            float z = xin[i];
            if (scale != nullptr) z = fmaf(z, scale[c], shift[c]);
            if (!(z > 0.f)) g = 0.f;
        }
        if (add_full != nullptr) g += add_full[i];
        if (add_half != nullptr && !(h & 1) && !(w & 1)) g += add_half[((f * Ho + (h >> 1)) * Wo + (w >> 1)) * C + c];
        dz[i] = g;
    }
}

// k < 9: dw[c][k] += sum_p dD[p] * act(x)[p + tap k];  k = 9: bnsum[0][c] += sum dz;  k = 10: bnsum[1][c] += sum dz*x
// grid (ceil(C/32), 11 or 9), block = 32 channels x 8 pixel lanes
__global__ void __launch_bounds__(256)
f32_dw_bwd_red_kernel(const float* __restrict__ dD, const float* __restrict__ xin, const float* __restrict__ scale,
                      const float* __restrict__ shift, int relu, const float* __restrict__ dz, float* dw, float* bnsum, int F, int H,
                      int W, int C, int c_real) {
    __shared__ double sh[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int k = blockIdx.y;
    const long long M = (long long)F * H * W;
    double a = 0.0;
    if (c < C) {
        if (k < 9) {
            const int kh = k / 3, kw = k % 3;
            for (long long p = ry; p < M; p += 8) {
                long long t = p;
                const int w = (int)(t % W); t /= W;
                const int h = (int)(t % H);
                const long long f = t / H;
                const int hi = h + kh - 1, wi = w + kw - 1;
                if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                a += (double)dD[p * C + c] * (double)act_of(xin[((f * H + hi) * W + wi) * C + c], scale, shift, c, relu);
            }
        } else {
            for (long long p = ry; p < M; p += 8) {
                const float d = dz[p * C + c];
                a += (k == 9) ? (double)d : (double)d * (double)xin[p * C + c];
            }
        }
    }
    sh[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int r = 1; r < 8; ++r) a += sh[r][cx];
        if (k < 9) { if (c < c_real) dw[(long long)c * 9 + k] += (float)a; }
        else bnsum[(long long)(k - 9) * C + c] += (float)a;
    }
}

// ---------------------------------------------------------------------------------------------- GEMM weight gradient
// dW[P,Q] += dY[R,P]^T X[R,Q]: 64x64 tile, 16-row slices, fp32 inside a 512-row chunk, double across chunks
__global__ void __launch_bounds__(256)
f32_gemm_wgrad_kernel(const float* __restrict__ dY, long long ld_dy, const float* __restrict__ X, long long ld_x, float* dW,
                      long long ld_dw, long long R, int P, int Q) {
    __shared__ float sa[16][64 + 1], sb[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int p0 = blockIdx.y * 64, q0 = blockIdx.x * 64;
    double acc[4][4] = {};
    float part[4][4] = {};
    for (long long r0 = 0; r0 < R; r0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int k = e >> 6, j = e & 63;
            const long long r = r0 + k;
            sa[k][j] = (r < R && p0 + j < P) ? dY[r * ld_dy + p0 + j] : 0.f;
            sb[k][j] = (r < R && q0 + j < Q) ? X[r * ld_x + q0 + j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; bv[i] = sb[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
        }
        __syncthreads();
        if (((r0 >> 4) & 31) == 31 || r0 + 16 >= R) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc[i][j] += (double)part[i][j]; part[i][j] = 0.f; }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = p0 + ty * 4 + i, q = q0 + tx * 4 + j;
            if (p < P && q < Q) dW[(long long)p * ld_dw + q] += (float)acc[i][j];
        }
}

// ---------------------------------------------------------------------------------------------- dense 3x3 (stem) backward
// dx[f,h,w,ci] = sum_{kh,kw,co} dy[f,h-kh,w-kw,co] * w[co,ci,kh,kw]   (stride 1, padding 0; dy [F,Ho,Wo,Co], x/dx NHWC)
__global__ void __launch_bounds__(256)
f32_conv3x3_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int F, int H, int W, int Ci,
                         int Co, int Ho, int Wo) {
    const long long total = (long long)F * H * W * Ci;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ci = (int)(i % Ci);
        long long p = i / Ci;
        const int x = (int)(p % W); p /= W;
        const int y = (int)(p % H);
        const long long f = p / H;
        float acc = 0.f;
        for (int kh = 0; kh < 3; ++kh) {
            const int ho = y - kh;
            if (ho < 0 || ho >= Ho) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int wo = x - kw;
                if (wo < 0 || wo >= Wo) continue;
                const float* dp = dy + ((f * Ho + ho) * Wo + wo) * Co;
                for (int co = 0; co < Co; ++co) acc = fmaf(dp[co], w[((co * Ci + ci) * 3 + kh) * 3 + kw], acc);
            }
        }
        dx[i] = acc;
    }
}

// dw[co,ci,kh,kw] += sum_{f,ho,wo} dy[f,ho,wo,co] * x[f, ho*s+kh, wo*s+kw, ci]; block = (ci,kh,kw), threads = Co x lanes
__global__ void __launch_bounds__(256)
f32_conv3x3_wgrad_kernel(const float* __restrict__ x, int x_nchw, const float* __restrict__ dy, float* dw, int F, int H, int W, int Ci,
                         int Co, int stride, int Ho, int Wo) {
    __shared__ double sh[256];
    const int ci = blockIdx.x / 9, k = blockIdx.x % 9, kh = k / 3, kw = k % 3;
    const int lanes = 256 / Co;                      // Co in {32, 64}: 8 or 4 pixel lanes
    const int co = threadIdx.x % Co, ln = threadIdx.x / Co;
    const long long M = (long long)F * Ho * Wo;
    double a = 0.0;
    if (ln < lanes)
        for (long long p = ln; p < M; p += lanes) {
            long long t = p;
            const int wo = (int)(t % Wo); t /= Wo;
            const int ho = (int)(t % Ho);
            const long long f = t / Ho;
            const int hi = ho * stride + kh, wi = wo * stride + kw;
            const float xv = x_nchw ? x[((f * Ci + ci) * H + hi) * W + wi] : x[((f * H + hi) * W + wi) * Ci + ci];
            a += (double)dy[p * Co + co] * (double)xv;
        }
    sh[threadIdx.x] = a;
    __syncthreads();
    if (ln == 0) {
        for (int r = 1; r < lanes; ++r) a += sh[r * Co + co];
        dw[((co * Ci + ci) * 3 + kh) * 3 + kw] += (float)a;
    }
}

// ---------------------------------------------------------------------------------------------- 3-way bf16 split
// x = h + m + l with h = bf16(x), m = bf16(x - h), l = bf16(x - h - m) (24 mantissa bits).  The six products
// h.h + h.m + m.h + h.l + m.m + l.h of two split operands reproduce the fp32 product to ~2^-24, so a bf16 tensor-core GEMM
// over the 6-fold concatenation computes an fp32-grade result:  side 0 ("A") emits (h,h,m,h,m,l), side 1 ("B") emits
// (h,m,h,l,m,h).  along_rows = 0: out[rows][6*cols] (K concatenation, for D = A.B^T); 1: out[6*rows][cols] (for the
// weight gradient dW = dY^T.X, whose reduction runs over rows).
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int cols, int side, int along_rows) {
    const long long total = rows * cols;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16(v);
        const float r1 = v - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16(r1);
        const __nv_bfloat16 l = __float2bfloat16(r1 - __bfloat162float(m));
        const __nv_bfloat16 pa[6] = {h, h, m, h, m, l}, pb[6] = {h, m, h, l, m, h};
        const long long r = i / cols;
        const int c = (int)(i % cols);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const __nv_bfloat16 o = side == 0 ? pa[j] : pb[j];
            if (along_rows) out[((long long)j * rows + r) * cols + c] = o;
            else out[r * 6LL * cols + (long long)j * cols + c] = o;
        }
    }
}
}  // namespace

// ================================================================================================ C ABI
extern "C" int xcp_f32_dw3x3_fused(const float* x, const float* w9, const float* scale, const float* shift, int relu, float* out,
                                   int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && (scale == nullptr) == (shift == nullptr), "xcp_f32_dw3x3_fused: bad arguments");
    XCP_CUDA(cudaSetDevice(device));
    f32_dw_fused_kernel<<<grid_for((long long)F * H * W * C, 256), 256, 0, ST>>>(x, w9, scale, shift, relu, out, F, H, W, C);
    return check_cuda(cudaGetLastError(), "f32_dw_fused launch");
}

extern "C" int xcp_f32_pool_add_fused(const float* y, const float* scale, const float* shift, const float* ys, const float* scale_s,
                                      const float* shift_s, float* out, void* idx, float* ymax, int F, int H, int W, int C, int device,
                                      void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && scale != nullptr && shift != nullptr && ys != nullptr, "xcp_f32_pool_add_fused: bad arguments");
    XCP_CUDA(cudaSetDevice(device));
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    f32_pool_add_fused_kernel<<<grid_for((long long)F * Ho * Wo * C, 256), 256, 0, ST>>>(y, scale, shift, ys, scale_s, shift_s, out,
                                                                                      (unsigned char*)idx, ymax, F, H, W, C, Ho, Wo);
    return check_cuda(cudaGetLastError(), "f32_pool_add_fused launch");
}

extern "C" int xcp_f32_bn_add(const float* y, const float* scale, const float* shift, const float* skip, const float* scale_s,
                              const float* shift_s, float* out, long long n, int C, int device, void* stream) {
    XCP_REQUIRE(n > 0 && C > 0 && n % C == 0, "xcp_f32_bn_add: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_bn_add_kernel<<<grid_for(n, 256), 256, 0, ST>>>(y, scale, shift, skip, scale_s, shift_s, out, n, C);
    return check_cuda(cudaGetLastError(), "f32_bn_add launch");
}

extern "C" int xcp_f32_bn_relu_gap(const float* y, const float* scale, const float* shift, float* feat, int F, int HW, int C,
                                   int device, void* stream) {
    XCP_REQUIRE(F > 0 && HW > 0 && C > 0, "xcp_f32_bn_relu_gap: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_bn_relu_gap_kernel<<<grid_for((long long)F * C, 256), 256, 0, ST>>>(y, scale, shift, feat, F, HW, C);
    return check_cuda(cudaGetLastError(), "f32_bn_relu_gap launch");
}

extern "C" int xcp_f32_bn_bwd(int mode, const float* y, const float* G, const void* idx, const float* dfeat, const float* scale,
                              const float* shift, const float* gamma, const float* mean, const float* rstd, int training,
                              const float* presums, float* sums_ws, float* coef, float* dgamma, float* dbeta, float* dy, int F, int H,
                              int W, int C, int c_real, int grid_w, int grid_h, int device, void* stream) {
    XCP_REQUIRE(mode >= 0 && mode <= 3 && F > 0 && H > 0 && W > 0 && C > 0 && c_real > 0 && c_real <= C, "xcp_f32_bn_bwd: bad shape");
    XCP_REQUIRE(coef != nullptr && (presums != nullptr || sums_ws != nullptr), "xcp_f32_bn_bwd: workspace missing");
    XCP_CUDA(cudaSetDevice(device));
    BnSrc s{mode, G, (const unsigned char*)idx, dfeat, scale, shift, F, H, W, C};
    const float* sums = presums;
    if (sums == nullptr) {
        f32_bn_bwd_reduce_kernel<<<(C + 31) / 32, 256, 0, ST>>>(y, s, sums_ws);
        sums = sums_ws;
    }
    f32_bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(sums, C, c_real, (double)F * H * W, gamma, mean, rstd, training, coef,
                                                                 dgamma, dbeta);
    if (dy != nullptr)
        f32_bn_bwd_apply_kernel<<<grid_for((long long)F * H * W * C, 256), 256, 0, ST>>>(y, s, coef, dy, grid_w, grid_h);
    return check_cuda(cudaGetLastError(), "f32_bn_bwd launch");
}

// ~ xcp_bn_bwd_sums: sums[0][c] = sum G, sums[1][c] = sum G*y over [n_pix, C]
extern "C" int xcp_f32_bn_bwd_sums(const float* y, const float* G, float* sums, long long n_pix, int C, int device, void* stream) {
    XCP_REQUIRE(n_pix > 0 && n_pix < (1LL << 31) && C > 0, "xcp_f32_bn_bwd_sums: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    BnSrc s{SRC_DIRECT, G, nullptr, nullptr, nullptr, nullptr, 1, 1, (int)n_pix, C};
    f32_bn_bwd_reduce_kernel<<<(C + 31) / 32, 256, 0, ST>>>(y, s, sums);
    return check_cuda(cudaGetLastError(), "f32_bn_bwd_sums launch");
}

extern "C" int xcp_f32_dw3x3_bwd(const float* dD, const float* xin, const float* w9, const float* scale, const float* shift, int relu,
                                 float* dz, const float* add_full, const float* add_half, float* dw, float* bnsum, int F, int H, int W,
                                 int C, int c_real, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && c_real > 0 && c_real <= C, "xcp_f32_dw3x3_bwd: bad shape");
    XCP_REQUIRE(dw != nullptr && (scale == nullptr || bnsum != nullptr), "xcp_f32_dw3x3_bwd: dw / bnsum missing");
    XCP_CUDA(cudaSetDevice(device));
    f32_dw_bwd_dz_kernel<<<grid_for((long long)F * H * W * C, 256), 256, 0, ST>>>(dD, xin, w9, scale, shift, relu, dz, add_full, add_half,
                                                                                F, H, W, C);
    dim3 grid((unsigned)((C + 31) / 32), scale != nullptr ? 11u : 9u);
    f32_dw_bwd_red_kernel<<<grid, 256, 0, ST>>>(dD, xin, scale, shift, relu, dz, dw, bnsum, F, H, W, C, c_real);
    return check_cuda(cudaGetLastError(), "f32_dw3x3_bwd launch");
}

extern "C" int xcp_f32_gemm_wgrad(const float* dY, long long ld_dy, const float* X, long long ld_x, float* dW, long long ld_dw,
                                  long long R, int P, int Q, int device, void* stream) {
    XCP_REQUIRE(R > 0 && P > 0 && Q > 0 && ld_dy >= P && ld_x >= Q && ld_dw >= Q, "xcp_f32_gemm_wgrad: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((unsigned)((Q + 63) / 64), (unsigned)((P + 63) / 64));
    f32_gemm_wgrad_kernel<<<grid, 256, 0, ST>>>(dY, ld_dy, X, ld_x, dW, ld_dw, R, P, Q);
    return check_cuda(cudaGetLastError(), "f32_gemm_wgrad launch");
}

extern "C" int xcp_f32_conv3x3_dgrad(const float* dy, const float* w, float* dx, int F, int H, int W, int Ci, int Co, int device,
                                     void* stream) {
    XCP_REQUIRE(F > 0 && H >= 3 && W >= 3 && Ci > 0 && Co > 0, "xcp_f32_conv3x3_dgrad: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_conv3x3_dgrad_kernel<<<grid_for((long long)F * H * W * Ci, 256), 256, 0, ST>>>(dy, w, dx, F, H, W, Ci, Co, H - 2, W - 2);
    return check_cuda(cudaGetLastError(), "f32_conv3x3_dgrad launch");
}

extern "C" int xcp_f32_conv3x3_wgrad(const float* x, int x_nchw, const float* dy, float* dw, int F, int H, int W, int Ci, int Co,
                                     int stride, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H >= 3 && W >= 3 && Ci > 0 && (Co == 32 || Co == 64 || Co == 128 || Co == 256) && stride > 0,
                "xcp_f32_conv3x3_wgrad: bad shape (Co must divide 256)");
    XCP_CUDA(cudaSetDevice(device));
    const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
    f32_conv3x3_wgrad_kernel<<<Ci * 9, 256, 0, ST>>>(x, x_nchw, dy, dw, F, H, W, Ci, Co, stride, Ho, Wo);
    return check_cuda(cudaGetLastError(), "f32_conv3x3_wgrad launch");
}

extern "C" int xcp_split3_bf16(const float* x, void* out, long long rows, int cols, int side, int along_rows, int device,
                               void* stream) {
    XCP_REQUIRE(rows > 0 && cols > 0 && (side == 0 || side == 1), "xcp_split3_bf16: bad arguments");
    XCP_CUDA(cudaSetDevice(device));
    split3_kernel<<<grid_for(rows * cols, 256), 256, 0, ST>>>(x, (__nv_bfloat16*)out, rows, cols, side, along_rows);
    return check_cuda(cudaGetLastError(), "split3 launch");
}
