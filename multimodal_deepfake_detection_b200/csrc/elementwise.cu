// Memory-bound glue kernels of the Xception path on NHWC bf16 activations: stem conv1, BatchNorm
// statistics finalisation (train) / affine folding (eval), BN+ReLU materialisation, stride-2 gather for the
// skip 1x1 convs, fused BN + MaxPool(3,2,1) + skip-BN + residual add, fused BN + identity residual add,
// BN + ReLU + global-average-pool, and the backward counterparts (two-pass BatchNorm backward with the
// max-pool / GAP / ReLU gradient routing folded into its loads), layout converters and weight packing.
//
// Reference call sites: Xception.py:118-123,168-174 (stem), :56,67,73,78 (BN), :86 (MaxPool2d(3,2,1)),
// :92-98 (skip + add), :197 (adaptive_avg_pool2d).  BatchNorm conventions: SURVEY.md App. E.
#include "common.cuh"

namespace xcp {

// (stem conv1: csrc/stem_conv1.cu)

// ======================================================================================== BN finalize
// partials: [nparts][2][C].  Train mode: batch mean / biased var -> scale, shift, saved mean, rstd; running stats
// updated with momentum and the unbiased variance.  block = 32 channels x 32 part-lanes, so the
// reduction over (up to ~1400) partial rows is ~nparts/32 independent coalesced loads per thread.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partials, int nparts, int C, int C_real, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps,
                   float* scale, float* shift, float* mean_out, float* rstd_out) {
    // block = 8 channels x 128 part-lanes (the first version used 32 channels x 32 part-lanes: 24 CTAs for 768 channels, ~23
    // dependent load batches per thread over the 722 per-tile partial rows of a middle-flow layer, 11.8 us per launch x 40
    // launches per step, profiles/r2p_launches_B16.md; now 96 CTAs x <= 6 rows per thread)
    __shared__ double s1[128][9], s2[128][9];
    const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cx;
    double a = 0.0, b = 0.0;
    if (c < C) {
        int pi = ry;
        for (; pi + 128 < nparts; pi += 256) {
            const float a0 = partials[((long long)pi * 2 + 0) * C + c], b0 = partials[((long long)pi * 2 + 1) * C + c];
            const float a1 = partials[((long long)(pi + 128) * 2 + 0) * C + c], b1 = partials[((long long)(pi + 128) * 2 + 1) * C + c];
            a += (double)a0 + (double)a1;
            b += (double)b0 + (double)b1;
        }
        for (; pi < nparts; pi += 128) {
            a += (double)partials[((long long)pi * 2 + 0) * C + c];
            b += (double)partials[((long long)pi * 2 + 1) * C + c];
        }
    }
    s1[ry][cx] = a; s2[ry][cx] = b;
    __syncthreads();
    // warp w < 8 sums channel w's 128 lane partials
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (w >= 8) return;
    a = (s1[ln][w] + s1[ln + 32][w]) + (s1[ln + 64][w] + s1[ln + 96][w]);
    b = (s2[ln][w] + s2[ln + 32][w]) + (s2[ln + 64][w] + s2[ln + 96][w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    const int cc = blockIdx.x * 8 + w;
    const int cx0 = ln;     // lane 0 of the warp finalizes
    if (cx0 == 0 && cc >= C_real && cc < C) {       // zero-padded channels (physical width > logical): stay exactly 0
        scale[cc] = 0.f; shift[cc] = 0.f; mean_out[cc] = 0.f; rstd_out[cc] = 0.f;
    }
    if (cx0 == 0 && cc < C_real) {
        const double mean = a / count;
        double var = b / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = gamma[cc] * rstd;
        scale[cc] = sc;
        shift[cc] = beta[cc] - (float)mean * sc;
        mean_out[cc] = (float)mean;
        rstd_out[cc] = rstd;
        if (running_mean != nullptr) {
            const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
            running_mean[cc] = (1.f - momentum) * running_mean[cc] + momentum * (float)mean;
            running_var[cc] = (1.f - momentum) * running_var[cc] + momentum * (float)unbiased;
        }
    }
}

__global__ void bn_eval_affine_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                      float* scale, float* shift, float* mean_out, float* rstd_out, int C, int C_real) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (c >= C_real) {
        scale[c] = 0.f; shift[c] = 0.f;
        if (mean_out) mean_out[c] = 0.f;
        if (rstd_out) rstd_out[c] = 0.f;
        return;
    }
    const float rstd = rsqrtf(rv[c] + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - rm[c] * sc;
    if (mean_out) mean_out[c] = rm[c];
    if (rstd_out) rstd_out[c] = rstd;
}

// ======================================================================================== elementwise fwd
XCP_DEVINL void load_affine8(const float* scale, const float* shift, int c0, float (&sc)[8], float (&sh)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(scale + c0), b = *reinterpret_cast<const float4*>(scale + c0 + 4);
    const float4 c = *reinterpret_cast<const float4*>(shift + c0), d = *reinterpret_cast<const float4*>(shift + c0 + 4);
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
    sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
}

// out = relu?(scale*y + shift)      (n8 = number of 8-channel vectors, C % 8 == 0)
__global__ void bn_act_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                              int relu, uint4* __restrict__ out, long long n8, int C) {
    const int ncg = C >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % ncg) * 8;
        float v[8], sc[8], sh[8];
        unpack8(ldg_nc_v4(y + i), v);
        load_affine8(scale, shift, c0, sc, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = fmaf(v[j], sc[j], sh[j]);
            if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        out[i] = pack8(v);
    }
}

// out[f,ho,wo,:] = act(x[f,2ho,2wo,:])   (the input sampling of a 1x1 stride-2 conv)
__global__ void gather_s2_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                 int relu, uint4* __restrict__ out, int F, int H, int W, int C) {
    const int ncg = C >> 3, Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const long long n8 = (long long)F * Ho * Wo * ncg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % ncg);
        long long t = i / ncg;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int f = (int)(t / Ho);
        float v[8];
        unpack8(ldg_nc_v4(x + (((long long)f * H + 2 * ho) * W + 2 * wo) * ncg + cg), v);
        if (scale != nullptr) {
            float sc[8], sh[8];
            load_affine8(scale, shift, cg * 8, sc, sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        out[i] = pack8(v);
    }
}

// out[f,ho,wo,:] = maxpool3x3s2p1( scale*y + shift )[f,ho,wo,:] + (scale_s*ys + shift_s)[f,ho,wo,:]
// idx (uint8 per element) records the arg-max tap (first maximum in row-major window order) for backward.
// IDX = false (inference: no arg-max, no y[arg-max]) drops the index bookkeeping from the instruction stream.
template <bool IDX>
__global__ void __launch_bounds__(256, 4) pool_add_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const uint4* __restrict__ ys, const float* __restrict__ scale_s,
                                    const float* __restrict__ shift_s, uint4* __restrict__ out, uint2* __restrict__ idx,
                                    uint4* __restrict__ ymax, int F, int H, int W, int C, int fast) {
    const int ncg = C >> 3, Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long n8 = (long long)F * Ho * Wo * ncg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % ncg);
        long long t = i / ncg;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int f = (int)(t / Ho);
        float sc[8], sh[8];
        load_affine8(scale, shift, cg * 8, sc, sh);
        // z = sc*y + sh is monotonic in y, so the window maximum is taken on the RAW bf16 values, two channels per
        // instruction (HMNMX2 + a packed compare mask for the arg-max), and the affine is applied once to the winner.
        // Channels with a negative scale have their sign flipped on load so that "max" is right for them too.
        // (The scalar version spent 45 instructions per output element, 5 per tap and channel, and ran at 40 % of HBM.)
        uint32_t sgn[4], best[4], bi[4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            sgn[pp] = (sc[2 * pp] < 0.f ? 0x8000u : 0u) | (sc[2 * pp + 1] < 0.f ? 0x80000000u : 0u);
            bi[pp] = 0u;
        }
        // all 9 taps are loaded up front from clamped (always valid) addresses and masked afterwards: nine independent
        // 16-byte loads in flight per thread instead of a chain of bounds-checked ones
        uint4 raw[9];
        bool ok[9];
        const bool interior = fast && ho > 0 && wo > 0 && 2 * ho + 1 < H && 2 * wo + 1 < W;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h = 2 * ho - 1 + kh;
            const int hc = min(max(h, 0), H - 1);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w = 2 * wo - 1 + kw;
                const int wc = min(max(w, 0), W - 1);
                ok[kh * 3 + kw] = (h == hc) && (w == wc);
                raw[kh * 3 + kw] = __ldg(y + (((long long)f * H + hc) * W + wc) * ncg + cg);
            }
        }
        if (interior && (sgn[0] | sgn[1] | sgn[2] | sgn[3]) == 0u) {
            // whole window inside the image, no negative scale in this channel group (all but a few per cent of the threads):
            // no border select, no sign flip -- compare-mask, max and index merge are the only per-tap instructions (3 of 5)
            best[0] = raw[0].x; best[1] = raw[0].y; best[2] = raw[0].z; best[3] = raw[0].w;
#pragma unroll
            for (int k = 1; k < 9; ++k) {
                const uint32_t wv[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
                const uint32_t kk = (uint32_t)k * 0x00010001u;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    __nv_bfloat162 vb, bb;
                    *reinterpret_cast<uint32_t*>(&vb) = wv[pp];
                    *reinterpret_cast<uint32_t*>(&bb) = best[pp];
                    if (IDX) {
                        const uint32_t m = __hgt2_mask(vb, bb);                // 0xffff per half where v > best (first maximum wins ties)
                        bi[pp] = (bi[pp] & ~m) | (kk & m);
                    }
                    const __nv_bfloat162 mx = __hmax2(vb, bb);
                    best[pp] = *reinterpret_cast<const uint32_t*>(&mx);
                }
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) best[pp] = 0xff80ff80u;         // (-inf, -inf)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const uint32_t wv[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
                const uint32_t kk = (uint32_t)k * 0x00010001u;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const uint32_t v = ok[k] ? (wv[pp] ^ sgn[pp]) : 0xff80ff80u;
                    __nv_bfloat162 vb, bb;
                    *reinterpret_cast<uint32_t*>(&vb) = v;
                    *reinterpret_cast<uint32_t*>(&bb) = best[pp];
                    const uint32_t m = __hgt2_mask(vb, bb);                    // 0xffff per half where v > best (first maximum wins ties)
                    const __nv_bfloat162 mx = __hmax2(vb, bb);
                    best[pp] = *reinterpret_cast<const uint32_t*>(&mx);
                    bi[pp] = (bi[pp] & ~m) | (kk & m);
                }
            }
        }
        float bestf[8];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const uint32_t yv = best[pp] ^ sgn[pp];
            bestf[2 * pp] = fmaf(bf16_lo(yv), sc[2 * pp], sh[2 * pp]);
            bestf[2 * pp + 1] = fmaf(bf16_hi(yv), sc[2 * pp + 1], sh[2 * pp + 1]);
        }
        float s[8];
        unpack8(ldg_nc_v4(ys + i), s);
        load_affine8(scale_s, shift_s, cg * 8, sc, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) bestf[j] += fmaf(s[j], sc[j], sh[j]);
        out[i] = pack8(bestf);
        if (IDX) {
            if (idx != nullptr) {       // bi[pp] = (tap of channel 2pp) | (tap of channel 2pp+1) << 16  ->  one byte per channel
                uint2 o;
                o.x = __byte_perm(bi[0], bi[1], 0x6420);
                o.y = __byte_perm(bi[2], bi[3], 0x6420);
                idx[i] = o;
            }
            // the RAW winner y[arg-max]: the BatchNorm backward through the pool needs sum dz*y = sum_windows G * y[arg-max],
            // which it can then take from this quarter-size tensor instead of re-reading all of y (xcp_bn_bwd_sums)
            if (ymax != nullptr) ymax[i] = make_uint4(best[0] ^ sgn[0], best[1] ^ sgn[1], best[2] ^ sgn[2], best[3] ^ sgn[3]);
        }
    }
}

// out = scale*y + shift + (scale_s*skip + shift_s)      (scale_s/shift_s optional: identity skip)
__global__ void bn_add_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                  const uint4* __restrict__ skip, const float* __restrict__ scale_s,
                                  const float* __restrict__ shift_s, uint4* __restrict__ out, long long n8, int C) {
    const int ncg = C >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % ncg) * 8;
        float v[8], s[8], sc[8], sh[8];
        unpack8(ldg_nc_v4(y + i), v);
        unpack8(ldg_nc_v4(skip + i), s);
        if (scale_s != nullptr) {
            load_affine8(scale_s, shift_s, c0, sc, sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = fmaf(s[j], sc[j], sh[j]);
        }
        load_affine8(scale, shift, c0, sc, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]) + s[j];
        out[i] = pack8(v);
    }
}

// feat[f,c] = mean_hw relu(scale*y + shift)       grid = (ceil(ncg/32), F), block = (32 cgs, 8 pixel lanes)
__global__ void __launch_bounds__(256)
bn_relu_gap_kernel(const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                   float* __restrict__ feat, int HW, int C) {
    __shared__ float s_acc[8][32][8];
    const int ncg = C >> 3;
    const int cgl = threadIdx.x & 31, lane_p = threadIdx.x >> 5;
    const int cg = blockIdx.x * 32 + cgl;
    const int f = blockIdx.y;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (cg < ncg) {
        float sc[8], sh[8];
        load_affine8(scale, shift, cg * 8, sc, sh);
        for (int pidx = lane_p; pidx < HW; pidx += 8) {
            float v[8];
            unpack8(ldg_nc_v4(y + ((long long)f * HW + pidx) * ncg + cg), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[lane_p][cgl][j] = acc[j];
    __syncthreads();
    if (lane_p == 0 && cg < ncg) {
        const float inv = 1.f / (float)HW;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = 0.f;
            for (int r = 0; r < 8; ++r) s += s_acc[r][cgl][j];
            feat[(long long)f * C + cg * 8 + j] = s * inv;
        }
    }
}

// ======================================================================================== BN backward
// Source of dz (the gradient wrt the BN output z = scale*y + shift):
//   0 DIRECT     dz = G
//   1 RELU       dz = G * [z > 0]                       (x = relu(bn(y)) was consumed, G = dL/dx)
//   2 POOL       dz[h,w] = sum over the <=4 pooling windows containing (h,w) whose arg-max is (h,w) of G[window]
//   3 GAP_RELU   dz = dfeat[f,c] / HW * [z > 0]         (GAP over relu(bn(y)))
enum { SRC_DIRECT = 0, SRC_RELU = 1, SRC_POOL = 2, SRC_GAP_RELU = 3 };

struct BnBwdSrc {
    int mode;
    const __nv_bfloat16* G;     // modes 0,1: [F,H,W,C] ; mode 2: [F,Ho,Wo,C]
    const uint8_t* idx;         // mode 2: [F,Ho,Wo,C]
    const float* dfeat;         // mode 3: [F,C]
    const float* scale;         // modes 1,3
    const float* shift;
    int F, H, W, C;
};

XCP_DEVINL void bnbwd_dz8(const BnBwdSrc& s, long long i, int cg, const float (&yv)[8], float (&dz)[8]) {
    const int ncg = s.C >> 3;
    if (s.mode == SRC_DIRECT || s.mode == SRC_RELU) {
        unpack8(ldg_nc_v4(reinterpret_cast<const uint4*>(s.G) + i), dz);
    } else if (s.mode == SRC_POOL) {
        long long t = i / ncg;
        const int w = (int)(t % s.W); t /= s.W;
        const int h = (int)(t % s.H);
        const int f = (int)(t / s.H);
        const int Ho = (s.H - 1) / 2 + 1, Wo = (s.W - 1) / 2 + 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = 0.f;
        const int oh0 = h >> 1, ow0 = w >> 1;          // window o covers inputs 2o-1 .. 2o+1
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int oh = oh0 + a;
            if (a == 1 && !(h & 1)) continue;          // even h belongs to exactly one window row
            if (oh >= Ho) continue;
            const int kh = h - (2 * oh - 1);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ow = ow0 + b;
                if (b == 1 && !(w & 1)) continue;
                if (ow >= Wo) continue;
                const int kw = w - (2 * ow - 1);
                const long long o = (((long long)f * Ho + oh) * Wo + ow) * ncg + cg;
                const uint2 id = __ldg(reinterpret_cast<const uint2*>(s.idx) + o);
                float g[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(s.G) + o), g);
                const uint32_t want = (uint32_t)(kh * 3 + kw);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t got = ((j < 4 ? id.x : id.y) >> ((j & 3) * 8)) & 0xffu;
                    if (got == want) dz[j] += g[j];
                }
            }
        }
    } else {  // SRC_GAP_RELU
        const long long pix = i / ncg;
        const int f = (int)(pix / ((long long)s.H * s.W));
        const float inv = 1.f / (float)(s.H * s.W);
        const float4 a = *reinterpret_cast<const float4*>(s.dfeat + (long long)f * s.C + cg * 8);
        const float4 b = *reinterpret_cast<const float4*>(s.dfeat + (long long)f * s.C + cg * 8 + 4);
        dz[0] = a.x * inv; dz[1] = a.y * inv; dz[2] = a.z * inv; dz[3] = a.w * inv;
        dz[4] = b.x * inv; dz[5] = b.y * inv; dz[6] = b.z * inv; dz[7] = b.w * inv;
    }
    if (s.mode == SRC_RELU || s.mode == SRC_GAP_RELU) {
        float sc[8], sh[8];
        load_affine8(s.scale, s.shift, cg * 8, sc, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (!(fmaf(yv[j], sc[j], sh[j]) > 0.f)) dz[j] = 0.f;
    }
}

// ReLU mask of modes 1,3 applied to an already loaded dz vector
XCP_DEVINL void bnbwd_mask8(const BnBwdSrc& s, int cg, const float (&yv)[8], float (&dz)[8]) {
    float sc[8], sh[8];
    load_affine8(s.scale, s.shift, cg * 8, sc, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (!(fmaf(yv[j], sc[j], sh[j]) > 0.f)) dz[j] = 0.f;
}

// pass 1: per-channel (sum dz, sum dz*y) -> partials[gridDim.x][2][C].
// A thread stays on one channel group (stride S is a multiple of ncg) and keeps BNBWD_U independent 16-byte loads of y
// (and of G in the direct / ReLU modes) in flight: the first version had one load pair per thread outstanding and
// reached 47 % of the DRAM peak with the issue slots 18 % busy (ncu, profiles/r1j), i.e. it was latency bound.
constexpr int BNBWD_U = 4;
__global__ void __launch_bounds__(256, 2)
bnbwd_reduce_kernel(const uint4* __restrict__ y, const BnBwdSrc s, float* __restrict__ partials, long long n8) {
    extern __shared__ float s_acc[];   // [2][C]
    const int C = s.C, ncg = C >> 3;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long S = T - (T % ncg);                 // stride that keeps a thread on one channel group
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < S) {
        const int cg = (int)(gid % ncg);
        float a1[8], a2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
        const bool paired = (s.mode == SRC_DIRECT || s.mode == SRC_RELU);      // G has y's shape and index
        const uint4* Gv = reinterpret_cast<const uint4*>(s.G);
        long long i = gid;
        for (; i + (BNBWD_U - 1) * S < n8; i += BNBWD_U * S) {
            uint4 yr[BNBWD_U], gr[BNBWD_U];
#pragma unroll
            for (int u = 0; u < BNBWD_U; ++u) yr[u] = ldg_nc_v4(y + i + u * S);
            if (paired) {
#pragma unroll
                for (int u = 0; u < BNBWD_U; ++u) gr[u] = ldg_nc_v4(Gv + i + u * S);
            }
#pragma unroll
            for (int u = 0; u < BNBWD_U; ++u) {
                float yv[8], dz[8];
                unpack8(yr[u], yv);
                if (paired) {
                    unpack8(gr[u], dz);
                    if (s.mode == SRC_RELU) bnbwd_mask8(s, cg, yv, dz);
                } else {
                    bnbwd_dz8(s, i + u * S, cg, yv, dz);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { a1[j] += dz[j]; a2[j] = fmaf(dz[j], yv[j], a2[j]); }
            }
        }
        for (; i < n8; i += S) {
            float yv[8], dz[8];
            unpack8(ldg_nc_v4(y + i), yv);
            bnbwd_dz8(s, i, cg, yv, dz);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a1[j] += dz[j]; a2[j] = fmaf(dz[j], yv[j], a2[j]); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[cg * 8 + j], a1[j]); atomicAdd(&s_acc[C + cg * 8 + j], a2[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partials[(long long)blockIdx.x * 2 * C + i] = s_acc[i];
}

// MaxPool(3,2,1) source, quad formulation.  One thread owns a 2x2 quad of input pixels x 8 channels: the quad (2a..2a+1, 2b..2b+1)
// is covered by exactly the four pooling windows (a,b), (a,b+1), (a+1,b), (a+1,b+1), so 4 (idx, G) loads serve 4 input pixels
// (9 window/tap combinations) with no per-pixel parity branches.  The generic gather above loads up to 4 windows PER PIXEL behind
// divergent branches and ran the b1/b2/b3/b12 BatchNorm backward passes at ~25 % of the HBM roofline (1.04 ms for 147^2x128 at
// 128 frames, profiles/r1x); both passes (APPLY = false: column sums, true: dy) use this kernel.
template <bool APPLY>
__global__ void __launch_bounds__(256, 2)
bnbwd_pool_kernel(const uint4* __restrict__ y, const BnBwdSrc s, float* __restrict__ partials, const float* __restrict__ coefA,
                  const float* __restrict__ coefB, const float* __restrict__ coefC, uint4* __restrict__ dy, long long nq8) {
    extern __shared__ float s_acc[];   // [2][C] (reduce pass)
    const int C = s.C, ncg = C >> 3, H = s.H, W = s.W;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, Hq = (H + 1) / 2, Wq = (W + 1) / 2;
    if (!APPLY) {
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
        __syncthreads();
    }
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long S = T - (T % ncg);                 // a thread stays on one channel group
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < S) {
        const int cg = (int)(gid % ncg);
        float a1[8], a2[8], A[8], B[8], Cc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
        if (APPLY) {
            load_affine8(coefA, coefB, cg * 8, A, B);
            const float4 c0 = *reinterpret_cast<const float4*>(coefC + cg * 8), c1 = *reinterpret_cast<const float4*>(coefC + cg * 8 + 4);
            Cc[0] = c0.x; Cc[1] = c0.y; Cc[2] = c0.z; Cc[3] = c0.w; Cc[4] = c1.x; Cc[5] = c1.y; Cc[6] = c1.z; Cc[7] = c1.w;
        }
        const uint2* idxv = reinterpret_cast<const uint2*>(s.idx);
        const uint4* Gv = reinterpret_cast<const uint4*>(s.G);
        for (long long qi = gid; qi < nq8; qi += S) {
            long long t = qi / ncg;
            const int b = (int)(t % Wq); t /= Wq;
            const int a = (int)(t % Hq);
            const long long f = t / Hq;
            const int h0 = 2 * a, w0 = 2 * b;
            const bool h1 = h0 + 1 < H, w1 = w0 + 1 < W;          // second row / column of the quad inside the image
            const bool oh1 = a + 1 < Ho, ow1 = b + 1 < Wo;        // windows (a+1, .), (., b+1) exist
            // ---- all loads first (independent addresses)
            const long long p00 = ((f * H + h0) * W + w0) * ncg + cg;
            uint4 yr[4];
            yr[0] = ldg_nc_v4(y + p00);
            yr[1] = w1 ? ldg_nc_v4(y + p00 + ncg) : make_uint4(0, 0, 0, 0);
            yr[2] = h1 ? ldg_nc_v4(y + p00 + (long long)W * ncg) : make_uint4(0, 0, 0, 0);
            yr[3] = (h1 && w1) ? ldg_nc_v4(y + p00 + (long long)W * ncg + ncg) : make_uint4(0, 0, 0, 0);
            const long long o00 = ((f * Ho + a) * Wo + b) * ncg + cg;
            const bool wv[4] = {true, ow1, oh1, oh1 && ow1};
            const long long wo[4] = {o00, o00 + ncg, o00 + (long long)Wo * ncg, o00 + (long long)Wo * ncg + ncg};
            uint2 id[4]; uint4 gr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                id[k] = wv[k] ? __ldg(idxv + wo[k]) : make_uint2(0xffffffffu, 0xffffffffu);     // 0xff matches no tap
                gr[k] = wv[k] ? __ldg(Gv + wo[k]) : make_uint4(0, 0, 0, 0);
            }
            // ---- route: dz of the 4 quad pixels from the (window, tap) pairs that can select them
            float g[4][8];
#pragma unroll
            for (int k = 0; k < 4; ++k) unpack8(gr[k], g[k]);
            float dz[4][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t tp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tp[k] = ((j < 4 ? id[k].x : id[k].y) >> ((j & 3) * 8)) & 0xffu;
                dz[0][j] = (tp[0] == 4u ? g[0][j] : 0.f);
                dz[1][j] = (tp[0] == 5u ? g[0][j] : 0.f) + (tp[1] == 3u ? g[1][j] : 0.f);
                dz[2][j] = (tp[0] == 7u ? g[0][j] : 0.f) + (tp[2] == 1u ? g[2][j] : 0.f);
                dz[3][j] = (tp[0] == 8u ? g[0][j] : 0.f) + (tp[1] == 6u ? g[1][j] : 0.f) + (tp[2] == 2u ? g[2][j] : 0.f) +
                           (tp[3] == 0u ? g[3][j] : 0.f);
            }
            const bool pv[4] = {true, w1, h1, h1 && w1};
            const long long po[4] = {p00, p00 + ncg, p00 + (long long)W * ncg, p00 + (long long)W * ncg + ncg};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!pv[k]) continue;
                float yv[8];
                unpack8(yr[k], yv);
                if (APPLY) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], dz[k][j], fmaf(B[j], yv[j], Cc[j]));
                    dy[po[k]] = pack8(o);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a1[j] += dz[k][j]; a2[j] = fmaf(dz[k][j], yv[j], a2[j]); }
                }
            }
        }
        if (!APPLY) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[cg * 8 + j], a1[j]); atomicAdd(&s_acc[C + cg * 8 + j], a2[j]); }
        }
    }
    if (!APPLY) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partials[(long long)blockIdx.x * 2 * C + i] = s_acc[i];
    }
}

// pass 1b: fold partials -> coefficients of dy = A*dz + B*y + Cc and the BN parameter gradients (accumulated).
// nparts == 1: `partials` is already the [2][C] sums (e.g. produced by the depthwise backward kernel).
// block = 32 channels x 8 part-lanes.
__global__ void __launch_bounds__(256)
bnbwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, int C_real, double count,
                      const float* __restrict__ gamma, const float* __restrict__ mean,
                      const float* __restrict__ rstd, int training, float* coefA, float* coefB, float* coefC,
                      float* dgamma, float* dbeta) {
    __shared__ double sh1[8][32], sh2[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        int pi = ry;
        for (; pi + 24 < nparts; pi += 32) {
            const float a0 = partials[(long long)pi * 2 * C + c], b0 = partials[(long long)pi * 2 * C + C + c];
            const float a1 = partials[(long long)(pi + 8) * 2 * C + c], b1 = partials[(long long)(pi + 8) * 2 * C + C + c];
            const float a2 = partials[(long long)(pi + 16) * 2 * C + c], b2 = partials[(long long)(pi + 16) * 2 * C + C + c];
            const float a3 = partials[(long long)(pi + 24) * 2 * C + c], b3 = partials[(long long)(pi + 24) * 2 * C + C + c];
            s1 += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
            s2 += ((double)b0 + (double)b1) + ((double)b2 + (double)b3);
        }
        for (; pi < nparts; pi += 8) {
            s1 += (double)partials[(long long)pi * 2 * C + c];
            s2 += (double)partials[(long long)pi * 2 * C + C + c];
        }
    }
    sh1[ry][cx] = s1; sh2[ry][cx] = s2;
    __syncthreads();
    if (ry != 0 || c >= C) return;
    if (c >= C_real) { coefA[c] = 0.f; coefB[c] = 0.f; coefC[c] = 0.f; return; }     // zero-padded channels
#pragma unroll
    for (int r = 1; r < 8; ++r) { s1 += sh1[r][cx]; s2 += sh2[r][cx]; }
    const double m = mean[c], r = rstd[c], g = gamma[c];
    const double dg = r * (s2 - m * s1);      // sum dz * xhat
    const double A = g * r;
    double B = 0.0, Cc = 0.0;
    if (training) {
        B = -g * r * r * dg / count;
        Cc = -B * m - A * s1 / count;
    }
    coefA[c] = (float)A; coefB[c] = (float)B; coefC[c] = (float)Cc;
    if (dgamma != nullptr) dgamma[c] += (float)dg;
    if (dbeta != nullptr) dbeta[c] += (float)s1;
}

// pass 2: dy = A*dz + B*y + Cc  (bf16).  conv_grid_w > 0: scatter rows onto a zero-initialised
// (conv_grid_h x conv_grid_w) "input grid" layout used by the stem implicit-GEMM backward.
// When the per-channel sums are already complete (one [2][C] row: produced by the depthwise backward or xcp_bn_bwd_sums), the
// finalize step -- a few flops per channel -- runs inside the apply kernel: every thread derives the coefficients of its own 8
// channels, the first thread of each channel group adds dgamma / dbeta.  One launch (and its ~7 us) less per BatchNorm.
struct BnBwdFin {
    const float* sums;            // [2][C] or null (= read the coefficients the finalize kernel wrote)
    const float* gamma; const float* mean; const float* rstd;
    float* dgamma; float* dbeta;
    double count; int training, c_real;
};

__global__ void __launch_bounds__(256, 2)
bnbwd_apply_kernel(const uint4* __restrict__ y, const BnBwdSrc s, const float* __restrict__ coefA,
                   const float* __restrict__ coefB, const float* __restrict__ coefC, uint4* __restrict__ dy, long long n8,
                   int grid_w, int grid_h, const BnBwdFin fin) {
    const int ncg = s.C >> 3;
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long S = T - (T % ncg);                 // a thread stays on one channel group: coefficients live in registers
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= S) return;
    const int cg = (int)(gid % ncg);
    float A[8], B[8], Cc[8];
    if (fin.sums != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            A[j] = 0.f; B[j] = 0.f; Cc[j] = 0.f;
            if (c < fin.c_real) {                      // same arithmetic as bnbwd_finalize_kernel (fp64 on a handful of values)
                const double s1 = fin.sums[c], s2 = fin.sums[s.C + c];
                const double m = fin.mean[c], r = fin.rstd[c], g = fin.gamma[c];
                const double dg = r * (s2 - m * s1);
                const double a = g * r;
                double b = 0.0, cc = 0.0;
                if (fin.training) { b = -g * r * r * dg / fin.count; cc = -b * m - a * s1 / fin.count; }
                A[j] = (float)a; B[j] = (float)b; Cc[j] = (float)cc;
                if (gid < ncg) {                       // one thread per channel group owns the parameter gradients
                    if (fin.dgamma != nullptr) fin.dgamma[c] += (float)dg;
                    if (fin.dbeta != nullptr) fin.dbeta[c] += (float)s1;
                }
            }
        }
    } else {
    load_affine8(coefA, coefB, cg * 8, A, B);
    {
        const float4 c0 = *reinterpret_cast<const float4*>(coefC + cg * 8), c1 = *reinterpret_cast<const float4*>(coefC + cg * 8 + 4);
        Cc[0] = c0.x; Cc[1] = c0.y; Cc[2] = c0.z; Cc[3] = c0.w; Cc[4] = c1.x; Cc[5] = c1.y; Cc[6] = c1.z; Cc[7] = c1.w;
    }
    }
    const bool paired = (s.mode == SRC_DIRECT || s.mode == SRC_RELU);
    const uint4* Gv = reinterpret_cast<const uint4*>(s.G);
    auto emit = [&](long long i, const float (&yv)[8], float (&dz)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = fmaf(A[j], dz[j], fmaf(B[j], yv[j], Cc[j]));
        long long o = i;
        if (grid_w > 0) {
            long long t = i / ncg;
            const int w = (int)(t % s.W); t /= s.W;
            const int h = (int)(t % s.H);
            const long long f = t / s.H;
            o = ((f * grid_h + h) * grid_w + w) * ncg + cg;
        }
        dy[o] = pack8(dz);
    };
    long long i = gid;
    for (; i + (BNBWD_U - 1) * S < n8; i += BNBWD_U * S) {
        uint4 yr[BNBWD_U], gr[BNBWD_U];
#pragma unroll
        for (int u = 0; u < BNBWD_U; ++u) yr[u] = ldg_nc_v4(y + i + u * S);
        if (paired) {
#pragma unroll
            for (int u = 0; u < BNBWD_U; ++u) gr[u] = ldg_nc_v4(Gv + i + u * S);
        }
#pragma unroll
        for (int u = 0; u < BNBWD_U; ++u) {
            float yv[8], dz[8];
            unpack8(yr[u], yv);
            if (paired) {
                unpack8(gr[u], dz);
                if (s.mode == SRC_RELU) bnbwd_mask8(s, cg, yv, dz);
            } else {
                bnbwd_dz8(s, i + u * S, cg, yv, dz);
            }
            emit(i + u * S, yv, dz);
        }
    }
    for (; i < n8; i += S) {
        float yv[8], dz[8];
        unpack8(ldg_nc_v4(y + i), yv);
        bnbwd_dz8(s, i, cg, yv, dz);
        emit(i, yv, dz);
    }
}

// ======================================================================================== layout converters
// NCHW fp32 -> NHWC bf16 via a 32x32 smem transpose over (C, HW) per frame.
// Cp >= C is the physical (padded) channel pitch of the NHWC tensor; pad channels are written as zeros / ignored.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int Cp, int HW) {
    __shared__ float t[32][33];
    const int f = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, pp = p0 + threadIdx.x;
        t[r][threadIdx.x] = (c < C && pp < HW) ? x[((long long)f * C + c) * HW + pp] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int pp = p0 + r, c = c0 + threadIdx.x;
        if (pp < HW && c < Cp) out[((long long)f * HW + pp) * Cp + c] = __float2bfloat16(t[threadIdx.x][r]);
    }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int C, int Cp, int HW) {
    __shared__ float t[32][33];
    const int f = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int pp = p0 + r, c = c0 + threadIdx.x;
        t[r][threadIdx.x] = (c < C && pp < HW) ? __bfloat162float(x[((long long)f * HW + pp) * Cp + c]) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, pp = p0 + threadIdx.x;
        if (pp < HW && c < C) out[((long long)f * C + c) * HW + pp] = t[threadIdx.x][r];
    }
}

// fp32 [R, Cc] -> bf16 [Rp, Cp] (zero padded) and (optionally) its transpose bf16 [Cp, Rp]
// row_scale (optional, [Rp]): every row is multiplied by its entry before rounding -- the inference plan folds the BatchNorm
// scale gamma * rstd of an output channel into the pointwise weights this way (test_visual.py:609-624)
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_t,
                                   int R, int Cc, int Rp, int Cp, const float* __restrict__ row_scale = nullptr) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int rr = r0 + r, cc = c0 + threadIdx.x;
        float v = (rr < R && cc < Cc) ? w[(long long)rr * Cc + cc] : 0.f;
        if (row_scale != nullptr && rr < R) v *= row_scale[rr];
        t[r][threadIdx.x] = v;
        if (rr < Rp && cc < Cp && out != nullptr) out[(long long)rr * Cp + cc] = __float2bfloat16(v);
    }
    __syncthreads();
    if (out_t != nullptr) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int cc = c0 + r, rr = r0 + threadIdx.x;
            if (rr < Rp && cc < Cp) out_t[(long long)cc * Rp + rr] = __float2bfloat16(t[threadIdx.x][r]);
        }
    }
}

// Multi-tensor form of pack_weight / pack_dw: one launch re-packs every pointwise and depthwise weight the optimizer just
// changed (~165 tensors per step, each a 3-4 us launch on its own).  Blocks find their tensor by binary search over the
// table's running tile offsets; kind 0 = 32x32 transpose tiles of a [R,Cc] -> bf16 [Rp,Cp] (+ [Cp,Rp]) matrix,
// kind 1 = 1024-element tiles of a depthwise [C,1,3,3] -> fp32 [9][Cp].
struct PackTensor {
    const float* src;
    void* out;
    void* out_t;
    int R, Cc, Rp, Cp;
    int kind, tile0;
};
static_assert(sizeof(PackTensor) == 48, "PackTensor layout is mirrored on the host (executor.PackCache.prefetch)");

__global__ void __launch_bounds__(256) pack_multi_kernel(const PackTensor* __restrict__ tab, int n) {
    __shared__ float t[32][33];
    const int b = blockIdx.x;
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].tile0 <= b) lo = mid; else hi = mid - 1;
    }
    const PackTensor T = tab[lo];
    const int tile = b - T.tile0;
    if (T.kind == 1) {
        const int C = T.R, Cp = T.Cp, end = min((tile + 1) * 1024, 9 * Cp);
        float* w9 = (float*)T.out;
        for (int i = tile * 1024 + threadIdx.x; i < end; i += 256) {
            const int c = i / 9, k = i % 9;
            w9[(long long)k * Cp + c] = c < C ? T.src[i] : 0.f;
        }
        return;
    }
    const int tiles_x = (T.Cp + 31) >> 5;
    const int r0 = (tile / tiles_x) * 32, c0 = (tile % tiles_x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __nv_bfloat16* out = (__nv_bfloat16*)T.out;
    __nv_bfloat16* out_t = (__nv_bfloat16*)T.out_t;
    for (int r = ty; r < 32; r += 8) {
        const int rr = r0 + r, cc = c0 + tx;
        const float v = (rr < T.R && cc < T.Cc) ? T.src[(long long)rr * T.Cc + cc] : 0.f;
        t[r][tx] = v;
        if (rr < T.Rp && cc < T.Cp && out != nullptr) out[(long long)rr * T.Cp + cc] = __float2bfloat16(v);
    }
    __syncthreads();
    if (out_t != nullptr) {
        for (int r = ty; r < 32; r += 8) {
            const int cc = c0 + r, rr = r0 + tx;
            if (rr < T.Rp && cc < T.Cp) out_t[(long long)cc * T.Rp + rr] = __float2bfloat16(t[tx][r]);
        }
    }
}

// depthwise weights [C,1,3,3] fp32 -> tap-major [9][C] fp32 ; and the reverse accumulation for gradients
__global__ void pack_dw_kernel(const float* __restrict__ w, float* __restrict__ w9, int C, int Cp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * Cp) { const int c = i / 9, k = i % 9; w9[(long long)k * Cp + c] = c < C ? w[i] : 0.f; }
}
__global__ void unpack_dw_grad_kernel(const float* __restrict__ g9, float* __restrict__ gw, int C, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * C) {
        const int c = i / 9, k = i % 9;
        const float v = g9[(long long)k * C + c];
        gw[i] = accumulate ? gw[i] + v : v;
    }
}

// dense 3x3 weights [O,I,3,3] fp32 -> bf16 [O][tap*I + i]   (K-major "tap-major" packing for the implicit GEMM)
// and (optionally) bf16 [I][tap*O + o] for the data-gradient implicit GEMM.
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wk, __nv_bfloat16* __restrict__ wk_t,
                                    int O, int I) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= O * I * 9) return;
    const int tap = idx % 9, i = (idx / 9) % I, o = idx / (9 * I);
    const float v = w[idx];
    wk[(long long)o * (9 * I) + tap * I + i] = __float2bfloat16(v);
    if (wk_t != nullptr) wk_t[(long long)i * (9 * O) + tap * O + o] = __float2bfloat16(v);
}

// gradient of pack_conv3x3: gk fp32 [O][tap*I + i]  ->  gw [O,I,3,3] (+=)
__global__ void unpack_conv3x3_grad_kernel(const float* __restrict__ gk, float* __restrict__ gw, int O, int I) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= O * I * 9) return;
    const int tap = idx % 9, i = (idx / 9) % I, o = idx / (9 * I);
    gw[idx] += gk[(long long)o * (9 * I) + tap * I + i];
}

// bilinear (align_corners=False) upsample of [F,C,n,1] fp32 to [F,C,S,S] fp32 (XceptionLSTMA.py:45-46).
// With input width 1 every output column equals the row value, so only the vertical lerp is computed.
__global__ void bilinear_up_kernel(const float* __restrict__ x, float* __restrict__ out, long long planes, int n, int S) {
    const long long total = planes * S * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)((i / S) % S);
        const long long pl = i / ((long long)S * S);
        float src = ((float)r + 0.5f) * ((float)n / (float)S) - 0.5f;
        if (src < 0.f) src = 0.f;
        int i0 = (int)floorf(src);
        if (i0 > n - 1) i0 = n - 1;
        const int i1 = min(i0 + 1, n - 1);
        const float l1 = src - (float)i0, l0 = 1.f - l1;
        out[i] = l0 * x[pl * n + i0] + l1 * x[pl * n + i1];
    }
}

static int ew_grid(long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = 8LL * num_sms();
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace xcp

using namespace xcp;

#define ST ((cudaStream_t)stream)

extern "C" int xcp_bn_finalize(const float* partials, int nparts, int C, int c_real, double count, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                               float* mean_out, float* rstd_out, int device, void* stream) {
    XCP_REQUIRE(nparts > 0 && C > 0 && count > 0 && c_real > 0 && c_real <= C, "xcp_bn_finalize: bad args");
    XCP_CUDA(cudaSetDevice(device));
    bn_finalize_kernel<<<(C + 7) / 8, 1024, 0, ST>>>(partials, nparts, C, c_real, count, gamma, beta, running_mean, running_var,
                                                      momentum, eps, scale, shift, mean_out, rstd_out);
    return check_cuda(cudaGetLastError(), "bn_finalize launch");
}

extern "C" int xcp_bn_eval_affine(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                  float* scale, float* shift, float* mean_out, float* rstd_out, int C, int c_real, int device,
                                  void* stream) {
    XCP_REQUIRE(c_real > 0 && c_real <= C, "xcp_bn_eval_affine: bad args");
    XCP_CUDA(cudaSetDevice(device));
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, ST>>>(gamma, beta, rm, rv, eps, scale, shift, mean_out, rstd_out, C, c_real);
    return check_cuda(cudaGetLastError(), "bn_eval_affine launch");
}

extern "C" int xcp_bn_act(const void* y, const float* scale, const float* shift, int relu, void* out, long long n, int C,
                          int device, void* stream) {
    XCP_REQUIRE(C % 8 == 0 && n % C == 0, "xcp_bn_act: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    bn_act_kernel<<<ew_grid(n / 8, 256), 256, 0, ST>>>((const uint4*)y, scale, shift, relu, (uint4*)out, n / 8, C);
    return check_cuda(cudaGetLastError(), "bn_act launch");
}

extern "C" int xcp_gather_s2(const void* x, const float* scale, const float* shift, int relu, void* out, int F, int H, int W,
                             int C, int device, void* stream) {
    XCP_REQUIRE(C % 8 == 0, "xcp_gather_s2: C %% 8");
    XCP_CUDA(cudaSetDevice(device));
    const long long n8 = (long long)F * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    gather_s2_kernel<<<ew_grid(n8, 256), 256, 0, ST>>>((const uint4*)x, scale, shift, relu, (uint4*)out, F, H, W, C);
    return check_cuda(cudaGetLastError(), "gather_s2 launch");
}

extern "C" int xcp_pool_add_fwd(const void* y, const float* scale, const float* shift, const void* ys, const float* scale_s,
                                const float* shift_s, void* out, void* idx, void* ymax, int F, int H, int W, int C, int device,
                                void* stream) {
    XCP_REQUIRE(C % 8 == 0, "xcp_pool_add_fwd: C %% 8");
    XCP_CUDA(cudaSetDevice(device));
    const long long n8 = (long long)F * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * (C / 8);
    static int slow_env = -1;                                      // A/B hook: XCP_POOL_SLOW=1 = the round-1 instruction stream
    if (slow_env < 0) { const char* e = getenv("XCP_POOL_SLOW"); slow_env = e ? atoi(e) : 0; }
    if (idx != nullptr || ymax != nullptr || slow_env)
        pool_add_fwd_kernel<true><<<ew_grid(n8, 256), 256, 0, ST>>>((const uint4*)y, scale, shift, (const uint4*)ys, scale_s, shift_s,
                                                                    (uint4*)out, (uint2*)idx, (uint4*)ymax, F, H, W, C, !slow_env);
    else
        pool_add_fwd_kernel<false><<<ew_grid(n8, 256), 256, 0, ST>>>((const uint4*)y, scale, shift, (const uint4*)ys, scale_s, shift_s,
                                                                     (uint4*)out, nullptr, nullptr, F, H, W, C, 1);
    return check_cuda(cudaGetLastError(), "pool_add_fwd launch");
}

extern "C" int xcp_bn_add_fwd(const void* y, const float* scale, const float* shift, const void* skip, const float* scale_s,
                              const float* shift_s, void* out, long long n, int C, int device, void* stream) {
    XCP_REQUIRE(C % 8 == 0 && n % C == 0, "xcp_bn_add_fwd: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    bn_add_fwd_kernel<<<ew_grid(n / 8, 256), 256, 0, ST>>>((const uint4*)y, scale, shift, (const uint4*)skip, scale_s, shift_s,
                                                         (uint4*)out, n / 8, C);
    return check_cuda(cudaGetLastError(), "bn_add_fwd launch");
}

extern "C" int xcp_bn_relu_gap(const void* y, const float* scale, const float* shift, float* feat, int F, int HW, int C,
                               int device, void* stream) {
    XCP_REQUIRE(C % 8 == 0, "xcp_bn_relu_gap: C %% 8");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((C / 8 + 31) / 32, F);
    bn_relu_gap_kernel<<<grid, 256, 0, ST>>>((const uint4*)y, scale, shift, feat, HW, C);
    return check_cuda(cudaGetLastError(), "bn_relu_gap launch");
}

extern "C" int xcp_bnbwd_num_parts(void) { return 2 * 148; }     // resident CTAs of bnbwd_reduce_kernel on a B200 (2 per SM)

// Two-pass BatchNorm backward.  mode: 0 direct, 1 relu-masked, 2 through MaxPool(3,2,1) (G, idx at pooled
// resolution), 3 through GAP+ReLU (dfeat).  Writes dy (bf16) and accumulates dgamma/dbeta.  `presums`
// non-null skips pass 1 and uses those [2][C] sums (produced by xcp_dw3x3_bwd).
extern "C" int xcp_bn_bwd(int mode, const void* y, const void* G, const void* idx, const float* dfeat, const float* scale,
                          const float* shift, const float* gamma, const float* mean, const float* rstd, int training,
                          const float* presums, float* workspace, float* coef, float* dgamma, float* dbeta, void* dy, int F,
                          int H, int W, int C, int c_real, int grid_w, int grid_h, int device, void* stream) {
    XCP_REQUIRE(C % 8 == 0 && mode >= 0 && mode <= 3 && c_real > 0 && c_real <= C, "xcp_bn_bwd: bad args");
    XCP_REQUIRE(coef != nullptr && (presums != nullptr || workspace != nullptr), "xcp_bn_bwd: workspace");
    XCP_CUDA(cudaSetDevice(device));
    BnBwdSrc s{mode, (const __nv_bfloat16*)G, (const uint8_t*)idx, dfeat, scale, shift, F, H, W, C};
    const long long n8 = (long long)F * H * W * (C / 8);
    const double count = (double)F * H * W;
    int nparts = 1;
    const float* sums = presums;
    const bool pool = (mode == SRC_POOL) && grid_w <= 0;
    const long long nq8 = (long long)F * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);      // 2x2 quads x channel groups (POOL mode)
    if (presums == nullptr) {
        long long g = ((pool ? nq8 : n8) + 255) / 256;
        nparts = (int)(g < xcp_bnbwd_num_parts() ? g : xcp_bnbwd_num_parts());
        if ((long long)nparts * 256 < C / 8) nparts = (C / 8 + 255) / 256;      // at least one thread per channel group
        if (pool) bnbwd_pool_kernel<false><<<nparts, 256, 2 * C * sizeof(float), ST>>>((const uint4*)y, s, workspace, nullptr, nullptr, nullptr, nullptr, nq8);
        else bnbwd_reduce_kernel<<<nparts, 256, 2 * C * sizeof(float), ST>>>((const uint4*)y, s, workspace, n8);
        XCP_CUDA(cudaGetLastError());
        sums = workspace;
    }
    // complete sums + a plain apply pass: the finalize arithmetic runs inside the apply kernel (no finalize launch)
    static int nofuse_env = -1;                                    // A/B hook: XCP_BN_FIN_SEPARATE=1 keeps the separate finalize launch
    if (nofuse_env < 0) { const char* e = getenv("XCP_BN_FIN_SEPARATE"); nofuse_env = e ? atoi(e) : 0; }
    const bool fused_fin = presums != nullptr && dy != nullptr && !pool && !nofuse_env;
    BnBwdFin fin{fused_fin ? presums : nullptr, gamma, mean, rstd, dgamma, dbeta, count, training, c_real};
    if (!fused_fin) {
        bnbwd_finalize_kernel<<<(C + 31) / 32, 256, 0, ST>>>(sums, nparts, C, c_real, count, gamma, mean, rstd, training, coef, coef + C,
                                                               coef + 2 * C, dgamma, dbeta);
        XCP_CUDA(cudaGetLastError());
    }
    if (dy != nullptr) {
        // grid-stride with a channel-group preserving stride: the grid must hold at least one thread per channel group
        const long long work = pool ? nq8 : (n8 + BNBWD_U - 1) / BNBWD_U;
        long long ga_ = (work + 255) / 256;
        int ga = (int)(ga_ < 2LL * num_sms() ? (ga_ > 0 ? ga_ : 1) : 2LL * num_sms());      // persistent: 2 resident CTAs per SM
        if ((long long)ga * 256 < C / 8) ga = (C / 8 + 255) / 256;
        if (pool) bnbwd_pool_kernel<true><<<ga, 256, 0, ST>>>((const uint4*)y, s, nullptr, coef, coef + C, coef + 2 * C, (uint4*)dy, nq8);
        else bnbwd_apply_kernel<<<ga, 256, 0, ST>>>((const uint4*)y, s, coef, coef + C, coef + 2 * C, (uint4*)dy, n8, grid_w, grid_h, fin);
    }
    return check_cuda(cudaGetLastError(), "bn_bwd launch");
}

// partials [nparts][2][C] -> sums [2][C]
__global__ void bnbwd_fold_kernel(const float* __restrict__ partials, int nparts, int C, float* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * C) return;
    double a = 0.0;
    for (int p = 0; p < nparts; ++p) a += (double)partials[(long long)p * 2 * C + i];
    sums[i] = (float)a;
}

// Pass 1 of the BatchNorm backward on its own: sums[0][c] = sum G, sums[1][c] = sum G*y over an [n_pix, C] pair of tensors.
// Used for the max-pool blocks (Xception.py:86): with dz routed to the arg-max pixels, sum dz = sum G and
// sum dz*y = sum_windows G * y[arg-max], so the pass runs over the pooled-resolution pair (y_max from xcp_pool_add_fwd, G)
// -- a quarter of the bytes of walking y with the routing logic -- and its result goes to xcp_bn_bwd as `presums`.
extern "C" int xcp_bn_bwd_sums(const void* y, const void* G, float* workspace, float* sums, long long n_pix, int C, int device,
                               void* stream) {
    XCP_REQUIRE(C % 8 == 0 && n_pix > 0 && y != nullptr && G != nullptr && workspace != nullptr && sums != nullptr, "xcp_bn_bwd_sums: bad args");
    XCP_CUDA(cudaSetDevice(device));
    BnBwdSrc s{SRC_DIRECT, (const __nv_bfloat16*)G, nullptr, nullptr, nullptr, nullptr, 1, 1, 1, C};
    const long long n8 = n_pix * (C / 8);
    long long g = (n8 + 255) / 256;
    int nparts = (int)(g < xcp_bnbwd_num_parts() ? g : xcp_bnbwd_num_parts());
    if ((long long)nparts * 256 < C / 8) nparts = (C / 8 + 255) / 256;
    bnbwd_reduce_kernel<<<nparts, 256, 2 * C * sizeof(float), ST>>>((const uint4*)y, s, workspace, n8);
    XCP_CUDA(cudaGetLastError());
    bnbwd_fold_kernel<<<(2 * C + 255) / 256, 256, 0, ST>>>(workspace, nparts, C, sums);
    return check_cuda(cudaGetLastError(), "bn_bwd_sums launch");
}

extern "C" int xcp_nchw_to_nhwc(const float* x, void* out, int F, int C, int Cp, int HW, int device, void* stream) {
    XCP_REQUIRE(Cp >= C, "xcp_nchw_to_nhwc: channel pitch < channels");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((HW + 31) / 32, (Cp + 31) / 32, F), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, ST>>>(x, (__nv_bfloat16*)out, C, Cp, HW);
    return check_cuda(cudaGetLastError(), "nchw_to_nhwc launch");
}

extern "C" int xcp_nhwc_to_nchw(const void* x, float* out, int F, int C, int Cp, int HW, int device, void* stream) {
    XCP_REQUIRE(Cp >= C, "xcp_nhwc_to_nchw: channel pitch < channels");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((HW + 31) / 32, (C + 31) / 32, F), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, ST>>>((const __nv_bfloat16*)x, out, C, Cp, HW);
    return check_cuda(cudaGetLastError(), "nhwc_to_nchw launch");
}

extern "C" int xcp_pack_weight(const float* w, void* out, void* out_t, int R, int Cc, int Rp, int Cp, int device, void* stream) {
    XCP_REQUIRE(Rp >= R && Cp >= Cc, "xcp_pack_weight: padded dims smaller than the matrix");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((Cp + 31) / 32, (Rp + 31) / 32), block(32, 8);
    pack_weight_kernel<<<grid, block, 0, ST>>>(w, (__nv_bfloat16*)out, (__nv_bfloat16*)out_t, R, Cc, Rp, Cp);
    return check_cuda(cudaGetLastError(), "pack_weight launch");
}

extern "C" int xcp_pack_weight_scaled(const float* w, const float* row_scale, void* out, int R, int Cc, int Rp, int Cp, int device,
                                      void* stream) {
    XCP_REQUIRE(Rp >= R && Cp >= Cc && row_scale != nullptr, "xcp_pack_weight_scaled: bad arguments");
    XCP_CUDA(cudaSetDevice(device));
    dim3 grid((Cp + 31) / 32, (Rp + 31) / 32), block(32, 8);
    pack_weight_kernel<<<grid, block, 0, ST>>>(w, (__nv_bfloat16*)out, nullptr, R, Cc, Rp, Cp, row_scale);
    return check_cuda(cudaGetLastError(), "pack_weight_scaled launch");
}

extern "C" int xcp_pack_multi(const void* table, int n_tensors, int n_tiles, int device, void* stream) {
    XCP_REQUIRE(table != nullptr && n_tensors > 0 && n_tiles > 0, "xcp_pack_multi: empty table");
    XCP_CUDA(cudaSetDevice(device));
    pack_multi_kernel<<<n_tiles, 256, 0, ST>>>((const PackTensor*)table, n_tensors);
    return check_cuda(cudaGetLastError(), "pack_multi launch");
}

extern "C" int xcp_pack_dw(const float* w, float* w9, int C, int Cp, int device, void* stream) {
    XCP_REQUIRE(Cp >= C, "xcp_pack_dw: channel pitch < channels");
    XCP_CUDA(cudaSetDevice(device));
    pack_dw_kernel<<<(9 * Cp + 255) / 256, 256, 0, ST>>>(w, w9, C, Cp);
    return check_cuda(cudaGetLastError(), "pack_dw launch");
}

extern "C" int xcp_unpack_dw_grad(const float* g9, float* gw, int C, int accumulate, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    unpack_dw_grad_kernel<<<(9 * C + 255) / 256, 256, 0, ST>>>(g9, gw, C, accumulate);
    return check_cuda(cudaGetLastError(), "unpack_dw_grad launch");
}

extern "C" int xcp_pack_conv3x3(const float* w, void* wk, void* wk_t, int O, int I, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    pack_conv3x3_kernel<<<(O * I * 9 + 255) / 256, 256, 0, ST>>>(w, (__nv_bfloat16*)wk, (__nv_bfloat16*)wk_t, O, I);
    return check_cuda(cudaGetLastError(), "pack_conv3x3 launch");
}

extern "C" int xcp_unpack_conv3x3_grad(const float* gk, float* gw, int O, int I, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    unpack_conv3x3_grad_kernel<<<(O * I * 9 + 255) / 256, 256, 0, ST>>>(gk, gw, O, I);
    return check_cuda(cudaGetLastError(), "unpack_conv3x3_grad launch");
}

extern "C" int xcp_bilinear_up(const float* x, float* out, long long planes, int n, int S, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    bilinear_up_kernel<<<ew_grid(planes * S * S, 256), 256, 0, ST>>>(x, out, planes, n, S);
    return check_cuda(cudaGetLastError(), "bilinear_up launch");
}
