// Depthwise 3x3 backward on tiny maps (S x S, S = 2 / 4 / 8: the middle and exit flow of the audio model's 64x64 patches,
// XceptionLSTMA.py:46 -> 4x4x728 x26 layers and 2x2x1024/1536).  Same fused contract as dw3x3_bwd_kernel (dw.cu): dgrad,
// 9-tap weight gradient, BN-affine / ReLU prologue and mask, identity-skip / stride-2-skip gradient adds, BatchNorm-backward sums.
//
// The TMA halo-tile kernel stages (S+2)^2 pixels per S^2 image in 6x6 boxes and ran these layers at 12 % of HBM (92 us per launch
// at 960 patches: 19 % of the audio model's training step, bench.py --config c4).  Here a thread owns one channel pair of `fpt`
// whole images: the S^2 gradient pixels of an image sit in registers (coalesced 4-byte loads: a warp reads 128 contiguous bytes
// per pixel), every operand byte is read once, dz is written once, and the weight-gradient / BatchNorm sums stay in registers
// across the thread's images before one atomic per (tap, channel).
#include "common.cuh"

namespace xcp {

template <int S, bool AFFINE, bool RELU, int ADDM>
__global__ void __launch_bounds__(256)
dw3x3_small_bwd_kernel(const __nv_bfloat162* __restrict__ dD, const __nv_bfloat162* __restrict__ xin, const float* __restrict__ w9,
                       const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat162* __restrict__ dz,
                       const __nv_bfloat162* __restrict__ add_full, const __nv_bfloat162* __restrict__ add_half,
                       float* __restrict__ dw, float* __restrict__ bnsum, int F, int C, int c_real, int add_pre, int fpt) {
    constexpr int SH = (S + 1) / 2;                                   // stride-2 skip gradient: [F, SH, SH, C]
    // block = 32 channel pairs (lanes: 128 contiguous bytes per pixel) x 8 frame groups (warps): the eight warps' weight-gradient
    // and BatchNorm sums of the same channels are added in shared memory before the atomics (an eighth of them reaches L2; one
    // atomic per thread and (tap, channel) made the first version atomic bound: 61 us per launch at 960 patches, now 31 -> see DESIGN)
    __shared__ float s_red[8][22][32];
    const int C2 = C >> 1;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int ctiles = (C2 + 31) / 32;
    const int c2 = (int)(blockIdx.x % ctiles) * 32 + lane;
    const long long f0 = ((long long)(blockIdx.x / ctiles) * 8 + wrp) * fpt;
    const bool live = c2 < C2 && f0 < F;
    const long long f1 = live ? (f0 + fpt < F ? f0 + fpt : F) : f0;
    float2 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = live ? *reinterpret_cast<const float2*>(w9 + (long long)k * C + 2 * c2) : make_float2(0.f, 0.f);
    float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
    if (AFFINE && live) { sc = *reinterpret_cast<const float2*>(scale + 2 * c2); sh = *reinterpret_cast<const float2*>(shift + 2 * c2); }
    float2 dwa[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) dwa[k] = make_float2(0.f, 0.f);
    float2 sdz = make_float2(0.f, 0.f), sdzy = make_float2(0.f, 0.f);
    const bool pre = (ADDM != 0) && add_pre != 0;

    for (long long f = f0; f < f1; ++f) {
        const long long base = f * (S * S) * C2 + c2;
        // the image's gradient pixels: fp32 pairs up to 4x4, packed bf16 pairs (unpacked at use) at 8x8 to stay inside the register file
        constexpr bool PACKED = S > 4;
        float2 g[PACKED ? 1 : S][PACKED ? 1 : S];
        uint32_t gp[PACKED ? S : 1][PACKED ? S : 1];
        float2 yv[S][S];
#pragma unroll
        for (int y = 0; y < S; ++y)
#pragma unroll
            for (int x = 0; x < S; ++x) {
                if (PACKED) gp[PACKED ? y : 0][PACKED ? x : 0] = *reinterpret_cast<const uint32_t*>(dD + base + (long long)(y * S + x) * C2);
                else g[PACKED ? 0 : y][PACKED ? 0 : x] = __bfloat1622float2(dD[base + (long long)(y * S + x) * C2]);
                yv[y][x] = __bfloat1622float2(xin[base + (long long)(y * S + x) * C2]);
            }
#pragma unroll
        for (int y = 0; y < S; ++y)
#pragma unroll
            for (int x = 0; x < S; ++x) {
                const float2 v = yv[y][x];
                float2 z = v;
                if (AFFINE) { z.x = fmaf(v.x, sc.x, sh.x); z.y = fmaf(v.y, sc.y, sh.y); }
                float2 av = z;
                if (RELU) { av.x = fmaxf(z.x, 0.f); av.y = fmaxf(z.y, 0.f); }
                float2 d = make_float2(0.f, 0.f);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int yy = y + 1 - kh, xx = x + 1 - kw;            // the output pixel that read (y, x) through tap (kh, kw)
                        if (yy >= 0 && yy < S && xx >= 0 && xx < S) {
                            float2 gg;
                            if (PACKED) { const uint32_t q = gp[PACKED ? yy : 0][PACKED ? xx : 0]; gg = make_float2(bf16_lo(q), bf16_hi(q)); }
                            else gg = g[PACKED ? 0 : yy][PACKED ? 0 : xx];
                            d.x = fmaf(wk[kh * 3 + kw].x, gg.x, d.x); d.y = fmaf(wk[kh * 3 + kw].y, gg.y, d.y);
                            dwa[kh * 3 + kw].x = fmaf(av.x, gg.x, dwa[kh * 3 + kw].x); dwa[kh * 3 + kw].y = fmaf(av.y, gg.y, dwa[kh * 3 + kw].y);
                        }
                    }
                if (RELU && !pre) { d.x = z.x > 0.f ? d.x : 0.f; d.y = z.y > 0.f ? d.y : 0.f; }
                if (ADDM & 1) {
                    const float2 a = __bfloat1622float2(add_full[base + (long long)(y * S + x) * C2]);
                    d.x += a.x; d.y += a.y;
                }
                if ((ADDM & 2) && (y & 1) == 0 && (x & 1) == 0) {
                    const float2 a = __bfloat1622float2(add_half[(f * (SH * SH) + (y >> 1) * SH + (x >> 1)) * C2 + c2]);
                    d.x += a.x; d.y += a.y;
                }
                if (RELU && pre) { d.x = z.x > 0.f ? d.x : 0.f; d.y = z.y > 0.f ? d.y : 0.f; }
                if (AFFINE) { sdz.x += d.x; sdz.y += d.y; sdzy.x = fmaf(d.x, v.x, sdzy.x); sdzy.y = fmaf(d.y, v.y, sdzy.y); }
                dz[base + (long long)(y * S + x) * C2] = __floats2bfloat162_rn(d.x, d.y);
            }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) { s_red[wrp][2 * k][lane] = dwa[k].x; s_red[wrp][2 * k + 1][lane] = dwa[k].y; }
    s_red[wrp][18][lane] = sdz.x; s_red[wrp][19][lane] = sdz.y; s_red[wrp][20][lane] = sdzy.x; s_red[wrp][21][lane] = sdzy.y;
    __syncthreads();
    const int c = 2 * c2;
    for (int j = wrp; j < (AFFINE ? 22 : 18); j += 8) {              // warp w adds rows w, w+8, ...: lanes = channel pairs
        const float v = ((s_red[0][j][lane] + s_red[1][j][lane]) + (s_red[2][j][lane] + s_red[3][j][lane])) +
                        ((s_red[4][j][lane] + s_red[5][j][lane]) + (s_red[6][j][lane] + s_red[7][j][lane]));
        if (c2 >= C2) continue;
        const int ch = c + (j & 1);
        if (j < 18) { if (ch < c_real) atomicAdd(dw + (long long)ch * 9 + (j >> 1), v); }
        else atomicAdd(bnsum + (long long)((j - 18) >> 1) * C + ch, v);
    }
}

template <int S>
static int launch_small_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift, int relu, void* dz,
                            const void* add_full, const void* add_half, float* dw, float* bnsum, int F, int C, int c_real,
                            cudaStream_t st) {
    const int C2 = C / 2;
    // enough threads to fill the machine, few enough images per thread to keep the atomics down: ~64 K threads
    long long fpt = ((long long)F * C2) / 65536;
    if (fpt < 1) fpt = 1;
    if (fpt > 16) fpt = 16;
    const long long groups = (F + fpt - 1) / fpt;
    const int grid = (int)(((groups + 7) / 8) * ((C2 + 31) / 32));
    const int addm = (add_full != nullptr ? 1 : 0) | (add_half != nullptr ? 2 : 0);
    const int variant = ((scale != nullptr) ? 8 : 0) | (relu ? 4 : 0) | addm;
    const int add_pre = relu == 2 ? 1 : 0;
#define SMALL_BWD(A, R, M)                                                                                                            \
    case ((A ? 8 : 0) | (R ? 4 : 0) | M):                                                                                             \
        dw3x3_small_bwd_kernel<S, A, R, M><<<grid, 256, 0, st>>>((const __nv_bfloat162*)dD, (const __nv_bfloat162*)xin, w9, scale, shift, \
            (__nv_bfloat162*)dz, (const __nv_bfloat162*)add_full, (const __nv_bfloat162*)add_half, dw, bnsum, F, C, c_real, add_pre, (int)fpt); \
        break;
    switch (variant) {
        SMALL_BWD(false, false, 0) SMALL_BWD(false, false, 1) SMALL_BWD(false, false, 2) SMALL_BWD(false, false, 3)
        SMALL_BWD(false, true, 0) SMALL_BWD(false, true, 1) SMALL_BWD(false, true, 2) SMALL_BWD(false, true, 3)
        SMALL_BWD(true, false, 0) SMALL_BWD(true, false, 1) SMALL_BWD(true, false, 2) SMALL_BWD(true, false, 3)
        SMALL_BWD(true, true, 0) SMALL_BWD(true, true, 1) SMALL_BWD(true, true, 2) SMALL_BWD(true, true, 3)
        default: break;
    }
#undef SMALL_BWD
    return check_cuda(cudaGetLastError(), "dw3x3_small_bwd launch");
}

// -> handled = 1 and the launch status when the shape is one of the register-resident ones; handled = 0 otherwise
int dw_small_try_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift, int relu, void* dz,
                     const void* add_full, const void* add_half, float* dw, float* bnsum, int F, int H, int W, int C, int c_real,
                     cudaStream_t st, int* handled) {
    *handled = 0;
    if (H != W || (H != 2 && H != 4 && H != 8)) return 0;
    const char* e = getenv("XCP_DW_NO_SMALL_BWD");                                    // A/B hook
    if (e != nullptr && e[0] == '1') return 0;
    *handled = 1;
    if (H == 8) return launch_small_bwd<8>(dD, xin, w9, scale, shift, relu, dz, add_full, add_half, dw, bnsum, F, C, c_real, st);
    if (H == 2) return launch_small_bwd<2>(dD, xin, w9, scale, shift, relu, dz, add_full, add_half, dw, bnsum, F, C, c_real, st);
    return launch_small_bwd<4>(dD, xin, w9, scale, shift, relu, dz, add_full, add_half, dw, bnsum, F, C, c_real, st);
}

}  // namespace xcp
