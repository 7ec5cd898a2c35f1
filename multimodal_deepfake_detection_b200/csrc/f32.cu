// fp32 VALIDATION path of the Xception forward (north_star: "fp32 logits within 1e-4 relative of the reference").
//
// The production path computes in bf16 on tcgen05 with fp32 accumulation; these kernels restate the same forward in
// plain fp32 FMA arithmetic on NHWC fp32 activations so that the layout / indexing / BatchNorm bookkeeping of the
// plan can be checked against the reference at fp32 tolerance.  They are correctness instruments: straightforward
// tiling, no tensor cores, forward only, read the fp32 master parameters in torch's own layouts (no packing).
// Host side: fp32_plan.py.  Reference arithmetic: Models/Xception.py:44-47 (separable conv), :89-99 (block),
// :167-199 (network).
#include "common.cuh"

namespace {
using namespace xcp;
#define ST ((cudaStream_t)stream)

inline unsigned grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = 148LL * 32;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// dense 3x3 convolution, padding 0: x fp32 (NCHW when x_nchw, else NHWC), w [Co][Ci][3][3], out NHWC [F,Ho,Wo,Co]
__global__ void __launch_bounds__(256)
f32_conv3x3_kernel(const float* __restrict__ x, int x_nchw, const float* __restrict__ w, float* __restrict__ out, int F, int H, int W,
                   int Ci, int Co, int stride, int Ho, int Wo) {
    const long long total = (long long)F * Ho * Wo * Co;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int co = (int)(i % Co);
        long long p = i / Co;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int f = (int)(p / Ho);
        float acc = 0.f;
        for (int ci = 0; ci < Ci; ++ci)
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) {
                    const int hi = ho * stride + kh, wi = wo * stride + kw;
                    const float xv = x_nchw ? x[(((long long)f * Ci + ci) * H + hi) * W + wi] : x[(((long long)f * H + hi) * W + wi) * Ci + ci];
                    acc = fmaf(xv, w[((co * Ci + ci) * 3 + kh) * 3 + kw], acc);
                }
        out[i] = acc;
    }
}

// depthwise 3x3, stride 1, padding 1: x/out NHWC [F,H,W,C], w [C][1][3][3]
__global__ void __launch_bounds__(256)
f32_dw3x3_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, int F, int H, int W, int C) {
    const long long total = (long long)F * H * W * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % W); p /= W;
        const int ho = (int)(p % H);
        const int f = (int)(p / H);
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hi = ho + kh - 1, wi = wo + kw - 1;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) acc = fmaf(x[(((long long)f * H + hi) * W + wi) * C + c], w[c * 9 + kh * 3 + kw], acc);
            }
        out[i] = acc;
    }
}

// out[M,N] = a[M,K] . w[N,K]^T (+ bias[N]): 64x64 tile, 16-deep k slices, 4x4 outputs per thread
__global__ void __launch_bounds__(256)
f32_gemm_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int M,
                int N, int K) {
    __shared__ float sa[16][64 + 1], sw[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long m0 = (long long)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, k = e & 15;
            sa[k][r] = (m0 + r < M && k0 + k < K) ? a[(m0 + r) * K + k0 + k] : 0.f;
            sw[k][r] = (n0 + r < N && k0 + k < K) ? w[(long long)(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; wv[i] = sw[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long m = m0 + ty * 4 + i;
            const int n = n0 + tx * 4 + j;
            if (m < M && n < N) out[m * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
        }
}

// per-channel (sum, sum of squares) partials over row chunks: partials [nparts][2][C] (the layout xcp_bn_finalize reads)
__global__ void __launch_bounds__(256)
f32_bn_stats_kernel(const float* __restrict__ y, float* __restrict__ partials, long long M, int C, long long rows_per_part) {
    const long long r0 = blockIdx.x * rows_per_part;
    long long r1 = r0 + rows_per_part;
    if (r1 > M) r1 = M;
    for (int c = threadIdx.x; c < C; c += 256) {
        double s1 = 0.0, s2 = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double v = (double)y[r * C + c];
            s1 += v; s2 += v * v;
        }
        partials[((long long)blockIdx.x * 2 + 0) * C + c] = (float)s1;
        partials[((long long)blockIdx.x * 2 + 1) * C + c] = (float)s2;
    }
}

// out = relu?(scale[c] * y + shift[c]); scale == NULL: plain (optional) ReLU
__global__ void __launch_bounds__(256)
f32_affine_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                  float* __restrict__ out, long long n, int C) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float v = y[i];
        if (scale) { const int c = (int)(i % C); v = fmaf(v, scale[c], shift[c]); }
        out[i] = relu ? fmaxf(v, 0.f) : v;
    }
}

// MaxPool2d(3, 2, 1) of y [F,H,W,C] (+ skip [F,Ho,Wo,C] when given)
__global__ void __launch_bounds__(256)
f32_pool_add_kernel(const float* __restrict__ y, const float* __restrict__ skip, float* __restrict__ out, int F, int H, int W, int C,
                    int Ho, int Wo) {
    const long long total = (long long)F * Ho * Wo * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int f = (int)(p / Ho);
        float m = -INFINITY;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int hi = 2 * ho + kh - 1, wi = 2 * wo + kw - 1;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) m = fmaxf(m, y[(((long long)f * H + hi) * W + wi) * C + c]);
            }
        out[i] = m + (skip ? skip[i] : 0.f);
    }
}

__global__ void __launch_bounds__(256)
f32_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = a[i] + b[i];
}

// out[f,ho,wo,:] = x[f, ho*stride, wo*stride, :]  (input sampling of the strided 1x1 skip convolution)
__global__ void __launch_bounds__(256)
f32_gather_kernel(const float* __restrict__ x, float* __restrict__ out, int F, int H, int W, int C, int stride, int Ho, int Wo) {
    const long long total = (long long)F * Ho * Wo * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int f = (int)(p / Ho);
        out[i] = x[(((long long)f * H + ho * stride) * W + wo * stride) * C + c];
    }
}

// adaptive_avg_pool2d((1,1)): out[f,c] = mean over HW of x[f,:,c]
__global__ void __launch_bounds__(256)
f32_gap_kernel(const float* __restrict__ x, float* __restrict__ out, int F, int HW, int C) {
    const long long total = (long long)F * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        const int f = (int)(i / C);
        double s = 0.0;
        for (int p = 0; p < HW; ++p) s += (double)x[((long long)f * HW + p) * C + c];
        out[i] = (float)(s / (double)HW);
    }
}

// nn.LSTM(I, H, 1, batch_first) recurrence in fp32 (gate order i, f, g, o; zero initial state): one CTA per sequence,
// a warp per gate row (lanes over k, coalesced reads of the fp32 W_hh [4H][H]), then the cell update.
__global__ void __launch_bounds__(256)
f32_lstm_kernel(const float* __restrict__ xproj, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                const float* __restrict__ w_hh, float* __restrict__ h_out, float* __restrict__ hn, float* __restrict__ cn, int T, int H) {
    extern __shared__ float lstm_sm[];
    float* h = lstm_sm;
    float* c = lstm_sm + H;
    float* g = lstm_sm + 2 * H;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < H; j += 256) { h[j] = 0.f; c[j] = 0.f; }
    for (int t = 0; t < T; ++t) {
        __syncthreads();
        const float* xp = xproj + ((long long)b * T + t) * 4 * H;
        for (int row = warp; row < 4 * H; row += 8) {
            float acc = 0.f;
            for (int k = lane; k < H; k += 32) acc = fmaf(w_hh[(long long)row * H + k], h[k], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) g[row] = (xp[row] + b_ih[row]) + (acc + b_hh[row]);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < H; j += 256) {
            const float ig = 1.f / (1.f + expf(-g[j])), fg = 1.f / (1.f + expf(-g[H + j]));
            const float gg = tanhf(g[2 * H + j]), og = 1.f / (1.f + expf(-g[3 * H + j]));
            const float cc = fg * c[j] + ig * gg;
            const float hh = og * tanhf(cc);
            c[j] = cc; h[j] = hh;
            h_out[((long long)b * T + t) * H + j] = hh;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += 256) { hn[(long long)b * H + j] = h[j]; cn[(long long)b * H + j] = c[j]; }
}
}  // namespace

extern "C" int xcp_f32_conv3x3(const float* x, int x_nchw, const float* w, float* out, int F, int H, int W, int Ci, int Co, int stride,
                               int device, void* stream) {
    XCP_REQUIRE(F > 0 && H >= 3 && W >= 3 && Ci > 0 && Co > 0 && stride > 0, "xcp_f32_conv3x3: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1;
    f32_conv3x3_kernel<<<grid_for((long long)F * Ho * Wo * Co, 256), 256, 0, ST>>>(x, x_nchw, w, out, F, H, W, Ci, Co, stride, Ho, Wo);
    return check_cuda(cudaGetLastError(), "f32_conv3x3 launch");
}

extern "C" int xcp_f32_dw3x3(const float* x, const float* w, float* out, int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0, "xcp_f32_dw3x3: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_dw3x3_kernel<<<grid_for((long long)F * H * W * C, 256), 256, 0, ST>>>(x, w, out, F, H, W, C);
    return check_cuda(cudaGetLastError(), "f32_dw3x3 launch");
}

extern "C" int xcp_f32_gemm(const float* a, const float* w, const float* bias, float* out, long long M, int N, int K, int device,
                            void* stream) {
    XCP_REQUIRE(M > 0 && N > 0 && K > 0, "xcp_f32_gemm: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    XCP_REQUIRE((M + 63) / 64 <= 65535, "xcp_f32_gemm: M too large for the validation kernel (%lld rows)", M);
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
    f32_gemm_kernel<<<grid, 256, 0, ST>>>(a, w, bias, out, (int)M, N, K);
    return check_cuda(cudaGetLastError(), "f32_gemm launch");
}

extern "C" int xcp_f32_bn_stats_parts(long long M) {
    long long parts = (M + 255) / 256;
    return (int)(parts > 2048 ? 2048 : (parts < 1 ? 1 : parts));
}

extern "C" int xcp_f32_bn_stats(const float* y, float* partials, long long M, int C, int device, void* stream) {
    XCP_REQUIRE(M > 0 && C > 0, "xcp_f32_bn_stats: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const int nparts = xcp_f32_bn_stats_parts(M);
    const long long rpp = (M + nparts - 1) / nparts;
    f32_bn_stats_kernel<<<nparts, 256, 0, ST>>>(y, partials, M, C, rpp);
    return check_cuda(cudaGetLastError(), "f32_bn_stats launch");
}

extern "C" int xcp_f32_affine(const float* y, const float* scale, const float* shift, int relu, float* out, long long n, int C,
                              int device, void* stream) {
    XCP_REQUIRE(n > 0 && C > 0 && n % C == 0, "xcp_f32_affine: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_affine_kernel<<<grid_for(n, 256), 256, 0, ST>>>(y, scale, shift, relu, out, n, C);
    return check_cuda(cudaGetLastError(), "f32_affine launch");
}

extern "C" int xcp_f32_pool_add(const float* y, const float* skip, float* out, int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0, "xcp_f32_pool_add: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    f32_pool_add_kernel<<<grid_for((long long)F * Ho * Wo * C, 256), 256, 0, ST>>>(y, skip, out, F, H, W, C, Ho, Wo);
    return check_cuda(cudaGetLastError(), "f32_pool_add launch");
}

extern "C" int xcp_f32_add(const float* a, const float* b, float* out, long long n, int device, void* stream) {
    XCP_REQUIRE(n > 0, "xcp_f32_add: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_add_kernel<<<grid_for(n, 256), 256, 0, ST>>>(a, b, out, n);
    return check_cuda(cudaGetLastError(), "f32_add launch");
}

extern "C" int xcp_f32_gather(const float* x, float* out, int F, int H, int W, int C, int stride, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && stride > 0, "xcp_f32_gather: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    f32_gather_kernel<<<grid_for((long long)F * Ho * Wo * C, 256), 256, 0, ST>>>(x, out, F, H, W, C, stride, Ho, Wo);
    return check_cuda(cudaGetLastError(), "f32_gather launch");
}

extern "C" int xcp_f32_gap(const float* x, float* out, int F, int HW, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && HW > 0 && C > 0, "xcp_f32_gap: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    f32_gap_kernel<<<grid_for((long long)F * C, 256), 256, 0, ST>>>(x, out, F, HW, C);
    return check_cuda(cudaGetLastError(), "f32_gap launch");
}

extern "C" int xcp_f32_lstm_fwd(const float* xproj, const float* b_ih, const float* b_hh, const float* w_hh, float* h_out, float* hn,
                                float* cn, int B, int T, int H, int device, void* stream) {
    XCP_REQUIRE(B > 0 && T > 0 && H > 0 && 6LL * H * 4 <= 200 * 1024, "xcp_f32_lstm_fwd: bad shape (B=%d T=%d H=%d)", B, T, H);
    XCP_CUDA(cudaSetDevice(device));
    const int smem = 6 * H * 4;
    if (smem > 48 * 1024) XCP_CUDA(cudaFuncSetAttribute(f32_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    f32_lstm_kernel<<<B, 256, smem, ST>>>(xproj, b_ih, b_hh, w_hh, h_out, hn, cn, T, H);
    return check_cuda(cudaGetLastError(), "f32_lstm launch");
}
