// Classifier head + loss as ONE kernel per direction (BASELINE north_star (3)).
//
// Reference: XceptionLSTMV.py:25-44 (fc_layers = 4 x [Linear -> ReLU -> Dropout(0.3)], fc_out, sigmoid), :66-70
// (lstm_out[:, -1, :] -> fc_layers -> sigmoid(fc_out)), nn.BCELoss on the sigmoid output (train_audio.py:20,39) and
// LabelSmoothingBCEWithLogitsLoss on the fc_out logits (train_au_patch.py:203-211).
//
// The five layers are dependent and tiny (B <= 32 rows, 13 MB of fp32 weights): what they cost as separate launches is
// launch gaps, not math.  Here the forward is one launch -- last-step select (a row stride / index into the LSTM output),
// the four hidden layers, fc_out, sigmoid, the loss and dL/dz -- and the backward is one launch (sigmoid', all five
// weight / bias gradients, the gradient wrt the selected LSTM rows).  Layers are separated by a grid-wide barrier on a
// global counter; the grids (<= 128 CTAs of 256 threads, a few KB of shared memory) are co-resident on the 148 SMs by
// construction, which the host wrapper checks with the occupancy API.  The dropout keep-mask is drawn in the kernel from a
// counter-based hash of (seed, launch counter, layer, row, neuron); the launch counter lives in device memory and is
// advanced by the kernel, so a CUDA-graph replay draws fresh masks.  The backward needs no mask: a kept, positive
// activation is > 0 and everything else is exactly 0.
#include "common.cuh"

namespace xcp {

constexpr int HEAD_MAXB = 32;
constexpr int HEAD_BCH = 8;      // batch rows per accumulator chunk
constexpr int HEAD_NS = 8;       // neurons per backward block
constexpr int HEAD_FWD_CTAS = 128;
constexpr int HEAD_BWD_CTAS = 128;

struct HeadFwdArgs {
    const float* x; long long row_stride; const long long* row_index;   // input row b = x + b*row_stride + row_index[b]*H
    const float* W[5]; const float* bias[5];
    const uint8_t* mask;            // optional keep masks [4][B][Wd] (tests); else drawn from rng when p_drop > 0
    unsigned long long* rng;        // {seed, launch counter} in device memory, or null
    float p_drop, drop_scale;
    float* acts;                    // [4][B][Wd] layer outputs (post ReLU + dropout), saved for the backward
    float* z; float* prob;          // [B]
    int loss_mode;                  // 0 none, 1 BCELoss on the probability, 2 BCE-with-logits on smoothed targets
    const float* y; float smoothing; float* loss; float* dz;
    unsigned* bar;                  // {arrivals, exits}, zero before the first launch, left zero by every launch
    int B, H, Wd;
};

struct HeadBwdArgs {
    const float* dsrc; const float* prob; const float* gscale;   // dz[b] = dsrc[b] * (prob ? p(1-p) : 1) * (gscale ? *gscale : 1)
    const float* x; long long row_stride; const long long* row_index;
    const float* acts; float drop_scale;
    const float* W[5]; float* dW[5]; float* db[5];
    float* dacts;                   // [4][B][Wd] scratch: gradients wrt the layer outputs
    float* dx; float* dx_base; long long dx_zero_n;   // gradient wrt the LSTM output: rows addressed like x; [dx_base, +dx_zero_n) zeroed first
    unsigned* bar;
    int B, H, Wd;
};

XCP_DEVINL void grid_sync(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v, polls = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            __nanosleep(64);
            if (++polls > (1u << 24)) __trap();     // a lost CTA would otherwise hang the device: fail loudly instead
        }
        __threadfence();
    }
    __syncthreads();
}

// last CTA out resets the counters (every CTA has passed the last barrier before it signs out)
XCP_DEVINL void grid_exit(unsigned* bar, unsigned long long* rng, unsigned long long next_counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned e = atomicAdd(bar + 1, 1u);
        if (e == gridDim.x - 1) {
            bar[0] = 0; bar[1] = 0;
            if (rng) rng[1] = next_counter;
            __threadfence();
        }
    }
}

XCP_DEVINL bool drop_keep(unsigned long long seed, unsigned long long counter, int layer, int b, int n, float p_drop) {
    unsigned long long v = seed + 0x9E3779B97F4A7C15ull * (counter + 1);
    v ^= ((unsigned long long)layer << 48) | ((unsigned long long)b << 32) | (unsigned long long)(unsigned)n;
    v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull;       // splitmix64 finaliser
    v ^= v >> 27; v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    const float u = (float)(v >> 40) * (1.0f / 16777216.0f);
    return u >= p_drop;
}

// ------------------------------------------------------------------------------------------------ forward
// Hidden layers: warp per output neuron.  After the grid barrier every CTA stages the layer's whole input [B][K] (<= 128 KB,
// written by other CTAs: ld.global.cg) into shared memory once; the weight row (K <= 1024 floats) sits in registers as
// 8 float4 per lane and is reused for all B rows; the NEXT layer's weight row is requested before the barrier so the HBM
// latency of the weights hides behind the barrier wait.
__global__ void __launch_bounds__(256)
head_mlp_fwd_kernel(const HeadFwdArgs p) {
    extern __shared__ float4 s_act[];            // [B][K/4]
    __shared__ float s_loss[HEAD_MAXB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + wib, nw = gridDim.x * 8;
    const int B = p.B, Wd = p.Wd;
    unsigned long long seed = 0, counter = 0;
    const bool draw = p.mask == nullptr && p.rng != nullptr && p.p_drop > 0.f;
    if (p.rng) { seed = p.rng[0]; counter = p.rng[1]; }

    float4 wv[8];
    auto load_row = [&](int l, int n) {
        const int K = l == 0 ? p.H : Wd;
        const float4* wrow = reinterpret_cast<const float4*>(p.W[l] + (long long)n * K);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 32 + lane;
            wv[j] = c < (K >> 2) ? __ldg(wrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    bool have = false;
    if (gw < Wd) { load_row(0, gw); have = true; }

    for (int l = 0; l < 4; ++l) {
        const int K = l == 0 ? p.H : Wd;
        const int K4 = K >> 2;
        const float* __restrict__ bias = p.bias[l];
        float* out = p.acts + (long long)l * B * Wd;
        const float* in = l == 0 ? nullptr : p.acts + (long long)(l - 1) * B * Wd;
        const uint8_t* mask = p.mask ? p.mask + (long long)l * B * Wd : nullptr;
        for (int i = threadIdx.x; i < B * K4; i += 256) {       // all copies in flight at once (cp.async.cg: L2, never L1)
            const int b = i / K4, c = i - b * K4;
            const float* arow = l == 0 ? p.x + (long long)b * p.row_stride + (p.row_index ? p.row_index[b] * p.H : 0) : in + (long long)b * Wd;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_act + i)), "l"(reinterpret_cast<const float4*>(arow) + c) : "memory");
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        for (int n = gw; n < Wd; n += nw) {
            if (!(have && n == gw)) load_row(l, n);
            have = false;
            const float bn = bias ? bias[n] : 0.f;
            for (int b0 = 0; b0 < B; b0 += HEAD_BCH) {
                float acc[HEAD_BCH];
#pragma unroll
                for (int b = 0; b < HEAD_BCH; ++b) {
                    acc[b] = 0.f;
                    if (b0 + b < B) {
                        const float4* ar = s_act + (b0 + b) * K4;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = j * 32 + lane;
                            if (c < K4) {
                                const float4 av = ar[c];
                                acc[b] = fmaf(wv[j].x, av.x, acc[b]); acc[b] = fmaf(wv[j].y, av.y, acc[b]);
                                acc[b] = fmaf(wv[j].z, av.z, acc[b]); acc[b] = fmaf(wv[j].w, av.w, acc[b]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < HEAD_BCH; ++b) {
                    if (b0 + b < B) {
                        float v = warp_sum(acc[b]);
                        if (lane == 0) {
                            v = fmaxf(v + bn, 0.f);
                            if (mask) v = mask[(long long)(b0 + b) * Wd + n] ? v * p.drop_scale : 0.f;
                            else if (draw) v = drop_keep(seed, counter, l, b0 + b, n, p.p_drop) ? v * p.drop_scale : 0.f;
                            out[(long long)(b0 + b) * Wd + n] = v;
                        }
                    }
                }
            }
        }
        if (l < 3 && gw < Wd) { load_row(l + 1, gw); have = true; }
        grid_sync(p.bar, (unsigned)(l + 1) * gridDim.x);
    }

    if (blockIdx.x == 0) {      // fc_out (one neuron) + sigmoid + loss: warp per row
        const float* a3 = p.acts + (long long)3 * B * Wd;
        const float4* w4 = reinterpret_cast<const float4*>(p.W[4]);
        const int K4 = Wd >> 2;
        for (int b = wib; b < B; b += 8) {
            const float4* ar = reinterpret_cast<const float4*>(a3 + (long long)b * Wd);
            float acc = 0.f;
            for (int c = lane; c < K4; c += 32) {
                const float4 av = __ldcg(ar + c), wv = __ldg(w4 + c);
                acc = fmaf(wv.x, av.x, acc); acc = fmaf(wv.y, av.y, acc); acc = fmaf(wv.z, av.z, acc); acc = fmaf(wv.w, av.w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                const float zz = acc + (p.bias[4] ? p.bias[4][0] : 0.f);
                const float pp = 1.f / (1.f + expf(-zz));
                p.z[b] = zz; p.prob[b] = pp;
                float term = 0.f, dzz = 0.f;
                if (p.loss_mode == 1) {          // nn.BCELoss: logs clamped at -100; torch's backward clamps p(1-p) at 1e-12
                    const float t = p.y[b];
                    term = -(t * fmaxf(logf(pp), -100.f) + (1.f - t) * fmaxf(logf(1.f - pp), -100.f));
                    dzz = (pp - t) / fmaxf((1.f - pp) * pp, 1e-12f) / (float)B * (pp * (1.f - pp));
                } else if (p.loss_mode == 2) {   // BCE-with-logits on y(1-s) + s/2
                    const float t = p.y[b] * (1.f - p.smoothing) + 0.5f * p.smoothing;
                    const float sp = log1pf(expf(-fabsf(zz)));
                    term = t * (fmaxf(-zz, 0.f) + sp) + (1.f - t) * (fmaxf(zz, 0.f) + sp);
                    dzz = (pp - t) / (float)B;
                }
                s_loss[b] = term;
                if (p.dz) p.dz[b] = dzz;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && p.loss_mode != 0 && p.loss) {
            float tot = 0.f;
            for (int b = 0; b < B; ++b) tot += s_loss[b];
            *p.loss = tot / (float)B;
        }
    }
    grid_exit(p.bar, p.rng, counter + 1);
}

// ------------------------------------------------------------------------------------------------ backward
// One hidden layer: delta = dout * (out > 0 ? drop_scale : 0); dW += delta^T in; db += sum_b delta; din += delta W (RED).
// Virtual block = HEAD_NS neurons x a 4*KT-wide column chunk; thread = (row group, float4 column): the rows of W and dW
// are streamed once with 16-byte accesses, the din partials stay in registers until one vector RED per (row, column).
template <int KT>
XCP_DEVINL void head_bwd_layer(const float* dout, const float* out_act, float drop_scale, const float* in, long long in_stride,
                               const long long* s_off, const float* __restrict__ W, float* dW, float* db, float* din, int B, int N,
                               int K, float (*s_d)[HEAD_BCH]) {
    constexpr int RG = 256 / KT;
    const int kc = threadIdx.x % KT, rg = threadIdx.x / KT;
    const int K4 = K >> 2;
    const int ny = (K4 + KT - 1) / KT;
    const int nvb = (N / HEAD_NS) * ny;
    for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        const int n0 = (vb / ny) * HEAD_NS, by = vb % ny;
        const int c4 = by * KT + kc;
        const bool col_ok = c4 < K4;
        for (int b0 = 0; b0 < B; b0 += HEAD_BCH) {
            __syncthreads();
            if (threadIdx.x < HEAD_NS * HEAD_BCH) {
                const int nn = threadIdx.x / HEAD_BCH, b = threadIdx.x % HEAD_BCH;
                float v = 0.f;
                if (b0 + b < B) {
                    const long long o = (long long)(b0 + b) * N + n0 + nn;
                    v = __ldcg(dout + o);
                    v = out_act[o] > 0.f ? v * drop_scale : 0.f;
                }
                s_d[nn][b] = v;
            }
            __syncthreads();
            if (db != nullptr && by == 0 && threadIdx.x < HEAD_NS) {
                float t = 0.f;
#pragma unroll
                for (int b = 0; b < HEAD_BCH; ++b) t += s_d[threadIdx.x][b];
                db[n0 + threadIdx.x] += t;
            }
            if (!col_ok) continue;
            float4 av[HEAD_BCH], acc[HEAD_BCH];
#pragma unroll
            for (int b = 0; b < HEAD_BCH; ++b) {
                acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                av[b] = acc[b];
                if (b0 + b < B)
                    av[b] = __ldcg(reinterpret_cast<const float4*>(in + (long long)(b0 + b) * in_stride + (s_off ? s_off[b0 + b] : 0)) + c4);
            }
#pragma unroll 4
            for (int nn = rg; nn < HEAD_NS; nn += RG) {
                const int n = n0 + nn;
                const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (long long)n * K) + c4);
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dW) g = reinterpret_cast<const float4*>(dW + (long long)n * K)[c4];
#pragma unroll
                for (int b = 0; b < HEAD_BCH; ++b) {
                    const float d = s_d[nn][b];
                    g.x = fmaf(d, av[b].x, g.x); g.y = fmaf(d, av[b].y, g.y); g.z = fmaf(d, av[b].z, g.z); g.w = fmaf(d, av[b].w, g.w);
                    acc[b].x = fmaf(d, wv.x, acc[b].x); acc[b].y = fmaf(d, wv.y, acc[b].y);
                    acc[b].z = fmaf(d, wv.z, acc[b].z); acc[b].w = fmaf(d, wv.w, acc[b].w);
                }
                if (dW) reinterpret_cast<float4*>(dW + (long long)n * K)[c4] = g;
            }
            if (din != nullptr) {
#pragma unroll
                for (int b = 0; b < HEAD_BCH; ++b) {
                    if (b0 + b < B) {
                        float* dp = din + (long long)(b0 + b) * in_stride + (s_off ? s_off[b0 + b] : 0) + c4 * 4;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(acc[b].x), "f"(acc[b].y),
                                     "f"(acc[b].z), "f"(acc[b].w) : "memory");
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
head_mlp_bwd_kernel(const HeadBwdArgs p) {
    __shared__ float s_d[HEAD_NS][HEAD_BCH];
    __shared__ float s_dz[HEAD_MAXB];
    __shared__ long long s_off[HEAD_MAXB];
    const int B = p.B, Wd = p.Wd, H = p.H;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;

    // ---- phase 0: dz, zero the RED targets, fc_out backward (dW4, db4, gradient wrt the last hidden layer written directly)
    if (threadIdx.x < HEAD_MAXB) {
        float v = 0.f;
        if (threadIdx.x < B) {
            v = p.dsrc[threadIdx.x];
            if (p.prob) { const float pp = p.prob[threadIdx.x]; v *= pp * (1.f - pp); }
            if (p.gscale) v *= p.gscale[0];
        }
        s_dz[threadIdx.x] = v;
        s_off[threadIdx.x] = (threadIdx.x < B && p.row_index) ? p.row_index[threadIdx.x] * (long long)H : 0;
    }
    __syncthreads();
    {
        float4* z0 = reinterpret_cast<float4*>(p.dacts);
        const long long n0 = (long long)3 * B * Wd / 4;
        for (long long i = gtid; i < n0; i += gthreads) z0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dx) {
            float4* z1 = reinterpret_cast<float4*>(p.dx_base);
            for (long long i = gtid; i < p.dx_zero_n / 4; i += gthreads) z1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* a3 = p.acts + (long long)3 * B * Wd;
        float* da3 = p.dacts + (long long)3 * B * Wd;
        for (long long k = gtid; k < Wd; k += gthreads) {
            const float wk = p.W[4][k];
            float g = 0.f;
            for (int b = 0; b < B; ++b) {
                g = fmaf(s_dz[b], a3[(long long)b * Wd + k], g);
                da3[(long long)b * Wd + k] = s_dz[b] * wk;
            }
            if (p.dW[4]) p.dW[4][k] += g;
        }
        if (gtid == 0 && p.db[4]) {
            float t = 0.f;
            for (int b = 0; b < B; ++b) t += s_dz[b];
            p.db[4][0] += t;
        }
    }
    grid_sync(p.bar, gridDim.x);

    // ---- hidden layers 3, 2, 1 (inputs and RED targets are the dense [B][Wd] buffers)
    for (int l = 3; l >= 1; --l) {
        head_bwd_layer<256>(p.dacts + (long long)l * B * Wd, p.acts + (long long)l * B * Wd, p.drop_scale,
                            p.acts + (long long)(l - 1) * B * Wd, Wd, nullptr, p.W[l], p.dW[l], p.db[l],
                            p.dacts + (long long)(l - 1) * B * Wd, B, Wd, Wd, s_d);
        grid_sync(p.bar, (unsigned)(5 - l) * gridDim.x);
    }
    // ---- layer 0: input = the selected LSTM rows, RED target = the same rows of dx
    const int K4 = H >> 2;
    if (K4 <= 32) head_bwd_layer<32>(p.dacts, p.acts, p.drop_scale, p.x, p.row_stride, s_off, p.W[0], p.dW[0], p.db[0], p.dx, B, Wd, H, s_d);
    else if (K4 <= 64) head_bwd_layer<64>(p.dacts, p.acts, p.drop_scale, p.x, p.row_stride, s_off, p.W[0], p.dW[0], p.db[0], p.dx, B, Wd, H, s_d);
    else if (K4 <= 128) head_bwd_layer<128>(p.dacts, p.acts, p.drop_scale, p.x, p.row_stride, s_off, p.W[0], p.dW[0], p.db[0], p.dx, B, Wd, H, s_d);
    else head_bwd_layer<256>(p.dacts, p.acts, p.drop_scale, p.x, p.row_stride, s_off, p.W[0], p.dW[0], p.db[0], p.dx, B, Wd, H, s_d);
    grid_exit(p.bar, nullptr, 0);
}

static int head_grid(const void* kernel, int want, size_t smem, int device, int* grid) {
    int per_sm = 0, sms = 0;
    if (smem > 48 * 1024) XCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem));
    XCP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int cap = per_sm * sms;
    XCP_REQUIRE(cap >= 8, "head kernel: the device cannot hold 8 co-resident CTAs (%d)", cap);
    *grid = want < cap ? want : cap;
    return 0;
}

}  // namespace xcp

using namespace xcp;
#define ST ((cudaStream_t)stream)

static int head_check(int B, int H, int Wd, const char* who) {
    XCP_REQUIRE(B > 0 && B <= HEAD_MAXB, "%s: 1..%d rows per call (got %d)", who, HEAD_MAXB, B);
    XCP_REQUIRE(H > 0 && H % 4 == 0 && H <= 1024, "%s: hidden size must be a multiple of 4 and <= 1024 (got %d)", who, H);
    XCP_REQUIRE(Wd > 0 && Wd % HEAD_NS == 0 && Wd <= 1024, "%s: layer width must be a multiple of %d and <= 1024 (got %d)", who, HEAD_NS, Wd);
    return 0;
}

extern "C" int xcp_head_mlp_fwd(const float* x, long long row_stride, const long long* row_index, const void* const* wb,
                                const void* mask, void* rng, float p_drop, float* acts, float* z, float* prob, int loss_mode,
                                const float* y, float smoothing, float* loss, float* dz, void* bar, int B, int H, int Wd, int device,
                                void* stream) {
    if (int rc = head_check(B, H, Wd, "xcp_head_mlp_fwd")) return rc;
    XCP_REQUIRE(x && wb && acts && z && prob && bar, "xcp_head_mlp_fwd: null pointer");
    XCP_REQUIRE(row_stride % 4 == 0 && ((uintptr_t)x & 15) == 0, "xcp_head_mlp_fwd: input rows must be 16-byte aligned");
    XCP_REQUIRE(loss_mode >= 0 && loss_mode <= 2 && (loss_mode == 0 || (y && loss)), "xcp_head_mlp_fwd: loss mode %d needs targets and a loss slot", loss_mode);
    XCP_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "xcp_head_mlp_fwd: dropout probability %f", p_drop);
    XCP_CUDA(cudaSetDevice(device));
    HeadFwdArgs a{};
    a.x = x; a.row_stride = row_stride; a.row_index = row_index;
    for (int l = 0; l < 5; ++l) { a.W[l] = (const float*)wb[2 * l]; a.bias[l] = (const float*)wb[2 * l + 1]; XCP_REQUIRE(a.W[l], "xcp_head_mlp_fwd: weight %d is null", l); }
    a.mask = (const uint8_t*)mask; a.rng = (unsigned long long*)rng;
    const bool dropping = p_drop > 0.f && (mask || rng);
    a.p_drop = dropping ? p_drop : 0.f; a.drop_scale = dropping ? 1.f / (1.f - p_drop) : 1.f;
    a.acts = acts; a.z = z; a.prob = prob; a.loss_mode = loss_mode; a.y = y; a.smoothing = smoothing; a.loss = loss; a.dz = dz;
    a.bar = (unsigned*)bar; a.B = B; a.H = H; a.Wd = Wd;
    int grid = 0;
    const size_t smem = (size_t)B * (H > Wd ? H : Wd) * sizeof(float);
    if (int rc = head_grid((const void*)head_mlp_fwd_kernel, HEAD_FWD_CTAS, smem, device, &grid)) return rc;
    head_mlp_fwd_kernel<<<grid, 256, smem, ST>>>(a);
    return check_cuda(cudaGetLastError(), "head_mlp_fwd launch");
}

extern "C" int xcp_head_mlp_bwd(const float* dsrc, const float* prob, const float* gscale, const float* x, long long row_stride,
                                const long long* row_index, const float* acts, float drop_scale, const void* const* wb,
                                void* const* dwb, float* dacts, float* dx, float* dx_base, long long dx_zero_n, void* bar, int B, int H, int Wd,
                                int device, void* stream) {
    if (int rc = head_check(B, H, Wd, "xcp_head_mlp_bwd")) return rc;
    XCP_REQUIRE(dsrc && x && acts && wb && dwb && dacts && bar, "xcp_head_mlp_bwd: null pointer");
    XCP_REQUIRE(row_stride % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dx_base & 15) == 0 && dx_zero_n % 4 == 0 && (!dx || dx_base),
                "xcp_head_mlp_bwd: rows must be 16-byte aligned");
    XCP_CUDA(cudaSetDevice(device));
    HeadBwdArgs a{};
    a.dsrc = dsrc; a.prob = prob; a.gscale = gscale; a.x = x; a.row_stride = row_stride; a.row_index = row_index;
    a.acts = acts; a.drop_scale = drop_scale;
    for (int l = 0; l < 5; ++l) {
        a.W[l] = (const float*)wb[2 * l]; a.dW[l] = (float*)dwb[2 * l]; a.db[l] = (float*)dwb[2 * l + 1];
        XCP_REQUIRE(a.W[l], "xcp_head_mlp_bwd: weight %d is null", l);
    }
    a.dacts = dacts; a.dx = dx; a.dx_base = dx_base; a.dx_zero_n = dx ? dx_zero_n : 0; a.bar = (unsigned*)bar; a.B = B; a.H = H; a.Wd = Wd;
    int grid = 0;
    if (int rc = head_grid((const void*)head_mlp_bwd_kernel, HEAD_BWD_CTAS, 0, device, &grid)) return rc;
    head_mlp_bwd_kernel<<<grid, 256, 0, ST>>>(a);
    return check_cuda(cudaGetLastError(), "head_mlp_bwd launch");
}
