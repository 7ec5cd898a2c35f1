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
                    v = (out_act == nullptr || out_act[o] > 0.f) ? v * drop_scale : 0.f;     // out_act null: a linear output, plain scale
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

// ================================================================================================ fusion head (train_au_face)
// The fused region of train_au_face.py:659-674 as ONE launch per direction: token mean-pooling of both streams, concat,
// embed_head (Linear -> ReLU -> Dropout(0.2) -> Linear), ArcFace margin logits (:423-442), class-balanced focal loss
// (:445-458) + 0.2 * mse(v_pool, au_pool) + 0.1 * 0.5 * (temporal smoothness of both token streams) (:669-674).
// Forward: pooled, hidden, embedding, logits, the loss and the unit-loss gradients wrt the embedding and the ArcFace weight
// (they fall out of the same per-sample arithmetic).  Backward: everything scaled by the upstream scalar read from device
// memory: ArcFace weight gradient, both Linear layers, the pooling and the regularisers' token gradients.
struct FusionFwdArgs {
    const float* v; const float* a; int B, Tv, Ta, D;         // tokens [B,Tv,D], [B,Ta,D]
    const float* W0; const float* b0; const float* W3; const float* b3; int N0, N3;      // embed_head[0]: [N0, 2D], embed_head[3]: [N3, N0]
    const float* arc_w; const long long* labels; float s, m; int loss_mode; const float* class_w; float gamma;
    float la, lt;
    const uint8_t* mask; unsigned long long* rng; float p_drop, drop_scale;
    float* pooled; float* h; float* e; float* logits; float* loss; float* de; float* darc; float* rows;   // rows: [B][2] scratch, darc scratch [B][2][N3]
    unsigned* bar;
};

struct FusionBwdArgs {
    const float* gscale;
    const float* v; const float* a; int B, Tv, Ta, D;
    const float* W0; const float* W3; int N0, N3;
    const float* pooled; const float* h; const float* de; const float* darc_unit; float drop_scale; float la, lt;
    float* dW0; float* db0; float* dW3; float* db3; float* darc;      // accumulated into
    float* dh; float* dpooled;                                        // scratch [B,N0], [B,2D]
    float* dv; float* da;                                             // token gradients (written), nullable
    unsigned* bar;
};

// out[b][n] = act(in[b] . W[n] + bias[n]) (* dropout): warp per neuron, weight row in registers, the layer input staged in shared memory
XCP_DEVINL void fusion_fwd_layer(const float* in, const float* __restrict__ W, const float* __restrict__ bias, float* out, int B,
                                 int N, int K, bool relu, const uint8_t* mask, bool draw, unsigned long long seed,
                                 unsigned long long counter, float p_drop, float drop_scale, float4* s_act) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K4 = K >> 2;
    for (int i = threadIdx.x; i < B * K4; i += 256)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_act + i)), "l"(reinterpret_cast<const float4*>(in) + i) : "memory");
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int n = blockIdx.x * 8 + wib; n < N; n += gridDim.x * 8) {
        float4 wv[8];
        const float4* wrow = reinterpret_cast<const float4*>(W + (long long)n * K);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j * 32 + lane;
            wv[j] = c < K4 ? __ldg(wrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float bn = bias ? bias[n] : 0.f;
        for (int b = 0; b < B; ++b) {
            const float4* ar = s_act + b * K4;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = j * 32 + lane;
                if (c < K4) {
                    const float4 av = ar[c];
                    acc = fmaf(wv[j].x, av.x, acc); acc = fmaf(wv[j].y, av.y, acc); acc = fmaf(wv[j].z, av.z, acc); acc = fmaf(wv[j].w, av.w, acc);
                }
            }
            float val = warp_sum(acc);
            if (lane == 0) {
                val += bn;
                if (relu) val = fmaxf(val, 0.f);
                if (mask) val = mask[(long long)b * N + n] ? val * drop_scale : 0.f;
                else if (draw) val = drop_keep(seed, counter, 0, b, n, p_drop) ? val * drop_scale : 0.f;
                out[(long long)b * N + n] = val;
            }
        }
    }
}

// mean over T of one token stream -> pooled columns [off, off + D); returns this thread's share of the temporal-smoothness loss
XCP_DEVINL float fusion_pool_stream(const float* tok, int b, int T, int D, float* pooled_row, int off, float lt_w) {
    float local = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float* col = tok + (long long)b * T * D + d;
        float sum = 0.f, prev = 0.f;
        for (int t = 0; t < T; ++t) {
            const float x = col[(long long)t * D];
            sum += x;
            if (t > 0) { const float e = x - prev; local = fmaf(lt_w * e, e, local); }
            prev = x;
        }
        pooled_row[off + d] = sum / (float)T;
    }
    return local;
}

__global__ void __launch_bounds__(256)
fusion_head_fwd_kernel(const FusionFwdArgs p) {
    extern __shared__ float4 s_act[];
    __shared__ float s_red[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int B = p.B, D = p.D, K0 = 2 * p.D;
    unsigned long long seed = 0, counter = 0;
    const bool draw = p.mask == nullptr && p.rng != nullptr && p.p_drop > 0.f;
    if (p.rng) { seed = p.rng[0]; counter = p.rng[1]; }

    // ---- phase 0: pooled = [mean_t v | mean_t a]; per-sample regulariser terms (fixed summation order: deterministic loss)
    const float ltv = p.Tv > 1 ? p.lt * 0.5f / ((float)B * (p.Tv - 1) * D) : 0.f;
    const float lta = p.Ta > 1 ? p.lt * 0.5f / ((float)B * (p.Ta - 1) * D) : 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        float* prow = p.pooled + (long long)b * K0;
        float local = fusion_pool_stream(p.v, b, p.Tv, D, prow, 0, ltv) + fusion_pool_stream(p.a, b, p.Ta, D, prow, D, lta);
        __syncthreads();                                   // pooled row complete (written by this CTA)
        for (int d = threadIdx.x; d < D; d += blockDim.x) { const float diff = prow[d] - prow[D + d]; local = fmaf(p.la / ((float)B * D) * diff, diff, local); }
        local = warp_sum(local);
        if (lane == 0) s_red[wib] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int i = 0; i < 8; ++i) tot += s_red[i];
            p.rows[b * 2 + 1] = tot;                       // regulariser share of sample b
        }
        __syncthreads();
    }
    grid_sync(p.bar, gridDim.x);
    // ---- embed_head: Linear(2D, N0) -> ReLU -> Dropout ; Linear(N0, N3)
    fusion_fwd_layer(p.pooled, p.W0, p.b0, p.h, B, p.N0, K0, true, p.mask, draw, seed, counter, p.p_drop, p.drop_scale, s_act);
    grid_sync(p.bar, 2 * gridDim.x);
    fusion_fwd_layer(p.h, p.W3, p.b3, p.e, B, p.N3, p.N0, false, nullptr, false, 0, 0, 0.f, 1.f, s_act);
    grid_sync(p.bar, 3 * gridDim.x);

    // ---- ArcFace margin logits + CE / class-balanced focal loss: warp per sample (CTA 0), then fixed-order sums
    if (blockIdx.x == 0) {
        const int E = p.N3;
        const float* w = p.arc_w;
        for (int b = wib; b < B; b += 8) {
            const float* xb = p.e + (long long)b * E;
            float xx = 0.f, w0w0 = 0.f, w1w1 = 0.f, xw0 = 0.f, xw1 = 0.f;
            for (int k = lane; k < E; k += 32) {
                const float xv = __ldcg(xb + k), a = w[k], c = w[E + k];
                xx = fmaf(xv, xv, xx); w0w0 = fmaf(a, a, w0w0); w1w1 = fmaf(c, c, w1w1); xw0 = fmaf(xv, a, xw0); xw1 = fmaf(xv, c, xw1);
            }
            xx = warp_sum(xx); w0w0 = warp_sum(w0w0); w1w1 = warp_sum(w1w1); xw0 = warp_sum(xw0); xw1 = warp_sum(xw1);
            const float nx = fmaxf(sqrtf(xx), 1e-12f), nw0 = fmaxf(sqrtf(w0w0), 1e-12f), nw1 = fmaxf(sqrtf(w1w1), 1e-12f);
            const float cosv[2] = {xw0 / (nx * nw0), xw1 / (nx * nw1)};
            const long long y = p.labels ? p.labels[b] : -1;
            float lg[2] = {p.s * cosv[0], p.s * cosv[1]};
            float dt_dcos = 1.f;
            if (y >= 0) {
                const float cy = cosv[y];
                const float cl = fminf(fmaxf(cy, -1.f + 1e-7f), 1.f - 1e-7f);
                const float th = acosf(cl);
                lg[y] = p.s * cosf(th + p.m);
                const bool inside = (cy >= -1.f + 1e-7f) && (cy <= 1.f - 1e-7f);
                dt_dcos = inside ? sinf(th + p.m) / sqrtf(fmaxf(1.f - cl * cl, 1e-30f)) : 0.f;
            }
            if (lane == 0) { p.logits[b * 2] = lg[0]; p.logits[b * 2 + 1] = lg[1]; }
            if (y < 0) continue;
            const float mx = fmaxf(lg[0], lg[1]);
            const float e0 = __expf(lg[0] - mx), e1 = __expf(lg[1] - mx);
            const float den = e0 + e1;
            const float pr[2] = {e0 / den, e1 / den};
            const float ce_plain = -(lg[y] - mx - logf(den));
            float dl[2], row_loss;
            if (p.loss_mode == 0) {
                row_loss = ce_plain;
                dl[0] = pr[0] - (y == 0 ? 1.f : 0.f); dl[1] = pr[1] - (y == 1 ? 1.f : 0.f);
            } else {            // ce = w_y * ce_plain ; pt = exp(-ce) ; loss = (1-pt)^gamma * ce   (train_au_face.py:455-458)
                const float wy = p.class_w[y];
                const float ce = wy * ce_plain;
                const float pt = __expf(-ce);
                const float om = 1.f - pt;
                row_loss = powf(om, p.gamma) * ce;
                const float dloss_dce = p.gamma * powf(fmaxf(om, 1e-30f), p.gamma - 1.f) * pt * ce + powf(om, p.gamma);
                dl[0] = dloss_dce * wy * (pr[0] - (y == 0 ? 1.f : 0.f)); dl[1] = dloss_dce * wy * (pr[1] - (y == 1 ? 1.f : 0.f));
            }
            const float inv = 1.f / (float)B;
            if (lane == 0) p.rows[b * 2] = row_loss * inv;
            const float dc0 = p.s * dl[0] * inv * (y == 0 ? dt_dcos : 1.f), dc1 = p.s * dl[1] * inv * (y == 1 ? dt_dcos : 1.f);
            for (int k = lane; k < E; k += 32) {
                const float xv = __ldcg(xb + k) / nx, a = w[k] / nw0, c = w[E + k] / nw1;
                p.de[(long long)b * E + k] = (dc0 * (a - cosv[0] * xv) + dc1 * (c - cosv[1] * xv)) / nx;
                p.darc[((long long)b * 2 + 0) * E + k] = dc0 * (xv - cosv[0] * a) / nw0;      // per-sample rows: summed in order below
                p.darc[((long long)b * 2 + 1) * E + k] = dc1 * (xv - cosv[1] * c) / nw1;
            }
        }
        __syncthreads();
        if (p.labels != nullptr) {
            for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {      // unit-loss gradient of the ArcFace weight -> row 0 of the scratch
                float t = 0.f;                                           // column i of every row is touched by this thread only
                for (int b = 0; b < B; ++b) t += p.darc[(long long)b * 2 * E + i];
                p.darc[i] = t;
            }
            if (threadIdx.x == 0) {
                float tot = 0.f;
                for (int b = 0; b < B; ++b) tot += __ldcg(p.rows + b * 2);
                for (int b = 0; b < B; ++b) tot += __ldcg(p.rows + b * 2 + 1);       // written by other CTAs in phase 0
                *p.loss = tot;
            }
        }
    }
    grid_exit(p.bar, p.rng, counter + 1);
}

__global__ void __launch_bounds__(256)
fusion_head_bwd_kernel(const FusionBwdArgs p) {
    __shared__ float s_d[HEAD_NS][HEAD_BCH];
    const int B = p.B, D = p.D, K0 = 2 * p.D;
    const float g = p.gscale ? p.gscale[0] : 1.f;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long long)gridDim.x * blockDim.x;
    // ---- phase 0: zero the RED targets, ArcFace weight gradient, regulariser gradients of the tokens (written, not added)
    for (long long i = gtid; i < (long long)B * p.N0; i += gthreads) p.dh[i] = 0.f;
    for (long long i = gtid; i < (long long)B * K0; i += gthreads) p.dpooled[i] = 0.f;
    if (p.darc) for (long long i = gtid; i < 2LL * p.N3; i += gthreads) p.darc[i] += p.darc_unit[i] * g;
    for (int which = 0; which < 2; ++which) {
        const float* tok = which ? p.a : p.v;
        float* dt = which ? p.da : p.dv;
        const int T = which ? p.Ta : p.Tv;
        if (dt == nullptr) continue;
        const float sgn = which ? -1.f : 1.f;
        const float c_al = p.la * 2.f / ((float)B * D) / (float)T * g;
        const float c_tm = T > 1 ? p.lt / ((float)B * (T - 1) * D) * g : 0.f;
        const long long n = (long long)B * T * D;
        for (long long i = gtid; i < n; i += gthreads) {
            const int d = (int)(i % D);
            const long long bt = i / D;
            const int t = (int)(bt % T), b = (int)(bt / T);
            const float diff = p.pooled[(long long)b * K0 + d] - p.pooled[(long long)b * K0 + D + d];
            float gr = sgn * c_al * diff;
            float tv = 0.f;
            if (t + 1 < T) tv -= tok[i + D] - tok[i];
            if (t > 0) tv += tok[i] - tok[i - D];
            dt[i] = fmaf(c_tm, tv, gr);
        }
    }
    grid_sync(p.bar, gridDim.x);
    // ---- embed_head[3] (linear output): delta = de * g
    {
        const int K4 = p.N0 >> 2;
        if (K4 <= 32) head_bwd_layer<32>(p.de, nullptr, g, p.h, p.N0, nullptr, p.W3, p.dW3, p.db3, p.dh, B, p.N3, p.N0, s_d);
        else if (K4 <= 64) head_bwd_layer<64>(p.de, nullptr, g, p.h, p.N0, nullptr, p.W3, p.dW3, p.db3, p.dh, B, p.N3, p.N0, s_d);
        else if (K4 <= 128) head_bwd_layer<128>(p.de, nullptr, g, p.h, p.N0, nullptr, p.W3, p.dW3, p.db3, p.dh, B, p.N3, p.N0, s_d);
        else head_bwd_layer<256>(p.de, nullptr, g, p.h, p.N0, nullptr, p.W3, p.dW3, p.db3, p.dh, B, p.N3, p.N0, s_d);
    }
    grid_sync(p.bar, 2 * gridDim.x);
    // ---- embed_head[0] (ReLU + dropout): delta = dh * (h > 0 ? drop_scale : 0)
    {
        const int K4 = K0 >> 2;
        if (K4 <= 32) head_bwd_layer<32>(p.dh, p.h, p.drop_scale, p.pooled, K0, nullptr, p.W0, p.dW0, p.db0, p.dpooled, B, p.N0, K0, s_d);
        else if (K4 <= 64) head_bwd_layer<64>(p.dh, p.h, p.drop_scale, p.pooled, K0, nullptr, p.W0, p.dW0, p.db0, p.dpooled, B, p.N0, K0, s_d);
        else if (K4 <= 128) head_bwd_layer<128>(p.dh, p.h, p.drop_scale, p.pooled, K0, nullptr, p.W0, p.dW0, p.db0, p.dpooled, B, p.N0, K0, s_d);
        else head_bwd_layer<256>(p.dh, p.h, p.drop_scale, p.pooled, K0, nullptr, p.W0, p.dW0, p.db0, p.dpooled, B, p.N0, K0, s_d);
    }
    grid_sync(p.bar, 3 * gridDim.x);
    // ---- mean-pool backward: every token of a clip receives dpooled / T
    for (int which = 0; which < 2; ++which) {
        float* dt = which ? p.da : p.dv;
        const int T = which ? p.Ta : p.Tv;
        if (dt == nullptr) continue;
        const long long n = (long long)B * T * D;
        const float invT = 1.f / (float)T;
        for (long long i = gtid; i < n; i += gthreads) {
            const int d = (int)(i % D);
            const int b = (int)(i / ((long long)T * D));
            dt[i] = fmaf(__ldcg(p.dpooled + (long long)b * K0 + which * D + d), invT, dt[i]);
        }
    }
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

// ---- fusion head (train_au_face.py:659-674), one launch per direction
static int fusion_check(int B, int Tv, int Ta, int D, int N0, int N3, const char* who) {
    XCP_REQUIRE(B > 0 && B <= HEAD_MAXB, "%s: 1..%d clips per call (got %d)", who, HEAD_MAXB, B);
    XCP_REQUIRE(Tv > 0 && Ta > 0, "%s: empty token stream", who);
    XCP_REQUIRE(D > 0 && D % 2 == 0 && 2 * D <= 1024, "%s: token width must be even and <= 512 (got %d)", who, D);
    XCP_REQUIRE(N0 > 0 && N0 % HEAD_NS == 0 && N0 <= 1024 && N3 > 0 && N3 % HEAD_NS == 0 && N3 <= 1024,
                "%s: embed widths must be multiples of %d and <= 1024 (got %d, %d)", who, HEAD_NS, N0, N3);
    return 0;
}

extern "C" int xcp_fusion_head_fwd(const float* v, const float* a, int B, int Tv, int Ta, int D, const float* W0, const float* b0,
                                   const float* W3, const float* b3, int N0, int N3, const float* arc_w, const long long* labels,
                                   float s, float m, int loss_mode, const float* class_w, float gamma, float lambda_align,
                                   float lambda_temp, const void* mask, void* rng, float p_drop, float* pooled, float* h, float* e,
                                   float* logits, float* loss, float* de, float* darc_scratch, float* rows, void* bar, int device,
                                   void* stream) {
    if (int rc = fusion_check(B, Tv, Ta, D, N0, N3, "xcp_fusion_head_fwd")) return rc;
    XCP_REQUIRE(v && a && W0 && W3 && arc_w && pooled && h && e && logits && rows && bar, "xcp_fusion_head_fwd: null pointer");
    XCP_REQUIRE(labels == nullptr || (loss && de && darc_scratch), "xcp_fusion_head_fwd: labels need loss / gradient slots");
    XCP_REQUIRE(loss_mode == 0 || class_w != nullptr, "xcp_fusion_head_fwd: class-balanced focal loss needs class weights");
    XCP_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "xcp_fusion_head_fwd: dropout probability %f", p_drop);
    XCP_CUDA(cudaSetDevice(device));
    FusionFwdArgs q{};
    q.v = v; q.a = a; q.B = B; q.Tv = Tv; q.Ta = Ta; q.D = D; q.W0 = W0; q.b0 = b0; q.W3 = W3; q.b3 = b3; q.N0 = N0; q.N3 = N3;
    q.arc_w = arc_w; q.labels = labels; q.s = s; q.m = m; q.loss_mode = loss_mode; q.class_w = class_w; q.gamma = gamma;
    q.la = lambda_align; q.lt = lambda_temp;
    const bool dropping = p_drop > 0.f && (mask || rng);
    q.mask = (const uint8_t*)mask; q.rng = (unsigned long long*)rng; q.p_drop = dropping ? p_drop : 0.f; q.drop_scale = dropping ? 1.f / (1.f - p_drop) : 1.f;
    q.pooled = pooled; q.h = h; q.e = e; q.logits = logits; q.loss = loss; q.de = de; q.darc = darc_scratch; q.rows = rows;
    q.bar = (unsigned*)bar;
    const size_t smem = (size_t)B * (2 * D > N0 ? 2 * D : N0) * sizeof(float);
    int grid = 0;
    if (int rc = head_grid((const void*)fusion_head_fwd_kernel, 32, smem, device, &grid)) return rc;
    fusion_head_fwd_kernel<<<grid, 256, smem, ST>>>(q);
    return check_cuda(cudaGetLastError(), "fusion_head_fwd launch");
}

extern "C" int xcp_fusion_head_bwd(const float* gscale, const float* v, const float* a, int B, int Tv, int Ta, int D, const float* W0,
                                   const float* W3, int N0, int N3, const float* pooled, const float* h, const float* de,
                                   const float* darc_unit, float drop_scale, float lambda_align, float lambda_temp, float* dW0,
                                   float* db0, float* dW3, float* db3, float* darc, float* dh, float* dpooled, float* dv, float* da,
                                   void* bar, int device, void* stream) {
    if (int rc = fusion_check(B, Tv, Ta, D, N0, N3, "xcp_fusion_head_bwd")) return rc;
    XCP_REQUIRE(v && a && W0 && W3 && pooled && h && de && darc_unit && dh && dpooled && bar, "xcp_fusion_head_bwd: null pointer");
    XCP_CUDA(cudaSetDevice(device));
    FusionBwdArgs q{};
    q.gscale = gscale; q.v = v; q.a = a; q.B = B; q.Tv = Tv; q.Ta = Ta; q.D = D; q.W0 = W0; q.W3 = W3; q.N0 = N0; q.N3 = N3;
    q.pooled = pooled; q.h = h; q.de = de; q.darc_unit = darc_unit; q.drop_scale = drop_scale; q.la = lambda_align; q.lt = lambda_temp;
    q.dW0 = dW0; q.db0 = db0; q.dW3 = dW3; q.db3 = db3; q.darc = darc; q.dh = dh; q.dpooled = dpooled; q.dv = dv; q.da = da;
    q.bar = (unsigned*)bar;
    int grid = 0;
    if (int rc = head_grid((const void*)fusion_head_bwd_kernel, 32, 0, device, &grid)) return rc;
    fusion_head_bwd_kernel<<<grid, 256, 0, ST>>>(q);
    return check_cuda(cudaGetLastError(), "fusion_head_bwd launch");
}
