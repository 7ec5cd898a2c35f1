// nn.LSTM recurrence for the wide hidden sizes (H = 256 / 512; XceptionLSTMA(hidden_dim=512), train_audio.py:15) on a
// thread-block CLUSTER, forward and BPTT.
//
// The single-CTA kernels in lstm_head.cu keep W_hh in shared memory, which only fits up to H = 128 (128 KB in bf16); at
// H = 512 (2 MB) every one of the T = 120 steps re-streamed W_hh from L2 into ONE SM per clip: 52 us per step, 77 % of the
// whole audio training step.  Here W_hh never moves after the prologue:
//   * a cluster of CS = H / 32 CTAs (16 at H = 512) serves up to 8 clips; CTA r owns hidden units [32 r, 32 r + 32), i.e. the
//     128 gate rows {g H + 32 r + u};
//   * its 128 x H slice of W_hh lives in REGISTERS as mma.m16n8k16 A fragments (H / 4 = 128 registers per thread, 8 warps x 16
//     gate rows); the batch is the n = 8 of the MMA, so one step is (H / 16) x 2 tensor-core instructions per warp -- h is fed
//     as a bf16 hi + lo pair, which keeps the recurrent input at ~fp32 accuracy (the weights are bf16, as before);
//   * the cell update runs on (unit, clip) threads with c in registers, and the new h slice is pushed into every peer's
//     shared memory with 16-byte DSMEM stores, followed by one cluster barrier per step (double-buffered h, no second barrier).
// The backward mirrors it: CTA r owns dh[:, 32 r : 32 r + 32] and the gate gradients of those units; dh_{t-1} = W_hh^T dg is
// computed as per-CTA partials over all H units (K = the CTA's own 128 gate rows) that are reduce-scattered over DSMEM.
// Reference: nn.LSTM(2048, H, 1, batch_first=True), XceptionLSTMV.py:18-23,67-68; XceptionLSTMA.py:14-19,56-57.
#include <cstdlib>
#include "common.cuh"

namespace xcp {
namespace {
constexpr int LSTM_CLUSTER_NA = -1000;
constexpr int LC_NB = 8;          // clips per cluster = n of mma.m16n8k16
constexpr int LC_THREADS = 256;   // 8 warps
constexpr int LC_U = 32;          // hidden units per CTA

XCP_DEVINL void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
XCP_DEVINL void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
XCP_DEVINL uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}
XCP_DEVINL float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------- forward
// wt = W_hh^T bf16 [H][4H] (the pack xcp_lstm_fwd already receives).  Outputs as lstm_fwd_kernel.
template <int H>
__global__ void __launch_bounds__(LC_THREADS, 1)
lstm_fwd_cluster_kernel(const float* __restrict__ xproj, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                        const __nv_bfloat16* __restrict__ wt, float* __restrict__ h_out, float* __restrict__ gates_out,
                        float* __restrict__ c_out, float* __restrict__ hn, float* __restrict__ cn, int B, int T) {
    constexpr int G = 4 * H, CS = H / LC_U, KS = H / 16, HP = H + 8, GP = 128 + 4;
    extern __shared__ __align__(16) uint8_t lc_smem[];
    __nv_bfloat16* s_h = reinterpret_cast<__nv_bfloat16*>(lc_smem);                         // [2 buf][hi, lo][NB][HP]
    float* s_g = reinterpret_cast<float*>(lc_smem + (size_t)2 * 2 * LC_NB * HP * 2);        // [NB][GP] gate pre-activations
    __nv_bfloat16* s_stage = reinterpret_cast<__nv_bfloat16*>(s_g + LC_NB * GP);            // [hi, lo][NB][32] new h slice
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t r = cluster_ctarank();
    const int cl = blockIdx.x / CS;

    // W_hh slice -> A fragments.  Local gate row c = 16 w + i  <->  gate g = w / 2, unit 32 r + 16 (w & 1) + i.
    uint32_t wf[KS][4];
    {
        const int j0 = (w >> 1) * H + (int)r * LC_U + (w & 1) * 16 + (lane >> 2);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = ks * 16 + (lane & 3) * 2;
            wf[ks][0] = pack_bf16(wt[(long long)k0 * G + j0], wt[(long long)(k0 + 1) * G + j0]);
            wf[ks][1] = pack_bf16(wt[(long long)k0 * G + j0 + 8], wt[(long long)(k0 + 1) * G + j0 + 8]);
            wf[ks][2] = pack_bf16(wt[(long long)(k0 + 8) * G + j0], wt[(long long)(k0 + 9) * G + j0]);
            wf[ks][3] = pack_bf16(wt[(long long)(k0 + 8) * G + j0 + 8], wt[(long long)(k0 + 9) * G + j0 + 8]);
        }
    }
    for (int i = tid; i < 2 * 2 * LC_NB * HP / 2; i += LC_THREADS) reinterpret_cast<uint32_t*>(s_h)[i] = 0u;
    // cell-update role: thread = (unit ul, clip b)
    const int ul = tid & 31, b = tid >> 5, u = (int)r * LC_U + ul;
    const int bg = cl * LC_NB + b;
    const bool valid = bg < B;
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = b_ih[g * H + u] + b_hh[g * H + u];
    float c_reg = 0.f;
    cluster_sync_all();                                   // every CTA's h buffers are zeroed before anybody writes into them

    float xn[4] = {0.f, 0.f, 0.f, 0.f};                  // input projection of the NEXT step (loads stay in flight over a step)
    if (valid) {
        const float* xr = xproj + ((long long)bg * T) * G + u;
#pragma unroll
        for (int g = 0; g < 4; ++g) xn[g] = xr[g * H];
    }
    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        float xp[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) xp[g] = xn[g];
        if (valid && t + 1 < T) {
            const float* xr = xproj + ((long long)bg * T + t + 1) * G + u;
#pragma unroll
            for (int g = 0; g < 4; ++g) xn[g] = xr[g * H];
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const __nv_bfloat16* hh = s_h + ((size_t)(buf * 2 + 0) * LC_NB + (lane >> 2)) * HP + (lane & 3) * 2;
            const __nv_bfloat16* hl = s_h + ((size_t)(buf * 2 + 1) * LC_NB + (lane >> 2)) * HP + (lane & 3) * 2;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint32_t h0 = *reinterpret_cast<const uint32_t*>(hh + ks * 16), h1 = *reinterpret_cast<const uint32_t*>(hh + ks * 16 + 8);
                const uint32_t l0 = *reinterpret_cast<const uint32_t*>(hl + ks * 16), l1 = *reinterpret_cast<const uint32_t*>(hl + ks * 16 + 8);
                mma_16816(acc, wf[ks], h0, h1);
                mma_16816(acc, wf[ks], l0, l1);
            }
        }
        {
            const int c = 16 * w + (lane >> 2), n = (lane & 3) * 2;
            s_g[n * GP + c] = acc[0]; s_g[(n + 1) * GP + c] = acc[1];
            s_g[n * GP + c + 8] = acc[2]; s_g[(n + 1) * GP + c + 8] = acc[3];
        }
        __syncthreads();
        {
            const float ig = sigm(s_g[b * GP + ul] + xp[0] + bias[0]);
            const float fg = sigm(s_g[b * GP + 32 + ul] + xp[1] + bias[1]);
            const float gg = tanhf(s_g[b * GP + 64 + ul] + xp[2] + bias[2]);
            const float og = sigm(s_g[b * GP + 96 + ul] + xp[3] + bias[3]);
            const float c = fg * c_reg + ig * gg;
            c_reg = c;
            const float h = valid ? og * tanhf(c) : 0.f;
            if (valid) {
                const long long o = (long long)bg * T + t;
                gates_out[o * G + u] = ig; gates_out[o * G + H + u] = fg; gates_out[o * G + 2 * H + u] = gg; gates_out[o * G + 3 * H + u] = og;
                c_out[o * H + u] = c;
                h_out[o * H + u] = h;
                if (t == T - 1) { hn[(long long)bg * H + u] = h; cn[(long long)bg * H + u] = c; }
            }
            const __nv_bfloat16 hi = __float2bfloat16(h);
            s_stage[b * 32 + ul] = hi;
            s_stage[(LC_NB + b) * 32 + ul] = __float2bfloat16(h - __bfloat162float(hi));
        }
        __syncthreads();
        // push the new slice (2 x 8 x 32 bf16 = 64 chunks of 16 bytes) into every CTA's next-step buffer
        const int nxt = buf ^ 1;
        for (int i = tid; i < 64 * CS; i += LC_THREADS) {
            const int peer = i >> 6, q = i & 63, a = q >> 5, nb = (q & 31) >> 2, part = q & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(s_stage + (a * LC_NB + nb) * 32 + part * 8);
            const uint32_t dst = smem_u32(s_h + ((size_t)(nxt * 2 + a) * LC_NB + nb) * HP + (int)r * LC_U + part * 8);
            st_cluster_v4(mapa_cluster(dst, (uint32_t)peer), v);
        }
        cluster_sync_all();
    }
}

// ---------------------------------------------------------------------------------------------------------- backward
// wb = W_hh bf16 [4H][H] (the pack xcp_lstm_bwd already receives).  Outputs as lstm_bwd_kernel.
// dh_{t-1}[k] = sum_j W_hh[j][k] dg_t[j].  CTA r owns the gate gradients of its own 32 units (128 rows j), so it multiplies
// them -- straight from its own shared memory -- with its [H x 128] slice of W_hh^T (registers) into a PARTIAL dh over all H
// units, and the partials are reduce-scattered: the 32-unit segment that CTA p owns goes into p's receive buffer (fp32,
// 1 KB per sender), and after the cluster barrier every CTA adds its CS incoming partials.  Per step a CTA pushes CS x 1 KB
// over DSMEM -- the same volume as the forward's h exchange, and a quarter of broadcasting the 4H gate gradients.
template <int H>
__global__ void __launch_bounds__(LC_THREADS, 1)
lstm_bwd_cluster_kernel(const float* __restrict__ dout, const float* __restrict__ dhn, const float* __restrict__ dcn,
                        const float* __restrict__ gates, const float* __restrict__ cst, const float* __restrict__ hst,
                        const __nv_bfloat16* __restrict__ wb, __nv_bfloat16* __restrict__ dgates, __nv_bfloat16* __restrict__ hprev,
                        float* __restrict__ dbias_ih, float* __restrict__ dbias_hh, int B, int T) {
    constexpr int G = 4 * H, CS = H / LC_U, MTW = H / 16 / 8, JP = 128 + 8, HP = H + 4;
    extern __shared__ __align__(16) uint8_t lc_smem[];
    float* s_recv = reinterpret_cast<float*>(lc_smem);                                      // [2 buf][CS senders][NB][32]
    float* s_out = s_recv + 2 * CS * LC_NB * 32;                                            // [NB][HP] partial dh over all units
    __nv_bfloat16* s_dg = reinterpret_cast<__nv_bfloat16*>(s_out + LC_NB * HP);            // [hi, lo][NB][JP] own gate gradients
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t r = cluster_ctarank();
    const int cl = blockIdx.x / CS;

    // A[k][jl] = W_hh[j(jl)][k]: rows = ALL hidden units (warp w: m16 tiles w*MTW ..), K = the CTA's 128 gate rows
    // (local gate row jl = g*32 + ul  <->  j = g*H + 32 r + ul)
    uint32_t wf[MTW][8][4];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
        const int k0 = (w * MTW + mt) * 16 + (lane >> 2);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int jl = ks * 16 + (lane & 3) * 2;                                       // jl, jl+1 stay inside one gate (even)
            const long long j0 = (long long)((jl >> 5) * H + (int)r * LC_U + (jl & 31));
            const long long j8 = (long long)(((jl + 8) >> 5) * H + (int)r * LC_U + ((jl + 8) & 31));
            wf[mt][ks][0] = pack_bf16(wb[j0 * H + k0], wb[(j0 + 1) * H + k0]);
            wf[mt][ks][1] = pack_bf16(wb[j0 * H + k0 + 8], wb[(j0 + 1) * H + k0 + 8]);
            wf[mt][ks][2] = pack_bf16(wb[j8 * H + k0], wb[(j8 + 1) * H + k0]);
            wf[mt][ks][3] = pack_bf16(wb[j8 * H + k0 + 8], wb[(j8 + 1) * H + k0 + 8]);
        }
    }
    const int ul = tid & 31, b = tid >> 5, u = (int)r * LC_U + ul;
    const int bg = cl * LC_NB + b;
    const bool valid = bg < B;
    float dh_rec = (valid && dhn) ? dhn[(long long)bg * H + u] : 0.f;
    float dc_reg = (valid && dcn) ? dcn[(long long)bg * H + u] : 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    // saved state of the step being processed; the loads for step t-1 are issued while step t's MMAs + exchange run
    float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, c = 0.f, cprev = 0.f, hp = 0.f, dov = 0.f;
    if (valid) {
        const long long o = (long long)bg * T + T - 1;
        ig = gates[o * G + u]; fg = gates[o * G + H + u]; gg = gates[o * G + 2 * H + u]; og = gates[o * G + 3 * H + u];
        c = cst[o * H + u];
        if (T > 1) { cprev = cst[(o - 1) * H + u]; hp = hst[(o - 1) * H + u]; }
        if (dout) dov = dout[o * H + u];
    }
    cluster_sync_all();

    for (int t = T - 1; t >= 0; --t) {
        const int buf = t & 1;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        float n_ig = 0.f, n_fg = 0.f, n_gg = 0.f, n_og = 0.f, n_cprev = 0.f, n_hp = 0.f, n_dov = 0.f;
        if (valid) {
            const long long o = (long long)bg * T + t;
            if (t > 0) {
                n_ig = gates[(o - 1) * G + u]; n_fg = gates[(o - 1) * G + H + u];
                n_gg = gates[(o - 1) * G + 2 * H + u]; n_og = gates[(o - 1) * G + 3 * H + u];
                if (t > 1) { n_cprev = cst[(o - 2) * H + u]; n_hp = hst[(o - 2) * H + u]; }
                if (dout) n_dov = dout[(o - 1) * H + u];
            }
            const float tc = tanhf(c);
            const float dh = dh_rec + dov;
            const float dc = dc_reg + dh * og * (1.f - tc * tc);
            d[0] = dc * gg * ig * (1.f - ig);
            d[1] = dc * cprev * fg * (1.f - fg);
            d[2] = dc * ig * (1.f - gg * gg);
            d[3] = dh * tc * og * (1.f - og);
            dc_reg = dc * fg;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                dgates[o * G + g * H + u] = __float2bfloat16(d[g]);
                db[g] += d[g];
            }
            hprev[o * H + u] = __float2bfloat16(hp);
        }
        if (t == 0) break;                                 // dh_{-1} is not needed (uniform across the cluster)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const __nv_bfloat16 hi = __float2bfloat16(d[g]);
            s_dg[b * JP + g * 32 + ul] = hi;
            s_dg[(LC_NB + b) * JP + g * 32 + ul] = __float2bfloat16(d[g] - __bfloat162float(hi));
        }
        __syncthreads();
        {
            float acc[MTW][4];
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) { acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f; }
            const __nv_bfloat16* gh = s_dg + (lane >> 2) * JP + (lane & 3) * 2;
            const __nv_bfloat16* gl = gh + LC_NB * JP;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t h0 = *reinterpret_cast<const uint32_t*>(gh + ks * 16), h1 = *reinterpret_cast<const uint32_t*>(gh + ks * 16 + 8);
                const uint32_t l0 = *reinterpret_cast<const uint32_t*>(gl + ks * 16), l1 = *reinterpret_cast<const uint32_t*>(gl + ks * 16 + 8);
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) {
                    mma_16816(acc[mt], wf[mt][ks], h0, h1);
                    mma_16816(acc[mt], wf[mt][ks], l0, l1);
                }
            }
            const int n = (lane & 3) * 2;
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                const int k = (w * MTW + mt) * 16 + (lane >> 2);
                s_out[n * HP + k] = acc[mt][0]; s_out[(n + 1) * HP + k] = acc[mt][1];
                s_out[n * HP + k + 8] = acc[mt][2]; s_out[(n + 1) * HP + k + 8] = acc[mt][3];
            }
        }
        __syncthreads();
        // reduce-scatter: the segment of units owned by CTA p (8 clips x 32 floats = 64 chunks of 16 bytes) -> p's slot r
        for (int i = tid; i < 64 * CS; i += LC_THREADS) {
            const int peer = i >> 6, q = i & 63, nb = q >> 3, part = q & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(s_out + nb * HP + peer * LC_U + part * 4);
            const uint32_t dst = smem_u32(s_recv + (((size_t)buf * CS + r) * LC_NB + nb) * 32 + part * 4);
            st_cluster_v4(mapa_cluster(dst, (uint32_t)peer), v);
        }
        cluster_sync_all();
        {
            const float* rp = s_recv + (size_t)buf * CS * LC_NB * 32 + b * 32 + ul;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int sdx = 0; sdx < CS; sdx += 2) { a0 += rp[sdx * LC_NB * 32]; a1 += rp[(sdx + 1) * LC_NB * 32]; }
            dh_rec = a0 + a1;
        }
        ig = n_ig; fg = n_fg; gg = n_gg; og = n_og; c = cprev; cprev = n_cprev; hp = n_hp; dov = n_dov;
    }
    if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (dbias_ih) atomicAdd(&dbias_ih[g * H + u], db[g]);
            if (dbias_hh) atomicAdd(&dbias_hh[g * H + u], db[g]);
        }
    }
    cluster_sync_all();                                    // nobody leaves while a peer may still be writing into its smem
}

// One-time per (kernel, device): opt in to the shared-memory size / non-portable cluster size and ask the driver whether a
// cluster of `cs` CTAs can be co-scheduled.  Cached, because none of this may run while a stream is being captured into a
// CUDA graph (the first call always happens in an eager warm-up step).
template <typename K>
int launch_cluster(K kern, int* state /* [64], per device: 0 unknown, 1 usable, -1 not usable */, int cs, int B, size_t smem,
                   cudaStream_t st, void** args, const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return LSTM_CLUSTER_NA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * ((B + LC_NB - 1) / LC_NB)));
    cfg.blockDim = dim3(LC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (state[dev] == 0) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return LSTM_CLUSTER_NA; }
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && cs > 8) e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int nclusters = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&nclusters, (const void*)kern, &cfg);
        if (e != cudaSuccess) cudaGetLastError();
        state[dev] = (e == cudaSuccess && nclusters >= 1) ? 1 : -1;
    }
    if (state[dev] < 0) return LSTM_CLUSTER_NA;
    return check_cuda(cudaLaunchKernelExC(&cfg, (const void*)kern, args), what);
}

bool cluster_disabled() {
    const char* e = getenv("XCP_LSTM_NO_CLUSTER");
    return e && e[0] == '1';
}
}  // namespace

// Returns 0 = launched, LSTM_CLUSTER_NA = not applicable (caller falls back to the single-CTA kernel), else a CUDA error.
int lstm_fwd_cluster(const float* xproj, const float* b_ih, const float* b_hh, const void* w_hh_t, float* h_out, float* gates,
                     float* cstate, float* hn, float* cn, int B, int T, int H, cudaStream_t st) {
    if (cluster_disabled() || (H != 128 && H != 256 && H != 512)) return LSTM_CLUSTER_NA;
    void* args[] = {&xproj, &b_ih, &b_hh, &w_hh_t, &h_out, &gates, &cstate, &hn, &cn, &B, &T};
    const size_t smem = (size_t)2 * 2 * LC_NB * (H + 8) * 2 + (size_t)LC_NB * 132 * 4 + 2 * LC_NB * 32 * 2;
    static int st512[64] = {0}, st256[64] = {0}, st128[64] = {0};
    if (H == 128) return launch_cluster(lstm_fwd_cluster_kernel<128>, st128, 4, B, smem, st, args, "lstm_fwd_cluster launch");
    if (H == 512) return launch_cluster(lstm_fwd_cluster_kernel<512>, st512, 16, B, smem, st, args, "lstm_fwd_cluster launch");
    return launch_cluster(lstm_fwd_cluster_kernel<256>, st256, 8, B, smem, st, args, "lstm_fwd_cluster launch");
}

int lstm_bwd_cluster(const float* dout, const float* dhn, const float* dcn, const float* gates, const float* cstate,
                     const float* hstate, const void* w_hh, void* dgates, void* hprev, float* dbias_ih, float* dbias_hh, int B,
                     int T, int H, cudaStream_t st) {
    if (cluster_disabled() || (H != 128 && H != 256 && H != 512)) return LSTM_CLUSTER_NA;
    void* args[] = {&dout, &dhn, &dcn, &gates, &cstate, &hstate, &w_hh, &dgates, &hprev, &dbias_ih, &dbias_hh, &B, &T};
    const size_t smem = (size_t)2 * (H / LC_U) * LC_NB * 32 * 4 + (size_t)LC_NB * (H + 4) * 4 + (size_t)2 * LC_NB * (128 + 8) * 2;
    static int st512[64] = {0}, st256[64] = {0}, st128[64] = {0};
    if (H == 128) return launch_cluster(lstm_bwd_cluster_kernel<128>, st128, 4, B, smem, st, args, "lstm_bwd_cluster launch");
    if (H == 512) return launch_cluster(lstm_bwd_cluster_kernel<512>, st512, 16, B, smem, st, args, "lstm_bwd_cluster launch");
    return launch_cluster(lstm_bwd_cluster_kernel<256>, st256, 8, B, smem, st, args, "lstm_bwd_cluster launch");
}
}  // namespace xcp
