// Depthwise 3x3 (stride 1, pad 1) forward and backward on NHWC bf16 activations.
//
// Replaces nn.Conv2d(groups=C) in SeparableConv2d.conv1 (Xception.py:41,45) together with the ReLU that
// precedes it (Xception.py:61-76) and the previous layer's BatchNorm affine (Xception.py:67,73,78), which
// are applied on the fly to the staged tile ("consumer prologue" fusion).  HBM-bound: every input element
// is fetched once by TMA into a shared-memory halo tile (OOB zero fill = the conv's zero padding), every
// output element is written once with 16-byte stores.
//
// Tile = (TH x TW pixels) x 64 channels of one frame; a CTA is persistent over tiles with a 2-deep TMA ring.
// Thread = (4-channel group, column pair, row slice): it streams down its two columns keeping a 3-row window
// of partial sums in registers, so each staged vector is read from shared memory ~2x (not 9x).
#include "common.cuh"

namespace xcp {

struct DwGeom {
    int F, H, W, C;
    int TH, TW, n_h, n_w, c_tiles;   // tiling
    int RS, rows_per_slice;          // row slices per tile (thread utilisation for narrow tiles)
    long long num_tiles;
};

static DwGeom make_geom(int F, int H, int W, int C, int cols_per_cta, int max_tile_pixels, bool pair_cols) {
    DwGeom g{};
    g.F = F; g.H = H; g.W = W; g.C = C;
    g.n_w = (W + cols_per_cta - 1) / cols_per_cta;
    g.TW = (W + g.n_w - 1) / g.n_w;
    int th_max = max_tile_pixels / (g.TW + 2) - 2;
    if (th_max < 1) th_max = 1;
    g.n_h = (H + th_max - 1) / th_max;
    g.TH = (H + g.n_h - 1) / g.n_h;
    g.c_tiles = (C + 63) / 64;
    g.RS = pair_cols ? (cols_per_cta / 2) / ((g.TW + 1) / 2) : cols_per_cta / g.TW;
    if (g.RS < 1) g.RS = 1;
    if (g.RS > g.TH) g.RS = g.TH;
    g.rows_per_slice = (g.TH + g.RS - 1) / g.RS;
    g.num_tiles = (long long)F * g.n_h * g.n_w * g.c_tiles;
    return g;
}

struct DwFwdParams {
    DwGeom g;
    const float* w9;      // [9][C] tap-major depthwise weights
    const float* scale;   // [C] pending BN affine of the producer (AFFINE) or null
    const float* shift;
    int relu;
    __nv_bfloat16* out;   // [F,H,W,C]
};

template <bool AFFINE>
__global__ void __launch_bounds__(256, 2)
dw3x3_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const DwFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const DwGeom& g = p.g;
    const int halo_w = g.TW + 2, halo_h = g.TH + 2;
    const uint32_t stage_bytes = (uint32_t)halo_w * halo_h * 128u;
    __shared__ uint64_t full[2];

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
    }
    __syncthreads();

    auto issue = [&](long long tile, int s) {
        long long t = tile;
        const int ct = (int)(t % g.c_tiles); t /= g.c_tiles;
        const int tw = (int)(t % g.n_w); t /= g.n_w;
        const int th = (int)(t % g.n_h); t /= g.n_h;
        const int f = (int)t;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_4d(smem + s * stage_bytes, &tmX, &full[s], ct * 64, tw * g.TW - 1, th * g.TH - 1, f);
    };

    // thread = (4-channel group, column pair, row slice)
    const int cg = threadIdx.x & 15;
    const int rest = threadIdx.x >> 4;            // 0..15
    const int pairs = (g.TW + 1) >> 1;
    const int x0 = (rest % pairs) * 2;            // first of the two tile-local columns
    const int slice = rest / pairs;
    const bool thread_active = slice < g.RS;

    long long tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < g.num_tiles) issue(tile, 0);

    for (int it = 0; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const long long next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < g.num_tiles) issue(next, s ^ 1);

        long long t = tile;
        const int ct = (int)(t % g.c_tiles); t /= g.c_tiles;
        const int tw = (int)(t % g.n_w); t /= g.n_w;
        const int th = (int)(t % g.n_h); t /= g.n_h;
        const int f = (int)t;
        const int c0 = ct * 64 + cg * 4;
        const int gx0 = tw * g.TW + x0;
        const bool active = thread_active && c0 < g.C && gx0 < g.W && x0 < g.TW;
        const bool second = (x0 + 1 < g.TW) && (gx0 + 1 < g.W);

        float wgt[9][4];
        float sc[4], sh[4];
        if (active) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(p.w9 + (long long)k * g.C + c0);
                wgt[k][0] = a.x; wgt[k][1] = a.y; wgt[k][2] = a.z; wgt[k][3] = a.w;
            }
            if (AFFINE) {
                const float4 a = *reinterpret_cast<const float4*>(p.scale + c0);
                const float4 c = *reinterpret_cast<const float4*>(p.shift + c0);
                sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w;
                sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w;
            }
        }

        mbar_wait(&full[s], (it >> 1) & 1);

        if (active) {
            const uint8_t* tile_smem = smem + s * stage_bytes;
            const int r0 = slice * g.rows_per_slice;
            const int r1 = min(r0 + g.rows_per_slice, g.TH);
            // in-image validity of the four staged columns gx0-1 .. gx0+2 (AFFINE must keep the padding at 0)
            bool cv[4];
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) cv[dx] = (gx0 - 1 + dx >= 0) && (gx0 - 1 + dx < g.W);
            float a0[2][4], a1[2][4], a2[2][4];
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int j = 0; j < 4; ++j) { a0[o][j] = 0.f; a1[o][j] = 0.f; a2[o][j] = 0.f; }
            // input (tile-local) rows r0-1 .. r1 ; smem row index = local row + 1
            for (int ir = r0 - 1; ir <= r1; ++ir) {
                const int gh = th * g.TH + ir;
                const bool row_valid = gh >= 0 && gh < g.H;
                float v[4][4];
                const uint8_t* rp = tile_smem + ((ir + 1) * halo_w + x0) * 128 + cg * 8;
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) {
                    // the 4th column may lie outside the staged box for an odd-width tile's last pair
                    uint2 raw = make_uint2(0u, 0u);
                    if (dx < 3 || x0 + 3 < halo_w) raw = *reinterpret_cast<const uint2*>(rp + dx * 128);
                    v[dx][0] = bf16_lo(raw.x); v[dx][1] = bf16_hi(raw.x); v[dx][2] = bf16_lo(raw.y); v[dx][3] = bf16_hi(raw.y);
                    if (AFFINE) {
                        const bool ok = row_valid && cv[dx];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float z = fmaf(v[dx][j], sc[j], sh[j]);
                            if (p.relu) z = fmaxf(z, 0.f);
                            v[dx][j] = ok ? z : 0.f;
                        }
                    } else if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[dx][j] = fmaxf(v[dx][j], 0.f);
                    }
                }
                // this input row is tap row kh=2 of output ir-1, kh=1 of output ir, kh=0 of output ir+1
#pragma unroll
                for (int o = 0; o < 2; ++o) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        a0[o][j] = fmaf(wgt[6][j], v[o + 0][j], a0[o][j]);
                        a0[o][j] = fmaf(wgt[7][j], v[o + 1][j], a0[o][j]);
                        a0[o][j] = fmaf(wgt[8][j], v[o + 2][j], a0[o][j]);
                        a1[o][j] = fmaf(wgt[3][j], v[o + 0][j], a1[o][j]);
                        a1[o][j] = fmaf(wgt[4][j], v[o + 1][j], a1[o][j]);
                        a1[o][j] = fmaf(wgt[5][j], v[o + 2][j], a1[o][j]);
                        a2[o][j] = wgt[0][j] * v[o + 0][j];
                        a2[o][j] = fmaf(wgt[1][j], v[o + 1][j], a2[o][j]);
                        a2[o][j] = fmaf(wgt[2][j], v[o + 2][j], a2[o][j]);
                    }
                }
                const int orow = ir - 1;
                const int oh = th * g.TH + orow;
                if (orow >= r0 && oh < g.H) {
                    __nv_bfloat16* op = p.out + (((long long)f * g.H + oh) * g.W + gx0) * g.C + c0;
                    *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16(a0[0][0], a0[0][1]), pack_bf16(a0[0][2], a0[0][3]));
                    if (second)
                        *reinterpret_cast<uint2*>(op + g.C) =
                            make_uint2(pack_bf16(a0[1][0], a0[1][1]), pack_bf16(a0[1][2], a0[1][3]));
                }
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { a0[o][j] = a1[o][j]; a1[o][j] = a2[o][j]; }
            }
        }
        __syncthreads();   // everyone is done with stage s before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  For every pixel p and channel c (a = the activated DW input, dD = grad wrt the DW output):
//     da[p]        = sum_{kh,kw} w[kh][kw] * dD[p + (1-kh, 1-kw)]
//     dw[kh][kw]  += a[p] * dD[p + (1-kh, 1-kw)]
// so both need the same 3x3 neighbourhood of dD (TMA halo tile) and only the centre value of the input.
// The kernel also applies the ReLU mask / BN-affine chain rule and accumulates the per-channel sums the
// BatchNorm backward needs (sum dz, sum dz*y), and can add the residual-branch gradient(s).
struct DwBwdParams {
    DwGeom g;
    const float* w9;              // [9][C]
    const __nv_bfloat16* xin;     // DW forward input source (raw y if AFFINE else materialised x), [F,H,W,C]
    const float* scale;           // AFFINE only
    const float* shift;
    int relu;
    __nv_bfloat16* dz;            // out: grad wrt the pre-activation (z if AFFINE, x otherwise) [F,H,W,C]
    const __nv_bfloat16* add_full;   // optional: same-shape gradient to add (identity skip), after masking
    const __nv_bfloat16* add_half;   // optional: [F,ceil(H/2),ceil(W/2),C] gradient of the stride-2 skip gather
    float* partials;              // [gridDim.x][11][64]: 9 dw taps, sum dz, sum dz*y  (per CTA, fixed channel tile)
};

template <bool AFFINE>
__global__ void __launch_bounds__(256, 2)
dw3x3_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const DwBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 127u) & ~127u) - raw_addr);
    const DwGeom& g = p.g;
    const int halo_w = g.TW + 2, halo_h = g.TH + 2;
    const uint32_t stage_bytes = (uint32_t)halo_w * halo_h * 128u;
    __shared__ uint64_t full[2];
    __shared__ float s_red[11][64];

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmG);
    }
    for (int i = threadIdx.x; i < 11 * 64; i += blockDim.x) (&s_red[0][0])[i] = 0.f;
    __syncthreads();

    // A CTA owns one channel tile for its whole life so the weight/BN partial sums stay in registers.
    const int ct = blockIdx.x % g.c_tiles;
    const long long sp_tiles = (long long)g.F * g.n_h * g.n_w;
    const int sp_stride = gridDim.x / g.c_tiles;          // host guarantees gridDim.x % c_tiles == 0
    long long sp = blockIdx.x / g.c_tiles;

    auto issue = [&](long long sp_tile, int s) {
        long long t = sp_tile;
        const int tw = (int)(t % g.n_w); t /= g.n_w;
        const int th = (int)(t % g.n_h); t /= g.n_h;
        const int f = (int)t;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_4d(smem + s * stage_bytes, &tmG, &full[s], ct * 64, tw * g.TW - 1, th * g.TH - 1, f);
    };

    // thread = (4-channel group, column, row slice)
    const int cg = threadIdx.x & 15;
    const int rest = threadIdx.x >> 4;            // 0..15
    const int x = rest % g.TW;
    const int slice = rest / g.TW;
    const int c0 = ct * 64 + cg * 4;
    const bool thread_active = slice < g.RS && c0 < g.C;

    float wgt[9][4], dwa[9][4], sdz[4], sdzy[4], sc[4], sh[4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) { wgt[k][j] = 0.f; dwa[k][j] = 0.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { sdz[j] = 0.f; sdzy[j] = 0.f; sc[j] = 1.f; sh[j] = 0.f; }
    if (thread_active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(p.w9 + (long long)k * g.C + c0);
            wgt[k][0] = a.x; wgt[k][1] = a.y; wgt[k][2] = a.z; wgt[k][3] = a.w;
        }
        if (AFFINE) {
            const float4 a = *reinterpret_cast<const float4*>(p.scale + c0);
            const float4 b = *reinterpret_cast<const float4*>(p.shift + c0);
            sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w;
            sh[0] = b.x; sh[1] = b.y; sh[2] = b.z; sh[3] = b.w;
        }
    }

    if (threadIdx.x == 0 && sp < sp_tiles) issue(sp, 0);

    for (int it = 0; sp < sp_tiles; sp += sp_stride, ++it) {
        const int s = it & 1;
        const long long next = sp + sp_stride;
        if (threadIdx.x == 0 && next < sp_tiles) issue(next, s ^ 1);

        long long t = sp;
        const int tw = (int)(t % g.n_w); t /= g.n_w;
        const int th = (int)(t % g.n_h); t /= g.n_h;
        const int f = (int)t;
        const int gx = tw * g.TW + x;
        const bool active = thread_active && gx < g.W;

        mbar_wait(&full[s], (it >> 1) & 1);

        if (active) {
            const uint8_t* tile_smem = smem + s * stage_bytes;
            const int r0 = slice * g.rows_per_slice;
            const int r1 = min(r0 + g.rows_per_slice, g.TH);
            for (int r = r0; r < r1; ++r) {
                const int gh = th * g.TH + r;
                if (gh >= g.H) break;
                const long long pix = (((long long)f * g.H + gh) * g.W + gx) * g.C + c0;
                // centre input value -> activated a, mask
                const uint2 xr = *reinterpret_cast<const uint2*>(p.xin + pix);
                float yv[4] = {bf16_lo(xr.x), bf16_hi(xr.x), bf16_lo(xr.y), bf16_hi(xr.y)};
                float a[4];
                bool pos[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = AFFINE ? fmaf(yv[j], sc[j], sh[j]) : yv[j];
                    pos[j] = p.relu ? (z > 0.f) : true;
                    a[j] = pos[j] ? z : 0.f;
                }
                float da[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        // neighbour p + (1-kh, 1-kw): tile-local row r+1-kh -> smem row r+2-kh ; col x+2-kw
                        const uint2 gr = *reinterpret_cast<const uint2*>(
                            tile_smem + ((r + 2 - kh) * halo_w + (x + 2 - kw)) * 128 + cg * 8);
                        const float gv[4] = {bf16_lo(gr.x), bf16_hi(gr.x), bf16_lo(gr.y), bf16_hi(gr.y)};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            da[j] = fmaf(wgt[kh * 3 + kw][j], gv[j], da[j]);
                            dwa[kh * 3 + kw][j] = fmaf(a[j], gv[j], dwa[kh * 3 + kw][j]);
                        }
                    }
                }
                float dzv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) dzv[j] = pos[j] ? da[j] : 0.f;
                if (p.add_full != nullptr) {
                    const uint2 ar = *reinterpret_cast<const uint2*>(p.add_full + pix);
                    dzv[0] += bf16_lo(ar.x); dzv[1] += bf16_hi(ar.x); dzv[2] += bf16_lo(ar.y); dzv[3] += bf16_hi(ar.y);
                }
                if (p.add_half != nullptr && ((gh | gx) & 1) == 0) {
                    const int Ho = (g.H + 1) / 2, Wo = (g.W + 1) / 2;
                    const uint2 ar = *reinterpret_cast<const uint2*>(
                        p.add_half + (((long long)f * Ho + (gh >> 1)) * Wo + (gx >> 1)) * g.C + c0);
                    dzv[0] += bf16_lo(ar.x); dzv[1] += bf16_hi(ar.x); dzv[2] += bf16_lo(ar.y); dzv[3] += bf16_hi(ar.y);
                }
                if (AFFINE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { sdz[j] += dzv[j]; sdzy[j] = fmaf(dzv[j], yv[j], sdzy[j]); }
                }
                uint2 o;
                o.x = pack_bf16(dzv[0], dzv[1]);
                o.y = pack_bf16(dzv[2], dzv[3]);
                *reinterpret_cast<uint2*>(p.dz + pix) = o;
            }
        }
        __syncthreads();
    }

    // CTA reduction of the weight-gradient / BN sums (shared atomics: <= 16 contributors per address)
    if (thread_active) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&s_red[k][cg * 4 + j], dwa[k][j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&s_red[9][cg * 4 + j], sdz[j]); atomicAdd(&s_red[10][cg * 4 + j], sdzy[j]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 11 * 64; i += blockDim.x)
        p.partials[(long long)blockIdx.x * (11 * 64) + i] = (&s_red[0][0])[i];
}

// Reduce the per-CTA partials of dw3x3_bwd: dw9[9][C] (+=), and optional BN sums out[2][C] (=).
__global__ void dw_bwd_finalize_kernel(const float* partials, int grid, int c_tiles, int C, float* dw9, float* bnsum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int ct = c / 64, cl = c % 64;
    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.0;
    for (int b = ct; b < grid; b += c_tiles) {
        const float* pp = partials + (long long)b * (11 * 64) + cl;
#pragma unroll
        for (int k = 0; k < 11; ++k) acc[k] += (double)pp[k * 64];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) dw9[(long long)k * C + c] += (float)acc[k];
    if (bnsum != nullptr) { bnsum[c] = (float)acc[9]; bnsum[C + c] = (float)acc[10]; }
}

static int make_dw_tmap(CUtensorMap* m, const void* base, const DwGeom& g) {
    const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.F};
    const uint64_t strides[3] = {(uint64_t)g.C * 2, (uint64_t)g.W * g.C * 2, (uint64_t)g.H * g.W * g.C * 2};
    const uint32_t box[4] = {64, (uint32_t)(g.TW + 2), (uint32_t)(g.TH + 2), 1};
    return make_tmap_4d(m, base, dims, strides, box, 0);
}

}  // namespace xcp

using namespace xcp;

// out[F,H,W,C] = depthwise3x3( act(x) ),  act(x) = relu?( scale*x + shift ) with scale/shift optional.
extern "C" int xcp_dw3x3_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out,
                             int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "xcp_dw3x3_fwd: bad shape F=%d H=%d W=%d C=%d", F, H, W, C);
    XCP_REQUIRE((scale == nullptr) == (shift == nullptr), "xcp_dw3x3_fwd: scale/shift must both be given or both null");
    XCP_CUDA(cudaSetDevice(device));
    DwGeom g = make_geom(F, H, W, C, 32, 384, true);
    CUtensorMap tm;
    if (int e = make_dw_tmap(&tm, x, g)) return e;
    DwFwdParams p{g, w9, scale, shift, relu, (__nv_bfloat16*)out};
    const int smem = 2 * (g.TW + 2) * (g.TH + 2) * 128 + 128;
    long long grid = 2LL * num_sms();
    if (grid > g.num_tiles) grid = g.num_tiles;
    if (scale != nullptr) {
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dw3x3_fwd_kernel<true><<<(int)grid, 256, smem, (cudaStream_t)stream>>>(tm, p);
    } else {
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dw3x3_fwd_kernel<false><<<(int)grid, 256, smem, (cudaStream_t)stream>>>(tm, p);
    }
    return check_cuda(cudaGetLastError(), "dw3x3_fwd launch");
}

extern "C" long long xcp_dw3x3_bwd_workspace_floats(int C) {
    const int c_tiles = (C + 63) / 64;
    long long grid = 2LL * 160;   // upper bound on 2 * SM count
    grid = (grid / c_tiles + 1) * c_tiles;
    return grid * 11 * 64;
}

// Backward of xcp_dw3x3_fwd.  dz = mask * conv_transpose(dD) [+ add_full] [+ add_half at even pixels];
// dw9[9][C] += weight gradient; bnsum[2][C] = (sum dz, sum dz*x) per channel when scale/shift are given.
extern "C" int xcp_dw3x3_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift,
                             int relu, void* dz, const void* add_full, const void* add_half, float* dw9, float* bnsum,
                             float* workspace, int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "xcp_dw3x3_bwd: bad shape");
    XCP_REQUIRE((scale == nullptr) == (shift == nullptr), "xcp_dw3x3_bwd: scale/shift");
    XCP_REQUIRE(workspace != nullptr && dw9 != nullptr, "xcp_dw3x3_bwd: workspace / dw9 missing");
    XCP_CUDA(cudaSetDevice(device));
    DwGeom g = make_geom(F, H, W, C, 16, 384, false);
    CUtensorMap tm;
    if (int e = make_dw_tmap(&tm, dD, g)) return e;
    DwBwdParams p{g, w9, (const __nv_bfloat16*)xin, scale, shift, relu, (__nv_bfloat16*)dz,
                  (const __nv_bfloat16*)add_full, (const __nv_bfloat16*)add_half, workspace};
    const int smem = 2 * (g.TW + 2) * (g.TH + 2) * 128 + 128;
    const long long sp_tiles = (long long)F * g.n_h * g.n_w;
    long long per_ct = (2LL * num_sms()) / g.c_tiles;
    if (per_ct < 1) per_ct = 1;
    if (per_ct > sp_tiles) per_ct = sp_tiles;
    const int grid = (int)(per_ct * g.c_tiles);
    cudaStream_t st = (cudaStream_t)stream;
    if (scale != nullptr) {
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dw3x3_bwd_kernel<true><<<grid, 256, smem, st>>>(tm, p);
    } else {
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dw3x3_bwd_kernel<false><<<grid, 256, smem, st>>>(tm, p);
    }
    XCP_CUDA(cudaGetLastError());
    dw_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(workspace, grid, g.c_tiles, C, dw9, scale != nullptr ? bnsum : nullptr);
    return check_cuda(cudaGetLastError(), "dw3x3_bwd launch");
}
