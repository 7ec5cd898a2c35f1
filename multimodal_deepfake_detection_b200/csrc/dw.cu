// Depthwise 3x3 (stride 1, pad 1) forward and backward on NHWC bf16 activations.
//
// Replaces nn.Conv2d(groups=C) in SeparableConv2d.conv1 (Xception.py:41,45) together with the ReLU that
// precedes it (Xception.py:61-76) and the previous layer's BatchNorm affine (Xception.py:67,73,78), which
// are applied on the fly to the staged tile ("consumer prologue" fusion).  HBM-bound by design: every input
// element is fetched once by TMA into a shared-memory halo tile (OOB zero fill = the conv's zero padding) and
// every output element is written once.
//
// Mapping (v2, chosen after the first ncu pass showed the v1 kernels issue- and latency-bound, not DRAM-bound):
//   tile   = (TH x TW pixels) x 64 channels of one frame, 2-deep TMA ring per persistent CTA
//   warp   = one pair of adjacent pixel columns (x one row slice)      -> border tests are warp-uniform branches
//   lane   = one pair of adjacent channels, held as a packed f32x2     -> all math is FFMA2 (fma.rn.f32x2)
// so one warp instruction touches 32 lanes x 4 B = one pixel's 128 contiguous bytes in shared and global memory.
// The warp streams down its columns with a 3-row register window (forward: partial sums; backward: dD values).
#include "common.cuh"

namespace xcp {

XCP_DEVINL u64 bf2_to_f2(uint32_t v) { return pk2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
XCP_DEVINL uint32_t f2_to_bf2(u64 v) { float lo, hi; upk2(v, lo, hi); return pack_bf16(lo, hi); }
XCP_DEVINL uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
XCP_DEVINL u64 relu2(u64 v) { float lo, hi; upk2(v, lo, hi); return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }

struct DwGeom {
    int F, H, W, C;
    int TH, TW, n_h, n_w, c_tiles;   // tiling
    int pairs, RS, rows_per_slice;   // warps per tile = pairs * RS
    long long num_tiles;
};

static DwGeom make_geom(int F, int H, int W, int C, int max_halo_pixels, int max_warps) {
    DwGeom g{};
    g.F = F; g.H = H; g.W = W; g.C = C;
    g.n_w = (W + 2 * max_warps - 1) / (2 * max_warps);
    g.TW = (W + g.n_w - 1) / g.n_w;
    int th_max = max_halo_pixels / (g.TW + 2) - 2;
    if (th_max < 1) th_max = 1;
    g.n_h = (H + th_max - 1) / th_max;
    g.TH = (H + g.n_h - 1) / g.n_h;
    g.c_tiles = (C + 63) / 64;
    g.pairs = (g.TW + 1) / 2;
    g.RS = max_warps / g.pairs;
    if (g.RS < 1) g.RS = 1;
    if (g.RS > g.TH) g.RS = g.TH;
    g.rows_per_slice = (g.TH + g.RS - 1) / g.RS;
    g.RS = (g.TH + g.rows_per_slice - 1) / g.rows_per_slice;
    g.num_tiles = (long long)F * g.n_h * g.n_w * g.c_tiles;
    return g;
}

struct DwFwdParams {
    DwGeom g;
    const float* w9;      // [9][C] tap-major depthwise weights
    const float* scale;   // [C] pending BN affine of the producer (AFFINE) or null
    const float* shift;
    __nv_bfloat16* out;   // [F,H,W,C]
};

template <bool AFFINE, bool RELU>
__global__ void __launch_bounds__(256, 3)
dw3x3_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const DwFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t sbase = (raw_addr + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - raw_addr);
    const DwGeom& g = p.g;
    const int halo_w = g.TW + 2;
    const uint32_t row_stride = (uint32_t)halo_w * 128u;
    const uint32_t stage_bytes = row_stride * (uint32_t)(g.TH + 2);
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
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_4d(smem + s * stage_bytes, &tmX, &full[s], ct * 64, tw * g.TW - 1, th * g.TH - 1, (int)t);
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = (warp % g.pairs) * 2;          // first of the warp's two tile-local columns
    const int slice = warp / g.pairs;
    const int r0 = slice * g.rows_per_slice;
    const int r1 = min(r0 + g.rows_per_slice, g.TH);
    const u64 NEG = pk2(-1e30f, -1e30f);

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
        const int c0 = ct * 64 + lane * 2;
        const int gx0 = tw * g.TW + x0;
        const bool active = slice < g.RS && c0 < g.C && gx0 < g.W;      // gx0 < W is warp-uniform, c0 < C per lane
        const bool second = (x0 + 1 < g.TW) && (gx0 + 1 < g.W);

        u64 wg[9];
        u64 sc = 0, shc[4] = {0, 0, 0, 0};
        if (active) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float2 a = *reinterpret_cast<const float2*>(p.w9 + (long long)k * g.C + c0);
                wg[k] = pk2(a.x, a.y);
            }
            if (AFFINE) {
                const float2 a = *reinterpret_cast<const float2*>(p.scale + c0);
                const float2 b = *reinterpret_cast<const float2*>(p.shift + c0);
                sc = pk2(a.x, a.y);
                const u64 sh = pk2(b.x, b.y);
                // Out-of-image columns must stay exactly 0 after the affine: with the ReLU fused, a hugely negative
                // shift does that branch-free (TMA zero-filled the raw value, scale*0 = 0).  Without ReLU the
                // generic (select) path below is used.
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) {
                    const bool cv = (gx0 - 1 + dx >= 0) && (gx0 - 1 + dx < g.W);
                    shc[dx] = (cv || !RELU) ? sh : NEG;
                }
            }
        }

        mbar_wait(&full[s], (it >> 1) & 1);

        if (active) {
            const uint32_t tb = sbase + s * stage_bytes + (uint32_t)x0 * 128u + (uint32_t)lane * 4u;
            bool cvn[4];
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) cvn[dx] = (gx0 - 1 + dx >= 0) && (gx0 - 1 + dx < g.W);
            // load + transform one staged row (halo-tile row index hr) into a 4-column register row
            auto load_row = [&](int hr, u64 (&w)[4]) {
                const int gh = th * g.TH + hr - 1;
                const uint32_t a = tb + (uint32_t)hr * row_stride;
                if (AFFINE && (gh < 0 || gh >= g.H)) {          // warp-uniform: rows outside the image stay 0
                    w[0] = 0; w[1] = 0; w[2] = 0; w[3] = 0;
                    return;
                }
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) {
                    u64 z = bf2_to_f2(lds32(a + dx * 128));
                    if (AFFINE) z = fma2(z, sc, shc[dx]);
                    if (RELU) z = relu2(z);
                    if (AFFINE && !RELU && !cvn[dx]) z = 0;
                    w[dx] = z;
                }
            };
            u64 win[3][4];
            load_row(r0, win[0]);
            load_row(r0 + 1, win[1]);
            __nv_bfloat16* outp = p.out + (((long long)f * g.H + th * g.TH) * g.W + gx0) * g.C + c0;
            const long long out_row = (long long)g.W * g.C;
            const int rmax = min(r1, g.H - th * g.TH);       // output rows of this slice that are inside the image

#define DW_STEP(WA, WB, WC)                                                                               \
    {                                                                                                     \
        load_row(o + 2, WC);                                                                              \
        u64 a0 = mul2(wg[0], WA[0]), a1 = mul2(wg[0], WA[1]);                                             \
        a0 = fma2(wg[1], WA[1], a0); a1 = fma2(wg[1], WA[2], a1);                                         \
        a0 = fma2(wg[2], WA[2], a0); a1 = fma2(wg[2], WA[3], a1);                                         \
        a0 = fma2(wg[3], WB[0], a0); a1 = fma2(wg[3], WB[1], a1);                                         \
        a0 = fma2(wg[4], WB[1], a0); a1 = fma2(wg[4], WB[2], a1);                                         \
        a0 = fma2(wg[5], WB[2], a0); a1 = fma2(wg[5], WB[3], a1);                                         \
        a0 = fma2(wg[6], WC[0], a0); a1 = fma2(wg[6], WC[1], a1);                                         \
        a0 = fma2(wg[7], WC[1], a0); a1 = fma2(wg[7], WC[2], a1);                                         \
        a0 = fma2(wg[8], WC[2], a0); a1 = fma2(wg[8], WC[3], a1);                                         \
        __nv_bfloat16* op = outp + o * out_row;                                                           \
        *reinterpret_cast<uint32_t*>(op) = f2_to_bf2(a0);                                                 \
        if (second) *reinterpret_cast<uint32_t*>(op + g.C) = f2_to_bf2(a1);                               \
    }
            int o = r0;
            while (o < rmax) {
                DW_STEP(win[0], win[1], win[2]); if (++o >= rmax) break;
                DW_STEP(win[1], win[2], win[0]); if (++o >= rmax) break;
                DW_STEP(win[2], win[0], win[1]); ++o;
            }
#undef DW_STEP
        }
        __syncthreads();   // everyone is done with stage s before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  For every pixel p and channel c (a = the activated DW input, dD = grad wrt the DW output):
//     da[p]        = sum_{kh,kw} w[kh][kw] * dD[p + (1-kh, 1-kw)]
//     dw[kh][kw]  += a[p] * dD[p + (1-kh, 1-kw)]
// Both need the same 3x3 neighbourhood of dD (TMA halo tile, zero fill = correct padding) and only the centre value
// of the forward input (second TMA tile, no halo).  The warp keeps a 3-row x 4-column window of dD in registers.
// The kernel also applies the ReLU mask / BN-affine chain rule, accumulates the per-channel sums the BatchNorm
// backward of the producer needs (sum dz, sum dz*y), and can add the residual-branch gradient(s).
struct DwBwdParams {
    DwGeom g;
    const float* w9;              // [9][C]
    const float* scale;           // AFFINE only
    const float* shift;
    int relu;
    __nv_bfloat16* dz;            // out: grad wrt the pre-activation (z if AFFINE, x otherwise) [F,H,W,C]
    const __nv_bfloat16* add_full;   // optional: same-shape gradient to add (identity skip), after masking
    const __nv_bfloat16* add_half;   // optional: [F,ceil(H/2),ceil(W/2),C] gradient of the stride-2 skip gather
    float* dw;                    // [C][9]  (the nn.Conv2d weight-gradient layout), accumulated with RED
    float* bnsum;                 // [2][C]  (sum dz, sum dz*y), accumulated with RED (caller zero-fills); AFFINE only
    int c_real;                   // logical channel count (<= C, the physical pitch): dw has c_real rows
};

template <bool AFFINE, bool RELU>
__global__ void __launch_bounds__(256, 2)
dw3x3_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const DwBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t sbase = (raw_addr + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - raw_addr);
    const DwGeom& g = p.g;
    const int halo_w = g.TW + 2;
    const uint32_t row_stride = (uint32_t)halo_w * 128u;
    const uint32_t g_bytes = row_stride * (uint32_t)(g.TH + 2);
    const uint32_t x_row = (uint32_t)g.TW * 128u;
    const uint32_t x_bytes = x_row * (uint32_t)g.TH;
    const uint32_t stage_bytes = g_bytes + x_bytes;
    __shared__ uint64_t full[2];
    __shared__ float s_red[11][64];

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmX);
    }
    for (int i = threadIdx.x; i < 11 * 64; i += blockDim.x) (&s_red[0][0])[i] = 0.f;
    __syncthreads();

    // A CTA owns one channel tile for its whole life so the weight / BN partial sums stay in registers.
    const int ct = blockIdx.x % g.c_tiles;
    const long long sp_tiles = (long long)g.F * g.n_h * g.n_w;
    const int sp_stride = gridDim.x / g.c_tiles;          // host guarantees gridDim.x % c_tiles == 0
    long long sp = blockIdx.x / g.c_tiles;

    auto issue = [&](long long sp_tile, int s) {
        long long t = sp_tile;
        const int tw = (int)(t % g.n_w); t /= g.n_w;
        const int th = (int)(t % g.n_h); t /= g.n_h;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        tma_load_4d(smem + s * stage_bytes, &tmG, &full[s], ct * 64, tw * g.TW - 1, th * g.TH - 1, (int)t);
        tma_load_4d(smem + s * stage_bytes + g_bytes, &tmX, &full[s], ct * 64, tw * g.TW, th * g.TH, (int)t);
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = (warp % g.pairs) * 2;
    const int slice = warp / g.pairs;
    const int r0 = slice * g.rows_per_slice;
    const int r1 = min(r0 + g.rows_per_slice, g.TH);
    const int c0 = ct * 64 + lane * 2;
    const bool lane_active = slice < g.RS && c0 < g.C;

    u64 wg[9], dwa[9];
    u64 sdz = 0, sdzy = 0, sc = pk2(1.f, 1.f), sh = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { wg[k] = 0; dwa[k] = 0; }
    if (lane_active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(p.w9 + (long long)k * g.C + c0);
            wg[k] = pk2(a.x, a.y);
        }
        if (AFFINE) {
            const float2 a = *reinterpret_cast<const float2*>(p.scale + c0);
            const float2 b = *reinterpret_cast<const float2*>(p.shift + c0);
            sc = pk2(a.x, a.y); sh = pk2(b.x, b.y);
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
        const int gx0 = tw * g.TW + x0;
        const bool active = lane_active && gx0 < g.W;
        const bool second = (x0 + 1 < g.TW) && (gx0 + 1 < g.W);

        mbar_wait(&full[s], (it >> 1) & 1);

        if (active) {
            const uint32_t gb = sbase + s * stage_bytes + (uint32_t)x0 * 128u + (uint32_t)lane * 4u;   // dD halo tile
            const uint32_t xb = gb + g_bytes;                                                          // fwd-input centre tile
            auto load_g = [&](int hr, u64 (&w)[4]) {
                const uint32_t a = gb + (uint32_t)hr * row_stride;
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) w[dx] = bf2_to_f2(lds32(a + dx * 128));
            };
            u64 win[3][4];
            load_g(r0, win[0]);          // dD row r0-1
            load_g(r0 + 1, win[1]);      // dD row r0
            const int gh0 = th * g.TH;
            const int rmax = min(r1, g.H - gh0);
            __nv_bfloat16* dzp = p.dz + (((long long)f * g.H + gh0) * g.W + gx0) * g.C + c0;
            const __nv_bfloat16* afp = p.add_full ? p.add_full + (((long long)f * g.H + gh0) * g.W + gx0) * g.C + c0 : nullptr;
            const long long row_el = (long long)g.W * g.C;

            // centre row rc: window rows (rc-1, rc, rc+1) = (WA, WB, WC); WC is loaded here.  The neighbour
            // p + (1-kh, 1-kw) of centre (rc, o) sits in window row 2-kh (WC, WB, WA for kh = 0, 1, 2), staged column o+2-kw.
#define DWB_PIX(O, WA, WB, WC)                                                                             \
    {                                                                                                      \
        const u64 yv = bf2_to_f2(lds32(xb + (uint32_t)rc * x_row + (O) * 128));                            \
        const u64 z = AFFINE ? fma2(yv, sc, sh) : yv;                                                      \
        float zl, zh; upk2(z, zl, zh);                                                                     \
        const bool pl = RELU ? zl > 0.f : true, ph = RELU ? zh > 0.f : true;                               \
        const u64 av = RELU ? pk2(fmaxf(zl, 0.f), fmaxf(zh, 0.f)) : z;                                     \
        u64 da = mul2(wg[0], WC[(O) + 2]);                                                                 \
        da = fma2(wg[1], WC[(O) + 1], da); da = fma2(wg[2], WC[(O)], da);                                  \
        da = fma2(wg[3], WB[(O) + 2], da); da = fma2(wg[4], WB[(O) + 1], da); da = fma2(wg[5], WB[(O)], da); \
        da = fma2(wg[6], WA[(O) + 2], da); da = fma2(wg[7], WA[(O) + 1], da); da = fma2(wg[8], WA[(O)], da); \
        dwa[0] = fma2(av, WC[(O) + 2], dwa[0]); dwa[1] = fma2(av, WC[(O) + 1], dwa[1]); dwa[2] = fma2(av, WC[(O)], dwa[2]); \
        dwa[3] = fma2(av, WB[(O) + 2], dwa[3]); dwa[4] = fma2(av, WB[(O) + 1], dwa[4]); dwa[5] = fma2(av, WB[(O)], dwa[5]); \
        dwa[6] = fma2(av, WA[(O) + 2], dwa[6]); dwa[7] = fma2(av, WA[(O) + 1], dwa[7]); dwa[8] = fma2(av, WA[(O)], dwa[8]); \
        float dl, dh; upk2(da, dl, dh);                                                                    \
        dl = pl ? dl : 0.f; dh = ph ? dh : 0.f;                                                            \
        const long long eo = rc * row_el + (O) * g.C;                                                      \
        if (afp != nullptr) {                                                                              \
            const uint32_t ar = __ldg(reinterpret_cast<const uint32_t*>(afp + eo));                        \
            dl += bf16_lo(ar); dh += bf16_hi(ar);                                                          \
        }                                                                                                  \
        if (p.add_half != nullptr && (((gh0 + rc) | (gx0 + (O))) & 1) == 0) {                              \
            const int Ho = (g.H + 1) / 2, Wo = (g.W + 1) / 2;                                              \
            const uint32_t ar = __ldg(reinterpret_cast<const uint32_t*>(                                   \
                p.add_half + (((long long)f * Ho + ((gh0 + rc) >> 1)) * Wo + ((gx0 + (O)) >> 1)) * g.C + c0)); \
            dl += bf16_lo(ar); dh += bf16_hi(ar);                                                          \
        }                                                                                                  \
        if (AFFINE) { const u64 dzv = pk2(dl, dh); sdz = add2(sdz, dzv); sdzy = fma2(dzv, yv, sdzy); }     \
        *reinterpret_cast<uint32_t*>(dzp + eo) = pack_bf16(dl, dh);                                        \
    }
#define DWB_STEP(WA, WB, WC)                                                                               \
    {                                                                                                      \
        load_g(rc + 2, WC);                                                                                \
        DWB_PIX(0, WA, WB, WC)                                                                             \
        if (second) DWB_PIX(1, WA, WB, WC)                                                                 \
    }
            int rc = r0;
            while (rc < rmax) {
                DWB_STEP(win[0], win[1], win[2]); if (++rc >= rmax) break;
                DWB_STEP(win[1], win[2], win[0]); if (++rc >= rmax) break;
                DWB_STEP(win[2], win[0], win[1]); ++rc;
            }
#undef DWB_STEP
#undef DWB_PIX
        }
        __syncthreads();
    }

    // CTA reduction (shared-memory atomics, once per CTA lifetime) then one RED per (tap, channel) to global.
    if (lane_active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float lo, hi; upk2(dwa[k], lo, hi);
            atomicAdd(&s_red[k][lane * 2], lo); atomicAdd(&s_red[k][lane * 2 + 1], hi);
        }
        if (AFFINE) {
            float lo, hi;
            upk2(sdz, lo, hi); atomicAdd(&s_red[9][lane * 2], lo); atomicAdd(&s_red[9][lane * 2 + 1], hi);
            upk2(sdzy, lo, hi); atomicAdd(&s_red[10][lane * 2], lo); atomicAdd(&s_red[10][lane * 2 + 1], hi);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 11 * 64; i += blockDim.x) {
        const int k = i / 64, c = ct * 64 + (i % 64);
        if (c >= g.C) continue;
        const float v = s_red[k][i % 64];
        if (k < 9) { if (c < p.c_real) atomicAdd(&p.dw[(long long)c * 9 + k], v); }
        else if (AFFINE) atomicAdd(&p.bnsum[(long long)(k - 9) * g.C + c], v);
    }
}

static int make_dw_tmap(CUtensorMap* m, const void* base, const DwGeom& g, int halo) {
    const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.F};
    const uint64_t strides[3] = {(uint64_t)g.C * 2, (uint64_t)g.W * g.C * 2, (uint64_t)g.H * g.W * g.C * 2};
    const uint32_t box[4] = {64, (uint32_t)(g.TW + 2 * halo), (uint32_t)(g.TH + 2 * halo), 1};
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
    DwGeom g = make_geom(F, H, W, C, 272, 8);
    CUtensorMap tm;
    if (int e = make_dw_tmap(&tm, x, g, 1)) return e;
    DwFwdParams p{g, w9, scale, shift, (__nv_bfloat16*)out};
    const int smem = 2 * (g.TW + 2) * (g.TH + 2) * 128 + 384;
    const int threads = 32 * g.pairs * g.RS;
    long long grid = 3LL * num_sms();
    if (grid > g.num_tiles) grid = g.num_tiles;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FWD(A, R)                                                                                          \
    {                                                                                                             \
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_fwd_kernel<A, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        dw3x3_fwd_kernel<A, R><<<(int)grid, threads, smem, st>>>(tm, p);                                          \
    }
    if (scale != nullptr) { if (relu) LAUNCH_FWD(true, true) else LAUNCH_FWD(true, false) }
    else { if (relu) LAUNCH_FWD(false, true) else LAUNCH_FWD(false, false) }
#undef LAUNCH_FWD
    return check_cuda(cudaGetLastError(), "dw3x3_fwd launch");
}

// Backward of xcp_dw3x3_fwd.  dz = mask * conv_transpose(dD) [+ add_full] [+ add_half at even pixels];
// dw[C][9] += weight gradient (nn.Conv2d layout); bnsum[2][C] += (sum dz, sum dz*x) per channel when
// scale/shift are given (the caller zero-fills bnsum).
extern "C" int xcp_dw3x3_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift,
                             int relu, void* dz, const void* add_full, const void* add_half, float* dw, float* bnsum,
                             int F, int H, int W, int C, int c_real, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && c_real > 0 && c_real <= C, "xcp_dw3x3_bwd: bad shape");
    XCP_REQUIRE((scale == nullptr) == (shift == nullptr), "xcp_dw3x3_bwd: scale/shift");
    XCP_REQUIRE(dw != nullptr && (scale == nullptr || bnsum != nullptr), "xcp_dw3x3_bwd: dw / bnsum missing");
    XCP_CUDA(cudaSetDevice(device));
    DwGeom g = make_geom(F, H, W, C, 240, 8);
    CUtensorMap tmG, tmX;
    if (int e = make_dw_tmap(&tmG, dD, g, 1)) return e;
    if (int e = make_dw_tmap(&tmX, xin, g, 0)) return e;
    DwBwdParams p{g, w9, scale, shift, relu, (__nv_bfloat16*)dz, (const __nv_bfloat16*)add_full,
                  (const __nv_bfloat16*)add_half, dw, bnsum, c_real};
    const int smem = 2 * ((g.TW + 2) * (g.TH + 2) + g.TW * g.TH) * 128 + 384;
    const int threads = 32 * g.pairs * g.RS;
    const long long sp_tiles = (long long)F * g.n_h * g.n_w;
    long long per_ct = (2LL * num_sms()) / g.c_tiles;
    if (per_ct < 1) per_ct = 1;
    if (per_ct > sp_tiles) per_ct = sp_tiles;
    const int grid = (int)(per_ct * g.c_tiles);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_BWD(A, R)                                                                                          \
    {                                                                                                             \
        XCP_CUDA(cudaFuncSetAttribute(dw3x3_bwd_kernel<A, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        dw3x3_bwd_kernel<A, R><<<grid, threads, smem, st>>>(tmG, tmX, p);                                         \
    }
    if (scale != nullptr) { if (relu) LAUNCH_BWD(true, true) else LAUNCH_BWD(true, false) }
    else { if (relu) LAUNCH_BWD(false, true) else LAUNCH_BWD(false, false) }
#undef LAUNCH_BWD
    return check_cuda(cudaGetLastError(), "dw3x3_bwd launch");
}
