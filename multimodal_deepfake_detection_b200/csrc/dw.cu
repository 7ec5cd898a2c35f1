// Depthwise 3x3 (stride 1, pad 1) forward and backward on NHWC bf16 activations.
//
// Replaces nn.Conv2d(groups=C) in SeparableConv2d.conv1 (Xception.py:41,45) together with the ReLU that
// precedes it (Xception.py:61-76) and the previous layer's BatchNorm affine (Xception.py:67,73,78), which
// are applied on the fly to the staged tile ("consumer prologue" fusion).  HBM-bound by design: every input
// element is fetched once by TMA into a shared-memory halo tile (OOB zero fill = the conv's zero padding) and
// every output element is written once.
//
// Mapping (v3).  ncu on v2 (warp = 2 pixel columns) showed the kernels instruction-issue bound, not DRAM bound:
// 28 (forward) / 43 (backward) lane-instructions per element at ~60 % issue-slot utilisation (profiles/r1j), because a
// 2-column warp loads and transforms 4 staged columns per 2 outputs and pays the address / predicate / loop overhead
// every 4 elements.  v3:
//   tile   = (TH x TW pixels) x 64 channels of one frame, TW a multiple of 4, 2-deep TMA ring per persistent CTA
//   warp   = a STRIP of 4 adjacent pixel columns x one row slice  -> 6 staged columns per 4 outputs (1.5x instead of 2x)
//   lane   = one pair of adjacent channels, held as a packed f32x2   -> all math is FFMA2 (fma.rn.f32x2)
// so one warp instruction touches 32 lanes x 4 B = one pixel's 128 contiguous bytes in shared and global memory.
// The warp streams down its strip with a 3-row x 6-column register window; border handling is warp-uniform (tiles
// overhang the image, the overhang is TMA zero fill on the way in and a uniform predicate on the way out).
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace xcp {

XCP_DEVINL u64 bf2_to_f2(uint32_t v) { return pk2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
XCP_DEVINL uint32_t f2_to_bf2(u64 v) { float lo, hi; upk2(v, lo, hi); return pack_bf16(lo, hi); }
XCP_DEVINL uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
XCP_DEVINL u64 relu2(u64 v) { float lo, hi; upk2(v, lo, hi); return pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }

constexpr int SW = 4;            // output columns per warp strip
constexpr int LW = SW + 2;       // staged columns a strip reads

struct DwGeom {
    int F, H, W, C;
    int TH, TW, n_h, n_w, c_tiles;   // tiling (TW = SW * strips)
    int strips, RS, rows_per_slice;  // compute warps per tile = strips * RS
    int sp_tiles;                    // spatial tiles = F * n_h * n_w (one CTA keeps ONE channel tile for its whole life)
    int stages;                      // depth of the TMA ring
};

// Pick the tile shape that minimises the estimated lane-instructions per useful output element (staged-column and
// warm-up-row redundancy of the register window, tile overhang past the image) with a mild penalty for halo re-reads
// and for CTAs with few warps.  max_halo_pixels bounds the staged tile (shared memory per stage).
static DwGeom make_geom(int F, int H, int W, int C, int max_halo_pixels, int max_warps) {
    DwGeom best{};
    double best_cost = 1e30;
    {   // tuning hook: XCP_DW_GEOM="strips,slices,tile_rows" forces a geometry (tools/dw_tune.py); unset in production
        const char* e = getenv("XCP_DW_GEOM");
        int ns = 0, rs = 0, th = 0;
        if (e != nullptr && sscanf(e, "%d,%d,%d", &ns, &rs, &th) == 3 && ns > 0 && rs > 0 && th > 0) {
            DwGeom g{};
            g.F = F; g.H = H; g.W = W; g.C = C;
            g.strips = ns; g.TW = SW * ns; g.n_w = (W + g.TW - 1) / g.TW;
            g.n_h = (H + th - 1) / th; g.TH = (H + g.n_h - 1) / g.n_h;
            g.rows_per_slice = (g.TH + rs - 1) / rs;
            g.RS = (g.TH + g.rows_per_slice - 1) / g.rows_per_slice;
            g.c_tiles = (C + 63) / 64; g.sp_tiles = F * g.n_h * g.n_w;
            return g;
        }
    }
    for (int ns = 1; ns <= max_warps; ++ns) {
        if (ns > 1 && SW * (ns - 1) >= W) break;          // already wider than the image
        for (int rs_try = 1; rs_try * ns <= max_warps; ++rs_try) {
            DwGeom g{};
            g.F = F; g.H = H; g.W = W; g.C = C;
            g.strips = ns;
            g.TW = SW * ns;
            g.n_w = (W + g.TW - 1) / g.TW;
            const int th_max = max_halo_pixels / (g.TW + 2) - 2;
            if (th_max < 1) continue;
            g.n_h = (H + th_max - 1) / th_max;
            g.TH = (H + g.n_h - 1) / g.n_h;
            g.RS = rs_try;
            if (g.RS > g.TH) break;
            g.rows_per_slice = (g.TH + g.RS - 1) / g.RS;
            if ((g.TH + g.rows_per_slice - 1) / g.rows_per_slice != g.RS) continue;     // this slice count leaves a slice empty
            g.c_tiles = (C + 63) / 64;
            g.sp_tiles = F * g.n_h * g.n_w;
            const double rps = g.rows_per_slice;
            const double col_waste = (double)g.n_w * g.TW / W;
            const double row_waste = (double)g.n_h * g.RS * rps / H;
            const double instr = 3.0 * ((double)LW / SW) * (rps + 2.0) / rps + 6.5;
            const double halo = (double)(g.TW + 2) * (g.TH + 2) / ((double)g.TW * g.TH);
            const int warps = ns * g.RS;
            // few warps per SM cannot hide the LDS -> FMA -> STG latency chain: strong penalty below ~3/4 of the budget
            const double cost = instr * col_waste * row_waste * (1.0 + 0.25 * (halo - 1.0)) *
                                (1.0 + 0.6 * (double)(max_warps - warps) / max_warps);
            if (cost < best_cost) { best_cost = cost; best = g; }
        }
    }
    return best;
}

// Ring protocol shared by both kernels.  Warps 0..NW-1 compute, warp NW is the TMA producer (one elected lane):
//   producer : wait empty[s] -> arm full[s] with the stage's byte count -> issue the TMA box(es)
//   consumer : wait full[s]  -> compute from the staged tile -> (warp) arrive on empty[s]
// so a fast warp never waits for a slow one at a CTA-wide barrier (the v2/v3.0 kernels spent 1.5 issue-slots per issued
// instruction stalled at __syncthreads, profiles/r1l) and the prefetch runs `stages - 1` tiles ahead.
struct DwRing {
    uint64_t* full;
    uint64_t* empty;
    int stages;
};

struct DwFwdParams {
    DwGeom g;
    const float* w9;      // [9][C] tap-major depthwise weights
    const float* scale;   // [C] pending BN affine of the producer (AFFINE) or null
    const float* shift;
    __nv_bfloat16* out;   // [F,H,W,C]
};

constexpr int DW_MAX_STAGES = 4;

template <bool AFFINE, bool RELU, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? 512 : 256, MINB)
dw3x3_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const DwFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t sbase = (raw_addr + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - raw_addr);
    const int gH = p.g.H, gW = p.g.W, gC = p.g.C, TH = p.g.TH, TW = p.g.TW;
    const int n_h = p.g.n_h, n_w = p.g.n_w, c_tiles = p.g.c_tiles;
    const int sp_tiles = p.g.sp_tiles, stages = p.g.stages;
    const uint32_t row_stride = (uint32_t)(TW + 2) * 128u;
    const uint32_t stage_bytes = row_stride * (uint32_t)(TH + 2);
    __shared__ uint64_t full[DW_MAX_STAGES], empty[DW_MAX_STAGES];
    const int NW = p.g.strips * p.g.RS;                 // compute warps

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NW); }
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x % c_tiles;                 // host guarantees gridDim.x % c_tiles == 0
    const int sp_stride = gridDim.x / c_tiles;
    const int sp0 = blockIdx.x / c_tiles;

    // ---------------------------------------------------------------------- compute warps
    const int x0 = (warp % p.g.strips) * SW;       // first of the warp's tile-local columns
    const int slice = warp / p.g.strips;
    const int r0 = slice * p.g.rows_per_slice;
    const int r1 = min(r0 + p.g.rows_per_slice, TH);
    const int c0 = ct * 64 + lane * 2;
    const bool lane_ok = c0 < gC;
    const long long pix_b = (long long)gC * 2;      // bytes per pixel / per image row in global memory
    const long long row_b = (long long)gW * pix_b;

    // the CTA's channel tile never changes: weights and the fused affine live in registers for the whole kernel
    //
    // FOLD (BN affine + ReLU in front of the conv, 22 of the 34 depthwise layers):  relu(s*y + b) = |s| * max(sgn*y, sgn*t) + b
    // with t = -b/s, sgn = sign(s).  The scale is folded into the taps (w'_k = w_k * s) and
    // the shift becomes the constant b * sum_k w_k the accumulator starts from, so the activation costs one FMNMX per
    // staged element instead of FMNMX + FMA -- the ncu pass on v3.1 showed this instantiation math-pipe throttled
    // (42 FFMA2 = 84 FMA-pipe cycles per row step).  Zero padding: an out-of-image tap must contribute 0 = w'_k * t + b * w_k,
    // i.e. out-of-image staged values are replaced by t (border strips / rows only; warp-uniform).
    constexpr bool FOLD = AFFINE && RELU;
    u64 wg[9];
    u64 sc = 0, sh = 0, tpair = 0, cst = 0;
    uint32_t yoob = 0;
    bool pos0 = true, pos1 = true;
#pragma unroll
    for (int k = 0; k < 9; ++k) wg[k] = 0;
    if (lane_ok) {
        float2 wv[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) wv[k] = *reinterpret_cast<const float2*>(p.w9 + (long long)k * gC + c0);
        if (AFFINE) {
            const float2 a = *reinterpret_cast<const float2*>(p.scale + c0);
            const float2 b = *reinterpret_cast<const float2*>(p.shift + c0);
            sc = pk2(a.x, a.y);
            sh = pk2(b.x, b.y);
            if (FOLD) {
                float s0 = a.x, s1 = a.y, t0 = 0.f, t1 = 0.f;
                // s == 0: constant activation relu(b); a vanishing scale of the right sign reproduces it (and its zero padding)
                if (s0 == 0.f) s0 = (b.x == 0.f) ? 0.f : 1e-30f;
                if (s1 == 0.f) s1 = (b.y == 0.f) ? 0.f : 1e-30f;
                if (s0 != 0.f) t0 = -b.x / s0;
                if (s1 != 0.f) t1 = -b.y / s1;
                pos0 = s0 >= 0.f; pos1 = s1 >= 0.f;
                // negative scale: s * min(y, t) = |s| * max(-y, -t); the staged value gets its sign bit flipped on unpack
                tpair = pk2(pos0 ? t0 : -t0, pos1 ? t1 : -t1);
                // bf16 stand-in for out-of-image pixels: rounded away from the kept side so that max(sgn*y_oob, sgn*t) = sgn*t
                auto oob = [](float t, bool pos) -> uint32_t {
                    uint32_t b = __float_as_uint(t) >> 16;
                    const float v = __uint_as_float(b << 16);
                    if (pos ? (v > t) : (v < t)) b += 1;            // truncation moved it the wrong way: one bf16 ulp outwards
                    return b & 0xffffu;
                };
                yoob = oob(t0, pos0) | (oob(t1, pos1) << 16);
                float w0 = 0.f, w1 = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) { w0 += wv[k].x; w1 += wv[k].y; wv[k].x *= fabsf(s0); wv[k].y *= fabsf(s1); }
                cst = pk2(b.x * w0, b.y * w1);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) wg[k] = pk2(wv[k].x, wv[k].y);
    }

    // block-uniform: does any channel of this tile have a negative BN scale?  (gamma is initialised to 1 and is almost always
    // positive; the sign-free instantiation saves one instruction per staged element)
    const bool any_neg = FOLD ? (__syncthreads_or((warp < NW && lane_ok && !(pos0 && pos1)) ? 1 : 0) != 0) : false;
    const uint32_t sg0 = pos0 ? 0u : 0x80000000u, sg1 = pos1 ? 0u : 0x80000000u;

    if (warp == NW) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int sp = sp0; sp < sp_tiles; sp += sp_stride) {
                const int tw = sp % n_w; const int t2 = sp / n_w;
                const int th = t2 % n_h; const int f = t2 / n_h;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                tma_load_4d(smem + s * stage_bytes, &tmX, &full[s], ct * 64, tw * TW - 1, th * TH - 1, f);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    int s = 0; uint32_t ph = 0;
    for (int sp = sp0; sp < sp_tiles; sp += sp_stride) {
        const int tw = sp % n_w; const int t2 = sp / n_w;
        const int th = t2 % n_h; const int f = t2 / n_h;
        const int gx0 = tw * TW + x0;
        const int gh0 = th * TH;
        const bool active = lane_ok && gx0 < gW;                    // gx0 < W is warp-uniform, c0 < C per lane
        const int ncols = min(SW, gW - gx0);                        // valid output columns of this strip (warp-uniform)

        mbar_wait(&full[s], ph);

        if (gx0 < gW) {                 // warp-uniform; per-lane channel validity (lane_ok) is applied to the math only
            const uint32_t tb = sbase + s * stage_bytes + (uint32_t)x0 * 128u + (uint32_t)lane * 4u + (uint32_t)r0 * row_stride;
            char* op = reinterpret_cast<char*>(p.out) + ((long long)f * gH + gh0 + r0) * row_b + (long long)gx0 * pix_b +
                       (long long)c0 * 2;
            const int rmax = min(r1, gH - gh0);              // output rows of this slice that are inside the image
            if (FOLD && rmax > r0) {
                // Halo fix-up.  Staged pixels outside the image were zero-filled by TMA, but the folded activation needs them
                // to read as "t" (they must contribute exactly w'_k * t + b * w_k = 0).  Each warp patches the part of the
                // halo tile IT reads (<= 1 row above, 1 row below, 1 column left, <= 4 columns right; border tiles only) with
                // the bf16 value y_oob chosen so that max(sgn*y_oob, sgn*t) == sgn*t.  Neighbouring warps may patch the same
                // pixel with the same value.  After this the row loop has no border logic at all.
                const int nrows = rmax - r0 + 2;                                  // staged rows this warp reads
                if (gh0 + r0 == 0) {
#pragma unroll
                    for (int dx = 0; dx < LW; ++dx) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tb + dx * 128), "r"(yoob) : "memory");
                }
                if (gh0 + rmax == gH) {
                    const uint32_t a = tb + (uint32_t)(nrows - 1) * row_stride;
#pragma unroll
                    for (int dx = 0; dx < LW; ++dx) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + dx * 128), "r"(yoob) : "memory");
                }
                if (gx0 == 0) {
                    for (int r = 0; r < nrows; ++r) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tb + (uint32_t)r * row_stride), "r"(yoob) : "memory");
                }
                if (gx0 + SW >= gW) {
                    const int dx0 = gW - gx0 + 1;                                 // first staged column outside the image
                    for (int r = 0; r < nrows; ++r)
                        for (int dx = dx0; dx < LW; ++dx)
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(tb + (uint32_t)r * row_stride + dx * 128), "r"(yoob) : "memory");
                }
                __syncwarp();
                auto run = [&](auto signed_tag) {
                    if (!lane_ok) return;
                    constexpr bool SIGNED = decltype(signed_tag)::value;
                    float t0, t1; upk2(tpair, t0, t1);
                    auto load_row = [&](uint32_t a, u64 (&w)[LW]) {
#pragma unroll
                        for (int dx = 0; dx < LW; ++dx) {
                            const uint32_t v = lds32(a + dx * 128);
                            uint32_t ul = v << 16, uh = v & 0xffff0000u;
                            if (SIGNED) { ul ^= sg0; uh ^= sg1; }
                            w[dx] = pk2(fmaxf(__uint_as_float(ul), t0), fmaxf(__uint_as_float(uh), t1));
                        }
                    };
                    u64 win[3][LW];
                    uint32_t a = tb;
                    load_row(a, win[0]); a += row_stride;
                    load_row(a, win[1]); a += row_stride;
                    char* o_ptr = op;
#define DW_STEP(WA, WB, WC)                                                                                 \
    {                                                                                                       \
        load_row(a, WC); a += row_stride;                                                                   \
        _Pragma("unroll") for (int px = 0; px < SW; ++px) {                                                 \
            u64 a0 = fma2(wg[0], WA[px], cst);                                                              \
            a0 = fma2(wg[1], WA[px + 1], a0); a0 = fma2(wg[2], WA[px + 2], a0);                             \
            a0 = fma2(wg[3], WB[px], a0); a0 = fma2(wg[4], WB[px + 1], a0); a0 = fma2(wg[5], WB[px + 2], a0); \
            a0 = fma2(wg[6], WC[px], a0); a0 = fma2(wg[7], WC[px + 1], a0); a0 = fma2(wg[8], WC[px + 2], a0); \
            if (px < ncols) *reinterpret_cast<uint32_t*>(o_ptr + px * pix_b) = f2_to_bf2(a0);               \
        }                                                                                                   \
        o_ptr += row_b;                                                                                     \
    }
                    int o = r0;
                    while (o < rmax) {
                        DW_STEP(win[0], win[1], win[2]); if (++o >= rmax) break;
                        DW_STEP(win[1], win[2], win[0]); if (++o >= rmax) break;
                        DW_STEP(win[2], win[0], win[1]); ++o;
                    }
#undef DW_STEP
                };
                if (!any_neg) run(std::false_type{}); else run(std::true_type{});
            } else if (!FOLD && lane_ok) {
            // EDGE = the strip's 6 staged columns reach outside the image (TMA zero-filled them): after a fused affine
            // they must be forced back to exactly 0 (the conv's zero padding applies to the activated input).
            auto run = [&](auto edge_tag) {
                constexpr bool EDGE = decltype(edge_tag)::value;
                bool cvn[LW];
#pragma unroll
                for (int dx = 0; dx < LW; ++dx) cvn[dx] = (gx0 - 1 + dx >= 0) && (gx0 - 1 + dx < gW);
                // load + transform one staged row (a = its shared-memory address, gh = its image row) into registers
                auto load_row = [&](uint32_t a, int gh, u64 (&w)[LW]) {
                    // rows outside the image must stay exactly 0 after the fused affine: scale = shift = 0
                    const bool row_in = (gh >= 0) && (gh < gH);
                    const u64 scr = (AFFINE && !row_in) ? 0ull : sc;
                    const u64 shr = (AFFINE && !row_in) ? 0ull : sh;
#pragma unroll
                    for (int dx = 0; dx < LW; ++dx) {
                        u64 z = bf2_to_f2(lds32(a + dx * 128));
                        if (AFFINE) z = fma2(z, scr, shr);
                        if (RELU) z = relu2(z);
                        if (AFFINE && EDGE && !cvn[dx]) z = 0;
                        w[dx] = z;
                    }
                };
                u64 win[3][LW];
                uint32_t a = tb;
                int gh = gh0 + r0 - 1;
                load_row(a, gh, win[0]); a += row_stride; ++gh;
                load_row(a, gh, win[1]); a += row_stride; ++gh;
                char* o_ptr = op;

#define DW_STEP(WA, WB, WC)                                                                                 \
    {                                                                                                       \
        load_row(a, gh, WC); a += row_stride; ++gh;                                                         \
        _Pragma("unroll") for (int px = 0; px < SW; ++px) {                                                 \
            u64 a0 = mul2(wg[0], WA[px]);                                                                   \
            a0 = fma2(wg[1], WA[px + 1], a0); a0 = fma2(wg[2], WA[px + 2], a0);                             \
            a0 = fma2(wg[3], WB[px], a0); a0 = fma2(wg[4], WB[px + 1], a0); a0 = fma2(wg[5], WB[px + 2], a0); \
            a0 = fma2(wg[6], WC[px], a0); a0 = fma2(wg[7], WC[px + 1], a0); a0 = fma2(wg[8], WC[px + 2], a0); \
            if (px < ncols) *reinterpret_cast<uint32_t*>(o_ptr + px * pix_b) = f2_to_bf2(a0);               \
        }                                                                                                   \
        o_ptr += row_b;                                                                                     \
    }
                int o = r0;
                while (o < rmax) {
                    DW_STEP(win[0], win[1], win[2]); if (++o >= rmax) break;
                    DW_STEP(win[1], win[2], win[0]); if (++o >= rmax) break;
                    DW_STEP(win[2], win[0], win[1]); ++o;
                }
#undef DW_STEP
            };
            if (!AFFINE || (gx0 >= 1 && gx0 + SW < gW)) run(std::false_type{});
            else run(std::true_type{});
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);     // this warp is done with stage s
        if (++s == stages) { s = 0; ph ^= 1; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  For every pixel p and channel c (a = the activated DW input, dD = grad wrt the DW output):
//     da[p]        = sum_{kh,kw} w[kh][kw] * dD[p + (1-kh, 1-kw)]
//     dw[kh][kw]  += a[p] * dD[p + (1-kh, 1-kw)]
// Both need the same 3x3 neighbourhood of dD (TMA halo tile, zero fill = correct padding) and only the centre value
// of the forward input (second TMA tile, no halo).  The warp keeps a 3-row x 6-column window of dD in registers.
// The kernel also applies the ReLU mask / BN-affine chain rule, accumulates the per-channel sums the BatchNorm
// backward of the producer needs (sum dz, sum dz*y), and can add the residual-branch gradient(s).
struct DwBwdParams {
    DwGeom g;
    const float* w9;              // [9][C]
    const float* scale;           // AFFINE only
    const float* shift;
    __nv_bfloat16* dz;            // out: grad wrt the pre-activation (z if AFFINE, x otherwise) [F,H,W,C]
    const __nv_bfloat16* add_full;   // ADDM & 1: same-shape gradient to add (identity skip), after masking
    const __nv_bfloat16* add_half;   // ADDM & 2: [F,ceil(H/2),ceil(W/2),C] gradient of the stride-2 skip gather
    float* dw;                    // [C][9]  (the nn.Conv2d weight-gradient layout), accumulated with RED
    float* bnsum;                 // [2][C]  (sum dz, sum dz*y), accumulated with RED (caller zero-fills); AFFINE only
    int c_real;                   // logical channel count (<= C, the physical pitch): dw has c_real rows
    int add_pre;                  // the residual gradients are added BEFORE the ReLU mask (they are gradients wrt the activated input:
                                  // block 1 reading relu(bn2(y2)) straight from the stem's raw conv2 output, executor.py)
};

template <bool AFFINE, bool RELU, int ADDM, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? 512 : 256, MINB)
dw3x3_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmF, const DwBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t sbase = (raw_addr + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - raw_addr);
    const int gH = p.g.H, gW = p.g.W, gC = p.g.C, TH = p.g.TH, TW = p.g.TW;
    const int n_h = p.g.n_h, n_w = p.g.n_w, c_tiles = p.g.c_tiles;
    const int sp_tiles = p.g.sp_tiles, stages = p.g.stages;
    const uint32_t row_stride = (uint32_t)(TW + 2) * 128u;
    const uint32_t g_bytes = row_stride * (uint32_t)(TH + 2);
    const uint32_t x_row = (uint32_t)TW * 128u;
    const uint32_t x_bytes = x_row * (uint32_t)TH;
    // stage = [dD halo tile][forward-input centre tile][identity-skip gradient centre tile (ADDM & 1)]: the residual gradient is
    // staged by TMA like the other operands -- read with per-pixel __ldg it made the block-entry layers 2.6x slower than the
    // others (190 vs 73 us at 19x19x768, profiles/r1x)
    const uint32_t stage_bytes = g_bytes + x_bytes * ((ADDM & 1) ? 2u : 1u);
    __shared__ uint64_t full[DW_MAX_STAGES], empty[DW_MAX_STAGES];
    __shared__ float s_red[11][64];
    const int NW = p.g.strips * p.g.RS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NW); }
        fence_barrier_init();
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmX);
        if (ADDM & 1) tma_prefetch_desc(&tmF);
    }
    for (int i = threadIdx.x; i < 11 * 64; i += blockDim.x) (&s_red[0][0])[i] = 0.f;
    __syncthreads();

    // A CTA owns one channel tile for its whole life so the weight / BN partial sums stay in registers.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct = blockIdx.x % c_tiles;                 // host guarantees gridDim.x % c_tiles == 0
    const int sp_stride = gridDim.x / c_tiles;
    const int sp0 = blockIdx.x / c_tiles;

    const int x0 = (warp % p.g.strips) * SW;
    const int slice = warp / p.g.strips;
    const int r0 = slice * p.g.rows_per_slice;
    const int r1 = min(r0 + p.g.rows_per_slice, TH);
    const int c0 = ct * 64 + lane * 2;
    const bool lane_ok = warp < NW && c0 < gC;
    const long long pix_b = (long long)gC * 2;
    const long long row_b = (long long)gW * pix_b;
    const int Ho = (gH + 1) / 2, Wo = (gW + 1) / 2;

    u64 wg[9], dwa[9];
    u64 sdz = 0, sdzy = 0, sc = pk2(1.f, 1.f), sh = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { wg[k] = 0; dwa[k] = 0; }

    if (warp == NW) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int sp = sp0; sp < sp_tiles; sp += sp_stride) {
                const int tw = sp % n_w; const int t2 = sp / n_w;
                const int th = t2 % n_h; const int f = t2 / n_h;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                tma_load_4d(smem + s * stage_bytes, &tmG, &full[s], ct * 64, tw * TW - 1, th * TH - 1, f);
                tma_load_4d(smem + s * stage_bytes + g_bytes, &tmX, &full[s], ct * 64, tw * TW, th * TH, f);
                if (ADDM & 1) tma_load_4d(smem + s * stage_bytes + g_bytes + x_bytes, &tmF, &full[s], ct * 64, tw * TW, th * TH, f);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        if (lane_ok) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float2 a = *reinterpret_cast<const float2*>(p.w9 + (long long)k * gC + c0);
                wg[k] = pk2(a.x, a.y);
            }
            if (AFFINE) {
                const float2 a = *reinterpret_cast<const float2*>(p.scale + c0);
                const float2 b = *reinterpret_cast<const float2*>(p.shift + c0);
                sc = pk2(a.x, a.y); sh = pk2(b.x, b.y);
            }
        }
        int s = 0; uint32_t ph = 0;
        for (int sp = sp0; sp < sp_tiles; sp += sp_stride) {
            const int tw = sp % n_w; const int t2 = sp / n_w;
            const int th = t2 % n_h; const int f = t2 / n_h;
            const int gx0 = tw * TW + x0;                 // always even (TW and x0 are multiples of 4)
            const int gh0 = th * TH;
            const bool active = lane_ok && gx0 < gW;
            const int ncols = min(SW, gW - gx0);

            mbar_wait(&full[s], ph);

            if (active) {
                const uint32_t gb = sbase + s * stage_bytes + (uint32_t)x0 * 128u + (uint32_t)lane * 4u + (uint32_t)r0 * row_stride;
                const uint32_t xb = sbase + s * stage_bytes + g_bytes + (uint32_t)x0 * 128u + (uint32_t)lane * 4u + (uint32_t)r0 * x_row;
                const int rmax = min(r1, gH - gh0);
                const long long e0 = ((long long)f * gH + gh0 + r0) * row_b + (long long)gx0 * pix_b + (long long)c0 * 2;
                const bool pre = (ADDM != 0) && p.add_pre != 0;
                auto run = [&](auto edge_tag) {
                    constexpr bool EDGE = decltype(edge_tag)::value;      // partial strip: ncols < SW
                    auto load_g = [&](uint32_t a, u64 (&w)[LW]) {
#pragma unroll
                        for (int dx = 0; dx < LW; ++dx) w[dx] = bf2_to_f2(lds32(a + dx * 128));
                    };
                    u64 win[3][LW];
                    uint32_t ga = gb, xa = xb;
                    load_g(ga, win[0]); ga += row_stride;        // dD row r0-1
                    load_g(ga, win[1]); ga += row_stride;        // dD row r0
                    char* dzp = reinterpret_cast<char*>(p.dz) + e0;
                    uint32_t fa = xb + x_bytes;                  // identity-skip gradient, centre tile (ADDM & 1)
                    // stride-2 skip gradient: lives at the even (row, column) pixels; gx0 is even, so columns px = 0, 2
                    const char* hfp = nullptr;
                    const long long hrow_b = (long long)Wo * pix_b;
                    if (ADDM & 2) hfp = reinterpret_cast<const char*>(p.add_half) + ((long long)f * Ho * Wo + (gx0 >> 1)) * pix_b + (long long)c0 * 2;
                    int gh = gh0 + r0;

                    // centre row: window rows (rc-1, rc, rc+1) = (WA, WB, WC); WC is loaded here.  The neighbour
                    // p + (1-kh, 1-kw) of centre (rc, px) sits in window row 2-kh (WC, WB, WA for kh = 0, 1, 2), staged column px+2-kw.
#define DWB_STEP(WA, WB, WC)                                                                               \
    {                                                                                                      \
        load_g(ga, WC); ga += row_stride;                                                                  \
        const char* hrow = nullptr;                                                                        \
        if ((ADDM & 2) && (gh & 1) == 0) hrow = hfp + (long long)(gh >> 1) * hrow_b;                       \
        _Pragma("unroll") for (int px = 0; px < SW; ++px) {                                                \
            if (!EDGE || px < ncols) {                                                                     \
                const u64 yv = bf2_to_f2(lds32(xa + px * 128));                                            \
                const u64 z = AFFINE ? fma2(yv, sc, sh) : yv;                                              \
                float zl, zh; upk2(z, zl, zh);                                                             \
                const u64 av = RELU ? pk2(fmaxf(zl, 0.f), fmaxf(zh, 0.f)) : z;                             \
                u64 da = mul2(wg[0], WC[px + 2]);                                                          \
                da = fma2(wg[1], WC[px + 1], da); da = fma2(wg[2], WC[px], da);                            \
                da = fma2(wg[3], WB[px + 2], da); da = fma2(wg[4], WB[px + 1], da); da = fma2(wg[5], WB[px], da); \
                da = fma2(wg[6], WA[px + 2], da); da = fma2(wg[7], WA[px + 1], da); da = fma2(wg[8], WA[px], da); \
                dwa[0] = fma2(av, WC[px + 2], dwa[0]); dwa[1] = fma2(av, WC[px + 1], dwa[1]); dwa[2] = fma2(av, WC[px], dwa[2]); \
                dwa[3] = fma2(av, WB[px + 2], dwa[3]); dwa[4] = fma2(av, WB[px + 1], dwa[4]); dwa[5] = fma2(av, WB[px], dwa[5]); \
                dwa[6] = fma2(av, WA[px + 2], dwa[6]); dwa[7] = fma2(av, WA[px + 1], dwa[7]); dwa[8] = fma2(av, WA[px], dwa[8]); \
                float dl, dh; upk2(da, dl, dh);                                                            \
                if (RELU && !pre) { dl = zl > 0.f ? dl : 0.f; dh = zh > 0.f ? dh : 0.f; }                  \
                if (ADDM & 1) {                                                                            \
                    const uint32_t ar = lds32(fa + px * 128);                                              \
                    dl += bf16_lo(ar); dh += bf16_hi(ar);                                                  \
                }                                                                                          \
                if ((ADDM & 2) && (px & 1) == 0 && hrow != nullptr) {                                      \
                    const uint32_t ar = __ldg(reinterpret_cast<const uint32_t*>(hrow + (px >> 1) * pix_b)); \
                    dl += bf16_lo(ar); dh += bf16_hi(ar);                                                  \
                }                                                                                          \
                if (RELU && pre) { dl = zl > 0.f ? dl : 0.f; dh = zh > 0.f ? dh : 0.f; }                   \
                if (AFFINE) { const u64 dzv = pk2(dl, dh); sdz = add2(sdz, dzv); sdzy = fma2(dzv, yv, sdzy); } \
                *reinterpret_cast<uint32_t*>(dzp + px * pix_b) = pack_bf16(dl, dh);                        \
            }                                                                                              \
        }                                                                                                  \
        xa += x_row; dzp += row_b; ++gh;                                                                   \
        if (ADDM & 1) fa += x_row;                                                                         \
    }
                    int rc = r0;
                    while (rc < rmax) {
                        DWB_STEP(win[0], win[1], win[2]); if (++rc >= rmax) break;
                        DWB_STEP(win[1], win[2], win[0]); if (++rc >= rmax) break;
                        DWB_STEP(win[2], win[0], win[1]); ++rc;
                    }
#undef DWB_STEP
                };
                if (ncols == SW) run(std::false_type{});
                else run(std::true_type{});
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    }

    // CTA reduction (shared-memory atomics, once per CTA lifetime) then one RED per (tap, channel) to global.
    if (lane_ok) {
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
        if (c >= gC) continue;
        const float v = s_red[k][i % 64];
        if (k < 9) { if (c < p.c_real) atomicAdd(&p.dw[(long long)c * 9 + k], v); }
        else if (AFFINE) atomicAdd(&p.bnsum[(long long)(k - 9) * gC + c], v);
    }
}

static int make_dw_tmap(CUtensorMap* m, const void* base, const DwGeom& g, int halo) {
    const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.F};
    const uint64_t strides[3] = {(uint64_t)g.C * 2, (uint64_t)g.W * g.C * 2, (uint64_t)g.H * g.W * g.C * 2};
    const uint32_t box[4] = {64, (uint32_t)(g.TW + 2 * halo), (uint32_t)(g.TH + 2 * halo), 1};
    return make_tmap_4d(m, base, dims, strides, box, 0);
}

// persistent grid: `ctas_per_sm` resident CTAs per SM, rounded down to a multiple of the channel tiles (>= c_tiles)
static int dw_grid(const DwGeom& g, int ctas_per_sm) {
    long long per_ct = ((long long)ctas_per_sm * num_sms()) / g.c_tiles;
    if (per_ct < 1) per_ct = 1;
    if (per_ct > g.sp_tiles) per_ct = g.sp_tiles;
    return (int)(per_ct * g.c_tiles);
}

int dw_small_try_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift, int relu, void* dz,
                     const void* add_full, const void* add_half, float* dw, float* bnsum, int F, int H, int W, int C, int c_real,
                     cudaStream_t st, int* handled);
int dws_try_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H, int W,
                int C, cudaStream_t st, int* handled);          // dw_stream.cu

}  // namespace xcp

using namespace xcp;

// out[F,H,W,C] = depthwise3x3( act(x) ),  act(x) = relu?( scale*x + shift ) with scale/shift optional.
// ---------------------------------------------------------------------------------------------------------------------
// Tiny images (S x S, S <= 8: the 8x8 / 4x4 / 2x2 maps of the audio model's 64x64 patches, XceptionLSTMA.py:46; measured at
// 960 patches, tools/dw_rows_ab.py: 8x8x768 76.8 -> 47.1 us, 4x4x768 66.6 -> 17.4 us, 2x2x1536 101.4 -> 13.3 us).  A TMA halo
// tile per image would move (S+2)^2 / S^2 = 2.25x the bytes at S = 4 in 6x6-pixel boxes; instead one thread owns a whole
// image of one channel pair in registers: S^2 coalesced 4-byte loads (a warp reads 128 contiguous bytes per pixel), the
// BN-affine / ReLU prologue once per pixel, all S^2 outputs from registers with compile-time border handling.  Every
// input byte is read once and every output byte written once.
template <int S, bool AFFINE, bool RELU>
__global__ void __launch_bounds__(128)
dw3x3_small_fwd_kernel(const __nv_bfloat162* __restrict__ x, const float* __restrict__ w9, const float* __restrict__ scale,
                       const float* __restrict__ shift, __nv_bfloat162* __restrict__ out, long long n_items, int C) {
    const int C2 = C >> 1;
    const long long idx = blockIdx.x * 128LL + threadIdx.x;
    if (idx >= n_items) return;
    const int c2 = (int)(idx % C2);
    const long long f = idx / C2;
    float2 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float2*>(w9 + (long long)k * C + 2 * c2);
    float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
    if (AFFINE) { sc = *reinterpret_cast<const float2*>(scale + 2 * c2); sh = *reinterpret_cast<const float2*>(shift + 2 * c2); }
    const __nv_bfloat162* xp = x + f * (S * S) * C2 + c2;
    float2 a[S][S];
#pragma unroll
    for (int y = 0; y < S; ++y)
#pragma unroll
        for (int xx = 0; xx < S; ++xx) {
            float2 v = __bfloat1622float2(xp[(long long)(y * S + xx) * C2]);
            if (AFFINE) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); }
            if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
            a[y][xx] = v;
        }
    __nv_bfloat162* op = out + f * (S * S) * C2 + c2;
#pragma unroll
    for (int y = 0; y < S; ++y)
#pragma unroll
        for (int xx = 0; xx < S; ++xx) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int yy = y + ky - 1, xq = xx + kx - 1;
                    if (yy >= 0 && yy < S && xq >= 0 && xq < S) {
                        acc.x = fmaf(a[yy][xq].x, wk[ky * 3 + kx].x, acc.x);
                        acc.y = fmaf(a[yy][xq].y, wk[ky * 3 + kx].y, acc.y);
                    }
                }
            op[(long long)(y * S + xx) * C2] = __floats2bfloat162_rn(acc.x, acc.y);
        }
}

// Narrow images (W = 10 / 15: the exit flow at 299x299, block 2 of the audio model).  One thread owns a channel
// pair of one image and walks down its rows with a three-row fp32 window in registers (rows y-1, y, y+1, already through the
// BN-affine / ReLU prologue) while the raw loads of row y+2 are in flight: every input element is loaded exactly once with
// warp-coalesced 128-byte accesses, no halo is re-read, nothing is staged in shared memory, and there is no tile geometry
// (the TMA halo-tile kernel above moves 6-wide boxes for 4 output columns and is latency bound on such small images).
template <int W, bool AFFINE, bool RELU>
__global__ void __launch_bounds__(128, 3)
dw3x3_rows_fwd_kernel(const __nv_bfloat162* __restrict__ x, const float* __restrict__ w9, const float* __restrict__ scale,
                      const float* __restrict__ shift, __nv_bfloat162* __restrict__ out, long long n_items, int C, int H) {
    const int C2 = C >> 1;
    const long long idx = blockIdx.x * 128LL + threadIdx.x;
    if (idx >= n_items) return;
    const int c2 = (int)(idx % C2);
    const long long f = idx / C2;
    float2 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float2*>(w9 + (long long)k * C + 2 * c2);
    float2 sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
    if (AFFINE) { sc = *reinterpret_cast<const float2*>(scale + 2 * c2); sh = *reinterpret_cast<const float2*>(shift + 2 * c2); }
    const __nv_bfloat162* xp = x + f * H * W * C2 + c2;
    __nv_bfloat162* op = out + f * H * W * C2 + c2;
    const long long rs = (long long)W * C2;                 // row stride in channel pairs

    float2 r0[W], r1[W], r2[W];
    __nv_bfloat162 raw[W];
    auto fetch = [&](int y) {                                // issue the loads of row y (zeros outside the image)
#pragma unroll
        for (int i = 0; i < W; ++i) raw[i] = (y < H) ? xp[y * rs + (long long)i * C2] : __floats2bfloat162_rn(0.f, 0.f);
    };
    auto convert = [&](float2 (&dst)[W], int y) {            // prologue of row y into the window (zero row outside the image)
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float2 v = __bfloat1622float2(raw[i]);
            if (AFFINE) { v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); }
            if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
            if (y >= H) v = make_float2(0.f, 0.f);
            dst[i] = v;
        }
    };
    auto emit = [&](const float2 (&a)[W], const float2 (&b)[W], const float2 (&c)[W], int y) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int j = i + kx - 1;
                if (j >= 0 && j < W) {
                    acc.x = fmaf(a[j].x, wk[kx].x, acc.x); acc.y = fmaf(a[j].y, wk[kx].y, acc.y);
                    acc.x = fmaf(b[j].x, wk[3 + kx].x, acc.x); acc.y = fmaf(b[j].y, wk[3 + kx].y, acc.y);
                    acc.x = fmaf(c[j].x, wk[6 + kx].x, acc.x); acc.y = fmaf(c[j].y, wk[6 + kx].y, acc.y);
                }
            }
            op[y * rs + (long long)i * C2] = __floats2bfloat162_rn(acc.x, acc.y);
        }
    };
#pragma unroll
    for (int i = 0; i < W; ++i) r0[i] = make_float2(0.f, 0.f);          // row -1
    fetch(0); convert(r1, 0);
    fetch(1); convert(r2, 1);
    fetch(2);
    // three steps per trip so the window rotates by renaming, not by copying
    for (int y = 0; y < H; y += 3) {
        emit(r0, r1, r2, y);
        convert(r0, y + 2); fetch(y + 3);
        if (y + 1 < H) emit(r1, r2, r0, y + 1);
        convert(r1, y + 3); fetch(y + 4);
        if (y + 2 < H) emit(r2, r0, r1, y + 2);
        convert(r2, y + 4); fetch(y + 5);
    }
}

// Same walk with the channel pitch C as a compile-time constant and packed f32x2 arithmetic: every load / store of a row is
// one base register + an immediate offset (no per-access 64-bit address arithmetic), the prologue is one FFMA2 + two FMNMX
// and each tap one FFMA2 per channel pair: 17 instructions per pixel pair instead of 39 (ncu r2w: the generic kernel above
// was issue-limited at 48 % issue-active with 10 warps per SM).
template <int W, int C, bool AFFINE, bool RELU, int MINB>
__global__ void __launch_bounds__(128, MINB)
dw3x3_rowsc_fwd_kernel(const uint32_t* __restrict__ x, const float* __restrict__ w9, const float* __restrict__ scale,
                       const float* __restrict__ shift, uint32_t* __restrict__ out, long long n_items, int H) {
    constexpr int C2 = C / 2;
    const long long idx = blockIdx.x * 128LL + threadIdx.x;
    if (idx >= n_items) return;
    const int c2 = (int)(idx % C2);
    const long long f = idx / C2;
    u64 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const float2 t = *reinterpret_cast<const float2*>(w9 + k * C + 2 * c2); wk[k] = pk2(t.x, t.y); }
    u64 sc = pk2(1.f, 1.f), sh = pk2(0.f, 0.f);
    if (AFFINE) {
        const float2 a = *reinterpret_cast<const float2*>(scale + 2 * c2), b = *reinterpret_cast<const float2*>(shift + 2 * c2);
        sc = pk2(a.x, a.y); sh = pk2(b.x, b.y);
    }
    const uint32_t* xrow = x + f * H * (W * C2) + c2;       // row being FETCHED
    uint32_t* orow = out + f * H * (W * C2) + c2;           // row being EMITTED

    u64 r0[W], r1[W], r2[W];
    uint32_t raw[W];
    int yf = 0;                                              // next row to fetch
    auto fetch = [&]() {
        if (yf < H) {
#pragma unroll
            for (int i = 0; i < W; ++i) raw[i] = xrow[i * C2];
        }
        xrow += W * C2; ++yf;
    };
    auto convert = [&](u64 (&dst)[W], bool inside) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            u64 v = pk2(__uint_as_float(raw[i] << 16), __uint_as_float(raw[i] & 0xffff0000u));
            if (AFFINE) v = fma2(v, sc, sh);
            if (RELU) { float lo, hi; upk2(v, lo, hi); v = pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
            dst[i] = inside ? v : 0ull;
        }
    };
    auto emit = [&](const u64 (&a)[W], const u64 (&b)[W], const u64 (&c)[W]) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            u64 acc = 0ull;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int j = i + kx - 1;
                if (j >= 0 && j < W) {
                    acc = fma2(a[j], wk[kx], acc);
                    acc = fma2(b[j], wk[3 + kx], acc);
                    acc = fma2(c[j], wk[6 + kx], acc);
                }
            }
            float lo, hi;
            upk2(acc, lo, hi);
            const __nv_bfloat162 o = __floats2bfloat162_rn(lo, hi);
            orow[i * C2] = *reinterpret_cast<const uint32_t*>(&o);
        }
        orow += W * C2;
    };
#pragma unroll
    for (int i = 0; i < W; ++i) r0[i] = 0ull;               // row -1
    fetch(); convert(r1, true);                             // row 0
    fetch(); convert(r2, 1 < H);                            // row 1
    fetch();                                                // row 2 in flight
    for (int y = 0; y < H; y += 3) {
        emit(r0, r1, r2);
        convert(r0, y + 2 < H); fetch();
        if (y + 1 < H) emit(r1, r2, r0);
        convert(r1, y + 3 < H); fetch();
        if (y + 2 < H) emit(r2, r0, r1);
        convert(r2, y + 4 < H); fetch();
    }
}

template <int W, int C, int MINB>
static int launch_dw_rowsc(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H,
                           cudaStream_t st) {
    const long long n_items = (long long)F * (C / 2);
    const unsigned grid = (unsigned)((n_items + 127) / 128);
    const uint32_t* xi = (const uint32_t*)x;
    uint32_t* o = (uint32_t*)out;
    if (scale != nullptr) {
        if (relu) dw3x3_rowsc_fwd_kernel<W, C, true, true, MINB><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, H);
        else dw3x3_rowsc_fwd_kernel<W, C, true, false, MINB><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, H);
    } else {
        if (relu) dw3x3_rowsc_fwd_kernel<W, C, false, true, MINB><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, H);
        else dw3x3_rowsc_fwd_kernel<W, C, false, false, MINB><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, H);
    }
    return check_cuda(cudaGetLastError(), "dw3x3_rowsc_fwd launch");
}

template <int W>
static int launch_dw_rows(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H,
                          int C, cudaStream_t st) {
    const long long n_items = (long long)F * (C / 2);
    const unsigned grid = (unsigned)((n_items + 127) / 128);
    const __nv_bfloat162* xi = (const __nv_bfloat162*)x;
    __nv_bfloat162* o = (__nv_bfloat162*)out;
    if (scale != nullptr) {
        if (relu) dw3x3_rows_fwd_kernel<W, true, true><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C, H);
        else dw3x3_rows_fwd_kernel<W, true, false><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C, H);
    } else {
        if (relu) dw3x3_rows_fwd_kernel<W, false, true><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C, H);
        else dw3x3_rows_fwd_kernel<W, false, false><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C, H);
    }
    return check_cuda(cudaGetLastError(), "dw3x3_rows_fwd launch");
}

template <int S>
static int launch_dw_small(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int C,
                           cudaStream_t st) {
    const long long n_items = (long long)F * (C / 2);
    const unsigned grid = (unsigned)((n_items + 127) / 128);
    const __nv_bfloat162* xi = (const __nv_bfloat162*)x;
    __nv_bfloat162* o = (__nv_bfloat162*)out;
    if (scale != nullptr) {
        if (relu) dw3x3_small_fwd_kernel<S, true, true><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C);
        else dw3x3_small_fwd_kernel<S, true, false><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C);
    } else {
        if (relu) dw3x3_small_fwd_kernel<S, false, true><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C);
        else dw3x3_small_fwd_kernel<S, false, false><<<grid, 128, 0, st>>>(xi, w9, scale, shift, o, n_items, C);
    }
    return check_cuda(cudaGetLastError(), "dw3x3_small_fwd launch");
}

extern "C" int xcp_dw3x3_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out,
                             int F, int H, int W, int C, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "xcp_dw3x3_fwd: bad shape F=%d H=%d W=%d C=%d", F, H, W, C);
    XCP_REQUIRE((scale == nullptr) == (shift == nullptr), "xcp_dw3x3_fwd: scale/shift must both be given or both null");
    XCP_REQUIRE((long long)F * H * W < (1LL << 30), "xcp_dw3x3_fwd: too many pixels for 32-bit tile indices");
    XCP_CUDA(cudaSetDevice(device));
    {   // row-stream kernels (dw_stream.cu) for the shapes they are instantiated for
        int handled = 0;
        const int r = dws_try_fwd(x, w9, scale, shift, relu, out, F, H, W, C, (cudaStream_t)stream, &handled);
        if (handled) return r;
    }
    if (H == W && H <= 8) {
        const char* e = getenv("XCP_DW_NO_SMALL");                                    // A/B hook
        if (!(e && e[0] == '1')) {
            cudaStream_t st = (cudaStream_t)stream;
            switch (H) {
                case 1: return launch_dw_small<1>(x, w9, scale, shift, relu, out, F, C, st);
                case 2: return launch_dw_small<2>(x, w9, scale, shift, relu, out, F, C, st);
                case 3: return launch_dw_small<3>(x, w9, scale, shift, relu, out, F, C, st);
                case 4: return launch_dw_small<4>(x, w9, scale, shift, relu, out, F, C, st);
                case 5: return launch_dw_small<5>(x, w9, scale, shift, relu, out, F, C, st);
                case 6: return launch_dw_small<6>(x, w9, scale, shift, relu, out, F, C, st);
                case 7: return launch_dw_small<7>(x, w9, scale, shift, relu, out, F, C, st);
                default: return launch_dw_small<8>(x, w9, scale, shift, relu, out, F, C, st);
            }
        }
    }
    // generic-pitch variant (tools/dw_rows_ab.py, gpurun r2q, 256 / 960 frames): 10x10x1024 50.3 -> 43.9 us, 15x15x256
    // 74.8 -> 65.5 us; at W = 19 it only ties the TMA kernel (94 vs 91 us: spills at 168 registers), so other 19-wide pitches
    // stay on the tile kernel.
    // compile-time-pitch FFMA2 variant for the three shapes of the 299x299 plan (tools/dw_rows_ab.py, gpurun r2x, 256 frames):
    // 19x19x768 89.1 -> 78.8 us (3.6 TB/s; 252 registers, 8 warps/SM), 10x10x1024 50.1 -> 32.8 us, 10x10x1536 70.8 -> 45.2 us
    if ((W == 10 || W == 19) && H <= 64) {
        const char* e = getenv("XCP_DW_NO_ROWS");                                     // A/B hook
        const char* e2 = getenv("XCP_DW_NO_ROWSC");
        if (!(e && e[0] == '1') && !(e2 && e2[0] == '1')) {
            cudaStream_t st = (cudaStream_t)stream;
            // (a column-split form -- two threads per row, 12-column windows, 166 registers, 12 warps per SM instead of 8 --
            //  was slower: 84.1 us vs 79.9 us, gpurun r3a; the extra halo loads and instructions outweigh the occupancy)
            if (W == 19 && C == 768) return launch_dw_rowsc<19, 768, 2>(x, w9, scale, shift, relu, out, F, H, st);
            if (W == 10 && C == 1024) return launch_dw_rowsc<10, 1024, 4>(x, w9, scale, shift, relu, out, F, H, st);
            if (W == 10 && C == 1536) return launch_dw_rowsc<10, 1536, 4>(x, w9, scale, shift, relu, out, F, H, st);
        }
    }
    if ((W == 10 || W == 15) && H <= 64) {
        const char* e = getenv("XCP_DW_NO_ROWS");                                     // A/B hook
        if (!(e && e[0] == '1')) {
            cudaStream_t st = (cudaStream_t)stream;
            if (W == 10) return launch_dw_rows<10>(x, w9, scale, shift, relu, out, F, H, C, st);
            return launch_dw_rows<15>(x, w9, scale, shift, relu, out, F, H, C, st);
        }
    }
    // resident CTAs per SM: ONE 512-thread CTA (15 compute warps, tiles up to 800 staged pixels) or two 256-thread CTAs
    // (7 compute warps each, tiles up to 400 pixels); both run spill-free at <= 128 registers (three 80-register CTAs spilled
    // and were 10-20 % slower, gpurun r1m).  The single large CTA has less halo and tile-rounding waste and wins on the big
    // entry-flow images (276 vs 369 us at 147x147x128, 128 frames) and at 19x19; the pair wins at 37x37 and 10x10 (gpurun r2e).
    int dbg_minb = 0;
    { const char* e = getenv("XCP_DW_MINB"); dbg_minb = e ? atoi(e) : 0; }          // tuning hook (tools/dw_tune.py)
    const int hw = H > W ? H : W;
    const int minb = dbg_minb > 0 ? dbg_minb : ((hw >= 64 || (hw > 12 && hw <= 24)) ? 1 : 2);
    DwGeom g = make_geom(F, H, W, C, minb == 1 ? 800 : 400, minb == 1 ? 15 : 7);
    const int stage_bytes = (g.TW + 2) * (g.TH + 2) * 128;
    g.stages = ((minb == 1 ? 220 : 110) * 1024) / stage_bytes;
    if (g.stages > DW_MAX_STAGES) g.stages = DW_MAX_STAGES;
    if (g.stages < 2) g.stages = 2;
    CUtensorMap tm;
    if (int e = make_dw_tmap(&tm, x, g, 1)) return e;
    DwFwdParams p{g, w9, scale, shift, (__nv_bfloat16*)out};
    const int smem = g.stages * stage_bytes + 256;
    const int threads = 32 * (g.strips * g.RS + 1);
    const int grid = dw_grid(g, minb);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FWD(A, R)                                                                                                 \
    {                                                                                                                    \
        if (minb == 1) {                                                                                                 \
            XCP_CUDA(cudaFuncSetAttribute(dw3x3_fwd_kernel<A, R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            dw3x3_fwd_kernel<A, R, 1><<<grid, threads, smem, st>>>(tm, p);                                               \
        } else {                                                                                                         \
            XCP_CUDA(cudaFuncSetAttribute(dw3x3_fwd_kernel<A, R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            dw3x3_fwd_kernel<A, R, 2><<<grid, threads, smem, st>>>(tm, p);                                               \
        }                                                                                                                \
    }
    if (scale != nullptr) { if (relu) LAUNCH_FWD(true, true) else LAUNCH_FWD(true, false) }
    else { if (relu) LAUNCH_FWD(false, true) else LAUNCH_FWD(false, false) }
#undef LAUNCH_FWD
    return check_cuda(cudaGetLastError(), "dw3x3_fwd launch");
}

// Backward of xcp_dw3x3_fwd.  dz = mask * conv_transpose(dD) [+ add_full] [+ add_half at even pixels]; relu = 2: the adds are
// gradients wrt the ACTIVATED input and go inside the mask: dz = mask * (conv_transpose(dD) + add_full + add_half);
// dw[c_real][9] += weight gradient (nn.Conv2d layout); bnsum[2][C] += (sum dz, sum dz*x) per channel when
// scale/shift are given (the caller zero-fills bnsum).
extern "C" int xcp_dw3x3_bwd(const void* dD, const void* xin, const float* w9, const float* scale, const float* shift,
                             int relu, void* dz, const void* add_full, const void* add_half, float* dw, float* bnsum,
                             int F, int H, int W, int C, int c_real, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && c_real > 0 && c_real <= C, "xcp_dw3x3_bwd: bad shape");
    XCP_REQUIRE((scale == nullptr) == (shift == nullptr), "xcp_dw3x3_bwd: scale/shift");
    XCP_REQUIRE(dw != nullptr && (scale == nullptr || bnsum != nullptr), "xcp_dw3x3_bwd: dw / bnsum missing");
    XCP_REQUIRE((long long)F * H * W < (1LL << 30), "xcp_dw3x3_bwd: too many pixels for 32-bit tile indices");
    XCP_CUDA(cudaSetDevice(device));
    {   // 2x2 / 4x4 maps (audio model): register-resident whole-image kernel (dw_small_bwd.cu)
        int handled = 0;
        const int r = dw_small_try_bwd(dD, xin, w9, scale, shift, relu, dz, add_full, add_half, dw, bnsum, F, H, W, C, c_real,
                                       (cudaStream_t)stream, &handled);
        if (handled) return r;
    }
    // one 512-thread CTA per SM (15 compute warps) measured faster than two 256-thread CTAs on every shape (477 vs 623 us at
    // 147x147x128, 213 vs 309 us at 37x37x728, 65 vs 76 us at 19x19x728; gpurun r2e) except with the staged identity-skip tile
    int dbg_minb = 0;
    { const char* e = getenv("XCP_DW_MINB"); dbg_minb = e ? atoi(e) : 0; }          // tuning hook (tools/dw_tune.py)
    const int minb = dbg_minb > 0 ? dbg_minb : (add_full != nullptr ? 2 : 1);
    const int n_centre = add_full != nullptr ? 2 : 1;
    DwGeom g = make_geom(F, H, W, C, minb == 1 ? (add_full != nullptr ? 330 : 470) : (add_full != nullptr ? 170 : 200), minb == 1 ? 15 : 7);
    const int stage_bytes = ((g.TW + 2) * (g.TH + 2) + g.TW * g.TH * n_centre) * 128;
    g.stages = ((minb == 1 ? 216 : 108) * 1024) / stage_bytes;
    if (g.stages > DW_MAX_STAGES) g.stages = DW_MAX_STAGES;
    if (g.stages < 2) g.stages = 2;
    CUtensorMap tmG, tmX, tmF;
    if (int e = make_dw_tmap(&tmG, dD, g, 1)) return e;
    if (int e = make_dw_tmap(&tmX, xin, g, 0)) return e;
    if (int e = make_dw_tmap(&tmF, add_full != nullptr ? add_full : xin, g, 0)) return e;
    DwBwdParams p{g, w9, scale, shift, (__nv_bfloat16*)dz, (const __nv_bfloat16*)add_full,
                  (const __nv_bfloat16*)add_half, dw, bnsum, c_real, relu == 2 ? 1 : 0};
    const int smem = g.stages * stage_bytes + 256;
    const int threads = 32 * (g.strips * g.RS + 1);
    const int grid = dw_grid(g, minb);
    const int addm = (add_full != nullptr ? 1 : 0) | (add_half != nullptr ? 2 : 0);
    const int variant = ((scale != nullptr) ? 8 : 0) | (relu ? 4 : 0) | addm;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_BWD(A, R, M)                                                                                                 \
    case ((A ? 8 : 0) | (R ? 4 : 0) | M): {                                                                                 \
        if (minb == 1) {                                                                                                    \
            XCP_CUDA(cudaFuncSetAttribute(dw3x3_bwd_kernel<A, R, M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            dw3x3_bwd_kernel<A, R, M, 1><<<grid, threads, smem, st>>>(tmG, tmX, tmF, p);                                    \
        } else {                                                                                                            \
            XCP_CUDA(cudaFuncSetAttribute(dw3x3_bwd_kernel<A, R, M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            dw3x3_bwd_kernel<A, R, M, 2><<<grid, threads, smem, st>>>(tmG, tmX, tmF, p);                                    \
        }                                                                                                                   \
    } break;
    switch (variant) {
        LAUNCH_BWD(false, false, 0) LAUNCH_BWD(false, false, 1) LAUNCH_BWD(false, false, 2) LAUNCH_BWD(false, false, 3)
        LAUNCH_BWD(false, true, 0) LAUNCH_BWD(false, true, 1) LAUNCH_BWD(false, true, 2) LAUNCH_BWD(false, true, 3)
        LAUNCH_BWD(true, false, 0) LAUNCH_BWD(true, false, 1) LAUNCH_BWD(true, false, 2) LAUNCH_BWD(true, false, 3)
        LAUNCH_BWD(true, true, 0) LAUNCH_BWD(true, true, 1) LAUNCH_BWD(true, true, 2) LAUNCH_BWD(true, true, 3)
        default: break;
    }
#undef LAUNCH_BWD
    return check_cuda(cudaGetLastError(), "dw3x3_bwd launch");
}
