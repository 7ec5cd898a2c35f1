// Depthwise 3x3 forward / backward, "row stream" kernels for narrow images (19x19, 10x10 and column strips of wider ones).
//
// Same operator as dw.cu (nn.Conv2d(groups=C) of SeparableConv2d.conv1, Xception.py:41,45, with the preceding ReLU and
// the pending BatchNorm affine of the producer fused in front, Xception.py:61-78), different decomposition.  The tile
// kernels of dw.cu pay per-tile fixed costs on small maps (halo, warm-up rows of the register window, 4-column strips of
// a 19-wide image) and the register-window kernels that replaced them at 19x19 / 10x10 move every byte once but have only
// one image row per thread in flight (8 warps per SM x 19 x 128 B = 19 KB per SM against the ~35 KB Little's law asks for
// at 6.4 TB/s), which left them at 0.53 of the HBM roofline.  Here:
//
//   * work item   = (frame, 64-channel tile, column strip of <= WS pixels); its H rows plus ONE virtual zero row form a
//                   virtual stream of H+1 rows.  Items follow each other, so the zero row is the bottom padding of one image
//                   and the top padding of the next.  The concatenated stream of all items is cut into equal contiguous
//                   ranges, one per warp: perfectly balanced statically, one warm-up row above and below each range.
//   * each warp   owns a private ring of NSLOTS row slots in shared memory and its own mbarriers.  Lane 0 requests rows with
//                   TMA (box = 64 channels x (WS+2) pixels x 1 row, out-of-image columns zero-filled = the conv padding)
//                   NSLOTS rows ahead; there is no inter-warp synchronisation at all.  Bytes in flight per SM = 8 warps x
//                   8 rows x 2.7 KB = 170 KB, independent of the register budget.  The tensor map uses 128-byte L2 promotion:
//                   with the default 256 B every 64-channel box row dragged in the neighbouring channel tile, which this
//                   traversal consumes much later (ncu r2k: 217 MB read for 142 MB algorithmic, 75.8 vs 63.5 us).
//   * lane        = one pair of adjacent channels (packed f32x2 math); the warp keeps a 3-row x (WS+2)-column window of the
//                   activated input (forward) / of dD (backward) in registers and slides it down the stream: every staged
//                   element is read from shared memory once, converted once and used for 9 (18) FFMA2.
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace xcp {

struct DwsParams {
    int F, H, W, C;                    // C = channel pitch
    int c_tiles, n_strips, base_w, rem_w;   // strip s covers base_w (+1 if s < rem_w) columns
    int n_items;
    int fblock;                        // frames per block of the item order (block, channel tile, frame in block, strip)
    long long total_v;                 // n_items * (H + 1)
    long long per_warp;                // virtual rows per warp
    const float* w9;                   // [9][C]
    const float* scale;                // AFFINE
    const float* shift;
    void* out;                         // forward: d [F,H,W,C]; backward: dz
    const void* add_half;              // unused here
    float* dw;                         // backward: [c_real][9]
    float* bnsum;                      // backward, AFFINE: [2][C]
    int c_real;
};

struct DwsCursor {
    int item, y;           // y == H: the virtual zero row (also: before the first / after the last item)
    int f, ct, x0, ncols;
};

// item -> (frame, channel tile, first column, columns); out of line: runs once per item and contains two integer divisions
__device__ __noinline__ int4 dws_decode(int item, int n_strips, int c_tiles, int base_w, int rem_w, int fblock) {
    const int strip = item % n_strips;
    const int t = item / n_strips;
    const int fi = t % fblock, t2 = t / fblock;
    return make_int4((t2 / c_tiles) * fblock + fi, t2 % c_tiles, strip * base_w + min(strip, rem_w), base_w + (strip < rem_w ? 1 : 0));
}
XCP_DEVINL void dws_enter(DwsCursor& c, const DwsParams& p) {
    if (c.item < 0 || c.item >= p.n_items) { c.y = p.H; c.f = c.ct = c.x0 = c.ncols = 0; return; }
    const int4 d = dws_decode(c.item, p.n_strips, p.c_tiles, p.base_w, p.rem_w, p.fblock);
    c.f = d.x; c.ct = d.y; c.x0 = d.z; c.ncols = d.w;
}
XCP_DEVINL void dws_seek(DwsCursor& c, long long v, const DwsParams& p) {
    if (v < 0) { c.item = -1; c.y = p.H; }
    else { c.item = (int)(v / (p.H + 1)); c.y = (int)(v % (p.H + 1)); }
    dws_enter(c, p);
}
XCP_DEVINL void dws_advance(DwsCursor& c, const DwsParams& p) {
    if (++c.y > p.H) { c.y = 0; ++c.item; dws_enter(c, p); }
}

XCP_DEVINL uint32_t dws_lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
XCP_DEVINL u64 dws_unpack(uint32_t v) { return pk2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
__device__ __noinline__ void dws_wait_slow(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 26)) { printf("xcp: dws mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
    }
}
XCP_DEVINL void dws_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) dws_wait_slow(bar, parity);
}

// ---------------------------------------------------------------------------------------------------------------- forward
template <int WS, int CP, bool AFFINE, bool RELU, int NWARPS, int NSLOTS>
__global__ void __maxnreg__(((65536 / (NWARPS * 32)) / 8) * 8 > 255 ? 255 : ((65536 / (NWARPS * 32)) / 8) * 8)
dws_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const DwsParams p) {
    constexpr int LW = WS + 2;
    constexpr uint32_t ROWB = LW * 128u;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full[NWARPS][NSLOTS];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t sbase = (raw_addr + 127u) & ~127u;         // [0, ROWB): a row of zeros; then the per-warp rings
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        for (int i = 0; i < NSLOTS; ++i) mbar_init(&full[warp][i], 1);
        fence_barrier_init();
        if (warp == 0) tma_prefetch_desc(&tmX);
    }
    for (uint32_t i = threadIdx.x; i < ROWB / 4; i += blockDim.x) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + i * 4), "r"(0u) : "memory");
    __syncthreads();

    const long long gw = (long long)blockIdx.x * NWARPS + warp;
    const long long v0 = gw * p.per_warp;
    const long long v1 = min(v0 + p.per_warp, p.total_v);
    if (v0 >= v1) return;
    const uint32_t wbase = sbase + ROWB + (uint32_t)warp * NSLOTS * ROWB;
    const uint32_t bar0 = smem_u32(&full[warp][0]);
    const int gH = p.H, gW = p.W;

    // producer cursor: rows v0-1 .. v1 (inclusive), zero rows skipped
    DwsCursor pc;
    dws_seek(pc, v0 - 1, p);
    int pleft = (int)(v1 - v0 + 2);          // virtual rows the producer still has to visit
    int pslot = 0;
    auto produce = [&]() {                   // request the next real row (if any) into pslot
        while (pleft > 0) {
            const bool real = pc.y < gH;
            if (real && lane == 0) {
                const uint32_t bar = bar0 + (uint32_t)pslot * 8u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROWB) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(wbase + (uint32_t)pslot * ROWB), "l"(&tmX), "r"(bar), "r"(pc.ct * 64), "r"(pc.x0 - 1), "r"(pc.y), "r"(pc.f)
                    : "memory");
            }
            dws_advance(pc, p); --pleft;
            if (real) { if (++pslot == NSLOTS) pslot = 0; return; }
        }
    };
    for (int i = 0; i < NSLOTS; ++i) produce();

    // channel-tile constants are reloaded when the stream enters a row of another channel tile
    u64 wk[9], sc = pk2(1.f, 1.f), sh = 0ull;
    int cur_ct = -1;
    auto load_consts = [&](int ct) {
        const int c0 = ct * 64 + lane * 2;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const float2 t = *reinterpret_cast<const float2*>(p.w9 + (long long)k * p.C + c0); wk[k] = pk2(t.x, t.y); }
        if (AFFINE) {
            const float2 a = *reinterpret_cast<const float2*>(p.scale + c0), b = *reinterpret_cast<const float2*>(p.shift + c0);
            sc = pk2(a.x, a.y); sh = pk2(b.x, b.y);
        }
        cur_ct = ct;
    };

    DwsCursor lc, ec;
    dws_seek(lc, v0 - 1, p);
    ec = lc; ec.y = gH;                      // "no centre row yet"
    int cslot = 0; uint32_t cph = 0;
    int free_slot = -1;                      // slot of the previous loaded row: released (refilled) at the end of the next step
    int nleft = (int)(v1 - v0 + 2);          // steps to go
    bool warm = true;                        // the first loaded row is the warm-up row above the range: never a centre
    u64 w0[LW], w1[LW], w2[LW];
    uint32_t* const out32 = reinterpret_cast<uint32_t*>(p.out);

    // load + activate the row under the load cursor; a virtual zero row reads the zero row of shared memory with a zeroed affine.
    // (Fetching the raw values one step ahead, so that the shared-memory latency hides under the previous row's FFMA2s, was
    // measured slower: 67.6 vs 63.5 us at 19x19x768 -- 233 instead of 206 registers, gpurun r2m.)
    auto load_row = [&](u64 (&dst)[LW]) -> int {
        const bool real = lc.y < gH;
        int slot = -1;
        uint32_t a = sbase + (uint32_t)lane * 4u;
        if (real) {
            slot = cslot;
            dws_wait(bar0 + (uint32_t)slot * 8u, cph);
            if (++cslot == NSLOTS) { cslot = 0; cph ^= 1; }
            a = wbase + (uint32_t)slot * ROWB + (uint32_t)lane * 4u;
            if (lc.ct != cur_ct) load_consts(lc.ct);
        }
        uint32_t raw[LW];
#pragma unroll
        for (int j = 0; j < LW; ++j) raw[j] = dws_lds32(a + j * 128);
        const u64 scr = real ? sc : 0ull, shr = real ? sh : 0ull;
        const bool left = lc.x0 == 0, right = lc.x0 + lc.ncols == gW;
        const int rc = lc.ncols + 1;
#pragma unroll
        for (int j = 0; j < LW; ++j) {
            u64 v = dws_unpack(raw[j]);
            if (AFFINE) v = fma2(v, scr, shr);
            if (RELU) { float lo, hi; upk2(v, lo, hi); v = pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
            if (AFFINE) {
                if (j == 0 && left) v = 0ull;
                if (j >= WS && j == rc && right) v = 0ull;        // ncols is WS or WS - 1
            }
            dst[j] = v;
        }
        return slot;
    };
    auto emit = [&](const u64 (&a)[LW], const u64 (&b)[LW], const u64 (&c)[LW]) {
        uint32_t* orow = out32 + ((((long long)ec.f * gH + ec.y) * gW + ec.x0) * CP + ec.ct * 64) / 2 + lane;
        const bool fullw = ec.ncols == WS;
#pragma unroll
        for (int i = 0; i < WS; ++i) {
            u64 acc = mul2(a[i], wk[0]);
            acc = fma2(a[i + 1], wk[1], acc); acc = fma2(a[i + 2], wk[2], acc);
            acc = fma2(b[i], wk[3], acc); acc = fma2(b[i + 1], wk[4], acc); acc = fma2(b[i + 2], wk[5], acc);
            acc = fma2(c[i], wk[6], acc); acc = fma2(c[i + 1], wk[7], acc); acc = fma2(c[i + 2], wk[8], acc);
            float lo, hi; upk2(acc, lo, hi);
            if (i < WS - 1 || fullw) orow[i * (CP / 2)] = pack_bf16(lo, hi);
        }
    };

#define DWS_STEP(WA, WB, WC)                                                             \
    {                                                                                    \
        const int sl = load_row(WC);                                                     \
        if (ec.y < gH) emit(WA, WB, WC);                                                 \
        __syncwarp();                                                                    \
        if (free_slot >= 0) produce();                                                   \
        free_slot = sl;                                                                  \
        ec = lc;                                                                         \
        if (warm) { ec.y = gH; warm = false; }                                           \
        dws_advance(lc, p);                                                              \
        if (--nleft == 0) break;                                                         \
    }
    while (true) {
        DWS_STEP(w0, w1, w2)
        DWS_STEP(w1, w2, w0)
        DWS_STEP(w2, w0, w1)
    }
#undef DWS_STEP
}

// ---------------------------------------------------------------------------------------------------------------- host
static int dws_make_params(DwsParams& p, int F, int H, int W, int C, int WS, int nwarps, int* grid) {
    p.F = F; p.H = H; p.W = W; p.C = C;
    p.c_tiles = C / 64;
    p.n_strips = (W + WS - 1) / WS;
    p.base_w = W / p.n_strips;
    p.rem_w = W % p.n_strips;
    p.n_items = F * p.c_tiles * p.n_strips;
    p.fblock = 1;
    p.total_v = (long long)p.n_items * (H + 1);
    const long long max_warps = (long long)num_sms() * nwarps;
    long long per = (p.total_v + max_warps - 1) / max_warps;
    if (per < 12) per = 12;
    p.per_warp = per;
    const long long warps = (p.total_v + per - 1) / per;
    *grid = (int)((warps + nwarps - 1) / nwarps);
    return 0;
}

static int dws_tmap(CUtensorMap* m, const void* base, int F, int H, int W, int C, int box_w) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)F};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)box_w, 1, 1};
    int promo = 128;
    { const char* e = getenv("XCP_DWS_L2PROMO"); if (e) promo = atoi(e); }             // tuning hook
    return make_tmap_4d_l2(m, base, dims, strides, box, 0, promo);
}

template <int WS, int CP, int NWARPS, int NSLOTS>
static int dws_launch_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H,
                          int W, cudaStream_t st) {
    DwsParams p{};
    int grid = 0;
    dws_make_params(p, F, H, W, CP, WS, NWARPS, &grid);
    p.w9 = w9; p.scale = scale; p.shift = shift; p.out = out;
    CUtensorMap tm;
    if (int e = dws_tmap(&tm, x, F, H, W, CP, WS + 2)) return e;
    const int smem = (NWARPS * NSLOTS + 1) * (WS + 2) * 128 + 128;
#define DWS_LAUNCH(A, R)                                                                                                        \
    {                                                                                                                           \
        auto k = dws_fwd_kernel<WS, CP, A, R, NWARPS, NSLOTS>;                                                                  \
        XCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                                   \
        k<<<grid, NWARPS * 32, smem, st>>>(tm, p);                                                                              \
    }
    if (scale != nullptr) { if (relu) DWS_LAUNCH(true, true) else DWS_LAUNCH(true, false) }
    else { if (relu) DWS_LAUNCH(false, true) else DWS_LAUNCH(false, false) }
#undef DWS_LAUNCH
    return check_cuda(cudaGetLastError(), "dws_fwd launch");
}

// *handled = 0 when the shape is not instantiated here (the caller falls through to the other kernels); otherwise the
// return value is the launch status
int dws_try_fwd(const void* x, const float* w9, const float* scale, const float* shift, int relu, void* out, int F, int H, int W,
                int C, cudaStream_t st, int* handled) {
    *handled = 0;
    const char* e = getenv("XCP_DW_NO_STREAM");                                       // A/B hook
    if (e && e[0] == '1') return 0;
    *handled = 1;
    // measured against the kernels of dw.cu (tools/dw_stream_ab.py, 256 frames, gpurun r2m): 19x19x768 80 -> 63.5 us (0.69 of
    // HBM), 37x37x768 affine+relu 285 -> 245 us, 74x74x256 affine+relu 161 -> 150 us; ties or loses elsewhere (10x10: 37 vs 33 us,
    // 147x147x128: 146 vs 146 us, plain-ReLU 37x37 / 74x74: 251 vs 245, 144 vs 144 us), so only these are routed here
    if (W == 19 && C == 768) return dws_launch_fwd<19, 768, 8, 8>(x, w9, scale, shift, relu, out, F, H, W, st);
    if (W == 37 && C == 768 && scale != nullptr && relu) return dws_launch_fwd<19, 768, 8, 8>(x, w9, scale, shift, relu, out, F, H, W, st);
    if (W == 74 && C == 256 && scale != nullptr && relu) return dws_launch_fwd<19, 256, 8, 8>(x, w9, scale, shift, relu, out, F, H, W, st);
    *handled = 0;
    return 0;
}

}  // namespace xcp
