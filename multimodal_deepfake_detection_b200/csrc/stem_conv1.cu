// Stem conv1 (Xception.py:118,168: nn.Conv2d(3, 32, 3, stride 2, pad 0) on the input frames) forward and weight gradient.
//
// K = 27 taps: far too thin for the tcgen05 pipeline (an im2col operand would cost as many bytes as the output), and on
// plain FFMA the first version was instruction bound (thread = pixel x 32 channels, 27 x 32 FFMA + 27 x 8 LDS per pixel:
// 451 us forward / 360 us weight gradient per 256 frames against a ~100 us HBM floor, profiles r2m).  Both kernels now run
// the reduction on warp-level tensor-core MMAs with the operands built straight from the staged input rows:
//   forward : mma.m16n8k8 tf32 (the fp32 frames are rounded to tf32 once while they are staged; weights tf32, fp32 accumulate)
//             D[16 pixels x 32 ch] = patch[16 x 32(27)] * W^T: 16 LDS + 16 MMA per 16 pixels instead of ~1100 instructions;
//   wgrad   : mma.m16n8k16 bf16, D[32(27) taps x 32 ch] += patch^T[taps x 16 pixels] * dY[16 pixels x 32 ch]; dY comes out
//             of shared memory with ldmatrix.trans (the [pixel][channel] rows are exactly the transposed B operand), patches are
//             rounded to bf16 like the im2col matrix of the first version.
// A CTA stages the 2n+1 input rows (x 3 channels) of a group of n <= 4 output rows in shared memory; every input byte is read
// from HBM once per group (the one row two neighbouring groups share comes from L2).  Input formats: fp32 NCHW in [0,1]
// (video_dataloader.py:35) or raw uint8 NHWC frames scaled by 1/255 while staging.
#include "common.cuh"

namespace xcp {

constexpr int S1_ROWS = 4;                       // output rows per group
constexpr int S1_NIN = 2 * S1_ROWS + 1;          // staged input rows per channel

struct StemGeom {
    int F, H, W, H1, W1;
    int pitch;                                   // floats per staged input row
    int gpf;                                     // groups per frame
    unsigned rcp_w1;                             // ceil(2^32 / W1): p / W1 == __umulhi(p, rcp_w1) for p < 2^16
    long long n_groups;
};

XCP_DEVINL uint32_t f32_to_tf32(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return r; }

// stage the 3 x S1_NIN input rows of the group starting at output row ho0 as fp32 (CVT_TF32: rounded to tf32) into
// s_in[ic][r][pitch] (rows past the image repeat the last one: only groups with fewer output rows reach them, unused).
// Flat index over (row, column) with 8 independent loads in flight per thread: the first form walked the 27 rows one after the
// other (two dependent global-load latencies per row) and made the whole kernel latency bound (1024 us for 256 frames).
template <bool U8, bool CVT_TF32>
XCP_DEVINL void stem_stage(const void* __restrict__ x, float* s_in, const StemGeom& g, long long f, int ho0) {
    const int total = 3 * S1_NIN * g.W;
    int idx = threadIdx.x;
    int q = idx / g.W, c = idx - q * g.W;
    const int step = blockDim.x;
    while (idx < total) {
        float v[8];
        int dst[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            dst[u] = -1;
            if (idx < total) {
                const int ic = q / S1_NIN, r = q - ic * S1_NIN;
                const int h = min(2 * ho0 + r, g.H - 1);
                if (U8) v[u] = (float)__ldg(reinterpret_cast<const uint8_t*>(x) + (((long long)f * g.H + h) * g.W + c) * 3 + ic) * (1.f / 255.f);
                else v[u] = __ldg(reinterpret_cast<const float*>(x) + (((long long)f * 3 + ic) * g.H + h) * g.W + c);
                dst[u] = q * g.pitch + c;
            }
            idx += step; c += step;
            while (c >= g.W) { c -= g.W; ++q; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (dst[u] >= 0) s_in[dst[u]] = CVT_TF32 ? __uint_as_float(f32_to_tf32(v[u])) : v[u];
    }
}

// fp32 NCHW frames: the same rows with 4-byte cp.async (the 299-float rows are not 16-byte aligned): no registers, no waiting --
// the whole group is in flight at once and the caller overlaps it with the previous group's MMAs.  One row at a time so that an
// element costs an add and a compare instead of index arithmetic.
XCP_DEVINL void stem_stage_async(const float* __restrict__ x, float* s_in, const StemGeom& g, long long f, int ho0) {
    const uint32_t s0 = smem_u32(s_in);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll 1
    for (int q = warp; q < 3 * S1_NIN; q += nwarps) {          // a warp per staged row: the row's base is computed once per ~10 copies
        const int ic = q / S1_NIN, r = q - ic * S1_NIN;
        const int h = min(2 * ho0 + r, g.H - 1);
        const float* src = x + (((long long)f * 3 + ic) * g.H + h) * g.W + lane;
        uint32_t dst = s0 + (uint32_t)(q * g.pitch + lane) * 4u;
        for (int c = lane; c < g.W; c += 32) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            src += 32; dst += 128;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
XCP_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// offset (floats) of tap k = (ic*3 + kh)*3 + kw relative to a pixel's base (2*row_local*pitch + 2*wo); taps >= 27 -> 0
XCP_DEVINL int stem_tap_off(int k, int pitch) {
    if (k >= 27) return 0;
    const int ic = k / 9, kh = (k % 9) / 3, kw = k % 3;
    return (ic * S1_NIN + kh) * pitch + kw;
}

// ---------------------------------------------------------------------------------------------------------------- forward
template <bool U8>
__global__ void __launch_bounds__(256, 2)
stem_conv1_fwd_mma_kernel(const void* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
                          float* __restrict__ partials, const float* __restrict__ scale, const float* __restrict__ shift,
                          const StemGeom g) {
    extern __shared__ float s_dyn[];
    float* s_in = s_dyn;                                                  // [3][S1_NIN][pitch]
    uint8_t* s_out = reinterpret_cast<uint8_t*>(s_in + (U8 ? 1 : 2) * 3 * S1_NIN * g.pitch);   // [8 warps][16 rows x 80 B]
    __shared__ float s_stat[8][2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t = lane & 3;

    // B fragments (weights, tf32): b[s][j] = (W[8j+gq][8s+t], W[8j+gq][8s+t+4]), W[n][k] = w[n*27 + k]
    uint32_t b0[4][4], b1[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = 8 * j + gq, k0 = 8 * s + t, k1 = k0 + 4;
            const float fold = scale ? scale[n] : 1.f;     // inference: eval-mode bn1 scale folded into the filter rows
            b0[s][j] = f32_to_tf32(k0 < 27 ? w[n * 27 + k0] * fold : 0.f);
            b1[s][j] = f32_to_tf32(k1 < 27 ? w[n * 27 + k1] * fold : 0.f);
        }
    float bsh[4][2];                                  // inference epilogue: + bn1 shift, ReLU (channels 8j + 2t + {0,1})
#pragma unroll
    for (int j = 0; j < 4; ++j) { bsh[j][0] = shift ? shift[8 * j + 2 * t] : 0.f; bsh[j][1] = shift ? shift[8 * j + 2 * t + 1] : 0.f; }
    int off0[4], off1[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { off0[s] = 4 * stem_tap_off(8 * s + t, g.pitch); off1[s] = 4 * stem_tap_off(8 * s + t + 4, g.pitch); }   // bytes

    float st1[4][2], st2[4][2];                       // per-thread column statistics: channels 8j + 2t + {0,1}
#pragma unroll
    for (int j = 0; j < 4; ++j) { st1[j][0] = st1[j][1] = st2[j][0] = st2[j][1] = 0.f; }
    const uint32_t stg = smem_u32(s_out) + (uint32_t)warp * 1280u;

    // fp32 frames: two staging buffers; group i+1 streams in (cp.async) while the MMAs of group i run.  uint8 frames: one
    // buffer, staged through registers (byte loads cannot be cp.async'ed into floats).
    const int buf_floats = 3 * S1_NIN * g.pitch;
    int it = 0;
    if (!U8 && (long long)blockIdx.x < g.n_groups) {
        const long long f0 = blockIdx.x / g.gpf;
        stem_stage_async(reinterpret_cast<const float*>(x), s_in, g, f0, (int)(blockIdx.x - f0 * g.gpf) * S1_ROWS);
    }
    for (long long grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x, ++it) {
        const long long f = grp / g.gpf;
        const int ho0 = (int)(grp - f * g.gpf) * S1_ROWS;
        const int nrows = min(S1_ROWS, g.H1 - ho0);
        const int npix = nrows * g.W1;
        float* s_cur = s_in;
        if (U8) {
            __syncthreads();                          // the previous group's tiles are done with s_in
            stem_stage<true, false>(x, s_in, g, f, ho0);
            __syncthreads();
        } else {
            s_cur = s_in + (it & 1) * buf_floats;
            cp_async_wait_all();
            __syncthreads();                          // this group's rows have landed; everyone is done with the other buffer
            const long long nxt = grp + gridDim.x;
            if (nxt < g.n_groups) {
                const long long fn = nxt / g.gpf;
                stem_stage_async(reinterpret_cast<const float*>(x), s_in + ((it + 1) & 1) * buf_floats, g, fn, (int)(nxt - fn * g.gpf) * S1_ROWS);
            }
        }
        const long long gp0 = ((long long)f * g.H1 + ho0) * g.W1;
        for (int tile = warp; tile * 16 < npix; tile += 8) {
            const int p0 = tile * 16 + gq, p1 = p0 + 8;
            const bool v0 = p0 < npix, v1 = p1 < npix;
            const int q0 = v0 ? p0 : npix - 1, q1 = v1 ? p1 : npix - 1;
            const int r0 = (int)__umulhi((unsigned)q0, g.rcp_w1), r1 = (int)__umulhi((unsigned)q1, g.rcp_w1);
            const uint32_t cur_s = smem_u32(s_cur);
            const uint32_t a0b = cur_s + (uint32_t)(2 * r0 * g.pitch + 2 * (q0 - r0 * g.W1)) * 4u;
            const uint32_t a1b = cur_s + (uint32_t)(2 * r1 * g.pitch + 2 * (q1 - r1 * g.W1)) * 4u;
            float c[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f; }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                // round (not truncate) to tf32: the MMA drops the low 13 mantissa bits, so adding half an ulp of tf32 to the bit
                // pattern is round-to-nearest (frames are finite, in [0,1]: no exponent overflow to guard)
                uint32_t a0, a1, a2, a3;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(a0b + off0[s]));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(a1b + off0[s]));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a2) : "r"(a0b + off1[s]));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a3) : "r"(a1b + off1[s]));
                a0 += 0x1000u; a1 += 0x1000u; a2 += 0x1000u; a3 += 0x1000u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0[s][j]), "r"(b1[s][j]));
            }
            // statistics on the fp32 accumulators (rows past the group: masked), bf16 rows through a padded staging tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (shift != nullptr) {               // y = relu(bn1(conv1(x))) written directly, no statistics
                    c[j][0] = fmaxf(c[j][0] + bsh[j][0], 0.f); c[j][1] = fmaxf(c[j][1] + bsh[j][1], 0.f);
                    c[j][2] = fmaxf(c[j][2] + bsh[j][0], 0.f); c[j][3] = fmaxf(c[j][3] + bsh[j][1], 0.f);
                }
                if (v0) { st1[j][0] += c[j][0]; st1[j][1] += c[j][1]; st2[j][0] = fmaf(c[j][0], c[j][0], st2[j][0]); st2[j][1] = fmaf(c[j][1], c[j][1], st2[j][1]); }
                if (v1) { st1[j][0] += c[j][2]; st1[j][1] += c[j][3]; st2[j][0] = fmaf(c[j][2], c[j][2], st2[j][0]); st2[j][1] = fmaf(c[j][3], c[j][3], st2[j][1]); }
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(stg + gq * 80 + j * 16 + t * 4), "r"(pack_bf16(c[j][0], c[j][1])) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(stg + (gq + 8) * 80 + j * 16 + t * 4), "r"(pack_bf16(c[j][2], c[j][3])) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = lane + 32 * h, row = ch >> 2, part = ch & 3;
                if (tile * 16 + row < npix) {
                    uint4 v;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + row * 80 + part * 16));
                    *reinterpret_cast<uint4*>(y + (gp0 + tile * 16 + row) * 32 + part * 8) = v;
                }
            }
            __syncwarp();
        }
    }
    // column statistics: lanes with the same t hold partial sums of the same 8 channels
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float a = st1[j][e], b = st2[j][e];
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (gq == 0) { s_stat[warp][0][8 * j + 2 * t + e] = a; s_stat[warp][1][8 * j + 2 * t + e] = b; }
        }
    __syncthreads();
    if (threadIdx.x < 64 && partials != nullptr) {
        const int st = threadIdx.x >> 5, chn = threadIdx.x & 31;
        float a = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) a += s_stat[wq][st][chn];
        partials[((long long)blockIdx.x * 2 + st) * 32 + chn] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------- weight gradient
// D[tap][oc] += sum_pix patch[pix][tap] * dY[pix][oc]; per CTA over its groups, one atomicAdd pass at the end.
template <bool U8>
__global__ void __launch_bounds__(256, 2)
stem_conv1_wgrad_mma_kernel(const void* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dW, const StemGeom g) {
    extern __shared__ float s_dyn[];
    float* s_in = s_dyn;                                                          // [3][S1_NIN][pitch] fp32
    const int max_pix = S1_ROWS * g.W1;
    const int pix_pad = (max_pix + 15) & ~15;
    uint8_t* s_dy = reinterpret_cast<uint8_t*>(s_in + 3 * S1_NIN * g.pitch);       // [pix_pad][80 B]: 32 bf16 channels + pad
    __shared__ float s_acc[32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;

    // A = patch^T [tap (2 m-tiles of 16) x pixel (16)]: a0 = (tap gq, pix 2t..2t+1), a1 = (tap gq+8, same), a2 = (tap gq, pix 2t+8..9),
    // a3 = (tap gq+8, pix 2t+8..9); tap offsets of this thread's four rows
    int toff[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m) { toff[m][0] = stem_tap_off(16 * m + gq, g.pitch); toff[m][1] = stem_tap_off(16 * m + gq + 8, g.pitch); }
    const bool tap_ok[2][2] = {{gq < 27, gq + 8 < 27}, {16 + gq < 27, 24 + gq < 27}};
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f; }
    const uint32_t dy_s = smem_u32(s_dy);

    for (long long grp = blockIdx.x; grp < g.n_groups; grp += gridDim.x) {
        const long long f = grp / g.gpf;
        const int ho0 = (int)(grp - f * g.gpf) * S1_ROWS;
        const int nrows = min(S1_ROWS, g.H1 - ho0);
        const int npix = nrows * g.W1;
        const int npad = (npix + 15) & ~15;
        __syncthreads();
        const long long gp0 = ((long long)f * g.H1 + ho0) * g.W1;
        for (int i = threadIdx.x; i < npad * 4; i += blockDim.x) {               // dY rows (64 B each) -> 80 B pitch, zero past the group
            const int p = i >> 2, part = i & 3;
            const uint32_t dst = dy_s + p * 80 + part * 16;
            if (p < npix) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(dy + (gp0 + p) * 32 + part * 8) : "memory");
            else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
        if (U8) stem_stage<true, false>(x, s_in, g, f, ho0);
        else stem_stage_async(reinterpret_cast<const float*>(x), s_in, g, f, ho0);
        asm volatile("cp.async.commit_group;" ::: "memory");
        cp_async_wait_all();
        __syncthreads();
        for (int tile = warp; tile * 16 < npix; tile += 8) {
            // B = dY [16 pixels (k) x 32 channels (n)] row-major in shared memory -> col-major fragments with ldmatrix.trans:
            // matrices (k half h, n tile j): lane l of an x4 load supplies the address of row (l & 7) of matrix (l >> 3)
            uint32_t bf[4][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int mat = lane >> 3, rr = lane & 7;               // mat = 2 * (j - 2jj) + h
                const int j = 2 * jj + (mat >> 1), h = mat & 1;
                const uint32_t addr = dy_s + (uint32_t)(tile * 16 + 8 * h + rr) * 80u + (uint32_t)j * 16u;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bf[2 * jj][0]), "=r"(bf[2 * jj][1]), "=r"(bf[2 * jj + 1][0]), "=r"(bf[2 * jj + 1][1]) : "r"(addr));
            }
            // patches: four pixels per thread (2t, 2t+1, 2t+8, 2t+9), clamped inside the group (their dY rows are zero)
            const float* pb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int p = tile * 16 + 2 * t + (e & 1) + 8 * (e >> 1);
                p = min(p, npix - 1);
                const int r = (int)__umulhi((unsigned)p, g.rcp_w1);
                pb[e] = s_in + 2 * r * g.pitch + 2 * (p - r * g.W1);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                uint32_t a[4];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)                           // hh = pixel half (2t.. / 2t+8..)
#pragma unroll
                    for (int rw = 0; rw < 2; ++rw) {                     // rw = tap row gq / gq + 8
                        const int o = toff[m][rw];
                        const float lo = tap_ok[m][rw] ? pb[2 * hh][o] : 0.f, hi = tap_ok[m][rw] ? pb[2 * hh + 1][o] : 0.f;
                        a[2 * hh + rw] = pack_bf16(lo, hi);
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(acc[m][j][0]), "+f"(acc[m][j][1]), "+f"(acc[m][j][2]), "+f"(acc[m][j][3])
                                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bf[j][0]), "r"(bf[j][1]));
            }
        }
    }
    // D fragment: (tap 16m + gq [+8], channel 8j + 2t [+1]); CTA reduction in shared memory, then one RED per (oc, tap)
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&s_acc[16 * m + gq][8 * j + 2 * t], acc[m][j][0]);
            atomicAdd(&s_acc[16 * m + gq][8 * j + 2 * t + 1], acc[m][j][1]);
            atomicAdd(&s_acc[16 * m + gq + 8][8 * j + 2 * t], acc[m][j][2]);
            atomicAdd(&s_acc[16 * m + gq + 8][8 * j + 2 * t + 1], acc[m][j][3]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
        const int tap = i >> 5, oc = i & 31;
        atomicAdd(&dW[oc * 27 + tap], s_acc[tap][oc]);
    }
}

static StemGeom stem_geom(int F, int H, int W) {
    StemGeom g;
    g.F = F; g.H = H; g.W = W;
    g.H1 = (H - 3) / 2 + 1; g.W1 = (W - 3) / 2 + 1;
    g.pitch = (W + 4) & ~3;
    g.gpf = (g.H1 + S1_ROWS - 1) / S1_ROWS;
    g.rcp_w1 = (unsigned)(((1ULL << 32) + (unsigned)g.W1 - 1) / (unsigned)g.W1);
    g.n_groups = (long long)F * g.gpf;
    return g;
}
static int stem_grid(const StemGeom& g) {
    const long long cap = 2LL * num_sms();
    return (int)(g.n_groups < cap ? g.n_groups : cap);
}

}  // namespace xcp

using namespace xcp;

#define ST ((cudaStream_t)stream)

// number of partial rows xcp_stem_conv1_fwd writes ( = its grid size)
extern "C" int xcp_stem_conv1_parts(int F, int H, int W, int device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    return stem_grid(stem_geom(F, H, W));
}

static int stem_conv1_launch(const void* x, int x_u8_nhwc, const float* w, void* y, float* partials, const float* scale,
                             const float* shift, int F, int H, int W, int device, void* stream) {
    XCP_REQUIRE(F > 0 && H >= 3 && W >= 3, "xcp_stem_conv1_fwd: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const StemGeom g = stem_geom(F, H, W);
    const int smem = (x_u8_nhwc ? 1 : 2) * 3 * S1_NIN * g.pitch * 4 + 8 * 1280;
    XCP_REQUIRE(smem <= 200 * 1024, "xcp_stem_conv1_fwd: frames wider than %d pixels are not supported", (200 * 1024 - 8 * 1280) / (3 * S1_NIN * 4) - 4);
    const int grid = stem_grid(g);
    if (x_u8_nhwc) {
        XCP_CUDA(cudaFuncSetAttribute(stem_conv1_fwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        stem_conv1_fwd_mma_kernel<true><<<grid, 256, smem, ST>>>(x, w, (__nv_bfloat16*)y, partials, scale, shift, g);
    } else {
        XCP_CUDA(cudaFuncSetAttribute(stem_conv1_fwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        stem_conv1_fwd_mma_kernel<false><<<grid, 256, smem, ST>>>(x, w, (__nv_bfloat16*)y, partials, scale, shift, g);
    }
    return check_cuda(cudaGetLastError(), "stem_conv1_fwd launch");
}

extern "C" int xcp_stem_conv1_fwd(const void* x, int x_u8_nhwc, const float* w, void* y, float* partials, int F, int H, int W,
                                  int device, void* stream) {
    XCP_REQUIRE(partials != nullptr, "xcp_stem_conv1_fwd: partials");
    return stem_conv1_launch(x, x_u8_nhwc, w, y, partials, nullptr, nullptr, F, H, W, device, stream);
}

// inference form: out = relu(scale * conv1(x) + shift) with eval-mode bn1 folded (scale into the filter rows, shift + ReLU in the
// epilogue): relu(bn1(conv1(x))) of Xception.py:168-170 in one pass, no raw conv output, no statistics
extern "C" int xcp_stem_conv1_fwd_affine(const void* x, int x_u8_nhwc, const float* w, const float* scale, const float* shift, void* out,
                                         int F, int H, int W, int device, void* stream) {
    XCP_REQUIRE(scale != nullptr && shift != nullptr, "xcp_stem_conv1_fwd_affine: scale / shift");
    return stem_conv1_launch(x, x_u8_nhwc, w, out, nullptr, scale, shift, F, H, W, device, stream);
}

// (the first version needed an im2col scratch matrix; kept in the ABI, a token size is enough now)
extern "C" long long xcp_stem_conv1_wgrad_ws_bytes(int F, int H, int W) {
    (void)F; (void)H; (void)W;
    return 256;
}

extern "C" int xcp_stem_conv1_wgrad(const void* x, int x_u8_nhwc, const void* dy, float* dW, void* workspace, int F, int H, int W,
                                    int device, void* stream) {
    (void)workspace;
    XCP_REQUIRE(F > 0 && H >= 3 && W >= 3, "xcp_stem_conv1_wgrad: bad shape");
    XCP_CUDA(cudaSetDevice(device));
    const StemGeom g = stem_geom(F, H, W);
    const int pix_pad = (S1_ROWS * g.W1 + 15) & ~15;
    const int smem = 3 * S1_NIN * g.pitch * 4 + pix_pad * 80;
    XCP_REQUIRE(smem <= 200 * 1024, "xcp_stem_conv1_wgrad: frames too wide for the staged rows (%d bytes of shared memory)", smem);
    const int grid = stem_grid(g);
    if (x_u8_nhwc) {
        XCP_CUDA(cudaFuncSetAttribute(stem_conv1_wgrad_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        stem_conv1_wgrad_mma_kernel<true><<<grid, 256, smem, ST>>>(x, (const __nv_bfloat16*)dy, dW, g);
    } else {
        XCP_CUDA(cudaFuncSetAttribute(stem_conv1_wgrad_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        stem_conv1_wgrad_mma_kernel<false><<<grid, 256, smem, ST>>>(x, (const __nv_bfloat16*)dy, dW, g);
    }
    return check_cuda(cudaGetLastError(), "stem_conv1_wgrad launch");
}
