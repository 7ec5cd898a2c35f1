// Weight gradient of the dense 3x3 stem convolution conv2 (Xception.py:122: 32 -> 64 channels, k3 s1 p0) on tcgen05.
//
//   gk[o][(kh*3 + kw)*32 + i] += sum_r dY[r][o] * X[r + kh*Wg + kw][i]        r over the F*Hg*Wg rows of the input grid
//
// (dY lives on the input grid, zero outside the valid output window, so the shifted pairing is exact.)  The first version
// ran this as a generic MN-major split-K GEMM with the 9 taps as N tiles: 128-row MMAs for 64 real rows, 64-column boxes for
// 32 real columns, a RED epilogue every 4096 grid rows and each tap's operand box fetched separately from L2 -- 1.14 ms per
// step at 256 frames, 0.15 of its HBM roofline (bench r2m).  This kernel:
//   * one persistent CTA per SM owns a contiguous range of the reduction (grid rows) and ALL nine taps: three TMEM accumulators
//     [64 x 96] (one per filter row kh; the 96 columns are kw x 32 input channels), a single RED epilogue per CTA;
//   * per 64-row K block the producer stages the dY rows once (64B x 128 box, 128B swizzle) and, per kh, ONE box of 72 X rows
//     (64B swizzle).  The three kw taps of a filter row are not separate loads: they are the same staged rows viewed one / two
//     rows further down, i.e. the B descriptor is MN-major with a leading-dimension byte offset of one row (64 B) between the
//     three 32-column atoms.  That is valid because the swizzle is a function of the absolute shared-memory address (the same
//     property the halo mode of the forward implicit GEMM relies on, gemm.cu) -- checked by test_conv3x3_implicit_gemm;
//   * the A operand (dY, MN-major) is padded from 64 to the 128 rows of the MMA by pointing its second 64-column atom at a
//     block of zeros in shared memory (descriptor LBO = distance to that block), so no zero bytes are fetched.
// Per K block and SM: 8 KB + 3 x 4.5 KB from L2 for 12 MMAs of 128 x 96 x 16 (576 clk): 38 B/clk, far below the L2->SM rate.
#include "common.cuh"
#include <stdlib.h>

namespace xcp {

struct ConvWgradParams {
    float* gk;            // [64][288] fp32, accumulated with RED
    int num_k_blocks;     // ceil(R / 64)
    int Wg;               // grid width: filter row kh reads X rows shifted by kh * Wg
};

constexpr int CW_DY_BYTES = 64 * 128;         // 64 K rows x 64 channels bf16
constexpr int CW_X_ROWS = 72;                 // 64 + 2 (kw shift), rounded up to the 8-row swizzle period
constexpr int CW_X_BYTES = CW_X_ROWS * 64;    // x 32 channels bf16
constexpr int CW_STAGE_BYTES = CW_DY_BYTES + 3 * CW_X_BYTES;

template <int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad32_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const ConvWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sDY = smem;                                  // [STAGES][8192]
    uint8_t* sX = sDY + STAGES * CW_DY_BYTES;             // [STAGES][3][4608]
    uint8_t* sZero = sX + STAGES * 3 * CW_X_BYTES;        // 8192 bytes of zeros (second MN atom of the A operand); 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(sZero + CW_DY_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
    }
    for (int i = threadIdx.x; i < CW_DY_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sZero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();                 // the zeros are read by the tensor core (async proxy)
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int per = (p.num_k_blocks + (int)gridDim.x - 1) / (int)gridDim.x;
    const int kb0 = (int)blockIdx.x * per;
    const int kb1 = min(kb0 + per, p.num_k_blocks);

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------ TMA producer
        int s = 0; uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], CW_STAGE_BYTES);
            const int r0 = kb * 64;
            tma_load_2d(sDY + s * CW_DY_BYTES, &tmDY, &full[s], 0, r0);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) tma_load_2d(sX + (s * 3 + kh) * CW_X_BYTES, &tmX, &full[s], 0, r0 + kh * p.Wg);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------ MMA issuer: D_kh[128 x 96] += dY^T[128 x 16] * Xview_kh[16 x 96]
        constexpr uint32_t idesc = make_idesc_bf16(128, 96, 1, 1);
        const uint32_t dy_base = smem_u32(sDY), x_base = smem_u32(sX), z_base = smem_u32(sZero);
        int s = 0; uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a0 = dy_base + s * CW_DY_BYTES;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // A: MN-major, 128B swizzle: atoms of 64 (MN) x 8 (K) at 1024 B per K group; atom 1 = the zero block
                    const uint64_t adesc = make_smem_desc(a0 + k * (16 * 128), z_base - a0, 1024, 2);
                    // B: MN-major, 64B swizzle: atoms of 32 (MN) x 8 (K) at 512 B per K group; atoms kw = 0,1,2 are the same rows
                    // shifted by kw rows (LBO = 64 B)
                    const uint64_t bdesc = make_smem_desc(x_base + (s * 3 + kh) * CW_X_BYTES + k * (16 * 64), 64, 512, 4);
                    umma_bf16(tmem_base + kh * 96, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
            }
            umma_commit(&empty[s]);
            if (kb == kb1 - 1) umma_commit(tmem_full);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp >= 2 && kb1 > kb0) {
        // ------------------------------------------------ epilogue (once): rows o = TMEM lanes 0..63, columns = gk columns
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        if (q < 2) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
            float* orow = p.gk + (long long)(q * 32 + lane) * 288;
#pragma unroll 1
            for (int c = 0; c < 9; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c * 32 + g * 4),
                                 "f"(__uint_as_float(r[g * 4 + 0])), "f"(__uint_as_float(r[g * 4 + 1])),
                                 "f"(__uint_as_float(r[g * 4 + 2])), "f"(__uint_as_float(r[g * 4 + 3])) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// returns 0 / error; *handled = 1 when this kernel took the problem
int conv3x3_wgrad32_try(const void* dy_grid, const void* x, float* gk, int F, int Hg, int Wg, int Cin, int Cout, cudaStream_t st,
                        int* handled) {
    *handled = 0;
    if (Cin != 32 || Cout != 64) return 0;
    { const char* e = getenv("XCP_CONV_WGRAD_OLD"); if (e && e[0] == '1') return 0; }        // A/B hook (tools/kernel_bench.py)
    *handled = 1;
    const long long R = (long long)F * Hg * Wg;
    CUtensorMap tmDY, tmX;
    if (int e = make_tmap_2d(&tmDY, dy_grid, 64, (uint64_t)R, 128, 64, 64, 128)) return e;
    if (int e = make_tmap_2d(&tmX, x, 32, (uint64_t)R, 64, 32, CW_X_ROWS, 64)) return e;
    ConvWgradParams p{gk, (int)((R + 63) / 64), Wg};
    constexpr int STAGES = 8;
    const int smem = 1024 + STAGES * CW_STAGE_BYTES + CW_DY_BYTES + 256;
    auto k = conv3x3_wgrad32_kernel<STAGES>;
    XCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = num_sms();
    if (grid > p.num_k_blocks) grid = p.num_k_blocks;
    k<<<grid, 192, smem, st>>>(tmDY, tmX, p);
    return check_cuda(cudaGetLastError(), "conv3x3_wgrad32 launch");
}

}  // namespace xcp
