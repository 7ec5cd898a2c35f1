// GPU audio front-end: waveform -> MFCC, replacing the offline
//   librosa.feature.mfcc(y, sr=16000, n_mfcc=13, n_fft=400, hop_length=160).T        (wavfake_audio_dataset.py:17-19,40-44)
// so that XceptionLSTMA can be fed from waveforms resident in HBM.  Two launches per batch of waveforms:
//   1. mfcc_logmel_kernel : one CTA per frame.  Centre-padded framing + periodic Hann window, a direct 400-point real DFT
//      from a shared-memory twiddle table, folded over the x[n] / x[N-n] pairs (n_fft = 400 is not a power of two; 201 bins
//      x 199 pairs x 2 FMA per frame is ~80 kFMA, three orders of magnitude below one Xception frame), |X|^2, the 128
//      Slaney mel filters, 10*log10(max(amin, .)), and the per-waveform maximum (power_to_db's top_db clip is relative to the maximum of the WHOLE spectrogram).
//   2. mfcc_dct_kernel    : clip at (max - top_db), orthonormal DCT-II, first n_mfcc coefficients.
// The mel filterbank is a constant of (sr, n_fft, n_mels) and is built once on the host (audio_frontend.py).
#include "common.cuh"

namespace {
using namespace xcp;
#define ST ((cudaStream_t)stream)

constexpr int MFCC_MAX_FFT = 1024;
constexpr int MFCC_MAX_MELS = 256;

XCP_DEVINL int float_key(float f) {            // order-preserving float -> int map for atomicMax
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
XCP_DEVINL float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
mfcc_logmel_kernel(const float* __restrict__ wav, int L, const float* __restrict__ melfb_t, int n_fft, int hop, int n_mels,
                   int pad_reflect, float amin, float* __restrict__ logmel, int* __restrict__ gmax, int T) {
    __shared__ float xs[MFCC_MAX_FFT], pw[MFCC_MAX_FFT / 2 + 1];
    __shared__ float2 tw[MFCC_MAX_FFT], pe[MFCC_MAX_FFT / 2 + 1];
    __shared__ float red[8];
    const int t = blockIdx.x, b = blockIdx.y;
    const float* y = wav + (long long)b * L;
    const int half = n_fft / 2, n_bins = half + 1;
    for (int n = threadIdx.x; n < n_fft; n += 256) {
        int j = t * hop + n - half;
        float v = 0.f;
        if (pad_reflect) {
            if (j < 0) j = -j;
            if (j >= L) j = 2 * (L - 1) - j;
            v = (j >= 0 && j < L) ? y[j] : 0.f;
        } else if (j >= 0 && j < L) {
            v = y[j];
        }
        const float ph = 2.f * (float)n / (float)n_fft;           // in units of pi
        float sn, cs;
        sincospif(ph, &sn, &cs);
        tw[n] = make_float2(cs, sn);
        xs[n] = v * (0.5f - 0.5f * cs);                           // periodic Hann
    }
    __syncthreads();
    // real input: x[n] and x[N-n] share a twiddle (cos even, sin odd), so the DFT runs over the (N-1)/2 folded pairs
    //   Re X[k] = x[0] (+ (-1)^k x[N/2]) + sum_n (x[n] + x[N-n]) cos(2 pi k n / N),  Im X[k] = -sum_n (x[n] - x[N-n]) sin(...)
    const int pairs = (n_fft - 1) / 2;
    for (int n = 1 + threadIdx.x; n <= pairs; n += 256) pe[n] = make_float2(xs[n] + xs[n_fft - n], xs[n] - xs[n_fft - n]);
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += 256) {
        float re0 = xs[0], re1 = 0.f, im0 = 0.f, im1 = 0.f;
        if ((n_fft & 1) == 0) re1 = (k & 1) ? -xs[half] : xs[half];
        int i0 = k, i1 = (2 * k) % n_fft;                         // (k * n) mod n_fft for n = 1, 2, then += 2k
        const int step = i1;
        int n = 1;
        for (; n + 1 <= pairs; n += 2) {
            const float2 p0 = pe[n], p1 = pe[n + 1], w0 = tw[i0], w1 = tw[i1];
            re0 = fmaf(p0.x, w0.x, re0); im0 = fmaf(p0.y, w0.y, im0);
            re1 = fmaf(p1.x, w1.x, re1); im1 = fmaf(p1.y, w1.y, im1);
            i0 += step; if (i0 >= n_fft) i0 -= n_fft;
            i1 += step; if (i1 >= n_fft) i1 -= n_fft;
        }
        if (n <= pairs) { const float2 p0 = pe[n], w0 = tw[i0]; re0 = fmaf(p0.x, w0.x, re0); im0 = fmaf(p0.y, w0.y, im0); }
        const float re = re0 + re1, im = im0 + im1;
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    float mx = -3.0e38f;
    for (int m = threadIdx.x; m < n_mels; m += 256) {
        float a0 = 0.f, a1 = 0.f;
        int k = 0;
        for (; k + 1 < n_bins; k += 2) {
            a0 = fmaf(melfb_t[(long long)k * n_mels + m], pw[k], a0);
            a1 = fmaf(melfb_t[(long long)(k + 1) * n_mels + m], pw[k + 1], a1);
        }
        if (k < n_bins) a0 = fmaf(melfb_t[(long long)k * n_mels + m], pw[k], a0);
        const float db = 10.f * log10f(fmaxf(amin, a0 + a1));
        logmel[((long long)b * T + t) * n_mels + m] = db;
        mx = fmaxf(mx, db);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
        atomicMax(gmax + b, float_key(mx));
    }
}

__global__ void __launch_bounds__(128)
mfcc_dct_kernel(const float* __restrict__ logmel, const int* __restrict__ gmax, float top_db, int n_mels, int n_mfcc,
                float* __restrict__ out, int T) {
    __shared__ float row[MFCC_MAX_MELS];
    const int t = blockIdx.x, b = blockIdx.y;
    const float floor_db = key_float(gmax[b]) - top_db;
    for (int m = threadIdx.x; m < n_mels; m += 128) row[m] = fmaxf(logmel[((long long)b * T + t) * n_mels + m], floor_db);
    __syncthreads();
    // one warp per coefficient: lanes over the mel bands, cos(pi (2m+1) c / (2 n_mels)) with exact argument reduction
    const int lane = threadIdx.x & 31;
    for (int c = threadIdx.x >> 5; c < n_mfcc; c += 4) {
        float acc = 0.f;
        for (int m = lane; m < n_mels; m += 32) {
            const int num = ((2 * m + 1) * c) % (4 * n_mels);               // angle = pi * num / (2 n_mels), period 4 n_mels
            acc = fmaf(row[m], cospif((float)num / (float)(2 * n_mels)), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[((long long)b * T + t) * n_mfcc + c] = acc * sqrtf((c == 0 ? 1.f : 2.f) / (float)n_mels);
    }
}
}  // namespace

extern "C" int xcp_mfcc_frames(int L, int hop) { return (L < 0 || hop <= 0) ? 0 : 1 + L / hop; }

extern "C" int xcp_mfcc(const float* wav, int B, int L, const float* melfb_t, int n_fft, int hop, int n_mels, int n_mfcc,
                        int pad_reflect, float amin, float top_db, float* logmel_ws, int* gmax_ws, float* out, int device,
                        void* stream) {
    XCP_REQUIRE(B > 0 && L > 0 && hop > 0, "xcp_mfcc: empty batch (B=%d L=%d hop=%d)", B, L, hop);
    XCP_REQUIRE(n_fft >= 2 && n_fft <= MFCC_MAX_FFT && n_mels > 0 && n_mels <= MFCC_MAX_MELS && n_mfcc > 0 && n_mfcc <= n_mels,
                "xcp_mfcc: n_fft <= %d, n_mels <= %d, n_mfcc <= n_mels (got %d, %d, %d)", MFCC_MAX_FFT, MFCC_MAX_MELS, n_fft, n_mels, n_mfcc);
    XCP_REQUIRE(!pad_reflect || L > n_fft / 2, "xcp_mfcc: reflect padding needs more than n_fft/2 samples");
    XCP_REQUIRE(B <= 65535, "xcp_mfcc: at most 65535 waveforms per call");
    XCP_CUDA(cudaSetDevice(device));
    const int T = xcp_mfcc_frames(L, hop);
    XCP_CUDA(cudaMemsetAsync(gmax_ws, 0x80, sizeof(int) * B, ST));              // key of about -3.4e38
    mfcc_logmel_kernel<<<dim3(T, B), 256, 0, ST>>>(wav, L, melfb_t, n_fft, hop, n_mels, pad_reflect, amin, logmel_ws, gmax_ws, T);
    XCP_CUDA(cudaGetLastError());
    mfcc_dct_kernel<<<dim3(T, B), 128, 0, ST>>>(logmel_ws, gmax_ws, top_db, n_mels, n_mfcc, out, T);
    return check_cuda(cudaGetLastError(), "mfcc launch");
}
