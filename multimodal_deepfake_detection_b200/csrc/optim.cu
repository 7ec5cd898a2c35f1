// Fused optimizer-side kernels: global grad-norm, clip + Adam/AdamW over a flat fp32
// arena in one launch, no host synchronisation.  Reference semantics: torch.optim.Adam(lr, weight_decay) with
// L2-style decay added to the gradient (train_visual.py:533), AdamW decoupled decay (train_au_face.py:616-619),
// clip_grad_norm_(…, 1.0) (train_visual.py:575).
#include "common.cuh"

namespace xcp {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    __shared__ float s[32];
    float l = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        l = fmaf(g[i], g[i], l);
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s[i];
        atomicAdd(out, tot);
    }
}

// p, g, m, v: flat fp32 [n].  sumsq: device scalar with the squared global grad norm (or null = no clipping).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                            float bc1, float bc2, const float* __restrict__ sumsq, float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (sumsq != nullptr) {
        const float norm = sqrtf(*sumsq) * grad_scale;
        const float c = max_norm / (norm + 1e-6f);
        if (c < 1.f) clip *= c;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * clip;
        float pi = p[i];
        if (weight_decay != 0.f) {
            if (decoupled) pi *= (1.f - lr * weight_decay);
            else gi = fmaf(weight_decay, pi, gi);
        }
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

}  // namespace xcp
using namespace xcp;

extern "C" int xcp_grad_sumsq(const float* g, long long n, float* out, int zero_first, int device, void* stream) {
    XCP_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (zero_first) XCP_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    long long grid = (n + 255) / 256;
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    sumsq_kernel<<<(int)grid, 256, 0, st>>>(g, n, out);
    return check_cuda(cudaGetLastError(), "sumsq launch");
}

extern "C" int xcp_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int decoupled, int step, const float* sumsq, float max_norm,
                             float grad_scale, int device, void* stream) {
    XCP_REQUIRE(step >= 1, "xcp_adam_step: step must start at 1");
    XCP_CUDA(cudaSetDevice(device));
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    long long grid = (n + 255) / 256;
    if (grid > 2368) grid = 2368;
    if (grid < 1) grid = 1;
    adam_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, bc1, bc2,
                                                            sumsq, max_norm, grad_scale);
    return check_cuda(cudaGetLastError(), "adam launch");
}
