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

// ---- multi-tensor variant: one launch walks every parameter tensor of the model through a device-side table.
// Used by FusedAdam (optim.py): ~300 parameter tensors -> 2 launches per step instead of ~350, and the step counter
// lives in device memory so the launch is CUDA-graph capturable (bias corrections are computed on the device).
struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; int* step; };   // 48 B
constexpr int ADAM_CHUNK = 8192;     // elements per work item

__global__ void __launch_bounds__(256)
sumsq_multi_kernel(const AdamTensor* __restrict__ tt, const int2* __restrict__ chunks, int n_chunks, float* __restrict__ out) {
    __shared__ float s[8];
    float l = 0.f;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int2 ch = chunks[c];
        const AdamTensor t = tt[ch.x];
        const long long i0 = (long long)ch.y * ADAM_CHUNK;
        const long long i1 = min(i0 + ADAM_CHUNK, t.n);
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) l = fmaf(t.g[i], t.g[i], l);
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += s[i];
        atomicAdd(out, tot);
    }
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ tt, const int2* __restrict__ chunks, int n_chunks,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled,
                  const float* __restrict__ sumsq, float max_norm, float grad_scale, const float* __restrict__ hyper) {
    if (hyper != nullptr) { lr = hyper[0]; weight_decay = hyper[1]; }     // device-resident: a graph replay sees scheduler updates
    float clip = grad_scale;
    if (sumsq != nullptr) {
        const float norm = sqrtf(*sumsq) * grad_scale;
        const float c = max_norm / (norm + 1e-6f);
        if (c < 1.f) clip *= c;
    }
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int2 ch = chunks[c];
        const AdamTensor t = tt[ch.x];
        // torch.optim.Adam keeps one step count per parameter (a parameter that had no gradient for a while, e.g. the
        // backbone frozen for the first epochs of train_visual.py:551-556, starts its bias correction late)
        const float stepf = (float)(*t.step);            // already incremented for this step (>= 1)
        const float bc1 = 1.f - powf(beta1, stepf), bc2 = 1.f - powf(beta2, stepf);
        const float rs_bc2 = rsqrtf(bc2), step_size = lr / bc1;
        const long long i0 = (long long)ch.y * ADAM_CHUNK;
        const long long i1 = min(i0 + ADAM_CHUNK, t.n);
        for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
            float gi = t.g[i] * clip;
            float pi = t.p[i];
            if (weight_decay != 0.f) {
                if (decoupled) pi *= (1.f - lr * weight_decay);
                else gi = fmaf(weight_decay, pi, gi);
            }
            const float mi = fmaf(beta1, t.m[i], (1.f - beta1) * gi);
            const float vi = fmaf(beta2, t.v[i], (1.f - beta2) * gi * gi);
            t.m[i] = mi; t.v[i] = vi;
            const float denom = sqrtf(vi) * rs_bc2 + eps;
            t.p[i] = pi - step_size * (mi / denom);
        }
    }
}

__global__ void adam_prologue_kernel(const AdamTensor* __restrict__ tt, int n_tensors, float* sumsq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tensors) *tt[i].step += 1;
    if (i == 0 && sumsq != nullptr) *sumsq = 0.f;
}

}  // namespace xcp
using namespace xcp;

// table: n_tensors x {p, g, m, v, n, step*} (48 B each, device memory; every tensor's own int step counter is incremented here
// first); chunks: n_chunks x {tensor index, chunk index} (8192 elements per chunk).  max_norm > 0: clip by the global gradient
// norm (sumsq_ws = device float scratch).  3 launches (2 without clipping), no host synchronisation, graph-capturable.
// hyper: optional DEVICE pointer to {lr, weight_decay}; when non-NULL it overrides the two host scalars, so a captured CUDA
// graph does not bake the learning rate in (the host refreshes the two floats before a replay).
extern "C" int xcp_adam_multi(const void* table, int n_tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int decoupled, float* sumsq_ws, float max_norm, float grad_scale,
                              const float* hyper, int device, void* stream) {
    XCP_REQUIRE(table != nullptr && chunks != nullptr && n_tensors > 0 && n_chunks > 0, "xcp_adam_multi: bad table");
    XCP_REQUIRE(max_norm <= 0.f || sumsq_ws != nullptr, "xcp_adam_multi: clipping needs the sumsq scratch");
    XCP_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool clipping = max_norm > 0.f;
    adam_prologue_kernel<<<(n_tensors + 255) / 256, 256, 0, st>>>((const AdamTensor*)table, n_tensors, clipping ? sumsq_ws : nullptr);
    int grid = n_chunks < 8 * num_sms() ? n_chunks : 8 * num_sms();
    if (clipping) sumsq_multi_kernel<<<grid, 256, 0, st>>>((const AdamTensor*)table, (const int2*)chunks, n_chunks, sumsq_ws);
    adam_multi_kernel<<<grid, 256, 0, st>>>((const AdamTensor*)table, (const int2*)chunks, n_chunks, lr, beta1, beta2, eps,
                                            weight_decay, decoupled, clipping ? sumsq_ws : nullptr, max_norm, grad_scale, hyper);
    return check_cuda(cudaGetLastError(), "adam_multi launch");
}

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
