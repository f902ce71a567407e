// Multi-tensor Adam (torch.optim.Adam semantics, L2 weight-decay form) in ONE launch over every parameter tensor.
//
// The host uploads a table of tensor descriptors and a list of (tensor, chunk) work items; each CTA owns one chunk of
// kAdamChunk contiguous elements of one tensor, so the grid covers the whole 71 M-parameter model with ~4.4 k CTAs
// (30 per SM).  Pure HBM streaming: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter, 128-bit accesses
// when the four pointers are 16 B aligned (DDP bucket views may not be: scalar path).
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 16384;  // elements per CTA: 16 float4 per thread

struct AdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
};

struct AdamHyper {
    float lr, beta1, beta2, eps, weight_decay;
    float omb1, omb2;       // 1 - beta1, 1 - beta2 evaluated in double on the host (as torch does)
    float bias_corr1;       // 1 - beta1^t
    float inv_sqrt_bc2;     // 1 / sqrt(1 - beta2^t)
    float grad_scale;       // multiplies g before use (1 / world size when the all-reduce summed)
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h) {
    g = g * h.grad_scale + h.weight_decay * p;
    m = h.beta1 * m + h.omb1 * g;
    v = h.beta2 * v + h.omb2 * g * g;
    const float denom = sqrtf(v) * h.inv_sqrt_bc2 + h.eps;
    p -= (h.lr / h.bias_corr1) * (m / denom);
}

// step_dev (optional): the 1-based step count in device memory.  A launch replayed from a CUDA graph has its host
// arguments frozen at capture; the bias corrections are then evaluated here from the live counter (in double, as the
// host path and torch do).
__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const AdamTensor* __restrict__ tensors,
                                                                  const int2* __restrict__ work, AdamHyper h,
                                                                  const long long* __restrict__ step_dev) {
    if (step_dev != nullptr) {
        const double t = (double)__ldg(step_dev);
        h.bias_corr1 = (float)(1.0 - pow((double)h.beta1, t));
        h.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)h.beta2, t)));
    }
    const int2 wi = work[blockIdx.x];
    const AdamTensor t = tensors[wi.x];
    const long long beg = (long long)wi.y * kAdamChunk;
    const long long end = min(t.n, beg + kAdamChunk);
    const bool aligned = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                           reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
    if (aligned) {
        const long long vend = beg + ((end - beg) & ~3LL);
        for (long long i = beg + 4LL * threadIdx.x; i < vend; i += 4LL * kAdamThreads) {
            float4 p = *reinterpret_cast<const float4*>(t.p + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(t.g + i));
            float4 m = *reinterpret_cast<const float4*>(t.m + i);
            float4 v = *reinterpret_cast<const float4*>(t.v + i);
            adam_update(p.x, g.x, m.x, v.x, h);
            adam_update(p.y, g.y, m.y, v.y, h);
            adam_update(p.z, g.z, m.z, v.z, h);
            adam_update(p.w, g.w, m.w, v.w, h);
            *reinterpret_cast<float4*>(t.p + i) = p;
            *reinterpret_cast<float4*>(t.m + i) = m;
            *reinterpret_cast<float4*>(t.v + i) = v;
        }
        for (long long i = vend + threadIdx.x; i < end; i += kAdamThreads) adam_update(t.p[i], t.g[i], t.m[i], t.v[i], h);
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kAdamThreads) adam_update(t.p[i], t.g[i], t.m[i], t.v[i], h);
    }
}

// dst_i[0..n) = src_i[0..n) for MANY fp32 tensors in one launch (same table / work-list scheme): gathers the per-parameter
// gradients into ONE flat buffer so that the data-parallel step needs a single all-reduce.
struct CopyTensor {
    const float* src;
    float* dst;
    long long n;
};
__global__ void __launch_bounds__(kAdamThreads) copy_multi_kernel(const CopyTensor* __restrict__ tensors,
                                                                  const int2* __restrict__ work) {
    const int2 wi = work[blockIdx.x];
    const CopyTensor t = tensors[wi.x];
    const long long beg = (long long)wi.y * kAdamChunk;
    const long long end = min(t.n, beg + kAdamChunk);
    const bool aligned = ((reinterpret_cast<uintptr_t>(t.src) | reinterpret_cast<uintptr_t>(t.dst)) & 15) == 0;
    if (aligned) {
        const long long vend = beg + ((end - beg) & ~3LL);
        for (long long i = beg + 4LL * threadIdx.x; i < vend; i += 4LL * kAdamThreads)
            *reinterpret_cast<float4*>(t.dst + i) = __ldg(reinterpret_cast<const float4*>(t.src + i));
        for (long long i = vend + threadIdx.x; i < end; i += kAdamThreads) t.dst[i] = t.src[i];
    } else {
        for (long long i = beg + threadIdx.x; i < end; i += kAdamThreads) t.dst[i] = t.src[i];
    }
}

}  // namespace s2s
