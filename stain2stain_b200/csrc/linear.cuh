// Small fp32 GEMMs of the embedding path: timestep-embedding MLP, label embedding add and the 22 FiLM
// projections Linear(512 -> 2*Cout) of the ResBlocks, forward and backward.  M = batch (<= a few hundred), K, N <= 13.5 k:
// 0.9 GFLOP per step in total, i.e. bound by reading the fp32 weights once (28 MB) and by launch count -- so MANY GEMMs
// run in ONE launch: the job table travels as a kernel parameter (capturable in a CUDA graph, no device-side table),
// each CTA owns one 64 x 64 output tile of one job.  Plain fp32 FMA in registers (exactly the reference's arithmetic,
// no tensor cores: the values feed normalisation scales and must not be rounded to 16 bit).
//
//   C[m][n] = bias[n] + add[m][n] + sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn]        (generic element strides)
//   C2[m][n] = silu(C[m][n])                                                          (optional second output)
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int kLinTile = 64, kLinK = 16, kLinThreads = 256, kLinMaxJobs = 36;

struct GemmJob {
    const float* A;
    const float* B;
    float* C;
    const float* bias;  // [N] or nullptr
    const float* add;   // [M][ld_add] or nullptr
    float* C2;          // silu(C), same geometry as C, or nullptr
    int M, N, K;
    int ldc, ld_add;
    int sam, sak, sbk, sbn;
};
struct GemmBatch {
    int njobs;
    int tile_end[kLinMaxJobs];  // exclusive prefix sum of the jobs' tile counts
    GemmJob jobs[kLinMaxJobs];
};
static_assert(sizeof(GemmBatch) <= 4000, "the job table must fit the kernel parameter space");

__global__ void __launch_bounds__(kLinThreads) linear_multi_kernel(const __grid_constant__ GemmBatch batch) {
    __shared__ float As[kLinK][kLinTile + 4];
    __shared__ float Bs[kLinK][kLinTile + 4];
    int j = 0;
    while (j + 1 < batch.njobs && (int)blockIdx.x >= batch.tile_end[j]) ++j;
    const GemmJob& jb = batch.jobs[j];
    const int local = (int)blockIdx.x - (j == 0 ? 0 : batch.tile_end[j - 1]);
    const int ntn = (jb.N + kLinTile - 1) / kLinTile;
    const int m0 = (local / ntn) * kLinTile, n0 = (local % ntn) * kLinTile;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // thread owns rows ty*4.., columns tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    for (int k0 = 0; k0 < jb.K; k0 += kLinK) {
        // stage a [16 k][64 m] slab of A and a [16 k][64 n] slab of B; the fastest thread index follows the unit stride
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = threadIdx.x + r * kLinThreads;
            int kk, mm;
            if (jb.sak == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < jb.M && k < jb.K) ? __ldg(jb.A + (size_t)m * jb.sam + (size_t)k * jb.sak) : 0.f;
            int kb, nn;
            if (jb.sbk == 1) { kb = i & 15; nn = i >> 4; } else { nn = i & 63; kb = i >> 6; }
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < jb.N && k2 < jb.K) ? __ldg(jb.B + (size_t)k2 * jb.sbk + (size_t)n * jb.sbn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kLinK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][e] = fmaf(av[i], bv[e], acc[i][e]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= jb.M) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n = n0 + tx * 4 + e;
            if (n >= jb.N) continue;
            float v = acc[i][e];
            if (jb.bias) v += __ldg(jb.bias + n);
            if (jb.add) v += __ldg(jb.add + (size_t)m * jb.ld_add + n);
            jb.C[(size_t)m * jb.ldc + n] = v;
            if (jb.C2) jb.C2[(size_t)m * jb.ldc + n] = v / (1.f + expf(-v));
        }
    }
}

// out[i] = (sum_{j < nparts} parts[j*n + i]) * silu'(z[i])   (z == nullptr: plain sum): folds the per-block partial
// gradients of the shared embedding and takes them through the SiLU in the same pass; exact sigmoid (fp32 path).
__global__ void sum_parts_silu_bwd_kernel(const float* __restrict__ parts, int nparts, long long n, const float* __restrict__ z,
                                          float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int j = 0; j < nparts; ++j) v += parts[(size_t)j * n + i];
        if (z != nullptr) {
            const float zz = z[i];
            const float s = 1.f / (1.f + expf(-zz));
            v *= s * (1.f + zz * (1.f - s));
        }
        out[i] = v;
    }
}

// emb[b][0:half] = cos(t_b * f_i), emb[b][half:2*half] = sin(t_b * f_i), f_i = exp(-ln(max_period) * i / half)
// (guided-diffusion `timestep_embedding`; t is used raw, no x1000 -- SURVEY.md A.3); an odd dim gets a trailing zero.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim, float max_period, float* __restrict__ emb) {
    const int half = dim / 2;
    const int total = B * dim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / dim, c = i % dim;
        float v = 0.f;
        if (c < 2 * half) {
            const int f = c < half ? c : c - half;
            const float freq = expf(-logf(max_period) * (float)f / (float)half);
            const float arg = t[b] * freq;
            v = c < half ? cosf(arg) : sinf(arg);
        }
        emb[i] = v;
    }
}

}  // namespace s2s
