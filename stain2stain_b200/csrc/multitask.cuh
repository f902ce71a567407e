// Kernels specific to the reference's multitask model (config M: src/models/components/shared_encoder.py,
// task_decoders.py, conditional_flow_matching_multitask_multiclassloss.py): train-mode BatchNorm coefficient folds
// (the streaming passes are the GroupNorm kernels with ReLU), 2x2 max-pool, bilinear x2 (align_corners) and their
// adjoints, the softmax Dice + cross-entropy segmentation loss.  16-bit NHWC activations, fp32 math.
#pragma once
#include "elementwise.cuh"

namespace s2s {

// ------------------------------------------------------------------------------------------------ BatchNorm2d (train)
// stats: [B][nchunks][C] (sum, sumsq) partials written by gn_stats_kernel.  One thread column per channel, 8 row lanes
// fold the B*nchunks partials in a fixed order (deterministic).
//   coef[b][c]      = (A, Bc),  A = gamma*rstd, Bc = beta - mean*A     (same for every sample b)
//   mean_rstd[b][c] = (mean, rstd)                                      (layout gn_bwd_reduce expects with G = C)
//   running_mean/var <- (1-m)*running + m*(mean, unbiased var)          (torch.nn.BatchNorm2d, momentum m)
constexpr int kBnCh = 32, kBnRows = 8;
__global__ void __launch_bounds__(kBnCh * kBnRows) bn_coef_kernel(const float2* __restrict__ stats, int nparts, int B,
                                                                  int C, long long count, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps,
                                                                  float momentum, float* __restrict__ running_mean,
                                                                  float* __restrict__ running_var,
                                                                  float2* __restrict__ coef,
                                                                  float2* __restrict__ mean_rstd) {
    __shared__ float s_a[kBnRows][kBnCh + 1], s_q[kBnRows][kBnCh + 1];
    const int ci = threadIdx.x % kBnCh, r = threadIdx.x / kBnCh;
    const int c = blockIdx.x * kBnCh + ci;
    float a = 0.f, q = 0.f;
    if (c < C)
        for (int k = r; k < nparts; k += kBnRows) {
            const float2 t = stats[(size_t)k * C + c];
            a += t.x;
            q += t.y;
        }
    s_a[r][ci] = a;
    s_q[r][ci] = q;
    __syncthreads();
    if (r == 0 && c < C) {
        a = q = 0.f;
#pragma unroll
        for (int k = 0; k < kBnRows; ++k) {
            a += s_a[k][ci];
            q += s_q[k][ci];
        }
        const float n = (float)count;
        const float mean = a / n;
        const float var = fmaxf(q / n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        const float A = gamma[c] * rstd;
        const float Bc = beta[c] - mean * A;
        for (int b = 0; b < B; ++b) {
            coef[(size_t)b * C + c] = make_float2(A, Bc);
            mean_rstd[(size_t)b * C + c] = make_float2(mean, rstd);
        }
        if (running_mean != nullptr) {
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (n / fmaxf(n - 1.f, 1.f));
        }
    }
}

// Fold of many partials in two levels.  The statistics of a 512^2 layer arrive as 2048 sub-tile partials per sample (32768
// per channel at B = 16) and the coefficient kernels have only C / 32 CTAs: folding there was a serial walk of 51 us on
// average (36 calls per multitask step).  This kernel cuts the parts into gridDim.y slices, each folded in the same fixed
// order into sums[slice][C]; the coefficient kernels then fold the slices (deterministic).  With one slice it is the
// per-rank total SyncBatchNorm all-reduces (torch.nn.SyncBatchNorm, what Lightning's `sync_batchnorm: True` of
// configs/trainer/ddp.yaml:9 installs): the totals of all ranks come back to the coefficient kernels as ONE part with the
// global element count.
__global__ void __launch_bounds__(kBnCh * kBnRows) bn_fold_kernel(const float2* __restrict__ parts, int nparts, int C,
                                                                  float2* __restrict__ sums) {
    __shared__ float s_a[kBnRows][kBnCh + 1], s_q[kBnRows][kBnCh + 1];
    const int ci = threadIdx.x % kBnCh, r = threadIdx.x / kBnCh;
    const int c = blockIdx.x * kBnCh + ci;
    const int per = (nparts + gridDim.y - 1) / gridDim.y;
    const int k0 = blockIdx.y * per, k1 = min(nparts, k0 + per);
    float a = 0.f, q = 0.f;
    if (c < C) {
        int k = k0 + r;
        for (; k + 3 * kBnRows < k1; k += 4 * kBnRows) {  // four independent loads in flight
            const float2 t0 = parts[(size_t)k * C + c], t1 = parts[(size_t)(k + kBnRows) * C + c];
            const float2 t2 = parts[(size_t)(k + 2 * kBnRows) * C + c], t3 = parts[(size_t)(k + 3 * kBnRows) * C + c];
            a += t0.x; q += t0.y; a += t1.x; q += t1.y; a += t2.x; q += t2.y; a += t3.x; q += t3.y;
        }
        for (; k < k1; k += kBnRows) {
            const float2 t = parts[(size_t)k * C + c];
            a += t.x;
            q += t.y;
        }
    }
    s_a[r][ci] = a;
    s_q[r][ci] = q;
    __syncthreads();
    if (r == 0 && c < C) {
        a = q = 0.f;
#pragma unroll
        for (int k = 0; k < kBnRows; ++k) {
            a += s_a[k][ci];
            q += s_q[k][ci];
        }
        sums[(size_t)blockIdx.y * C + c] = make_float2(a, q);
    }
}

// red: [B][nchunks][C] (sum dz, sum dz*xhat) partials from gn_bwd_reduce_kernel.
//   dx = dz*P + x*Q + R,  P = gamma*rstd, Q = -gamma*rstd^2*S2/N, R = -gamma*rstd*S1/N + gamma*mean*rstd^2*S2/N
//   dgamma += S2, dbeta += S1
__global__ void __launch_bounds__(kBnCh * kBnRows) bn_bwd_coef_kernel(const float2* __restrict__ red, int nparts, int B,
                                                                      int C, long long count,
                                                                      const float2* __restrict__ mean_rstd,
                                                                      const float* __restrict__ gamma,
                                                                      float4* __restrict__ pqr, float* __restrict__ dgamma,
                                                                      float* __restrict__ dbeta) {
    __shared__ float s_a[kBnRows][kBnCh + 1], s_q[kBnRows][kBnCh + 1];
    const int ci = threadIdx.x % kBnCh, r = threadIdx.x / kBnCh;
    const int c = blockIdx.x * kBnCh + ci;
    float a = 0.f, q = 0.f;
    if (c < C)
        for (int k = r; k < nparts; k += kBnRows) {
            const float2 t = red[(size_t)k * C + c];
            a += t.x;
            q += t.y;
        }
    s_a[r][ci] = a;
    s_q[r][ci] = q;
    __syncthreads();
    if (r == 0 && c < C) {
        a = q = 0.f;
#pragma unroll
        for (int k = 0; k < kBnRows; ++k) {
            a += s_a[k][ci];
            q += s_q[k][ci];
        }
        const float2 mr = mean_rstd[c];
        const float n = (float)count, ga = gamma[c];
        const float P = ga * mr.y;
        const float Q = -ga * mr.y * mr.y * q / n;
        const float R = -ga * mr.y * a / n + ga * mr.x * mr.y * mr.y * q / n;
        for (int b = 0; b < B; ++b) pqr[(size_t)b * C + c] = make_float4(P, Q, R, 0.f);
        dgamma[c] += q;
        dbeta[c] += a;
    }
}

// ------------------------------------------------------------------------------------------------ MaxPool2d(2)
__device__ __forceinline__ uint4 max8(const uint4& a, const uint4& b, int fmt) {
    float fa[8], fb[8];
    cvt8_in(a, fmt, fa);
    cvt8_in(b, fmt, fb);
#pragma unroll
    for (int e = 0; e < 8; ++e) fa[e] = fmaxf(fa[e], fb[e]);
    return cvt8_out(fa, fmt);  // exact: the result is one of the inputs
}
// out[b, y, x, :] = max over the 2x2 window of in (H, W = OUTPUT spatial dims)
__global__ void maxpool2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int vpp,
                                 int fmt) {
    const long long total = (long long)B * H * W * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int x = (int)(r % W);  r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const uint4* base = in + (((size_t)b * 2 * H + 2 * y) * (2 * W) + 2 * x) * vpp + v;
        const uint4 m0 = max8(ldg_stream(base), ldg_stream(base + vpp), fmt);
        const uint4 m1 = max8(ldg_stream(base + (size_t)2 * W * vpp), ldg_stream(base + (size_t)2 * W * vpp + vpp), fmt);
        stg_stream(out + i, max8(m0, m1, fmt));
    }
}
// dx[b, 2y+i, 2x+j, c] = g[b, y, x, c] at the FIRST position (row-major scan, like ATen) holding the window maximum
__global__ void maxpool2x_bwd_kernel(const uint4* __restrict__ xin, const uint4* __restrict__ g, uint4* __restrict__ dx,
                                     int B, int H, int W, int vpp, int xfmt, int gfmt) {
    const long long total = (long long)B * H * W * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int x = (int)(r % W);  r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const size_t o00 = (((size_t)b * 2 * H + 2 * y) * (2 * W) + 2 * x) * vpp + v;
        const size_t offs[4] = {o00, o00 + vpp, o00 + (size_t)2 * W * vpp, o00 + (size_t)2 * W * vpp + vpp};
        float f[4][8], gf[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) cvt8_in(ldg_stream(xin + offs[k]), xfmt, f[k]);
        cvt8_in(ldg_stream(g + i), gfmt, gf);
        float o[4][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int best = 0;
            float m = f[0][e];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (f[k][e] > m) {
                    m = f[k][e];
                    best = k;
                }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][e] = (k == best) ? gf[e] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) stg_stream(dx + offs[k], cvt8_out(o[k], gfmt));
    }
}

// ------------------------------------------------------------------------------------------------ bilinear x2, align_corners=True
// ATen's index rule for align_corners: src = dst * (in - 1) / (out - 1) evaluated in fp32; i0 = (int)src,
// i1 = i0 + (i0 < in - 1), lambda1 = src - i0.
__device__ __forceinline__ void bilin_src(int dst, float scale, int in, int& i0, int& i1, float& w1) {
    const float s = scale * (float)dst;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    w1 = s - (float)i0;
}
// out [B, 2H, 2W, C] from in [B, H, W, C].  Grid (x: vector columns of one output row, y: output row, z: sample): the row's
// two source rows and weight are block-uniform and every index is 32-bit (the first version decoded a flat 64-bit index
// with three 64-bit divisions per vector and ran at 0.27 of HBM).  Four rows per CTA instead of one was measured slower
// (forward 1.50 -> 1.56 ms, adjoint 1.86 -> 2.88 ms per multitask step): the row grid stays.
template <int F>
__global__ void __launch_bounds__(256) bilinear2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W,
                                                         int vpp) {
    const float sy = (H > 1) ? (float)(H - 1) / (float)(2 * H - 1) : 0.f;
    const float sx = (W > 1) ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
    const int y = blockIdx.y, b = blockIdx.z;
    int y0, y1;
    float wy;
    bilin_src(y, sy, H, y0, y1, wy);
    const int cols = 2 * W * vpp;
    const uint4* r0 = in + ((size_t)b * H + y0) * W * vpp;
    const uint4* r1 = in + ((size_t)b * H + y1) * W * vpp;
    uint4* orow = out + ((size_t)b * 2 * H + y) * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += gridDim.x * blockDim.x) {
        const int x = i / vpp, v = i - x * vpp;
        int x0, x1;
        float wx;
        bilin_src(x, sx, W, x0, x1, wx);
        float a[8], c[8], d[8], e8[8], o[8];
        cvt8_in_t<F>(__ldg(r0 + x0 * vpp + v), a);
        cvt8_in_t<F>(__ldg(r0 + x1 * vpp + v), c);
        cvt8_in_t<F>(__ldg(r1 + x0 * vpp + v), d);
        cvt8_in_t<F>(__ldg(r1 + x1 * vpp + v), e8);
#pragma unroll
        for (int e = 0; e < 8; ++e)
            o[e] = (1.f - wy) * ((1.f - wx) * a[e] + wx * c[e]) + wy * ((1.f - wx) * d[e] + wx * e8[e]);
        stg_stream(orow + i, cvt8_out_t<F>(o));
    }
}
// adjoint as a gather: din[b, yi, xi, :] = sum over the output pixels whose footprint touches (yi, xi).  Same grid shape
// (y: input row); the six candidate output rows and their weights are block-uniform.
template <int F>
__global__ void __launch_bounds__(256) bilinear2x_bwd_kernel(const uint4* __restrict__ g, uint4* __restrict__ din, int H,
                                                             int W, int vpp) {
    const float sy = (H > 1) ? (float)(H - 1) / (float)(2 * H - 1) : 0.f;
    const float sx = (W > 1) ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
    const int yi = blockIdx.y, b = blockIdx.z;
    // candidate outputs: src ~ dst / 2 (slightly less), so dst in [2*i - 2, 2*i + 3] covers every contributor
    float wrow[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int yo = 2 * yi - 2 + k;
        wrow[k] = 0.f;
        if (yo >= 0 && yo < 2 * H) {
            int y0, y1;
            float w;
            bilin_src(yo, sy, H, y0, y1, w);
            if (y0 == yi) wrow[k] += 1.f - w;
            if (y1 == yi) wrow[k] += w;
        }
    }
    const int cols = W * vpp, ocols = 2 * W * vpp;
    const uint4* gb = g + (size_t)b * 2 * H * ocols;
    uint4* drow = din + ((size_t)b * H + yi) * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += gridDim.x * blockDim.x) {
        const int xi = i / vpp, v = i - xi * vpp;
        float wcol[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int xo = 2 * xi - 2 + k;
            wcol[k] = 0.f;
            if (xo >= 0 && xo < 2 * W) {
                int x0, x1;
                float w;
                bilin_src(xo, sx, W, x0, x1, w);
                if (x0 == xi) wcol[k] += 1.f - w;
                if (x1 == xi) wcol[k] += w;
            }
        }
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 6; ++ky) {
            if (wrow[ky] == 0.f) continue;  // block-uniform
            const uint4* grow = gb + (size_t)(2 * yi - 2 + ky) * ocols + v;
#pragma unroll
            for (int kx = 0; kx < 6; ++kx) {
                if (wcol[kx] == 0.f) continue;
                float f[8];
                cvt8_in_t<F>(__ldg(grow + (2 * xi - 2 + kx) * vpp), f);
                const float w = wrow[ky] * wcol[kx];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, f[e], acc[e]);
            }
        }
        stg_stream(drow + i, cvt8_out_t<F>(acc));
    }
}

// ------------------------------------------------------------------------------------------------ layout glue
// fp32 NCHW [B, C, HW] -> 16-bit NHWC [B, HW, Cpad] (channels >= C are zero): narrow image-space gradients as GEMM operands
// Thread = one 16-byte vector (8 channels) of one pixel, consecutive threads = consecutive vectors of the NHWC row: stores are
// fully coalesced, only the vectors that hold real channels read (the element-per-thread version ran at 0.08 of the HBM rate).
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc16_pad_kernel(const float* __restrict__ in, uint16_t* __restrict__ out,
                                                                     int B, int C, int Cpad, int HW, int fmt) {
    const int vpp = Cpad >> 3;
    const long long total = (long long)B * HW * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        const long long r = i / vpp;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (v * 8 < C) {
            const int p = (int)(r % HW);
            const int b = (int)(r / HW);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = v * 8 + e;
                f[e] = c < C ? __ldg(in + ((size_t)b * C + c) * HW + p) : 0.f;
            }
            o = cvt8_out(f, fmt);
        }
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------ segmentation loss
// MulticlassDiceLoss + CrossEntropyLoss of the reference (conditional_flow_matching_multitask_multiclassloss.py:31-83,
// 231-236) on fp32 NCHW logits [B, C, HW] (C <= 8) and int64 targets [B, HW].
// sums (double): [0..C) I_c = sum p_c [t==c] m, [C..2C) P_c = sum p_c m, [2C..3C) T_c = sum [t==c] m,
//                [3C] = sum -log p_t over non-ignored pixels, [3C+1] = number of non-ignored pixels.
// m = Dice's valid mask: (t != ignore_index) when ignore_index >= 0, else 1 (the reference's rule).
constexpr int kSegMaxC = 8;
template <int kBlock>
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < kBlock / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(256) seg_loss_sums_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ target, int B, int C, int HW,
                                                            long long ignore_index, double* __restrict__ sums) {
    __shared__ double sh[8];
    float I[kSegMaxC], P[kSegMaxC], T[kSegMaxC];
    float ce = 0.f, cnt = 0.f;
#pragma unroll
    for (int c = 0; c < kSegMaxC; ++c) I[c] = P[c] = T[c] = 0.f;
    const long long total = (long long)B * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int b = (int)(i / HW);
        const long long t = target[i];
        float z[kSegMaxC], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                z[c] = logits[((size_t)b * C + c) * HW + p];
                mx = fmaxf(mx, z[c]);
            }
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                z[c] = __expf(z[c] - mx);
                den += z[c];
            }
        const float inv = 1.f / den;
        const bool ce_valid = (t != ignore_index);
        const float m = (ignore_index >= 0 && !ce_valid) ? 0.f : 1.f;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                const float pc = z[c] * inv;
                const float hit = (t == c) ? 1.f : 0.f;
                I[c] += pc * hit * m;
                P[c] += pc * m;
                T[c] += hit * m;
                if (ce_valid && t == c) ce -= __logf(pc);
            }
        if (ce_valid) cnt += 1.f;
    }
    for (int c = 0; c < C; ++c) {
        const double a = block_sum_d<256>((double)I[c], sh), bsum = block_sum_d<256>((double)P[c], sh),
                     d = block_sum_d<256>((double)T[c], sh);
        if (threadIdx.x == 0) {
            atomicAdd(sums + c, a);
            atomicAdd(sums + C + c, bsum);
            atomicAdd(sums + 2 * C + c, d);
        }
    }
    const double a = block_sum_d<256>((double)ce, sh), bsum = block_sum_d<256>((double)cnt, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums + 3 * C, a);
        atomicAdd(sums + 3 * C + 1, bsum);
    }
}
// dlogits[b, k, p] = gscale * ( w_dice * dDice/dlogit + w_ce * dCE/dlogit ),  gscale read from device memory
__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ target, int B, int C, int HW,
                                                           long long ignore_index, const double* __restrict__ sums,
                                                           float smooth, float w_dice, float w_ce,
                                                           const float* __restrict__ gscale,
                                                           float* __restrict__ dlogits) {
    __shared__ float s_a[kSegMaxC], s_b[kSegMaxC];  // dDice/dp_c = a_c * [t==c] + b_c  (times the valid mask)
    __shared__ float s_ce;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        const double I = sums[c], U = sums[C + c] + sums[2 * C + c];
        const double den = U + (double)smooth;
        // dice_c = (2I + s)/(U + s); loss = 1 - mean_c dice_c ; dI/dp = [t==c], dU/dp = 1
        s_a[c] = (float)(-(2.0 / den) / (double)C);
        s_b[c] = (float)(((2.0 * I + (double)smooth) / (den * den)) / (double)C);
    }
    if (threadIdx.x == 0) s_ce = sums[3 * C + 1] > 0.0 ? (float)(1.0 / sums[3 * C + 1]) : 0.f;
    __syncthreads();
    const float gs = gscale ? *gscale : 1.f;
    const long long total = (long long)B * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int b = (int)(i / HW);
        const long long t = target[i];
        float z[kSegMaxC], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                z[c] = logits[((size_t)b * C + c) * HW + p];
                mx = fmaxf(mx, z[c]);
            }
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                z[c] = __expf(z[c] - mx);
                den += z[c];
            }
        const float inv = 1.f / den;
        const bool ce_valid = (t != ignore_index);
        const float m = (ignore_index >= 0 && !ce_valid) ? 0.f : 1.f;
        float dp[kSegMaxC], dot = 0.f;
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                z[c] *= inv;  // p_c
                dp[c] = m * (s_a[c] * ((t == c) ? 1.f : 0.f) + s_b[c]);
                dot += z[c] * dp[c];
            }
#pragma unroll
        for (int c = 0; c < kSegMaxC; ++c)
            if (c < C) {
                float d = w_dice * z[c] * (dp[c] - dot);
                if (ce_valid) d += w_ce * s_ce * (z[c] - ((t == c) ? 1.f : 0.f));
                dlogits[((size_t)b * C + c) * HW + p] = gs * d;
            }
    }
}

}  // namespace s2s
