// Image-space kernels either side of the UNet: all HBM-bound streaming kernels over
// 3/4-channel tile tensors (a 256^2 tile is 0.8 MB in fp32 -- three orders of magnitude below the UNet's activations).
//
//   patch_pack_kernel<CT>     f3  3x3 patch operand of the stem conv for 3 + 1 channel inputs (tile + condition mask)
//   fm_loss_weighted_kernel   f3  mask-weighted flow-matching MSE (conditional_flow_matching_masked.py:76-92)
//   roi_charbonnier_kernel    f3  ROI Charbonnier term (conditional_flow_matching_ROI_loss.py:73-97)
//   tile_prep_kernel          f1  uint8 HWC tile pair -> crop + h/v flip + to_tensor + Normalize(0.5, 0.5) -> fp32 NCHW
//   resample_u8_{h,v}_kernel  f1  Pillow's antialiased 8-bit resampling passes (TF.resize of the eval path)
//   denorm_u8_kernel          f4  (x*0.5+0.5).clamp(0,1) -> uint8 HWC
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace s2s {

// ------------------------------------------------------------------------------------------------ stem operand, CT channels
// x0/x1: fp32 NCHW [B,Cx,H,W]; extra: fp32 [B,1,H,W] or nullptr (the condition mask, never interpolated); CT = Cx +
// (extra != nullptr) <= 7.  dst: 16-bit NHWC [B,H,W,64], column tap*CT + c = src_c[y + dy - 1][x + dx - 1] (zero outside
// the image, zero for columns >= 9*CT).  With x1: src_c = (1 - t_b) x0_c + t_b x1_c for c < Cx.
template <int CT>
__global__ void patch_pack_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ t,
                                  const float* __restrict__ extra, int Cx, int B, int H, int W,
                                  __nv_bfloat16* __restrict__ dst, int fmt) {
    static_assert(9 * CT <= 64, "patch columns must fit one 64-wide k-block");
    constexpr int NV = (9 * CT + 7) / 8;  // 16-byte vectors that hold real columns
    const long long npix = (long long)B * H * W;
    for (long long pidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; pidx < npix;
         pidx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pidx % W);
        const int y = (int)((pidx / W) % H);
        const int b = (int)(pidx / ((long long)W * H));
        const float tb = (x1 != nullptr) ? t[b] : 0.f;
        float v[NV * 8];
#pragma unroll
        for (int j = 0; j < NV * 8; ++j) v[j] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + (tap / 3 - 1), xx = x + (tap % 3 - 1);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    float s;
                    if (c < Cx) {
                        const size_t o = (((size_t)b * Cx + c) * H + yy) * W + xx;
                        s = __ldg(x0 + o);
                        if (x1 != nullptr) s = (1.f - tb) * s + tb * __ldg(x1 + o);
                    } else {
                        s = __ldg(extra + ((size_t)b * H + yy) * W + xx);
                    }
                    v[tap * CT + c] = s;
                }
            }
        }
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)pidx * 64);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = v[j * 8 + e];
            d[j] = cvt8_out(f, fmt);
        }
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = NV; j < 8; ++j) d[j] = z;
    }
}

// ------------------------------------------------------------------------------------------------ block reduction helper
template <int N>
__device__ __forceinline__ void block_add(float (&acc)[N], float* __restrict__ out) {
    __shared__ float part[N][kEwThreads / 32];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float s = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float s = threadIdx.x < (blockDim.x >> 5) ? part[k][threadIdx.x] : 0.f;
            s = warp_sum(s);
            if (threadIdx.x == 0) atomicAdd(out + k, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------ mask-weighted FM loss
// w = 1 + lam * mask[b,0,p] broadcast over C channels;  sums[0] += sum w (v - (x1 - x0))^2, sums[1] += sum w,
// dv = 2 w (v - (x1 - x0))  (the caller scales by 1 / (sums[1] + 1e-8)).
__global__ void __launch_bounds__(kEwThreads) fm_loss_weighted_kernel(
    const float* __restrict__ v, const float* __restrict__ x0, const float* __restrict__ x1,
    const float* __restrict__ mask, float lam, int B, int C, int HW, float* __restrict__ sums, float* __restrict__ dv) {
    float acc[2] = {0.f, 0.f};
    const long long n = (long long)B * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int b = (int)(i / ((long long)HW * C));
        const float w = fmaf(lam, __ldg(mask + (size_t)b * HW + p), 1.f);
        const float d = v[i] - (x1[i] - x0[i]);
        acc[0] = fmaf(w * d, d, acc[0]);
        acc[1] += w;
        if (dv) dv[i] = 2.f * w * d;
    }
    block_add<2>(acc, sums);
}

// ------------------------------------------------------------------------------------------------ ROI Charbonnier
// xt - x1 = (1 - t_b)(x0 - x1) for sigma = 0;  sums[0] += sum_{b,c,p} sqrt(diff^2 + eps^2) m[b,p], sums[1] += sum_{b,p} m.
__global__ void __launch_bounds__(kEwThreads) roi_charbonnier_kernel(
    const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ t,
    const float* __restrict__ mask, int B, int C, int HW, float eps, float* __restrict__ sums) {
    float acc[2] = {0.f, 0.f};
    const long long n = (long long)B * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int b = (int)(i / HW);
        const float m = __ldg(mask + i);
        const float tb = __ldg(t + b);
        acc[1] += m;
        for (int c = 0; c < C; ++c) {
            const size_t o = ((size_t)b * C + c) * HW + p;
            const float a = x0[o], e = x1[o];
            const float xt = tb * e + (1.f - tb) * a;  // torchcfm sample_xt order: t*x1 + (1-t)*x0
            const float diff = xt - e;
            acc[0] = fmaf(sqrtf(fmaf(diff, diff, eps * eps)), m, acc[0]);
        }
    }
    block_add<2>(acc, sums);
}

// ------------------------------------------------------------------------------------------------ tile preparation (input side)
// src/tgt: uint8 HWC [B,Hs,Ws,3] (channel order BGR if bgr != 0, as cv2.imread returns it); params: int32 [B][4] =
// (top, left, hflip, vflip).  out0/out1: fp32 NCHW [B,3,S,S]:
//   out[b,c,y,x] = (u8[b, top + (vflip ? S-1-y : y), left + (hflip ? S-1-x : x), c] / 255 - 0.5) / 0.5
// i.e. TF.crop -> TF.hflip -> TF.vflip -> TF.to_tensor -> Normalize(0.5, 0.5) (paired_data_module.py:171-199), with
// IEEE division / subtraction so that the result is bit-identical to the torchvision chain.
// mask (optional): uint8 [B,Hs,Ws] -> outm fp32 [B,1,S,S] = mask / 255 (to_tensor of a single-channel PIL image), or the
// raw byte value as a float when flags bit 1 is set (class-id masks: `torch.from_numpy(np.array(mask)).float()`,
// src/data/paired_data_multiclassmask.py:113-128).  flags bit 0: the colour bytes are BGR.
__global__ void __launch_bounds__(kEwThreads) tile_prep_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ tgt,
                                                               const uint8_t* __restrict__ mask,
                                                               const int* __restrict__ params, int B, int Hs, int Ws, int S,
                                                               int flags, float* __restrict__ out0, float* __restrict__ out1,
                                                               float* __restrict__ outm) {
    const int bgr = flags & 1, mask_raw = flags & 2;
    const long long n = (long long)B * S * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % S);
        const int y = (int)((i / S) % S);
        const int b = (int)(i / ((long long)S * S));
        const int4 pr = *reinterpret_cast<const int4*>(params + 4 * b);
        const int sy = pr.x + (pr.w ? S - 1 - y : y);
        const int sx = pr.y + (pr.z ? S - 1 - x : x);
        const size_t so = (((size_t)b * Hs + sy) * Ws + sx) * 3;
        const size_t plane = (size_t)S * S;
        const size_t oo = (size_t)b * 3 * plane + (size_t)y * S + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int sc = bgr ? 2 - c : c;
            const float a = __fsub_rn(__fdiv_rn((float)src[so + sc], 255.f), 0.5f) * 2.f;
            out0[oo + c * plane] = a;
            if (tgt != nullptr) out1[oo + c * plane] = __fsub_rn(__fdiv_rn((float)tgt[so + sc], 255.f), 0.5f) * 2.f;
        }
        if (mask != nullptr) {
            const float mv = (float)mask[((size_t)b * Hs + sy) * Ws + sx];
            outm[(size_t)b * plane + (size_t)y * S + x] = mask_raw ? mv : __fdiv_rn(mv, 255.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ Pillow 8-bit resampling
// One pass of Pillow's ImagingResample for 8-bit images (what TF.resize does to a PIL image): out = clip8((2^21 +
// sum_k in[first + k] * kk[k]) >> 22) with the integer coefficient table built on the host exactly like Pillow's
// precompute_coeffs + normalize_coeffs_8bpc.  bounds: int32 [n_out][2] = (first, count); kk: int32 [n_out][ksize].
// in: uint8 [B,Hin,Win,C]; horizontal pass -> [B,Hin,n_out,C]; vertical pass -> [B,n_out,Win,C].
__device__ __forceinline__ uint8_t clip8_22(int v) {
    v >>= 22;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
__global__ void __launch_bounds__(kEwThreads) resample_u8_h_kernel(const uint8_t* __restrict__ in, int B, int Hin, int Win,
                                                                   int C, const int* __restrict__ bounds,
                                                                   const int* __restrict__ kk, int ksize, int n_out,
                                                                   uint8_t* __restrict__ out) {
    const long long n = (long long)B * Hin * n_out * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int xo = (int)((i / C) % n_out);
        const long long row = i / ((long long)C * n_out);  // b * Hin + y
        const int first = bounds[2 * xo], cnt = bounds[2 * xo + 1];
        const uint8_t* src = in + ((size_t)row * Win + first) * C + c;
        int ss = 1 << 21;
        for (int k = 0; k < cnt; ++k) ss += (int)src[(size_t)k * C] * __ldg(kk + (size_t)xo * ksize + k);
        out[i] = clip8_22(ss);
    }
}
__global__ void __launch_bounds__(kEwThreads) resample_u8_v_kernel(const uint8_t* __restrict__ in, int B, int Hin, int Win,
                                                                   int C, const int* __restrict__ bounds,
                                                                   const int* __restrict__ kk, int ksize, int n_out,
                                                                   uint8_t* __restrict__ out) {
    const long long rowlen = (long long)Win * C;
    const long long n = (long long)B * n_out * rowlen;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long xc = i % rowlen;
        const int yo = (int)((i / rowlen) % n_out);
        const int b = (int)(i / (rowlen * n_out));
        const int first = bounds[2 * yo], cnt = bounds[2 * yo + 1];
        const uint8_t* src = in + ((size_t)b * Hin + first) * rowlen + xc;
        int ss = 1 << 21;
        for (int k = 0; k < cnt; ++k) ss += (int)src[(size_t)k * rowlen] * __ldg(kk + (size_t)yo * ksize + k);
        out[i] = clip8_22(ss);
    }
}

// ------------------------------------------------------------------------------------------------ output side
// x: fp32 NCHW [B,C,H,W] -> uint8 NHWC [B,H,W,C] = floor(clamp(x*0.5 + 0.5, 0, 1) * 255 + 0.5)
// (denormalize of src/infer_simple_flowmatching.py:37-38, then the 8-bit quantisation an image writer applies).
__global__ void __launch_bounds__(kEwThreads) denorm_u8_kernel(const float* __restrict__ x, int B, int C, int HW,
                                                               uint8_t* __restrict__ out) {
    const long long n = (long long)B * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int b = (int)(i / HW);
        for (int c = 0; c < C; ++c) {
            float v = __fadd_rn(__fmul_rn(x[((size_t)b * C + c) * HW + p], 0.5f), 0.5f);
            v = fminf(fmaxf(v, 0.f), 1.f);
            out[(size_t)i * C + c] = (uint8_t)__fadd_rn(__fmul_rn(v, 255.f), 0.5f);
        }
    }
}

}  // namespace s2s
