// Bandwidth-bound kernels of the flow-matching hot path: NHWC bf16 activations, 128-bit vector loads/stores,
// fp32 statistics.  One CTA never straddles two samples, so per-(sample, channel) coefficients live in registers.
//
// Thread mapping shared by the GroupNorm kernels: a pixel row of C channels is C/8 16-byte vectors ("vpp" <= 256).
// The block has vpp * floor(256 / vpp) threads; thread t owns vector slot (t % vpp) and walks pixels (t / vpp),
// (t / vpp) + blockDim / vpp, ... of its CTA's pixel chunk: warps touch whole contiguous pixel rows (coalesced), and
// the per-channel coefficients of the thread's 8 channels stay in registers for the whole chunk.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace s2s {

constexpr int kEwThreads = 256;

__device__ __forceinline__ void cvt8_in(const uint4& u, int fmt, float (&f)[8]) {
    float2 t;
    t = unpack2(u.x, fmt); f[0] = t.x; f[1] = t.y;
    t = unpack2(u.y, fmt); f[2] = t.x; f[3] = t.y;
    t = unpack2(u.z, fmt); f[4] = t.x; f[5] = t.y;
    t = unpack2(u.w, fmt); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 cvt8_out(const float (&f)[8], int fmt) {
    return make_uint4(pack2(f[0], f[1], fmt), pack2(f[2], f[3], fmt), pack2(f[4], f[5], fmt), pack2(f[6], f[7], fmt));
}
// compile-time format variants (the hot kernels are instantiated per format: a run-time format costs two predicated
// conversion sequences per element, and these kernels are instruction-issue bound next to the HBM roofline)
template <int F>
__device__ __forceinline__ float2 unpack2_t(uint32_t u) {
    return F == kFmtF16 ? unpack_f16x2(u) : unpack_bf16x2(u);
}
template <int F>
__device__ __forceinline__ uint32_t pack2_t(float lo, float hi) {
    return F == kFmtF16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
template <int F>
__device__ __forceinline__ void cvt8_in_t(const uint4& u, float (&f)[8]) {
    float2 t;
    t = unpack2_t<F>(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack2_t<F>(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack2_t<F>(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack2_t<F>(u.w); f[6] = t.x; f[7] = t.y;
}
template <int F>
__device__ __forceinline__ uint4 cvt8_out_t(const float (&f)[8]) {
    return make_uint4(pack2_t<F>(f[0], f[1]), pack2_t<F>(f[2], f[3]), pack2_t<F>(f[4], f[5]), pack2_t<F>(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ Philox4x32-7
// Counter-based dropout mask: the 8 consecutive elements starting at element index 8*e8 draw from counter e8, one
// 16-bit lane each (keep iff lane >= round(p * 65536): p is honoured to 2^-16).  The mask is a pure function of
// (seed, element index), so backward regenerates it instead of storing it.  7 rounds is the smallest Philox4x32
// variant that passes BigCrush; one call per 8 elements keeps the integer work at ~7 ops / element, inside the
// HBM-roofline instruction budget of the normalisation kernels (~20 lane-ops per 4 bytes moved).
__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t c2 = 0x243F6A88u, c3 = 0x85A308D3u;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t dropout_thresh16(float p) { return min((uint32_t)(p * 65536.0f + 0.5f), 65535u); }
// keep-mask bits for 8 consecutive elements starting at element index e8*8 (bit j = keep element j)
__device__ __forceinline__ uint32_t dropout_keep8(unsigned long long seed, unsigned long long e8, uint32_t thresh16) {
    const uint4 r = philox4x32_7((uint32_t)e8, (uint32_t)(e8 >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t m = 0;
    m |= ((r.x & 0xffffu) >= thresh16) << 0; m |= ((r.x >> 16) >= thresh16) << 1;
    m |= ((r.y & 0xffffu) >= thresh16) << 2; m |= ((r.y >> 16) >= thresh16) << 3;
    m |= ((r.z & 0xffffu) >= thresh16) << 4; m |= ((r.z >> 16) >= thresh16) << 5;
    m |= ((r.w & 0xffffu) >= thresh16) << 6; m |= ((r.w >> 16) >= thresh16) << 7;
    return m;
}

// the same keep bits, applied to 8 fp32 values on the way (the compare predicates select the values directly; building
// the byte first and testing its bits again costs two more integer ops per element)
__device__ __forceinline__ uint32_t dropout_apply8(unsigned long long seed, unsigned long long e8, uint32_t thresh16,
                                                   float (&f)[8]) {
    const uint4 r = philox4x32_7((uint32_t)e8, (uint32_t)(e8 >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // high lane: w >= T * 2^16  <=>  (w >> 16) >= T exactly, no extraction; T <= 65535 here (p < 1)
        const bool k0 = (w[j] & 0xffffu) >= thresh16, k1 = w[j] >= (thresh16 << 16);
        f[2 * j] = k0 ? f[2 * j] : 0.f;
        f[2 * j + 1] = k1 ? f[2 * j + 1] : 0.f;
        m |= (k0 ? 1u : 0u) << (2 * j);
        m |= (k1 ? 1u : 0u) << (2 * j + 1);
    }
    return m;
}

// ------------------------------------------------------------------------------------------------ weight packing
// OIHW fp32 -> 16-bit [rows_pad][ld_k] K-major operand of the implicit GEMM.
//   forward : dst[co][k_off + tap*ci_count + (ci-ci_begin)]           = w[co][ci][tap]          rows = Cout
//   dgrad   : dst[ci-ci_begin][k_off + tap*Cout + co]                 = w[co][ci][taps-1-tap]   rows = ci_count
// mode = 1 + phase (phase = py*2 + px), 3x3 weights only -- the phase-decomposed Upsample conv: nearest x2 followed by a
// 3x3/pad-1 conv is, for the output pixels (2y+py, 2x+px), a 2x2 conv over the LOW-resolution input whose taps are sums
// of the 3x3 taps that land on the same low-resolution pixel.  Logical tap ti = ai*2 + bi covers the filter rows
// R(py, ai) and columns R(px, bi) with R(0,0) = {0}, R(0,1) = {1,2}, R(1,0) = {0,1}, R(1,1) = {2} (summed in fp32
// before rounding to 16 bit):
//   forward : dst[co][k_off + ti*ci_count + (ci-ci_begin)]  = sum w[co][ci][R x R]
//   dgrad   : dst[ci-ci_begin][k_off + ti*Cout + co]        = the same sum (tap positions are mirrored by the tap map
//                                                              of the launch, not by the packing)
struct PackJob {
    const float* w;
    uint16_t* dst;
    int Cout, Cin, taps, ci_begin, ci_count, ld_k, k_off, transpose_flip, fmt, mode;
};
__device__ __forceinline__ void phase_range(int parity, int idx, int& lo, int& hi) {  // R(parity, idx) = [lo, hi]
    if (parity == 0) { lo = idx == 0 ? 0 : 1; hi = idx == 0 ? 0 : 2; }
    else             { lo = idx == 0 ? 0 : 2; hi = idx == 0 ? 1 : 2; }
}
// Work decomposition: tiles of kPackRows destination rows x kPackInner destination-contiguous elements x all (logical)
// taps.  The source of a tile is a set of CONTIGUOUS runs of the OIHW tensor either way -- forward: per output channel the
// 64 input channels x taps of the tile; transposed: per output channel (64 of them) the 16 input channels x taps -- so
// a warp stages one run at a time with coalesced loads into shared memory and the tile leaves as 128-byte rows.  (The
// element-per-thread kernel this replaces read the transposed operand with one cache line per lane: 0.14 of the HBM rate
// for the 158 operands of a training step.)
constexpr int kPackRows = 16, kPackInner = 64;
constexpr int kPackSmemFloats = kPackInner * (kPackRows * 9 + 1);  // >= kPackRows * (kPackInner * 9 + 1)
__host__ __device__ inline int pack_tile_count(int Cout, int ci_count, int transpose_flip) {
    const int rows = transpose_flip ? ci_count : Cout, inner = transpose_flip ? Cout : ci_count;
    return ((rows + kPackRows - 1) / kPackRows) * ((inner + kPackInner - 1) / kPackInner);
}
__device__ __forceinline__ void pack_conv_weight_tile(const PackJob& jb, int tile, float* S) {
    const bool tf = jb.transpose_flip != 0;
    const int rows = tf ? jb.ci_count : jb.Cout, inner = tf ? jb.Cout : jb.ci_count;
    const int tiles_i = (inner + kPackInner - 1) / kPackInner;
    const int r0 = (tile / tiles_i) * kPackRows, i0 = (tile % tiles_i) * kPackInner;
    if (r0 >= rows) return;  // (uniform per CTA)
    const int nr = min(kPackRows, rows - r0), ni = min(kPackInner, inner - i0);
    const int taps = jb.taps, ltaps = jb.mode == 0 ? taps : 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int pitch = (tf ? kPackRows : kPackInner) * taps + 1;  // odd: conflict-poor strided reads below
    // forward   : run r = output channel r0 + r, input channels [ci_begin + i0, + ni) x taps
    // transposed: run c = output channel i0 + c, input channels [ci_begin + r0, + nr) x taps
    const int n_runs = tf ? ni : nr, run_len = (tf ? nr : ni) * taps;
    const size_t run0 = tf ? ((size_t)i0 * jb.Cin + jb.ci_begin + r0) * taps : ((size_t)r0 * jb.Cin + jb.ci_begin + i0) * taps;
    const size_t run_stride = (size_t)jb.Cin * taps;
    const bool vec_ok = ((run_len | (int)(run_stride & 3) | (int)(run0 & 3)) & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(jb.w) & 15) == 0;  // every run 16-byte aligned, whole float4s
    if (vec_ok) {
        const int per = run_len >> 2;  // float4s per run; (run, j) pairs flattened so that all 256 threads load
        for (int idx = threadIdx.x; idx < n_runs * per; idx += blockDim.x) {
            const int run = idx / per, j = idx - run * per;
            const float4 v = __ldg(reinterpret_cast<const float4*>(jb.w + run0 + (size_t)run * run_stride) + j);
            float* d = S + run * pitch + 4 * j;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int run = warp; run < n_runs; run += nwarps) {
            const float* src = jb.w + run0 + (size_t)run * run_stride;
            for (int j = lane; j < run_len; j += 32) S[run * pitch + j] = __ldg(src + j);
        }
    }
    __syncthreads();
    auto value = [&](int r, int lt, int il) -> float {
        const float* sp = tf ? S + il * pitch + r * taps : S + r * pitch + il * taps;
        if (jb.mode == 0) return sp[tf ? (taps - 1 - lt) : lt];
        const int phase = jb.mode - 1;
        int y0, y1, x0, x1;
        phase_range(phase >> 1, lt >> 1, y0, y1);
        phase_range(phase & 1, lt & 1, x0, x1);
        float v = 0.f;
        for (int dy = y0; dy <= y1; ++dy)
            for (int dx = x0; dx <= x1; ++dx) v += sp[dy * 3 + dx];
        return v;
    };
    const bool pair_ok = ((jb.ld_k | jb.k_off | inner) & 1) == 0;  // 32-bit stores of two elements
    for (int item = warp; item < nr * ltaps; item += nwarps) {
        const int r = item / ltaps, lt = item - r * ltaps;
        uint16_t* drow = jb.dst + (size_t)(r0 + r) * jb.ld_k + (size_t)jb.k_off + (size_t)lt * inner + i0;
        for (int il = 2 * lane; il < ni; il += 64) {
            const uint16_t a = pack1(value(r, lt, il), jb.fmt);
            if (il + 1 < ni) {
                const uint16_t c = pack1(value(r, lt, il + 1), jb.fmt);
                if (pair_ok) *reinterpret_cast<uint32_t*>(drow + il) = (uint32_t)a | ((uint32_t)c << 16);
                else { drow[il] = a; drow[il + 1] = c; }
            } else {
                drow[il] = a;
            }
        }
    }
}
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(PackJob jb) {
    __shared__ float S[kPackSmemFloats];
    pack_conv_weight_tile(jb, blockIdx.x, S);
}

// The same packing for MANY weights in one launch (after an optimizer step every cached GEMM operand is stale: 158
// tensors per training step).  Host uploads a job table + (job, tile) work list, as for the multi-tensor Adam.
__global__ void __launch_bounds__(256) pack_conv_weight_multi_kernel(const PackJob* __restrict__ jobs,
                                                                     const int2* __restrict__ work) {
    __shared__ float S[kPackSmemFloats];
    const int2 wi = work[blockIdx.x];
    const PackJob jb = jobs[wi.x];
    pack_conv_weight_tile(jb, wi.y, S);
}

// Weight gradient of the phase-decomposed Upsample conv: src = [16 = phase*4 + ti][M][N] fp32 (what the four phase wgrad
// launches accumulated) -> OIHW [M][N][9]: each 3x3 tap belongs, in every phase, to exactly one logical tap, so
// dW[m][n][dy][dx] = sum_phase src[phase][ti(phase, dy, dx)][m][n].
__global__ void upconv_unpack_wgrad_kernel(const float* __restrict__ src, int M, int N, float* __restrict__ grad) {
    const long long total = (long long)M * N * 9;
    const size_t plane = (size_t)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(i % 9);
        const long long mn = i / 9;
        const int dy = tap / 3, dx = tap % 3;
        float v = 0.f;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int py = ph >> 1, px = ph & 1;
            const int ai = py == 0 ? (dy == 0 ? 0 : 1) : (dy == 2 ? 1 : 0);
            const int bi = px == 0 ? (dx == 0 ? 0 : 1) : (dx == 2 ? 1 : 0);
            v += src[(size_t)(ph * 4 + ai * 2 + bi) * plane + mn];
        }
        grad[i] = v;
    }
}

// [taps][M][ldn] fp32 wgrad buffer -> += into OIHW fp32 gradient.  dst[m][n_begin + n][tap] += src[tap][m][n_off + n]
// One thread per (m, n): its `taps` reads are coalesced across the warp (n fastest), its `taps` consecutive floats of the
// OIHW row join the neighbours' into one contiguous run per warp.
__global__ void unpack_wgrad_kernel(const float* __restrict__ src, int taps, int M, int ldn, int n_off, int n_count,
                                    float* __restrict__ grad, int Cin_total, int n_begin, float beta) {
    const long long total = (long long)M * n_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % n_count);
        const int m = (int)(i / n_count);
        float* g = grad + ((size_t)m * Cin_total + n_begin + n) * taps;
        const float* sp = src + (size_t)m * ldn + n_off + n;
        if (taps == 9) {
            float v[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) v[t] = sp[(size_t)t * M * ldn];
#pragma unroll
            for (int t = 0; t < 9; ++t) g[t] = (beta == 0.f) ? v[t] : fmaf(beta, g[t], v[t]);
        } else {
            for (int t = 0; t < taps; ++t) {
                const float v = sp[(size_t)t * M * ldn];
                g[t] = (beta == 0.f) ? v : fmaf(beta, g[t], v);  // beta == 0: the destination may be uninitialised
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ 3x3 patch pack
// fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H,W,64] with channel j = tap*3 + c holding src[b, c, y + sgn*(dy-1),
// x + sgn*(dx-1)] (zero outside the image, zeros for j >= 27).  Turns the K=27 stem conv (sgn=+1) and the dgrad/wgrad
// of the N=3 head conv (sgn=-1) into 64-wide GEMM operands.  Optional fused flow-matching interpolation:
// src = (1 - t_b) * x0 + t_b * x1  (torchcfm sample_xt with sigma = 0).
// (A thread-per-vector mapping with coalesced 512-byte stores was measured 3x SLOWER: the 27 scattered source reads of a
// pixel then spread over four lanes and the kernel became load/store-unit bound -- 0.10 vs 0.33 of the HBM rate.)
__global__ void patch27_pack_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                    const float* __restrict__ t, int B, int H, int W, int sgn,
                                    __nv_bfloat16* __restrict__ dst, float* __restrict__ xt_out, int fmt) {
    const long long npix = (long long)B * H * W;
    for (long long pidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; pidx < npix;
         pidx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pidx % W);
        const int y = (int)((pidx / W) % H);
        const int b = (int)(pidx / ((long long)W * H));
        const float tb = (x1 != nullptr) ? t[b] : 0.f;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + sgn * (tap / 3 - 1), xx = x + sgn * (tap % 3 - 1);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const size_t o = (((size_t)b * 3 + c) * H + yy) * W + xx;
                    float s = __ldg(x0 + o);
                    if (x1 != nullptr) s = (1.f - tb) * s + tb * __ldg(x1 + o);
                    v[tap * 3 + c] = s;
                    if (xt_out != nullptr && tap == 4) xt_out[o] = s;
                }
            }
        }
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)pidx * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = v[j * 8 + e];
            d[j] = cvt8_out(f, fmt);
        }
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 4; j < 8; ++j) d[j] = z;
    }
}

// Addressing of the streaming loops.  A thread walks pixels p, p + pstep, ... of its CTA's chunk at a fixed vector slot, so
// every tensor is ONE byte pointer advanced by a constant step (2 integer ops per tensor and iteration instead of a 64-bit
// multiply-add chain per access: these kernels sit next to the instruction-issue roofline once the SM clock drops under
// the power cap of a training step).  kFull (blockDim.x == 256 and every tensor has row length C): the thread's next pixel
// row is a COMPILE-TIME 256 vectors = 4096 B away (its keep byte 256 B), so the second access of an unrolled iteration is
// an immediate offset of the same pointer.
constexpr int kRowStepBytes = kEwThreads * 16;
template <bool kFull>
__device__ __forceinline__ const char* row_next(const char* q, size_t step) {
    return kFull ? q + kRowStepBytes : q + step;
}
template <bool kFull>
__device__ __forceinline__ char* row_next(char* q, size_t step) {
    return kFull ? q + kRowStepBytes : q + step;
}
template <bool kFull>
__device__ __forceinline__ const uint8_t* mask_next(const uint8_t* q, size_t step) {
    return kFull ? q + kEwThreads : q + step;
}
__device__ __forceinline__ uint4 ldg_stream_b(const char* q) { return ldg_stream(reinterpret_cast<const uint4*>(q)); }
__device__ __forceinline__ void stg_stream_b(char* q, const uint4& v) { stg_stream(reinterpret_cast<uint4*>(q), v); }

// ------------------------------------------------------------------------------------------------ GroupNorm stats
constexpr int kGnUnroll = 4;
enum Act : int { kActNone = 0, kActSilu = 1, kActRelu = 2 };  // activation fused behind the normalisation  // independent 16 B loads in flight per thread and per input tensor

// stats[b][chunk][c_off + c] = (sum, sumsq) over the pixels of one chunk of the sample (no atomics: deterministic).
template <int XF>
__global__ void __launch_bounds__(kEwThreads, 4) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int C, int HW,
                                                                 int pix_per_cta, float2* __restrict__ stats, int Ctot,
                                                                 int c_off) {
    __shared__ float red[kEwThreads][17];
    const int vpp = C >> 3;
    const int slot = threadIdx.x % vpp, prow = threadIdx.x / vpp, pstep = blockDim.x / vpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)b * HW * C);
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
    int p = p0 + prow;
    for (; p + (kGnUnroll - 1) * pstep < p1; p += kGnUnroll * pstep) {
        uint4 u[kGnUnroll];
#pragma unroll
        for (int i = 0; i < kGnUnroll; ++i) u[i] = ldg_stream(src + (size_t)(p + i * pstep) * vpp + slot);
#pragma unroll
        for (int i = 0; i < kGnUnroll; ++i) {
            float f[8];
            cvt8_in_t<XF>(u[i], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s[e] += f[e];
                ss[e] = fmaf(f[e], f[e], ss[e]);
            }
        }
    }
    for (; p < p1; p += pstep) {
        float f[8];
        cvt8_in_t<XF>(ldg_stream(src + (size_t)p * vpp + slot), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            s[e] += f[e];
            ss[e] = fmaf(f[e], f[e], ss[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[threadIdx.x][e] = s[e];
        red[threadIdx.x][8 + e] = ss[e];
    }
    __syncthreads();
    for (int ci = threadIdx.x; ci < C; ci += blockDim.x) {
        const int sl = ci >> 3, e = ci & 7;
        float a = 0.f, q = 0.f;
        for (int r = 0; r < pstep; ++r) {
            a += red[r * vpp + sl][e];
            q += red[r * vpp + sl][8 + e];
        }
        // deterministic: one partial per (sample, pixel chunk); gn_coef_kernel sums the chunks in order
        stats[((size_t)b * gridDim.x + blockIdx.x) * Ctot + c_off + ci] = make_float2(a, q);
    }
}

// Per-(sample, channel) affine coefficients of the fused normalisation:
//   y = silu?( x * A + Bc ),  A = rstd_g * gamma_c * (1 + scale_bc),  Bc = (beta_c - mean_g*rstd_g*gamma_c)*(1+scale_bc) + shift_bc
// Also records (mean, rstd) per (sample, group) for backward.  film: [B][2C] fp32 (scale | shift) or nullptr.
// One CTA per sample; dynamic smem = 2 * C floats (per-channel totals).
__global__ void __launch_bounds__(256) gn_coef_kernel(const float2* __restrict__ stats, int nchunks,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ film, int C, int G, int HW, float eps,
                                                      float2* __restrict__ coef, float2* __restrict__ mean_rstd) {
    extern __shared__ float s_tot[];  // [C] sums, [C] sums of squares
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.x;
    const int cpg = C / G;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float2* sp = stats + (size_t)b * nchunks * C + c;
        float a = 0.f, q = 0.f;
        int k = 0;
        for (; k + 4 <= nchunks; k += 4) {  // fixed summation order (deterministic), 4 loads in flight
            const float2 t0 = sp[(size_t)(k + 0) * C], t1 = sp[(size_t)(k + 1) * C];
            const float2 t2 = sp[(size_t)(k + 2) * C], t3 = sp[(size_t)(k + 3) * C];
            a += t0.x; q += t0.y; a += t1.x; q += t1.y; a += t2.x; q += t2.y; a += t3.x; q += t3.y;
        }
        for (; k < nchunks; ++k) {
            const float2 t = sp[(size_t)k * C];
            a += t.x; q += t.y;
        }
        s_tot[c] = a;
        s_tot[C + c] = q;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float a = 0.f, q = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            a += s_tot[c];
            q += s_tot[C + c];
        }
        const float n = (float)cpg * (float)HW;
        const float mean = a / n;
        const float var = fmaxf(q / n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        s_mean[g] = mean;
        s_rstd[g] = rstd;
        mean_rstd[(size_t)b * G + g] = make_float2(mean, rstd);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float ga = gamma[c], be = beta[c];
        float A = s_rstd[g] * ga;
        float Bc = be - s_mean[g] * A;
        if (film != nullptr) {
            const float sc = 1.f + film[(size_t)b * 2 * C + c];
            const float sh = film[(size_t)b * 2 * C + C + c];
            A *= sc;
            Bc = Bc * sc + sh;
        }
        coef[(size_t)b * C + c] = make_float2(A, Bc);
    }
}

// Same fold for partial statistics that arrive per SOURCE of a channel concat (written by the producing convs'
// epilogues): source i has its own [B][nchunks][Ci] buffer.  Channel c of the concat is channel c of source 0 for
// c < C0, else channel c - C0 of source 1.
// One CTA per sample.  The fold over the (up to 512) sub-tile partials is latency-bound if every channel walks its
// column alone (measured 40 us per call at 256^2 -- 6 % of a sampling evaluation): the chunk range is cut into `slices`
// walked by different threads, two channels (one 16-byte load) per thread, eight loads in flight, and the slice partials
// are folded in a fixed order in shared memory (deterministic).
__global__ void __launch_bounds__(1024) gn_coef_parts_kernel(const float2* __restrict__ stats0, int C0,
                                                             const float2* __restrict__ stats1, int C1, int nchunks,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ film, int G, int HW, float eps,
                                                             float2* __restrict__ coef, float2* __restrict__ mean_rstd,
                                                             int slices) {
    extern __shared__ float s_dyn[];            // [2C] totals, then [slices][2C] slice partials
    float* s_tot = s_dyn;
    float* s_part = s_dyn + 2 * (C0 + C1);
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.x;
    const int C = C0 + C1;
    const int cpg = C / G;
    const int half = C >> 1;                    // channel pairs (C0, C1 even)
    const int per = (nchunks + slices - 1) / slices;
    for (int w = threadIdx.x; w < half * slices; w += blockDim.x) {
        const int pair = w % half, sl = w / half;
        const int c = 2 * pair;
        const bool first = c < C0;
        const int Ci = first ? C0 : C1;
        const float4* sp = reinterpret_cast<const float4*>((first ? stats0 : stats1) + (size_t)b * nchunks * Ci +
                                                           (first ? c : c - C0));
        const size_t stride = (size_t)Ci >> 1;  // float4 units per chunk row
        const int k0 = sl * per, k1 = min(nchunks, k0 + per);
        float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
        int k = k0;
        for (; k + 8 <= k1; k += 8) {
            float4 t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __ldg(sp + (size_t)(k + i) * stride);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a0 += t[i].x; q0 += t[i].y; a1 += t[i].z; q1 += t[i].w;
            }
        }
        for (; k < k1; ++k) {
            const float4 t = __ldg(sp + (size_t)k * stride);
            a0 += t.x; q0 += t.y; a1 += t.z; q1 += t.w;
        }
        float* dst = s_part + (size_t)sl * 2 * C;
        dst[c] = a0; dst[c + 1] = a1; dst[C + c] = q0; dst[C + c + 1] = q1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += s_part[(size_t)sl * 2 * C + i];
        s_tot[i] = v;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float a = 0.f, q = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            a += s_tot[c];
            q += s_tot[C + c];
        }
        const float n = (float)cpg * (float)HW;
        const float mean = a / n;
        const float var = fmaxf(q / n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        s_mean[g] = mean;
        s_rstd[g] = rstd;
        mean_rstd[(size_t)b * G + g] = make_float2(mean, rstd);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float A = s_rstd[g] * gamma[c];
        float Bc = beta[c] - s_mean[g] * A;
        if (film != nullptr) {
            const float sc = 1.f + film[(size_t)b * 2 * C + c];
            const float sh = film[(size_t)b * 2 * C + C + c];
            A *= sc;
            Bc = Bc * sc + sh;
        }
        coef[(size_t)b * C + c] = make_float2(A, Bc);
    }
}

// y[b, p, c_off + c] = dropout( silu( x[b, p, c] * A + Bc ) ), 16-bit NHWC in / out (out row stride ld_out channels).
// kDual: the same values are also written as bf16 to y2 (same geometry) -- the operand the weight-gradient GEMM of the
// consuming conv needs (its MMA cannot mix fp16 x bf16), produced here for +2 B/element instead of a 4 B/element
// conversion pass in backward.
template <int kAct, bool kDrop, int XF, int YF, bool kDual, int kU>
__global__ void __launch_bounds__(kEwThreads, kU > 4 ? 3 : 4) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int C, int HW,
                                                                 int pix_per_cta, const float2* __restrict__ coef,
                                                                 int Ctot, int c_off, __nv_bfloat16* __restrict__ y,
                                                                 __nv_bfloat16* __restrict__ y2, int ld_out,
                                                                 float drop_p, unsigned long long seed,
                                                                 uint8_t* __restrict__ mask_out,
                                                                 const unsigned long long* __restrict__ seed_step) {
    // seed_step (optional): a device-resident step counter mixed into the seed, so that a launch replayed from a CUDA
    // graph (whose host-side seed argument is frozen at capture) still draws a fresh mask every training step
    if (kDrop && seed_step != nullptr) seed += __ldg(seed_step) * 0x9E3779B97F4A7C15ull;
    const int vpp = C >> 3;
    const int slot = threadIdx.x % vpp, prow = threadIdx.x / vpp, pstep = blockDim.x / vpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    // dropout: the 1/(1-p) factor of the kept elements is folded into the affine coefficients (act(z) * k = zs * sigmoid(zs / k)
    // for SiLU with zs = k z -- only the exponent's constant changes --, relu / identity commute with k > 0): no multiply
    // per element behind the activation
    const uint32_t thresh = kDrop ? dropout_thresh16(drop_p) : 0u;
    const float keep_scale = kDrop ? 1.f / (1.f - drop_p) : 1.f;
    const float silu_c = kDrop ? -1.4426950408889634f * (1.f - drop_p) : -1.4426950408889634f;
    float A[8], Bc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float2 cf = coef[(size_t)b * Ctot + c_off + slot * 8 + e];
        A[e] = cf.x * keep_scale;
        Bc[e] = cf.y * keep_scale;
    }
    const unsigned long long e8_base = (unsigned long long)b * HW * (unsigned long long)(ld_out >> 3) +
                                       (unsigned long long)((c_off >> 3) + slot);
    // one byte pointer per tensor, advanced by a constant step; kFull: the unrolled accesses are immediate offsets
    // (see row_next)
    auto run = [&](auto full_c) {
        constexpr bool kFull = decltype(full_c)::value;
        const size_t pix0 = (size_t)b * HW + (size_t)(p0 + prow);
        const char* xp = reinterpret_cast<const char*>(x) + (pix0 * C + slot * 8) * 2;
        char* yp = reinterpret_cast<char*>(y) + (pix0 * ld_out + c_off + slot * 8) * 2;
        char* y2p = kDual ? reinterpret_cast<char*>(y2) + (pix0 * ld_out + c_off + slot * 8) * 2 : nullptr;
        const size_t sx = (size_t)pstep * C * 2, sy = (size_t)pstep * ld_out * 2;
        unsigned long long e8 = e8_base + (unsigned long long)(p0 + prow) * (unsigned long long)(ld_out >> 3);
        const unsigned long long se8 = (unsigned long long)pstep * (unsigned long long)(ld_out >> 3);
        auto body = [&](const uint4& u, char* dst, char* dst2) {
            float f[8];
            cvt8_in_t<XF>(u, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float z = fmaf(f[e], A[e], Bc[e]);  // kDrop: z * 1/(1-p), see the coefficient load
                f[e] = kAct == kActSilu ? silu_scaled_f(z, silu_c) : (kAct == kActRelu ? fmaxf(z, 0.f) : z);
            }
            if (kDrop) {
                const uint32_t m = dropout_apply8(seed, e8, thresh, f);
                if (mask_out != nullptr) mask_out[e8] = (uint8_t)m;  // 1 bit / element: backward reads it instead of re-hashing
                e8 += se8;
            }
            stg_stream_b(dst, cvt8_out_t<YF>(f));
            if (kDual) stg_stream_b(dst2, cvt8_out_t<kFmtBF16>(f));
        };
        int p = p0 + prow;
        for (; p + (kU - 1) * pstep < p1; p += kU * pstep) {
            uint4 u[kU];
            const char* q = xp;
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                u[i] = ldg_stream_b(q);
                q = row_next<kFull>(q, sx);
            }
            xp = q;
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                body(u[i], yp, y2p);
                yp = row_next<kFull>(yp, sy);
                if (kDual) y2p = row_next<kFull>(y2p, sy);
            }
        }
        for (; p < p1; p += pstep) {
            body(ldg_stream_b(xp), yp, y2p);
            xp = row_next<kFull>(xp, sx);
            yp = row_next<kFull>(yp, sy);
            if (kDual) y2p = row_next<kFull>(y2p, sy);
        }
    };
    if (blockDim.x == kEwThreads && ld_out == C) run(std::true_type{});
    else run(std::false_type{});
}

// ------------------------------------------------------------------------------------------------ GroupNorm backward
// Forward: z = x*A + Bc, a = dropout(silu(z)).  Given g = dL/da:  dz = g * mask/keep * silu'(z).
// The sigmoid inside silu' uses one tanh.approx (relative error 2^-11, far below the bf16 gradient storage); the
// forward keeps the two-MUFU exact form because its result is stored with an 11-bit significand.
// Pass 1 (this kernel): per chunk, (sum_p dz, sum_p dz * xhat) with xhat = (x - mean_g) * rstd_g, accumulated as
// (sum dz, sum dz*x) in the streaming loop and centred once per channel at the end.
// Per-thread channel coefficients of z = x*A + Bc.  fp16 activations: held as 4 half2 pairs of (A/2, Bc/2) (u = z/2 and
// silu' are then evaluated with packed half2 arithmetic and ONE tanh.approx.f16x2 per two elements -- z only feeds silu',
// whose result multiplies a bf16 gradient, so 11 significant bits are ample); otherwise 8 + 8 fp32 values.
template <bool kHalf>
struct GnZCoef {
    float A[8], Bc[8];
    __device__ __forceinline__ void load(const float2* __restrict__ cf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float2 c = cf[e];
            A[e] = c.x;
            Bc[e] = c.y;
        }
    }
};
template <>
struct GnZCoef<true> {
    __half2 A[4], Bc[4];  // halved: u = x * (A/2) + Bc/2 = z/2 (exact scaling by a power of two)
    __device__ __forceinline__ void load(const float2* __restrict__ cf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 c0 = cf[2 * e], c1 = cf[2 * e + 1];
            A[e] = __floats2half2_rn(0.5f * c0.x, 0.5f * c1.x);
            Bc[e] = __floats2half2_rn(0.5f * c0.y, 0.5f * c1.y);
        }
    }
};
// packed-half2 evaluation of silu' is used for fp16 activations behind a SiLU; everything else keeps fp32 coefficients
template <int kAct, int XF>
__host__ __device__ constexpr bool gn_half_path() { return kAct == kActSilu && XF == kFmtF16; }
__device__ __forceinline__ __half2 tanh_approx_h2(__half2 x) {
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<uint32_t*>(&x)));
    return *reinterpret_cast<__half2*>(&r);
}
// silu'(z) for two elements in half2 from u = z/2, t = tanh(u):  sigmoid = (1 + t)/2 and sigmoid (1 - sigmoid) = (1 - t^2)/4, so
//   silu'(z) = sigmoid + z sigmoid (1 - sigmoid) = 0.5 (1 + t + u (1 - t^2))       -- three HFMA2 behind the tanh
__device__ __forceinline__ __half2 silu_grad_h2_u(__half2 u) {
    const __half2 half = __float2half2_rn(0.5f), one = __float2half2_rn(1.0f);
    const __half2 t = tanh_approx_h2(u);
    const __half2 w = __hfma2(__hneg2(t), t, one);
    return __hfma2(__hfma2(u, w, t), half, half);
}

// xf = x as fp32, dz = g * keep/(1-p) * act'(x*A + Bc) for 8 consecutive channels of one pixel
// (the 1/(1-p) factor of the kept elements is NOT applied here: it is linear, the callers fold it into their per-channel
// sums / coefficients -- one multiply per channel instead of one per element)
template <int kAct, bool kDrop, int XF, int GF>
__device__ __forceinline__ void gn_dz8(const uint4& xu, const uint4& gu, const GnZCoef<gn_half_path<kAct, XF>()>& cf,
                                       uint32_t keep, float (&xf)[8], float (&dz)[8]) {
    cvt8_in_t<XF>(xu, xf);
    cvt8_in_t<GF>(gu, dz);
    if (kDrop) {
#pragma unroll
        for (int e = 0; e < 8; ++e) dz[e] = (keep & (1u << e)) ? dz[e] : 0.f;  // R2P + one FSEL per element
    }
    if constexpr (kAct == kActSilu) {
        if constexpr (XF == kFmtF16) {
            const uint32_t xs[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half2 u = __hfma2(*reinterpret_cast<const __half2*>(&xs[e]), cf.A[e], cf.Bc[e]);
                const float2 gr = __half22float2(silu_grad_h2_u(u));
                dz[2 * e] *= gr.x;
                dz[2 * e + 1] *= gr.y;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) dz[e] *= silu_grad_fast(fmaf(xf[e], cf.A[e], cf.Bc[e]));
        }
    } else if constexpr (kAct == kActRelu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) dz[e] = fmaf(xf[e], cf.A[e], cf.Bc[e]) > 0.f ? dz[e] : 0.f;  // same fp32 z as forward
    }
}

// kDrop: 0 = no dropout, 1 = keep bits re-generated from the Philox counter, 2 = keep bits read from the stored mask (what the
// ResBlock node uses).  A compile-time choice: with both paths in one kernel the Philox code's registers made the 64-register
// variant spill in its main loop (profiles/r02_ncu_norm_attention_summary.txt).
// kX16: also write x as bf16 (x_bf16_out).  A template argument because the extra pointer and store pushed the 64-register
// main loop into local-memory spills; the variant with the side product runs at 3 CTAs per SM instead.
template <int kAct, int kDrop, int XF, int GF, bool kX16, int kU>
__global__ void __launch_bounds__(kEwThreads, (kX16 || kU > 2) ? 3 : 4) gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      const __nv_bfloat16* __restrict__ g, int ld_g,
                                                                      int C, int HW, int pix_per_cta,
                                                                      const float2* __restrict__ coef,
                                                                      const float2* __restrict__ mean_rstd, int G,
                                                                      int Ctot, int c_off, float2* __restrict__ red_out,
                                                                      float drop_p, unsigned long long seed,
                                                                      const uint8_t* __restrict__ mask_in,
                                                                      __nv_bfloat16* __restrict__ x_bf16_out) {
    __shared__ float red[kEwThreads][17];
    const int vpp = C >> 3;
    const int slot = threadIdx.x % vpp, prow = threadIdx.x / vpp, pstep = blockDim.x / vpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    GnZCoef<gn_half_path<kAct, XF>()> cf;
    cf.load(coef + (size_t)b * Ctot + c_off + slot * 8);
    const uint32_t thresh = kDrop ? dropout_thresh16(drop_p) : 0u;
    const float keep_scale = kDrop ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long e8_base = (unsigned long long)b * HW * (unsigned long long)(ld_g >> 3) +
                                       (unsigned long long)((c_off >> 3) + slot);
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
    // stored keep bits: the byte is fetched TOGETHER with the x / g vectors of its pixel (a dependent 1-byte load inside
    // the body exposed a full memory latency per pixel: the dropout variants ran at 0.53 of HBM peak, issue-stalled)
    constexpr bool stored = kDrop == 2;
    constexpr bool want_x16 = kX16;
    auto run = [&](auto full_c) {
        constexpr bool kFull = decltype(full_c)::value;
        const size_t pix0 = (size_t)b * HW + (size_t)(p0 + prow);
        const char* xp = reinterpret_cast<const char*>(x) + (pix0 * C + slot * 8) * 2;
        const char* gp = reinterpret_cast<const char*>(g) + (pix0 * ld_g + c_off + slot * 8) * 2;
        const uint8_t* mp = stored ? mask_in + e8_base + (size_t)(p0 + prow) * (size_t)(ld_g >> 3) : nullptr;
        char* xo = want_x16 ? reinterpret_cast<char*>(x_bf16_out) + (pix0 * C + slot * 8) * 2 : nullptr;
        const size_t sx = (size_t)pstep * C * 2, sg = (size_t)pstep * ld_g * 2, sm = (size_t)pstep * (ld_g >> 3);
        auto body = [&](const uint4& xu, const uint4& gu, uint32_t m, int p, char* xo_p) {
            if (kDrop == 1)
                m = dropout_keep8(seed, e8_base + (unsigned long long)p * (unsigned long long)(ld_g >> 3), thresh);
            float xf[8], dz[8];
            gn_dz8<kAct, kDrop != 0, XF, GF>(xu, gu, cf, m, xf, dz);
            // optional side product: x in bf16 (the weight-gradient operand of a 1x1 skip conv over the raw block input)
            // -- +2 B/element here instead of a 4 B/element conversion pass
            if (want_x16) stg_stream_b(xo_p, cvt8_out_t<kFmtBF16>(xf));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s1[e] += dz[e];
                s2[e] = fmaf(dz[e], xf[e], s2[e]);
            }
        };
        int p = p0 + prow;
        for (; p + (kU - 1) * pstep < p1; p += kU * pstep) {
            uint4 xu[kU], gu[kU];
            uint32_t mk[kU];
#pragma unroll
            for (int i = 0; i < kU; ++i) {  // kFull: the i-th row is an immediate offset of the same pointer
                xu[i] = ldg_stream_b(xp);
                gu[i] = ldg_stream_b(gp);
                mk[i] = stored ? (uint32_t)__ldg(mp) : 0xffu;
                xp = row_next<kFull>(xp, sx);
                gp = row_next<kFull>(gp, sg);
                if (stored) mp = mask_next<kFull>(mp, sm);
            }
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                body(xu[i], gu[i], mk[i], p + i * pstep, xo);
                if (want_x16) xo = row_next<kFull>(xo, sx);
            }
        }
        for (; p < p1; p += pstep) {
            body(ldg_stream_b(xp), ldg_stream_b(gp), stored ? (uint32_t)__ldg(mp) : 0xffu, p, xo);
            xp = row_next<kFull>(xp, sx);
            gp = row_next<kFull>(gp, sg);
            if (stored) mp = mask_next<kFull>(mp, sm);
            if (want_x16) xo = row_next<kFull>(xo, sx);
        }
    };
    if (blockDim.x == kEwThreads && ld_g == C) run(std::true_type{});
    else run(std::false_type{});
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[threadIdx.x][e] = s1[e] * keep_scale;  // dz of the kept elements carries 1/(1-p): applied once per channel here
        red[threadIdx.x][8 + e] = s2[e] * keep_scale;
    }
    __syncthreads();
    const int cpg = Ctot / G;
    for (int ci = threadIdx.x; ci < C; ci += blockDim.x) {
        const int sl = ci >> 3, e = ci & 7;
        float a = 0.f, q = 0.f;
        for (int r = 0; r < pstep; ++r) {
            a += red[r * vpp + sl][e];
            q += red[r * vpp + sl][8 + e];
        }
        const float2 mr = mean_rstd[(size_t)b * G + (c_off + ci) / cpg];
        red_out[((size_t)b * gridDim.x + blockIdx.x) * Ctot + c_off + ci] = make_float2(a, (q - mr.x * a) * mr.y);
    }
}

// Pass 2 coefficients + parameter gradients.  One CTA per sample.
//   dx = dz * P + x * Q + R,  P = rstd*gamma',  Q = -rstd^2 * m2,  R = -rstd*m1 + mean*rstd^2*m2
//   m1 = sum_{c in g} gamma'_c S1_c / N,  m2 = sum_{c in g} gamma'_c S2_c / N,  gamma' = gamma * (1 + scale)
//   dgamma_c += sum_b S2 (1+scale)   dbeta_c += sum_b S1 (1+scale)   dscale_bc = gamma S2 + beta S1   dshift_bc = S1
// The fold over the (up to 111) chunk partials used to be one thread per channel walking its column four loads at a
// time -- a chain of ~28 dependent global-memory round trips, 20 us per call and 51 calls per training step.  Now, as in
// gn_coef_parts: the chunk range is cut into `slices` walked by different threads, two channels (one 16-byte load) per
// thread, eight loads in flight, slice partials folded in a fixed order in shared memory (deterministic).
__global__ void __launch_bounds__(1024) gn_bwd_coef_kernel(const float2* __restrict__ red_part, int nchunks,
                                                           float2* __restrict__ red, const float2* __restrict__ mean_rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ film, int C, int G, int HW,
                                                           float4* __restrict__ pqr, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float* __restrict__ dfilm, int slices) {
    extern __shared__ float s_dyn[];  // [2C] totals (S1 | S2), then [slices][2C] slice partials
    float* s_tot = s_dyn;
    float* s_part = s_dyn + 2 * C;
    __shared__ float s_m1[64], s_m2[64];
    const int b = blockIdx.x;
    const int cpg = C / G;
    const int half = C >> 1;
    const int per = (nchunks + slices - 1) / slices;
    for (int w = threadIdx.x; w < half * slices; w += blockDim.x) {
        const int pair = w % half, sl = w / half;
        const int c = 2 * pair;
        const float4* rp = reinterpret_cast<const float4*>(red_part + (size_t)b * nchunks * C + c);
        const size_t stride = (size_t)C >> 1;  // float4 units per chunk row
        const int k0 = sl * per, k1 = min(nchunks, k0 + per);
        float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
        int k = k0;
        for (; k + 8 <= k1; k += 8) {
            float4 t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __ldg(rp + (size_t)(k + i) * stride);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a0 += t[i].x; q0 += t[i].y; a1 += t[i].z; q1 += t[i].w;
            }
        }
        for (; k < k1; ++k) {
            const float4 t = __ldg(rp + (size_t)k * stride);
            a0 += t.x; q0 += t.y; a1 += t.z; q1 += t.w;
        }
        float* dst = s_part + (size_t)sl * 2 * C;
        dst[c] = a0; dst[c + 1] = a1; dst[C + c] = q0; dst[C + c + 1] = q1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += s_part[(size_t)sl * 2 * C + i];
        s_tot[i] = v;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        float m1 = 0.f, m2 = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float sc = film ? 1.f + film[(size_t)b * 2 * C + c] : 1.f;
            const float gp = gamma[c] * sc;
            m1 += gp * s_tot[c];
            m2 += gp * s_tot[C + c];
        }
        const float n = (float)cpg * (float)HW;
        s_m1[g] = m1 / n;
        s_m2[g] = m2 / n;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float2 mr = mean_rstd[(size_t)b * G + g];
        const float sc = film ? 1.f + film[(size_t)b * 2 * C + c] : 1.f;
        const float2 r = make_float2(s_tot[c], s_tot[C + c]);
        if (red != nullptr) red[(size_t)b * C + c] = r;
        const float P = mr.y * gamma[c] * sc;
        const float Q = -mr.y * mr.y * s_m2[g];
        const float R = -mr.y * s_m1[g] + mr.x * mr.y * mr.y * s_m2[g];
        pqr[(size_t)b * C + c] = make_float4(P, Q, R, 0.f);
        atomicAdd(dgamma + c, r.y * sc);
        atomicAdd(dbeta + c, r.x * sc);
        if (dfilm != nullptr) {
            dfilm[(size_t)b * 2 * C + c] = gamma[c] * r.y + beta[c] * r.x;
            dfilm[(size_t)b * 2 * C + C + c] = r.x;
        }
    }
}

// Pass 2: dx[b,p,c] = dz*P + x*Q + R (+ add[b,p,c]) ; 16-bit NHWC.
// kU: pixels per thread and loop iteration = independent 16-byte loads in flight per tensor.  ncu (profiles/
// r02_ncu_norm_final_summary.txt): 10-14 warps wait on the long scoreboard per issued instruction, the issue slots are 35 %
// used -- memory-latency bound, so the lever is bytes in flight, not instructions.
template <int kAct, int kDrop, bool kAdd, int XF, int GF, int kU>
__global__ void __launch_bounds__(kEwThreads, kU > 2 ? 2 : 3) gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                                     const __nv_bfloat16* __restrict__ g, int ld_g,
                                                                     int C, int HW, int pix_per_cta,
                                                                     const float2* __restrict__ coef,
                                                                     const float4* __restrict__ pqr, int Ctot, int c_off,
                                                                     const __nv_bfloat16* __restrict__ add,
                                                                     __nv_bfloat16* __restrict__ dx, float drop_p,
                                                                     unsigned long long seed,
                                                                     const uint8_t* __restrict__ mask_in) {
    const int vpp = C >> 3;
    const int slot = threadIdx.x % vpp, prow = threadIdx.x / vpp, pstep = blockDim.x / vpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    GnZCoef<gn_half_path<kAct, XF>()> cf;
    cf.load(coef + (size_t)b * Ctot + c_off + slot * 8);
    float P[8], Q[8], R[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 t = pqr[(size_t)b * Ctot + c_off + slot * 8 + e];
        P[e] = t.x;
        Q[e] = t.y;
        R[e] = t.z;
    }
    const uint32_t thresh = kDrop ? dropout_thresh16(drop_p) : 0u;
    const float keep_scale = kDrop ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned long long e8_base = (unsigned long long)b * HW * (unsigned long long)(ld_g >> 3) +
                                       (unsigned long long)((c_off >> 3) + slot);
    constexpr bool stored = kDrop == 2;  // keep bits fetched together with the pixel's vectors (see gn_bwd_reduce)
    if (kDrop != 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) P[e] *= keep_scale;  // dz only enters through dz * P: 1/(1-p) folded into P
    }
    auto run = [&](auto full_c) {
        constexpr bool kFull = decltype(full_c)::value;
        const size_t pix0 = (size_t)b * HW + (size_t)(p0 + prow);
        const size_t xoff = (pix0 * C + slot * 8) * 2;
        const char* xp = reinterpret_cast<const char*>(x) + xoff;
        const char* ap = kAdd ? reinterpret_cast<const char*>(add) + xoff : nullptr;
        char* dp = reinterpret_cast<char*>(dx) + xoff;
        const char* gp = reinterpret_cast<const char*>(g) + (pix0 * ld_g + c_off + slot * 8) * 2;
        const uint8_t* mp = stored ? mask_in + e8_base + (size_t)(p0 + prow) * (size_t)(ld_g >> 3) : nullptr;
        const size_t sx = (size_t)pstep * C * 2, sg = (size_t)pstep * ld_g * 2, sm = (size_t)pstep * (ld_g >> 3);
        auto body = [&](const uint4& xu, const uint4& gu, const uint4& au, uint32_t m, int p, char* dst) {
            if (kDrop == 1)
                m = dropout_keep8(seed, e8_base + (unsigned long long)p * (unsigned long long)(ld_g >> 3), thresh);
            float xf[8], dz[8], o[8];
            gn_dz8<kAct, kDrop != 0, XF, GF>(xu, gu, cf, m, xf, dz);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(dz[e], P[e], fmaf(xf[e], Q[e], R[e]));
            if (kAdd) {
                float af[8];
                cvt8_in_t<GF>(au, af);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += af[e];
            }
            stg_stream_b(dst, cvt8_out_t<GF>(o));
        };
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        int p = p0 + prow;
        for (; p + (kU - 1) * pstep < p1; p += kU * pstep) {
            uint4 xu[kU], gu[kU], au[kU];
            uint32_t mk[kU];
#pragma unroll
            for (int i = 0; i < kU; ++i) {  // kFull: the i-th row is an immediate offset of the same pointer
                xu[i] = ldg_stream_b(xp);
                gu[i] = ldg_stream_b(gp);
                au[i] = kAdd ? ldg_stream_b(ap) : zero4;
                mk[i] = stored ? (uint32_t)__ldg(mp) : 0xffu;
                xp = row_next<kFull>(xp, sx);
                gp = row_next<kFull>(gp, sg);
                if (kAdd) ap = row_next<kFull>(ap, sx);
                if (stored) mp = mask_next<kFull>(mp, sm);
            }
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                body(xu[i], gu[i], au[i], mk[i], p + i * pstep, dp);
                dp = row_next<kFull>(dp, sx);
            }
        }
        for (; p < p1; p += pstep) {
            body(ldg_stream_b(xp), ldg_stream_b(gp), kAdd ? ldg_stream_b(ap) : zero4, stored ? (uint32_t)__ldg(mp) : 0xffu, p, dp);
            xp = row_next<kFull>(xp, sx);
            gp = row_next<kFull>(gp, sg);
            dp = row_next<kFull>(dp, sx);
            if (kAdd) ap = row_next<kFull>(ap, sx);
            if (stored) mp = mask_next<kFull>(mp, sm);
        }
    };
    if (blockDim.x == kEwThreads && ld_g == C) run(std::true_type{});
    else run(std::false_type{});
}

// Pass 2 with the inputs streamed through shared memory by 1-D bulk copies (cp.async.bulk, the TMA unit) instead of register
// loads.  ncu on the register version (profiles/r02_ncu_norm_final_summary.txt): 10-14 warps per issued instruction wait on the
// long scoreboard, issue slots 35 % used -- the kernel is bound by the bytes it can keep in flight, and every 16 bytes in
// flight cost four registers of a thread that also holds 32 per-channel coefficients.  Here a CTA keeps kStages tiles of
// 512 vectors per tensor in flight (kStages * 8 KB * 2-3 tensors) whatever its register count; the consumer threads read
// their two vectors per tile from shared memory (the same (slot, pixel-row) ownership as above) and release the stage as
// soon as the data is in registers.  One thread issues: a tensor whose pixel rows are contiguous (row length == C) is one
// copy per tile, a channel slice of a wider tensor (the concat gradient, row stride ld_g) one copy per pixel row.
// Requires blockDim.x == 256 (C / 8 divides 256).
constexpr int kBulkStages = 4;
constexpr int kBulkTileBytes = 2 * kEwThreads * 16;  // 512 vectors = 8 KB per tensor and stage
template <bool kAdd>
constexpr size_t gn_bwd_apply_bulk_smem() { return (size_t)kBulkStages * (kAdd ? 3 : 2) * kBulkTileBytes + 2 * kBulkStages * 8 + 128; }

template <int kAct, int kDrop, bool kAdd, int XF, int GF>
__global__ void __launch_bounds__(kEwThreads, kAdd ? 2 : 3) gn_bwd_apply_bulk_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, int ld_g, int C, int HW, int pix_per_cta,
    const float2* __restrict__ coef, const float4* __restrict__ pqr, int Ctot, int c_off, const __nv_bfloat16* __restrict__ add,
    __nv_bfloat16* __restrict__ dx, float drop_p, unsigned long long seed, const uint8_t* __restrict__ mask_in) {
    constexpr int NT = kAdd ? 3 : 2, S = kBulkStages;
    extern __shared__ uint8_t bulk_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bulk_smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * NT * kBulkTileBytes);
    uint64_t* empty = full + S;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kEwThreads);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int vpp = C >> 3;
    const int slot = tid % vpp, prow = tid / vpp, pstep = kEwThreads / vpp;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    const int tile_pix = 2 * pstep;
    const int ntiles = (p1 - p0 + tile_pix - 1) / tile_pix;
    const size_t pix_base = (size_t)b * HW + (size_t)p0;
    const char* xb = reinterpret_cast<const char*>(x) + pix_base * C * 2;
    const char* ab = kAdd ? reinterpret_cast<const char*>(add) + pix_base * C * 2 : nullptr;
    const char* gb = reinterpret_cast<const char*>(g) + (pix_base * ld_g + c_off) * 2;
    const bool g_rows_contiguous = ld_g == C;
    auto issue = [&](int t) {  // one thread: the copies of tile t into stage t % S
        const int s = t % S;
        const int pix = t * tile_pix, np = min(tile_pix, p1 - p0 - pix);
        const uint32_t bytes = (uint32_t)np * (uint32_t)C * 2u;
        uint8_t* st = smem + (size_t)s * NT * kBulkTileBytes;
        mbar_arrive_expect_tx(&full[s], NT * bytes);
        bulk_load_1d(st, xb + (size_t)pix * C * 2, bytes, &full[s]);
        if (g_rows_contiguous) {
            bulk_load_1d(st + kBulkTileBytes, gb + (size_t)pix * C * 2, bytes, &full[s]);
        } else {
            for (int r = 0; r < np; ++r)
                bulk_load_1d(st + kBulkTileBytes + (size_t)r * C * 2, gb + (size_t)(pix + r) * ld_g * 2, (uint32_t)C * 2u, &full[s]);
        }
        if (kAdd) bulk_load_1d(st + 2 * kBulkTileBytes, ab + (size_t)pix * C * 2, bytes, &full[s]);
    };
    if (tid == 0)
        for (int t = 0; t < min(S, ntiles); ++t) issue(t);
    // per-channel coefficients (while the first tiles fly)
    GnZCoef<gn_half_path<kAct, XF>()> cf;
    cf.load(coef + (size_t)b * Ctot + c_off + slot * 8);
    float P[8], Q[8], R[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float4 t4 = pqr[(size_t)b * Ctot + c_off + slot * 8 + e];
        P[e] = t4.x;
        Q[e] = t4.y;
        R[e] = t4.z;
    }
    const uint32_t thresh = kDrop ? dropout_thresh16(drop_p) : 0u;
    const float keep_scale = kDrop ? 1.f / (1.f - drop_p) : 1.f;
    if (kDrop != 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) P[e] *= keep_scale;
    }
    constexpr bool stored = kDrop == 2;
    const unsigned long long e8_base = (unsigned long long)b * HW * (unsigned long long)(ld_g >> 3) +
                                       (unsigned long long)((c_off >> 3) + slot);
    const size_t mrow = (size_t)(ld_g >> 3);
    char* dp = reinterpret_cast<char*>(dx) + ((pix_base + prow) * C + slot * 8) * 2;
    const size_t sx = (size_t)pstep * C * 2;
    // keep bytes of the stored dropout mask: fetched one tile ahead (two bytes per thread and tile)
    auto mask_of = [&](int p) -> uint32_t {
        return (stored && p < p1) ? (uint32_t)__ldg(mask_in + e8_base + (size_t)p * mrow) : 0xffu;
    };
    uint32_t mk0 = mask_of(p0 + prow), mk1 = mask_of(p0 + prow + pstep);
    auto body = [&](const uint4& xu, const uint4& gu, const uint4& au, uint32_t m, int p, char* dst) {
        if (kDrop == 1) m = dropout_keep8(seed, e8_base + (unsigned long long)p * (unsigned long long)(ld_g >> 3), thresh);
        float xf[8], dz[8], o[8];
        gn_dz8<kAct, kDrop != 0, XF, GF>(xu, gu, cf, m, xf, dz);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(dz[e], P[e], fmaf(xf[e], Q[e], R[e]));
        if (kAdd) {
            float af[8];
            cvt8_in_t<GF>(au, af);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += af[e];
        }
        stg_stream_b(dst, cvt8_out_t<GF>(o));
    };
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        const uint32_t ph = (uint32_t)(t / S) & 1u;
        const int p = p0 + t * tile_pix + prow;  // this thread's first pixel of the tile; the second is p + pstep
        const uint32_t nm0 = mask_of(p + tile_pix), nm1 = mask_of(p + tile_pix + pstep);  // next tile's keep bytes
        const uint8_t* st = smem + (size_t)s * NT * kBulkTileBytes + (size_t)tid * 16;
        mbar_wait(&full[s], ph);
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        const uint4 xu0 = *reinterpret_cast<const uint4*>(st), xu1 = *reinterpret_cast<const uint4*>(st + kEwThreads * 16);
        const uint4 gu0 = *reinterpret_cast<const uint4*>(st + kBulkTileBytes);
        const uint4 gu1 = *reinterpret_cast<const uint4*>(st + kBulkTileBytes + kEwThreads * 16);
        const uint4 au0 = kAdd ? *reinterpret_cast<const uint4*>(st + 2 * kBulkTileBytes) : zero4;
        const uint4 au1 = kAdd ? *reinterpret_cast<const uint4*>(st + 2 * kBulkTileBytes + kEwThreads * 16) : zero4;
        mbar_arrive(&empty[s]);  // the stage is free as soon as everybody holds its vectors in registers
        if (tid == 0 && t + S < ntiles) {
            mbar_wait(&empty[s], ph);
            issue(t + S);
        }
        if (p < p1) body(xu0, gu0, au0, mk0, p, dp);
        if (p + pstep < p1) body(xu1, gu1, au1, mk1, p + pstep, dp + sx);
        dp += 2 * sx;
        mk0 = nm0;
        mk1 = nm1;
    }
}

// ------------------------------------------------------------------------------------------------ resampling
// nearest x2 upsample, NHWC bf16: out[b, y, x, :] = in[b, y/2, x/2, :]
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int vpp) {
    const long long total = (long long)B * (2 * H) * (2 * W) * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int x = (int)(r % (2 * W));  r /= (2 * W);
        const int y = (int)(r % (2 * H));
        const int b = (int)(r / (2 * H));
        stg_stream(out + i, ldg_stream(in + (((size_t)b * H + (y >> 1)) * W + (x >> 1)) * vpp + v));
    }
}
// backward of nearest x2: out[b, y, x, :] = sum of the 2x2 block of in (fp32 accumulate)
__global__ void sumpool2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int vpp,
                                 int fmt) {
    const long long total = (long long)B * H * W * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int x = (int)(r % W);  r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float f[8];
                cvt8_in(ldg_stream(in + (((size_t)b * 2 * H + 2 * y + dy) * (2 * W) + 2 * x + dx) * vpp + v), fmt, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += f[e];
            }
        stg_stream(out + i, cvt8_out(acc, fmt));
    }
}
// zero-insertion (transposed stride-2 conv as a stride-1 conv): out[b, 2y, 2x, :] = in[b, y, x, :], 0 elsewhere
__global__ void zero_insert2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int vpp) {
    const long long total = (long long)B * (2 * H) * (2 * W) * vpp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vpp);
        long long r = i / vpp;
        const int x = (int)(r % (2 * W));  r /= (2 * W);
        const int y = (int)(r % (2 * H));
        const int b = (int)(r / (2 * H));
        uint4 val = make_uint4(0, 0, 0, 0);
        if (((x | y) & 1) == 0) val = ldg_stream(in + (((size_t)b * H + (y >> 1)) * W + (x >> 1)) * vpp + v);
        stg_stream(out + i, val);
    }
}

// ------------------------------------------------------------------------------------------------ misc reductions
// per-channel sum over all pixels of an NHWC bf16 tensor (bias gradients): out[c] += sum_{b,p} x[b,p,c]
__global__ void __launch_bounds__(kEwThreads) channel_sum_kernel(const __nv_bfloat16* __restrict__ x, int C,
                                                                 long long npix, int pix_per_cta,
                                                                 float* __restrict__ out, int fmt) {
    __shared__ float red[kEwThreads][9];
    const int vpp = C >> 3;
    const int slot = threadIdx.x % vpp, prow = threadIdx.x / vpp, pstep = blockDim.x / vpp;
    const long long p0 = (long long)blockIdx.x * pix_per_cta;
    const long long p1 = min(npix, p0 + pix_per_cta);
    const uint4* src = reinterpret_cast<const uint4*>(x);
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (long long p = p0 + prow; p < p1; p += pstep) {
        float f[8];
        cvt8_in(ldg_stream(src + p * vpp + slot), fmt, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = s[e];
    __syncthreads();
    for (int ci = threadIdx.x; ci < C; ci += blockDim.x) {
        const int sl = ci >> 3, e = ci & 7;
        float a = 0.f;
        for (int r = 0; r < pstep; ++r) a += red[r * vpp + sl][e];
        atomicAdd(out + ci, a);
    }
}

// Flow-matching loss: loss += sum (v - (x1 - x0))^2 * inv_n ; dv = 2 (v - (x1 - x0)) * inv_n * gscale.  fp32 NCHW.
__global__ void __launch_bounds__(kEwThreads) fm_loss_kernel(const float* __restrict__ v, const float* __restrict__ x0,
                                                             const float* __restrict__ x1, long long n, float inv_n,
                                                             float* __restrict__ loss, float* __restrict__ dv) {
    float acc = 0.f;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(v)[i];
        const float4 s = reinterpret_cast<const float4*>(x0)[i];
        const float4 t = reinterpret_cast<const float4*>(x1)[i];
        float4 d = make_float4(a.x - (t.x - s.x), a.y - (t.y - s.y), a.z - (t.z - s.z), a.w - (t.w - s.w));
        acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (dv) reinterpret_cast<float4*>(dv)[i] = make_float4(2.f * inv_n * d.x, 2.f * inv_n * d.y, 2.f * inv_n * d.z, 2.f * inv_n * d.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = n4 << 2; i < n; ++i) {
            const float d = v[i] - (x1[i] - x0[i]);
            acc += d * d;
            if (dv) dv[i] = 2.f * inv_n * d;
        }
    }
    __shared__ float wsum[kEwThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < kEwThreads / 32 ? wsum[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(loss, t * inv_n);
    }
}

// 16-bit format conversion (fp16 <-> bf16), n8 vectors of 8 elements
__global__ void convert16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n8, int in_fmt,
                                 int out_fmt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float f[8];
        cvt8_in(ldg_stream(in + i), in_fmt, f);
        stg_stream(out + i, cvt8_out(f, out_fmt));
    }
}

// layout changes between the reference's fp32 NCHW tensors and the engine's bf16 NHWC tensors
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int B,
                                             int C, int HW, int fmt) {
    const long long total = (long long)B * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int p = (int)(r % HW);
        const int b = (int)(r / HW);
        out[i] = pack1(in[((size_t)b * C + c) * HW + p], fmt);
    }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int B,
                                             int C, int HW, int fmt) {
    const long long total = (long long)B * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const long long r = i / HW;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        out[i] = unpack1(in[((size_t)b * HW + p) * C + c], fmt);
    }
}

}  // namespace s2s
