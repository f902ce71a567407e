// Multi-head softmax attention core of torchcfm's AttentionBlock (QKVAttentionLegacy / QKVAttention):
//   w = softmax_fp32((q * s)(k * s)^T),  s = ch^(-1/4);   a = w v          per (sample, head), T = H*W tokens, ch = 32 or 64
// forward and backward, reading q / k / v straight out of the qkv conv's NHWC output [B, T, 3C] (legacy interleave
// channel = head*3*ch + {q,k,v}*ch + c, or the "new order" {q,k,v}*C + head*ch + c) and writing a [B, T, C] / d_qkv [B, T, 3C]
// in place of the reference's reshape / split / einsum / softmax / einsum chain: the T x T matrix never leaves registers.
//
// This is 0.5 % of the model's FLOPs (3.1 % in the mask-conditioned variant).  It is written on the register-level tensor
// instruction (mma.sync m16n8k16, fp32 accumulate) rather than tcgen05: the flash-style recurrence keeps S, P and dS in the
// accumulator registers of the warp that owns the 16 query (or key) rows and feeds them back as the A operand of the next
// MMA without touching shared memory; tcgen05 takes both operands from shared memory, so P and dS (in two transposes for
// the backward) would have to be staged there for every 128 x N tile.  Softmax statistics are fp32, exp2-domain.
//
//   forward : grid (ceil(T/64), B*heads), 4 warps x 16 queries; K / V blocks of 64 keys staged in shared memory
//   backward: attn_bwd_prep (D = rowsum(dO * O)), attn_bwd_kv (CTA = 64 keys: dK, dV over all query blocks, S^T form),
//             attn_bwd_q (CTA = 64 queries: dQ over all key blocks).  No atomics: deterministic.
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int kAttnBlock = 64;     // queries per CTA / keys per staged block
constexpr int kAttnThreads = 128;  // 4 warps x 16 rows

struct AttnParams {
    const uint16_t* qkv;  // [B][T][ld]   forward activations (fmt a_fmt)
    uint16_t* out;        // [B][T][C]    fwd: a (a_fmt)
    float* lse;           // [B*heads][T] log2-domain logsumexp of the scaled scores
    const uint16_t* d_out;  // bwd: dL/da [B][T][C] (g_fmt)
    const uint16_t* o_fwd;  // bwd: a from the forward pass (a_fmt)
    float* dvec;            // bwd: [B*heads][T] rowsum(dO * O)
    uint16_t* d_qkv;        // bwd: [B][T][ld] (g_fmt)
    int B, T, heads, ld, C;
    int head_stride, which_stride;  // channel of (head h, which w in {q,k,v}, c) = h*head_stride + w*which_stride + c
    float scale_log2;               // log2(e) / sqrt(ch)
    float scale;                    // 1 / sqrt(ch)
};

// ---------------------------------------------------------------------------------------------- fragment helpers
template <int F>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (F == kFmtF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// A tile [rows][D] (row pitch D + 8 elements: conflict-free ldmatrix) in shared memory
template <int D>
struct AttnTile {
    static constexpr int kPitch = D + 8;
    uint16_t v[kAttnBlock * kPitch];
    __device__ __forceinline__ uint32_t addr(int row, int col) const { return smem_u32(v + row * kPitch + col); }
};

// stage rows [r0, r0 + 64) x D channels starting at `src` (row stride ld) into `dst` as format OF (zero rows beyond T)
template <int D, int IF, int OF>
__device__ __forceinline__ void attn_stage(AttnTile<D>& dst, const uint16_t* src, int ld, int r0, int T) {
    constexpr int kVec = D / 8;
    for (int i = threadIdx.x; i < kAttnBlock * kVec; i += kAttnThreads) {
        const int r = i / kVec, c8 = i % kVec;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (r0 + r < T) u = ldg_nc16(src + (size_t)(r0 + r) * ld + c8 * 8);
        if (IF != OF) {
            float f[8];
            float2 t;
            t = unpack2(u.x, IF); f[0] = t.x; f[1] = t.y;
            t = unpack2(u.y, IF); f[2] = t.x; f[3] = t.y;
            t = unpack2(u.z, IF); f[4] = t.x; f[5] = t.y;
            t = unpack2(u.w, IF); f[6] = t.x; f[7] = t.y;
            u = make_uint4(pack2(f[0], f[1], OF), pack2(f[2], f[3], OF), pack2(f[4], f[5], OF), pack2(f[6], f[7], OF));
        }
        *reinterpret_cast<uint4*>(dst.v + r * AttnTile<D>::kPitch + c8 * 8) = u;
    }
}

// A fragments (16 rows x D) of the warp's rows [wrow, wrow+16) from a staged tile
template <int D>
__device__ __forceinline__ void attn_load_a(uint32_t (&a)[D / 16][4], const AttnTile<D>& t, int wrow) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) ldsm_x4(a[ks], t.addr(wrow + (lane & 15), ks * 16 + (lane >> 4) * 8));
}
// acc[nt] (16 x 8 each, nt over 64 columns) = A (16 x D) * Bt^T, Bt staged as [n = 64][k = D]
template <int D, int F>
__device__ __forceinline__ void attn_mma_nk(float (&acc)[8][4], const uint32_t (&a)[D / 16][4], const AttnTile<D>& bt) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t b0, b1;
            ldsm_x2(b0, b1, bt.addr(nt * 8 + (lane & 7), ks * 16 + ((lane >> 3) & 1) * 8));
            mma16816<F>(acc[nt], a[ks], b0, b1);
        }
    }
}
// acc[nt] (16 x 8 each, nt over D columns) += P (16 x 64, accumulator layout packed to 16-bit) * Bm, Bm staged [k = 64][n = D]
template <int D, int F>
__device__ __forceinline__ void attn_mma_kn(float (&acc)[D / 8][4], const uint32_t (&p)[4][4], const AttnTile<D>& bm) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            uint32_t b0, b1;
            ldsm_x2_t(b0, b1, bm.addr(ks * 16 + (lane & 15), nt * 8));
            mma16816<F>(acc[nt], p[ks], b0, b1);
        }
    }
}
// accumulator tiles (16 x 64 fp32) -> A fragments (4 k-steps of 16) in format F
template <int F>
__device__ __forceinline__ void attn_pack_a(uint32_t (&p)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        p[j][0] = pack2(s[2 * j][0], s[2 * j][1], F);
        p[j][1] = pack2(s[2 * j][2], s[2 * j][3], F);
        p[j][2] = pack2(s[2 * j + 1][0], s[2 * j + 1][1], F);
        p[j][3] = pack2(s[2 * j + 1][2], s[2 * j + 1][3], F);
    }
}

// ---------------------------------------------------------------------------------------------- forward
template <int D, int AF>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(16) AttnTile<D> sQ, sK, sV;
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    attn_stage<D, AF, AF>(sQ, base, p.ld, q0, p.T);
    __syncthreads();
    uint32_t qa[D / 16][4];
    attn_load_a<D>(qa, sQ, warp * 16);
    float o[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    for (int k0 = 0; k0 < p.T; k0 += kAttnBlock) {
        __syncthreads();  // previous block's reads of sK / sV are done
        attn_stage<D, AF, AF>(sK, base + p.which_stride, p.ld, k0, p.T);
        attn_stage<D, AF, AF>(sV, base + 2 * p.which_stride, p.ld, k0, p.T);
        __syncthreads();
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        attn_mma_nk<D, AF>(s, qa, sK);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = k0 + nt * 8 + 2 * t4 + (e & 1);
                s[nt][e] = key < p.T ? s[nt][e] * p.scale_log2 : -INFINITY;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
            }
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(mrow[r], mx[r]);
            alpha[r] = ex2f(mrow[r] - mnew);  // first block: exp2(-inf) = 0
            mrow[r] = mnew;
            lrow[r] *= alpha[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[nt][e] = ex2f(s[nt][e] - mrow[e >> 1]);
                lrow[e >> 1] += s[nt][e];
            }
        }
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
            o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
        }
        uint32_t pa[4][4];
        attn_pack_a<AF>(pa, s);
        attn_mma_kn<D, AF>(o, pa, sV);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        if (q >= p.T) continue;
        const float inv = 1.f / lrow[r];
        uint16_t* dst = p.out + ((size_t)b * p.T + q) * p.C + h * D;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt)
            *reinterpret_cast<uint32_t*>(dst + nt * 8 + 2 * t4) = pack2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv, AF);
        if (t4 == 0 && p.lse != nullptr) p.lse[(size_t)bh * p.T + q] = mrow[r] + log2f(lrow[r]);
    }
}

// ---------------------------------------------------------------------------------------------- backward
// D[bh][t] = sum_c dO[b,t,h*ch+c] * O[b,t,h*ch+c]
template <int D, int AF, int GF>
__global__ void attn_bwd_prep_kernel(const AttnParams p) {
    const long long n = (long long)p.B * p.heads * p.T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % p.T);
        const int bh = (int)(i / p.T), b = bh / p.heads, h = bh % p.heads;
        const size_t off = ((size_t)b * p.T + t) * p.C + h * D;
        float acc = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < D / 8; ++c8) {
            const uint4 a = ldg_nc16(p.d_out + off + c8 * 8), o = ldg_nc16(p.o_fwd + off + c8 * 8);
            const uint32_t av[4] = {a.x, a.y, a.z, a.w}, ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 x = unpack2(av[e], GF), y = unpack2(ov[e], AF);
                acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
            }
        }
        p.dvec[i] = acc;
    }
}

// CTA = 64 keys of one (sample, head): dK, dV accumulated over all query blocks in the transposed (keys x queries) form
template <int D, int AF, int GF>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_kv_kernel(const AttnParams p) {
    __shared__ __align__(16) AttnTile<D> sKa, sVg, sQa, sQg, sdO;  // a = forward format (S), g = gradient format
    __shared__ float sL[kAttnBlock], sD[kAttnBlock];
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int k0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    const uint16_t* dob = p.d_out + (size_t)b * p.T * p.C + h * D;
    attn_stage<D, AF, AF>(sKa, base + p.which_stride, p.ld, k0, p.T);
    attn_stage<D, AF, GF>(sVg, base + 2 * p.which_stride, p.ld, k0, p.T);
    __syncthreads();
    uint32_t ka[D / 16][4], va[D / 16][4];
    attn_load_a<D>(ka, sKa, warp * 16);
    attn_load_a<D>(va, sVg, warp * 16);
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[nt][e] = dv[nt][e] = 0.f;
    for (int q0 = 0; q0 < p.T; q0 += kAttnBlock) {
        __syncthreads();
        attn_stage<D, AF, AF>(sQa, base, p.ld, q0, p.T);
        attn_stage<D, AF, GF>(sQg, base, p.ld, q0, p.T);
        attn_stage<D, GF, GF>(sdO, dob, p.C, q0, p.T);
        if (threadIdx.x < kAttnBlock) {
            const int q = q0 + threadIdx.x;
            sL[threadIdx.x] = q < p.T ? p.lse[(size_t)bh * p.T + q] : INFINITY;  // exp2(s - inf) = 0: padded queries vanish
            sD[threadIdx.x] = q < p.T ? p.dvec[(size_t)bh * p.T + q] : 0.f;
        }
        __syncthreads();
        float st[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) st[nt][e] = dp[nt][e] = 0.f;
        attn_mma_nk<D, AF>(st, ka, sQa);   // S^T  = K Q^T   (keys x queries)
        attn_mma_nk<D, GF>(dp, va, sdO);   // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qc = nt * 8 + 2 * t4 + (e & 1);            // query = column
                const int key = k0 + warp * 16 + g + (e >> 1) * 8;   // key = row
                const float pt = key < p.T ? ex2f(st[nt][e] * p.scale_log2 - sL[qc]) : 0.f;
                st[nt][e] = pt;
                dp[nt][e] = pt * (dp[nt][e] - sD[qc]) * p.scale;     // dS^T (softmax scale folded in)
            }
        }
        uint32_t pa[4][4];
        attn_pack_a<GF>(pa, st);
        attn_mma_kn<D, GF>(dv, pa, sdO);   // dV += P^T dO
        attn_pack_a<GF>(pa, dp);
        attn_mma_kn<D, GF>(dk, pa, sQg);   // dK += dS^T Q
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = k0 + warp * 16 + g + r * 8;
        if (key >= p.T) continue;
        uint16_t* dst = p.d_qkv + ((size_t)b * p.T + key) * p.ld + h * p.head_stride;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            *reinterpret_cast<uint32_t*>(dst + p.which_stride + nt * 8 + 2 * t4) = pack2(dk[nt][2 * r], dk[nt][2 * r + 1], GF);
            *reinterpret_cast<uint32_t*>(dst + 2 * p.which_stride + nt * 8 + 2 * t4) = pack2(dv[nt][2 * r], dv[nt][2 * r + 1], GF);
        }
    }
}

// CTA = 64 queries of one (sample, head): dQ accumulated over all key blocks
template <int D, int AF, int GF>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_q_kernel(const AttnParams p) {
    __shared__ __align__(16) AttnTile<D> sQa, sdO, sKa, sKg, sVg;
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    const uint16_t* dob = p.d_out + (size_t)b * p.T * p.C + h * D;
    attn_stage<D, AF, AF>(sQa, base, p.ld, q0, p.T);
    attn_stage<D, GF, GF>(sdO, dob, p.C, q0, p.T);
    __syncthreads();
    uint32_t qa[D / 16][4], da[D / 16][4];
    attn_load_a<D>(qa, sQa, warp * 16);
    attn_load_a<D>(da, sdO, warp * 16);
    float lrow[2], drow[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        lrow[r] = q < p.T ? p.lse[(size_t)bh * p.T + q] : INFINITY;
        drow[r] = q < p.T ? p.dvec[(size_t)bh * p.T + q] : 0.f;
    }
    float dq[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
    for (int k0 = 0; k0 < p.T; k0 += kAttnBlock) {
        __syncthreads();
        attn_stage<D, AF, AF>(sKa, base + p.which_stride, p.ld, k0, p.T);
        attn_stage<D, AF, GF>(sKg, base + p.which_stride, p.ld, k0, p.T);
        attn_stage<D, AF, GF>(sVg, base + 2 * p.which_stride, p.ld, k0, p.T);
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
        attn_mma_nk<D, AF>(s, qa, sKa);    // S  = Q K^T
        attn_mma_nk<D, GF>(dp, da, sVg);   // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = k0 + nt * 8 + 2 * t4 + (e & 1);
                const float pv = key < p.T ? ex2f(s[nt][e] * p.scale_log2 - lrow[e >> 1]) : 0.f;
                dp[nt][e] = pv * (dp[nt][e] - drow[e >> 1]) * p.scale;
            }
        }
        uint32_t pa[4][4];
        attn_pack_a<GF>(pa, dp);
        attn_mma_kn<D, GF>(dq, pa, sKg);   // dQ += dS K
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        if (q >= p.T) continue;
        uint16_t* dst = p.d_qkv + ((size_t)b * p.T + q) * p.ld + h * p.head_stride;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt)
            *reinterpret_cast<uint32_t*>(dst + nt * 8 + 2 * t4) = pack2(dq[nt][2 * r], dq[nt][2 * r + 1], GF);
    }
}

}  // namespace s2s
