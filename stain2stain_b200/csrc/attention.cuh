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

// Staging of a [64 rows][D] block goes through registers so that the NEXT block's global loads are in flight while the
// current block is being multiplied (and so that a block can be written in two formats from one load).
template <int D>
struct AttnRegs {
    static constexpr int kN = kAttnBlock * (D / 8) / kAttnThreads;  // uint4 per thread: 2 (D = 32) or 4 (D = 64)
    uint4 u[kN];
};
template <int D>
__device__ __forceinline__ void attn_fetch(AttnRegs<D>& r, const uint16_t* src, int ld, int r0, int T) {
    constexpr int kVec = D / 8;
#pragma unroll
    for (int i = 0; i < AttnRegs<D>::kN; ++i) {
        const int idx = threadIdx.x + i * kAttnThreads;
        const int row = idx / kVec, c8 = idx % kVec;
        r.u[i] = make_uint4(0, 0, 0, 0);  // rows beyond T: zeros
        if (r0 + row < T) r.u[i] = ldg_nc16(src + (size_t)(r0 + row) * ld + c8 * 8);
    }
}
template <int D, int IF, int OF>
__device__ __forceinline__ void attn_put(AttnTile<D>& dst, const AttnRegs<D>& r) {
    constexpr int kVec = D / 8;
#pragma unroll
    for (int i = 0; i < AttnRegs<D>::kN; ++i) {
        const int idx = threadIdx.x + i * kAttnThreads;
        const int row = idx / kVec, c8 = idx % kVec;
        uint4 u = r.u[i];
        if (IF != OF) {
            float2 t0 = unpack2(u.x, IF), t1 = unpack2(u.y, IF), t2 = unpack2(u.z, IF), t3 = unpack2(u.w, IF);
            u = make_uint4(pack2(t0.x, t0.y, OF), pack2(t1.x, t1.y, OF), pack2(t2.x, t2.y, OF), pack2(t3.x, t3.y, OF));
        }
        *reinterpret_cast<uint4*>(dst.v + row * AttnTile<D>::kPitch + c8 * 8) = u;
    }
}

// A fragments (16 rows x D) of the warp's rows [wrow, wrow+16) from a staged tile
template <int D>
__device__ __forceinline__ void attn_load_a(uint32_t (&a)[D / 16][4], const AttnTile<D>& t, int wrow) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) ldsm_x4(a[ks], t.addr(wrow + (lane & 15), ks * 16 + (lane >> 4) * 8));
}
// acc[nt] (16 x 8 each, nt over 64 columns) = A (16 x D) * Bt^T, Bt staged as [n = 64][k = D].  One ldmatrix.x4 fetches the
// B fragments of one n-tile for 32 k (two k-steps).
template <int D, int F>
__device__ __forceinline__ void attn_mma_nk(float (&acc)[8][4], const uint32_t (&a)[D / 16][4], const AttnTile<D>& bt) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int k2 = 0; k2 < D / 32; ++k2) {
            uint32_t b[4];
            ldsm_x4(b, bt.addr(nt * 8 + (lane & 7), k2 * 32 + (lane >> 3) * 8));
            mma16816<F>(acc[nt], a[2 * k2], b[0], b[1]);
            mma16816<F>(acc[nt], a[2 * k2 + 1], b[2], b[3]);
        }
    }
}
// acc[nt] (16 x 8 each, nt over D columns) += P (16 x 64, accumulator layout packed to 16-bit) * Bm, Bm staged [k = 64][n = D].
// One ldmatrix.x4.trans fetches the B fragments of two n-tiles for one k-step.
template <int D, int F>
__device__ __forceinline__ void attn_mma_kn(float (&acc)[D / 8][4], const uint32_t (&p)[4][4], const AttnTile<D>& bm) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int n2 = 0; n2 < D / 16; ++n2) {
            uint32_t b[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                         : "r"(bm.addr(ks * 16 + (lane & 15), n2 * 16 + (lane >> 4) * 8)));
            mma16816<F>(acc[2 * n2], p[ks], b[0], b[1]);
            mma16816<F>(acc[2 * n2 + 1], p[ks], b[2], b[3]);
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
template <int N>
__device__ __forceinline__ void attn_zero(float (&a)[N][4]) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
}
// double-buffered staging when it fits the 48 KB static shared-memory window (ch = 32), single-buffered otherwise
template <int D>
struct AttnBuf { static constexpr int kN = D <= 32 ? 2 : 1; };

// ---------------------------------------------------------------------------------------------- forward
template <int D, int AF>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p) {
    constexpr int NB = AttnBuf<D>::kN;
    __shared__ __align__(16) AttnTile<D> sQ, sK[NB], sV[NB];
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    AttnRegs<D> rk, rv;
    attn_fetch<D>(rk, base, p.ld, q0, p.T);
    attn_put<D, AF, AF>(sQ, rk);
    attn_fetch<D>(rk, base + p.which_stride, p.ld, 0, p.T);
    attn_fetch<D>(rv, base + 2 * p.which_stride, p.ld, 0, p.T);
    attn_put<D, AF, AF>(sK[0], rk);
    attn_put<D, AF, AF>(sV[0], rv);
    __syncthreads();
    uint32_t qa[D / 16][4];
    attn_load_a<D>(qa, sQ, warp * 16);
    float o[D / 8][4];
    attn_zero(o);
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};  // mrow in the scaled (log2) domain
    const int nblk = (p.T + kAttnBlock - 1) / kAttnBlock;
    for (int j = 0; j < nblk; ++j) {
        const int k0 = j * kAttnBlock, cur = NB == 2 ? (j & 1) : 0;
        const bool more = j + 1 < nblk;
        if (more) {  // the next block's loads fly while this block is multiplied
            attn_fetch<D>(rk, base + p.which_stride, p.ld, k0 + kAttnBlock, p.T);
            attn_fetch<D>(rv, base + 2 * p.which_stride, p.ld, k0 + kAttnBlock, p.T);
        }
        float s[8][4];
        attn_zero(s);
        attn_mma_nk<D, AF>(s, qa, sK[cur]);
        if (k0 + kAttnBlock > p.T) {  // ragged last block: keys beyond T never win the max and weigh 0
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (k0 + nt * 8 + 2 * t4 + (e & 1) >= p.T) s[nt][e] = -INFINITY;
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        float alpha[2], nm[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(mrow[r], mx[r] * p.scale_log2);  // scale > 0: max commutes with it
            alpha[r] = ex2f(mrow[r] - mnew);  // first block: exp2(-inf) = 0
            mrow[r] = mnew;
            nm[r] = -mnew;
            lrow[r] *= alpha[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[nt][e] = ex2f(fmaf(s[nt][e], p.scale_log2, nm[e >> 1]));  // one FFMA + one MUFU per score
                lrow[e >> 1] += s[nt][e];
            }
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
            o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
        }
        uint32_t pa[4][4];
        attn_pack_a<AF>(pa, s);
        attn_mma_kn<D, AF>(o, pa, sV[cur]);
        if (NB == 1) __syncthreads();  // single buffer: everybody is done reading before it is overwritten
        if (more) {
            attn_put<D, AF, AF>(sK[NB == 2 ? cur ^ 1 : 0], rk);
            attn_put<D, AF, AF>(sV[NB == 2 ? cur ^ 1 : 0], rv);
        }
        __syncthreads();
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
    constexpr int NB = AttnBuf<D>::kN;
    __shared__ __align__(16) AttnTile<D> sKa, sVg, sQa[NB], sQg[NB], sdO[NB];  // a = forward format (S), g = gradient format
    __shared__ float sL[NB][kAttnBlock], sD[NB][kAttnBlock];
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int k0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    const uint16_t* dob = p.d_out + (size_t)b * p.T * p.C + h * D;
    const float* lse = p.lse + (size_t)bh * p.T;
    const float* dvec = p.dvec + (size_t)bh * p.T;
    AttnRegs<D> rq, rd;
    float rl = 0.f, rdv = 0.f;
    auto fetch = [&](int q0) {
        attn_fetch<D>(rq, base, p.ld, q0, p.T);
        attn_fetch<D>(rd, dob, p.C, q0, p.T);
        if (threadIdx.x < kAttnBlock) {
            const int q = q0 + threadIdx.x;
            rl = q < p.T ? lse[q] : INFINITY;  // exp2(s - inf) = 0: padded queries vanish
            rdv = q < p.T ? dvec[q] : 0.f;
        }
    };
    auto put = [&](int buf) {
        attn_put<D, AF, AF>(sQa[buf], rq);
        attn_put<D, AF, GF>(sQg[buf], rq);
        attn_put<D, GF, GF>(sdO[buf], rd);
        if (threadIdx.x < kAttnBlock) {
            sL[buf][threadIdx.x] = rl;
            sD[buf][threadIdx.x] = rdv;
        }
    };
    attn_fetch<D>(rq, base + p.which_stride, p.ld, k0, p.T);
    attn_fetch<D>(rd, base + 2 * p.which_stride, p.ld, k0, p.T);
    attn_put<D, AF, AF>(sKa, rq);
    attn_put<D, AF, GF>(sVg, rd);
    fetch(0);
    put(0);
    __syncthreads();
    uint32_t ka[D / 16][4], va[D / 16][4];
    attn_load_a<D>(ka, sKa, warp * 16);
    attn_load_a<D>(va, sVg, warp * 16);
    float dk[D / 8][4], dv[D / 8][4];
    attn_zero(dk);
    attn_zero(dv);
    const int nblk = (p.T + kAttnBlock - 1) / kAttnBlock;
    const bool key_ok[2] = {k0 + warp * 16 + g < p.T, k0 + warp * 16 + g + 8 < p.T};
    for (int j = 0; j < nblk; ++j) {
        const int cur = NB == 2 ? (j & 1) : 0;
        const bool more = j + 1 < nblk;
        if (more) fetch((j + 1) * kAttnBlock);
        float st[8][4], dp[8][4];
        attn_zero(st);
        attn_zero(dp);
        attn_mma_nk<D, AF>(st, ka, sQa[cur]);   // S^T  = K Q^T   (keys x queries)
        attn_mma_nk<D, GF>(dp, va, sdO[cur]);   // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 l2 = *reinterpret_cast<const float2*>(&sL[cur][nt * 8 + 2 * t4]);   // query = column
            const float2 d2 = *reinterpret_cast<const float2*>(&sD[cur][nt * 8 + 2 * t4]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pt = key_ok[e >> 1] ? ex2f(fmaf(st[nt][e], p.scale_log2, -((e & 1) ? l2.y : l2.x))) : 0.f;
                st[nt][e] = pt;
                dp[nt][e] = pt * (dp[nt][e] - ((e & 1) ? d2.y : d2.x)) * p.scale;   // dS^T (softmax scale folded in)
            }
        }
        uint32_t pa[4][4];
        attn_pack_a<GF>(pa, st);
        attn_mma_kn<D, GF>(dv, pa, sdO[cur]);   // dV += P^T dO
        attn_pack_a<GF>(pa, dp);
        attn_mma_kn<D, GF>(dk, pa, sQg[cur]);   // dK += dS^T Q
        if (NB == 1) __syncthreads();
        if (more) put(NB == 2 ? cur ^ 1 : 0);
        __syncthreads();
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
    constexpr int NB = AttnBuf<D>::kN;
    __shared__ __align__(16) AttnTile<D> sQa, sdO, sKa[NB], sKg[NB], sVg[NB];
    const int bh = blockIdx.y, b = bh / p.heads, h = bh % p.heads;
    const int q0 = blockIdx.x * kAttnBlock;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint16_t* base = p.qkv + (size_t)b * p.T * p.ld + h * p.head_stride;
    const uint16_t* dob = p.d_out + (size_t)b * p.T * p.C + h * D;
    AttnRegs<D> rk, rv;
    attn_fetch<D>(rk, base, p.ld, q0, p.T);
    attn_fetch<D>(rv, dob, p.C, q0, p.T);
    attn_put<D, AF, AF>(sQa, rk);
    attn_put<D, GF, GF>(sdO, rv);
    attn_fetch<D>(rk, base + p.which_stride, p.ld, 0, p.T);
    attn_fetch<D>(rv, base + 2 * p.which_stride, p.ld, 0, p.T);
    attn_put<D, AF, AF>(sKa[0], rk);
    attn_put<D, AF, GF>(sKg[0], rk);
    attn_put<D, AF, GF>(sVg[0], rv);
    __syncthreads();
    uint32_t qa[D / 16][4], da[D / 16][4];
    attn_load_a<D>(qa, sQa, warp * 16);
    attn_load_a<D>(da, sdO, warp * 16);
    float nl[2], drow[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int q = q0 + warp * 16 + g + r * 8;
        nl[r] = q < p.T ? -p.lse[(size_t)bh * p.T + q] : -INFINITY;
        drow[r] = q < p.T ? p.dvec[(size_t)bh * p.T + q] : 0.f;
    }
    float dq[D / 8][4];
    attn_zero(dq);
    const int nblk = (p.T + kAttnBlock - 1) / kAttnBlock;
    for (int j = 0; j < nblk; ++j) {
        const int k0 = j * kAttnBlock, cur = NB == 2 ? (j & 1) : 0;
        const bool more = j + 1 < nblk;
        if (more) {
            attn_fetch<D>(rk, base + p.which_stride, p.ld, k0 + kAttnBlock, p.T);
            attn_fetch<D>(rv, base + 2 * p.which_stride, p.ld, k0 + kAttnBlock, p.T);
        }
        float s[8][4], dp[8][4];
        attn_zero(s);
        attn_zero(dp);
        attn_mma_nk<D, AF>(s, qa, sKa[cur]);    // S  = Q K^T
        attn_mma_nk<D, GF>(dp, da, sVg[cur]);   // dP = dO V^T
        const bool ragged = k0 + kAttnBlock > p.T;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float pv = ex2f(fmaf(s[nt][e], p.scale_log2, nl[e >> 1]));
                if (ragged && k0 + nt * 8 + 2 * t4 + (e & 1) >= p.T) pv = 0.f;
                dp[nt][e] = pv * (dp[nt][e] - drow[e >> 1]) * p.scale;
            }
        }
        uint32_t pa[4][4];
        attn_pack_a<GF>(pa, dp);
        attn_mma_kn<D, GF>(dq, pa, sKg[cur]);   // dQ += dS K
        if (NB == 1) __syncthreads();
        if (more) {
            attn_put<D, AF, AF>(sKa[NB == 2 ? cur ^ 1 : 0], rk);
            attn_put<D, AF, GF>(sKg[NB == 2 ? cur ^ 1 : 0], rk);
            attn_put<D, AF, GF>(sVg[NB == 2 ? cur ^ 1 : 0], rv);
        }
        __syncthreads();
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
