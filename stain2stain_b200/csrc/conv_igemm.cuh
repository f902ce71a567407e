// Implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA, NHWC bf16 activations.
//
//   out[b, y, x, n] = bias[n] + residual[b, y, x, n]
//                   + sum_{segment s} sum_{tap} sum_{c} A_s[b, y*stride + dy, x*stride + dx, c] * W[n, k(s, tap, c)]
//
// One GEMM row = one output pixel; an M tile is an 8 x 16 pixel patch (128 rows) of one sample, fetched per filter tap
// as ONE 5-D TMA box {64 ch, 16 w, 1, 8 h, 1 b} whose out-of-bounds rows/columns are zero-filled by the TMA unit
// (the conv padding is therefore free).  A box lands in shared memory as [128 pixels][64 ch] with the 128-byte
// swizzle, i.e. exactly the canonical K-major UMMA operand.  Weights are pre-packed [N][K] (K-major) and fetched as
// 2-D boxes {64 k, BN n}.  Several "segments" share one accumulator: the two halves of a channel concat, or a 3x3
// conv plus the 1x1 skip-connection conv of a ResBlock (the residual add then costs nothing).
//
// Warp roles (256 threads, persistent over tiles, 1 CTA / SM):
//   warp 0   : TMA producer (one lane)            smem ring: full[]/empty[] mbarriers
//   warp 1   : tcgen05.mma issuer (one lane)      accumulators: 2 TMEM stages of BN fp32 columns, tfull[]/tempty[]
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue  TMEM -> registers -> (+bias, +residual) -> bf16 -> swizzled smem -> TMA store (NHWC)
//              or fp32 NCHW direct stores for the narrow head conv (optionally fused with the Euler update).
//
// Stride-2 convs read the input through a parity view {2C, W/2, 2, H/2, B}: pixel (2y+dy-1, 2x+dx-1) is
// (channel offset pw*C, column x + (dx==0 ? -1 : 0), parity ph, row y + (dy==0 ? -1 : 0)).
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int kTileW = 16, kTileH = 8, kTileM = kTileW * kTileH;  // 128 output pixels per tile
constexpr int kBlockK = 64;                                      // bf16 channels per k-block = one 128 B swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;                    // 16 KB
constexpr int kOutStageBytes = kTileM * 64 * 2;                  // 16 KB staging per 64-channel output chunk
constexpr int kMaxSeg = 4;
constexpr int kConvThreads = 256;

enum ConvMode : int { kModeBf16Nhwc = 0, kModeF32Nchw = 1 };

struct ConvSeg {
    int taps;     // 9 (3x3 neighbourhood, pad 1) or 1 (1x1): the KIND of the segment (which box is fetched)
    int cblocks;  // ceil(C / 64)
    int stride;   // 1 or 2
    int C;        // channels of this segment's tensor
    // halo-tiled kernel only: the filter taps actually multiplied, in the order the packed weight stores them.  Entry i
    // (4 bits of tapmap) is the 3x3 position dy*3 + dx of logical tap i.  A plain 3x3 conv has ntaps = 9 and the identity
    // map; the phase-decomposed Upsample / transposed convs use 1-, 2- and 4-tap subsets (0 = plain).
    int ntaps;
    unsigned long long tapmap;
    unsigned long long wmap;  // tap slot of logical tap i inside the packed weight (bits [4i, 4i+4)); identity for a weight
                              // packed in logical order, a permutation when a plain 9-tap operand is reused
};
constexpr unsigned long long kTapIdentity = 0x876543210ull;

struct ConvParams {
    CUtensorMap tmA[kMaxSeg];
    CUtensorMap tmW;
    CUtensorMap tmOut;
    ConvSeg seg[kMaxSeg];
    int nseg;
    int B, Hout, Wout, Cout;
    int tiles_x, tiles_y, n_tiles_n, total_tiles;
    int BN;          // 16, 64, 128 or 256
    int mma_order;   // 0: one accumulator at a time; 1: alternate independent accumulators between consecutive MMAs
    int mt;          // M sub-tiles (8 x 16 pixel patches stacked in y) per CTA tile sharing one weight tile: 1 or 2
    int kblocks;     // total k-blocks per tile
    int num_stages;
    uint32_t tmem_cols;
    int mode;
    const float* bias;               // [Cout] or nullptr
    const __nv_bfloat16* residual;   // NHWC [B, Hout, Wout, Cout] (16-bit, res_fmt) or nullptr (NHWC mode only)
    float* out_f32;                  // NCHW [B, Cout, Hout, Wout] (fp32 mode)
    const float* axpy_x;             // fp32 mode, optional: out = axpy_x + axpy_a * (acc + bias)
    float axpy_a;
    int a_fmt, w_fmt, out_fmt, res_fmt;  // Fmt of activations (A), weights (B), NHWC output and residual
};

__device__ __forceinline__ void conv_decode_tile(const ConvParams& p, int tile, int& b, int& ty, int& tx, int& nt) {
    nt = tile % p.n_tiles_n;
    int m = tile / p.n_tiles_n;
    tx = m % p.tiles_x;
    m /= p.tiles_x;
    ty = m % p.tiles_y;
    b = m / p.tiles_y;
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024 B alignment is required by the 128 B swizzle atoms.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const int BN = p.BN;
    const uint32_t b_bytes = (uint32_t)(BN < 64 ? 64 : BN) * kBlockK * 2;  // B region per stage (>= 8 KB keeps 1 KB alignment)
    const int MT = p.mt;
    const uint32_t a_bytes = (uint32_t)MT * kABytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int S = p.num_stages;

    uint8_t* ring = smem;
    uint8_t* out_stage = smem + (size_t)S * stage_bytes;  // 2 x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + 2 * kOutStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmW);
        if (p.mode == kModeBf16Nhwc) tma_prefetch_desc(&p.tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================================================================== TMA producer (whole warp, elected issue)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int b, ty, tx, nt;
                conv_decode_tile(p, tile, b, ty, tx, nt);
                const int x0 = tx * kTileW, y0 = ty * kTileH * MT, n0 = nt * BN;
                int kb = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    for (int tap = 0; tap < sg.taps; ++tap) {
                        int cx, cy, cp = 0, coff = 0;
                        if (sg.taps == 1) {
                            cx = x0;
                            cy = y0;
                        } else if (sg.stride == 1) {
                            cx = x0 + (tap % 3) - 1;
                            cy = y0 + (tap / 3) - 1;
                        } else {
                            const int dx = tap % 3, dy = tap / 3;
                            cx = x0 + (dx == 0 ? -1 : 0);
                            cy = y0 + (dy == 0 ? -1 : 0);
                            cp = (dy == 1) ? 0 : 1;
                            coff = ((dx == 1) ? 0 : 1) * sg.C;
                        }
                        for (int cb = 0; cb < sg.cblocks; ++cb, ++kb) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            uint8_t* a_dst = ring + (size_t)stage * stage_bytes;
                            uint8_t* b_dst = a_dst + a_bytes;
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&full[stage], a_bytes + (uint32_t)BN * kBlockK * 2);
                                for (int h = 0; h < MT; ++h)
                                    tma_load_5d(a_dst + h * kABytes, &p.tmA[s], &full[stage], coff + cb * kBlockK, cx, cp,
                                                cy + h * kTileH, b);
                                tma_load_2d(b_dst, &p.tmW, &full[stage], kb * kBlockK, n0);
                            }
                            __syncwarp();
                            if (++stage == S) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp, elected issue)
        {
            const uint32_t idesc = umma_idesc_16b(kTileM, (uint32_t)BN, 0, 0, p.a_fmt, p.w_fmt);
            const uint32_t idesc_half = umma_idesc_16b(kTileM, (uint32_t)(BN / 2), 0, 0, p.a_fmt, p.w_fmt);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);  // epilogue drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BN);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = a_addr + a_bytes;
                    if (p.mma_order == 0) {
                        for (int h = 0; h < MT; ++h) {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t da = umma_smem_desc_sw128(a_addr + h * kABytes + k * 32, 16, 1024);
                                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                                if (elect_one()) umma_bf16(d_tmem + (uint32_t)(h * BN), da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                    } else if (MT == 2) {
                        // consecutive MMAs go to DIFFERENT accumulators (the two pixel sub-tiles)
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint64_t da = umma_smem_desc_sw128(a_addr + h * kABytes + k * 32, 16, 1024);
                                if (elect_one()) umma_bf16(d_tmem + (uint32_t)(h * BN), da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                    } else {
                        // one pixel tile, N split into two independent halves of BN/2 columns
                        const uint32_t hb = (uint32_t)(BN / 2);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t db = umma_smem_desc_sw128(b_addr + j * hb * 128 + k * 32, 16, 1024);
                                if (elect_one()) umma_bf16(d_tmem + j * hb, da, db, idesc_half, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    if (elect_one()) umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit(&tfull[as]);  // accumulator complete
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue
        const int q = warp & 3;           // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;    // GEMM row == pixel within the tile
        const int et = threadIdx.x - 128; // 0..127
        int it = 0;
        int ob = 0;                       // output staging buffer toggle
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            int b, ty, tx, nt;
            conv_decode_tile(p, tile, b, ty, tx, nt);
            const int x0 = tx * kTileW, y0 = ty * kTileH * MT, n0 = nt * BN;
            const int px = x0 + row % kTileW;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();

            if (p.mode == kModeF32Nchw) {  // narrow head conv (MT == 1)
                const int py = y0 + row / kTileW;
                const bool in_img = (py < p.Hout) && (px < p.Wout);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
                uint32_t v[16];
                tmem_ld_32x16(t_row, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
                if (in_img) {
                    for (int c = 0; c < p.Cout; ++c) {
                        float acc = __uint_as_float(v[c]) + (p.bias ? __ldg(p.bias + c) : 0.f);
                        const size_t o = (((size_t)b * p.Cout + c) * p.Hout + py) * p.Wout + px;
                        if (p.axpy_x) acc = __ldg(p.axpy_x + o) + p.axpy_a * acc;
                        p.out_f32[o] = acc;
                    }
                }
                continue;
            }

            const int nchunks = BN / 64;
            for (int h = 0; h < MT; ++h) {
                const int ys = y0 + h * kTileH;  // first image row of this sub-tile
                const int py = ys + row / kTileW;
                const bool in_img = (py < p.Hout) && (px < p.Wout);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MT * BN + h * BN);
                for (int ch = 0; ch < nchunks; ++ch) {
                    // the residual tile's eight 16-byte vectors are requested BEFORE the accumulator is read: loaded one by
                    // one inside the conversion loop they cost eight serialised memory latencies per 128x64 chunk, which made the
                    // epilogue 2-3x longer than the chunk's MMAs (conv2 of the identity-skip ResBlocks: 2.3 ms instead of 0.8 ms)
                    uint4 rres[8];
                    const bool has_res = p.residual != nullptr;
                    if (has_res) {
                        const int nb = n0 + ch * 64;
                        const bool ok = in_img;
                        const uint4* rp = reinterpret_cast<const uint4*>(
                            p.residual + (((size_t)b * p.Hout + (ok ? py : 0)) * p.Wout + (ok ? px : 0)) * p.Cout + nb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) rres[j] = ok ? ldg_nc16(rp + j) : make_uint4(0, 0, 0, 0);
                    }
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32(t_row + ch * 64, v0);
                    tmem_ld_32x32(t_row + ch * 64 + 32, v1);
                    tmem_ld_wait();
                    if (h == MT - 1 && ch == nchunks - 1) {
                        // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[as]);
                    }
                    const int nbase = n0 + ch * 64;
                    const bool res = has_res && in_img;
                    // staging buffer `ob` was last read by the TMA store issued two chunks ago; the converted vectors go
                    // straight into it (no register copy of the packed tile: the epilogue is register-bound)
                    if (et == 0) tma_store_wait_read<1>();
                    named_bar_sync(1, 128);
                    uint8_t* dst = out_stage + ob * kOutStageBytes + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 8 x (8 channels = 16 B)
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int col = j * 8 + e;
                            f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                            if (p.bias) f[e] += __ldg(p.bias + nbase + col);
                        }
                        if (res) {
                            const uint4 r = rres[j];
                            float2 t;
                            t = unpack2(r.x, p.res_fmt); f[0] += t.x; f[1] += t.y;
                            t = unpack2(r.y, p.res_fmt); f[2] += t.x; f[3] += t.y;
                            t = unpack2(r.z, p.res_fmt); f[4] += t.x; f[5] += t.y;
                            t = unpack2(r.w, p.res_fmt); f[6] += t.x; f[7] += t.y;
                        }
                        const int sw = j ^ (row & 7);  // 128 B swizzle: 16 B chunk index XOR (row mod 8)
                        *reinterpret_cast<uint4*>(dst + sw * 16) = make_uint4(pack2(f[0], f[1], p.out_fmt), pack2(f[2], f[3], p.out_fmt),
                                                                              pack2(f[4], f[5], p.out_fmt), pack2(f[6], f[7], p.out_fmt));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (et == 0) {
                        tma_store_5d(&p.tmOut, out_stage + ob * kOutStageBytes, nbase, x0, 0, ys, b);
                        tma_store_commit();
                    }
                    ob ^= 1;
                }
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ====================================================================================================== CTA-pair conv
// Same implicit GEMM, executed by CTA PAIRS (cluster of 2 = the two SMs of a TPC) with tcgen05.mma.cta_group::2:
// one MMA covers M = 256 output pixels (CTA r owns pixel tile 2j + r, i.e. its own A operand and its own 128
// accumulator rows in its own TMEM) x N = BN channels, and each CTA stages only HALF of the weight tile
// (rows r*BN/2 ...).  Per CTA and k-block that is 16 KB (A) + BN/2*128 B (B half) through TMA and shared memory
// instead of 16 KB + BN*128 B: the single-CTA kernel is bound by exactly that shared-memory / L2->SM feed
// (73 % tensor-pipe at BN = 256, 50 % at BN = 128; profiles/r01_ncu_kernels_summary.txt).
//
// Protocol (leader = even CTA): both producers wait on their OWN empty[s], the leader arms its full[s] with the bytes of
// BOTH CTAs and every TMA load of either CTA counts on the leader's full[s]; the leader's MMA thread issues the MMAs
// and commits with a multicast arrive on empty[s] / tfull[a] of both CTAs; the epilogue warps of both CTAs drain their
// own TMEM rows and arrive on the leader's tempty[a] (count 8).
struct Conv2Params {
    CUtensorMap tmA[kMaxSeg];
    CUtensorMap tmW;     // box {64 k, BN/2 rows}
    CUtensorMap tmOut;
    ConvSeg seg[kMaxSeg];
    int nseg;
    int B, Hout, Wout, Cout;
    int tiles_x, tiles_y, m_tiles, m_pairs, n_tiles_n, total_pairs;
    int BN, kblocks, num_stages;
    int mt;  // pixel sub-tiles (8 x 16, stacked in y) per CTA sharing one weight half-tile: 1 or 2
    uint32_t tmem_cols;
    const float* bias;
    const __nv_bfloat16* residual;
    uint16_t* out_direct;  // non-null: dense NHWC output written with plain 16-byte stores (launches without statistics)
    float2* stats;  // optional [B][stat_tiles][Cout] (sum, sumsq) of the STORED 16-bit outputs per 128-pixel sub-tile:
                    // the GroupNorm statistics of the consumer, taken from the epilogue's staging tile (no extra pass)
    int stat_tiles;
    int a_fmt, w_fmt, out_fmt, res_fmt;
};

template <int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
    conv_igemm_pair_kernel(const __grid_constant__ Conv2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // warp index through a shuffle: the compiler then knows it is warp-uniform, and the role branches below stay convergent
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int BN = p.BN, HB = p.BN / 2;
    constexpr uint32_t a_bytes = (uint32_t)MT * kABytes;
    const uint32_t b_bytes = (uint32_t)HB * kBlockK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int S = p.num_stages;
    const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for __cluster_dims__(2,1,1) on a 1-D grid (uniform)
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    uint8_t* ring = smem;
    uint8_t* out_stage = smem + (size_t)S * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + 2 * kOutStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    float2* stat_scratch = reinterpret_cast<float2*>(bars + 2 * S + 6);  // 4 warps x 64 channels x float2 (2 KB)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmW);
        tma_prefetch_desc(&p.tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);   // the leader's own arrive.expect_tx (the peer's copy is never used)
            mbar_init(&empty[i], 1);  // one multicast tcgen05.commit per use
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 8);  // 4 epilogue warps x 2 CTAs (leader's copy)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(tmem_ptr, p.tmem_cols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int pair, int& b, int& ty, int& tx, int& nt, bool& valid) {
        nt = pair % p.n_tiles_n;
        int m = (pair / p.n_tiles_n) * 2 + (int)rank;
        valid = m < p.m_tiles;
        tx = m % p.tiles_x;
        m /= p.tiles_x;
        ty = m % p.tiles_y;
        b = m / p.tiles_y;  // == B for the dummy tile of an odd tile count: TMA zero-fills loads and drops stores
    };

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        // whole warp, convergent; one elected lane issues (see the MMA issuer below for why)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters) {
                int b, ty, tx, nt;
                bool valid;
                decode(pair, b, ty, tx, nt, valid);
                const int x0 = tx * kTileW, y0 = ty * kTileH * MT, n0 = nt * BN + (int)rank * HB;
                int kb = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    for (int tap = 0; tap < sg.taps; ++tap) {
                        int cx, cy, cp = 0, coff = 0;
                        if (sg.taps == 1) {
                            cx = x0;
                            cy = y0;
                        } else if (sg.stride == 1) {
                            cx = x0 + (tap % 3) - 1;
                            cy = y0 + (tap / 3) - 1;
                        } else {
                            const int dx = tap % 3, dy = tap / 3;
                            cx = x0 + (dx == 0 ? -1 : 0);
                            cy = y0 + (dy == 0 ? -1 : 0);
                            cp = (dy == 1) ? 0 : 1;
                            coff = ((dx == 1) ? 0 : 1) * sg.C;
                        }
                        for (int cb = 0; cb < sg.cblocks; ++cb, ++kb) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            uint8_t* a_dst = ring + (size_t)stage * stage_bytes;
                            uint8_t* b_dst = a_dst + a_bytes;
                            if (elect_one()) {
                                if (leader) mbar_arrive_expect_tx(&full[stage], 2 * stage_bytes);
#pragma unroll
                                for (int h = 0; h < MT; ++h)
                                    tma_load_5d_2cta(a_dst + h * kABytes, &p.tmA[s], &full[stage], coff + cb * kBlockK, cx,
                                                     cp, cy + h * kTileH, b);
                                tma_load_2d_2cta(b_dst, &p.tmW, &full[stage], kb * kBlockK, n0);
                            }
                            __syncwarp();
                            if (++stage == S) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        // The WHOLE warp runs this loop (convergent code: stage / descriptor arithmetic lives in uniform registers) and one
        // elected lane issues each tcgen05 instruction.  Issuing from inside `if (lane == 0)` makes every operand
        // "potentially divergent": the compiler then wraps each MMA in an elect / R2UR.BROADCAST loop of ~20
        // instructions, and the M256 x N128 x K16 MMA (64 tensor cycles) becomes issue-bound.
        if (leader) {
            const uint32_t idesc = umma_idesc_16b(2 * kTileM, (uint32_t)BN, 0, 0, p.a_fmt, p.w_fmt);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);  // both CTAs' epilogues drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BN);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = a_addr + a_bytes;
                    const uint64_t da0 = umma_smem_desc_sw128(a_addr, 16, 1024);
                    const uint64_t db0 = umma_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
                        for (int h = 0; h < MT; ++h) {  // consecutive MMAs alternate between the sub-tiles' accumulators
                            // start-address field is (addr >> 4): advancing by k*32 B / h*16 KB is a plain add
                            const uint64_t da = da0 + (uint64_t)((h * kABytes + k * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((k * 32) >> 4);
                            if (elect_one())
                                umma_bf16_2cta(d_tmem + (uint32_t)(h * BN), da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (elect_one()) umma_commit_2cta(&empty[stage], 0x3);  // frees the slot in BOTH CTAs
                    __syncwarp();
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit_2cta(&tfull[as], 0x3);
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;
        int it = 0;
        int ob = 0;
        const int nchunks = BN / 64;
        for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters, ++it) {
            int b, ty, tx, nt;
            bool valid;
            decode(pair, b, ty, tx, nt, valid);
            const int x0 = tx * kTileW, y0 = ty * kTileH * MT, n0 = nt * BN;
            const int px = x0 + row % kTileW;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            for (int h = 0; h < MT; ++h) {
                const int ys = y0 + h * kTileH;
                const int py = ys + row / kTileW;
                const bool in_img = valid && (py < p.Hout) && (px < p.Wout);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MT * BN + h * BN);
                for (int ch = 0; ch < nchunks; ++ch) {
                    // the residual tile's eight 16-byte vectors are requested BEFORE the accumulator is read: loaded one by
                    // one inside the conversion loop they cost eight serialised memory latencies per 128x64 chunk, which made the
                    // epilogue 2-3x longer than the chunk's MMAs (conv2 of the identity-skip ResBlocks: 2.3 ms instead of 0.8 ms)
                    uint4 rres[8];
                    const bool has_res = p.residual != nullptr;
                    if (has_res) {
                        const int nb = n0 + ch * 64;
                        const bool ok = in_img;
                        const uint4* rp = reinterpret_cast<const uint4*>(
                            p.residual + (((size_t)b * p.Hout + (ok ? py : 0)) * p.Wout + (ok ? px : 0)) * p.Cout + nb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) rres[j] = ok ? ldg_nc16(rp + j) : make_uint4(0, 0, 0, 0);
                    }
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32(t_row + ch * 64, v0);
                    tmem_ld_32x32(t_row + ch * 64 + 32, v1);
                    tmem_ld_wait();
                    if (h == MT - 1 && ch == nchunks - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tempty[as]);
                    }
                    const int nbase = n0 + ch * 64;
                    const bool res = has_res && in_img;
                    if (p.out_direct != nullptr) {
                        // Launches that need no statistics and write a dense tensor (every dgrad conv, the 1x1 GEMMs) store
                        // their pixel row straight from registers: 8 x 16 B = the thread's own 128-byte line.  No staging
                        // tile, no named barriers, no TMA store to wait for -- the four epilogue warps run independently.
                        // (Through shared memory + TMA the epilogue took ~2800 cycles per 128 x 64 chunk, so every conv with
                        // K < 1400 was bound by it: the 1x1 skip-conv dgrads ran at 177 TFLOP/s.)
                        if (in_img) {
                            uint4* o = reinterpret_cast<uint4*>(p.out_direct + ((((size_t)b * p.Hout + py) * p.Wout + px) * p.Cout + nbase));
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float f[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const int col = j * 8 + e;
                                    f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                                    if (p.bias) f[e] += __ldg(p.bias + nbase + col);
                                }
                                if (res) {
                                    const uint4 r = rres[j];
                                    float2 t;
                                    t = unpack2(r.x, p.res_fmt); f[0] += t.x; f[1] += t.y;
                                    t = unpack2(r.y, p.res_fmt); f[2] += t.x; f[3] += t.y;
                                    t = unpack2(r.z, p.res_fmt); f[4] += t.x; f[5] += t.y;
                                    t = unpack2(r.w, p.res_fmt); f[6] += t.x; f[7] += t.y;
                                }
                                o[j] = make_uint4(pack2(f[0], f[1], p.out_fmt), pack2(f[2], f[3], p.out_fmt),
                                                  pack2(f[4], f[5], p.out_fmt), pack2(f[6], f[7], p.out_fmt));
                            }
                        }
                        continue;
                    }
                    // staging buffer `ob` was last read by the TMA store issued two chunks ago; the converted vectors go
                    // straight into it (no register copy of the packed tile: the epilogue is register-bound)
                    if (et == 0) tma_store_wait_read<1>();
                    named_bar_sync(1, 128);
                    uint8_t* dst = out_stage + ob * kOutStageBytes + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 8 x (8 channels = 16 B)
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int col = j * 8 + e;
                            f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                            if (p.bias) f[e] += __ldg(p.bias + nbase + col);
                        }
                        if (res) {
                            const uint4 r = rres[j];
                            float2 t;
                            t = unpack2(r.x, p.res_fmt); f[0] += t.x; f[1] += t.y;
                            t = unpack2(r.y, p.res_fmt); f[2] += t.x; f[3] += t.y;
                            t = unpack2(r.z, p.res_fmt); f[4] += t.x; f[5] += t.y;
                            t = unpack2(r.w, p.res_fmt); f[6] += t.x; f[7] += t.y;
                        }
                        const int sw = j ^ (row & 7);  // 128 B swizzle: 16 B chunk index XOR (row mod 8)
                        *reinterpret_cast<uint4*>(dst + sw * 16) = make_uint4(pack2(f[0], f[1], p.out_fmt), pack2(f[2], f[3], p.out_fmt),
                                                                              pack2(f[4], f[5], p.out_fmt), pack2(f[6], f[7], p.out_fmt));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (et == 0) {
                        tma_store_5d(&p.tmOut, out_stage + ob * kOutStageBytes, nbase, x0, 0, ys, b);
                        tma_store_commit();
                    }
                    if (p.stats != nullptr) {
                        // per-channel (sum, sumsq) over the 128 pixels of this sub-tile, read back from the staged tile (so
                        // the statistics are those of the rounded values the consumer reads): thread = (8-channel group,
                        // 8-row group), eight 128-bit shared loads, shuffle fold inside the warp, 2 KB scratch across warps
                        const int cg = et & 7, rg = et >> 3;
                        const uint8_t* tile = out_stage + ob * kOutStageBytes;
                        const bool full_tile = (ys + kTileH <= p.Hout) && (x0 + kTileW <= p.Wout);
                        float sm[8], sq[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) sm[e] = sq[e] = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rg * 8 + i;
                            if (!full_tile && ((ys + r / kTileW >= p.Hout) || (x0 + r % kTileW >= p.Wout))) continue;
                            const uint4 u = *reinterpret_cast<const uint4*>(tile + r * 128 + ((cg ^ i) << 4));
                            float f[8];
                            float2 t2;
                            t2 = unpack2(u.x, p.out_fmt); f[0] = t2.x; f[1] = t2.y;
                            t2 = unpack2(u.y, p.out_fmt); f[2] = t2.x; f[3] = t2.y;
                            t2 = unpack2(u.z, p.out_fmt); f[4] = t2.x; f[5] = t2.y;
                            t2 = unpack2(u.w, p.out_fmt); f[6] = t2.x; f[7] = t2.y;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                sm[e] += f[e];
                                sq[e] = fmaf(f[e], f[e], sq[e]);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 8);
                            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 8);
                            sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 16);
                            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 16);
                        }
                        if (lane < 8) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) stat_scratch[q * 64 + lane * 8 + e] = make_float2(sm[e], sq[e]);
                        }
                        named_bar_sync(3, 128);
                        if (et < 64 && valid) {
                            float2 o = stat_scratch[et];
#pragma unroll
                            for (int wq = 1; wq < 4; ++wq) {
                                const float2 t2 = stat_scratch[wq * 64 + et];
                                o.x += t2.x;
                                o.y += t2.y;
                            }
                            const int sub = (ty * MT + h) * p.tiles_x + tx;
                            p.stats[((size_t)b * p.stat_tiles + sub) * p.Cout + nbase + et] = o;
                        }
                    }
                    ob ^= 1;
                }
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA exits (or frees TMEM) while its partner may still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    }
}

// ====================================================================================================== halo-tiled CTA-pair conv
// The kernels above re-stage the A operand once per filter tap: nine overlapping 16 KB boxes per 64-channel block.  Here
// the pixel tile is 8 wide x 16 high, so one GEMM row group of 8 pixels is one image row segment, and the whole
// (16+2) x (8+2) halo of the tile is fetched by ONE TMA box {64 ch, 10 w, 1, 18 h, 1} per channel block (23 KB instead
// of 9 x 16 KB).  Its shared-memory image is [18 rows][10 px][64 ch] with the 128-byte swizzle, and the A operand of tap
// (dy, dx) is simply a UMMA descriptor on the SAME buffer: start + ((dy*10 + dx) * 128 B), stride between 8-row groups
// = one halo row = 1280 B, matrix base offset = (start >> 7) & 7 because the start is no longer 1024 B aligned.
// Weights still arrive per (tap, channel block) through their own ring.  1x1 segments (skip convs) use plain boxes.
constexpr int kHaloTW = 8, kHaloTH = 16, kHaloPitch = kHaloTW + 2;

__device__ __forceinline__ uint64_t umma_smem_desc_sw128_off(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                             int base_off_mode) {
    uint64_t d = umma_smem_desc_sw128(smem_addr, lbo_bytes, sbo_bytes);
    if (base_off_mode) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;
    return d;
}

struct Conv3Params {
    CUtensorMap tmA[kMaxSeg];  // 3x3 segments: halo box {64, 10, 1, 16*MT + 2, 1}; 1x1 segments: {64, 8, 1, 16, 1}
    CUtensorMap tmW;           // box {64 k, BN/2 rows}
    CUtensorMap tmOut;         // box {64, 8, 1, 16, 1}
    ConvSeg seg[kMaxSeg];
    int seg_kb[kMaxSeg];       // first k-block of each segment in the packed weight
    int nseg;
    int B, Hout, Wout, Cout;
    int tiles_x, tiles_y, m_tiles, n_tiles_n, total_pairs;
    int BN, sa, sb;            // A / B ring depths
    uint32_t a_slot, tmem_cols;
    int base_off_mode;
    int cols3;                 // 1: three column-shifted halo boxes {64, 8, 1, 16*MT+2, 1} per channel block (dx = -1, 0, +1):
                               // every tap's A descriptor is 1024 B aligned with an 8-row-group stride of exactly 1024 B
    const float* bias;
    const __nv_bfloat16* residual;
    uint16_t* out_direct;      // non-null: dense NHWC output written with plain 16-byte stores (launches without statistics)
    float2* stats;             // optional per-sub-tile (sum, sumsq) of the stored outputs, as in Conv2Params
    int stat_tiles, stat_off;  // sub-tiles per sample in the statistics buffer; first sub-tile this launch writes (the four
                               // phase launches of an Upsample conv fill one buffer)
    int a_fmt, w_fmt, out_fmt, res_fmt;
    // Fused normalisation prologue (inference): a segment with seg_coef != nullptr is NOT the activated tensor but the
    // raw one; warps 2-3 rewrite each landed halo box in place as act(x * A + Bc) with the per-(sample, channel)
    // coefficients of the GroupNorm (+FiLM) that precedes this conv, so the norm-apply pass over the tensor disappears.
    const float2* seg_coef[kMaxSeg];  // [B][seg_coef_ld] (A, Bc), or nullptr = segment is used as it is
    const uint16_t* seg_x[kMaxSeg];   // the segment's tensor (NHWC 16-bit): normalised segments are loaded by the transform
                                      // warps straight from global memory (no TMA, no second pass over shared memory)
    int seg_coef_ld[kMaxSeg], seg_coef_off[kMaxSeg];
    int prologue;                     // 1: A boxes go TMA -> own full barrier -> transform warps -> leader's ready barrier
    int act;                          // 0 none, 1 SiLU
};

template <int MT, bool PRO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
    conv_halo_pair_kernel(const __grid_constant__ Conv3Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const int BN = p.BN, HB = p.BN / 2;
    const uint32_t b_bytes = (uint32_t)HB * kBlockK * 2;
    constexpr uint32_t box3_bytes = (uint32_t)(kHaloTH * MT + 2) * kHaloTW * 128;  // one column-shifted box (cols3)
    const uint32_t halo_bytes = p.cols3 ? 3u * box3_bytes : (uint32_t)(kHaloTH * MT + 2) * kHaloPitch * 128;
    const int SA = p.sa, SB = p.sb;
    const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for __cluster_dims__(2,1,1) on a 1-D grid (uniform)
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    uint8_t* ringA = smem;
    uint8_t* ringB = ringA + (size_t)SA * p.a_slot;
    uint8_t* out_stage = ringB + (size_t)SB * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + 2 * kOutStageBytes);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + SA;
    uint64_t* fullB = emptyA + SA;
    uint64_t* emptyB = fullB + SB;
    uint64_t* tfull = emptyB + SB;
    uint64_t* tempty = tfull + 2;
    uint64_t* readyA = tempty + 2;  // prologue mode: 2 transform warps x 2 CTAs arrive on the leader's copy
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(readyA + SA);
    float2* stat_scratch = reinterpret_cast<float2*>(bars + 32);  // 4 warps x 64 channels x float2 (2 KB)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmW);
        tma_prefetch_desc(&p.tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(&fullA[i], 1);
            mbar_init(&emptyA[i], 1);
            mbar_init(&readyA[i], 4);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(&fullB[i], 1);
            mbar_init(&emptyB[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 8);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(tmem_ptr, p.tmem_cols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int pair, int& b, int& ty, int& tx, int& nt, bool& valid) {
        nt = pair % p.n_tiles_n;
        int m = (pair / p.n_tiles_n) * 2 + (int)rank;
        valid = m < p.m_tiles;
        tx = m % p.tiles_x;
        m /= p.tiles_x;
        ty = m % p.tiles_y;
        b = m / p.tiles_y;
    };

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs; whole warp, elected issue)
        {
            int ia = 0, ib = 0;
            uint32_t pha = 0, phb = 0;
            for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters) {
                int b, ty, tx, nt;
                bool valid;
                decode(pair, b, ty, tx, nt, valid);
                const int x0 = tx * kHaloTW, y0 = ty * kHaloTH * MT, n0 = nt * BN + (int)rank * HB;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    for (int cb = 0; cb < sg.cblocks; ++cb) {
                        mbar_wait(&emptyA[ia], pha ^ 1);
                        uint8_t* a_dst = ringA + (size_t)ia * p.a_slot;
                        if (PRO && p.seg_coef[s] != nullptr) {
                            // normalised segment: the transform warps fill this slot themselves (they wait on emptyA);
                            // only the weights of its taps are fetched here
                        } else if (PRO) {
                            // each CTA's box completes on its OWN barrier (its transform warps wait on it)
                            if (elect_one()) {
                                if (sg.taps == 9) {
                                    mbar_arrive_expect_tx(&fullA[ia], halo_bytes);
                                    tma_load_5d(a_dst, &p.tmA[s], &fullA[ia], cb * kBlockK, x0 - 1, 0, y0 - 1, b);
                                } else {
                                    mbar_arrive_expect_tx(&fullA[ia], MT * kABytes);
#pragma unroll
                                    for (int h = 0; h < MT; ++h)
                                        tma_load_5d(a_dst + h * kABytes, &p.tmA[s], &fullA[ia], cb * kBlockK, x0, 0,
                                                    y0 + h * kHaloTH, b);
                                }
                            }
                        } else if (elect_one()) {
                            if (sg.taps == 9) {
                                if (leader) mbar_arrive_expect_tx(&fullA[ia], 2 * halo_bytes);
                                if (p.cols3) {
#pragma unroll
                                    for (int j = 0; j < 3; ++j)
                                        tma_load_5d_2cta(a_dst + j * box3_bytes, &p.tmA[s], &fullA[ia], cb * kBlockK,
                                                         x0 - 1 + j, 0, y0 - 1, b);
                                } else {
                                    tma_load_5d_2cta(a_dst, &p.tmA[s], &fullA[ia], cb * kBlockK, x0 - 1, 0, y0 - 1, b);
                                }
                            } else {
                                if (leader) mbar_arrive_expect_tx(&fullA[ia], 2 * MT * kABytes);
#pragma unroll
                                for (int h = 0; h < MT; ++h)
                                    tma_load_5d_2cta(a_dst + h * kABytes, &p.tmA[s], &fullA[ia], cb * kBlockK, x0, 0,
                                                     y0 + h * kHaloTH, b);
                            }
                        }
                        __syncwarp();
                        if (++ia == SA) {
                            ia = 0;
                            pha ^= 1;
                        }
                        for (int tap = 0; tap < sg.ntaps; ++tap) {  // logical taps, in packed-weight order
                            mbar_wait(&emptyB[ib], phb ^ 1);
                            if (elect_one()) {
                                if (leader) mbar_arrive_expect_tx(&fullB[ib], 2 * b_bytes);
                                tma_load_2d_2cta(ringB + (size_t)ib * b_bytes, &p.tmW, &fullB[ib],
                                                 (p.seg_kb[s] + (int)((sg.wmap >> (4 * tap)) & 15ull) * sg.cblocks + cb) * kBlockK,
                                                 n0);
                            }
                            __syncwarp();
                            if (++ib == SB) {
                                ib = 0;
                                phb ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        if (leader) {  // whole warp, convergent; one elected lane issues each tcgen05 instruction
            const uint32_t idesc = umma_idesc_16b(2 * kTileM, (uint32_t)BN, 0, 0, p.a_fmt, p.w_fmt);
            int ia = 0, ib = 0;
            uint32_t pha = 0, phb = 0;
            int it = 0;
            for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BN);
                uint32_t acc = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    for (int cb = 0; cb < sg.cblocks; ++cb) {
                        if (PRO)
                            mbar_wait_cluster(&readyA[ia], pha);
                        else
                            mbar_wait(&fullA[ia], pha);
                        const uint32_t a_base = smem_u32(ringA + (size_t)ia * p.a_slot);
                        for (int ti = 0; ti < sg.ntaps; ++ti) {
                            const int tap = (int)((sg.tapmap >> (4 * ti)) & 15ull);  // 3x3 position of logical tap ti
                            mbar_wait(&fullB[ib], phb);
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(ringB + (size_t)ib * b_bytes);
                            uint32_t tap_off = 0u, a_sbo = 1024u, a_sub = (uint32_t)kABytes;
                            if (sg.taps == 9 && p.cols3) {
                                tap_off = (uint32_t)(tap % 3) * box3_bytes + (uint32_t)(tap / 3) * 1024u;
                                a_sub = (uint32_t)(kHaloTH * 1024);
                            } else if (sg.taps == 9) {
                                tap_off = (uint32_t)(((tap / 3) * kHaloPitch + (tap % 3)) * 128);
                                a_sbo = (uint32_t)(kHaloPitch * 128);
                                a_sub = (uint32_t)(kHaloTH * kHaloPitch * 128);
                            }
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
#pragma unroll
                                for (int h = 0; h < MT; ++h) {
                                    const uint64_t da = umma_smem_desc_sw128_off(a_base + h * a_sub + tap_off + k * 32, 16,
                                                                                 a_sbo, p.base_off_mode);
                                    if (elect_one()) umma_bf16_2cta(d_tmem + (uint32_t)(h * BN), da, db, idesc, acc);
                                }
                                acc = 1;
                            }
                            if (elect_one()) umma_commit_2cta(&emptyB[ib], 0x3);
                            __syncwarp();
                            if (++ib == SB) {
                                ib = 0;
                                phb ^= 1;
                            }
                        }
                        if (elect_one()) umma_commit_2cta(&emptyA[ia], 0x3);
                        __syncwarp();
                        if (++ia == SA) {
                            ia = 0;
                            pha ^= 1;
                        }
                    }
                }
                if (elect_one()) umma_commit_2cta(&tfull[as], 0x3);
                __syncwarp();
            }
        }
    } else if (warp < 4) {
        // ===================================================================== norm prologue (warps 2-3, both CTAs)
        if (PRO) {
            const int t = (int)threadIdx.x - 64;   // 0..63
            const int j = t & 7, rr = t >> 3;      // physical 16-byte chunk of the 128-byte row; row index mod 8
            const int c16 = j ^ rr;                // logical channel chunk behind the 128-byte swizzle (slots are 1 KB aligned)
            constexpr int rows9 = (kHaloTH * MT + 2) * kHaloPitch;
            int ia = 0;
            uint32_t pha = 0, full_bits = 0;
            for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters) {
                int b, ty, tx, nt;
                bool valid;
                decode(pair, b, ty, tx, nt, valid);
                const int x0 = tx * kHaloTW, y0 = ty * kHaloTH * MT;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    const float2* cf = p.seg_coef[s];
                    for (int cb = 0; cb < sg.cblocks; ++cb) {
                        const int c0 = cb * kBlockK + c16 * 8;
                        if (cf == nullptr) {
                            // raw segment (1x1 skip conv): TMA filled the slot; hand it on unchanged.  fullA[ia] completes only
                            // in the rounds in which the slot holds a raw segment, so its parity is tracked per slot.
                            mbar_wait(&fullA[ia], (full_bits >> ia) & 1u);
                            full_bits ^= 1u << ia;
                        } else {
                            // normalised 3x3 segment: global -> registers -> silu(x*A + Bc) -> swizzled shared memory.
                            // (Rewriting a TMA-filled tile in place was measured first: the extra shared-memory read + write
                            // competes with the MMAs' operand reads, which already run near the port's limit -- conv 22.9 ->
                            // 27 ms per evaluation even with the arithmetic removed.)
                            float Ah[8], Bh[8];
                            const bool chan_ok = valid && c0 + 8 <= sg.C;
                            if (chan_ok) {
                                const float4* src = reinterpret_cast<const float4*>(
                                    cf + (size_t)b * p.seg_coef_ld[s] + p.seg_coef_off[s] + c0);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float4 v = __ldg(src + e);
                                    Ah[2 * e] = 0.5f * v.x; Bh[2 * e] = 0.5f * v.y;
                                    Ah[2 * e + 1] = 0.5f * v.z; Bh[2 * e + 1] = 0.5f * v.w;
                                }
                            }
                            const uint16_t* gx = p.seg_x[s] + (size_t)b * p.Hout * p.Wout * sg.C + c0;
                            const uint32_t sbase = smem_u32(ringA + (size_t)ia * p.a_slot) + (uint32_t)(j * 16);
                            constexpr int U = 8;                              // rows per group and thread
                            constexpr int G = (rows9 + 8 * U - 1) / (8 * U);  // groups per box
                            int h = 0, w = rr;                                // box row / column of the next row to load
                            // two register buffers: the loads of group g+1 are in flight while group g is transformed and
                            // stored, and group 0 is requested BEFORE waiting for the slot (loads do not touch it)
                            auto load_group = [&](int g, uint4 (&v)[U], uint32_t& okm) {
                                okm = 0;
#pragma unroll
                                for (int i = 0; i < U; ++i) {
                                    const int r = rr + 8 * (g * U + i);
                                    const int yy = y0 - 1 + h, xx = x0 - 1 + w;
                                    const bool ok = chan_ok && r < rows9 && yy >= 0 && yy < p.Hout && xx >= 0 && xx < p.Wout;
                                    v[i] = make_uint4(0, 0, 0, 0);  // conv padding / channel tail / dummy tile: zeros
                                    if (ok) v[i] = ldg_nc16(gx + ((size_t)yy * p.Wout + xx) * sg.C);
                                    okm |= (uint32_t)ok << i;
                                    w += 8;
                                    if (w >= kHaloPitch) {
                                        w -= kHaloPitch;
                                        ++h;
                                    }
                                }
                            };
                            auto finish_group = [&](int g, uint4 (&v)[U], uint32_t okm) {
#pragma unroll
                                for (int i = 0; i < U; ++i) {
                                    if (!((okm >> i) & 1u)) continue;
                                    uint32_t in[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        // silu(z) = z/2 + z/2 * tanh(z/2): hz in fp32, one tanh.approx.f16x2 + one HFMA2 per pair
                                        const float2 xf = unpack_f16x2(in[e]);
                                        const uint32_t hz = pack_f16x2(fmaf(xf.x, Ah[2 * e], Bh[2 * e]),
                                                                       fmaf(xf.y, Ah[2 * e + 1], Bh[2 * e + 1]));
                                        uint32_t th;
                                        asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(hz));
                                        asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(in[e]) : "r"(hz), "r"(th));
                                    }
                                    v[i] = make_uint4(in[0], in[1], in[2], in[3]);
                                }
#pragma unroll
                                for (int i = 0; i < U; ++i) {
                                    const int r = rr + 8 * (g * U + i);
                                    if (r < rows9)
                                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)r * 128u),
                                                     "r"(v[i].x), "r"(v[i].y), "r"(v[i].z), "r"(v[i].w)
                                                     : "memory");
                                }
                            };
                            uint4 va[U], vb[U];
                            uint32_t oka, okb;
                            load_group(0, va, oka);
                            mbar_wait(&emptyA[ia], pha ^ 1);  // the MMAs that read this slot last have completed
#pragma unroll 1
                            for (int g = 0; g < G; g += 2) {
                                if (g + 1 < G) load_group(g + 1, vb, okb);
                                finish_group(g, va, oka);
                                if (g + 1 < G) {
                                    if (g + 2 < G) load_group(g + 2, va, oka);
                                    finish_group(g + 1, vb, okb);
                                }
                            }
                        }
                        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader_release(&readyA[ia]);
                        if (++ia == SA) {
                            ia = 0;
                            pha ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (both CTAs, own 128 rows)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;
        int it = 0;
        int ob = 0;
        const int nchunks = BN / 64;
        for (int pair = cluster_id; pair < p.total_pairs; pair += num_clusters, ++it) {
            int b, ty, tx, nt;
            bool valid;
            decode(pair, b, ty, tx, nt, valid);
            const int x0 = tx * kHaloTW, y0 = ty * kHaloTH * MT, n0 = nt * BN;
            const int px = x0 + row % kHaloTW;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < MT; ++h) {
                const int ys = y0 + h * kHaloTH;
                const int py = ys + row / kHaloTW;
                const bool in_img = valid && (py < p.Hout) && (px < p.Wout);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MT * BN + h * BN);
                for (int ch = 0; ch < nchunks; ++ch) {
                    // the residual tile's eight 16-byte vectors are requested BEFORE the accumulator is read: loaded one by
                    // one inside the conversion loop they cost eight serialised memory latencies per 128x64 chunk, which made the
                    // epilogue 2-3x longer than the chunk's MMAs (conv2 of the identity-skip ResBlocks: 2.3 ms instead of 0.8 ms)
                    uint4 rres[8];
                    const bool has_res = p.residual != nullptr;
                    if (has_res) {
                        const int nb = n0 + ch * 64;
                        const bool ok = in_img;
                        const uint4* rp = reinterpret_cast<const uint4*>(
                            p.residual + (((size_t)b * p.Hout + (ok ? py : 0)) * p.Wout + (ok ? px : 0)) * p.Cout + nb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) rres[j] = ok ? ldg_nc16(rp + j) : make_uint4(0, 0, 0, 0);
                    }
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32(t_row + ch * 64, v0);
                    tmem_ld_32x32(t_row + ch * 64 + 32, v1);
                    tmem_ld_wait();
                    if (h == MT - 1 && ch == nchunks - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tempty[as]);
                    }
                    const int nbase = n0 + ch * 64;
                    const bool res = has_res && in_img;
                    if (p.out_direct != nullptr) {
                        // Launches that need no statistics and write a dense tensor (every dgrad conv, the 1x1 GEMMs) store
                        // their pixel row straight from registers: 8 x 16 B = the thread's own 128-byte line.  No staging
                        // tile, no named barriers, no TMA store to wait for -- the four epilogue warps run independently.
                        // (Through shared memory + TMA the epilogue took ~2800 cycles per 128 x 64 chunk, so every conv with
                        // K < 1400 was bound by it: the 1x1 skip-conv dgrads ran at 177 TFLOP/s.)
                        if (in_img) {
                            uint4* o = reinterpret_cast<uint4*>(p.out_direct + ((((size_t)b * p.Hout + py) * p.Wout + px) * p.Cout + nbase));
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float f[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const int col = j * 8 + e;
                                    f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                                    if (p.bias) f[e] += __ldg(p.bias + nbase + col);
                                }
                                if (res) {
                                    const uint4 r = rres[j];
                                    float2 t;
                                    t = unpack2(r.x, p.res_fmt); f[0] += t.x; f[1] += t.y;
                                    t = unpack2(r.y, p.res_fmt); f[2] += t.x; f[3] += t.y;
                                    t = unpack2(r.z, p.res_fmt); f[4] += t.x; f[5] += t.y;
                                    t = unpack2(r.w, p.res_fmt); f[6] += t.x; f[7] += t.y;
                                }
                                o[j] = make_uint4(pack2(f[0], f[1], p.out_fmt), pack2(f[2], f[3], p.out_fmt),
                                                  pack2(f[4], f[5], p.out_fmt), pack2(f[6], f[7], p.out_fmt));
                            }
                        }
                        continue;
                    }
                    // staging buffer `ob` was last read by the TMA store issued two chunks ago; the converted vectors go
                    // straight into it (no register copy of the packed tile: the epilogue is register-bound)
                    if (et == 0) tma_store_wait_read<1>();
                    named_bar_sync(1, 128);
                    uint8_t* dst = out_stage + ob * kOutStageBytes + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 8 x (8 channels = 16 B)
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int col = j * 8 + e;
                            f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                            if (p.bias) f[e] += __ldg(p.bias + nbase + col);
                        }
                        if (res) {
                            const uint4 r = rres[j];
                            float2 t;
                            t = unpack2(r.x, p.res_fmt); f[0] += t.x; f[1] += t.y;
                            t = unpack2(r.y, p.res_fmt); f[2] += t.x; f[3] += t.y;
                            t = unpack2(r.z, p.res_fmt); f[4] += t.x; f[5] += t.y;
                            t = unpack2(r.w, p.res_fmt); f[6] += t.x; f[7] += t.y;
                        }
                        const int sw = j ^ (row & 7);  // 128 B swizzle: 16 B chunk index XOR (row mod 8)
                        *reinterpret_cast<uint4*>(dst + sw * 16) = make_uint4(pack2(f[0], f[1], p.out_fmt), pack2(f[2], f[3], p.out_fmt),
                                                                              pack2(f[4], f[5], p.out_fmt), pack2(f[6], f[7], p.out_fmt));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (et == 0) {
                        tma_store_5d(&p.tmOut, out_stage + ob * kOutStageBytes, nbase, x0, 0, ys, b);
                        tma_store_commit();
                    }
                    if (p.stats != nullptr) {
                        // as in conv_igemm_pair_kernel: per-channel (sum, sumsq) of the staged (rounded) tile; here the
                        // sub-tile is 8 px wide x 16 px high, row r = (r / 8, r % 8)
                        const int cg = et & 7, rg = et >> 3;
                        const uint8_t* tile = out_stage + ob * kOutStageBytes;
                        const bool full_tile = (ys + kHaloTH <= p.Hout) && (x0 + kHaloTW <= p.Wout);
                        float sm[8], sq[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) sm[e] = sq[e] = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rg * 8 + i;
                            if (!full_tile && ((ys + r / kHaloTW >= p.Hout) || (x0 + r % kHaloTW >= p.Wout))) continue;
                            const uint4 u = *reinterpret_cast<const uint4*>(tile + r * 128 + ((cg ^ i) << 4));
                            float f[8];
                            float2 t2;
                            t2 = unpack2(u.x, p.out_fmt); f[0] = t2.x; f[1] = t2.y;
                            t2 = unpack2(u.y, p.out_fmt); f[2] = t2.x; f[3] = t2.y;
                            t2 = unpack2(u.z, p.out_fmt); f[4] = t2.x; f[5] = t2.y;
                            t2 = unpack2(u.w, p.out_fmt); f[6] = t2.x; f[7] = t2.y;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                sm[e] += f[e];
                                sq[e] = fmaf(f[e], f[e], sq[e]);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 8);
                            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 8);
                            sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 16);
                            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 16);
                        }
                        if (lane < 8) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) stat_scratch[q * 64 + lane * 8 + e] = make_float2(sm[e], sq[e]);
                        }
                        named_bar_sync(3, 128);
                        if (et < 64 && valid) {
                            float2 o = stat_scratch[et];
#pragma unroll
                            for (int wq = 1; wq < 4; ++wq) {
                                const float2 t2 = stat_scratch[wq * 64 + et];
                                o.x += t2.x;
                                o.y += t2.y;
                            }
                            const int sub = p.stat_off + (ty * MT + h) * p.tiles_x + tx;
                            p.stats[((size_t)b * p.stat_tiles + sub) * p.Cout + nbase + et] = o;
                        }
                    }
                    ob ^= 1;
                }
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    }
}

// ====================================================================================================== wgrad
//   dW[tap][m][n] += sum_{pixels in this CTA's slice} P[pixel, m] * Q[pixel (+tap shift), n]
// P = the tensor indexing the GEMM M dimension (dY for a conv weight gradient), Q = the tensor indexing N (the conv
// input).  Both operands are "MN-major": a TMA box {64 ch, 16 w, 1, 8 h, 1} lands as [128 pixels][64 ch], 128 B
// swizzled, which is the canonical MN-major UMMA atom layout with K = pixels (SBO = 1024 B per 8 pixels,
// LBO = 16 KB to the next 64-channel atom).  Split-K over pixel tiles; partial sums leave through vectorised fp32
// reductions (red.global.add.v4.f32) into a [tap][M][N] fp32 buffer.
struct WgradParams {
    CUtensorMap tmP;  // 5-D {Cm, W, 1, H, B}      (output-side tensor, never shifted)
    CUtensorMap tmQ;  // 5-D parity view of the input-side tensor
    int taps, stride, Cq;       // Cq = channels of Q (parity offset for stride 2)
    unsigned long long tapmap;  // taps > 1: 3x3 position (dy*3 + dx) of logical tap i in bits [4i, 4i+4)
    int Mtot, Ntot;             // real extents (Cout, Cin of this segment)
    int BN;                     // 64, 128 or 256
    int m_tiles, n_tiles, splits;
    int tiles_x, tiles_y, pix_tiles;  // pixel tiles of the OUTPUT-side geometry
    int num_stages;
    uint32_t tmem_cols;
    float* dw;                  // [taps][Mtot][ldn] fp32, pre-zeroed or accumulating
    int ldn, n_off;             // row length of dw and column offset of this segment
    int mma_order;              // 1: split N into two independent accumulator halves, alternate between them
    int p_fmt, q_fmt;           // Fmt of P (usually bf16 gradients) and Q (usually fp16 saved activations)
};

__global__ void __launch_bounds__(kConvThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const int BN = p.BN;
    const uint32_t a_bytes = 2 * kABytes;                 // M = 128 -> two 64-channel atoms
    const uint32_t b_bytes = (uint32_t)(BN / 64) * kABytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int S = p.num_stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    // work item decode
    // tap fastest: the CTAs that share a pixel range (same split) and differ only in tap / m-tile / n-tile are
    // launched next to each other, run concurrently and walk the same dY / X tiles, so the 9x (taps) re-reads of both
    // operands are L2 hits instead of HBM traffic.
    int w = blockIdx.x;
    const int tap = w % p.taps;      w /= p.taps;
    const int mt = w % p.m_tiles;    w /= p.m_tiles;
    const int nt = w % p.n_tiles;    w /= p.n_tiles;
    const int split = w;
    const int per = (p.pix_tiles + p.splits - 1) / p.splits;
    const int kbeg = split * per;
    const int kend = min(p.pix_tiles, kbeg + per);
    const int nk = max(0, kend - kbeg);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmP);
        tma_prefetch_desc(&p.tmQ);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        {  // whole warp, convergent; one elected lane issues the TMA loads
            int sx = 0, sy = 0, cp = 0, coff = 0;
            if (p.taps > 1) {
                const int pos = (int)((p.tapmap >> (4 * tap)) & 15ull);
                const int dx = pos % 3, dy = pos / 3;
                if (p.stride == 1) {
                    sx = dx - 1;
                    sy = dy - 1;
                } else {
                    sx = (dx == 0 ? -1 : 0);
                    sy = (dy == 0 ? -1 : 0);
                    cp = (dy == 1) ? 0 : 1;
                    coff = ((dx == 1) ? 0 : 1) * p.Cq;
                }
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int k = kbeg; k < kend; ++k) {
                int m = k;
                const int tx = m % p.tiles_x;  m /= p.tiles_x;
                const int ty = m % p.tiles_y;
                const int b = m / p.tiles_y;
                const int x0 = tx * kTileW, y0 = ty * kTileH;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
                uint8_t* b_dst = a_dst + a_bytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[stage], stage_bytes);
                    tma_load_5d(a_dst, &p.tmP, &full[stage], mt * 128, x0, 0, y0, b);
                    tma_load_5d(a_dst + kABytes, &p.tmP, &full[stage], mt * 128 + 64, x0, 0, y0, b);
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_5d(b_dst + j * kABytes, &p.tmQ, &full[stage], coff + nt * BN + j * 64, x0 + sx, cp, y0 + sy,
                                    b);
                }
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (nk > 0) {  // whole warp, convergent; one elected lane issues each tcgen05 instruction
            const uint32_t idesc = umma_idesc_16b(128, (uint32_t)BN, 1, 1, p.p_fmt, p.q_fmt);
            const uint32_t idesc_half = umma_idesc_16b(128, (uint32_t)(BN / 2), 1, 1, p.p_fmt, p.q_fmt);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < nk; ++k) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
                if (p.mma_order == 0 || BN < 128) {
#pragma unroll
                    for (int kk = 0; kk < kTileM / 16; ++kk) {  // 16 pixels per MMA
                        const uint64_t da = umma_smem_desc_sw128(a_addr + kk * 2048, kABytes, 1024);
                        const uint64_t db = umma_smem_desc_sw128(b_addr + kk * 2048, kABytes, 1024);
                        if (elect_one()) umma_bf16(tmem_base, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
                    }
                } else {
                    const uint32_t hb = (uint32_t)(BN / 2);  // N halves = whole 64-channel atoms of the Q tile
#pragma unroll
                    for (int kk = 0; kk < kTileM / 16; ++kk) {
                        const uint64_t da = umma_smem_desc_sw128(a_addr + kk * 2048, kABytes, 1024);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t db = umma_smem_desc_sw128(b_addr + j * (hb / 64) * kABytes + kk * 2048, kABytes, 1024);
                            if (elect_one()) umma_bf16(tmem_base + j * hb, da, db, idesc_half, (k | kk) != 0 ? 1u : 0u);
                        }
                    }
                }
                if (elect_one()) umma_commit(&empty[stage]);
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit(tfull);
            __syncwarp();
        }
    } else if (warp >= 4 && nk > 0) {
        const int q = warp & 3;
        const int row = q * 32 + lane;  // m within the tile
        const int m = mt * 128 + row;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        float* dst_row = p.dw + ((size_t)tap * p.Mtot + m) * p.ldn + p.n_off + nt * BN;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c0, v);
            tmem_ld_wait();
            if (m < p.Mtot) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = nt * BN + c0 + j;
                    if (n + 3 < p.Ntot) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c0 + j),
                                     "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                     "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (n + e < p.Ntot) atomicAdd(dst_row + c0 + j + e, __uint_as_float(v[j + e]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ====================================================================================================== CTA-pair wgrad
// Weight gradient by CTA pairs (tcgen05 cta_group::2): one MMA covers M = 256 output channels (CTA r owns rows
// m0 + r*128 ... of dW and the matching half of the dY tile) x N = BN input channels, each CTA staging only HALF of the
// X tile.  Per CTA and pixel tile: 32 KB (dY half) + BN/2 * 256 B (X half) instead of 32 KB + BN * 256 B.
struct Wgrad2Params {
    CUtensorMap tmP, tmQ;
    int taps, stride, Cq;
    unsigned long long tapmap;
    int Mtot, Ntot, BN;
    int m_pairs, n_tiles, splits;
    int tiles_x, tiles_y, pix_tiles;
    int num_stages;
    uint32_t tmem_cols;
    float* dw;
    int ldn, n_off;
    int p_fmt, q_fmt;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
    conv_wgrad_pair_kernel(const __grid_constant__ Wgrad2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const int BN = p.BN;
    const int q_atoms = BN / 128;                           // 64-channel atoms of the X half this CTA stages
    const uint32_t a_bytes = 2 * kABytes;                   // 128 dY channels
    const uint32_t stage_bytes = a_bytes + (uint32_t)q_atoms * kABytes;
    const int S = p.num_stages;
    const uint32_t rank = blockIdx.x & 1u;  // == %cluster_ctarank for __cluster_dims__(2,1,1) on a 1-D grid (uniform)
    const bool leader = rank == 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

    int w = blockIdx.x >> 1;  // cluster id; tap fastest (see conv_wgrad_kernel)
    const int tap = w % p.taps;      w /= p.taps;
    const int mp = w % p.m_pairs;    w /= p.m_pairs;
    const int nt = w % p.n_tiles;    w /= p.n_tiles;
    const int split = w;
    const int per = (p.pix_tiles + p.splits - 1) / p.splits;
    const int kbeg = split * per;
    const int kend = min(p.pix_tiles, kbeg + per);
    const int nk = max(0, kend - kbeg);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmP);
        tma_prefetch_desc(&p.tmQ);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(tmem_ptr, p.tmem_cols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        {  // whole warp, convergent; one elected lane issues the TMA loads
            int sx = 0, sy = 0, cp = 0, coff = 0;
            if (p.taps > 1) {
                const int pos = (int)((p.tapmap >> (4 * tap)) & 15ull);
                const int dx = pos % 3, dy = pos / 3;
                if (p.stride == 1) {
                    sx = dx - 1;
                    sy = dy - 1;
                } else {
                    sx = (dx == 0 ? -1 : 0);
                    sy = (dy == 0 ? -1 : 0);
                    cp = (dy == 1) ? 0 : 1;
                    coff = ((dx == 1) ? 0 : 1) * p.Cq;
                }
            }
            const int m0 = mp * 256 + (int)rank * 128;
            const int n0 = coff + nt * BN + (int)rank * (BN / 2);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = kbeg; k < kend; ++k) {
                int m = k;
                const int tx = m % p.tiles_x;  m /= p.tiles_x;
                const int ty = m % p.tiles_y;
                const int b = m / p.tiles_y;
                const int x0 = tx * kTileW, y0 = ty * kTileH;
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
                uint8_t* b_dst = a_dst + a_bytes;
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * stage_bytes);
                    tma_load_5d_2cta(a_dst, &p.tmP, &full[stage], m0, x0, 0, y0, b);
                    tma_load_5d_2cta(a_dst + kABytes, &p.tmP, &full[stage], m0 + 64, x0, 0, y0, b);
                    for (int j = 0; j < q_atoms; ++j)
                        tma_load_5d_2cta(b_dst + j * kABytes, &p.tmQ, &full[stage], n0 + j * 64, x0 + sx, cp, y0 + sy, b);
                }
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (leader && nk > 0) {  // whole warp, convergent; one elected lane issues each tcgen05 instruction
            const uint32_t idesc = umma_idesc_16b(256, (uint32_t)BN, 1, 1, p.p_fmt, p.q_fmt);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < nk; ++k) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int kk = 0; kk < kTileM / 16; ++kk) {
                    const uint64_t da = umma_smem_desc_sw128(a_addr + kk * 2048, kABytes, 1024);
                    const uint64_t db = umma_smem_desc_sw128(b_addr + kk * 2048, kABytes, 1024);
                    if (elect_one()) umma_bf16_2cta(tmem_base, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
                }
                if (elect_one()) umma_commit_2cta(&empty[stage], 0x3);
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit_2cta(tfull, 0x3);
            __syncwarp();
        }
    } else if (warp >= 4 && nk > 0) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = mp * 256 + (int)rank * 128 + row;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        float* dst_row = p.dw + ((size_t)tap * p.Mtot + m) * p.ldn + p.n_off + nt * BN;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + c0, v);
            tmem_ld_wait();
            if (m < p.Mtot) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = nt * BN + c0 + j;
                    if (n + 3 < p.Ntot) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c0 + j),
                                     "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                     "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
                    } else {
                        for (int e = 0; e < 4; ++e)
                            if (n + e < p.Ntot) atomicAdd(dst_row + c0 + j + e, __uint_as_float(v[j + e]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    }
}

}  // namespace s2s
