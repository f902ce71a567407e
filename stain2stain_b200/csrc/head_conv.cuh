// The UNet's output projection `out = conv3x3(SiLU(GN(h)))`, C -> 3 channels (torchcfm UNetModel.out[2]), as
// its own kernel: fp32 NCHW output, optional fused Euler update x <- x + dt * v (the sampler's state update).
//
// N = 3 is tensor-core hostile and the layer is bound by reading its 128-channel input (1.07 GB at B = 64): the generic
// implicit-GEMM kernel fetched one A tile per filter tap -- nine reads of the input through L2 (1.2 ms per call, 7x the HBM
// time).  Here a CTA stages the (8+2) x (16+2) halo of its 8 x 16 pixel tile ONCE in shared memory (zero-padded borders) and
// all nine taps read it from there with ldmatrix; the 3 (padded to 8) output channels are the N of mma.sync.m16n8k16, one
// M = 16 tile = 16 pixels of an image row.  Persistent CTAs: the weights are converted to the activation format and laid out
// [tap][8][C] in shared memory once per CTA.
#pragma once
#include "attention.cuh"  // mma16816 / ldmatrix helpers
#include "common.cuh"

namespace s2s {

constexpr int kHeadTW = 16, kHeadTH = 8, kHeadThreads = 128;

struct HeadConvParams {
    const uint16_t* a;   // [B][H][W][C] activations (a_fmt)
    const float* w;      // [Cout][C][3][3] fp32 (the nn.Conv2d parameter itself)
    const float* bias;   // [Cout] or nullptr
    float* out;          // [B][Cout][H][W] fp32
    const float* axpy_x; // optional: out = axpy_x + axpy_a * (conv + bias)
    float axpy_a;
    int B, H, W, C, Cout;
    int tiles_x, tiles_y, total_tiles;
};

template <int AF>
__global__ void __launch_bounds__(kHeadThreads) head_conv_kernel(const HeadConvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int pitch = p.C + 8;                                 // halfs per pixel row (+16 B: conflict-free ldmatrix)
    uint16_t* halo = reinterpret_cast<uint16_t*>(smem_raw);    // [(8+2)*(16+2)][pitch]
    uint16_t* wsm = halo + (kHeadTH + 2) * (kHeadTW + 2) * pitch;  // [9][8][pitch]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int vec = p.C >> 3;
    // weights: fp32 OIHW -> [tap][n (8, rows >= Cout are zero)][c] in the activation format
    for (int i = threadIdx.x; i < 9 * 8 * p.C; i += kHeadThreads) {
        const int c = i % p.C, n = (i / p.C) & 7, tap = i / (8 * p.C);
        const float v = n < p.Cout ? __ldg(p.w + ((size_t)n * p.C + c) * 9 + tap) : 0.f;
        wsm[(tap * 8 + n) * pitch + c] = pack1(v, AF);
    }
    const int ksteps = p.C >> 4;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int m = tile;
        const int tx = m % p.tiles_x;  m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = m / p.tiles_y;
        const int x0 = tx * kHeadTW, y0 = ty * kHeadTH;
        __syncthreads();  // previous tile's reads of the halo are done (and, first time, the weights are visible)
        const uint16_t* src = p.a + (size_t)b * p.H * p.W * p.C;
        // halo -> shared memory, eight 16-byte loads in flight per thread (a load-then-store loop waits a full memory
        // latency per element: 22 serialised latencies per tile made the first version slower than the 9x re-read it replaces)
        const int nvec = (kHeadTH + 2) * (kHeadTW + 2) * vec;
        for (int base = threadIdx.x; base < nvec; base += kHeadThreads * 8) {
            uint4 u[8];
            int off[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + j * kHeadThreads;
                u[j] = make_uint4(0, 0, 0, 0);  // conv padding
                off[j] = -1;
                if (i < nvec) {
                    const int pix = i / vec, c8 = i - pix * vec;
                    const int hy = pix / (kHeadTW + 2), hx = pix - hy * (kHeadTW + 2);
                    const int yy = y0 - 1 + hy, xx = x0 - 1 + hx;
                    off[j] = pix * pitch + c8 * 8;
                    if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) u[j] = ldg_nc16(src + ((size_t)yy * p.W + xx) * p.C + c8 * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (off[j] >= 0) *reinterpret_cast<uint4*>(halo + off[j]) = u[j];
        }
        __syncthreads();
        // warp w: image rows y0 + 2w and y0 + 2w + 1, 16 pixels each; two accumulators per row (independent MMA chains)
        float acc[2][2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q) acc[r][q][0] = acc[r][q][1] = acc[r][q][2] = acc[r][q][3] = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint32_t wrow = smem_u32(wsm + (tap * 8 + (lane & 7)) * pitch + ((lane >> 3) & 1) * 8);
            const uint32_t arow0 = smem_u32(halo + ((2 * warp + dy) * (kHeadTW + 2) + dx + (lane & 15)) * pitch + (lane >> 4) * 8);
            const uint32_t arow1 = arow0 + (uint32_t)((kHeadTW + 2) * pitch * 2);
#pragma unroll 4
            for (int ks = 0; ks < ksteps; ++ks) {
                uint32_t b0, b1, a0[4], a1[4];
                ldsm_x2(b0, b1, wrow + ks * 32);
                ldsm_x4(a0, arow0 + ks * 32);
                ldsm_x4(a1, arow1 + ks * 32);
                mma16816<AF>(acc[0][ks & 1], a0, b0, b1);
                mma16816<AF>(acc[1][ks & 1], a1, b0, b1);
            }
        }
        // accumulator: rows g / g+8 = pixels x0+g / x0+g+8, columns 2*t4, 2*t4+1 = output channels
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int yy = y0 + 2 * warp + r;
            if (yy >= p.H) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int xx = x0 + g + (e >> 1) * 8, co = 2 * t4 + (e & 1);
                if (xx >= p.W || co >= p.Cout) continue;
                float v = acc[r][0][e] + acc[r][1][e] + (p.bias ? __ldg(p.bias + co) : 0.f);
                const size_t o = (((size_t)b * p.Cout + co) * p.H + yy) * p.W + xx;
                if (p.axpy_x) v = __ldg(p.axpy_x + o) + p.axpy_a * v;
                p.out[o] = v;
            }
        }
    }
}

}  // namespace s2s
