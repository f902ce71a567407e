// C-ABI entry points (include/s2s_b200.h): argument validation, TMA tensor-map construction and kernel launches.
#include "../../include/s2s_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "conv_igemm.cuh"
#include "attention.cuh"
#include "elementwise.cuh"
#include "head_conv.cuh"
#include "linear.cuh"
#include "multitask.cuh"
#include "optim.cuh"
#include "tiles.cuh"

using namespace s2s;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) return fail(S2S_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                               \
    do {                                                                                                 \
        cudaError_t _e = cudaGetLastError();                                                             \
        if (_e != cudaSuccess) return fail(S2S_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int mma_order() {  // S2S_MMA_ORDER=0 restores the one-accumulator-at-a-time issue order (A/B experiments)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_MMA_ORDER");
        v = e ? atoi(e) : 1;
    }
    return v;
}

int conv_pairs() {  // S2S_CONV_2CTA=0 forces the single-CTA conv kernel
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_CONV_2CTA");
        v = e ? atoi(e) : 1;
    }
    return v;
}

int wgrad_pairs() {  // S2S_WGRAD_2CTA=0 forces the single-CTA wgrad kernel
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_WGRAD_2CTA");
        v = e ? atoi(e) : 1;
    }
    return v;
}

int conv_halo() {  // S2S_CONV_HALO: 0 = off, 1 = halo-tiled A operand (matrix base offset set), 2 = same without base offset,
                   // 3 = three column-shifted halo boxes per channel block (all descriptors 1024 B aligned)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_CONV_HALO");
        v = e ? atoi(e) : 2;  // default: halo-tiled CTA-pair kernel for every stride-1 conv with a 3x3 segment
    }
    return v;
}

// Largest GEMM K for which a launch without statistics writes its output with plain 16-byte stores from registers instead of
// the shared-memory tile + TMA store (S2S_CONV_DIRECT_MAXK; 0 = never).  Short-K launches are bound by their epilogue and the
// register path removes its barriers and the TMA round trip (1x1 skip-conv dgrad 128 -> 128 at 256^2, B = 64: 0.775 -> 0.546
// ms); MMA-bound launches get slower with it (pixel-strided stores: the 3x3 dgrad 0.84 -> 0.92 ms), hence the limit.
int conv_direct_max_k() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_CONV_DIRECT_MAXK");
        v = e ? atoi(e) : 640;
    }
    return v;
}

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

// bf16 tensor map, 128 B swizzle.  dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(S2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(S2S_ERR_TMAP, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                    (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                    (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
                    rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return S2S_OK;
}

// Activation NHWC [B,H,W,C] bf16 as the 5-D parity view {C*s, W/s, s, H/s, B} (s = 1 or 2), box {64,16,1,8,1}.
int make_act_tmap(CUtensorMap* m, const void* x, int B, int H, int W, int C, int s) {
    if (C % 8) return fail(S2S_ERR_INVALID, "activation channels (%d) must be a multiple of 8", C);
    if (s != 1 && s != 2) return fail(S2S_ERR_INVALID, "stride %d unsupported", s);
    if (s == 2 && ((H | W) & 1)) return fail(S2S_ERR_INVALID, "stride-2 conv needs even H, W (got %d x %d)", H, W);
    cuuint64_t dims[5] = {(cuuint64_t)C * s, (cuuint64_t)W / s, (cuuint64_t)s, (cuuint64_t)H / s, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)C * s * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2 * s,
                         (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {64, kTileW, 1, kTileH, 1};
    return make_tmap(m, x, 5, dims, str, box);
}

uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return S2S_OK;
}

constexpr size_t kSmemBudget = 227 * 1024;

// run-time flag / format -> template argument
#define S2S_BOOL(cond, NAME, ...)                 \
    do {                                          \
        if (cond) {                               \
            constexpr bool NAME = true;           \
            __VA_ARGS__;                          \
        } else {                                  \
            constexpr bool NAME = false;          \
            __VA_ARGS__;                          \
        }                                         \
    } while (0)
#define S2S_DROP(drop_p, mask, NAME, ...)         \
    do {                                          \
        if (!((drop_p) > 0.f)) {                  \
            constexpr int NAME = 0;               \
            __VA_ARGS__;                          \
        } else if ((mask) == nullptr) {           \
            constexpr int NAME = 1;               \
            __VA_ARGS__;                          \
        } else {                                  \
            constexpr int NAME = 2;               \
            __VA_ARGS__;                          \
        }                                         \
    } while (0)
#define S2S_ACT(act, NAME, ...)                   \
    do {                                          \
        if ((act) == 1) {                         \
            constexpr int NAME = kActSilu;        \
            __VA_ARGS__;                          \
        } else if ((act) == 2) {                  \
            constexpr int NAME = kActRelu;        \
            __VA_ARGS__;                          \
        } else {                                  \
            constexpr int NAME = kActNone;        \
            __VA_ARGS__;                          \
        }                                         \
    } while (0)
#define S2S_FMT(fmt, NAME, ...)                   \
    do {                                          \
        if ((fmt) == S2S_FMT_F16) {               \
            constexpr int NAME = kFmtF16;         \
            __VA_ARGS__;                          \
        } else {                                  \
            constexpr int NAME = kFmtBF16;        \
            __VA_ARGS__;                          \
        }                                         \
    } while (0)

int ew_grid(long long work_items, int threads = kEwThreads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Pixels per CTA for the (chunk, sample) grids of the normalisation kernels.  The kernels run at 2, 3 or 4 CTAs per SM
// (register-bound), so the grid B x chunks is made a multiple of 148 * lcm(2,3,4) = 1776 CTAs whenever the tensor is
// big enough: every wave is full and there is no tail wave (a 2.05-wave grid costs 3 waves).  Depends on (B, HW) only,
// so that every source of a channel concat is cut into the same pixel chunks (their partial statistics line up).
// A CTA is never given fewer than 256 pixels (128 below 64^2): with the 111 chunks the wave rule asks for at B = 64, a
// 64^2 sample was cut into 37-pixel CTAs whose life is one latency chain (coefficients -> loads -> stores) -- 0.29-0.58
// of the HBM rate at 64^2 / 32^2 against 0.64-0.82 with fat CTAs (profiles/r02_kbench_gn_min_ppc.txt).
// Pixels per thread and loop iteration (= independent 16-byte loads in flight per tensor) of the streaming norm kernels, with
// the CTAs per SM the register count then allows.  Same-box A/B inside the training step (profiles/r02_ab_gn_unroll.txt):
//   gn_bwd_apply  2 @ 3 CTAs/SM: 0.82 of HBM   4 @ 2 CTAs/SM: 0.87  <- memory-latency bound, more loads in flight win
//   gn_bwd_reduce 2 @ 4 CTAs/SM: 0.84          4 @ 3 CTAs/SM: 0.78  <- occupancy wins
//   gn_apply      4 @ 4 CTAs/SM: 0.90 / 0.85 (dropout)   8 @ 3 CTAs/SM: 0.90 / 0.78
constexpr int kGnBwdApplyU = 4, kGnBwdReduceU = 2, kGnApplyU = 4;
int gn_bulk() {  // S2S_GN_BULK=0: register-load version of gn_bwd_apply everywhere (A/B; profiles/r02_ab_gn_bulk.txt)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_GN_BULK");
        v = e ? atoi(e) : 1;
    }
    return v;
}
int gn_bulk_min_hw() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_GN_BULK_MIN_HW");
        v = e ? atoi(e) : 16384;
    }
    return v;
}
int gn_min_ppc() {  // S2S_GN_MIN_PPC: fewest pixels a CTA of the normalisation kernels is given (experiments)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_GN_MIN_PPC");
        v = e ? atoi(e) : 0;  // 0 = the measured default below
    }
    return v;
}
int pick_pix_per_cta(int B, int HW, int /*C*/) {
    // pure geometry: without a device (host-only callers of s2s_gn_chunks, the CPU test tier) the B200's 148 SMs are assumed
    const long long wave = (long long)(num_sms() > 0 ? num_sms() : 148) * 12;
    long long a = B, b = wave;
    while (b) { long long t = a % b; a = b; b = t; }  // a = gcd(B, wave)
    const long long base = wave / a;                  // smallest chunk count with (B * chunks) % wave == 0
    long long k = (2 * wave + (long long)B * base - 1) / ((long long)B * base);  // aim at >= 2 * wave CTAs in total
    long long chunks = base * (k < 1 ? 1 : k);
    const int min_ppc = gn_min_ppc() > 0 ? gn_min_ppc() : (HW >= 4096 ? 256 : 128);
    const long long max_chunks = HW / min_ppc < 1 ? 1 : HW / min_ppc;
    if (chunks > max_chunks) chunks = max_chunks;
    long long ppc = (HW + chunks - 1) / chunks;
    if (ppc < 32) ppc = 32;  // tiny tensors: launch latency dominates, keep some work per CTA
    if (ppc > HW) ppc = HW;
    return (int)ppc;
}

int check_vec_layout(int C, const char* who) {
    if (C <= 0 || C % 8 || C / 8 > kEwThreads)
        return fail(S2S_ERR_INVALID, "%s: C = %d unsupported (need C %% 8 == 0 and C <= %d)", who, C, 8 * kEwThreads);
    return S2S_OK;
}
// block size of the (slot, pixel-row) mapping: vpp * floor(256 / vpp) threads
int vec_threads(int C) {
    const int vpp = C / 8;
    return vpp * (kEwThreads / vpp);
}

}  // namespace

extern "C" {

const char* s2s_last_error(void) { return g_err; }
int s2s_abi_version(void) { return 1; }
int s2s_num_sms(void) { return num_sms(); }

int s2s_pack_conv_weight(const float* w, int Cout, int Cin, int taps, int ci_begin, int ci_count, void* dst, int ld_k,
                         int k_off, int transpose_flip, int fmt, void* stream) {
    return s2s_pack_conv_weight_mode(w, Cout, Cin, taps, ci_begin, ci_count, dst, ld_k, k_off, transpose_flip, fmt, 0, stream);
}

int s2s_pack_conv_weight_mode(const float* w, int Cout, int Cin, int taps, int ci_begin, int ci_count, void* dst, int ld_k,
                              int k_off, int transpose_flip, int fmt, int mode, void* stream) {
    if (!w || !dst || Cout <= 0 || ci_count <= 0 || ci_begin < 0 || ci_begin + ci_count > Cin || (taps != 1 && taps != 9))
        return fail(S2S_ERR_INVALID, "pack_conv_weight: bad arguments");
    if (mode < 0 || mode > 4 || (mode != 0 && taps != 9))
        return fail(S2S_ERR_INVALID, "pack_conv_weight: mode %d (phase-summed taps) needs a 3x3 weight", mode);
    PackJob jb;
    jb.w = w; jb.dst = (uint16_t*)dst; jb.Cout = Cout; jb.Cin = Cin; jb.taps = taps; jb.ci_begin = ci_begin;
    jb.ci_count = ci_count; jb.ld_k = ld_k; jb.k_off = k_off; jb.transpose_flip = transpose_flip; jb.fmt = fmt; jb.mode = mode;
    pack_conv_weight_kernel<<<pack_tile_count(Cout, ci_count, transpose_flip), 256, 0, (cudaStream_t)stream>>>(jb);
    LAUNCH_CHECK("pack_conv_weight_kernel");
    return S2S_OK;
}

// geometry the CTA-pair conv kernel uses for an output [*, Hout, Wout, Cout]; false = that kernel does not apply
static bool pair_geometry(int Hout, int Wout, int Cout, int* BN, int* mt, int* tiles_x, int* tiles_y) {
    if (Cout % 128 != 0 || !conv_pairs()) return false;
    *BN = (Cout % 256 == 0) ? 256 : 128;
    *mt = (*BN == 128 && Hout >= 2 * kTileH) ? 2 : 1;
    *tiles_x = (Wout + kTileW - 1) / kTileW;
    *tiles_y = (Hout + kTileH * *mt - 1) / (kTileH * *mt);
    return true;
}

// the halo-tiled CTA-pair kernel takes stride-1 convs with at least one 3x3 segment (1x1 segments ride along)
static bool halo_eligible(const s2s_conv_src* srcs, int nsrc, int Cout) {
    if (!conv_halo() || !conv_pairs() || Cout % 128 != 0 || !srcs) return false;
    bool any3 = false;
    for (int s = 0; s < nsrc; ++s) {
        if (srcs[s].stride != 1 || (srcs[s].taps != 1 && srcs[s].taps != 9)) return false;
        any3 = any3 || srcs[s].taps == 9;
    }
    return any3;
}
static void halo_geometry(int Hout, int Wout, int Cout, int* BN, int* mt, int* tiles_x, int* tiles_y) {
    *BN = (Cout % 256 == 0) ? 256 : 128;
    *mt = (conv_halo() != 3 && *BN == 128 && Hout >= 2 * kHaloTH) ? 2 : 1;
    *tiles_x = (Wout + kHaloTW - 1) / kHaloTW;
    *tiles_y = (Hout + kHaloTH * *mt - 1) / (kHaloTH * *mt);
}

int s2s_conv_stat_tiles(int Hout, int Wout, int Cout) {
    int BN, mt, tx, ty;
    if (!pair_geometry(Hout, Wout, Cout, &BN, &mt, &tx, &ty)) return 0;
    return tx * ty * mt;
}

static int epi_stats_min_k() {  // S2S_EPI_STATS_MINK: smallest GEMM K for which the conv epilogue also emits the statistics
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("S2S_EPI_STATS_MINK");
        v = e ? atoi(e) : 0;  // measured neutral-to-negative at 1800 (same box: 179.1 vs 178.6 ms/step, 24.13 vs 24.22 tiles/s): off
    }
    return v;
}

// geometry only: sub-tiles of the statistics the kernel chosen for these segments CAN emit
int s2s_conv_stat_tiles_geom(const s2s_conv_src* srcs, int nsrc, int Hout, int Wout, int Cout) {
    int BN, mt, tx, ty;
    if (halo_eligible(srcs, nsrc, Cout)) {
        halo_geometry(Hout, Wout, Cout, &BN, &mt, &tx, &ty);
        return tx * ty * mt;
    }
    return s2s_conv_stat_tiles(Hout, Wout, Cout);
}

int s2s_conv_stat_tiles_for(const s2s_conv_src* srcs, int nsrc, int Hout, int Wout, int Cout) {
    // Experiment switch (S2S_EPI_STATS_MINK, default 0 = off): convs with a short GEMM K are bound by their epilogue (the
    // 128 -> 128 convs at 256^2, K = 1152, take 1.35 ms with the statistics and 0.84 ms without, a separate statistics pass
    // 0.18 ms), so such layers could answer "no statistics" and let the consumer run s2s_gn_stats.  Measured on one box,
    // back to back, the whole step / evaluation did not get faster (the extra pass re-reads tensors the epilogue had for
    // free under the power cap), so the rule is off.
    int ktot = 0;
    for (int s = 0; srcs && s < nsrc; ++s) ktot += srcs[s].taps * ((srcs[s].C + kBlockK - 1) / kBlockK) * kBlockK;
    if (srcs && ktot < epi_stats_min_k()) return 0;
    return s2s_conv_stat_tiles_geom(srcs, nsrc, Hout, Wout, Cout);
}

// A 16-bit NHWC tensor seen through arbitrary pixel strides: [B][H][W][C] with strides (sb, sy, sx) in ELEMENTS.  The
// phase views of an Upsample conv's output (pixels (2y+py, 2x+px)) are such views; so is any dense tensor.
struct ActView {
    const void* base;
    int B, H, W, C;
    long long sx, sy, sb;
};
static ActView dense_view(const void* x, int B, int H, int W, int C) {
    ActView v;
    v.base = x; v.B = B; v.H = H; v.W = W; v.C = C;
    v.sx = C; v.sy = (long long)W * C; v.sb = (long long)H * W * C;
    return v;
}
// pixels (2y + py, 2x + px) of a dense [B, 2H, 2W, C] tensor as a [B, H, W, C] view
static ActView phase_view(const void* x, int B, int H, int W, int C, int py, int px) {
    ActView v;
    v.base = (const uint8_t*)x + ((size_t)py * 2 * W + px) * C * 2;
    v.B = B; v.H = H; v.W = W; v.C = C;
    v.sx = 2LL * C; v.sy = 4LL * W * C; v.sb = 4LL * H * W * C;
    return v;
}
static int make_view_tmap(CUtensorMap* m, const ActView& v, int box_w, int box_h) {
    if (v.C % 8) return fail(S2S_ERR_INVALID, "activation channels (%d) must be a multiple of 8", v.C);
    cuuint64_t dims[5] = {(cuuint64_t)v.C, (cuuint64_t)v.W, 1, (cuuint64_t)v.H, (cuuint64_t)v.B};
    cuuint64_t str[4] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sb * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
    return make_tmap(m, v.base, 5, dims, str, box);
}

struct HaloSeg {
    ActView view;
    int kind;                   // 9: 3x3 neighbourhood (halo box), 1: 1x1
    int ntaps;                  // logical taps multiplied (kind 9: 1..9, kind 1: 1)
    unsigned long long tapmap;  // 3x3 position of each logical tap
    unsigned long long wmap;    // weight tap slot of each logical tap (kTapIdentity: packed in logical order)
    int wslots;                 // tap slots this segment occupies in the packed weight (0 = ntaps)
};

// One launch of the halo-tiled CTA-pair kernel over explicit views.  Output view dims = the GEMM's pixel space.
static int halo_launch(const HaloSeg* segs, int nseg, const ActView& outv, const void* w_packed, int Ktot, int kb_first,
                       int Cout, const float* bias, const void* residual, float* stats_out, int stat_tiles_total,
                       int stat_off, int a_fmt, int w_fmt, int out_fmt, int res_fmt, void* stream,
                       const s2s_conv_norm* norms, int act) {
    const int B = outv.B, Hout = outv.H, Wout = outv.W;
    Conv3Params q;
    memset(&q, 0, sizeof(q));
    const bool cols3 = conv_halo() == 3;
    int BN3, mt3, tx3, ty3;
    halo_geometry(Hout, Wout, Cout, &BN3, &mt3, &tx3, &ty3);
    q.cols3 = cols3 ? 1 : 0;
    q.nseg = nseg;
    int kb3 = kb_first;
    for (int s = 0; s < nseg; ++s) {
        const HaloSeg& sc = segs[s];
        if (sc.view.B != B || sc.view.H != Hout || sc.view.W != Wout)
            return fail(S2S_ERR_INVALID, "conv(halo): segment %d view does not match the output pixel space", s);
        int rc = sc.kind == 9 ? make_view_tmap(&q.tmA[s], sc.view, cols3 ? kHaloTW : kHaloPitch, kHaloTH * mt3 + 2)
                              : make_view_tmap(&q.tmA[s], sc.view, kHaloTW, kHaloTH);
        if (rc) return rc;
        q.seg[s].taps = sc.kind;
        q.seg[s].ntaps = sc.ntaps;
        q.seg[s].tapmap = sc.tapmap;
        q.seg[s].wmap = sc.wmap;
        q.seg[s].cblocks = (sc.view.C + kBlockK - 1) / kBlockK;
        q.seg[s].stride = 1;
        q.seg[s].C = sc.view.C;
        q.seg_kb[s] = kb3;
        kb3 += (sc.wslots > 0 ? sc.wslots : sc.ntaps) * q.seg[s].cblocks;
    }
    if (kb3 * kBlockK > Ktot)
        return fail(S2S_ERR_INVALID, "conv_fwd: Ktot = %d is smaller than the segments need (%d)", Ktot, kb3 * kBlockK);
    {
        cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
        cuuint64_t str[1] = {(cuuint64_t)Ktot * 2};
        cuuint32_t box[2] = {kBlockK, (cuuint32_t)(BN3 / 2)};
        int rc = make_tmap(&q.tmW, w_packed, 2, dims, str, box);
        if (rc) return rc;
    }
    int rc = make_view_tmap(&q.tmOut, outv, kHaloTW, kHaloTH);
    if (rc) return rc;
    q.B = B; q.Hout = Hout; q.Wout = Wout; q.Cout = Cout;
    q.tiles_x = tx3;
    q.tiles_y = ty3;
    q.stats = (float2*)stats_out;
    q.stat_tiles = stat_tiles_total > 0 ? stat_tiles_total : tx3 * ty3 * mt3;
    q.stat_off = stat_off;
    const bool dense_out = outv.sx == Cout && outv.sy == (long long)Wout * Cout && outv.sb == (long long)Hout * Wout * Cout;
    q.out_direct = (!stats_out && dense_out && (kb3 - kb_first) * kBlockK <= conv_direct_max_k())
                       ? (uint16_t*)const_cast<void*>(outv.base) : nullptr;
    if (norms) {
        q.prologue = 1;
        q.act = act;
        for (int s = 0; s < nseg; ++s) {
            q.seg_coef[s] = (const float2*)norms[s].coef;
            q.seg_x[s] = (const uint16_t*)segs[s].view.base;
            q.seg_coef_ld[s] = norms[s].ld;
            q.seg_coef_off[s] = norms[s].off;
        }
    }
    q.m_tiles = B * q.tiles_x * q.tiles_y;
    q.n_tiles_n = Cout / BN3;
    q.total_pairs = ((q.m_tiles + 1) / 2) * q.n_tiles_n;
    q.BN = BN3;
    q.tmem_cols = pow2_cols(2 * mt3 * BN3);
    q.base_off_mode = conv_halo() == 1 ? 1 : 0;
    q.bias = bias;
    q.residual = (const __nv_bfloat16*)residual;
    q.a_fmt = a_fmt; q.w_fmt = w_fmt; q.out_fmt = out_fmt; q.res_fmt = res_fmt;
    const size_t halo_bytes = cols3 ? (size_t)3 * (kHaloTH * mt3 + 2) * kHaloTW * 128
                                    : (size_t)(kHaloTH * mt3 + 2) * kHaloPitch * 128;
    size_t a_slot = halo_bytes > (size_t)mt3 * kABytes ? halo_bytes : (size_t)mt3 * kABytes;
    a_slot = (a_slot + 1023) / 1024 * 1024;
    q.a_slot = (uint32_t)a_slot;
    const size_t b_bytes = (size_t)(BN3 / 2) * kBlockK * 2;
    const size_t fixed = 2 * kOutStageBytes + 1024 + 3072;  // + alignment slack + barriers / statistics scratch
    q.sa = cols3 ? 2 : 3;
    int sb = (int)((kSmemBudget - fixed - (size_t)q.sa * a_slot) / b_bytes);
    if (sb > 8) sb = 8;
    if (sb < 2) return fail(S2S_ERR_INVALID, "conv_fwd(halo): tile does not fit in shared memory");
    q.sb = sb;
    const size_t smem = (size_t)q.sa * a_slot + (size_t)sb * b_bytes + fixed;
    int clusters = num_sms() / 2;
    if (clusters > q.total_pairs) clusters = q.total_pairs;
    const dim3 grid3(2 * clusters);
    cudaStream_t st3 = (cudaStream_t)stream;
#define S2S_HALO_LAUNCH(MTV, PROV)                                                          \
    do {                                                                                    \
        rc = set_smem(conv_halo_pair_kernel<MTV, PROV>, smem);                              \
        if (rc) return rc;                                                                  \
        conv_halo_pair_kernel<MTV, PROV><<<grid3, kConvThreads, smem, st3>>>(q);            \
    } while (0)
    if (mt3 == 2 && q.prologue) S2S_HALO_LAUNCH(2, true);
    else if (mt3 == 2) S2S_HALO_LAUNCH(2, false);
    else if (q.prologue) S2S_HALO_LAUNCH(1, true);
    else S2S_HALO_LAUNCH(1, false);
#undef S2S_HALO_LAUNCH
    LAUNCH_CHECK("conv_halo_pair_kernel");
    return S2S_OK;
}

static int conv_fwd_impl(const s2s_conv_src* srcs, int nsrc, int B, int Hout, int Wout, const void* w_packed, int Ktot,
                         int Cout, const float* bias, const void* residual, void* out_bf16, float* out_f32,
                         const float* axpy_x, float axpy_a, float* stats_out, int a_fmt, int w_fmt, int out_fmt,
                         int res_fmt, void* stream, const s2s_conv_norm* norms, int act);

int s2s_conv_fwd(const s2s_conv_src* srcs, int nsrc, int B, int Hout, int Wout, const void* w_packed, int Ktot,
                 int Cout, const float* bias, const void* residual, void* out_bf16, float* out_f32,
                 const float* axpy_x, float axpy_a, float* stats_out, int a_fmt, int w_fmt, int out_fmt, int res_fmt,
                 void* stream) {
    return conv_fwd_impl(srcs, nsrc, B, Hout, Wout, w_packed, Ktot, Cout, bias, residual, out_bf16, out_f32, axpy_x, axpy_a,
                         stats_out, a_fmt, w_fmt, out_fmt, res_fmt, stream, nullptr, 0);
}

int s2s_conv_norm_fusable(const s2s_conv_src* srcs, int nsrc, int Cout) {
    return (conv_halo() == 1 || conv_halo() == 2) && halo_eligible(srcs, nsrc, Cout) ? 1 : 0;
}

int s2s_conv_fwd_norm(const s2s_conv_src* srcs, const s2s_conv_norm* norms, int act, int nsrc, int B, int Hout, int Wout,
                      const void* w_packed, int Ktot, int Cout, const float* bias, const void* residual, void* out_bf16,
                      float* stats_out, int a_fmt, int w_fmt, int out_fmt, int res_fmt, void* stream) {
    if (!norms || act != 1 || a_fmt != S2S_FMT_F16)
        return fail(S2S_ERR_INVALID, "conv_fwd_norm: the fused prologue is built for SiLU on fp16 activations");
    if (!s2s_conv_norm_fusable(srcs, nsrc, Cout))
        return fail(S2S_ERR_INVALID, "conv_fwd_norm: this geometry does not run on the halo-tiled CTA-pair kernel");
    for (int s = 0; s < nsrc; ++s)
        if (norms[s].coef && (srcs[s].taps != 9 || norms[s].ld % 8 || norms[s].off % 8))
            return fail(S2S_ERR_INVALID, "conv_fwd_norm: a normalised segment must be 3x3 with 8-aligned coefficient offsets");
    return conv_fwd_impl(srcs, nsrc, B, Hout, Wout, w_packed, Ktot, Cout, bias, residual, out_bf16, nullptr, nullptr, 0.f,
                         stats_out, a_fmt, w_fmt, out_fmt, res_fmt, stream, norms, act);
}

static int conv_fwd_impl(const s2s_conv_src* srcs, int nsrc, int B, int Hout, int Wout, const void* w_packed, int Ktot,
                         int Cout, const float* bias, const void* residual, void* out_bf16, float* out_f32,
                         const float* axpy_x, float axpy_a, float* stats_out, int a_fmt, int w_fmt, int out_fmt,
                         int res_fmt, void* stream, const s2s_conv_norm* norms, int act) {
    if (stats_out && (!out_bf16 || axpy_x || s2s_conv_stat_tiles_geom(srcs, nsrc, Hout, Wout, Cout) == 0))
        return fail(S2S_ERR_INVALID, "conv_fwd: epilogue statistics need a CTA-pair path (s2s_conv_stat_tiles_for() > 0)");
    if (nsrc < 1 || nsrc > kMaxSeg) return fail(S2S_ERR_INVALID, "conv_fwd: nsrc = %d (1..%d)", nsrc, kMaxSeg);
    if (a_fmt != w_fmt)
        return fail(S2S_ERR_INVALID, "conv_fwd: activations and weights must share one 16-bit format (tcgen05 kind::f16 "
                                     "rejects mixed fp16 x bf16 operands)");
    if ((out_bf16 != nullptr) == (out_f32 != nullptr))
        return fail(S2S_ERR_INVALID, "conv_fwd: exactly one of out_bf16 / out_f32 must be given");
    // halo-tiled CTA-pair kernel: stride-1 convs with at least one 3x3 segment
    if (out_bf16 && !axpy_x && halo_eligible(srcs, nsrc, Cout)) {
        HaloSeg hs[kMaxSeg];
        int kb = 0;
        for (int s = 0; s < nsrc; ++s) {
            hs[s].view = dense_view(srcs[s].x, B, Hout, Wout, srcs[s].C);
            hs[s].kind = srcs[s].taps;
            hs[s].ntaps = srcs[s].taps;
            hs[s].tapmap = srcs[s].taps == 9 ? kTapIdentity : 0ull;
            hs[s].wmap = kTapIdentity;
            hs[s].wslots = 0;
            kb += srcs[s].taps * ((srcs[s].C + kBlockK - 1) / kBlockK);
        }
        if (kb * kBlockK != Ktot)
            return fail(S2S_ERR_INVALID, "conv_fwd: Ktot = %d does not match the segments (%d)", Ktot, kb * kBlockK);
        return halo_launch(hs, nsrc, dense_view(out_bf16, B, Hout, Wout, Cout), w_packed, Ktot, 0, Cout, bias, residual,
                           stats_out, 0, 0, a_fmt, w_fmt, out_fmt, res_fmt, stream, norms, act);
    }
    // CTA-pair kernel (tcgen05 cta_group::2): 16-bit NHWC outputs with Cout a multiple of 128
    if (out_bf16 && !axpy_x && Cout % 128 == 0 && conv_pairs()) {
        Conv2Params q;
        memset(&q, 0, sizeof(q));
        q.nseg = nsrc;
        int kb2 = 0;
        for (int s = 0; s < nsrc; ++s) {
            const s2s_conv_src& sc = srcs[s];
            if (sc.taps != 1 && sc.taps != 9) return fail(S2S_ERR_INVALID, "conv_fwd: taps = %d", sc.taps);
            if (sc.taps == 1 && sc.stride != 1) return fail(S2S_ERR_INVALID, "conv_fwd: strided 1x1 unsupported");
            int rc = make_act_tmap(&q.tmA[s], sc.x, B, Hout * sc.stride, Wout * sc.stride, sc.C, sc.stride);
            if (rc) return rc;
            q.seg[s].taps = sc.taps;
            q.seg[s].cblocks = (sc.C + kBlockK - 1) / kBlockK;
            q.seg[s].stride = sc.stride;
            q.seg[s].C = sc.C;
            kb2 += sc.taps * q.seg[s].cblocks;
        }
        if (kb2 * kBlockK != Ktot)
            return fail(S2S_ERR_INVALID, "conv_fwd: Ktot = %d does not match the segments (%d)", Ktot, kb2 * kBlockK);
        int BN2, mt2, tiles_x2, tiles_y2;
        pair_geometry(Hout, Wout, Cout, &BN2, &mt2, &tiles_x2, &tiles_y2);
        {
            cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
            cuuint64_t str[1] = {(cuuint64_t)Ktot * 2};
            cuuint32_t box[2] = {kBlockK, (cuuint32_t)(BN2 / 2)};
            int rc = make_tmap(&q.tmW, w_packed, 2, dims, str, box);
            if (rc) return rc;
        }
        int rc = make_act_tmap(&q.tmOut, out_bf16, B, Hout, Wout, Cout, 1);
        if (rc) return rc;
        q.B = B; q.Hout = Hout; q.Wout = Wout; q.Cout = Cout;
        q.tiles_x = tiles_x2;
        q.mt = mt2;
        q.tiles_y = tiles_y2;
        q.m_tiles = B * q.tiles_x * q.tiles_y;
        q.stats = (float2*)stats_out;
        q.out_direct = (!stats_out && Ktot <= conv_direct_max_k()) ? (uint16_t*)out_bf16 : nullptr;
        q.stat_tiles = tiles_x2 * tiles_y2 * mt2;
        q.m_pairs = (q.m_tiles + 1) / 2;
        q.n_tiles_n = Cout / BN2;
        q.total_pairs = q.m_pairs * q.n_tiles_n;
        q.BN = BN2;
        q.kblocks = kb2;
        q.tmem_cols = pow2_cols(2 * mt2 * BN2);
        q.bias = bias;
        q.residual = (const __nv_bfloat16*)residual;
        q.a_fmt = a_fmt; q.w_fmt = w_fmt; q.out_fmt = out_fmt; q.res_fmt = res_fmt;
        const size_t stage_bytes = (size_t)mt2 * kABytes + (size_t)(BN2 / 2) * kBlockK * 2;
        const size_t fixed = 2 * kOutStageBytes + 1024 + 3072;  // + alignment slack + barriers / statistics scratch
        int stages = (int)((kSmemBudget - fixed) / stage_bytes);
        if (stages > 8) stages = 8;
        q.num_stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + fixed;
        int clusters = num_sms() / 2;
        if (clusters > q.total_pairs) clusters = q.total_pairs;
        if (mt2 == 2) {
            rc = set_smem(conv_igemm_pair_kernel<2>, smem);
            if (rc) return rc;
            conv_igemm_pair_kernel<2><<<2 * clusters, kConvThreads, smem, (cudaStream_t)stream>>>(q);
        } else {
            rc = set_smem(conv_igemm_pair_kernel<1>, smem);
            if (rc) return rc;
            conv_igemm_pair_kernel<1><<<2 * clusters, kConvThreads, smem, (cudaStream_t)stream>>>(q);
        }
        LAUNCH_CHECK("conv_igemm_pair_kernel");
        return S2S_OK;
    }
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = nsrc;
    int kblocks = 0;
    for (int s = 0; s < nsrc; ++s) {
        const s2s_conv_src& sc = srcs[s];
        if (sc.taps != 1 && sc.taps != 9) return fail(S2S_ERR_INVALID, "conv_fwd: taps = %d", sc.taps);
        if (sc.taps == 1 && sc.stride != 1) return fail(S2S_ERR_INVALID, "conv_fwd: strided 1x1 unsupported");
        int rc = make_act_tmap(&p.tmA[s], sc.x, B, Hout * sc.stride, Wout * sc.stride, sc.C, sc.stride);
        if (rc) return rc;
        p.seg[s].taps = sc.taps;
        p.seg[s].cblocks = (sc.C + kBlockK - 1) / kBlockK;
        p.seg[s].stride = sc.stride;
        p.seg[s].C = sc.C;
        kblocks += sc.taps * p.seg[s].cblocks;
    }
    if (kblocks * kBlockK != Ktot)
        return fail(S2S_ERR_INVALID, "conv_fwd: Ktot = %d does not match the segments (%d)", Ktot, kblocks * kBlockK);
    int BN;
    if (out_f32) {
        if (Cout > 16) return fail(S2S_ERR_INVALID, "conv_fwd: fp32 NCHW output supports Cout <= 16 (got %d)", Cout);
        BN = 16;
        p.mode = kModeF32Nchw;
    } else {
        if (Cout % 64) return fail(S2S_ERR_INVALID, "conv_fwd: bf16 output needs Cout %% 64 == 0 (got %d)", Cout);
        BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
        p.mode = kModeBf16Nhwc;
        if (axpy_x) return fail(S2S_ERR_INVALID, "conv_fwd: axpy needs the fp32 output mode");
    }
    const int npad = (Cout + BN - 1) / BN * BN;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)npad};
        cuuint64_t str[1] = {(cuuint64_t)Ktot * 2};
        cuuint32_t box[2] = {kBlockK, (cuuint32_t)BN};
        int rc = make_tmap(&p.tmW, w_packed, 2, dims, str, box);
        if (rc) return rc;
    }
    if (out_bf16) {
        int rc = make_act_tmap(&p.tmOut, out_bf16, B, Hout, Wout, Cout, 1);
        if (rc) return rc;
    }
    p.B = B; p.Hout = Hout; p.Wout = Wout; p.Cout = Cout;
    // narrow-N layers (Cout <= 128) would re-fetch the weight tile once per 128 pixels and run into the shared-memory /
    // L2 feed limit (A + B = 32 KB per 256 MMA cycles); two pixel sub-tiles per CTA tile share one weight tile instead.
    const int mt = (!out_f32 && BN <= 128 && Hout >= 2 * kTileH) ? 2 : 1;
    p.mt = mt;
    // measured (kbench, B200): alternating the two pixel sub-tiles' accumulators is +4 %; splitting N = 256 into halves
    // is -12 % for the K-major forward operands (A is then read twice from shared memory)
    p.mma_order = (mt == 2) ? mma_order() : 0;
    p.tiles_x = (Wout + kTileW - 1) / kTileW;
    p.tiles_y = (Hout + kTileH * mt - 1) / (kTileH * mt);
    p.n_tiles_n = npad / BN;
    p.total_tiles = B * p.tiles_x * p.tiles_y * p.n_tiles_n;
    p.BN = BN;
    p.kblocks = kblocks;
    p.tmem_cols = pow2_cols(2 * mt * BN);
    p.bias = bias;
    p.residual = (const __nv_bfloat16*)residual;
    p.out_f32 = out_f32;
    p.axpy_x = axpy_x;
    p.axpy_a = axpy_a;
    p.a_fmt = a_fmt; p.w_fmt = w_fmt; p.out_fmt = out_fmt; p.res_fmt = res_fmt;
    const size_t b_bytes = (size_t)(BN < 64 ? 64 : BN) * kBlockK * 2;
    const size_t stage_bytes = (size_t)mt * kABytes + b_bytes;
    const size_t fixed = 2 * kOutStageBytes + 1024 /*alignment slack*/ + 512 /*barriers*/;
    int stages = (int)((kSmemBudget - fixed) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(S2S_ERR_INVALID, "conv_fwd: tile does not fit in shared memory");
    p.num_stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    int rc = set_smem(conv_igemm_kernel, smem);
    if (rc) return rc;
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    conv_igemm_kernel<<<grid, kConvThreads, smem, (cudaStream_t)stream>>>(p);
    LAUNCH_CHECK("conv_igemm_kernel");
    return S2S_OK;
}

static int wgrad_impl(const ActView& dyv, const void* x, int Cq, int taps, unsigned long long tapmap, int stride,
                      float* dw, int ldn, int n_off, int dy_fmt, int x_fmt, void* stream);

int s2s_conv_wgrad(const void* dy, int Cm, const void* x, int Cq, int taps, int stride, int B, int Hout, int Wout,
                   float* dw, int ldn, int n_off, int dy_fmt, int x_fmt, void* stream) {
    if (taps != 1 && taps != 9) return fail(S2S_ERR_INVALID, "conv_wgrad: taps = %d", taps);
    return wgrad_impl(dense_view(dy, B, Hout, Wout, Cm), x, Cq, taps, taps == 9 ? kTapIdentity : 0ull, stride, dw, ldn, n_off,
                      dy_fmt, x_fmt, stream);
}

// dyv: the output-side tensor as a (possibly strided) view; x: dense [B, H*stride, W*stride, Cq]; taps logical taps at the
// 3x3 positions of `tapmap`
static int wgrad_impl(const ActView& dyv, const void* x, int Cq, int taps, unsigned long long tapmap, int stride,
                      float* dw, int ldn, int n_off, int dy_fmt, int x_fmt, void* stream) {
    const int B = dyv.B, Hout = dyv.H, Wout = dyv.W, Cm = dyv.C;
    if (dy_fmt != x_fmt)
        return fail(S2S_ERR_INVALID, "conv_wgrad: dy and x must share one 16-bit format (convert with s2s_convert16)");
    if (Cq % 64) return fail(S2S_ERR_INVALID, "conv_wgrad: input channels must be a multiple of 64 (got %d)", Cq);
    if (ldn % 4 || n_off % 4) return fail(S2S_ERR_INVALID, "conv_wgrad: ldn / n_off must be multiples of 4");
    if (Cm % 256 == 0 && Cq % 128 == 0 && wgrad_pairs()) {  // CTA-pair kernel: M = 256 output channels per MMA
        Wgrad2Params q;
        memset(&q, 0, sizeof(q));
        int rc = make_view_tmap(&q.tmP, dyv, kTileW, kTileH);
        if (rc) return rc;
        rc = make_act_tmap(&q.tmQ, x, B, Hout * stride, Wout * stride, Cq, stride);
        if (rc) return rc;
        q.taps = taps; q.stride = stride; q.Cq = Cq; q.Mtot = Cm; q.Ntot = Cq;
        q.tapmap = tapmap;
        q.BN = (Cq % 256 == 0) ? 256 : 128;
        q.m_pairs = Cm / 256;
        q.n_tiles = Cq / q.BN;
        q.tiles_x = (Wout + kTileW - 1) / kTileW;
        q.tiles_y = (Hout + kTileH - 1) / kTileH;
        q.pix_tiles = B * q.tiles_x * q.tiles_y;
        const int mn = taps * q.m_pairs * q.n_tiles;
        int splits = (num_sms() / 2) / mn;
        if (splits > q.pix_tiles) splits = q.pix_tiles;
        if (splits < 1) splits = 1;
        q.splits = splits;
        q.tmem_cols = pow2_cols(q.BN);
        q.dw = dw; q.ldn = ldn; q.n_off = n_off;
        q.p_fmt = dy_fmt; q.q_fmt = x_fmt;
        const size_t stage_bytes = 2 * kABytes + (size_t)(q.BN / 128) * kABytes;
        const size_t fixed = 1024 + 512;
        int stages = (int)((kSmemBudget - fixed) / stage_bytes);
        if (stages > 6) stages = 6;
        q.num_stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + fixed;
        rc = set_smem(conv_wgrad_pair_kernel, smem);
        if (rc) return rc;
        conv_wgrad_pair_kernel<<<2 * mn * splits, kConvThreads, smem, (cudaStream_t)stream>>>(q);
        LAUNCH_CHECK("conv_wgrad_pair_kernel");
        return S2S_OK;
    }
    WgradParams p;
    memset(&p, 0, sizeof(p));
    int rc = make_view_tmap(&p.tmP, dyv, kTileW, kTileH);
    if (rc) return rc;
    rc = make_act_tmap(&p.tmQ, x, B, Hout * stride, Wout * stride, Cq, stride);
    if (rc) return rc;
    p.taps = taps; p.stride = stride; p.Cq = Cq; p.Mtot = Cm; p.Ntot = Cq;
    p.tapmap = tapmap;
    p.BN = (Cq % 256 == 0) ? 256 : (Cq % 128 == 0 ? 128 : 64);
    p.m_tiles = (Cm + 127) / 128;
    p.n_tiles = Cq / p.BN;
    p.tiles_x = (Wout + kTileW - 1) / kTileW;
    p.tiles_y = (Hout + kTileH - 1) / kTileH;
    p.pix_tiles = B * p.tiles_x * p.tiles_y;
    const int mn = taps * p.m_tiles * p.n_tiles;
    // one CTA per SM (shared memory bound): size the split-K factor so that the grid is at most ONE full wave --
    // 297 CTAs on 148 SMs would cost three waves for two waves of work.
    int splits = num_sms() / mn;
    if (splits > p.pix_tiles) splits = p.pix_tiles;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.tmem_cols = pow2_cols(p.BN);
    p.dw = dw; p.ldn = ldn; p.n_off = n_off;
    p.p_fmt = dy_fmt; p.q_fmt = x_fmt;
    // measured: two independent N = 128 halves are +12 % for the MN-major wgrad operands at BN = 256, -28 % at BN = 128
    p.mma_order = (p.BN == 256) ? mma_order() : 0;
    const size_t stage_bytes = 2 * kABytes + (size_t)(p.BN / 64) * kABytes;
    const size_t fixed = 1024 + 512;
    int stages = (int)((kSmemBudget - fixed) / stage_bytes);
    if (stages > 6) stages = 6;
    p.num_stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    rc = set_smem(conv_wgrad_kernel, smem);
    if (rc) return rc;
    conv_wgrad_kernel<<<mn * splits, kConvThreads, smem, (cudaStream_t)stream>>>(p);
    LAUNCH_CHECK("conv_wgrad_kernel");
    return S2S_OK;
}

int s2s_pack_tiles(int Cout, int ci_count, int transpose_flip) {
    if (Cout <= 0 || ci_count <= 0) return 0;
    return pack_tile_count(Cout, ci_count, transpose_flip);
}

int s2s_pack_conv_weight_multi(const s2s_pack_job* jobs_dev, const int* work_dev, int n_work, void* stream) {
    static_assert(sizeof(s2s_pack_job) == sizeof(PackJob), "ABI struct drifted from the kernel's");
    if (n_work <= 0) return S2S_OK;
    if (!jobs_dev || !work_dev) return fail(S2S_ERR_INVALID, "pack_conv_weight_multi: bad arguments");
    pack_conv_weight_multi_kernel<<<n_work, 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs_dev, (const int2*)work_dev);
    LAUNCH_CHECK("pack_conv_weight_multi_kernel");
    return S2S_OK;
}

int s2s_unpack_wgrad(const float* dw, int taps, int M, int ldn, int n_off, int n_count, float* grad, int Cin_total,
                     int n_begin, float beta, void* stream) {
    const long long total = (long long)M * n_count;  // one thread per (m, n), all taps
    unpack_wgrad_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>(dw, taps, M, ldn, n_off, n_count, grad,
                                                                                  Cin_total, n_begin, beta);
    LAUNCH_CHECK("unpack_wgrad_kernel");
    return S2S_OK;
}

// ------------------------------------------------------------------------------------------------ phase-decomposed Upsample conv
// torchcfm Upsample = F.interpolate(x, scale_factor=2, mode="nearest") -> conv3x3(pad 1).  For output pixels (2y+py, 2x+px)
// that is a 2x2 conv over the LOW-resolution tensor with tap-summed weights (s2s_pack_conv_weight_mode, mode 1 + phase):
// the 4x tensor is never materialised and 4/9 of the MACs remain -- in forward, dgrad and wgrad alike.
// Logical tap ti = ai*2 + bi of phase (py, px) reads low-res pixel (y + ai - 1 + py, x + bi - 1 + px).
static unsigned long long upconv_tapmap(int py, int px, bool mirrored) {
    unsigned long long m = 0;
    for (int ti = 0; ti < 4; ++ti) {
        const int ai = ti >> 1, bi = ti & 1;
        int ry = ai + py, rx = bi + px;  // 3x3 position (offset + 1) of the tap in the low-res neighbourhood
        if (mirrored) { ry = 2 - ry; rx = 2 - rx; }
        m |= (unsigned long long)(ry * 3 + rx) << (4 * ti);
    }
    return m;
}

int s2s_upconv_stat_tiles(int H, int W, int Cout) {
    if (!conv_halo() || !conv_pairs() || Cout % 128 != 0) return 0;
    int BN, mt, tx, ty;
    halo_geometry(H, W, Cout, &BN, &mt, &tx, &ty);
    return 4 * tx * ty * mt;
}

int s2s_upconv_supported(int C, int Cout) {
    return (conv_halo() && conv_pairs() && Cout % 128 == 0 && C % 128 == 0 && C % 64 == 0) ? 1 : 0;
}

int s2s_upconv_fwd(const void* x, int B, int H, int W, int C, const void* w_packed, int Cout, const float* bias, void* out,
                   float* stats_out, int a_fmt, int w_fmt, int out_fmt, void* stream) {
    if (!x || !w_packed || !out) return fail(S2S_ERR_INVALID, "upconv_fwd: null argument");
    if (!s2s_upconv_supported(C, Cout))
        return fail(S2S_ERR_INVALID, "upconv_fwd: needs C and Cout multiples of 128 and the halo CTA-pair kernel (C=%d Cout=%d)", C, Cout);
    if (a_fmt != w_fmt) return fail(S2S_ERR_INVALID, "upconv_fwd: activations and weights must share one 16-bit format");
    const int cblocks = C / kBlockK;
    const int Ktot = 16 * cblocks * kBlockK;
    const int st_total = stats_out ? s2s_upconv_stat_tiles(H, W, Cout) : 0;
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        HaloSeg seg;
        seg.view = dense_view(x, B, H, W, C);
        seg.kind = 9;
        seg.ntaps = 4;
        seg.tapmap = upconv_tapmap(py, px, false);
        seg.wmap = kTapIdentity;
        seg.wslots = 0;
        int rc = halo_launch(&seg, 1, phase_view(out, B, H, W, Cout, py, px), w_packed, Ktot, ph * 4 * cblocks, Cout, bias,
                             nullptr, stats_out, st_total, ph * (st_total / 4), a_fmt, w_fmt, out_fmt, out_fmt, stream,
                             nullptr, 0);
        if (rc) return rc;
    }
    return S2S_OK;
}

int s2s_upconv_dgrad(const void* dy, int B, int H, int W, int Cm, const void* w_packed, int Cin, void* dx, int a_fmt,
                     int w_fmt, int out_fmt, void* stream) {
    if (!dy || !w_packed || !dx) return fail(S2S_ERR_INVALID, "upconv_dgrad: null argument");
    if (!s2s_upconv_supported(Cm, Cin))
        return fail(S2S_ERR_INVALID, "upconv_dgrad: needs channel counts that are multiples of 128 (Cm=%d Cin=%d)", Cm, Cin);
    if (a_fmt != w_fmt) return fail(S2S_ERR_INVALID, "upconv_dgrad: gradients and weights must share one 16-bit format");
    const int cblocks = Cm / kBlockK;
    HaloSeg segs[4];
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        segs[ph].view = phase_view(dy, B, H, W, Cm, py, px);
        segs[ph].kind = 9;
        segs[ph].ntaps = 4;
        segs[ph].tapmap = upconv_tapmap(py, px, true);  // adjoint: the tap that read (y + a, x + b) scatters back from (y - a, x - b)
        segs[ph].wmap = kTapIdentity;
        segs[ph].wslots = 0;
    }
    return halo_launch(segs, 4, dense_view(dx, B, H, W, Cin), w_packed, 16 * cblocks * kBlockK, 0, Cin, nullptr, nullptr,
                       nullptr, 0, 0, a_fmt, w_fmt, out_fmt, out_fmt, stream, nullptr, 0);
}

int s2s_upconv_wgrad(const void* dy, int Cm, const void* x, int Cq, int B, int H, int W, float* dw16, int dy_fmt, int x_fmt,
                     void* stream) {
    if (!dy || !x || !dw16) return fail(S2S_ERR_INVALID, "upconv_wgrad: null argument");
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        int rc = wgrad_impl(phase_view(dy, B, H, W, Cm, py, px), x, Cq, 4, upconv_tapmap(py, px, false), 1,
                            dw16 + (size_t)ph * 4 * Cm * Cq, Cq, 0, dy_fmt, x_fmt, stream);
        if (rc) return rc;
    }
    return S2S_OK;
}

int s2s_upconv_unpack_wgrad(const float* dw16, int M, int N, float* grad_oihw, void* stream) {
    if (!dw16 || !grad_oihw) return fail(S2S_ERR_INVALID, "upconv_unpack_wgrad: null argument");
    upconv_unpack_wgrad_kernel<<<ew_grid((long long)M * N * 9), kEwThreads, 0, (cudaStream_t)stream>>>(dw16, M, N, grad_oihw);
    LAUNCH_CHECK("upconv_unpack_wgrad_kernel");
    return S2S_OK;
}

// ------------------------------------------------------------------------------------------------ dgrad of the stride-2 conv
// Downsample = conv3x3(stride 2, pad 1): input pixel i = 2*o + d - 1.  The data gradient at input pixels (2y+py, 2x+px)
// only receives the taps whose parity matches: py = 0 -> d = 1 (from output row y); py = 1 -> d = 0 (row y+1) and d = 2
// (row y).  So the four input phases are 1-, 2-, 2- and 4-tap convs over the LOW-resolution gradient, written through
// strided TMA store maps: the 9 algorithmic taps exactly, instead of 36 taps' worth over a zero-inserted 4x tensor.
// w_packed is the ordinary dgrad operand ([Cin][9*Cm], slot t = filter tap 8 - t): the tap slots are permuted by `wmap`.
int s2s_downconv_dgrad(const void* dy, int B, int H, int W, int Cm, const void* w_packed, int Cin, void* dx, int a_fmt,
                       int w_fmt, int out_fmt, void* stream) {
    if (!dy || !w_packed || !dx) return fail(S2S_ERR_INVALID, "downconv_dgrad: null argument");
    if (!s2s_upconv_supported(Cm, Cin))
        return fail(S2S_ERR_INVALID, "downconv_dgrad: needs channel counts that are multiples of 128 (Cm=%d Cin=%d)", Cm, Cin);
    if (a_fmt != w_fmt) return fail(S2S_ERR_INVALID, "downconv_dgrad: gradients and weights must share one 16-bit format");
    const int cblocks = Cm / kBlockK;
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        HaloSeg seg;
        seg.view = dense_view(dy, B, H, W, Cm);
        seg.kind = 9;
        seg.ntaps = 0;
        seg.tapmap = 0;
        seg.wmap = 0;
        seg.wslots = 9;
        for (int d = 0; d < 3; ++d) {
            if ((d & 1) == (py & 1)) continue;      // needs d == py + 1 (mod 2)
            const int oy = (py == 1 && d == 0) ? 1 : 0;  // output row offset the tap reads: y + oy
            for (int e = 0; e < 3; ++e) {
                if ((e & 1) == (px & 1)) continue;
                const int ox = (px == 1 && e == 0) ? 1 : 0;
                seg.tapmap |= (unsigned long long)((oy + 1) * 3 + (ox + 1)) << (4 * seg.ntaps);
                seg.wmap |= (unsigned long long)(8 - (d * 3 + e)) << (4 * seg.ntaps);
                ++seg.ntaps;
            }
        }
        int rc = halo_launch(&seg, 1, phase_view(dx, B, H, W, Cin, py, px), w_packed, 9 * cblocks * kBlockK, 0, Cin, nullptr,
                             nullptr, nullptr, 0, 0, a_fmt, w_fmt, out_fmt, out_fmt, stream, nullptr, 0);
        if (rc) return rc;
    }
    return S2S_OK;
}

int s2s_patch27_pack(const float* x0, const float* x1, const float* t, int B, int H, int W, int sgn, void* dst,
                     float* xt_out, int fmt, void* stream) {
    if (!x0 || !dst || (x1 && !t) || (sgn != 1 && sgn != -1)) return fail(S2S_ERR_INVALID, "patch27_pack: bad arguments");
    const long long npix = (long long)B * H * W;
    patch27_pack_kernel<<<ew_grid(npix, 128), 128, 0, (cudaStream_t)stream>>>(x0, x1, t, B, H, W, sgn,
                                                                              (__nv_bfloat16*)dst, xt_out, fmt);
    LAUNCH_CHECK("patch27_pack_kernel");
    return S2S_OK;
}

int s2s_gn_stats(const void* x, int B, int HW, int C, float* stats, int Ctot, int c_off, int x_fmt, void* stream) {
    int rc = check_vec_layout(C, "gn_stats");
    if (rc) return rc;
    const int ppc = pick_pix_per_cta(B, HW, C);
    dim3 grid((HW + ppc - 1) / ppc, B);
    S2S_FMT(x_fmt, XF, (gn_stats_kernel<XF><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
                           (const __nv_bfloat16*)x, C, HW, ppc, (float2*)stats, Ctot, c_off)));
    LAUNCH_CHECK("gn_stats_kernel");
    return S2S_OK;
}

int s2s_gn_chunks(int B, int HW) {
    const int ppc = pick_pix_per_cta(B, HW, 0);
    return (HW + ppc - 1) / ppc;
}

int s2s_gn_coef(const float* stats, const float* gamma, const float* beta, const float* film, int B, int C, int G,
                int HW, float eps, float* coef, float* mean_rstd, void* stream) {
    if (G > 64 || C % G) return fail(S2S_ERR_INVALID, "gn_coef: G = %d, C = %d unsupported", G, C);
    if (C > 4096) return fail(S2S_ERR_INVALID, "gn_coef: C = %d unsupported (<= 4096)", C);
    gn_coef_kernel<<<B, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>((const float2*)stats, s2s_gn_chunks(B, HW), gamma, beta, film,
                                                        C, G, HW, eps, (float2*)coef, (float2*)mean_rstd);
    LAUNCH_CHECK("gn_coef_kernel");
    return S2S_OK;
}

int s2s_gn_coef_parts(const float* stats0, int C0, const float* stats1, int C1, int nchunks, const float* gamma,
                      const float* beta, const float* film, int B, int G, int HW, float eps, float* coef,
                      float* mean_rstd, void* stream) {
    const int C = C0 + C1;
    if (G > 64 || C % G || !stats0 || C0 <= 0 || (C1 > 0 && !stats1))
        return fail(S2S_ERR_INVALID, "gn_coef_parts: G = %d, C = %d + %d unsupported", G, C0, C1);
    if (C > 4096) return fail(S2S_ERR_INVALID, "gn_coef_parts: C = %d unsupported (<= 4096)", C);
    if ((C0 | C1) & 1) return fail(S2S_ERR_INVALID, "gn_coef_parts: channel counts must be even");
    const int threads = 1024;
    int slices = threads / (C / 2);
    if (slices < 1) slices = 1;
    if (slices > 16) slices = 16;
    if (slices > nchunks) slices = nchunks;
    const size_t smem = (size_t)(2 * C) * (1 + slices) * sizeof(float);  // <= 2 * 4096 * 2 * 4 = 64 KB at slices = 1
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(gn_coef_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gn_coef_parts_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(
        (const float2*)stats0, C0, (const float2*)stats1, C1, nchunks, gamma, beta, film, G, HW, eps, (float2*)coef,
        (float2*)mean_rstd, slices);
    LAUNCH_CHECK("gn_coef_parts_kernel");
    return S2S_OK;
}

int s2s_gn_apply(const void* x, int B, int HW, int C, const float* coef, int Ctot, int c_off, void* y, void* y2_bf16,
                 int ld_out, int silu, float drop_p, uint64_t seed, void* mask_out, int x_fmt, int y_fmt, void* stream) {
    return s2s_gn_apply_step(x, B, HW, C, coef, Ctot, c_off, y, y2_bf16, ld_out, silu, drop_p, seed, nullptr, mask_out, x_fmt,
                             y_fmt, stream);
}

int s2s_gn_apply_step(const void* x, int B, int HW, int C, const float* coef, int Ctot, int c_off, void* y, void* y2_bf16,
                      int ld_out, int silu, float drop_p, uint64_t seed, const uint64_t* seed_step_dev, void* mask_out,
                      int x_fmt, int y_fmt, void* stream) {
    int rc = check_vec_layout(C, "gn_apply");
    if (rc) return rc;
    if (ld_out % 8 || c_off % 8) return fail(S2S_ERR_INVALID, "gn_apply: ld_out / c_off must be multiples of 8");
    const int ppc = pick_pix_per_cta(B, HW, C);
    dim3 grid((HW + ppc - 1) / ppc, B);
    S2S_ACT(silu, SILU, S2S_BOOL(drop_p > 0.f, DROP, S2S_BOOL(y2_bf16 != nullptr, DUAL, S2S_FMT(x_fmt, XF, S2S_FMT(y_fmt, YF,
        (gn_apply_kernel<SILU, DROP, XF, YF, DUAL, kGnApplyU><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, C, HW, ppc, (const float2*)coef, Ctot, c_off, (__nv_bfloat16*)y,
            (__nv_bfloat16*)y2_bf16, ld_out, drop_p, seed, (uint8_t*)mask_out,
            (const unsigned long long*)seed_step_dev)))))));
    LAUNCH_CHECK("gn_apply_kernel");
    return S2S_OK;
}

int s2s_gn_bwd_reduce_x2(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef,
                         const float* mean_rstd, int G, int Ctot, int c_off, float* red, int silu, float drop_p,
                         uint64_t seed, const void* mask_in, void* x_bf16_out, int x_fmt, int g_fmt, void* stream) {
    int rc = check_vec_layout(C, "gn_bwd_reduce");
    if (rc) return rc;
    const int ppc = pick_pix_per_cta(B, HW, C);
    dim3 grid((HW + ppc - 1) / ppc, B);
    // gradients are bf16 everywhere in the engine (fp32's range, no loss scaling): the fp16-gradient variants of the two
    // streaming backward kernels were never launched and doubled their 70-odd instantiations
    if (g_fmt != S2S_FMT_BF16) return fail(S2S_ERR_INVALID, "gn_bwd_reduce: the gradient format must be bf16");
    constexpr int GF = kFmtBF16;
    S2S_ACT(silu, SILU, S2S_DROP(drop_p, mask_in, DROP, S2S_BOOL(x_bf16_out != nullptr, X16, S2S_FMT(x_fmt, XF,
        (gn_bwd_reduce_kernel<SILU, DROP, XF, GF, X16, kGnBwdReduceU><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)g, ld_g, C, HW, ppc, (const float2*)coef,
            (const float2*)mean_rstd, G, Ctot, c_off, (float2*)red, drop_p, seed, (const uint8_t*)mask_in,
            (__nv_bfloat16*)x_bf16_out))))));
    LAUNCH_CHECK("gn_bwd_reduce_kernel");
    return S2S_OK;
}

int s2s_gn_bwd_reduce(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef,
                      const float* mean_rstd, int G, int Ctot, int c_off, float* red, int silu, float drop_p,
                      uint64_t seed, const void* mask_in, int x_fmt, int g_fmt, void* stream) {
    return s2s_gn_bwd_reduce_x2(x, g, ld_g, B, HW, C, coef, mean_rstd, G, Ctot, c_off, red, silu, drop_p, seed, mask_in,
                                nullptr, x_fmt, g_fmt, stream);
}

int s2s_gn_bwd_coef(const float* red_part, float* red, const float* mean_rstd, const float* gamma, const float* beta,
                    const float* film, int B, int C, int G, int HW, float* pqr, float* dgamma, float* dbeta,
                    float* dfilm, void* stream) {
    if (G > 64 || C % G || (C & 1)) return fail(S2S_ERR_INVALID, "gn_bwd_coef: G = %d, C = %d unsupported", G, C);
    if (C > 4096) return fail(S2S_ERR_INVALID, "gn_bwd_coef: C = %d unsupported (<= 4096)", C);
    const int nchunks = s2s_gn_chunks(B, HW);
    const int threads = 1024;
    int slices = threads / (C / 2);
    if (slices < 1) slices = 1;
    if (slices > 16) slices = 16;
    if (slices > nchunks) slices = nchunks;
    const size_t smem = (size_t)(2 * C) * (1 + slices) * sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(gn_bwd_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gn_bwd_coef_kernel<<<B, threads, smem, (cudaStream_t)stream>>>((const float2*)red_part, nchunks, (float2*)red,
                                                                   (const float2*)mean_rstd, gamma, beta, film, C, G, HW,
                                                                   (float4*)pqr, dgamma, dbeta, dfilm, slices);
    LAUNCH_CHECK("gn_bwd_coef_kernel");
    return S2S_OK;
}

int s2s_gn_bwd_apply(const void* x, const void* g, int ld_g, int B, int HW, int C, const float* coef, const float* pqr,
                     int Ctot, int c_off, const void* add, void* dx, int silu, float drop_p, uint64_t seed,
                     const void* mask_in, int x_fmt, int g_fmt, void* stream) {
    int rc = check_vec_layout(C, "gn_bwd_apply");
    if (rc) return rc;
    const int ppc = pick_pix_per_cta(B, HW, C);
    dim3 grid((HW + ppc - 1) / ppc, B);
    if (g_fmt != S2S_FMT_BF16) return fail(S2S_ERR_INVALID, "gn_bwd_apply: the gradient format must be bf16");
    constexpr int GF = kFmtBF16;
    // inputs streamed through shared memory by bulk copies (TMA unit): contiguous gradients of the large levels only (a channel
    // slice of the concat gradient needs one copy per pixel row -- measured slower than register loads --, and at <= 64^2 a
    // CTA lives for too few tiles to amortise the pipeline fill)
    if (gn_bulk() && vec_threads(C) == kEwThreads && ld_g == C && c_off == 0 && HW >= gn_bulk_min_hw()) {
        S2S_ACT(silu, SILU, S2S_DROP(drop_p, mask_in, DROP, S2S_BOOL(add != nullptr, ADD, S2S_FMT(x_fmt, XF, {
            auto kern = gn_bwd_apply_bulk_kernel<SILU, DROP, ADD, XF, GF>;
            const size_t smem = gn_bwd_apply_bulk_smem<ADD>();
            rc = set_smem(kern, smem);
            if (rc) return rc;
            kern<<<grid, kEwThreads, smem, (cudaStream_t)stream>>>(
                (const __nv_bfloat16*)x, (const __nv_bfloat16*)g, ld_g, C, HW, ppc, (const float2*)coef, (const float4*)pqr,
                Ctot, c_off, (const __nv_bfloat16*)add, (__nv_bfloat16*)dx, drop_p, seed, (const uint8_t*)mask_in);
        }))));
        LAUNCH_CHECK("gn_bwd_apply_bulk_kernel");
        return S2S_OK;
    }
    S2S_ACT(silu, SILU, S2S_DROP(drop_p, mask_in, DROP, S2S_BOOL(add != nullptr, ADD, S2S_FMT(x_fmt, XF,
        (gn_bwd_apply_kernel<SILU, DROP, ADD, XF, GF, kGnBwdApplyU><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)g, ld_g, C, HW, ppc, (const float2*)coef, (const float4*)pqr,
            Ctot, c_off, (const __nv_bfloat16*)add, (__nv_bfloat16*)dx, drop_p, seed, (const uint8_t*)mask_in))))));
    LAUNCH_CHECK("gn_bwd_apply_kernel");
    return S2S_OK;
}

int s2s_upsample2x(const void* in, void* out, int B, int H, int W, int C, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "upsample2x: C %% 8 != 0");
    const long long total = (long long)B * 4 * H * W * (C / 8);
    upsample2x_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, C / 8);
    LAUNCH_CHECK("upsample2x_kernel");
    return S2S_OK;
}
int s2s_sumpool2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "sumpool2x: C %% 8 != 0");
    const long long total = (long long)B * H * W * (C / 8);
    sumpool2x_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, C / 8, fmt);
    LAUNCH_CHECK("sumpool2x_kernel");
    return S2S_OK;
}
int s2s_zero_insert2x(const void* in, void* out, int B, int H, int W, int C, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "zero_insert2x: C %% 8 != 0");
    const long long total = (long long)B * 4 * H * W * (C / 8);
    zero_insert2x_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, C / 8);
    LAUNCH_CHECK("zero_insert2x_kernel");
    return S2S_OK;
}

int s2s_channel_sum(const void* x, long long npix, int C, float* out, int fmt, void* stream) {
    int rc = check_vec_layout(C, "channel_sum");
    if (rc) return rc;
    const int rows = kEwThreads / (C / 8);
    long long ctas = (long long)num_sms() * 8;
    long long ppc = (npix + ctas - 1) / ctas;
    ppc = (ppc + rows - 1) / rows * rows;
    if (ppc < rows * 8) ppc = rows * 8;
    const int grid = (int)((npix + ppc - 1) / ppc);
    channel_sum_kernel<<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, C, npix, (int)ppc, out, fmt);
    LAUNCH_CHECK("channel_sum_kernel");
    return S2S_OK;
}

int s2s_fm_loss(const float* v, const float* x0, const float* x1, long long n, float* loss, float* dv, void* stream) {
    int grid = ew_grid(n / 4 + 1);
    fm_loss_kernel<<<grid, kEwThreads, 0, (cudaStream_t)stream>>>(v, x0, x1, n, 1.0f / (float)n, loss, dv);
    LAUNCH_CHECK("fm_loss_kernel");
    return S2S_OK;
}

// ---- image-space kernels either side of the UNet (tiles.cuh) --------------------------------------------------------
int s2s_patch_pack(const float* x0, const float* x1, const float* t, const float* extra, int B, int Cx, int H, int W,
                   void* dst, int fmt, void* stream) {
    const int CT = Cx + (extra ? 1 : 0);
    if (!x0 || !dst || (x1 && !t) || Cx < 1) return fail(S2S_ERR_INVALID, "patch_pack: bad arguments");
    const long long npix = (long long)B * H * W;
    const int grid = ew_grid(npix, 128);
    cudaStream_t st = (cudaStream_t)stream;
    switch (CT) {
        case 3: patch_pack_kernel<3><<<grid, 128, 0, st>>>(x0, x1, t, extra, Cx, B, H, W, (__nv_bfloat16*)dst, fmt); break;
        case 4: patch_pack_kernel<4><<<grid, 128, 0, st>>>(x0, x1, t, extra, Cx, B, H, W, (__nv_bfloat16*)dst, fmt); break;
        default: return fail(S2S_ERR_INVALID, "patch_pack: %d input channels unsupported (3 or 4)", CT);
    }
    LAUNCH_CHECK("patch_pack_kernel");
    return S2S_OK;
}

int s2s_fm_loss_weighted(const float* v, const float* x0, const float* x1, const float* mask, float lam, int B, int C,
                         int HW, float* sums, float* dv, void* stream) {
    if (!v || !x0 || !x1 || !mask || !sums) return fail(S2S_ERR_INVALID, "fm_loss_weighted: bad arguments");
    fm_loss_weighted_kernel<<<ew_grid((long long)B * C * HW), kEwThreads, 0, (cudaStream_t)stream>>>(v, x0, x1, mask, lam, B,
                                                                                                  C, HW, sums, dv);
    LAUNCH_CHECK("fm_loss_weighted_kernel");
    return S2S_OK;
}

int s2s_roi_charbonnier(const float* x0, const float* x1, const float* t, const float* mask, int B, int C, int HW,
                        float eps, float* sums, void* stream) {
    if (!x0 || !x1 || !t || !mask || !sums) return fail(S2S_ERR_INVALID, "roi_charbonnier: bad arguments");
    roi_charbonnier_kernel<<<ew_grid((long long)B * HW), kEwThreads, 0, (cudaStream_t)stream>>>(x0, x1, t, mask, B, C, HW, eps,
                                                                                             sums);
    LAUNCH_CHECK("roi_charbonnier_kernel");
    return S2S_OK;
}

int s2s_tile_prep(const uint8_t* src, const uint8_t* tgt, const uint8_t* mask, const int* params, int B, int Hs, int Ws,
                  int S, int flags, float* out0, float* out1, float* outm, void* stream) {
    if (!src || !params || !out0 || (tgt && !out1) || (mask && !outm) || S <= 0 || S > Hs || S > Ws)
        return fail(S2S_ERR_INVALID, "tile_prep: bad arguments");
    tile_prep_kernel<<<ew_grid((long long)B * S * S), kEwThreads, 0, (cudaStream_t)stream>>>(src, tgt, mask, params, B, Hs, Ws,
                                                                                          S, flags, out0, out1, outm);
    LAUNCH_CHECK("tile_prep_kernel");
    return S2S_OK;
}

int s2s_resample_u8(const uint8_t* in, int B, int Hin, int Win, int C, const int* bounds, const int* kk, int ksize,
                    int n_out, int vertical, uint8_t* out, void* stream) {
    if (!in || !bounds || !kk || !out || ksize <= 0 || n_out <= 0) return fail(S2S_ERR_INVALID, "resample_u8: bad arguments");
    if (vertical)
        resample_u8_v_kernel<<<ew_grid((long long)B * n_out * Win * C), kEwThreads, 0, (cudaStream_t)stream>>>(
            in, B, Hin, Win, C, bounds, kk, ksize, n_out, out);
    else
        resample_u8_h_kernel<<<ew_grid((long long)B * Hin * n_out * C), kEwThreads, 0, (cudaStream_t)stream>>>(
            in, B, Hin, Win, C, bounds, kk, ksize, n_out, out);
    LAUNCH_CHECK("resample_u8_kernel");
    return S2S_OK;
}

int s2s_denorm_u8(const float* x, int B, int C, int HW, uint8_t* out, void* stream) {
    if (!x || !out) return fail(S2S_ERR_INVALID, "denorm_u8: bad arguments");
    denorm_u8_kernel<<<ew_grid((long long)B * HW), kEwThreads, 0, (cudaStream_t)stream>>>(x, B, C, HW, out);
    LAUNCH_CHECK("denorm_u8_kernel");
    return S2S_OK;
}

int s2s_convert16(const void* in, void* out, long long n, int in_fmt, int out_fmt, void* stream) {
    if (n % 8) return fail(S2S_ERR_INVALID, "convert16: n %% 8 != 0");
    convert16_kernel<<<ew_grid(n / 8), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, n / 8,
                                                                              in_fmt, out_fmt);
    LAUNCH_CHECK("convert16_kernel");
    return S2S_OK;
}

int s2s_nchw_f32_to_nhwc16(const float* in, void* out, int B, int C, int HW, int fmt, void* stream) {
    const long long total = (long long)B * C * HW;
    nchw_f32_to_nhwc_bf16_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>(in, (uint16_t*)out, B, C, HW, fmt);
    LAUNCH_CHECK("nchw_f32_to_nhwc_bf16_kernel");
    return S2S_OK;
}
int s2s_nhwc16_to_nchw_f32(const void* in, float* out, int B, int C, int HW, int fmt, void* stream) {
    const long long total = (long long)B * C * HW;
    nhwc_bf16_to_nchw_f32_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint16_t*)in, out, B, C, HW, fmt);
    LAUNCH_CHECK("nhwc_bf16_to_nchw_f32_kernel");
    return S2S_OK;
}

int s2s_bn_coef(const float* stats, int B, int nchunks, int C, int HW, const float* gamma, const float* beta, float eps,
                float momentum, float* running_mean, float* running_var, float* coef, float* mean_rstd, void* stream) {
    if (!stats || !gamma || !beta || !coef || !mean_rstd || C <= 0) return fail(S2S_ERR_INVALID, "bn_coef: bad arguments");
    bn_coef_kernel<<<(C + kBnCh - 1) / kBnCh, kBnCh * kBnRows, 0, (cudaStream_t)stream>>>(
        (const float2*)stats, B * nchunks, B, C, (long long)B * HW, gamma, beta, eps, momentum, running_mean, running_var,
        (float2*)coef, (float2*)mean_rstd);
    LAUNCH_CHECK("bn_coef_kernel");
    return S2S_OK;
}

int s2s_bn_bwd_coef(const float* red, int B, int nchunks, int C, int HW, const float* mean_rstd, const float* gamma,
                    float* pqr, float* dgamma, float* dbeta, void* stream) {
    if (!red || !mean_rstd || !gamma || !pqr || !dgamma || !dbeta) return fail(S2S_ERR_INVALID, "bn_bwd_coef: bad arguments");
    bn_bwd_coef_kernel<<<(C + kBnCh - 1) / kBnCh, kBnCh * kBnRows, 0, (cudaStream_t)stream>>>(
        (const float2*)red, B * nchunks, B, C, (long long)B * HW, (const float2*)mean_rstd, gamma, (float4*)pqr, dgamma,
        dbeta);
    LAUNCH_CHECK("bn_bwd_coef_kernel");
    return S2S_OK;
}

int s2s_bn_fold_slices(int nparts) {  // slices a fold of `nparts` partials is cut into (each at least 256 parts)
    int s = nparts / 256;
    return s < 1 ? 1 : (s > 128 ? 128 : s);
}

int s2s_bn_fold(const float* parts, int nparts, int C, float* sums, int nslices, void* stream) {
    if (!parts || !sums || nparts <= 0 || C <= 0 || nslices <= 0 || nslices > nparts)
        return fail(S2S_ERR_INVALID, "bn_fold: bad arguments");
    dim3 grid((C + kBnCh - 1) / kBnCh, nslices);
    bn_fold_kernel<<<grid, kBnCh * kBnRows, 0, (cudaStream_t)stream>>>((const float2*)parts, nparts, C, (float2*)sums);
    LAUNCH_CHECK("bn_fold_kernel");
    return S2S_OK;
}

int s2s_bn_coef_sums(const float* sums, int nparts, int C, long long count, int B, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var, float* coef, float* mean_rstd,
                     void* stream) {
    if (!sums || !gamma || !beta || !coef || !mean_rstd || C <= 0 || count <= 0 || B <= 0 || nparts <= 0)
        return fail(S2S_ERR_INVALID, "bn_coef_sums: bad arguments");
    bn_coef_kernel<<<(C + kBnCh - 1) / kBnCh, kBnCh * kBnRows, 0, (cudaStream_t)stream>>>(
        (const float2*)sums, nparts, B, C, count, gamma, beta, eps, momentum, running_mean, running_var, (float2*)coef,
        (float2*)mean_rstd);
    LAUNCH_CHECK("bn_coef_kernel");
    return S2S_OK;
}

int s2s_bn_bwd_coef_sums(const float* sums, int nparts, int C, long long count, int B, const float* mean_rstd,
                         const float* gamma, float* pqr, float* dgamma, float* dbeta, void* stream) {
    if (!sums || !mean_rstd || !gamma || !pqr || !dgamma || !dbeta || count <= 0 || B <= 0 || nparts <= 0)
        return fail(S2S_ERR_INVALID, "bn_bwd_coef_sums: bad arguments");
    bn_bwd_coef_kernel<<<(C + kBnCh - 1) / kBnCh, kBnCh * kBnRows, 0, (cudaStream_t)stream>>>(
        (const float2*)sums, nparts, B, C, count, (const float2*)mean_rstd, gamma, (float4*)pqr, dgamma, dbeta);
    LAUNCH_CHECK("bn_bwd_coef_kernel");
    return S2S_OK;
}

int s2s_maxpool2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "maxpool2x: C %% 8 != 0");
    const long long total = (long long)B * H * W * (C / 8);
    maxpool2x_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, C / 8, fmt);
    LAUNCH_CHECK("maxpool2x_kernel");
    return S2S_OK;
}
int s2s_maxpool2x_bwd(const void* x, const void* g, void* dx, int B, int H, int W, int C, int x_fmt, int g_fmt,
                      void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "maxpool2x_bwd: C %% 8 != 0");
    const long long total = (long long)B * H * W * (C / 8);
    maxpool2x_bwd_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)g, (uint4*)dx,
                                                                                   B, H, W, C / 8, x_fmt, g_fmt);
    LAUNCH_CHECK("maxpool2x_bwd_kernel");
    return S2S_OK;
}
int s2s_bilinear2x(const void* in, void* out, int B, int H, int W, int C, int fmt, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "bilinear2x: C %% 8 != 0");
    if (2 * H > 65535 || B > 65535) return fail(S2S_ERR_INVALID, "bilinear2x: H / B too large for the row grid");
    const int cols = 2 * W * (C / 8);
    const dim3 grid((cols + 255) / 256 > 8 ? 8 : (cols + 255) / 256, 2 * H, B);
    S2S_FMT(fmt, F, (bilinear2x_kernel<F><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, H, W, C / 8)));
    LAUNCH_CHECK("bilinear2x_kernel");
    return S2S_OK;
}
int s2s_bilinear2x_bwd(const void* g, void* din, int B, int H, int W, int C, int fmt, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "bilinear2x_bwd: C %% 8 != 0");
    if (H > 65535 || B > 65535) return fail(S2S_ERR_INVALID, "bilinear2x_bwd: H / B too large for the row grid");
    const int cols = W * (C / 8);
    const dim3 grid((cols + 255) / 256 > 8 ? 8 : (cols + 255) / 256, H, B);
    S2S_FMT(fmt, F, (bilinear2x_bwd_kernel<F><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)g, (uint4*)din, H, W, C / 8)));
    LAUNCH_CHECK("bilinear2x_bwd_kernel");
    return S2S_OK;
}
int s2s_nchw_f32_to_nhwc16_pad(const float* in, void* out, int B, int C, int Cpad, int HW, int fmt, void* stream) {
    if (C > Cpad || Cpad % 8) return fail(S2S_ERR_INVALID, "nchw_f32_to_nhwc16_pad: need C <= Cpad, Cpad %% 8 == 0");
    const long long total = (long long)B * (Cpad / 8) * HW;  // one thread per 16-byte vector
    nchw_f32_to_nhwc16_pad_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>(in, (uint16_t*)out, B, C, Cpad, HW, fmt);
    LAUNCH_CHECK("nchw_f32_to_nhwc16_pad_kernel");
    return S2S_OK;
}

int s2s_seg_loss_sums(const float* logits, const long long* target, int B, int C, int HW, long long ignore_index,
                      double* sums, void* stream) {
    if (C < 1 || C > kSegMaxC) return fail(S2S_ERR_INVALID, "seg_loss: C = %d (1..%d)", C, kSegMaxC);
    seg_loss_sums_kernel<<<ew_grid((long long)B * HW), 256, 0, (cudaStream_t)stream>>>(logits, target, B, C, HW, ignore_index, sums);
    LAUNCH_CHECK("seg_loss_sums_kernel");
    return S2S_OK;
}
int s2s_seg_loss_bwd(const float* logits, const long long* target, int B, int C, int HW, long long ignore_index,
                     const double* sums, float smooth, float w_dice, float w_ce, const float* gscale, float* dlogits,
                     void* stream) {
    if (C < 1 || C > kSegMaxC) return fail(S2S_ERR_INVALID, "seg_loss: C = %d (1..%d)", C, kSegMaxC);
    seg_loss_bwd_kernel<<<ew_grid((long long)B * HW), 256, 0, (cudaStream_t)stream>>>(
        logits, target, B, C, HW, ignore_index, sums, smooth, w_dice, w_ce, gscale, dlogits);
    LAUNCH_CHECK("seg_loss_bwd_kernel");
    return S2S_OK;
}

// ------------------------------------------------------------------------------------------------ attention core
static int attn_params(AttnParams* p, int B, int T, int heads, int ch, int new_order) {
    if (B < 1 || T < 1 || heads < 1 || (ch != 32 && ch != 64))
        return fail(S2S_ERR_INVALID, "attention: head channels = %d unsupported (32 or 64), B=%d T=%d heads=%d", ch, B, T, heads);
    memset(p, 0, sizeof(*p));
    p->B = B; p->T = T; p->heads = heads; p->C = heads * ch; p->ld = 3 * heads * ch;
    p->head_stride = new_order ? ch : 3 * ch;
    p->which_stride = new_order ? heads * ch : ch;
    p->scale = 1.0f / sqrtf((float)ch);
    p->scale_log2 = p->scale * 1.4426950408889634f;
    return S2S_OK;
}
int s2s_attn_supported(int ch) { return (ch == 32 || ch == 64) ? 1 : 0; }

int s2s_attn_fwd(const void* qkv, int B, int T, int heads, int ch, int new_order, void* out, float* lse, int a_fmt,
                 void* stream) {
    if (!qkv || !out) return fail(S2S_ERR_INVALID, "attn_fwd: null argument");
    AttnParams p;
    int rc = attn_params(&p, B, T, heads, ch, new_order);
    if (rc) return rc;
    p.qkv = (const uint16_t*)qkv; p.out = (uint16_t*)out; p.lse = lse;
    const dim3 grid((T + kAttnBlock - 1) / kAttnBlock, B * heads);
    cudaStream_t st = (cudaStream_t)stream;
    if (ch == 32) S2S_FMT(a_fmt, AF, (attn_fwd_kernel<32, AF><<<grid, kAttnThreads, 0, st>>>(p)));
    else S2S_FMT(a_fmt, AF, (attn_fwd_kernel<64, AF><<<grid, kAttnThreads, 0, st>>>(p)));
    LAUNCH_CHECK("attn_fwd_kernel");
    return S2S_OK;
}

int s2s_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, float* dvec, void* d_qkv, int B, int T,
                 int heads, int ch, int new_order, int a_fmt, int g_fmt, void* stream) {
    if (!qkv || !out || !d_out || !lse || !dvec || !d_qkv) return fail(S2S_ERR_INVALID, "attn_bwd: null argument");
    AttnParams p;
    int rc = attn_params(&p, B, T, heads, ch, new_order);
    if (rc) return rc;
    p.qkv = (const uint16_t*)qkv; p.o_fwd = (const uint16_t*)out; p.d_out = (const uint16_t*)d_out;
    p.lse = const_cast<float*>(lse); p.dvec = dvec; p.d_qkv = (uint16_t*)d_qkv;
    const dim3 grid((T + kAttnBlock - 1) / kAttnBlock, B * heads);
    cudaStream_t st = (cudaStream_t)stream;
    const int prep_grid = ew_grid((long long)B * heads * T);
#define S2S_ATTN_BWD(DV)                                                                                   \
    S2S_FMT(a_fmt, AF, S2S_FMT(g_fmt, GF, {                                                                \
        attn_bwd_prep_kernel<DV, AF, GF><<<prep_grid, kEwThreads, 0, st>>>(p);                             \
        attn_bwd_kv_kernel<DV, AF, GF><<<grid, kAttnThreads, 0, st>>>(p);                                  \
        attn_bwd_q_kernel<DV, AF, GF><<<grid, kAttnThreads, 0, st>>>(p);                                   \
    }))
    if (ch == 32) S2S_ATTN_BWD(32);
    else S2S_ATTN_BWD(64);
#undef S2S_ATTN_BWD
    LAUNCH_CHECK("attn_bwd kernels");
    return S2S_OK;
}

int s2s_head_conv(const void* a, int B, int H, int W, int C, const float* w_oihw, int Cout, const float* bias, float* out,
                  const float* axpy_x, float axpy_a, int a_fmt, void* stream) {
    if (!a || !w_oihw || !out) return fail(S2S_ERR_INVALID, "head_conv: null argument");
    if (C % 16 || C > 512 || Cout < 1 || Cout > 8)
        return fail(S2S_ERR_INVALID, "head_conv: needs C %% 16 == 0, C <= 512, Cout <= 8 (C=%d Cout=%d)", C, Cout);
    HeadConvParams p;
    memset(&p, 0, sizeof(p));
    p.a = (const uint16_t*)a; p.w = w_oihw; p.bias = bias; p.out = out; p.axpy_x = axpy_x; p.axpy_a = axpy_a;
    p.B = B; p.H = H; p.W = W; p.C = C; p.Cout = Cout;
    p.tiles_x = (W + kHeadTW - 1) / kHeadTW;
    p.tiles_y = (H + kHeadTH - 1) / kHeadTH;
    p.total_tiles = B * p.tiles_x * p.tiles_y;
    const size_t smem = (size_t)((kHeadTH + 2) * (kHeadTW + 2) + 72) * (C + 8) * 2;
    int per_sm = (int)(kSmemBudget / (smem + 1024));
    if (per_sm < 1) return fail(S2S_ERR_INVALID, "head_conv: tile does not fit in shared memory (C=%d)", C);
    if (per_sm > 4) per_sm = 4;
    int grid = num_sms() * per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    if (a_fmt == S2S_FMT_F16) {
        int rc = set_smem(head_conv_kernel<kFmtF16>, smem);
        if (rc) return rc;
        head_conv_kernel<kFmtF16><<<grid, kHeadThreads, smem, st>>>(p);
    } else {
        int rc = set_smem(head_conv_kernel<kFmtBF16>, smem);
        if (rc) return rc;
        head_conv_kernel<kFmtBF16><<<grid, kHeadThreads, smem, st>>>(p);
    }
    LAUNCH_CHECK("head_conv_kernel");
    return S2S_OK;
}

int s2s_linear_max_jobs(void) { return kLinMaxJobs; }

int s2s_linear_multi(const s2s_gemm_job* jobs, int njobs, void* stream) {
    static_assert(sizeof(s2s_gemm_job) == sizeof(GemmJob), "ABI struct drifted from the kernel's");
    if (njobs <= 0) return S2S_OK;
    if (!jobs) return fail(S2S_ERR_INVALID, "linear_multi: null job table");
    for (int first = 0; first < njobs; first += kLinMaxJobs) {
        GemmBatch b;
        memset(&b, 0, sizeof(b));
        b.njobs = njobs - first < kLinMaxJobs ? njobs - first : kLinMaxJobs;
        int tiles = 0;
        for (int j = 0; j < b.njobs; ++j) {
            const s2s_gemm_job& jb = jobs[first + j];
            if (!jb.A || !jb.B || !jb.C || jb.M <= 0 || jb.N <= 0 || jb.K <= 0)
                return fail(S2S_ERR_INVALID, "linear_multi: job %d has bad arguments", first + j);
            memcpy(&b.jobs[j], &jb, sizeof(GemmJob));
            tiles += ((jb.M + kLinTile - 1) / kLinTile) * ((jb.N + kLinTile - 1) / kLinTile);
            b.tile_end[j] = tiles;
        }
        linear_multi_kernel<<<tiles, kLinThreads, 0, (cudaStream_t)stream>>>(b);
        LAUNCH_CHECK("linear_multi_kernel");
    }
    return S2S_OK;
}

int s2s_sum_parts_silu_bwd(const float* parts, int nparts, long long n, const float* z, float* out, void* stream) {
    if (!parts || !out || nparts < 1 || n < 1) return fail(S2S_ERR_INVALID, "sum_parts_silu_bwd: bad arguments");
    sum_parts_silu_bwd_kernel<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(parts, nparts, n, z, out);
    LAUNCH_CHECK("sum_parts_silu_bwd_kernel");
    return S2S_OK;
}

int s2s_timestep_embedding(const float* t, int B, int dim, float max_period, float* emb, void* stream) {
    if (!t || !emb || B < 1 || dim < 2) return fail(S2S_ERR_INVALID, "timestep_embedding: bad arguments");
    timestep_embedding_kernel<<<ew_grid((long long)B * dim), kEwThreads, 0, (cudaStream_t)stream>>>(t, B, dim, max_period, emb);
    LAUNCH_CHECK("timestep_embedding_kernel");
    return S2S_OK;
}

int s2s_adam_chunk(void) { return kAdamChunk; }

int s2s_adam_multi(const s2s_adam_tensor* tensors_dev, const int* work_dev, int n_work, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int step, double grad_scale, void* stream) {
    return s2s_adam_multi_step(tensors_dev, work_dev, n_work, lr, beta1, beta2, eps, weight_decay, step, nullptr, grad_scale,
                               stream);
}

int s2s_copy_multi(const s2s_copy_tensor* tensors_dev, const int* work_dev, int n_work, void* stream) {
    static_assert(sizeof(s2s_copy_tensor) == sizeof(CopyTensor), "ABI struct drifted from the kernel's");
    if (n_work <= 0) return S2S_OK;
    if (!tensors_dev || !work_dev) return fail(S2S_ERR_INVALID, "copy_multi: bad arguments");
    copy_multi_kernel<<<n_work, kAdamThreads, 0, (cudaStream_t)stream>>>((const CopyTensor*)tensors_dev, (const int2*)work_dev);
    LAUNCH_CHECK("copy_multi_kernel");
    return S2S_OK;
}

int s2s_adam_multi_step(const s2s_adam_tensor* tensors_dev, const int* work_dev, int n_work, double lr, double beta1,
                        double beta2, double eps, double weight_decay, int step, const long long* step_dev,
                        double grad_scale, void* stream) {
    static_assert(sizeof(s2s_adam_tensor) == sizeof(AdamTensor), "ABI struct drifted from the kernel's");
    if (n_work <= 0) return S2S_OK;
    if (!tensors_dev || !work_dev || step < 1) return fail(S2S_ERR_INVALID, "adam_multi: bad arguments");
    AdamHyper h;
    h.lr = (float)lr; h.beta1 = (float)beta1; h.beta2 = (float)beta2; h.eps = (float)eps;
    h.weight_decay = (float)weight_decay;
    h.omb1 = (float)(1.0 - beta1);
    h.omb2 = (float)(1.0 - beta2);
    h.bias_corr1 = (float)(1.0 - pow(beta1, (double)step));
    h.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
    h.grad_scale = (float)grad_scale;
    adam_multi_kernel<<<n_work, kAdamThreads, 0, (cudaStream_t)stream>>>((const AdamTensor*)tensors_dev,
                                                                          (const int2*)work_dev, h, step_dev);
    LAUNCH_CHECK("adam_multi_kernel");
    return S2S_OK;
}

}  // extern "C"
