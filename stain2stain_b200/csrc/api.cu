// C-ABI entry points (include/s2s_b200.h): argument validation, TMA tensor-map construction and kernel launches.
#include "../../include/s2s_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "conv_igemm.cuh"
#include "elementwise.cuh"
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

// stride-1 activation view {C, W, 1, H, B} with a custom pixel box (halo-tiled conv)
int make_act_tmap_box(CUtensorMap* m, const void* x, int B, int H, int W, int C, int box_w, int box_h) {
    if (C % 8) return fail(S2S_ERR_INVALID, "activation channels (%d) must be a multiple of 8", C);
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
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
int pick_pix_per_cta(int B, int HW, int /*C*/) {
    const long long wave = (long long)num_sms() * 12;
    long long a = B, b = wave;
    while (b) { long long t = a % b; a = b; b = t; }  // a = gcd(B, wave)
    const long long base = wave / a;                  // smallest chunk count with (B * chunks) % wave == 0
    long long k = (2 * wave + (long long)B * base - 1) / ((long long)B * base);  // aim at >= 2 * wave CTAs in total
    long long chunks = base * (k < 1 ? 1 : k);
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
    if (!w || !dst || Cout <= 0 || ci_count <= 0 || ci_begin < 0 || ci_begin + ci_count > Cin || (taps != 1 && taps != 9))
        return fail(S2S_ERR_INVALID, "pack_conv_weight: bad arguments");
    const long long total = (long long)Cout * ci_count * taps;
    pack_conv_weight_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        w, Cout, Cin, taps, ci_begin, ci_count, (uint16_t*)dst, ld_k, k_off, transpose_flip, fmt);
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

int s2s_conv_stat_tiles_for(const s2s_conv_src* srcs, int nsrc, int Hout, int Wout, int Cout) {
    int BN, mt, tx, ty;
    if (halo_eligible(srcs, nsrc, Cout)) {
        halo_geometry(Hout, Wout, Cout, &BN, &mt, &tx, &ty);
        return tx * ty * mt;
    }
    return s2s_conv_stat_tiles(Hout, Wout, Cout);
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
    if (stats_out && (!out_bf16 || axpy_x || s2s_conv_stat_tiles_for(srcs, nsrc, Hout, Wout, Cout) == 0))
        return fail(S2S_ERR_INVALID, "conv_fwd: epilogue statistics need a CTA-pair path (s2s_conv_stat_tiles_for() > 0)");
    if (nsrc < 1 || nsrc > kMaxSeg) return fail(S2S_ERR_INVALID, "conv_fwd: nsrc = %d (1..%d)", nsrc, kMaxSeg);
    if (a_fmt != w_fmt)
        return fail(S2S_ERR_INVALID, "conv_fwd: activations and weights must share one 16-bit format (tcgen05 kind::f16 "
                                     "rejects mixed fp16 x bf16 operands)");
    if ((out_bf16 != nullptr) == (out_f32 != nullptr))
        return fail(S2S_ERR_INVALID, "conv_fwd: exactly one of out_bf16 / out_f32 must be given");
    // halo-tiled CTA-pair kernel: stride-1 convs with at least one 3x3 segment
    {
        const bool ok = out_bf16 && !axpy_x && halo_eligible(srcs, nsrc, Cout);
        const bool any3 = ok;
        if (ok && any3) {
            Conv3Params q;
            memset(&q, 0, sizeof(q));
            const bool cols3 = conv_halo() == 3;
            int BN3, mt3, tx3, ty3;
            halo_geometry(Hout, Wout, Cout, &BN3, &mt3, &tx3, &ty3);
            q.cols3 = cols3 ? 1 : 0;
            q.nseg = nsrc;
            int kb3 = 0;
            for (int s = 0; s < nsrc; ++s) {
                const s2s_conv_src& sc = srcs[s];
                int rc = sc.taps == 9 ? make_act_tmap_box(&q.tmA[s], sc.x, B, Hout, Wout, sc.C, cols3 ? kHaloTW : kHaloPitch,
                                                          kHaloTH * mt3 + 2)
                                      : make_act_tmap_box(&q.tmA[s], sc.x, B, Hout, Wout, sc.C, kHaloTW, kHaloTH);
                if (rc) return rc;
                q.seg[s].taps = sc.taps;
                q.seg[s].cblocks = (sc.C + kBlockK - 1) / kBlockK;
                q.seg[s].stride = 1;
                q.seg[s].C = sc.C;
                q.seg_kb[s] = kb3;
                kb3 += sc.taps * q.seg[s].cblocks;
            }
            if (kb3 * kBlockK != Ktot)
                return fail(S2S_ERR_INVALID, "conv_fwd: Ktot = %d does not match the segments (%d)", Ktot, kb3 * kBlockK);
            {
                cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
                cuuint64_t str[1] = {(cuuint64_t)Ktot * 2};
                cuuint32_t box[2] = {kBlockK, (cuuint32_t)(BN3 / 2)};
                int rc = make_tmap(&q.tmW, w_packed, 2, dims, str, box);
                if (rc) return rc;
            }
            int rc = make_act_tmap_box(&q.tmOut, out_bf16, B, Hout, Wout, Cout, kHaloTW, kHaloTH);
            if (rc) return rc;
            q.B = B; q.Hout = Hout; q.Wout = Wout; q.Cout = Cout;
            q.tiles_x = tx3;
            q.tiles_y = ty3;
            q.stats = (float2*)stats_out;
            q.stat_tiles = tx3 * ty3 * mt3;
            if (norms) {
                q.prologue = 1;
                q.act = act;
                for (int s = 0; s < nsrc; ++s) {
                    q.seg_coef[s] = (const float2*)norms[s].coef;
                    q.seg_x[s] = (const uint16_t*)srcs[s].x;
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

int s2s_conv_wgrad(const void* dy, int Cm, const void* x, int Cq, int taps, int stride, int B, int Hout, int Wout,
                   float* dw, int ldn, int n_off, int dy_fmt, int x_fmt, void* stream) {
    if (taps != 1 && taps != 9) return fail(S2S_ERR_INVALID, "conv_wgrad: taps = %d", taps);
    if (dy_fmt != x_fmt)
        return fail(S2S_ERR_INVALID, "conv_wgrad: dy and x must share one 16-bit format (convert with s2s_convert16)");
    if (Cq % 64) return fail(S2S_ERR_INVALID, "conv_wgrad: input channels must be a multiple of 64 (got %d)", Cq);
    if (ldn % 4 || n_off % 4) return fail(S2S_ERR_INVALID, "conv_wgrad: ldn / n_off must be multiples of 4");
    if (Cm % 256 == 0 && Cq % 128 == 0 && wgrad_pairs()) {  // CTA-pair kernel: M = 256 output channels per MMA
        Wgrad2Params q;
        memset(&q, 0, sizeof(q));
        int rc = make_act_tmap(&q.tmP, dy, B, Hout, Wout, Cm, 1);
        if (rc) return rc;
        rc = make_act_tmap(&q.tmQ, x, B, Hout * stride, Wout * stride, Cq, stride);
        if (rc) return rc;
        q.taps = taps; q.stride = stride; q.Cq = Cq; q.Mtot = Cm; q.Ntot = Cq;
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
    int rc = make_act_tmap(&p.tmP, dy, B, Hout, Wout, Cm, 1);
    if (rc) return rc;
    rc = make_act_tmap(&p.tmQ, x, B, Hout * stride, Wout * stride, Cq, stride);
    if (rc) return rc;
    p.taps = taps; p.stride = stride; p.Cq = Cq; p.Mtot = Cm; p.Ntot = Cq;
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

int s2s_pack_chunk(void) { return kPackChunk; }

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
    const long long total = (long long)M * n_count * taps;
    unpack_wgrad_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>(dw, taps, M, ldn, n_off, n_count, grad,
                                                                                  Cin_total, n_begin, beta);
    LAUNCH_CHECK("unpack_wgrad_kernel");
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
    int rc = check_vec_layout(C, "gn_apply");
    if (rc) return rc;
    if (ld_out % 8 || c_off % 8) return fail(S2S_ERR_INVALID, "gn_apply: ld_out / c_off must be multiples of 8");
    const int ppc = pick_pix_per_cta(B, HW, C);
    dim3 grid((HW + ppc - 1) / ppc, B);
    S2S_ACT(silu, SILU, S2S_BOOL(drop_p > 0.f, DROP, S2S_BOOL(y2_bf16 != nullptr, DUAL, S2S_FMT(x_fmt, XF, S2S_FMT(y_fmt, YF,
        (gn_apply_kernel<SILU, DROP, XF, YF, DUAL><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, C, HW, ppc, (const float2*)coef, Ctot, c_off, (__nv_bfloat16*)y,
            (__nv_bfloat16*)y2_bf16, ld_out, drop_p, seed, (uint8_t*)mask_out)))))));
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
    S2S_ACT(silu, SILU, S2S_BOOL(drop_p > 0.f, DROP, S2S_FMT(x_fmt, XF, S2S_FMT(g_fmt, GF,
        (gn_bwd_reduce_kernel<SILU, DROP, XF, GF><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
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
    if (G > 64 || C % G) return fail(S2S_ERR_INVALID, "gn_bwd_coef: G = %d, C = %d unsupported", G, C);
    gn_bwd_coef_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const float2*)red_part, s2s_gn_chunks(B, HW), (float2*)red,
                                                            (const float2*)mean_rstd, gamma, beta, film, C, G, HW,
                                                            (float4*)pqr, dgamma, dbeta, dfilm);
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
    S2S_ACT(silu, SILU, S2S_BOOL(drop_p > 0.f, DROP, S2S_BOOL(add != nullptr, ADD, S2S_FMT(x_fmt, XF, S2S_FMT(g_fmt, GF,
        (gn_bwd_apply_kernel<SILU, DROP, ADD, XF, GF><<<grid, vec_threads(C), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)g, ld_g, C, HW, ppc, (const float2*)coef, (const float4*)pqr,
            Ctot, c_off, (const __nv_bfloat16*)add, (__nv_bfloat16*)dx, drop_p, seed, (const uint8_t*)mask_in)))))));
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
    const long long total = (long long)B * 4 * H * W * (C / 8);
    bilinear2x_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, B, H, W, C / 8, fmt);
    LAUNCH_CHECK("bilinear2x_kernel");
    return S2S_OK;
}
int s2s_bilinear2x_bwd(const void* g, void* din, int B, int H, int W, int C, int fmt, void* stream) {
    if (C % 8) return fail(S2S_ERR_INVALID, "bilinear2x_bwd: C %% 8 != 0");
    const long long total = (long long)B * H * W * (C / 8);
    bilinear2x_bwd_kernel<<<ew_grid(total), kEwThreads, 0, (cudaStream_t)stream>>>((const uint4*)g, (uint4*)din, B, H, W, C / 8, fmt);
    LAUNCH_CHECK("bilinear2x_bwd_kernel");
    return S2S_OK;
}
int s2s_nchw_f32_to_nhwc16_pad(const float* in, void* out, int B, int C, int Cpad, int HW, int fmt, void* stream) {
    if (C > Cpad || Cpad % 8) return fail(S2S_ERR_INVALID, "nchw_f32_to_nhwc16_pad: need C <= Cpad, Cpad %% 8 == 0");
    const long long total = (long long)B * Cpad * HW;
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

int s2s_adam_chunk(void) { return kAdamChunk; }

int s2s_adam_multi(const s2s_adam_tensor* tensors_dev, const int* work_dev, int n_work, double lr, double beta1, double beta2,
                   double eps, double weight_decay, int step, double grad_scale, void* stream) {
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
                                                                          (const int2*)work_dev, h);
    LAUNCH_CHECK("adam_multi_kernel");
    return S2S_OK;
}

}  // extern "C"
