// K4: masked activation-guidance losses and their gradients (SURVEY.md 8(a) rows 10, 10b, 10c;
// losses.py:4-84 with patch_size = 1, evaluated by guided_stable_diffuser.py:417-434).
//
// One fused pass: for every (layer, channel) plane one CTA
//   1. stages the current and the recorded plane in shared memory (float4 loads),
//   2. bilinearly resizes both to the G x G loss grid when the layer is smaller (align_corners=False,
//      torch's upsample_bilinear2d arithmetic),
//   3. foreground term: sum_n |up(orig)[src_n] - up(cur)[dst_n]| with warp-shuffle reductions, and the
//      sign counts per destination cell accumulated as INTEGERS in shared memory - integer addition is
//      associative, so the gradient is bit-reproducible although the correspondence list has ~6x
//      duplicates (the reference's autograd index_put(accumulate=True) is not deterministic on CUDA),
//   4. background term (global_avg: |mean_bg_orig - mean_bg_trans|; local_avg: same as 3 over bg cells),
//   5. writes the weighted gradient at native resolution (transposed bilinear, gather form, no atomics).
// A second tiny kernel reduces the per-channel partial sums in a fixed order into the loss scalars.
#include "dh_common.cuh"

#include <string.h>

namespace dh {

constexpr int kLossThreads = 256;
constexpr int kMaxLossLayers = 8;
constexpr int kMaxG = 64;
constexpr int kMaxNative = 64;

struct LossLayerDev {
    const float* cur;
    const float* orig;
    float* grad;
    int C, h, w;
    float fgw, bgw;
    int chan_begin;     // first global channel id of this layer
};

struct LossParams {
    LossLayerDev lv[kMaxLossLayers];
    int n_layers, total_channels, G;
    const int32_t* fg_src; const int32_t* fg_dst; int n_fg;
    const int32_t* bg_orig; int n_bg_orig;
    const int32_t* bg_trans; int n_bg_trans;
    const int32_t* bg_common; int n_bg_common;
    int fg_kind;        // 0 = off, 1 = local_avg patch 1
    int bg_kind;        // 0 = off, 1 = global_avg, 2 = local_avg
    float* partial;     // [total_channels][2]: fg sum, bg term
};

__device__ __forceinline__ float block_sum_f(float v, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) sm[warp_id()] = v;
    __syncthreads();
    float t = lane_id() < (kLossThreads / 32) ? sm[lane_id()] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    __syncthreads();
    return t;
}

// torch area_pixel_compute_source_index (align_corners = False) for output index i
__device__ __forceinline__ void bilinear_tap(int i, int n_in, float scale, int& i0, int& i1, float& lam) {
    float src = scale * ((float)i + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = src - (float)i0;
}

__global__ void __launch_bounds__(kLossThreads) guidance_loss_kernel(const __grid_constant__ LossParams p) {
    extern __shared__ __align__(16) float smf[];
    __shared__ float red[32];
    const int G = p.G, GG = G * G;
    const int gc = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLossLayers; ++i)
        if (i < p.n_layers && gc >= p.lv[i].chan_begin) l = i;
    const LossLayerDev& L = p.lv[l];
    const int c = gc - L.chan_begin;
    const int h = L.h, w = L.w, hw = h * w;
    const bool resize = (h != G) || (w != G);
    const int tid = threadIdx.x;

    float* uc = smf;                   // up(cur)   G*G
    float* uo = smf + GG;              // up(orig)  G*G
    float* gu = smf + 2 * GG;          // dL/d up(cur), G*G (ints while counting)
    int* gi = reinterpret_cast<int*>(gu);
    float* nc = smf + 3 * GG;          // native planes (only when resizing)
    float* no = nc + hw;
    float* tmp = no + hw;              // h * G, separable transpose

    const float* cur = L.cur + (size_t)c * hw;
    const float* org = L.orig + (size_t)c * hw;
    {
        float* dc = resize ? nc : uc;
        float* dorg = resize ? no : uo;
        if ((hw & 3) == 0) {
            const float4* c4 = reinterpret_cast<const float4*>(cur);
            const float4* o4 = reinterpret_cast<const float4*>(org);
            for (int i = tid; i < hw / 4; i += kLossThreads) {
                reinterpret_cast<float4*>(dc)[i] = __ldg(c4 + i);
                reinterpret_cast<float4*>(dorg)[i] = __ldg(o4 + i);
            }
        } else {
            for (int i = tid; i < hw; i += kLossThreads) { dc[i] = cur[i]; dorg[i] = org[i]; }
        }
    }
    for (int i = tid; i < GG; i += kLossThreads) gi[i] = 0;
    __syncthreads();
    const float sy = (float)h / (float)G, sx = (float)w / (float)G;
    if (resize) {
        for (int i = tid; i < GG; i += kLossThreads) {
            const int r = i / G, s = i - r * G;
            int y0, y1, x0, x1;
            float ly, lx;
            bilinear_tap(r, h, sy, y0, y1, ly);
            bilinear_tap(s, w, sx, x0, x1, lx);
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            uc[i] = hy * (hx * nc[y0 * w + x0] + lx * nc[y0 * w + x1]) + ly * (hx * nc[y1 * w + x0] + lx * nc[y1 * w + x1]);
            uo[i] = hy * (hx * no[y0 * w + x0] + lx * no[y0 * w + x1]) + ly * (hx * no[y1 * w + x0] + lx * no[y1 * w + x1]);
        }
        __syncthreads();
    }

    // ---- foreground: local_average_feat_l1_loss with patch 1 (losses.py:80-82) ----
    float acc = 0.0f;
    for (int n = tid; p.fg_kind && n < p.n_fg; n += kLossThreads) {
        const int s = p.fg_src[n], d = p.fg_dst[n];
        const float df = uo[s] - uc[d];
        acc += fabsf(df);
        const int sg = (df > 0.0f) - (df < 0.0f);
        if (sg) atomicAdd(gi + d, -sg);
    }
    const float fg_sum = block_sum_f(acc, red);     // (contains the barrier that orders the atomics)
    const float fscale = p.fg_kind ? L.fgw / ((float)L.C * (float)p.n_fg) : 0.0f;
    for (int i = tid; i < GG; i += kLossThreads) gu[i] = p.fg_kind ? (float)gi[i] * fscale : 0.0f;
    __syncthreads();

    // ---- background ----
    float bg_term = 0.0f;
    if (p.bg_kind == 1) {       // average_feat_l1_loss (losses.py:46-49)
        float so = 0.0f, sc = 0.0f;
        for (int n = tid; n < p.n_bg_orig; n += kLossThreads) so += uo[p.bg_orig[n]];
        for (int n = tid; n < p.n_bg_trans; n += kLossThreads) sc += uc[p.bg_trans[n]];
        so = block_sum_f(so, red);
        sc = block_sum_f(sc, red);
        const float delta = so / (float)p.n_bg_orig - sc / (float)p.n_bg_trans;
        bg_term = fabsf(delta);
        const float sg = (float)((delta > 0.0f) - (delta < 0.0f));
        const float bscale = -sg * L.bgw / ((float)L.C * (float)p.n_bg_trans);
        // atomicAdd: a generic caller may list a cell twice; every addend is the same value, so the result does
        // not depend on the order
        for (int n = tid; n < p.n_bg_trans; n += kLossThreads) atomicAdd(gu + p.bg_trans[n], bscale);
    } else if (p.bg_kind == 2) {   // local_avg over the common background cells (losses.py:31-36)
        float a2 = 0.0f;
        const float bscale = L.bgw / ((float)L.C * (float)p.n_bg_common);
        for (int n = tid; n < p.n_bg_common; n += kLossThreads) {
            const int q = p.bg_common[n];
            const float df = uo[q] - uc[q];
            a2 += fabsf(df);
            atomicAdd(gu + q, -(float)((df > 0.0f) - (df < 0.0f)) * bscale);
        }
        bg_term = block_sum_f(a2, red);
    }
    __syncthreads();
    if (tid == 0) {
        p.partial[2 * gc] = fg_sum;
        p.partial[2 * gc + 1] = bg_term;
    }
    if (!L.grad) return;
    float* g = L.grad + (size_t)c * hw;
    if (!resize) {
        for (int i = tid; i < GG / 4; i += kLossThreads)
            reinterpret_cast<float4*>(g)[i] = reinterpret_cast<const float4*>(gu)[i];
        return;
    }
    // ---- transposed bilinear resize, separable gather: rows (G -> h) into tmp, then columns (G -> w) ----
    const int ky = (G + h - 1) / h + 2, kx = (G + w - 1) / w + 2;     // half-width of the candidate window
    for (int i = tid; i < h * G; i += kLossThreads) {
        const int yi = i / G, s = i - yi * G;
        const int rc = (int)(((float)yi + 0.5f) / sy);
        float a = 0.0f;
        for (int r = max(0, rc - ky); r <= min(G - 1, rc + ky); ++r) {
            int y0, y1;
            float ly;
            bilinear_tap(r, h, sy, y0, y1, ly);
            float wgt = 0.0f;
            if (y0 == yi) wgt += 1.0f - ly;
            if (y1 == yi) wgt += ly;
            if (wgt != 0.0f) a += wgt * gu[r * G + s];
        }
        tmp[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < hw; i += kLossThreads) {
        const int yi = i / w, xj = i - yi * w;
        const int sc0 = (int)(((float)xj + 0.5f) / sx);
        float a = 0.0f;
        for (int s = max(0, sc0 - kx); s <= min(G - 1, sc0 + kx); ++s) {
            int x0, x1;
            float lx;
            bilinear_tap(s, w, sx, x0, x1, lx);
            float wgt = 0.0f;
            if (x0 == xj) wgt += 1.0f - lx;
            if (x1 == xj) wgt += lx;
            if (wgt != 0.0f) a += wgt * tmp[yi * G + s];
        }
        g[i] = a;
    }
}

// Fixed-order reduction of the per-channel partials -> loss_out[0] = total, [1+2l] = fg_l, [2+2l] = bg_l.
__global__ void __launch_bounds__(256) guidance_loss_finalize_kernel(const __grid_constant__ LossParams p, float* __restrict__ loss_out) {
    __shared__ float red[32];
    __shared__ float total;
    if (threadIdx.x == 0) total = 0.0f;
    __syncthreads();
    for (int l = 0; l < p.n_layers; ++l) {
        const LossLayerDev& L = p.lv[l];
        float a = 0.0f, b = 0.0f;
        for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
            a += p.partial[2 * (L.chan_begin + c)];
            b += p.partial[2 * (L.chan_begin + c) + 1];
        }
        a = block_sum_f(a, red);
        b = block_sum_f(b, red);
        if (threadIdx.x == 0) {
            const float fg = p.fg_kind ? a / (float)p.n_fg / (float)L.C : 0.0f;
            const float bg = p.bg_kind == 2 ? b / (float)p.n_bg_common / (float)L.C : (p.bg_kind == 1 ? b / (float)L.C : 0.0f);
            loss_out[1 + 2 * l] = fg;
            loss_out[2 + 2 * l] = bg;
            if (p.fg_kind) total += L.fgw * fg;
            if (p.bg_kind) total += L.bgw * bg;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[0] = total;
}

__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ data, size_t n, const float* __restrict__ scale) {
    const float s = *scale;
    if (s == 1.0f) return;      // the common autograd.grad(loss, ...) case: nothing to do
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = n / 4;
    if (i < n4) {
        float4 v = reinterpret_cast<float4*>(data)[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(data)[i] = v;
    }
    if (i < (n & 3)) data[n4 * 4 + i] *= s;
}

}  // namespace dh

using namespace dh;

extern "C" {

size_t dh_guidance_loss_workspace_bytes(int n_layers, int max_channels) {
    if (n_layers < 1 || max_channels < 1) return 0;
    return sizeof(float) * 2 * (size_t)n_layers * max_channels;
}

int dh_guidance_loss(const dh_loss_layer* layers_host, int n_layers, int grid, const int32_t* fg_src, const int32_t* fg_dst,
                     int n_fg, const int32_t* bg_orig, int n_bg_orig, const int32_t* bg_trans, int n_bg_trans,
                     const int32_t* bg_common, int n_bg_common, int fg_kind, int bg_kind, float* loss_out, void* ws,
                     size_t ws_bytes, void* stream) {
    DH_REQUIRE(layers_host && n_layers >= 1 && n_layers <= kMaxLossLayers && loss_out && ws);
    DH_REQUIRE(grid >= 1 && grid <= kMaxG);
    DH_REQUIRE(n_fg >= 0 && n_bg_orig >= 0 && n_bg_trans >= 0 && n_bg_common >= 0);
    DH_REQUIRE((fg_src && fg_dst) || n_fg == 0 || fg_kind == 0);
    if (bg_kind != 0 && bg_kind != 1 && bg_kind != 2) return DH_ERR_INVALID_ARGUMENT;
    if (fg_kind != 0 && fg_kind != 1) return DH_ERR_INVALID_ARGUMENT;
    if (bg_kind == 1) DH_REQUIRE((bg_orig || n_bg_orig == 0) && (bg_trans || n_bg_trans == 0));
    if (bg_kind == 2) DH_REQUIRE(bg_common || n_bg_common == 0);
    LossParams p;
    memset(&p, 0, sizeof(p));
    int chan = 0, max_hw = 0;
    bool any_resize = false;
    for (int i = 0; i < n_layers; ++i) {
        const dh_loss_layer& s = layers_host[i];
        DH_REQUIRE(s.cur && s.orig && s.channels >= 1 && s.h >= 1 && s.w >= 1);
        if (s.h > kMaxNative || s.w > kMaxNative) return DH_ERR_UNSUPPORTED;
        if ((reinterpret_cast<uintptr_t>(s.cur) & 15) || (reinterpret_cast<uintptr_t>(s.orig) & 15) ||
            (s.grad && (reinterpret_cast<uintptr_t>(s.grad) & 15)))
            return DH_ERR_INVALID_ARGUMENT;
        LossLayerDev& L = p.lv[i];
        L.cur = s.cur; L.orig = s.orig; L.grad = s.grad;
        L.C = s.channels; L.h = s.h; L.w = s.w; L.fgw = s.fg_weight; L.bgw = s.bg_weight;
        L.chan_begin = chan;
        chan += s.channels;
        if (s.h != grid || s.w != grid) {
            any_resize = true;
            if (s.h * s.w > max_hw) max_hw = s.h * s.w;
        }
    }
    if (ws_bytes < sizeof(float) * 2 * (size_t)chan) return DH_ERR_WORKSPACE;
    p.n_layers = n_layers; p.total_channels = chan; p.G = grid;
    p.fg_src = fg_src; p.fg_dst = fg_dst; p.n_fg = n_fg;
    p.bg_orig = bg_orig; p.n_bg_orig = n_bg_orig; p.bg_trans = bg_trans; p.n_bg_trans = n_bg_trans;
    p.bg_common = bg_common; p.n_bg_common = n_bg_common; p.fg_kind = fg_kind; p.bg_kind = bg_kind;
    p.partial = static_cast<float*>(ws);
    size_t smem = sizeof(float) * 3 * (size_t)grid * grid;
    if (any_resize) smem += sizeof(float) * (2 * (size_t)max_hw + (size_t)kMaxNative * grid);
    DH_CUDA_CHECK(cudaFuncSetAttribute(guidance_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t st = as_stream(stream);
    guidance_loss_kernel<<<chan, kLossThreads, smem, st>>>(p);
    DH_LAUNCH_CHECK();
    guidance_loss_finalize_kernel<<<1, 256, 0, st>>>(p, loss_out);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_scale_inplace(float* data, size_t n, const float* scale, void* stream) {
    DH_REQUIRE(data && scale);
    if (n == 0) return DH_OK;
    if (reinterpret_cast<uintptr_t>(data) & 15) return DH_ERR_INVALID_ARGUMENT;
    const size_t threads = n / 4 > 3 ? n / 4 : 4;
    scale_inplace_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(data, n, scale);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
