// K4: masked activation-guidance losses and their gradients (SURVEY.md 8(a) rows 10, 10b, 10c;
// losses.py:4-84 with patch_size = 1, evaluated by guided_stable_diffuser.py:417-434).
//
// Two parts.
//
// (1) dh_build_loss_plan - once per edit.  The correspondence list has ~6x duplicates at the 64x64 loss grid
//     (SURVEY.md "hard parts"); the plan groups it by DESTINATION cell (counting sort) and collapses equal
//     (src, dst) cell pairs into one entry with a multiplicity, giving a CSR over destination cells
//     (row_ptr, 8-byte pairs: src | dst << 16, multiplicity) plus per-cell multiplicities of the three background lists.
//
// (2) dh_guidance_loss - every denoising step.  Two kernels of small persistent CTAs (256 threads, three per SM) pull
//     (layer, channel) planes from a dynamic queue and evaluate, per plane,
//         L_fg  = 1/(C N)  sum_pairs mult * |up(orig)[src] - up(cur)[dst]|
//         dL/dup(cur)[dst] = -1/(C N) sum_pairs mult * sign(...)
//     plus the background term.  The sign terms are accumulated as INTEGERS per destination cell (shared-memory
//     atomics): integer addition is associative, so the gradient is bit-reproducible (the reference's
//     index_put(accumulate=True) backward is not, on CUDA).  Layers smaller than the loss grid are resized
//     bilinearly inside the box of pair cells only, and their gradient is written at native resolution through
//     the transposed resize in gather form.  Loss value and gradient come out of the same pass: algorithmic
//     traffic = read cur + read orig + write grad.  The last CTA to finish reduces the per-channel partial sums in
//     a fixed order; the second kernel is a programmatic dependent launch so that it fills the SMs as the first drains.
#include "dh_common.cuh"
#include "dh_loss_plan.cuh"

#include <string.h>

namespace dh {

// ------------------------------------------------------------------------------------------------
// plan builder: one CTA of 1024 threads
// ------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;
constexpr int kPlanCountWarps = 8;

__global__ void __launch_bounds__(kPlanThreads) loss_plan_kernel(
    const int32_t* __restrict__ fg_src, const int32_t* __restrict__ fg_dst, int n_fg,
    const int32_t* __restrict__ bg_orig, int n_bg_orig, const int32_t* __restrict__ bg_trans, int n_bg_trans,
    const int32_t* __restrict__ bg_common, int n_bg_common, int grid, void* plan, int32_t* __restrict__ scratch) {
    extern __shared__ __align__(16) int psm[];
    __shared__ int scan_smem[33];
    const int cells = grid * grid;
    int* hist = psm;                       // cells
    int* start = psm + cells;              // cells + 1
    int* ucount = start + cells + 1;       // cells
    uint32_t* counters = reinterpret_cast<uint32_t*>(ucount + cells);   // kPlanCountWarps * cells
    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    PlanView pv = plan_view(plan, grid, n_fg);

    for (int i = tid; i < cells; i += kPlanThreads) { hist[i] = 0; ucount[i] = 0; }
    __syncthreads();
    for (int n = tid; n < n_fg; n += kPlanThreads) atomicAdd(hist + fg_dst[n], 1);
    __syncthreads();
    // exclusive scan of hist -> start (each thread owns a contiguous run of cells)
    const int per = (cells + kPlanThreads - 1) / kPlanThreads;
    const int c0 = min(cells, tid * per), c1 = min(cells, c0 + per);
    int s = 0;
    for (int i = c0; i < c1; ++i) s += hist[i];
    int total;
    int run = block_exclusive_scan(s, scan_smem, total);
    for (int i = c0; i < c1; ++i) { start[i] = run; run += hist[i]; }
    if (tid == 0) start[cells] = total;
    __syncthreads();
    for (int i = tid; i < cells; i += kPlanThreads) hist[i] = start[i];        // cursors
    __syncthreads();
    for (int n = tid; n < n_fg; n += kPlanThreads) scratch[atomicAdd(hist + fg_dst[n], 1)] = fg_src[n];
    __syncthreads();
    // per destination cell: count the sources (shared-memory counters, one private table per warp), then emit the
    // distinct sources in ascending order in place of the bucket
    if (wid < kPlanCountWarps) {
        uint32_t* cnt = counters + (size_t)wid * cells;
        for (int d = wid; d < cells; d += kPlanCountWarps) {
            const int b0 = start[d], b1 = start[d + 1];
            if (b0 == b1) continue;
            int lo = 0x7FFFFFFF, hi = -1;
            for (int k = b0 + lane; k < b1; k += 32) { const int v = scratch[k]; lo = min(lo, v); hi = max(hi, v); }
            lo = __reduce_min_sync(0xFFFFFFFFu, lo);
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            for (int i = lo + lane; i <= hi; i += 32) cnt[i] = 0;
            __syncwarp();
            for (int k = b0 + lane; k < b1; k += 32) atomicAdd(cnt + scratch[k], 1u);
            __syncwarp();
            int out = b0;
            for (int base = lo; base <= hi; base += 32) {
                const int i = base + lane;
                uint32_t c = i <= hi ? cnt[i] : 0u;
                while (__any_sync(0xFFFFFFFFu, c > 0)) {
                    const unsigned b = __ballot_sync(0xFFFFFFFFu, c > 0);
                    const uint32_t m = c > 65535u ? 65535u : c;
                    if (c > 0) scratch[out + __popc(b & ((1u << lane) - 1u))] = (int32_t)((uint32_t)i | (m << 16));
                    c -= m;
                    out += __popc(b);
                }
            }
            if (lane == 0) ucount[d] = out - b0;
            __syncwarp();
        }
    }
    __syncthreads();
    // row_ptr = exclusive scan of the distinct-pair counts, then compact the buckets
    s = 0;
    for (int i = c0; i < c1; ++i) s += ucount[i];
    run = block_exclusive_scan(s, scan_smem, total);
    for (int i = c0; i < c1; ++i) {
        pv.row_ptr[i] = run;
        const int b0 = start[i];
        for (int j = 0; j < ucount[i]; ++j) {
            const uint32_t e = (uint32_t)scratch[b0 + j];
            pv.pairs[run + j] = make_uint2((e & 0xFFFFu) | ((uint32_t)i << 16), e >> 16);
        }
        run += ucount[i];
    }
    if (tid == 0) {
        pv.row_ptr[cells] = total;
        PlanHeader h;
        h.n_pairs = total; h.n_fg = n_fg; h.n_bg_orig = n_bg_orig; h.n_bg_trans = n_bg_trans; h.n_bg_common = n_bg_common;
        h.grid = grid; h.cap = n_fg; h.reserved = 0;
        *pv.hdr = h;
    }
    __syncthreads();
    // background list multiplicities per cell (lists from np.nonzero never repeat a cell; generic callers may)
    int* co = psm;
    int* ct = psm + cells;
    int* cc = psm + 2 * cells;
    int* is_src = psm + 3 * cells;
    for (int i = tid; i < 4 * cells; i += kPlanThreads) psm[i] = 0;
    __syncthreads();
    for (int n = tid; n < n_fg; n += kPlanThreads) is_src[fg_src[n]] = 1;
    __shared__ int box_sm[4];
    __shared__ int nonbin_sm;
    if (tid == 0) { box_sm[0] = grid; box_sm[1] = -1; box_sm[2] = grid; box_sm[3] = -1; nonbin_sm = 0; }
    for (int n = tid; n < n_bg_orig; n += kPlanThreads) atomicAdd(co + bg_orig[n], 1);
    for (int n = tid; n < n_bg_trans; n += kPlanThreads) atomicAdd(ct + bg_trans[n], 1);
    for (int n = tid; n < n_bg_common; n += kPlanThreads) atomicAdd(cc + bg_common[n], 1);
    __syncthreads();
    for (int i = tid; i < cells; i += kPlanThreads) {
        ushort4 v;
        v.x = (unsigned short)min(co[i], 65535); v.y = (unsigned short)min(ct[i], 65535);
        v.z = (unsigned short)min(cc[i], 65535); v.w = (unsigned short)(is_src[i] ? 1 : 0);
        pv.bgcnt[i] = v;
        if (co[i] > 1 || ct[i] > 1 || cc[i] > 1) nonbin_sm = 1;
        if (is_src[i] || pv.row_ptr[i + 1] > pv.row_ptr[i]) {
            const int r = i / grid, c = i - r * grid;
            atomicMin(box_sm + 0, r); atomicMax(box_sm + 1, r); atomicMin(box_sm + 2, c); atomicMax(box_sm + 3, c);
        }
    }
    __syncthreads();
    if (tid == 0) {
        pv.hdr->box_r0 = box_sm[0]; pv.hdr->box_r1 = box_sm[1]; pv.hdr->box_s0 = box_sm[2]; pv.hdr->box_s1 = box_sm[3];
        pv.hdr->reserved = nonbin_sm ? 0 : 1;      // bit 0: every background multiplicity is 0 or 1
    }
}

// ------------------------------------------------------------------------------------------------
// fused loss + gradient
// ------------------------------------------------------------------------------------------------
// Two kernels share one design: small CTAs (256 threads), many per SM, each walking over (layer, channel) planes.
//   loss_flat_kernel    layers that already have the loss-grid resolution (64x64): the planes are read straight from
//                       global memory with 128-bit loads (every cell is needed exactly once for the background sums
//                       and the gradient), the pair gathers hit the same 32 KB in L1.
//   loss_resize_kernel  smaller layers (32x32, ...): both planes are staged in shared memory, resized bilinearly
//                       only inside the box that contains pair cells; the background sums are inner products with
//                       up^T(multiplicity) at native resolution; the gradient is gathered back through up^T.
// The sign terms of the foreground pairs are accumulated as INTEGERS (shared-memory atomics): integer addition is
// associative, so the gradient is bit-reproducible (the reference's index_put(accumulate=True) is not, on CUDA).
// The last CTA of the last kernel reduces the per-channel partial sums in a fixed order.
constexpr int kLossThreads = 256;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kOwnGroups = kMaxG * kMaxG / 4 / kLossThreads;   // float4 groups per thread at the 64x64 grid = 4
constexpr int kPairUnroll = 4;   // independent pair chains per thread and iteration in the flat kernel (6 and 8 measured equal)
constexpr int kWin = 16;      // up rows (columns) that can touch one native row (column): 2 * G / h <= 16 for h >= 8

struct LossLayerDev {
    const float* cur;
    const float* orig;
    float* grad;
    int C, h, w;
    float fgw, bgw;
    const void* tab;      // LayerTab of a resized layer (NULL for layers at the loss-grid resolution)
    int chan_begin;       // first channel id of this layer inside the kernel's own channel range
    int partial_begin;    // first channel id of this layer in the partial-sum array (all layers)
};

struct LossParams {
    LossLayerDev lv[kMaxLossLayers];
    int n_layers, total_channels, G;
    const void* plan;
    int plan_cap;
    int n_fg, n_bg_orig, n_bg_trans, n_bg_common;
    int fg_kind;        // 0 = off, 1 = local_avg patch 1
    int bg_kind;        // 0 = off, 1 = global_avg, 2 = local_avg
    float* partial;     // [all channels][2]: fg sum, bg term
    unsigned int* work_counter;   // zeroed before the launches: channels beyond the first gridDim.x are handed out dynamically
};

struct LossFinal {
    int n_layers, C[kMaxLossLayers], partial_begin[kMaxLossLayers];
    float fgw[kMaxLossLayers], bgw[kMaxLossLayers];
    unsigned int* done_counter;       // zeroed before the launches
    unsigned int total_ctas;          // CTAs of both kernels
    float* loss_out;
};

struct ResizeLayout {   // float offsets into dynamic shared memory
    int tab, wo, wt, planes, uc, uo, cnt, tmp, total;
    int box_cap;        // capacity (cells) of the box-local uc / uo / cnt arrays
};

__device__ __forceinline__ int layer_of(const LossParams& p, int gc) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLossLayers; ++i)
        if (i < p.n_layers && gc >= p.lv[i].chan_begin) l = i;
    return l;
}

// sum over the CTA of three values, result in every thread; fixed order (warp tree, then a tree over the warp
// partials that every warp evaluates identically) -> deterministic.  One barrier.
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float (*red)[4]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    }
    if (lane_id() == 0) { red[warp_id()][0] = a; red[warp_id()][1] = b; red[warp_id()][2] = c; }
    __syncthreads();
    const float4 r = *reinterpret_cast<const float4*>(red[lane_id() & (kLossWarps - 1)]);
    a = r.x; b = r.y; c = r.z;
#pragma unroll
    for (int o = kLossWarps / 2; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    }
}

__device__ __forceinline__ float block_sum1(float v, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) sm[warp_id()] = v;
    __syncthreads();
    float t = lane_id() < kLossWarps ? sm[lane_id()] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    __syncthreads();
    return t;
}

// Called by every CTA of both kernels when it is done; the last one reduces the per-channel partials in a fixed
// order -> loss_out[0] = total, [1+2l] = fg_l, [2+2l] = bg_l.
__device__ void loss_finish(const LossParams& p, const LossFinal& f, float* red32) {
    __shared__ unsigned int ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(f.done_counter, 1u);
    __syncthreads();
    if (ticket != f.total_ctas - 1) return;
    __threadfence();
    float total = 0.0f;
    for (int l = 0; l < f.n_layers; ++l) {
        float a = 0.0f, b = 0.0f;
        for (int c = threadIdx.x; c < f.C[l]; c += blockDim.x) {
            a += __ldcg(p.partial + 2 * (f.partial_begin[l] + c));
            b += __ldcg(p.partial + 2 * (f.partial_begin[l] + c) + 1);
        }
        a = block_sum1(a, red32);
        b = block_sum1(b, red32);
        const float fg = p.fg_kind ? a / (float)p.n_fg / (float)f.C[l] : 0.0f;
        const float bg = p.bg_kind == 2 ? b / (float)p.n_bg_common / (float)f.C[l] : (p.bg_kind == 1 ? b / (float)f.C[l] : 0.0f);
        if (threadIdx.x == 0) { f.loss_out[1 + 2 * l] = fg; f.loss_out[2 + 2 * l] = bg; }
        if (p.fg_kind) total += f.fgw[l] * fg;
        if (p.bg_kind) total += f.bgw[l] * bg;
    }
    if (threadIdx.x == 0) f.loss_out[0] = total;
}

__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


// ---- layers at the loss-grid resolution -------------------------------------------------------------
template <bool kBinary>
__global__ void __launch_bounds__(kLossThreads, 3) loss_flat_kernel(const __grid_constant__ LossParams p,
                                                                    const __grid_constant__ LossFinal fin) {
    __shared__ __align__(16) int cnt[kMaxG * kMaxG];
    __shared__ __align__(16) float red[kLossWarps][4];
    __shared__ float red32[32];
    __shared__ float lconst[kMaxLossLayers][4];
    const int tid = threadIdx.x;
    const int GG = p.G * p.G;
    const PlanView pv = plan_view(const_cast<void*>(p.plan), p.G, p.plan_cap);
    const int n_pairs = pv.hdr->n_pairs;
    const uint2* __restrict__ pairs = pv.pairs;
    // membership of the own cells in the three background lists as 16-bit masks (bit 4k+i = cell i of group k).
    // Lists that come from np.nonzero never repeat a cell; if a generic caller does, the multiplicities are re-read
    // from the plan in the (slower) general path.
    uint32_t mo = 0, mt = 0, mc = 0;
#pragma unroll
    for (int k = 0; k < kOwnGroups; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = 4 * (tid + k * kLossThreads) + i;
            if (q < GG) {
                const ushort4 bc = pv.bgcnt[q];
                mo |= (bc.x ? 1u : 0u) << (4 * k + i); mt |= (bc.y ? 1u : 0u) << (4 * k + i); mc |= (bc.z ? 1u : 0u) << (4 * k + i);
            }
        }
    auto wgt = [&](uint32_t mask, int k, int i, int which) -> float {     // multiplicity of own cell (k, i) in list `which`
        if (kBinary) return (float)((mask >> (4 * k + i)) & 1u);
        const ushort4 bc = pv.bgcnt[4 * (tid + k * kLossThreads) + i];
        return (float)(which == 0 ? bc.x : which == 1 ? bc.y : bc.z);
    };
    if (tid < p.n_layers) {
        const LossLayerDev& L = p.lv[tid];
        // an empty index list makes the reference's loss NaN (mean over nothing) but its gradient ZERO: scale 0, not inf
        lconst[tid][0] = p.fg_kind && p.n_fg > 0 ? L.fgw / ((float)L.C * (float)p.n_fg) : 0.0f;
        lconst[tid][1] = p.bg_kind == 2 && p.n_bg_common > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_common) : 0.0f;
        lconst[tid][2] = p.bg_kind == 1 && p.n_bg_trans > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_trans) : 0.0f;
    }
    const float inv_no = 1.0f / (float)p.n_bg_orig, inv_nt = 1.0f / (float)p.n_bg_trans;
    __syncthreads();

    // dynamic channel queue: the two kernels of one evaluation overlap (programmatic dependent launch), so CTAs start
    // at different times; the next channel index is fetched while the current channel is processed
    __shared__ int s_next;
    int gc = blockIdx.x;
    while (gc < p.total_channels) {
        int nxt = 0;
        if (tid == 0) nxt = (int)atomicAdd(p.work_counter, 1u) + (int)gridDim.x;
        const int l = layer_of(p, gc);
        const LossLayerDev& L = p.lv[l];
        const int c = gc - L.chan_begin;
        const float* __restrict__ cur = L.cur + (size_t)c * GG;
        const float* __restrict__ org = L.orig + (size_t)c * GG;
        // Own cells: four 128-bit groups per thread.  They are consumed right away - background sums, and for the local
        // background term the sign of (orig - cur) as two bits per cell - so that no plane data stays in registers across
        // the pair walk (the loads also bring both planes into L1 for the pair gathers).
        float s1 = 0.0f, s2 = 0.0f;
        uint32_t sign_pos = 0u, sign_neg = 0u;          // bit 4k+i: orig > cur / orig < cur at own cell i of group k
#pragma unroll
        for (int k = 0; k < kOwnGroups; ++k) {
            const int q = 4 * (tid + k * kLossThreads);
            if (q < GG) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cur + q));
                const float4 o4 = __ldg(reinterpret_cast<const float4*>(org + q));
                *reinterpret_cast<int4*>(cnt + q) = make_int4(0, 0, 0, 0);
                const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
                if (p.bg_kind == 1) {
                    // same association as before: s = fma(w0, v0, fma(w1, v1, fma(w2, v2, fma(w3, v3, s))))
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        s1 = fmaf(wgt(mo, k, i, 0), ov[i], s1);
                        s2 = fmaf(wgt(mt, k, i, 1), cv[i], s2);
                    }
                } else if (p.bg_kind == 2) {
#pragma unroll
                    for (int i = 3; i >= 0; --i) {
                        const float d = ov[i] - cv[i];
                        s1 = fmaf(wgt(mc, k, i, 2), fabsf(d), s1);
                        sign_pos |= (d > 0.0f ? 1u : 0u) << (4 * k + i);
                        sign_neg |= (d < 0.0f ? 1u : 0u) << (4 * k + i);
                    }
                }
            }
        }
        __syncthreads();
        float acc_f = 0.0f;
        if (p.fg_kind) {
            // kPairUnroll independent pair chains per thread and iteration (memory-level parallelism for the L1 gathers)
            for (int j0 = tid; j0 < n_pairs; j0 += kPairUnroll * kLossThreads) {
                uint2 e[kPairUnroll];
                float df[kPairUnroll];
#pragma unroll
                for (int u = 0; u < kPairUnroll; ++u) {
                    const int j = j0 + u * kLossThreads;
                    e[u] = j < n_pairs ? pairs[j] : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < kPairUnroll; ++u) df[u] = __ldg(org + (e[u].x & 0xFFFFu)) - __ldg(cur + (e[u].x >> 16));
#pragma unroll
                for (int u = 0; u < kPairUnroll; ++u) {
                    acc_f = fmaf((float)e[u].y, fabsf(df[u]), acc_f);
                    if (df[u] != 0.0f && e[u].y) atomicAdd(cnt + (e[u].x >> 16), df[u] > 0.0f ? -(int)e[u].y : (int)e[u].y);
                }
            }
        }
        block_sum3(acc_f, s1, s2, red);       // (its barrier also orders the cnt atomics)
        const float fscale = lconst[l][0], lscale = lconst[l][1], gscale = lconst[l][2];
        float bg_term = 0.0f, bscale = 0.0f;
        if (p.bg_kind == 1) {
            const float delta = s1 * inv_no - s2 * inv_nt;
            bg_term = fabsf(delta);
            bscale = -sgn(delta) * gscale;
        } else if (p.bg_kind == 2) {
            bg_term = s1;
        }
        if (tid == 0) { p.partial[2 * (L.partial_begin + c)] = acc_f; p.partial[2 * (L.partial_begin + c) + 1] = bg_term; }
        if (L.grad) {
            float* g = L.grad + (size_t)c * GG;
#pragma unroll
            for (int k = 0; k < kOwnGroups; ++k) {
                const int q = 4 * (tid + k * kLossThreads);
                if (q >= GG) break;
                const int4 ci = *reinterpret_cast<const int4*>(cnt + q);
                float v[4] = {(float)ci.x * fscale, (float)ci.y * fscale, (float)ci.z * fscale, (float)ci.w * fscale};
                if (p.bg_kind == 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = fmaf(wgt(mt, k, i, 1), bscale, v[i]);
                } else if (p.bg_kind == 2) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float sg = (float)((sign_pos >> (4 * k + i)) & 1u) - (float)((sign_neg >> (4 * k + i)) & 1u);
                        v[i] -= sg * wgt(mc, k, i, 2) * lscale;
                    }
                }
                st_cs_f4(g + q, make_float4(v[0], v[1], v[2], v[3]));
            }
        }
        if (tid == 0) s_next = nxt;
        __syncthreads();     // cnt / red are reused by the next channel
        gc = s_next;
    }
    loss_finish(p, fin, red32);
}

// ---- layers smaller than the loss grid -----------------------------------------------------------------
// Per-layer tables, built once per evaluation by loss_resize_setup_kernel (one CTA per resized layer) in global
// memory and copied to shared memory by every CTA of loss_resize_kernel:
//   bilinear taps of every up row / column, the up rows (columns) that touch each native row (column) with their
//   weights (the transposed resize in gather form), up^T(background multiplicities) at native resolution, and the
//   native box that the active up box touches.
struct LayerTabSmall {
    int ty0[kMaxG], ty1[kMaxG], tx0[kMaxG], tx1[kMaxG];
    float tly[kMaxG], tlx[kMaxG];
    int ylo[kMaxNative], yhi[kMaxNative], xlo[kMaxNative], xhi[kMaxNative];
    float wrow[kMaxNative * kWin], wcol[kMaxNative * kWin];
    int box[8];                        // up space: r0, r1, s0, s1; native: y0, y1, x0, x1
};
struct LayerTab : LayerTabSmall {
    float wo[kMaxNative * kMaxNative], wt[kMaxNative * kMaxNative];
};
static_assert(sizeof(LayerTabSmall) % 16 == 0 && sizeof(LayerTab) % 16 == 0, "tables are copied with 128-bit accesses");

struct SetupParams {
    const void* plan;
    int plan_cap, G, h, w, fg_kind, bg_kind;
};

__global__ void __launch_bounds__(256) loss_resize_setup_kernel(const __grid_constant__ SetupParams p, LayerTab* __restrict__ tabs) {
    __shared__ float tmp[kMaxNative * kMaxG];
    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    LayerTab& T = tabs[0];
    const int G = p.G, h = p.h, w = p.w;
    const PlanView pv = plan_view(const_cast<void*>(p.plan), G, p.plan_cap);
    const float sy = (float)h / (float)G, sx = (float)w / (float)G;
    if (tid < G) {
        bilinear_tap(tid, h, sy, T.ty0[tid], T.ty1[tid], T.tly[tid]);
        bilinear_tap(tid, w, sx, T.tx0[tid], T.tx1[tid], T.tlx[tid]);
    }
    __syncthreads();
    if (tid < h) {       // up rows that touch native row tid, and their weights
        // (taps with weight zero - the clamped rows at the top border have i1 = 1 with lambda = 0 - are not part of the window)
        int lo = G, hi = -1;
        for (int r = 0; r < G; ++r)
            if ((T.ty0[r] == tid && T.tly[r] < 1.0f) || (T.ty1[r] == tid && T.tly[r] > 0.0f)) { lo = min(lo, r); hi = max(hi, r); }
        hi = min(hi, lo + kWin - 1);
        T.ylo[tid] = lo; T.yhi[tid] = hi;
        for (int r = lo; r <= hi; ++r)
            T.wrow[tid * kWin + r - lo] = (T.ty0[r] == tid ? 1.0f - T.tly[r] : 0.0f) + (T.ty1[r] == tid ? T.tly[r] : 0.0f);
    }
    if (tid >= 64 && tid - 64 < w) {
        const int j = tid - 64;
        int lo = G, hi = -1;
        for (int s = 0; s < G; ++s)
            if ((T.tx0[s] == j && T.tlx[s] < 1.0f) || (T.tx1[s] == j && T.tlx[s] > 0.0f)) { lo = min(lo, s); hi = max(hi, s); }
        hi = min(hi, lo + kWin - 1);
        T.xlo[j] = lo; T.xhi[j] = hi;
        for (int s = lo; s <= hi; ++s)
            T.wcol[j * kWin + s - lo] = (T.tx0[s] == j ? 1.0f - T.tlx[s] : 0.0f) + (T.tx1[s] == j ? T.tlx[s] : 0.0f);
    }
    if (tid == 128) {
        const bool full = p.bg_kind == 2;      // local_avg needs up() on every background cell
        const bool none = !p.fg_kind && !full;
        const int r0 = full ? 0 : (none ? G : pv.hdr->box_r0), r1 = full ? G - 1 : (none ? -1 : pv.hdr->box_r1);
        const int s0 = full ? 0 : (none ? G : pv.hdr->box_s0), s1 = full ? G - 1 : (none ? -1 : pv.hdr->box_s1);
        const bool any = r1 >= r0;
        T.box[0] = r0; T.box[1] = r1; T.box[2] = s0; T.box[3] = s1;
        T.box[4] = any ? T.ty0[r0] : 0; T.box[5] = any ? T.ty1[r1] : -1;
        T.box[6] = any ? T.tx0[s0] : 0; T.box[7] = any ? T.tx1[s1] : -1;
    }
    __syncthreads();
    // wo / wt = up^T applied to the background multiplicities (separable, via tmp)
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? T.wo : T.wt;
        for (int yi = wid; yi < h; yi += 8)
            for (int s = lane; s < G; s += 32) {
                float a = 0.0f;
                for (int r = T.ylo[yi]; r <= T.yhi[yi]; ++r) {
                    const ushort4 bc = pv.bgcnt[r * G + s];
                    a = fmaf(T.wrow[yi * kWin + r - T.ylo[yi]], (float)(pass == 0 ? bc.x : bc.y), a);
                }
                tmp[yi * G + s] = a;
            }
        __syncthreads();
        for (int yi = wid; yi < h; yi += 8)
            for (int xj = lane; xj < w; xj += 32) {
                float a = 0.0f;
                for (int s = T.xlo[xj]; s <= T.xhi[xj]; ++s) a = fmaf(T.wcol[xj * kWin + s - T.xlo[xj]], tmp[yi * G + s], a);
                dst[yi * w + xj] = a;
            }
        __syncthreads();
    }
}

struct ResizeShared {
    float red[kLossWarps][4];
    float red32[32];
    float lconst[kMaxLossLayers][4];
};

// kG = 64: the loss grid of the reference (shifts instead of integer divisions); kG = 0: any grid <= 64.
template <int kG>
__global__ void __launch_bounds__(kLossThreads, 3) loss_resize_kernel(const __grid_constant__ LossParams p,
                                                                      const __grid_constant__ LossFinal fin,
                                                                      const __grid_constant__ ResizeLayout lay) {
    extern __shared__ __align__(16) float rsm[];
    __shared__ __align__(16) ResizeShared sh;
    // let the flat-layer kernel of the same evaluation (a programmatic dependent launch) start as soon as SM resources free up
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    const int G = kG ? kG : p.G, GG = G * G;
    const PlanView pv = plan_view(const_cast<void*>(p.plan), G, p.plan_cap);
    const int n_pairs = pv.hdr->n_pairs;
    const uint2* __restrict__ pairs = pv.pairs;
    LayerTabSmall& T = *reinterpret_cast<LayerTabSmall*>(rsm + lay.tab);
    float* const two = rsm + lay.wo;           // up^T(background multiplicities) of the current layer
    float* const twt = rsm + lay.wt;
    float* const planes = rsm + lay.planes;
    float* const suc = rsm + lay.uc;          // box-local: index (r - br0) * bw + (s - bs0)
    float* const suo = rsm + lay.uo;
    int* const cnt = reinterpret_cast<int*>(rsm + lay.cnt);
    float* const gu = rsm + lay.cnt;
    float* const tmp = rsm + lay.tmp;

    if (tid < p.n_layers) {
        const LossLayerDev& L = p.lv[tid];
        // an empty index list makes the reference's loss NaN (mean over nothing) but its gradient ZERO: scale 0, not inf
        sh.lconst[tid][0] = p.fg_kind && p.n_fg > 0 ? L.fgw / ((float)L.C * (float)p.n_fg) : 0.0f;
        sh.lconst[tid][1] = p.bg_kind == 2 && p.n_bg_common > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_common) : 0.0f;
        sh.lconst[tid][2] = p.bg_kind == 1 && p.n_bg_trans > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_trans) : 0.0f;
    }
    const float inv_no = 1.0f / (float)p.n_bg_orig, inv_nt = 1.0f / (float)p.n_bg_trans;
    // the active box is a property of the plan (identical in every layer's table)
    const LayerTab* tab0 = static_cast<const LayerTab*>(p.lv[0].tab);
    const int br0 = tab0->box[0], br1 = tab0->box[1], bs0 = tab0->box[2], bs1 = tab0->box[3];
    const bool any_box = br1 >= br0;
    const int bw = any_box ? bs1 - bs0 + 1 : 0, bh = any_box ? br1 - br0 + 1 : 0;
    const int bcells = bw * bh;
    __syncthreads();
    if (bcells > lay.box_cap) {      // the caller under-sized the box-local buffers: poison the result instead of corrupting memory
        for (int gc = blockIdx.x; gc < p.total_channels; gc += gridDim.x)
            if (tid == 0) {
                const int l = layer_of(p, gc);
                p.partial[2 * (p.lv[l].partial_begin + gc - p.lv[l].chan_begin)] = __int_as_float(0x7FC00000);
                p.partial[2 * (p.lv[l].partial_begin + gc - p.lv[l].chan_begin) + 1] = __int_as_float(0x7FC00000);
            }
        loss_finish(p, fin, sh.red32);
        return;
    }

    int cur_layer = -1;
    __shared__ int s_next;
    int gc = blockIdx.x;
    while (gc < p.total_channels) {
        int nxt = 0;
        if (tid == 0) nxt = (int)atomicAdd(p.work_counter, 1u) + (int)gridDim.x;
        const int l = layer_of(p, gc);
        const LossLayerDev& L = p.lv[l];
        const int c = gc - L.chan_begin;
        const int h = L.h, w = L.w, hw = h * w;
        float* const pc_ = planes;
        float* const po_ = planes + hw;
        {   // stage both planes (coalesced 128-bit loads)
            const float4* c4 = reinterpret_cast<const float4*>(L.cur + (size_t)c * hw);
            const float4* o4 = reinterpret_cast<const float4*>(L.orig + (size_t)c * hw);
            for (int i = tid; i < hw / 4; i += kLossThreads) {
                reinterpret_cast<float4*>(pc_)[i] = __ldg(c4 + i);
                reinterpret_cast<float4*>(po_)[i] = __ldg(o4 + i);
            }
            for (int i = tid; i < bcells; i += kLossThreads) cnt[i] = 0;
        }
        if (l != cur_layer) {       // (a CTA crosses a layer boundary at most n_layers times)
            const LayerTab* tl = static_cast<const LayerTab*>(L.tab);
            const float4* src = reinterpret_cast<const float4*>(static_cast<const LayerTabSmall*>(tl));
            float4* dst = reinterpret_cast<float4*>(&T);
            for (int i = tid; i < (int)(sizeof(LayerTabSmall) / 16); i += kLossThreads) dst[i] = src[i];
            if (p.bg_kind == 1)
                for (int i = tid; i < hw / 4; i += kLossThreads) {
                    reinterpret_cast<float4*>(two)[i] = reinterpret_cast<const float4*>(tl->wo)[i];
                    reinterpret_cast<float4*>(twt)[i] = reinterpret_cast<const float4*>(tl->wt)[i];
                }
            cur_layer = l;
        }
        __syncthreads();          // planes staged, cnt zeroed, tables loaded
        const int ny0 = T.box[4], ny1 = T.box[5], nx0 = T.box[6], nx1 = T.box[7];
        // up(cur), up(orig) inside the box
        for (int r = br0 + wid; r <= br1; r += kLossWarps) {
            const int y0 = T.ty0[r] * w, y1 = T.ty1[r] * w;
            const float ly = T.tly[r], hy = 1.0f - ly;
            for (int s = bs0 + lane; s <= bs1; s += 32) {
                const int x0 = T.tx0[s], x1 = T.tx1[s];
                const float lx = T.tlx[s], hx = 1.0f - lx;
                const int b = (r - br0) * bw + (s - bs0);
                suc[b] = hy * (hx * pc_[y0 + x0] + lx * pc_[y0 + x1]) + ly * (hx * pc_[y1 + x0] + lx * pc_[y1 + x1]);
                suo[b] = hy * (hx * po_[y0 + x0] + lx * po_[y0 + x1]) + ly * (hx * po_[y1 + x0] + lx * po_[y1 + x1]);
            }
        }
        __syncthreads();
        float acc_f = 0.0f, s1s = 0.0f, s2s = 0.0f;
        if (p.fg_kind) {
            const int boff = br0 * bw + bs0;
            for (int j = tid; j < n_pairs; j += kLossThreads) {
                const uint2 e = pairs[j];
                const int d = (int)(e.x >> 16), sc_ = (int)(e.x & 0xFFFFu);
                const int dr = kG ? d >> 6 : d / G, sr = kG ? sc_ >> 6 : sc_ / G;
                const int db = dr * bw + (d - dr * G) - boff, sb = sr * bw + (sc_ - sr * G) - boff;
                const float df = suo[sb] - suc[db];
                acc_f = fmaf((float)e.y, fabsf(df), acc_f);
                if (df != 0.0f) atomicAdd(cnt + db, df > 0.0f ? -(int)e.y : (int)e.y);
            }
        }
        const float fscale = sh.lconst[l][0], lscale = sh.lconst[l][1], gscale = sh.lconst[l][2];
        if (p.bg_kind == 1) {     // background sums at native resolution: <wo, orig>, <wt, cur>
            for (int i = tid * 4; i < hw; i += kLossThreads * 4) {
                const float4 a = *reinterpret_cast<const float4*>(two + i), b = *reinterpret_cast<const float4*>(po_ + i);
                const float4 e = *reinterpret_cast<const float4*>(twt + i), f = *reinterpret_cast<const float4*>(pc_ + i);
                s1s = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, s1s))));
                s2s = fmaf(e.x, f.x, fmaf(e.y, f.y, fmaf(e.z, f.z, fmaf(e.w, f.w, s2s))));
            }
        } else if (p.bg_kind == 2) {       // box = whole grid
            for (int q = tid; q < GG; q += kLossThreads) s1s = fmaf((float)pv.bgcnt[q].z, fabsf(suo[q] - suc[q]), s1s);
        }
        block_sum3(acc_f, s1s, s2s, sh.red);
        float bg_term = 0.0f, bscale = 0.0f;
        if (p.bg_kind == 1) {
            const float delta = s1s * inv_no - s2s * inv_nt;
            bg_term = fabsf(delta);
            bscale = -sgn(delta) * gscale;
        } else if (p.bg_kind == 2) {
            bg_term = s1s;
        }
        if (tid == 0) { p.partial[2 * (L.partial_begin + c)] = acc_f; p.partial[2 * (L.partial_begin + c) + 1] = bg_term; }
        if (L.grad) {
            float* g = L.grad + (size_t)c * hw;
            // gradient w.r.t. up(cur) inside the box (in place: integer -> float), then up^T in gather form
            for (int i = tid; i < bcells; i += kLossThreads) {
                float v = (float)cnt[i] * fscale;
                if (p.bg_kind == 2) v -= sgn(suo[i] - suc[i]) * (float)pv.bgcnt[i].z * lscale;
                gu[i] = v;
            }
            __syncthreads();
            for (int yi = ny0 + wid; yi <= ny1; yi += kLossWarps) {
                const int lo = T.ylo[yi];
                const int ra = max(lo, br0), rb = min(T.yhi[yi], br1);
                const float* wr = T.wrow + yi * kWin - lo;
                for (int s = bs0 + lane; s <= bs1; s += 32) {
                    float a = 0.0f;
                    const float* gp = gu + (s - bs0) - br0 * bw;
                    for (int r = ra; r <= rb; ++r) a = fmaf(wr[r], gp[r * bw], a);
                    tmp[yi * G + s] = a;
                }
            }
            __syncthreads();
            for (int yi = wid; yi < h; yi += kLossWarps) {
                const bool row_in = yi >= ny0 && yi <= ny1;
                for (int xj = lane; xj < w; xj += 32) {
                    float a = 0.0f;
                    if (row_in && xj >= nx0 && xj <= nx1) {
                        const int lo = T.xlo[xj];
                        const int sa = max(lo, bs0), sb = min(T.xhi[xj], bs1);
                        const float* wc = T.wcol + xj * kWin - lo;
                        const float* tp = tmp + yi * G;
                        for (int s = sa; s <= sb; ++s) a = fmaf(wc[s], tp[s], a);
                    }
                    if (p.bg_kind == 1) a = fmaf(bscale, twt[yi * w + xj], a);
                    g[yi * w + xj] = a;
                }
            }
        }
        if (tid == 0) s_next = nxt;
        __syncthreads();     // planes / cnt / uc / tmp are reused by the next channel
        gc = s_next;
    }
    loss_finish(p, fin, sh.red32);
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

struct ScaleMany {
    float* data[kMaxLossLayers];
    unsigned long long n[kMaxLossLayers];
    unsigned int block_begin[kMaxLossLayers + 1];
    int count;
};

// several tensors in one launch (autograd's backward of the fused multi-layer loss)
__global__ void __launch_bounds__(256) scale_many_kernel(const __grid_constant__ ScaleMany m, const float* __restrict__ scale) {
    const float s = *scale;
    if (s == 1.0f) return;
    int t = 0;
#pragma unroll
    for (int i = 1; i < kMaxLossLayers; ++i)
        if (i < m.count && blockIdx.x >= m.block_begin[i]) t = i;
    float* data = m.data[t];
    const size_t n = m.n[t], n4 = n / 4;
    const size_t i = (size_t)(blockIdx.x - m.block_begin[t]) * blockDim.x + threadIdx.x;
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

size_t dh_loss_plan_bytes(int grid, int n_fg) {
    if (grid < 1 || grid > kMaxG || n_fg < 0) return 0;
    size_t a, b, c;
    return plan_layout(grid, n_fg, &a, &b, &c);
}

size_t dh_loss_plan_workspace_bytes(int grid, int n_fg) {
    if (grid < 1 || grid > kMaxG || n_fg < 0) return 0;
    return sizeof(int32_t) * (size_t)(n_fg > 0 ? n_fg : 1);
}

int dh_build_loss_plan(const int32_t* fg_src, const int32_t* fg_dst, int n_fg, const int32_t* bg_orig, int n_bg_orig,
                       const int32_t* bg_trans, int n_bg_trans, const int32_t* bg_common, int n_bg_common, int grid,
                       void* plan, size_t plan_bytes, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(plan && ws && grid >= 1 && grid <= kMaxG);
    DH_REQUIRE(n_fg >= 0 && n_bg_orig >= 0 && n_bg_trans >= 0 && n_bg_common >= 0);
    DH_REQUIRE((fg_src && fg_dst) || n_fg == 0);
    DH_REQUIRE((bg_orig || n_bg_orig == 0) && (bg_trans || n_bg_trans == 0) && (bg_common || n_bg_common == 0));
    if (plan_bytes < dh_loss_plan_bytes(grid, n_fg) || ws_bytes < dh_loss_plan_workspace_bytes(grid, n_fg)) return DH_ERR_WORKSPACE;
    const int cells = grid * grid;
    const size_t smem = sizeof(int) * ((size_t)3 * cells + 1) + sizeof(uint32_t) * (size_t)kPlanCountWarps * cells;
    DH_CUDA_CHECK(cudaFuncSetAttribute(loss_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loss_plan_kernel<<<1, kPlanThreads, smem, as_stream(stream)>>>(fg_src, fg_dst, n_fg, bg_orig, n_bg_orig, bg_trans, n_bg_trans,
                                                                   bg_common, n_bg_common, grid, plan, static_cast<int32_t*>(ws));
    DH_LAUNCH_CHECK();
    return DH_OK;
}

static size_t loss_ws_layout(size_t channels, size_t* o_counter) {
    size_t o = (sizeof(float) * 2 * channels + 15) / 16 * 16;
    *o_counter = o; o += 16;
    return o;
}

size_t dh_guidance_loss_workspace_bytes(int n_layers, int max_channels) {
    if (n_layers < 1 || max_channels < 1) return 0;
    size_t a;
    return loss_ws_layout((size_t)n_layers * max_channels, &a);
}

size_t dh_loss_resize_tables_bytes(void) { return sizeof(LayerTab); }

int dh_build_loss_resize_tables(const void* plan, int n_fg, int grid, int h, int w, int fg_kind, int bg_kind, void* tables,
                                void* stream) {
    DH_REQUIRE(plan && tables && grid >= 1 && grid <= kMaxG && h >= 1 && w >= 1 && n_fg >= 0);
    if (h > grid || w > grid || h > kMaxNative || w > kMaxNative) return DH_ERR_UNSUPPORTED;
    if (2 * ((grid + h - 1) / h) > kWin || 2 * ((grid + w - 1) / w) > kWin) return DH_ERR_UNSUPPORTED;
    if (2 * ((grid + h - 1) / h) > kWin || 2 * ((grid + w - 1) / w) > kWin) return DH_ERR_UNSUPPORTED;
    SetupParams sp;
    sp.plan = plan; sp.plan_cap = n_fg; sp.G = grid; sp.h = h; sp.w = w; sp.fg_kind = fg_kind; sp.bg_kind = bg_kind;
    loss_resize_setup_kernel<<<1, 256, 0, as_stream(stream)>>>(sp, static_cast<LayerTab*>(tables));
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_loss_plan_info(const void* plan_header_host, int* n_pairs, int* box_cells, int* plan_flags) {
    DH_REQUIRE(plan_header_host);
    const PlanHeader* h = static_cast<const PlanHeader*>(plan_header_host);
    if (n_pairs) *n_pairs = h->n_pairs;
    if (plan_flags) *plan_flags = h->reserved & 1;
    if (box_cells) *box_cells = h->box_r1 >= h->box_r0 ? (h->box_r1 - h->box_r0 + 1) * (h->box_s1 - h->box_s0 + 1) : 0;
    return DH_OK;
}

int dh_guidance_loss(const dh_loss_layer* layers_host, int n_layers, int grid, const void* plan, int n_fg, int n_bg_orig,
                     int n_bg_trans, int n_bg_common, int box_cells, int plan_flags, int fg_kind, int bg_kind, float* loss_out,
                     void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(layers_host && n_layers >= 1 && n_layers <= kMaxLossLayers && loss_out && ws && plan);
    DH_REQUIRE(grid >= 1 && grid <= kMaxG);
    DH_REQUIRE(n_fg >= 0 && n_bg_orig >= 0 && n_bg_trans >= 0 && n_bg_common >= 0);
    if (bg_kind != 0 && bg_kind != 1 && bg_kind != 2) return DH_ERR_INVALID_ARGUMENT;
    if (fg_kind != 0 && fg_kind != 1) return DH_ERR_INVALID_ARGUMENT;
    LossParams pf, pr;      // flat (native == grid) and resized layers
    LossFinal fin;
    memset(&pf, 0, sizeof(pf));
    memset(&pr, 0, sizeof(pr));
    memset(&fin, 0, sizeof(fin));
    int chan = 0, hw_r = 0, h_r = 0;
    for (int i = 0; i < n_layers; ++i) {
        const dh_loss_layer& s = layers_host[i];
        DH_REQUIRE(s.cur && s.orig && s.channels >= 1 && s.h >= 1 && s.w >= 1);
        if (s.h > kMaxNative || s.w > kMaxNative || s.h > grid || s.w > grid) return DH_ERR_UNSUPPORTED;
        // the transposed-resize tables hold kWin up rows (columns) per native row (column): 2 * ceil(grid / h) of them are needed
        if (2 * ((grid + s.h - 1) / s.h) > kWin || 2 * ((grid + s.w - 1) / s.w) > kWin) return DH_ERR_UNSUPPORTED;
        if (((size_t)s.h * s.w) % 4 != 0) return DH_ERR_UNSUPPORTED;     // 128-bit plane loads
        if ((reinterpret_cast<uintptr_t>(s.cur) & 15) || (reinterpret_cast<uintptr_t>(s.orig) & 15) ||
            (s.grad && (reinterpret_cast<uintptr_t>(s.grad) & 15)))
            return DH_ERR_INVALID_ARGUMENT;
        const bool flat = s.h == grid && s.w == grid;
        if (flat && (grid * grid) % 4 != 0) return DH_ERR_UNSUPPORTED;
        if (!flat && (!s.resize_tables || (reinterpret_cast<uintptr_t>(s.resize_tables) & 15))) return DH_ERR_INVALID_ARGUMENT;
        LossParams& q = flat ? pf : pr;
        LossLayerDev& L = q.lv[q.n_layers++];
        L.cur = s.cur; L.orig = s.orig; L.grad = s.grad;
        L.C = s.channels; L.h = s.h; L.w = s.w; L.fgw = s.fg_weight; L.bgw = s.bg_weight;
        L.tab = flat ? nullptr : s.resize_tables;
        L.chan_begin = q.total_channels;
        L.partial_begin = chan;
        q.total_channels += s.channels;
        fin.C[i] = s.channels; fin.partial_begin[i] = chan; fin.fgw[i] = s.fg_weight; fin.bgw[i] = s.bg_weight;
        chan += s.channels;
        if (!flat) {
            if (s.h * s.w > hw_r) hw_r = s.h * s.w;
            if (s.h > h_r) h_r = s.h;
        }
    }
    fin.n_layers = n_layers;
    size_t o_counter;
    if (ws_bytes < loss_ws_layout((size_t)chan, &o_counter)) return DH_ERR_WORKSPACE;
    for (LossParams* q : {&pf, &pr}) {
        q->G = grid; q->plan = plan; q->plan_cap = n_fg;
        q->n_fg = n_fg; q->n_bg_orig = n_bg_orig; q->n_bg_trans = n_bg_trans; q->n_bg_common = n_bg_common;
        q->fg_kind = fg_kind; q->bg_kind = bg_kind;
        q->partial = static_cast<float*>(ws);
    }
    fin.done_counter = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + o_counter);
    pf.work_counter = fin.done_counter + 1;
    pr.work_counter = fin.done_counter + 2;
    fin.loss_out = loss_out;
    cudaStream_t st = as_stream(stream);
    // per-process caches of the launch geometry queries (this entry point runs every denoising step)
    struct LaunchCache {
        int sms, occ_flat[2], occ_resize[2];
        size_t occ_resize_smem[2], smem_attr_set[2];
    };
    static LaunchCache caches[64];          // zero-initialised; one entry per device (function attributes are per device)
    int dev = 0;
    DH_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DH_ERR_UNSUPPORTED;
    LaunchCache& lc = caches[dev];
    if (!lc.sms) DH_CUDA_CHECK(cudaDeviceGetAttribute(&lc.sms, cudaDevAttrMultiProcessorCount, dev));
    const int sms = lc.sms;
    int (&occ_flat)[2] = lc.occ_flat;
    int (&occ_resize)[2] = lc.occ_resize;
    size_t (&occ_resize_smem)[2] = lc.occ_resize_smem;
    size_t (&smem_attr_set)[2] = lc.smem_attr_set;
    // grids: persistent CTAs, as many as fit per SM
    int grid_f = 0, grid_r = 0;
    ResizeLayout lay;
    memset(&lay, 0, sizeof(lay));
    size_t smem_r = 0;
    // bit 0 of plan_flags: background multiplicities are all 0/1 (lists from np.nonzero) -> register bit masks
    auto flat_kernel = (plan_flags & 1) ? loss_flat_kernel<true> : loss_flat_kernel<false>;
    if (pf.total_channels) {
        int& per_sm = occ_flat[(plan_flags & 1) ? 1 : 0];
        if (!per_sm) {
            DH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flat_kernel, kLossThreads, 0));
            if (per_sm < 1) per_sm = 1;
        }
        grid_f = sms * per_sm;
        if (grid_f > pf.total_channels) grid_f = pf.total_channels;
    }
    auto resize_kernel = grid == 64 ? loss_resize_kernel<64> : loss_resize_kernel<0>;
    if (pr.total_channels) {
        const int GG = grid * grid;
        auto up4 = [](int v) { return (v + 3) / 4 * 4; };
        int cap = (box_cells > 0 && box_cells <= GG && bg_kind != 2) ? box_cells : GG;
        if (!fg_kind && bg_kind != 2) cap = 4;
        int o = 0;
        lay.tab = o;    o += (int)(sizeof(LayerTabSmall) / 4);
        lay.wo = o;     o += up4(hw_r);
        lay.wt = o;     o += up4(hw_r);
        lay.planes = o; o += up4(2 * hw_r);
        lay.uc = o;     o += up4(cap);
        lay.uo = o;     o += up4(cap);
        lay.cnt = o;    o += up4(cap);
        lay.tmp = o;    o += up4(h_r * grid);
        lay.total = o;
        lay.box_cap = cap;
        smem_r = sizeof(float) * (size_t)o;
        const int rk = grid == 64 ? 1 : 0;
        if (smem_r > smem_attr_set[rk]) {
            DH_CUDA_CHECK(cudaFuncSetAttribute(resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
            smem_attr_set[rk] = smem_r;
        }
        if (!occ_resize[rk] || occ_resize_smem[rk] != smem_r) {
            int per_sm = 1;
            DH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resize_kernel, kLossThreads, smem_r));
            occ_resize[rk] = per_sm < 1 ? 1 : per_sm;
            occ_resize_smem[rk] = smem_r;
        }
        grid_r = sms * occ_resize[rk];
        if (grid_r > pr.total_channels) grid_r = pr.total_channels;
    }
    fin.total_ctas = (unsigned int)(grid_f + grid_r);
    DH_CUDA_CHECK(cudaMemsetAsync(fin.done_counter, 0, 4 * sizeof(unsigned int), st));
    if (grid_r) {
        resize_kernel<<<grid_r, kLossThreads, smem_r, st>>>(pr, fin, lay);
        DH_LAUNCH_CHECK();
    }
    if (grid_f) {
        // The two kernels are independent (they only meet in loss_finish through an atomic ticket), so the second one is
        // a programmatic dependent launch: its CTAs become resident as the first kernel's CTAs retire instead of waiting
        // for the whole grid, which hides the launch gap and the tail of the first kernel.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid_f); cfg.blockDim = dim3(kLossThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = grid_r ? 1 : 0;
        DH_CUDA_CHECK(cudaLaunchKernelEx(&cfg, flat_kernel, pf, fin));
    }
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

int dh_scale_inplace_many(float* const* data_host, const size_t* n_host, int count, const float* scale, void* stream) {
    DH_REQUIRE(data_host && n_host && scale && count >= 0 && count <= kMaxLossLayers);
    ScaleMany m;
    memset(&m, 0, sizeof(m));
    unsigned int blocks = 0;
    for (int i = 0; i < count; ++i) {
        if (!data_host[i] || n_host[i] == 0) continue;
        if (reinterpret_cast<uintptr_t>(data_host[i]) & 15) return DH_ERR_INVALID_ARGUMENT;
        const size_t threads = n_host[i] / 4 > 3 ? n_host[i] / 4 : 4;
        m.data[m.count] = data_host[i]; m.n[m.count] = n_host[i]; m.block_begin[m.count] = blocks;
        blocks += (unsigned int)((threads + 255) / 256);
        ++m.count;
    }
    if (m.count == 0) return DH_OK;
    m.block_begin[m.count] = blocks;
    scale_many_kernel<<<blocks, 256, 0, as_stream(stream)>>>(m, scale);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
