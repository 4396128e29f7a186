// K4: masked activation-guidance losses and their gradients (SURVEY.md 8(a) rows 10, 10b, 10c;
// losses.py:4-84 with patch_size = 1, evaluated by guided_stable_diffuser.py:417-434).
//
// Two parts.
//
// (1) dh_build_loss_plan - once per edit.  The correspondence list has ~6x duplicates at the 64x64 loss grid
//     (SURVEY.md "hard parts"); the plan groups it by DESTINATION cell (counting sort) and collapses equal
//     (src, dst) cell pairs into one entry with a multiplicity, giving a CSR over destination cells
//     (row_ptr, pairs = src | mult << 16) plus per-cell multiplicities of the three background lists.
//
// (2) dh_guidance_loss - every denoising step.  One persistent CTA per SM walks over (layer, channel)
//     planes: the current and the recorded plane are staged in shared memory with TMA bulk copies
//     (double buffered, mbarrier complete_tx), optionally resized bilinearly to the loss grid, then every
//     thread owns a fixed, interleaved set of destination cells and evaluates
//         L_fg  = 1/(C N)  sum_pairs mult * |up(orig)[src] - up(cur)[dst]|
//         dL/dup(cur)[dst] = -1/(C N) sum_pairs mult * sign(...)          (INTEGER accumulation per cell)
//     and the background term, without any atomic: the gradient is gathered per destination cell, so it is
//     bit-reproducible (the reference's index_put(accumulate=True) backward is not, on CUDA).  The gradient is
//     written at native resolution (transposed bilinear resize in gather form).  Loss value and gradient come
//     out of the same pass: algorithmic traffic = read cur + read orig + write grad.
//     A tiny second kernel reduces the per-channel partial sums in a fixed order.
#include "dh_common.cuh"

#include <string.h>

namespace dh {

constexpr int kLossThreads = 1024;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kMaxLossLayers = 8;
constexpr int kMaxG = 64;
constexpr int kMaxCells = kMaxG * kMaxG;          // 4096
constexpr int kMaxNative = 64;

struct PlanHeader {
    int32_t n_pairs, n_fg, n_bg_orig, n_bg_trans, n_bg_common, grid, cap, reserved;
};

struct PlanView {
    PlanHeader* hdr;
    int32_t* row_ptr;      // cells + 1
    ushort4* bgcnt;        // cells: (count in bg_orig, bg_trans, bg_common, bit0 = cell is the source of a pair)
    uint2* pairs;          // cap entries: x = src | dst << 16, y = multiplicity
};

__host__ __device__ inline size_t plan_layout(int grid, int cap, size_t* o_row, size_t* o_bg, size_t* o_pairs) {
    const size_t cells = (size_t)grid * grid;
    size_t o = sizeof(PlanHeader);
    *o_row = o;   o += ((cells + 1) * sizeof(int32_t) + 15) / 16 * 16;
    *o_bg = o;    o += cells * sizeof(ushort4);
    *o_pairs = o; o += ((size_t)(cap > 0 ? cap : 1) * sizeof(uint2) + 15) / 16 * 16;
    return o;
}

__host__ __device__ inline PlanView plan_view(void* plan, int grid, int cap) {
    size_t a, b, c;
    plan_layout(grid, cap, &a, &b, &c);
    char* p = static_cast<char*>(plan);
    PlanView v;
    v.hdr = reinterpret_cast<PlanHeader*>(p);
    v.row_ptr = reinterpret_cast<int32_t*>(p + a);
    v.bgcnt = reinterpret_cast<ushort4*>(p + b);
    v.pairs = reinterpret_cast<uint2*>(p + c);
    return v;
}

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
    for (int n = tid; n < n_bg_orig; n += kPlanThreads) atomicAdd(co + bg_orig[n], 1);
    for (int n = tid; n < n_bg_trans; n += kPlanThreads) atomicAdd(ct + bg_trans[n], 1);
    for (int n = tid; n < n_bg_common; n += kPlanThreads) atomicAdd(cc + bg_common[n], 1);
    __syncthreads();
    for (int i = tid; i < cells; i += kPlanThreads) {
        ushort4 v;
        v.x = (unsigned short)min(co[i], 65535); v.y = (unsigned short)min(ct[i], 65535);
        v.z = (unsigned short)min(cc[i], 65535); v.w = (unsigned short)(is_src[i] ? 1 : 0);
        pv.bgcnt[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// fused loss + gradient
// ------------------------------------------------------------------------------------------------
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
    const void* plan;
    int plan_cap;
    int n_fg, n_bg_orig, n_bg_trans, n_bg_common;
    int fg_kind;        // 0 = off, 1 = local_avg patch 1
    int bg_kind;        // 0 = off, 1 = global_avg, 2 = local_avg
    float* partial;     // [total_channels][2]: fg sum, bg term
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
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

// Shared-memory carve-up (floats unless noted), sized by the launch from the layer shapes:
//   stage[2][2][hw_max]   double-buffered current / recorded plane at native resolution (TMA destination)
//   cnt[GG] (int)         per destination cell: - sum mult * sign(d); rewritten in place as the float gradient
//   pairs[cap] (uint2)    the plan's (src | dst << 16, mult) entries when they fit
//   -- only when a layer is smaller than the loss grid --
//   uc[GG], uo[GG]        up(cur), up(orig) inside the active box
//   tmp[h_max * G]        separable transposed resize
//   wo[hw_r], wt[hw_r]    up^T(background multiplicities) at native resolution
//   wrow[h_max][kWin], wcol[w_max][kWin]   transposed-resize weights per native row / column
struct LossSmemLayout {
    int stage, stage_stride, cnt, pairs, pairs_cap, uc, uo, tmp, wo, wt, wrow, wcol, total_floats;
};

constexpr int kWin = 16;      // up rows (columns) that can touch one native row (column): G/h * 2 <= 16 for h >= 8

struct LossTables {
    float red[kLossWarps][4];
    int ty0[kMaxG], ty1[kMaxG], tx0[kMaxG], tx1[kMaxG];
    float tly[kMaxG], tlx[kMaxG];
    int ylo[kMaxNative], yhi[kMaxNative], xlo[kMaxNative], xhi[kMaxNative];
    int box[8];                        // active box in up space: r0, r1, s0, s1; native: y0, y1, x0, x1
    int flags[2];
    int scratch[kLossWarps * 8];
    float lconst[kMaxLossLayers][4];   // per layer: fscale, lscale, bgw / (C * n_bg_trans), unused
    uint64_t full[2];
};

struct LossLaunch {
    LossSmemLayout lay;
    unsigned int* done_counter;       // zeroed before the launch; the last CTA reduces the partial sums
    float* loss_out;
};

__device__ __forceinline__ void layer_of(const LossParams& p, int gc, int& l) {
    l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLossLayers; ++i)
        if (i < p.n_layers && gc >= p.lv[i].chan_begin) l = i;
}

// sum over the CTA of three values, result in every thread; fixed order (warp tree, then a tree over the 32 warp
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

__device__ void loss_finalize(const LossParams& p, float* __restrict__ loss_out, float (*red)[4]);

__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One CTA of 1024 threads per SM.  Thread t owns the four consecutive loss-grid cells 4t .. 4t+3 (128-bit shared
// loads / global stores); foreground pairs are walked one per thread.
__global__ void __launch_bounds__(kLossThreads, 1) guidance_loss_kernel(const __grid_constant__ LossParams p,
                                                                        const __grid_constant__ LossLaunch lp) {
    extern __shared__ __align__(128) float lsm[];
    __shared__ __align__(16) LossTables tb;
    __shared__ unsigned int ticket;
    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    const int G = p.G, GG = G * G;
    const PlanView pv = plan_view(const_cast<void*>(p.plan), G, p.plan_cap);
    const int n_pairs = pv.hdr->n_pairs;
    int* const cnt = reinterpret_cast<int*>(lsm + lp.lay.cnt);
    float* const gu = lsm + lp.lay.cnt;
    float* const suc = lsm + lp.lay.uc;
    float* const suo = lsm + lp.lay.uo;
    float* const tmp = lsm + lp.lay.tmp;
    float* const wo = lsm + lp.lay.wo;
    float* const wt = lsm + lp.lay.wt;
    float* const wrow = lsm + lp.lay.wrow;
    float* const wcol = lsm + lp.lay.wcol;
    uint2* const spairs = reinterpret_cast<uint2*>(lsm + lp.lay.pairs);
    const bool pairs_in_smem = n_pairs <= lp.lay.pairs_cap;
    const uint2* const pairs = pairs_in_smem ? spairs : pv.pairs;

    if (tid == 0) {
        mbar_init(&tb.full[0], 1);
        mbar_init(&tb.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (pairs_in_smem)
        for (int i = tid; i < n_pairs; i += kLossThreads) spairs[i] = pv.pairs[i];
    // per-thread constants: 4-bit masks of the own cells in the three background lists (multiplicities are 0/1 for
    // lists that come from np.nonzero), and the box of the cells that are a source or a destination of a pair
    uint32_t mo = 0, mt = 0, mc = 0;
    bool binary = true;
    int r0 = G, r1 = -1, s0 = G, s1 = -1;
    const int q0 = 4 * tid;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q = q0 + k;
        if (q < GG) {
            const ushort4 bc = pv.bgcnt[q];
            mo |= (bc.x ? 1u : 0u) << k; mt |= (bc.y ? 1u : 0u) << k; mc |= (bc.z ? 1u : 0u) << k;
            binary = binary && bc.x <= 1 && bc.y <= 1 && bc.z <= 1;
            if (p.fg_kind && (pv.row_ptr[q + 1] > pv.row_ptr[q] || (bc.w & 1))) {
                const int r = q / G, s = q - r * G;
                r0 = min(r0, r); r1 = max(r1, r); s0 = min(s0, s); s1 = max(s1, s);
            }
        }
    }
    if (p.bg_kind == 2) { r0 = 0; r1 = G - 1; s0 = 0; s1 = G - 1; }      // local_avg needs up() on every background cell
    {
        r0 = __reduce_min_sync(0xFFFFFFFFu, r0); s0 = __reduce_min_sync(0xFFFFFFFFu, s0);
        r1 = __reduce_max_sync(0xFFFFFFFFu, r1); s1 = __reduce_max_sync(0xFFFFFFFFu, s1);
        const bool wbin = __all_sync(0xFFFFFFFFu, binary);
        int* bx = tb.scratch;
        if (lane == 0) { bx[wid * 8 + 0] = r0; bx[wid * 8 + 1] = r1; bx[wid * 8 + 2] = s0; bx[wid * 8 + 3] = s1; bx[wid * 8 + 4] = wbin ? 1 : 0; }
        __syncthreads();
        if (tid == 0) {
            bool allbin = true;
            for (int i = 0; i < kLossWarps; ++i) {
                r0 = min(r0, bx[i * 8]); r1 = max(r1, bx[i * 8 + 1]); s0 = min(s0, bx[i * 8 + 2]); s1 = max(s1, bx[i * 8 + 3]);
                allbin = allbin && bx[i * 8 + 4] != 0;
            }
            tb.box[0] = r0; tb.box[1] = r1; tb.box[2] = s0; tb.box[3] = s1;
            tb.flags[0] = allbin ? 1 : 0;
        }
        __syncthreads();
    }
    if (tid < p.n_layers) {
        const LossLayerDev& L = p.lv[tid];
        tb.lconst[tid][0] = p.fg_kind ? L.fgw / ((float)L.C * (float)p.n_fg) : 0.0f;
        tb.lconst[tid][1] = p.bg_kind == 2 ? L.bgw / ((float)L.C * (float)p.n_bg_common) : 0.0f;
        tb.lconst[tid][2] = p.bg_kind == 1 ? L.bgw / ((float)L.C * (float)p.n_bg_trans) : 0.0f;
        tb.lconst[tid][3] = 0.0f;
    }
    const float inv_no = 1.0f / (float)p.n_bg_orig, inv_nt = 1.0f / (float)p.n_bg_trans;
    __syncthreads();
    const int br0 = tb.box[0], br1 = tb.box[1], bs0 = tb.box[2], bs1 = tb.box[3];
    const bool any_box = br1 >= br0;
    const bool bg_binary = tb.flags[0] != 0;
    float fo[4], ft[4], fc[4];          // multiplicities of the own cells (general case: read once, kept in registers)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        fo[k] = (float)((mo >> k) & 1u); ft[k] = (float)((mt >> k) & 1u); fc[k] = (float)((mc >> k) & 1u);
        if (!bg_binary && q0 + k < GG) {
            const ushort4 bc = pv.bgcnt[q0 + k];
            fo[k] = (float)bc.x; ft[k] = (float)bc.y; fc[k] = (float)bc.z;
        }
    }

    auto issue = [&](int gc, int buf) {
        int l;
        layer_of(p, gc, l);
        const LossLayerDev& L = p.lv[l];
        const int hw = L.h * L.w;
        const size_t off = (size_t)(gc - L.chan_begin) * hw;
        float* st = lsm + lp.lay.stage + buf * lp.lay.stage_stride;
        mbar_expect_tx(&tb.full[buf], 2u * hw * 4u);
        tma_g2s(st, L.cur + off, hw * 4u, &tb.full[buf]);
        tma_g2s(st + hw, L.orig + off, hw * 4u, &tb.full[buf]);
    };

    // foreground pairs, one per thread: the integer sign term goes to cnt[dst] with a shared-memory integer atomic -
    // integer addition is associative, so the gradient does not depend on the order.
    auto walk_pairs = [&](const float* uc, const float* uo, float& acc) {
        for (int j = tid; j < n_pairs; j += kLossThreads) {
            const uint2 e = pairs[j];
            const int d = (int)(e.x >> 16);
            const float df = uo[e.x & 0xFFFFu] - uc[d];
            acc = fmaf((float)e.y, fabsf(df), acc);
            if (df != 0.0f) atomicAdd(cnt + d, df > 0.0f ? -(int)e.y : (int)e.y);
        }
    };

    int it = 0, cur_layer = -1;
    if (tid == 0 && (int)blockIdx.x < p.total_channels) issue(blockIdx.x, 0);
    for (int gc = blockIdx.x; gc < p.total_channels; gc += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nxt = gc + gridDim.x;
        if (tid == 0 && nxt < p.total_channels) issue(nxt, buf ^ 1);     // prefetch the next plane pair
        int l;
        layer_of(p, gc, l);
        const LossLayerDev& L = p.lv[l];
        const int c = gc - L.chan_begin;
        const int h = L.h, w = L.w, hw = h * w;
        const bool resize = (h != G) || (w != G);
        const float* const pc_ = lsm + lp.lay.stage + buf * lp.lay.stage_stride;   // current plane, native resolution
        const float* const po_ = pc_ + hw;                                          // recorded plane
        if (q0 < GG) *reinterpret_cast<int4*>(cnt + q0) = make_int4(0, 0, 0, 0);
        if (resize && l != cur_layer) {
            // ---- per-layer tables (a CTA crosses a layer boundary at most n_layers times) ----
            const float sy = (float)h / (float)G, sx = (float)w / (float)G;
            if (tid < G) {
                bilinear_tap(tid, h, sy, tb.ty0[tid], tb.ty1[tid], tb.tly[tid]);
                bilinear_tap(tid, w, sx, tb.tx0[tid], tb.tx1[tid], tb.tlx[tid]);
            }
            __syncthreads();
            if (tid < h) {       // up rows that touch native row tid, and their weights
                int lo = G, hi = -1;
                for (int r = 0; r < G; ++r)
                    if (tb.ty0[r] == tid || tb.ty1[r] == tid) { lo = min(lo, r); hi = max(hi, r); }
                hi = min(hi, lo + kWin - 1);
                tb.ylo[tid] = lo; tb.yhi[tid] = hi;
                for (int r = lo; r <= hi; ++r)
                    wrow[tid * kWin + r - lo] = (tb.ty0[r] == tid ? 1.0f - tb.tly[r] : 0.0f) + (tb.ty1[r] == tid ? tb.tly[r] : 0.0f);
            }
            if (tid >= 64 && tid - 64 < w) {
                const int j = tid - 64;
                int lo = G, hi = -1;
                for (int s = 0; s < G; ++s)
                    if (tb.tx0[s] == j || tb.tx1[s] == j) { lo = min(lo, s); hi = max(hi, s); }
                hi = min(hi, lo + kWin - 1);
                tb.xlo[j] = lo; tb.xhi[j] = hi;
                for (int s = lo; s <= hi; ++s)
                    wcol[j * kWin + s - lo] = (tb.tx0[s] == j ? 1.0f - tb.tlx[s] : 0.0f) + (tb.tx1[s] == j ? tb.tlx[s] : 0.0f);
            }
            if (tid == 128) {
                tb.box[4] = any_box ? tb.ty0[br0] : 0; tb.box[5] = any_box ? tb.ty1[br1] : -1;
                tb.box[6] = any_box ? tb.tx0[bs0] : 0; tb.box[7] = any_box ? tb.tx1[bs1] : -1;
            }
            __syncthreads();
            if (p.bg_kind == 1) {
                // wo / wt = up^T applied to the background multiplicities (separable, via tmp)
                for (int pass = 0; pass < 2; ++pass) {
                    float* dst = pass == 0 ? wo : wt;
                    for (int yi = wid; yi < h; yi += kLossWarps)
                        for (int s = lane; s < G; s += 32) {
                            float a = 0.0f;
                            for (int r = tb.ylo[yi]; r <= tb.yhi[yi]; ++r) {
                                const ushort4 bc = pv.bgcnt[r * G + s];
                                a = fmaf(wrow[yi * kWin + r - tb.ylo[yi]], (float)(pass == 0 ? bc.x : bc.y), a);
                            }
                            tmp[yi * G + s] = a;
                        }
                    __syncthreads();
                    for (int yi = wid; yi < h; yi += kLossWarps)
                        for (int xj = lane; xj < w; xj += 32) {
                            float a = 0.0f;
                            for (int s = tb.xlo[xj]; s <= tb.xhi[xj]; ++s) a = fmaf(wcol[xj * kWin + s - tb.xlo[xj]], tmp[yi * G + s], a);
                            dst[yi * w + xj] = a;
                        }
                    __syncthreads();
                }
            }
        }
        cur_layer = l;
        mbar_wait(&tb.full[buf], (it >> 1) & 1);
        __syncthreads();          // cnt is zero, the planes have landed
        float acc_f = 0.0f, so = 0.0f, sc = 0.0f;
        const float fscale = tb.lconst[l][0], lscale = tb.lconst[l][1], gscale = tb.lconst[l][2];
        float* g = L.grad ? L.grad + (size_t)c * hw : nullptr;

        if (!resize) {
            // ================= layer already at the loss grid =================
            if (p.fg_kind) walk_pairs(pc_, po_, acc_f);
            float4 vc = make_float4(0, 0, 0, 0), vo = vc;
            if (q0 < GG) {
                vc = *reinterpret_cast<const float4*>(pc_ + q0);
                vo = *reinterpret_cast<const float4*>(po_ + q0);
            }
            if (p.bg_kind == 1) {
                so = fmaf(fo[0], vo.x, fmaf(fo[1], vo.y, fmaf(fo[2], vo.z, fo[3] * vo.w)));
                sc = fmaf(ft[0], vc.x, fmaf(ft[1], vc.y, fmaf(ft[2], vc.z, ft[3] * vc.w)));
            } else if (p.bg_kind == 2) {
                so = fmaf(fc[0], fabsf(vo.x - vc.x), fmaf(fc[1], fabsf(vo.y - vc.y), fmaf(fc[2], fabsf(vo.z - vc.z), fc[3] * fabsf(vo.w - vc.w))));
            }
            float r0s = acc_f, r1s = so, r2s = sc;
            block_sum3(r0s, r1s, r2s, tb.red);       // (its barrier also orders the cnt atomics)
            float bg_term = 0.0f, bscale = 0.0f;
            if (p.bg_kind == 1) {
                const float delta = r1s * inv_no - r2s * inv_nt;
                bg_term = fabsf(delta);
                bscale = delta > 0.0f ? -gscale : (delta < 0.0f ? gscale : 0.0f);
            } else if (p.bg_kind == 2) {
                bg_term = r1s;
            }
            if (tid == 0) { p.partial[2 * gc] = r0s; p.partial[2 * gc + 1] = bg_term; }
            if (g && q0 < GG) {
                const int4 ci = *reinterpret_cast<const int4*>(cnt + q0);
                float4 v = make_float4((float)ci.x * fscale, (float)ci.y * fscale, (float)ci.z * fscale, (float)ci.w * fscale);
                if (p.bg_kind == 1) {
                    v.x = fmaf(ft[0], bscale, v.x); v.y = fmaf(ft[1], bscale, v.y); v.z = fmaf(ft[2], bscale, v.z); v.w = fmaf(ft[3], bscale, v.w);
                } else if (p.bg_kind == 2) {
                    const float dx = vo.x - vc.x, dy = vo.y - vc.y, dz = vo.z - vc.z, dw = vo.w - vc.w;
                    v.x -= (float)((dx > 0.0f) - (dx < 0.0f)) * fc[0] * lscale; v.y -= (float)((dy > 0.0f) - (dy < 0.0f)) * fc[1] * lscale;
                    v.z -= (float)((dz > 0.0f) - (dz < 0.0f)) * fc[2] * lscale; v.w -= (float)((dw > 0.0f) - (dw < 0.0f)) * fc[3] * lscale;
                }
                st_cs_f4(g + q0, v);
            }
        } else {
            // ================= smaller layer: bilinear resize restricted to the active box =================
            const int ny0 = tb.box[4], ny1 = tb.box[5], nx0 = tb.box[6], nx1 = tb.box[7];
            for (int r = br0 + wid; r <= br1; r += kLossWarps) {
                const int y0 = tb.ty0[r] * w, y1 = tb.ty1[r] * w;
                const float ly = tb.tly[r], hy = 1.0f - ly;
                for (int s = bs0 + lane; s <= bs1; s += 32) {
                    const int x0 = tb.tx0[s], x1 = tb.tx1[s];
                    const float lx = tb.tlx[s], hx = 1.0f - lx;
                    suc[r * G + s] = hy * (hx * pc_[y0 + x0] + lx * pc_[y0 + x1]) + ly * (hx * pc_[y1 + x0] + lx * pc_[y1 + x1]);
                    suo[r * G + s] = hy * (hx * po_[y0 + x0] + lx * po_[y0 + x1]) + ly * (hx * po_[y1 + x0] + lx * po_[y1 + x1]);
                }
            }
            __syncthreads();
            if (p.fg_kind) walk_pairs(suc, suo, acc_f);
            float4 vc = make_float4(0, 0, 0, 0), vo = vc;
            if (p.bg_kind == 1) {     // background sums at native resolution: <wo, orig>, <wt, cur>
                for (int i = tid * 4; i < hw; i += kLossThreads * 4) {
                    const float4 a = *reinterpret_cast<const float4*>(wo + i), b = *reinterpret_cast<const float4*>(po_ + i);
                    const float4 e = *reinterpret_cast<const float4*>(wt + i), f = *reinterpret_cast<const float4*>(pc_ + i);
                    so = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, so))));
                    sc = fmaf(e.x, f.x, fmaf(e.y, f.y, fmaf(e.z, f.z, fmaf(e.w, f.w, sc))));
                }
            } else if (p.bg_kind == 2 && q0 < GG) {
                vc = *reinterpret_cast<const float4*>(suc + q0);
                vo = *reinterpret_cast<const float4*>(suo + q0);
                so = fmaf(fc[0], fabsf(vo.x - vc.x), fmaf(fc[1], fabsf(vo.y - vc.y), fmaf(fc[2], fabsf(vo.z - vc.z), fc[3] * fabsf(vo.w - vc.w))));
            }
            float r0s = acc_f, r1s = so, r2s = sc;
            block_sum3(r0s, r1s, r2s, tb.red);
            float bg_term = 0.0f, bscale = 0.0f;
            if (p.bg_kind == 1) {
                const float delta = r1s * inv_no - r2s * inv_nt;
                bg_term = fabsf(delta);
                bscale = delta > 0.0f ? -gscale : (delta < 0.0f ? gscale : 0.0f);
            } else if (p.bg_kind == 2) {
                bg_term = r1s;
            }
            if (tid == 0) { p.partial[2 * gc] = r0s; p.partial[2 * gc + 1] = bg_term; }
            if (g) {
                // gradient w.r.t. up(cur) (zero outside the box), then the transposed resize in gather form
                if (q0 < GG) {
                    const int4 ci = *reinterpret_cast<const int4*>(cnt + q0);
                    float4 v = make_float4((float)ci.x * fscale, (float)ci.y * fscale, (float)ci.z * fscale, (float)ci.w * fscale);
                    if (p.bg_kind == 2) {
                        const float dx = vo.x - vc.x, dy = vo.y - vc.y, dz = vo.z - vc.z, dw = vo.w - vc.w;
                        v.x -= (float)((dx > 0.0f) - (dx < 0.0f)) * fc[0] * lscale; v.y -= (float)((dy > 0.0f) - (dy < 0.0f)) * fc[1] * lscale;
                        v.z -= (float)((dz > 0.0f) - (dz < 0.0f)) * fc[2] * lscale; v.w -= (float)((dw > 0.0f) - (dw < 0.0f)) * fc[3] * lscale;
                    }
                    *reinterpret_cast<float4*>(gu + q0) = v;      // own cells: integer read above, float written in place
                }
                __syncthreads();
                for (int yi = ny0 + wid; yi <= ny1; yi += kLossWarps) {
                    const int lo = tb.ylo[yi];
                    const int ra = max(lo, br0), rb = min(tb.yhi[yi], br1);
                    for (int s = bs0 + lane; s <= bs1; s += 32) {
                        float a = 0.0f;
                        for (int r = ra; r <= rb; ++r) a = fmaf(wrow[yi * kWin + r - lo], gu[r * G + s], a);
                        tmp[yi * G + s] = a;
                    }
                }
                __syncthreads();
                for (int yi = wid; yi < h; yi += kLossWarps) {
                    const bool row_in = yi >= ny0 && yi <= ny1;
                    for (int xj = lane; xj < w; xj += 32) {
                        float a = 0.0f;
                        if (row_in && xj >= nx0 && xj <= nx1) {
                            const int lo = tb.xlo[xj];
                            const int sa = max(lo, bs0), sb = min(tb.xhi[xj], bs1);
                            for (int s = sa; s <= sb; ++s) a = fmaf(wcol[xj * kWin + s - lo], tmp[yi * G + s], a);
                        }
                        if (p.bg_kind == 1) a = fmaf(bscale, wt[yi * w + xj], a);
                        g[yi * w + xj] = a;
                    }
                }
            }
        }
        __syncthreads();     // every read of the stage / cnt / uc / tmp is done before the next iteration reuses them
    }
    // ---- the last CTA to finish reduces the per-channel partial sums (fixed order) ----
    __threadfence();
    if (tid == 0) ticket = atomicAdd(lp.done_counter, 1u);
    __syncthreads();
    if (ticket == gridDim.x - 1) {
        __threadfence();
        loss_finalize(p, lp.loss_out, tb.red);
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) sm[warp_id()] = v;
    __syncthreads();
    float t = lane_id() < (int)(blockDim.x >> 5) ? sm[lane_id()] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    __syncthreads();
    return t;
}

// Fixed-order reduction of the per-channel partials -> loss_out[0] = total, [1+2l] = fg_l, [2+2l] = bg_l.
__device__ void loss_finalize(const LossParams& p, float* __restrict__ loss_out, float (*red)[4]) {
    float* sm = &red[0][0];      // kLossWarps * 4 floats >= 32
    float total = 0.0f;
    for (int l = 0; l < p.n_layers; ++l) {
        const LossLayerDev& L = p.lv[l];
        float a = 0.0f, b = 0.0f;
        for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
            a += __ldcg(p.partial + 2 * (L.chan_begin + c));
            b += __ldcg(p.partial + 2 * (L.chan_begin + c) + 1);
        }
        a = block_sum_256(a, sm);
        b = block_sum_256(b, sm);
        const float fg = p.fg_kind ? a / (float)p.n_fg / (float)L.C : 0.0f;
        const float bg = p.bg_kind == 2 ? b / (float)p.n_bg_common / (float)L.C : (p.bg_kind == 1 ? b / (float)L.C : 0.0f);
        if (threadIdx.x == 0) {
            loss_out[1 + 2 * l] = fg;
            loss_out[2 + 2 * l] = bg;
        }
        if (p.fg_kind) total += L.fgw * fg;
        if (p.bg_kind) total += L.bgw * bg;
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

size_t dh_guidance_loss_workspace_bytes(int n_layers, int max_channels) {
    if (n_layers < 1 || max_channels < 1) return 0;
    return sizeof(float) * 2 * (size_t)n_layers * max_channels + 16;
}

int dh_guidance_loss(const dh_loss_layer* layers_host, int n_layers, int grid, const void* plan, int n_fg, int n_bg_orig,
                     int n_bg_trans, int n_bg_common, int fg_kind, int bg_kind, float* loss_out, void* ws, size_t ws_bytes,
                     void* stream) {
    DH_REQUIRE(layers_host && n_layers >= 1 && n_layers <= kMaxLossLayers && loss_out && ws && plan);
    DH_REQUIRE(grid >= 1 && grid <= kMaxG);
    DH_REQUIRE(n_fg >= 0 && n_bg_orig >= 0 && n_bg_trans >= 0 && n_bg_common >= 0);
    if (bg_kind != 0 && bg_kind != 1 && bg_kind != 2) return DH_ERR_INVALID_ARGUMENT;
    if (fg_kind != 0 && fg_kind != 1) return DH_ERR_INVALID_ARGUMENT;
    LossParams p;
    memset(&p, 0, sizeof(p));
    int chan = 0, hw_max = 0, hw_r = 0, h_r = 0;
    for (int i = 0; i < n_layers; ++i) {
        const dh_loss_layer& s = layers_host[i];
        DH_REQUIRE(s.cur && s.orig && s.channels >= 1 && s.h >= 1 && s.w >= 1);
        if (s.h > kMaxNative || s.w > kMaxNative) return DH_ERR_UNSUPPORTED;
        if (((size_t)s.h * s.w) % 4 != 0) return DH_ERR_UNSUPPORTED;     // TMA bulk copies move multiples of 16 bytes
        if ((reinterpret_cast<uintptr_t>(s.cur) & 15) || (reinterpret_cast<uintptr_t>(s.orig) & 15) ||
            (s.grad && (reinterpret_cast<uintptr_t>(s.grad) & 15)))
            return DH_ERR_INVALID_ARGUMENT;
        LossLayerDev& L = p.lv[i];
        L.cur = s.cur; L.orig = s.orig; L.grad = s.grad;
        L.C = s.channels; L.h = s.h; L.w = s.w; L.fgw = s.fg_weight; L.bgw = s.bg_weight;
        L.chan_begin = chan;
        chan += s.channels;
        const int hw = s.h * s.w;
        if (hw > hw_max) hw_max = hw;
        if (s.h != grid || s.w != grid) {
            if (hw > hw_r) hw_r = hw;
            if (s.h > h_r) h_r = s.h;
        }
    }
    const size_t partial_bytes = sizeof(float) * 2 * (size_t)chan;
    if (ws_bytes < partial_bytes + 16) return DH_ERR_WORKSPACE;
    p.n_layers = n_layers; p.total_channels = chan; p.G = grid;
    p.plan = plan; p.plan_cap = n_fg;
    p.n_fg = n_fg; p.n_bg_orig = n_bg_orig; p.n_bg_trans = n_bg_trans; p.n_bg_common = n_bg_common;
    p.fg_kind = fg_kind; p.bg_kind = bg_kind;
    p.partial = static_cast<float*>(ws);
    LossLaunch lp;
    const int GG = grid * grid;
    auto up4 = [](int v) { return (v + 3) / 4 * 4; };
    int o = 0;
    lp.lay.stage = o; lp.lay.stage_stride = up4(2 * hw_max); o += 2 * lp.lay.stage_stride;
    lp.lay.cnt = o;   o += up4(GG);
    lp.lay.uc = o;    if (hw_r) o += up4(GG);
    lp.lay.uo = o;    if (hw_r) o += up4(GG);
    lp.lay.tmp = o;   if (hw_r) o += up4(h_r * grid);
    lp.lay.wo = o;    if (hw_r) o += up4(hw_r);
    lp.lay.wt = o;    if (hw_r) o += up4(hw_r);
    lp.lay.wrow = o;  if (hw_r) o += kMaxNative * kWin;
    lp.lay.wcol = o;  if (hw_r) o += kMaxNative * kWin;
    // whatever is left of the 227 KB goes to the pair list (8 bytes per entry); larger lists are read from global memory
    lp.lay.pairs = o;
    const int max_floats = (227 * 1024 - 8192) / 4;     // static tables + alignment slack
    int cap = (max_floats - o) / 2;
    if (cap > n_fg) cap = n_fg;
    if (cap < 0) cap = 0;
    lp.lay.pairs_cap = cap;
    o += up4(2 * cap);
    lp.lay.total_floats = o;
    lp.done_counter = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + (partial_bytes + 15) / 16 * 16);
    if (ws_bytes < (partial_bytes + 15) / 16 * 16 + sizeof(unsigned int)) return DH_ERR_WORKSPACE;
    lp.loss_out = loss_out;
    const size_t smem = sizeof(float) * (size_t)o;
    cudaStream_t st = as_stream(stream);
    DH_CUDA_CHECK(cudaFuncSetAttribute(guidance_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    DH_CUDA_CHECK(cudaGetDevice(&dev));
    DH_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid_dim = sms;
    if (grid_dim > chan) grid_dim = chan;
    DH_CUDA_CHECK(cudaMemsetAsync(lp.done_counter, 0, sizeof(unsigned int), st));
    guidance_loss_kernel<<<grid_dim, kLossThreads, smem, st>>>(p, lp);
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
