// K4: masked activation-guidance losses and their gradients (SURVEY.md 8(a) rows 10, 10b, 10c;
// losses.py:4-84 with patch_size = 1, evaluated by guided_stable_diffuser.py:417-434).
//
// Two parts.
//
// (1) dh_build_loss_plan - once per edit.  The correspondence list has ~6x duplicates at the 64x64 loss grid
//     (SURVEY.md "hard parts"); the plan groups it by DESTINATION cell (counting sort) and collapses equal
//     (src, dst) cell pairs into one entry with a multiplicity: a CSR over destination cells (row_ptr, 8-byte pairs)
//     for the general patch kernel, a sliced-ELL view of the same pairs for this file's kernel (dh_loss_plan.cuh),
//     plus per-cell multiplicities of the three background lists.
//
// (2) dh_guidance_loss - every denoising step.  ONE persistent launch for every layer: small CTAs (256 threads, four per
//     SM) pull work items - one 64x64 plane pair, or a few planes of a smaller layer - from a dynamic queue.  The item's
//     `cur` and `orig` planes are staged in shared memory by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) that
//     thread 0 issues for the NEXT item as soon as every warp is done reading the current one, so the HBM reads overlap
//     the gradient write-out of this CTA and the arithmetic of the other CTAs of the SM.
//     Per plane
//         L_fg  = 1/(C N)  sum_pairs mult * |up(orig)[src] - up(cur)[dst]|
//         dL/dup(cur)[dst] = -1/(C N) sum_pairs mult * sign(...)
//     plus the background term.  One THREAD owns one destination cell (sliced ELL): it reads cur[dst] once, walks the
//     cell's distinct sources and keeps the INTEGER sign count in a register - no atomics, and integer addition is
//     associative, so the gradient is bit-reproducible (the reference's index_put(accumulate=True) backward is not, on
//     CUDA).  Layers smaller than the loss grid are resized bilinearly inside the box of pair cells only, and their
//     gradient is written at native resolution through the transposed resize in gather form.  Loss value and gradient
//     come out of the same pass: algorithmic traffic = read cur + read orig + write grad.  The last CTA to finish reduces
//     the per-channel partial sums in a fixed order and re-zeroes the two queue counters for the next launch.
#include "dh_common.cuh"
#include "dh_loss_plan.cuh"
#include "dh_tma.cuh"

#include <stdlib.h>
#include <string.h>

namespace dh {

// ------------------------------------------------------------------------------------------------
// plan builder: one CTA of 1024 threads
// ------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;
constexpr int kPlanWarps = kPlanThreads / 32;
constexpr int kPlanTab = 1024;          // per-warp counter table: source cells lo .. lo + 1023 of one destination bucket per pass
constexpr int kLenClasses = 256;        // rows are ordered by min(length, 255), descending

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
    uint32_t* tabs = reinterpret_cast<uint32_t*>(ucount + cells);   // kPlanWarps * kPlanTab
    int32_t* bucket = scratch;             // n_fg: sources grouped by destination cell
    int32_t* uniq = scratch + n_fg;        // n_fg: per bucket, distinct sources (ascending) | multiplicity << 16
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
    for (int n = tid; n < n_fg; n += kPlanThreads) bucket[atomicAdd(hist + fg_dst[n], 1)] = fg_src[n];
    __syncthreads();
    // per destination cell (one warp each, all 32 warps): count the sources in a private shared-memory table that covers
    // kPlanTab consecutive source cells per pass (one pass for any rigid edit: a bucket's sources span a few grid rows),
    // then emit the distinct sources in ascending order
    {
        uint32_t* cnt = tabs + (size_t)wid * kPlanTab;
        for (int d = wid; d < cells; d += kPlanWarps) {
            const int b0 = start[d], b1 = start[d + 1];
            if (b0 == b1) continue;
            int lo = 0x7FFFFFFF, hi = -1;
            for (int k = b0 + lane; k < b1; k += 32) { const int v = bucket[k]; lo = min(lo, v); hi = max(hi, v); }
            lo = __reduce_min_sync(0xFFFFFFFFu, lo);
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            int out = b0;
            for (int w0 = lo; w0 <= hi; w0 += kPlanTab) {
                const int w1 = min(hi, w0 + kPlanTab - 1);
                for (int i = lane; i <= w1 - w0; i += 32) cnt[i] = 0;
                __syncwarp();
                for (int k = b0 + lane; k < b1; k += 32) {
                    const int v = bucket[k];
                    if (v >= w0 && v <= w1) atomicAdd(cnt + (v - w0), 1u);
                }
                __syncwarp();
                for (int base = w0; base <= w1; base += 32) {
                    const int i = base + lane;
                    uint32_t c = i <= w1 ? cnt[i - w0] : 0u;
                    while (__any_sync(0xFFFFFFFFu, c > 0)) {
                        const unsigned b = __ballot_sync(0xFFFFFFFFu, c > 0);
                        const uint32_t m = c > 65535u ? 65535u : c;
                        if (c > 0) uniq[out + __popc(b & ((1u << lane) - 1u))] = (int32_t)((uint32_t)i | (m << 16));
                        c -= m;
                        out += __popc(b);
                    }
                }
                __syncwarp();
            }
            if (lane == 0) ucount[d] = out - b0;
        }
    }
    __syncthreads();
    // row_ptr = exclusive scan of the distinct-pair counts, then compact the buckets into the CSR
    s = 0;
    for (int i = c0; i < c1; ++i) s += ucount[i];
    run = block_exclusive_scan(s, scan_smem, total);
    const int n_pairs = total;
    for (int i = c0; i < c1; ++i) {
        pv.row_ptr[i] = run;
        const int b0 = start[i];
        for (int j = 0; j < ucount[i]; ++j) {
            const uint32_t e = (uint32_t)uniq[b0 + j];
            pv.pairs[run + j] = make_uint2((e & 0xFFFFu) | ((uint32_t)i << 16), e >> 16);
        }
        run += ucount[i];
    }
    if (tid == 0) pv.row_ptr[cells] = n_pairs;
    __syncthreads();

    // ---- sliced-ELL view: rows (destination cells with pairs) in descending order of their length, ties in raster order ----
    int* whist = reinterpret_cast<int*>(tabs);                       // [kPlanWarps][kLenClasses] per-warp class histograms
    int* wbase = whist + kPlanWarps * kLenClasses;                    // same shape: running output positions
    int* order = hist;                                                // row -> destination cell
    int* soff = start;                                                // slice -> first entry group
    for (int i = tid; i < kPlanWarps * kLenClasses; i += kPlanThreads) whist[i] = 0;
    __syncthreads();
    const int cpw = (cells + kPlanWarps - 1) / kPlanWarps;            // contiguous cells per warp
    for (int i = lane; i < cpw; i += 32) {
        const int cell = wid * cpw + i;
        const int k = cell < cells ? min(ucount[cell], kLenClasses - 1) : 0;
        if (k > 0) atomicAdd(whist + wid * kLenClasses + k, 1);
    }
    __syncthreads();
    {   // exclusive scan in (class descending, warp ascending) order; element j = (255 - class) * 32 + warp
        constexpr int kPer = kPlanWarps * kLenClasses / kPlanThreads;        // 8
        int cs[kPer];
        s = 0;
#pragma unroll
        for (int t = 0; t < kPer; ++t) {
            const int j = tid * kPer + t, k = kLenClasses - 1 - (j / kPlanWarps), w = j % kPlanWarps;
            cs[t] = k > 0 ? whist[w * kLenClasses + k] : 0;
            s += cs[t];
        }
        run = block_exclusive_scan(s, scan_smem, total);
#pragma unroll
        for (int t = 0; t < kPer; ++t) {
            const int j = tid * kPer + t, k = kLenClasses - 1 - (j / kPlanWarps), w = j % kPlanWarps;
            wbase[w * kLenClasses + k] = run;
            run += cs[t];
        }
    }
    const int n_rows = total;
    const int n_slices = (n_rows + 31) / 32;
    __syncthreads();
    for (int i0 = 0; i0 < cpw; i0 += 32) {      // stable placement: cells of a warp in ascending order, 32 at a time
        const int cell = wid * cpw + i0 + lane;
        const int k = (i0 + lane < cpw && cell < cells) ? min(ucount[cell], kLenClasses - 1) : 0;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, k);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (k > 0) order[wbase[wid * kLenClasses + k] + rank] = cell;
        __syncwarp();
        if (k > 0 && rank == 0) wbase[wid * kLenClasses + k] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // slice widths (longest row of the slice) and their exclusive scan
        int w = 0;
        if (tid < n_slices)
            for (int r = tid * 32; r < min(n_rows, tid * 32 + 32); ++r) w = max(w, (ucount[order[r]] + 3) & ~3);   // rows are padded to 4
        run = block_exclusive_scan(w, scan_smem, total);
        if (tid < n_slices) { soff[tid] = run; pv.ell_off[tid] = run; }
        if (tid == 0) { soff[n_slices] = total; pv.ell_off[n_slices] = total; }
    }
    const int n_groups = total;
    __syncthreads();
    // box of the cells that are the source or the destination of a pair (resized layers work inside it)
    __shared__ int box_sm[4];
    __shared__ int nonbin_sm;
    int* is_src = reinterpret_cast<int*>(tabs) + 2 * kPlanWarps * kLenClasses;     // cells ints behind the ELL histograms
    for (int i = tid; i < cells; i += kPlanThreads) is_src[i] = 0;
    if (tid == 0) { box_sm[0] = grid; box_sm[1] = -1; box_sm[2] = grid; box_sm[3] = -1; nonbin_sm = 0; }
    __syncthreads();
    for (int n = tid; n < n_fg; n += kPlanThreads) is_src[fg_src[n]] = 1;
    __syncthreads();
    for (int i = tid; i < cells; i += kPlanThreads)
        if (is_src[i] || ucount[i] > 0) {
            const int r = i / grid, c = i - r * grid;
            atomicMin(box_sm + 0, r); atomicMax(box_sm + 1, r); atomicMin(box_sm + 2, c); atomicMax(box_sm + 3, c);
        }
    __syncthreads();
    const int box_r0 = box_sm[0], box_r1 = box_sm[1], box_s0 = box_sm[2], box_s1 = box_sm[3];
    const int box_w = box_s1 >= box_s0 ? box_s1 - box_s0 + 1 : 0;
    auto to_box = [&](int cell) { const int r = cell / grid; return (r - box_r0) * box_w + (cell - r * grid - box_s0); };
    for (int sl = wid; sl < n_slices; sl += kPlanWarps) {
        const int r = sl * 32 + lane;
        const int cell = r < n_rows ? order[r] : -1;
        const int len = cell >= 0 ? ucount[cell] : 0;
        const int base = cell >= 0 ? pv.row_ptr[cell] : 0;
        pv.row_desc[r] = cell >= 0 ? ((uint32_t)cell | ((uint32_t)len << 16)) : 0u;
        pv.row_desc_box[r] = cell >= 0 ? ((uint32_t)to_box(cell) | ((uint32_t)len << 16)) : 0u;
        uint32_t* dst = pv.ent + (size_t)soff[sl] * 32 + lane;
        uint32_t* dst_box = pv.ent_box + (size_t)soff[sl] * 32 + lane;
        for (int k = 0; k < len; ++k) {
            const uint2 e = pv.pairs[base + k];
            dst[(size_t)k * 32] = (e.x & 0xFFFFu) | (e.y << 12);
            dst_box[(size_t)k * 32] = (uint32_t)to_box((int)(e.x & 0xFFFFu)) | (e.y << 12);
        }
        // every row is padded to a multiple of four entries with (first source, multiplicity 0): the kernel walks whole
        // groups of four without per-entry predicates
        for (int k = len; k < ((len + 3) & ~3); ++k) {
            dst[(size_t)k * 32] = dst[0] & 0xFFFu;
            dst_box[(size_t)k * 32] = dst_box[0] & 0xFFFu;
        }
    }
    if (tid == 0) {
        PlanHeader h;
        h.n_pairs = n_pairs; h.n_fg = n_fg; h.n_bg_orig = n_bg_orig; h.n_bg_trans = n_bg_trans; h.n_bg_common = n_bg_common;
        h.grid = grid; h.cap = n_fg; h.reserved = 0;
        h.box_r0 = box_r0; h.box_r1 = box_r1; h.box_s0 = box_s0; h.box_s1 = box_s1;
        h.n_rows = n_rows; h.n_slices = n_slices; h.n_groups = n_groups; h.pad = 0;
        *pv.hdr = h;
    }
    __syncthreads();
    // background list multiplicities per cell (lists from np.nonzero never repeat a cell; generic callers may)
    int* co = psm;
    int* ct = psm + cells;
    int* cc = psm + 2 * cells;
    int* rowf = psm + 3 * cells;          // cell is a destination row
    for (int i = tid; i < 3 * cells; i += kPlanThreads) psm[i] = 0;
    for (int i = tid; i < cells; i += kPlanThreads) rowf[i] = pv.row_ptr[i + 1] > pv.row_ptr[i] ? 1 : 0;
    __syncthreads();
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
    }
    // membership bits of the 16 own cells of every loss-kernel thread (cell 4 * (t + 256 k) + i <-> bit 4 k + i):
    // x = bg_orig | bg_trans << 16, y = bg_common | destination row << 16
    if (tid < 256) {
        uint32_t m_ot = 0, m_cr = 0;
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 4; ++i) {
                const int q = 4 * (tid + k * 256) + i, b = 4 * k + i;
                if (q < cells) {
                    m_ot |= (co[q] ? 1u : 0u) << b; m_ot |= (ct[q] ? 1u : 0u) << (16 + b);
                    m_cr |= (cc[q] ? 1u : 0u) << b; m_cr |= (rowf[q] ? 1u : 0u) << (16 + b);
                }
            }
        pv.own_masks[tid] = make_uint2(m_ot, m_cr);
    }
    __syncthreads();
    if (tid == 0) pv.hdr->reserved = nonbin_sm ? 0 : 1;      // bit 0: every background multiplicity is 0 or 1
}

// ------------------------------------------------------------------------------------------------
// fused loss + gradient
// ------------------------------------------------------------------------------------------------
constexpr int kLossThreads = 256;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kMaxGroups = 3;                    // groups of kLossThreads threads per CTA (one CTA per SM): 80 registers per thread
constexpr int kPlaneCap = kMaxG * kMaxG;         // floats per tensor in the stage
constexpr int kStageFloats = 2 * kPlaneCap;      // [cur planes][orig planes] = 32 KB
constexpr int kOwnGroups = kPlaneCap / 4 / kLossThreads;   // float4 groups per thread at the 64x64 grid = 4
constexpr int kRowUnroll = 4;    // independent gathers in flight per thread in the row walk
constexpr int kWin = 16;      // up rows (columns) that can touch one native row (column): 2 * G / h <= 16 for h >= 8
constexpr int kMaxPlanesPerItem = 2;     // resized layers: one pair of planes per item (short items balance better)

struct ResizeLayout {   // float offsets into the scratch area
    int tab, wrow, wcol, uc, uo, cnt, tmp, flat_cnt, total;
    int box_cap;        // capacity (cells) of the box-local uc / uo / cnt arrays
};

struct FusedLayer {
    const float* cur;
    const float* orig;
    float* grad;
    int C, h, w;
    float fgw, bgw;
    const void* tab;      // LayerTab of a resized layer (NULL for layers at the loss-grid resolution)
    int item_begin;       // first item of this layer among the items of its kind (flat / resized)
    int ppi;              // planes per item
    int partial_begin;    // first channel of this layer in the partial-sum array
    int flat;
};

struct FusedParams {
    FusedLayer lv[kMaxLossLayers];
    int n_layers, G;
    int n_flat_items, n_small_items;
    PlanView pv;        // the plan's arrays (device pointers, resolved on the host)
    int n_fg, n_bg_orig, n_bg_trans, n_bg_common;
    int fg_kind;        // 0 = off, 1 = local_avg patch 1
    int bg_kind;        // 0 = off, 1 = global_avg, 2 = local_avg
    float* partial;     // [all channels][2]: fg sum, bg term
    unsigned int* counters;   // [0] work queue, [1] finished CTAs: zero on entry, re-zeroed by the last CTA
    float* loss_out;
    ResizeLayout lay;
    int scratch_floats;
    int ell_floats, ell_desc_at, ell_ent_at, ell_ent_cap;    // shared-memory copy of the sliced-ELL plan (0 = read it from global)
    int ell_slices;     // slices of the plan as the caller read them from the plan header (0 = unknown: read the header)
    unsigned long long* debug;     // optional (DH_LOSS_DEBUG_BUF): per CTA {start ns, end ns, items, sm id}
};

struct WorkItem {
    int layer, c0, planes;
};

// Longest items first: the plane pairs of the resized layers (arithmetic heavy), then the 64x64 planes - the tail of the
// launch is then made of the shortest items.
__device__ __forceinline__ WorkItem decode_item(const FusedParams& p, int item) {
    const int flat = item >= p.n_small_items;
    const int idx = flat ? item - p.n_small_items : item;
    int l = -1;
#pragma unroll
    for (int i = 0; i < kMaxLossLayers; ++i)
        if (i < p.n_layers && p.lv[i].flat == flat && idx >= p.lv[i].item_begin) l = i;
    WorkItem w;
    w.layer = l;
    w.c0 = (idx - p.lv[l].item_begin) * p.lv[l].ppi;
    w.planes = min(p.lv[l].ppi, p.lv[l].C - w.c0);
    return w;
}

__device__ __forceinline__ float block_sum1(float v, float* sm) {      // over the first kLossThreads threads of the CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) sm[warp_id()] = v;
    asm volatile("bar.sync 1, %0;" ::"n"(kLossThreads) : "memory");
    float t = lane_id() < kLossWarps ? sm[lane_id()] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    asm volatile("bar.sync 1, %0;" ::"n"(kLossThreads) : "memory");
    return t;
}

// Called by the compute warps of every CTA when the queue is empty; the last CTA reduces the per-channel partials in a
// fixed order -> loss_out[0] = total, [1+2l] = fg_l, [2+2l] = bg_l, and re-arms the counters.
__device__ void loss_finish(const FusedParams& p, float* red32, unsigned int* ticket) {
    __threadfence();
    __syncthreads();                      // every group of the CTA is done
    if (threadIdx.x == 0) *ticket = atomicAdd(p.counters + 1, 1u);
    __syncthreads();
    if (*ticket != gridDim.x - 1 || threadIdx.x >= kLossThreads) return;
    __threadfence();
    float total = 0.0f;
    for (int l = 0; l < p.n_layers; ++l) {
        const FusedLayer& L = p.lv[l];
        float a = 0.0f, b = 0.0f;
        for (int c = threadIdx.x; c < L.C; c += kLossThreads) {
            a += __ldcg(p.partial + 2 * (L.partial_begin + c));
            b += __ldcg(p.partial + 2 * (L.partial_begin + c) + 1);
        }
        a = block_sum1(a, red32);
        b = block_sum1(b, red32);
        const float fg = p.fg_kind ? a / (float)p.n_fg / (float)L.C : 0.0f;
        const float bg = p.bg_kind == 2 ? b / (float)p.n_bg_common / (float)L.C : (p.bg_kind == 1 ? b / (float)L.C : 0.0f);
        if (threadIdx.x == 0) { p.loss_out[1 + 2 * l] = fg; p.loss_out[2 + 2 * l] = bg; }
        if (p.fg_kind) total += L.fgw * fg;
        if (p.bg_kind) total += L.bgw * bg;
    }
    if (threadIdx.x == 0) {
        p.loss_out[0] = total;
        p.counters[0] = 0u;       // every group has drawn its last item: the queue can be re-armed for the next launch
        p.counters[1] = 0u;
    }
}

__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// snake order of the slices over the compute warps: rounds of kLossWarps slices, every other round reversed, so that the
// descending slice widths add up to about the same per warp
__device__ __forceinline__ int slice_of(int round, int wid) {
    return round * kLossWarps + ((round & 1) ? kLossWarps - 1 - wid : wid);
}

// ---- layers smaller than the loss grid -----------------------------------------------------------------
// Per-layer tables, built once per (plan, layer shape) by loss_resize_setup_kernel in global memory and copied to
// shared memory by the CTAs of loss_fused_kernel:
//   bilinear taps of every up row / column, the up rows (columns) that touch each native row (column) with their
//   weights (the transposed resize in gather form), up^T(background multiplicities) at native resolution, and the
//   native box that the active up box touches.
struct LayerTabHead {      // the part every CTA keeps in shared memory as is
    int ty0[kMaxG], ty1[kMaxG], tx0[kMaxG], tx1[kMaxG];
    float tly[kMaxG], tlx[kMaxG];
    int ylo[kMaxNative], yhi[kMaxNative], xlo[kMaxNative], xhi[kMaxNative];
    int box[8];                        // up space: r0, r1, s0, s1; native: y0, y1, x0, x1
    int win, pad_[3];                  // longest window of up rows (columns) that touch one native row (column)
};
struct LayerTabSmall : LayerTabHead {
    float wrow[kMaxNative * kWin], wcol[kMaxNative * kWin];   // stride kWin here, stride `win` in shared memory
};
struct LayerTab : LayerTabSmall {
    float wo[kMaxNative * kMaxNative], wt[kMaxNative * kMaxNative];
};
static_assert(kRowUnroll == 4, "the plan pads rows to four entries");
static_assert(sizeof(LayerTabHead) % 16 == 0 && sizeof(LayerTabSmall) % 16 == 0 && sizeof(LayerTab) % 16 == 0, "tables are copied with 128-bit accesses");

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
    if (tid == 0) {
        int win = 1;
        for (int i = 0; i < h; ++i) win = max(win, T.yhi[i] - T.ylo[i] + 1);
        for (int j = 0; j < w; ++j) win = max(win, T.xhi[j] - T.xlo[j] + 1);
        T.win = win;
    }
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

struct FusedShared {
    float red[kLossWarps][8];
    float red32[32];
    float lconst[kMaxLossLayers][4];
    unsigned int ticket;
    int item, layer, c0, planes;      // the item whose planes are (being) staged, decoded by thread 0
    uint64_t full;                    // mbarrier: the stage holds `item`
};

// sum over the CTA of kN values per thread, results in every thread; fixed order (warp tree, then a tree over the warp
// partials that every warp evaluates identically) -> deterministic.  One barrier.
__device__ __forceinline__ void group_sync(int gid) { asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(kLossThreads) : "memory"); }

template <int kN>
__device__ __forceinline__ void block_sum(float (&v)[kN], float (*red)[8], int gid) {
    static_assert(kN <= 8, "red holds 8 values per warp");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < kN; ++i) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], o);
    if (lane_id() == 0)
#pragma unroll
        for (int i = 0; i < kN; ++i) red[warp_id() & (kLossWarps - 1)][i] = v[i];
    group_sync(gid);
#pragma unroll
    for (int i = 0; i < kN; ++i) v[i] = red[lane_id() & (kLossWarps - 1)][i];
#pragma unroll
    for (int o = kLossWarps / 2; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < kN; ++i) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], o);
}

// One pair: acc += m |d|, cnt -= m sign(d), with t = copysign(m, d): m |d| = t d, and t is only counted when d != 0.
// The counts are integer valued floats (|sum| < 2^24, checked on the host): every partial sum is exact, so the order of
// the additions does not matter and the gradient stays bit-reproducible.
__device__ __forceinline__ void pair_term(float fm, float d, float& acc, float& cnt) {
    const float t = copysignf(fm, d);
    acc = fmaf(t, d, acc);
    if (d != 0.0f) cnt -= t;
}

// kG = 64: the loss grid of the reference; kG = 0: any grid <= 64.
// kBinary: every background multiplicity is 0 or 1 (lists from np.nonzero) -> register bit masks for the own cells.
// One CTA per SM, made of up to kMaxGroups independent GROUPS of 256 threads.  A group is what a small CTA would be - its own
// stage, scratch area, mbarrier and named barrier - but all groups of the SM share one copy of the sliced-ELL plan in shared
// memory (the pair entries are re-read for every plane; from L1 / L2 they were the dominant stall).
// kG = 64: the loss grid of the reference; kG = 0: any grid <= 64.
// kBinary: every background multiplicity is 0 or 1 (lists from np.nonzero) -> register bit masks for the own cells.
template <int kG, bool kBinary>
__global__ void __launch_bounds__(kLossThreads * kMaxGroups, 1) loss_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(128) float fsm[];
    __shared__ __align__(16) FusedShared shg[kMaxGroups];
    const int gid = threadIdx.x / kLossThreads, tid = threadIdx.x % kLossThreads, lane = lane_id(), wid = tid >> 5;
    FusedShared& sh = shg[gid];
#ifdef DH_LOSS_PHASE_TIMERS
    const long long dbg_clock0 = clock64();
#endif
    float* const gbase = fsm + p.ell_floats + (size_t)gid * (kStageFloats + p.scratch_floats);
    float* const st_cur = gbase;                      // the stage: [cur planes][orig planes]
    float* const st_org = gbase + kPlaneCap;
    float* const scratch = gbase + kStageFloats;
    const int n_items = p.n_flat_items + p.n_small_items;

    // Thread 0 is also the producer: it draws the next item from the queue while the current one is processed and, as soon
    // as every warp is done with the stage, issues the TMA bulk copies of the next item's planes (cp.async.bulk, completion
    // on the `full` mbarrier).  Several CTAs per SM keep HBM busy while one of them waits.
    auto stage_item = [&](int item) {
        sh.item = item;
        if (item >= n_items) {          // sentinel: the queue is empty
            mbar_arrive(&sh.full);
            return;
        }
        const WorkItem it = decode_item(p, item);
        const FusedLayer& L = p.lv[it.layer];
        sh.layer = it.layer; sh.c0 = it.c0; sh.planes = it.planes;
        const size_t off = (size_t)it.c0 * L.h * L.w;
        const uint32_t bytes = (uint32_t)(it.planes * L.h * L.w) * 4u;
        mbar_arrive_expect_tx(&sh.full, 2u * bytes);
        tma_bulk_g2s(st_cur, L.cur + off, bytes, &sh.full);
        tma_bulk_g2s(st_org, L.orig + off, bytes, &sh.full);
    };
    if (tid == 0) {
        mbar_init(&sh.full, 1);
        mbar_fence_init();
        stage_item((int)atomicAdd(p.counters, 1u));
    }

    const int G = kG ? kG : p.G, GG = G * G;
    const PlanView& pv = p.pv;
    const int n_slices = p.fg_kind ? (p.ell_slices ? p.ell_slices : pv.hdr->n_slices) : 0;
    const int n_rounds = (n_slices + kLossWarps - 1) / kLossWarps;
    // the sliced-ELL plan: one copy in shared memory for all groups when it fits (ell_floats != 0), else read through L1
    const int32_t* ell_off = pv.ell_off;
    const uint32_t* row_desc = pv.row_desc;
    const uint32_t* ent = pv.ent;
    if (p.ell_floats) {
        const int n_groups32 = p.fg_kind ? p.ell_ent_cap : 0;       // (the caller's ell_groups * 32)
        int32_t* s_off = reinterpret_cast<int32_t*>(fsm);
        uint32_t* s_desc = reinterpret_cast<uint32_t*>(fsm) + p.ell_desc_at;
        uint32_t* s_ent = reinterpret_cast<uint32_t*>(fsm) + p.ell_ent_at;
        {
            for (int i = threadIdx.x; i <= n_slices; i += blockDim.x) s_off[i] = pv.ell_off[i];
            for (int i = threadIdx.x; i < n_slices * 32; i += blockDim.x) s_desc[i] = pv.row_desc[i];
            for (int i = threadIdx.x; i < n_groups32 / 4; i += blockDim.x)
                reinterpret_cast<uint4*>(s_ent)[i] = reinterpret_cast<const uint4*>(pv.ent)[i];
            ell_off = s_off; row_desc = s_desc; ent = s_ent;
        }
    }
    // membership of the own cells (flat layers) in the three background lists and in the set of destination rows, as
    // 16-bit masks (bit 4k+i = cell i of group k), precomputed by the plan.  Lists that come from np.nonzero never repeat a
    // cell; if a generic caller does, the multiplicities are re-read from the plan in the (slower) general path.
    const uint2 own = pv.own_masks[tid];
    const uint32_t m_ot = own.x;                                   // low half: bg_orig, high half: bg_trans
    const uint32_t m_cr = p.fg_kind ? own.y : (own.y & 0xFFFFu);   // low half: bg_common, high half: destination rows
    // multiplicity of own cell (k, i) in list `which` (0 = bg_orig, 1 = bg_trans, 2 = bg_common)
    auto wgt = [&](int k, int i, int which) -> float {
        if (kBinary) return (float)(((which == 2 ? m_cr : m_ot) >> ((which == 1 ? 16 : 0) + 4 * k + i)) & 1u);
        const ushort4 bc = pv.bgcnt[4 * (tid + k * kLossThreads) + i];
        return (float)(which == 0 ? bc.x : which == 1 ? bc.y : bc.z);
    };
    if (tid < p.n_layers) {
        const FusedLayer& L = p.lv[tid];
        // an empty index list makes the reference's loss NaN (mean over nothing) but its gradient ZERO: scale 0, not inf
        sh.lconst[tid][0] = p.fg_kind && p.n_fg > 0 ? L.fgw / ((float)L.C * (float)p.n_fg) : 0.0f;
        sh.lconst[tid][1] = p.bg_kind == 2 && p.n_bg_common > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_common) : 0.0f;
        sh.lconst[tid][2] = p.bg_kind == 1 && p.n_bg_trans > 0 ? L.bgw / ((float)L.C * (float)p.n_bg_trans) : 0.0f;
    }
    const float inv_no = 1.0f / (float)p.n_bg_orig, inv_nt = 1.0f / (float)p.n_bg_trans;

    // resized layers: scratch views and the active box (identical in every layer's table).  Two planes are processed at a
    // time (float2 per cell), so that the index arithmetic of the resize and of the pair walk is paid once per pair of planes.
    const ResizeLayout& lay = p.lay;
    LayerTabHead& T = *reinterpret_cast<LayerTabHead*>(scratch + lay.tab);
    float* const swrow = scratch + lay.wrow;       // [h][win], [w][win]
    float* const swcol = scratch + lay.wcol;
    float2* const suc = reinterpret_cast<float2*>(scratch + lay.uc);     // box-local: index (r - br0) * bw + (s - bs0)
    float2* const suo = reinterpret_cast<float2*>(scratch + lay.uo);
    float2* const gu = reinterpret_cast<float2*>(scratch + lay.cnt);     // sign counts, then the gradient w.r.t. up(cur)
    float2* const tmp = reinterpret_cast<float2*>(scratch + lay.tmp);    // (overlays uc / uo, which are dead by then)
    float* const cntb = scratch + lay.flat_cnt;                          // sign counts of a flat plane
    int br0 = 0, br1 = -1, bs0 = 0, bs1 = -1;
    if (p.n_small_items) {
        const LayerTab* tab0 = nullptr;
#pragma unroll
        for (int i = kMaxLossLayers - 1; i >= 0; --i)
            if (i < p.n_layers && !p.lv[i].flat) tab0 = static_cast<const LayerTab*>(p.lv[i].tab);
        br0 = tab0->box[0]; br1 = tab0->box[1]; bs0 = tab0->box[2]; bs1 = tab0->box[3];
    }
    const bool any_box = br1 >= br0;
    const int bw = any_box ? bs1 - bs0 + 1 : 0, bh = any_box ? br1 - br0 + 1 : 0;
    const int bcells = bw * bh;
    const bool box_overflow = bcells > lay.box_cap;   // the caller under-sized the box-local buffers: poison, never corrupt
    const int boff = br0 * bw + bs0;
    const unsigned bw_magic = bw > 1 ? (unsigned)((0x100000000ull + bw - 1) / bw) : 0xFFFFFFFFu;     // floor(i / bw) = umulhi(i, magic), i < 2^20
    auto to_box = [&](int cell) { const int r = kG ? cell >> 6 : cell / G; return r * bw + (cell - r * G) - boff; };
    __syncthreads();

    int cur_layer = -1;          // resized layer whose tables are in shared memory
    uint32_t phase = 0;
    unsigned long long dbg_t0 = 0, dbg_items = 0;
    if (p.debug && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_t0));
#ifdef DH_LOSS_PHASE_TIMERS
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ph_t = dbg_clock0;
    { const long long t_ = clock64(); ph[6] = t_ - ph_t; ph_t = t_; }
#define DH_PH(i) { const long long t_ = clock64(); ph[i] += t_ - ph_t; ph_t = t_; }
#else
#define DH_PH(i)
#endif
    for (;;) {
        int next_item = 0;
        if (tid == 0) next_item = (int)atomicAdd(p.counters, 1u);     // (its latency hides behind this item's work)
        mbar_wait(&sh.full, phase);
        DH_PH(0)
        phase ^= 1;
        if (sh.item >= n_items) break;
        ++dbg_items;
        const int l = sh.layer, c0 = sh.c0, planes = sh.planes;
        const FusedLayer& L = p.lv[l];
        const float fscale = sh.lconst[l][0], lscale = sh.lconst[l][1], gscale = sh.lconst[l][2];
        if (L.flat) {
            // ------------------------------------------------------------------ a plane at the loss-grid resolution
            const int c = c0;
            // Own cells: four 128-bit groups per thread, consumed right away - background sums, and for the local
            // background term the sign of (orig - cur) as two bits per cell.
            float sums[3] = {0.0f, 0.0f, 0.0f};          // foreground sum, two background sums
            uint32_t sign_bits = 0u;          // bit 4k+i: orig > cur, bit 16+4k+i: orig < cur at own cell i of group k
            if (p.bg_kind) {
#pragma unroll
                for (int k = 0; k < kOwnGroups; ++k) {
                    const int q = 4 * (tid + k * kLossThreads);
                    if (q < GG) {
                        const float4 c4 = *reinterpret_cast<const float4*>(st_cur + q);
                        const float4 o4 = *reinterpret_cast<const float4*>(st_org + q);
                        const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
                        if (p.bg_kind == 1) {
#pragma unroll
                            for (int i = 3; i >= 0; --i) {
                                if (kBinary) {      // (adding x is fmaf(1, x, s); skipping it is fmaf(0, x, s) for finite x)
                                    if ((m_ot >> (4 * k + i)) & 1u) sums[1] += ov[i];
                                    if ((m_ot >> (16 + 4 * k + i)) & 1u) sums[2] += cv[i];
                                } else {
                                    sums[1] = fmaf(wgt(k, i, 0), ov[i], sums[1]);
                                    sums[2] = fmaf(wgt(k, i, 1), cv[i], sums[2]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 3; i >= 0; --i) {
                                const float d = ov[i] - cv[i];
                                sums[1] = fmaf(wgt(k, i, 2), fabsf(d), sums[1]);
                                sign_bits |= (d > 0.0f ? 1u : 0u) << (4 * k + i);
                                sign_bits |= (d < 0.0f ? 1u : 0u) << (16 + 4 * k + i);
                            }
                        }
                    }
                }
            }
            // the scratch area may still be read by slow warps (sign counts of the previous plane, a resized plane's buffers)
            group_sync(gid);
            // Destination rows: one thread per destination cell, its sources from the sliced-ELL plan
            for (int rd = 0; rd < n_rounds; ++rd) {
                const int sl = slice_of(rd, wid);
                if (sl >= n_slices) break;
                const uint32_t desc = row_desc[sl * 32 + lane];
                const int len = (int)(desc >> 16), dcell = (int)(desc & 0xFFFFu);
                const uint32_t* e_ptr = ent + ell_off[sl] * 32 + lane;
                const float cval = st_cur[dcell];
                float cn = 0.0f;
                for (int k = 0; k < len; k += kRowUnroll) {
                    uint32_t e[kRowUnroll];
                    float o[kRowUnroll];
#pragma unroll
                    for (int u = 0; u < kRowUnroll; ++u) e[u] = e_ptr[(k + u) * 32];      // (rows are padded to a multiple of kRowUnroll)
#pragma unroll
                    for (int u = 0; u < kRowUnroll; ++u) o[u] = st_org[e[u] & 0xFFFu];
#pragma unroll
                    for (int u = 0; u < kRowUnroll; ++u) pair_term((float)(e[u] >> 12), o[u] - cval, sums[0], cn);   // (m = 0: no-op)
                }
                if (len) cntb[dcell] = cn;
            }
            block_sum<3>(sums, sh.red, gid);       // (its barrier also orders the count stores and ends the reads of the stage)
            if (tid == 0) stage_item(next_item);
            float bg_term = 0.0f, bscale = 0.0f;
            if (p.bg_kind == 1) {
                const float delta = sums[1] * inv_no - sums[2] * inv_nt;
                bg_term = fabsf(delta);
                bscale = -sgn(delta) * gscale;
            } else if (p.bg_kind == 2) {
                bg_term = sums[1];
            }
            if (tid == 0) { p.partial[2 * (L.partial_begin + c)] = sums[0]; p.partial[2 * (L.partial_begin + c) + 1] = bg_term; }
            if (L.grad) {
                float* g = L.grad + (size_t)c * GG;
#pragma unroll
                for (int k = 0; k < kOwnGroups; ++k) {
                    const int q = 4 * (tid + k * kLossThreads);
                    if (q >= GG) break;
                    const float4 ci = *reinterpret_cast<const float4*>(cntb + q);     // cells that are no row hold stale data: masked
                    const float cv[4] = {ci.x, ci.y, ci.z, ci.w};
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = ((m_cr >> (16 + 4 * k + i)) & 1u) ? cv[i] * fscale : 0.0f;
                    if (p.bg_kind == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (kBinary) { if ((m_ot >> (16 + 4 * k + i)) & 1u) v[i] += bscale; }
                            else v[i] = fmaf(wgt(k, i, 1), bscale, v[i]);
                        }
                    } else if (p.bg_kind == 2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float sg = (float)((sign_bits >> (4 * k + i)) & 1u) - (float)((sign_bits >> (16 + 4 * k + i)) & 1u);
                            v[i] -= sg * wgt(k, i, 2) * lscale;
                        }
                    }
                    st_cs_f4(g + q, make_float4(v[0], v[1], v[2], v[3]));
                }
            }
            DH_PH(1)
        } else {
            // ------------------------------------------------------------------ planes of a layer below the loss grid
            const int h = L.h, w = L.w, hw = h * w;
            const LayerTab* const tl = static_cast<const LayerTab*>(L.tab);
            const float* __restrict__ gwo = tl->wo;      // up^T(background multiplicities), read through L1
            const float* __restrict__ gwt = tl->wt;
            if (l != cur_layer) {       // (a CTA crosses a layer boundary a handful of times)
                group_sync(gid);        // slow warps may still read the previous layer's tables
                const float4* src = reinterpret_cast<const float4*>(static_cast<const LayerTabHead*>(tl));
                float4* dst = reinterpret_cast<float4*>(&T);
                for (int i = tid; i < (int)(sizeof(LayerTabHead) / 16); i += kLossThreads) dst[i] = src[i];
                const int win = tl->win;
                for (int i = tid; i < h * win; i += kLossThreads) swrow[i] = tl->wrow[(i / win) * kWin + i % win];
                for (int i = tid; i < w * win; i += kLossThreads) swcol[i] = tl->wcol[(i / win) * kWin + i % win];
                cur_layer = l;
            }
            for (int pl = 0; pl < planes; pl += 2) {
                const int c = c0 + pl;
                const bool two = pl + 1 < planes;          // an odd plane at the end is paired with itself, its copy is not stored
                const float* const pc0 = st_cur + pl * hw;
                const float* const po0 = st_org + pl * hw;
                const float* const pc1 = two ? pc0 + hw : pc0;
                const float* const po1 = two ? po0 + hw : po0;
                const bool last_pair = pl + 2 >= planes;
                group_sync(gid);          // tables loaded; scratch of the previous pair / flat item no longer read
                if (box_overflow) {
                    if (tid == 0) {
                        for (int j = 0; j < (two ? 2 : 1); ++j) {
                            p.partial[2 * (L.partial_begin + c + j)] = __int_as_float(0x7FC00000);
                            p.partial[2 * (L.partial_begin + c + j) + 1] = __int_as_float(0x7FC00000);
                        }
                        if (last_pair) stage_item(next_item);
                    }
                    continue;
                }
                const int win = T.win;
                const int ny0 = T.box[4], ny1 = T.box[5], nx0 = T.box[6], nx1 = T.box[7];
                // up(cur), up(orig) inside the box (and the box-local sign counts start at zero)
                for (int i = tid; i < bcells; i += kLossThreads) gu[i] = make_float2(0.0f, 0.0f);
                for (int b = tid; b < bcells; b += kLossThreads) {        // (flat over the box: every lane busy)
                    const int rr = (int)__umulhi((unsigned)b, bw_magic), r = br0 + rr, s = bs0 + b - rr * bw;
                    const int y0 = T.ty0[r] * w, y1 = T.ty1[r] * w;
                    const float ly = T.tly[r], hy = 1.0f - ly;
                    const int x0 = T.tx0[s], x1 = T.tx1[s];
                    const float lx = T.tlx[s], hx = 1.0f - lx;
                    const int i00 = y0 + x0, i01 = y0 + x1, i10 = y1 + x0, i11 = y1 + x1;
                    suc[b] = make_float2(hy * (hx * pc0[i00] + lx * pc0[i01]) + ly * (hx * pc0[i10] + lx * pc0[i11]),
                                         hy * (hx * pc1[i00] + lx * pc1[i01]) + ly * (hx * pc1[i10] + lx * pc1[i11]));
                    suo[b] = make_float2(hy * (hx * po0[i00] + lx * po0[i01]) + ly * (hx * po0[i10] + lx * po0[i11]),
                                         hy * (hx * po1[i00] + lx * po1[i01]) + ly * (hx * po1[i10] + lx * po1[i11]));
                }
                group_sync(gid);
                DH_PH(2)
                float sums[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};     // per plane: foreground sum, two background sums
                for (int rd = 0; rd < n_rounds; ++rd) {
                    const int sl = slice_of(rd, wid);
                    if (sl >= n_slices) break;
                    const uint32_t desc = row_desc[sl * 32 + lane];
                    const int len = (int)(desc >> 16), db = len ? to_box((int)(desc & 0xFFFFu)) : 0;
                    const uint32_t* e_ptr = ent + ell_off[sl] * 32 + lane;
                    const float2 cval = suc[db];
                    float cn0 = 0.0f, cn1 = 0.0f;
                    for (int k = 0; k < len; k += kRowUnroll) {
                        uint32_t e[kRowUnroll];
                        float2 o[kRowUnroll];
#pragma unroll
                        for (int u = 0; u < kRowUnroll; ++u) e[u] = e_ptr[(k + u) * 32];      // (rows are padded to a multiple of kRowUnroll)
#pragma unroll
                        for (int u = 0; u < kRowUnroll; ++u) o[u] = suo[to_box((int)(e[u] & 0xFFFu))];
#pragma unroll
                        for (int u = 0; u < kRowUnroll; ++u) {
                            const float fm = (float)(e[u] >> 12);          // (m = 0 past the end of the row: no-op)
                            pair_term(fm, o[u].x - cval.x, sums[0], cn0);
                            pair_term(fm, o[u].y - cval.y, sums[3], cn1);
                        }
                    }
                    if (len) gu[db] = make_float2(cn0, cn1);
                }
                if (p.bg_kind == 1) {     // background sums at native resolution: <wo, orig>, <wt, cur>
                    for (int i = tid * 4; i < hw; i += kLossThreads * 4) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(gwo + i)), e = __ldg(reinterpret_cast<const float4*>(gwt + i));
                        const float4 b0 = *reinterpret_cast<const float4*>(po0 + i), f0 = *reinterpret_cast<const float4*>(pc0 + i);
                        const float4 b1 = *reinterpret_cast<const float4*>(po1 + i), f1 = *reinterpret_cast<const float4*>(pc1 + i);
                        sums[1] = fmaf(a.x, b0.x, fmaf(a.y, b0.y, fmaf(a.z, b0.z, fmaf(a.w, b0.w, sums[1]))));
                        sums[2] = fmaf(e.x, f0.x, fmaf(e.y, f0.y, fmaf(e.z, f0.z, fmaf(e.w, f0.w, sums[2]))));
                        sums[4] = fmaf(a.x, b1.x, fmaf(a.y, b1.y, fmaf(a.z, b1.z, fmaf(a.w, b1.w, sums[4]))));
                        sums[5] = fmaf(e.x, f1.x, fmaf(e.y, f1.y, fmaf(e.z, f1.z, fmaf(e.w, f1.w, sums[5]))));
                    }
                } else if (p.bg_kind == 2) {       // box = whole grid
                    for (int q = tid; q < GG; q += kLossThreads) {
                        const float mz = (float)pv.bgcnt[q].z;
                        const float2 uo = suo[q], uc = suc[q];
                        sums[1] = fmaf(mz, fabsf(uo.x - uc.x), sums[1]);
                        sums[4] = fmaf(mz, fabsf(uo.y - uc.y), sums[4]);
                    }
                }
                DH_PH(3)
                block_sum<6>(sums, sh.red, gid);
                DH_PH(4)
                if (last_pair && tid == 0) stage_item(next_item);      // every read of the staged planes is behind the barrier
                float bscale0 = 0.0f, bscale1 = 0.0f;
                {
                    float bg0 = 0.0f, bg1 = 0.0f;
                    if (p.bg_kind == 1) {
                        const float d0 = sums[1] * inv_no - sums[2] * inv_nt, d1 = sums[4] * inv_no - sums[5] * inv_nt;
                        bg0 = fabsf(d0); bg1 = fabsf(d1);
                        bscale0 = -sgn(d0) * gscale; bscale1 = -sgn(d1) * gscale;
                    } else if (p.bg_kind == 2) {
                        bg0 = sums[1]; bg1 = sums[4];
                    }
                    if (tid == 0) {
                        p.partial[2 * (L.partial_begin + c)] = sums[0]; p.partial[2 * (L.partial_begin + c) + 1] = bg0;
                        if (two) { p.partial[2 * (L.partial_begin + c + 1)] = sums[3]; p.partial[2 * (L.partial_begin + c + 1) + 1] = bg1; }
                    }
                }
                if (L.grad) {
                    float* g0 = L.grad + (size_t)c * hw;
                    // gradient w.r.t. up(cur) inside the box (in place: sign count -> gradient), then up^T in gather form
                    for (int i = tid; i < bcells; i += kLossThreads) {
                        float2 v = gu[i];
                        v.x *= fscale; v.y *= fscale;
                        if (p.bg_kind == 2) {
                            const float mz = (float)pv.bgcnt[i].z * lscale;
                            const float2 uo = suo[i], uc = suc[i];
                            v.x -= sgn(uo.x - uc.x) * mz; v.y -= sgn(uo.y - uc.y) * mz;
                        }
                        gu[i] = v;
                    }
                    group_sync(gid);
                    for (int i = tid; i < (ny1 - ny0 + 1) * bw; i += kLossThreads) {       // (flat over native rows x box columns)
                        const int yr = (int)__umulhi((unsigned)i, bw_magic), yi = ny0 + yr, sc = i - yr * bw;
                        const int lo = T.ylo[yi];
                        const int ra = max(lo, br0), rb = min(T.yhi[yi], br1);
                        const float* wr = swrow + yi * win - lo;
                        float a0 = 0.0f, a1 = 0.0f;
                        const float2* gp = gu + sc - br0 * bw;
                        for (int r = ra; r <= rb; ++r) {
                            const float2 gv = gp[r * bw];
                            a0 = fmaf(wr[r], gv.x, a0); a1 = fmaf(wr[r], gv.y, a1);
                        }
                        tmp[yi * G + bs0 + sc] = make_float2(a0, a1);
                    }
                    group_sync(gid);
                    DH_PH(5)
                    for (int yi = wid; yi < h; yi += kLossWarps) {
                        const bool row_in = yi >= ny0 && yi <= ny1;
                        for (int xj = lane; xj < w; xj += 32) {
                            float a0 = 0.0f, a1 = 0.0f;
                            if (row_in && xj >= nx0 && xj <= nx1) {
                                const int lo = T.xlo[xj];
                                const int sa = max(lo, bs0), sb = min(T.xhi[xj], bs1);
                                const float* wc = swcol + xj * win - lo;
                                const float2* tp = tmp + yi * G;
                                for (int s = sa; s <= sb; ++s) {
                                    const float2 tv = tp[s];
                                    a0 = fmaf(wc[s], tv.x, a0); a1 = fmaf(wc[s], tv.y, a1);
                                }
                            }
                            if (p.bg_kind == 1) {
                                const float wtv = __ldg(gwt + yi * w + xj);
                                a0 = fmaf(bscale0, wtv, a0); a1 = fmaf(bscale1, wtv, a1);
                            }
                            g0[yi * w + xj] = a0;
                            if (two) g0[hw + yi * w + xj] = a1;
                        }
                    }
                }
            }
            DH_PH(7)
        }
    }
    if (p.debug && tid == 0) {
        unsigned long long t1;
        unsigned int smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        const int slot = blockIdx.x * kMaxGroups + gid;
        p.debug[4 * slot + 0] = dbg_t0; p.debug[4 * slot + 1] = t1;
        p.debug[4 * slot + 2] = dbg_items; p.debug[4 * slot + 3] = smid;
#ifdef DH_LOSS_PHASE_TIMERS
        for (int i = 0; i < 8; ++i) p.debug[4 * 1024 + 8 * slot + i] = (unsigned long long)ph[i];
#endif
    }
    loss_finish(p, shg[0].red32, &shg[0].ticket);
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
    return plan_layout(grid, n_fg).total;
}

size_t dh_loss_plan_workspace_bytes(int grid, int n_fg) {
    if (grid < 1 || grid > kMaxG || n_fg < 0) return 0;
    return sizeof(int32_t) * 2 * (size_t)(n_fg > 0 ? n_fg : 1);
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
    size_t smem_ints = (size_t)3 * cells + 1 + (size_t)kPlanWarps * kPlanTab;
    if (smem_ints < (size_t)4 * cells) smem_ints = (size_t)4 * cells;
    const size_t smem = sizeof(int) * smem_ints;
    static_assert(kPlanWarps * kPlanTab >= 2 * kPlanWarps * kLenClasses, "the ELL histograms live in the counter tables");
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
    SetupParams sp;
    sp.plan = plan; sp.plan_cap = n_fg; sp.G = grid; sp.h = h; sp.w = w; sp.fg_kind = fg_kind; sp.bg_kind = bg_kind;
    loss_resize_setup_kernel<<<1, 256, 0, as_stream(stream)>>>(sp, static_cast<LayerTab*>(tables));
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_loss_plan_info(const void* plan_header_host, int* n_pairs, int* box_cells, int* plan_flags, int* ell_slices, int* ell_groups) {
    DH_REQUIRE(plan_header_host);
    const PlanHeader* h = static_cast<const PlanHeader*>(plan_header_host);
    if (n_pairs) *n_pairs = h->n_pairs;
    if (ell_slices) *ell_slices = h->n_slices;
    if (ell_groups) *ell_groups = h->n_groups;
    if (plan_flags) *plan_flags = h->reserved & 1;
    if (box_cells) *box_cells = h->box_r1 >= h->box_r0 ? (h->box_r1 - h->box_r0 + 1) * (h->box_s1 - h->box_s0 + 1) : 0;
    return DH_OK;
}

int dh_guidance_loss(const dh_loss_layer* layers_host, int n_layers, int grid, const void* plan, int n_fg, int n_bg_orig,
                     int n_bg_trans, int n_bg_common, int box_cells, int plan_flags, int ell_slices, int ell_groups, int fg_kind,
                     int bg_kind, float* loss_out, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(layers_host && n_layers >= 1 && n_layers <= kMaxLossLayers && loss_out && ws && plan);
    DH_REQUIRE(grid >= 1 && grid <= kMaxG);
    DH_REQUIRE(n_fg >= 0 && n_bg_orig >= 0 && n_bg_trans >= 0 && n_bg_common >= 0);
    if (bg_kind != 0 && bg_kind != 1 && bg_kind != 2) return DH_ERR_INVALID_ARGUMENT;
    if (fg_kind != 0 && fg_kind != 1) return DH_ERR_INVALID_ARGUMENT;
    FusedParams fp;
    memset(&fp, 0, sizeof(fp));
    const int GG = grid * grid;
    int chan = 0, h_r = 0, wrow_floats = 0, wcol_floats = 0;
    for (int i = 0; i < n_layers; ++i) {
        const dh_loss_layer& s = layers_host[i];
        DH_REQUIRE(s.cur && s.orig && s.channels >= 1 && s.h >= 1 && s.w >= 1);
        if (s.h > kMaxNative || s.w > kMaxNative || s.h > grid || s.w > grid) return DH_ERR_UNSUPPORTED;
        // the transposed-resize tables hold kWin up rows (columns) per native row (column): 2 * ceil(grid / h) of them are needed
        if (2 * ((grid + s.h - 1) / s.h) > kWin || 2 * ((grid + s.w - 1) / s.w) > kWin) return DH_ERR_UNSUPPORTED;
        if (((size_t)s.h * s.w) % 4 != 0) return DH_ERR_UNSUPPORTED;     // 16-byte bulk copies, 128-bit plane accesses
        if ((reinterpret_cast<uintptr_t>(s.cur) & 15) || (reinterpret_cast<uintptr_t>(s.orig) & 15) ||
            (s.grad && (reinterpret_cast<uintptr_t>(s.grad) & 15)))
            return DH_ERR_INVALID_ARGUMENT;
        const bool flat = s.h == grid && s.w == grid;
        if (!flat && (!s.resize_tables || (reinterpret_cast<uintptr_t>(s.resize_tables) & 15))) return DH_ERR_INVALID_ARGUMENT;
        FusedLayer& L = fp.lv[i];
        L.cur = s.cur; L.orig = s.orig; L.grad = s.grad;
        L.C = s.channels; L.h = s.h; L.w = s.w; L.fgw = s.fg_weight; L.bgw = s.bg_weight;
        L.tab = flat ? nullptr : s.resize_tables;
        L.flat = flat ? 1 : 0;
        L.partial_begin = chan;
        chan += s.channels;
        if (flat) {
            L.ppi = 1;
            L.item_begin = fp.n_flat_items;
            fp.n_flat_items += s.channels;
        } else {
            int ppi = kPlaneCap / (s.h * s.w);
            if (ppi > kMaxPlanesPerItem) ppi = kMaxPlanesPerItem;
            if (ppi < 1) ppi = 1;
            L.ppi = ppi;
            L.item_begin = fp.n_small_items;
            fp.n_small_items += (s.channels + ppi - 1) / ppi;
            if (s.h > h_r) h_r = s.h;
            // rows of the transposed-resize tables in shared memory: at most 2 * ceil(grid / h) up rows touch one native row
            const int wy = 2 * ((grid + s.h - 1) / s.h) < kWin ? 2 * ((grid + s.h - 1) / s.h) : kWin;
            const int wx = 2 * ((grid + s.w - 1) / s.w) < kWin ? 2 * ((grid + s.w - 1) / s.w) : kWin;
            const int wmax = wy > wx ? wy : wx;       // the tables share one stride
            if (s.h * wmax > wrow_floats) wrow_floats = s.h * wmax;
            if (s.w * wmax > wcol_floats) wcol_floats = s.w * wmax;
        }
    }
    fp.n_layers = n_layers;
    fp.G = grid;
    size_t o_counter;
    if (ws_bytes < loss_ws_layout((size_t)chan, &o_counter)) return DH_ERR_WORKSPACE;
    fp.pv = plan_view(const_cast<void*>(plan), grid, n_fg);
    if (n_fg >= (1 << 24)) return DH_ERR_UNSUPPORTED;      // sign counts are integer valued floats
    fp.n_fg = n_fg; fp.n_bg_orig = n_bg_orig; fp.n_bg_trans = n_bg_trans; fp.n_bg_common = n_bg_common;
    fp.fg_kind = fg_kind; fp.bg_kind = bg_kind;
    fp.partial = static_cast<float*>(ws);
    fp.counters = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + o_counter);
    fp.loss_out = loss_out;
    {
        const char* e = getenv("DH_LOSS_DEBUG_BUF");      // developer aid: address of a device buffer of 4 * grid u64
        fp.debug = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr;
    }
    // scratch: the tables of the current resized layer, then the per-plane buffers of a resized plane overlaid with the
    // sign-count buffer of a flat plane
    auto up4 = [](int v) { return (v + 3) / 4 * 4; };
    ResizeLayout& lay = fp.lay;
    int o = 0;
    if (fp.n_small_items) {
        int cap = (box_cells > 0 && box_cells <= GG && bg_kind != 2) ? box_cells : GG;
        if (!fg_kind && bg_kind != 2) cap = 4;
        lay.tab = o;    o += (int)(sizeof(LayerTabHead) / 4);
        lay.wrow = o;   o += up4(wrow_floats);
        lay.wcol = o;   o += up4(wcol_floats);
        // two planes at a time: float2 per box cell.  tmp (2 * h * grid floats) is written when uc / uo are dead
        const int plane_area = 4 * up4(cap) > 2 * up4(h_r * grid) ? 4 * up4(cap) : 2 * up4(h_r * grid);
        lay.uc = o;     lay.uo = o + 2 * up4(cap);  lay.tmp = o;   o += plane_area;
        lay.cnt = o;    o += 2 * up4(cap);
        lay.box_cap = cap;
    }
    lay.flat_cnt = fp.n_small_items ? lay.uc : 0;
    if (fp.n_flat_items && lay.flat_cnt + GG > o) o = lay.flat_cnt + GG;
    lay.total = o;
    fp.scratch_floats = up4(o);
    cudaStream_t st = as_stream(stream);
    // per-process caches of the launch geometry queries (this entry point runs every denoising step)
    struct LaunchCache {
        int sms, max_smem;
        size_t smem_attr_set[4];
    };
    static LaunchCache caches[64];          // zero-initialised; one entry per device (function attributes are per device)
    int dev = 0;
    DH_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DH_ERR_UNSUPPORTED;
    LaunchCache& lc = caches[dev];
    if (!lc.sms) {
        DH_CUDA_CHECK(cudaDeviceGetAttribute(&lc.sms, cudaDevAttrMultiProcessorCount, dev));
        DH_CUDA_CHECK(cudaDeviceGetAttribute(&lc.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    // shared memory: [one copy of the sliced-ELL plan][per group: stage + scratch]; as many groups (<= 3) as fit, the plan copy
    // is dropped (entries then come through L1) before the group count goes below two
    const size_t group_bytes = sizeof(float) * ((size_t)kStageFloats + fp.scratch_floats);
    const size_t static_bytes = 2048;       // FusedShared x kMaxGroups and the driver's reservation, rounded up
    const size_t budget = (size_t)lc.max_smem > static_bytes ? (size_t)lc.max_smem - static_bytes : 0;
    int ell_words = 0;
    if (fg_kind && ell_slices > 0 && ell_groups > 0) {
        fp.ell_desc_at = up4(ell_slices + 1);
        fp.ell_ent_at = fp.ell_desc_at + ell_slices * 32;
        fp.ell_ent_cap = ell_groups * 32;
        fp.ell_slices = ell_slices;
        ell_words = up4(fp.ell_ent_at + fp.ell_ent_cap);
        ell_words = (ell_words + 31) / 32 * 32;        // the groups' stages stay 128-byte aligned
    }
    int groups = 0;
    if (ell_words && sizeof(float) * ell_words + 2 * group_bytes <= budget) {
        groups = sizeof(float) * ell_words + 3 * group_bytes <= budget ? 3 : 2;
        fp.ell_floats = ell_words;
    } else {
        fp.ell_floats = 0;
        for (groups = kMaxGroups; groups > 1 && groups * group_bytes > budget; --groups) {}
    }
    const int n_items = fp.n_flat_items + fp.n_small_items;
    {
        const char* e = getenv("DH_LOSS_GROUPS");       // developer knob
        if (e && atoi(e) >= 1 && atoi(e) < groups) groups = atoi(e);
    }
    while (groups > 1 && lc.sms * (groups - 1) >= n_items) --groups;      // tiny problems: no idle groups
    const size_t smem = sizeof(float) * fp.ell_floats + groups * group_bytes;
    if (smem > budget + static_bytes) return DH_ERR_UNSUPPORTED;
    // bit 0 of plan_flags: background multiplicities are all 0/1 (lists from np.nonzero) -> register bit masks
    const int vi = (grid == 64 ? 2 : 0) + ((plan_flags & 1) ? 1 : 0);
    void (*kernel)(const FusedParams) = vi == 3 ? loss_fused_kernel<64, true> : vi == 2 ? loss_fused_kernel<64, false>
                                        : vi == 1 ? loss_fused_kernel<0, true> : loss_fused_kernel<0, false>;
    if (smem > lc.smem_attr_set[vi]) {
        DH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lc.smem_attr_set[vi] = smem;
    }
    int grid_x = lc.sms;
    if (grid_x * groups > n_items) grid_x = (n_items + groups - 1) / groups;
    kernel<<<grid_x, kLossThreads * groups, smem, st>>>(fp);
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
