// Shared definitions of the K4 loss kernels: the per-edit loss plan (dh_build_loss_plan) and the bilinear taps.
#pragma once
#include "dh_common.cuh"

namespace dh {

constexpr int kMaxLossLayers = 8;
constexpr int kMaxG = 64;
constexpr int kMaxNative = 64;

struct PlanHeader {
    int32_t n_pairs, n_fg, n_bg_orig, n_bg_trans, n_bg_common, grid, cap, reserved;
    int32_t box_r0, box_r1, box_s0, box_s1;   // box (loss-grid rows / columns) of the cells that are a pair source or destination
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

// torch area_pixel_compute_source_index (align_corners = False) for output index i
__device__ __forceinline__ void bilinear_tap(int i, int n_in, float scale, int& i0, int& i1, float& lam) {
    float src = scale * ((float)i + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = src - (float)i0;
}

__device__ __forceinline__ float sgn(float d) { return d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f); }

}  // namespace dh
