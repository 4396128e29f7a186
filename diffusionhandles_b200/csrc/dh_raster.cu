// Hard z-buffer triangle rasteriser for the reference's mesh mode (SURVEY.md 8(a) row 11, 8(f) rank 2;
// depth_transform.py:91-195 through pytorch3d_renderer.py:541-941).
//
// The reference calls pytorch3d's MeshRasterizer (faces_per_pixel = 1, blur_radius 1e-5, perspective-correct and
// clipped barycentrics, back-face culling).  pytorch3d is not installed here and no reference test pins its output, so
// this kernel follows the published semantics of pytorch3d's rasterize_meshes - PARITY UNPINNED:
//   * vertices: view = X R + T, NDC x = sx X / Z, y = sy Y / Z, NDC z = view z  (+x left, +y up);
//   * pixel (row i, col j) samples NDC (x, y) = (PixToNdc(W-1-j), PixToNdc(H-1-i)), pixel centres at half integers;
//   * per pixel and face: skip if the face is behind the camera, back-facing (when culling), of ~zero area or the pixel
//     is outside the bounding box grown by sqrt(blur_radius); barycentrics -> perspective correction -> clip and
//     renormalise; pz = sum bary * z; skip pz < 0; accept if inside (unclipped barycentrics > 0) or the squared
//     distance to the triangle is below blur_radius;
//   * the winner of a pixel is the lexicographic minimum of (pz, face index): ONE 64-bit atomicMin on
//     (float bits of pz << 32 | face index) - pz >= 0, so its bit pattern is order preserving.
// One thread per face walks the pixels of the face's bounding box; a second kernel recomputes the barycentrics of the
// winning face per pixel.  All arithmetic is fp32 with explicit round-to-nearest operations (no FMA contraction), so the
// NumPy restatement in the oracle reproduces it bit for bit.
#include "dh_common.cuh"

#include <math.h>
#include <string.h>

namespace dh {

constexpr float kRasterEps = 1e-8f;

struct RasterCamDev {
    float R[9];     // view = X R + T (row-vector convention, like pytorch3d)
    float T[3];
    float sx, sy;   // NDC = (sx X / Z, sy Y / Z)
};

struct RasterSettings {
    float blur_radius, blur_sqrt;
    int cull_backfaces, perspective_correct, clip_barycentric;
};

__device__ __forceinline__ float ndc_range(int S1, int S2) {
    float range = 2.0f;
    if (S1 > S2) range = __fdiv_rn(__fmul_rn((float)S1, range), (float)S2);
    return range;
}

__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
    const float range = ndc_range(S1, S2);
    const float offset = __fdiv_rn(range, 2.0f);
    return __fadd_rn(-offset, __fdiv_rn(__fadd_rn(__fmul_rn(range, (float)i), offset), (float)S1));
}

__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float bax = __fsub_rn(bx, ax), bay = __fsub_rn(by, ay);
    const float l2 = __fadd_rn(__fmul_rn(bax, bax), __fmul_rn(bay, bay));
    if (l2 <= kRasterEps) {
        const float dx = __fsub_rn(px, bx), dy = __fsub_rn(py, by);
        return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    }
    float t = __fdiv_rn(__fadd_rn(__fmul_rn(bax, __fsub_rn(px, ax)), __fmul_rn(bay, __fsub_rn(py, ay))), l2);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float qx = __fadd_rn(ax, __fmul_rn(t, bax)), qy = __fadd_rn(ay, __fmul_rn(t, bay));
    const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

struct Tri {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

// pytorch3d CheckPixelInsideFace for one (pixel, face).  Returns true when the face covers the pixel; pz and the clipped,
// perspective-corrected barycentrics are written.
__device__ __forceinline__ bool pixel_in_face(const Tri& t, float px, float py, const RasterSettings& st, float& pz, float& b0,
                                              float& b1, float& b2) {
    const float zmax = fmaxf(t.z0, fmaxf(t.z1, t.z2));
    const float xmin = fminf(t.x0, fminf(t.x1, t.x2)) - st.blur_sqrt, xmax = fmaxf(t.x0, fmaxf(t.x1, t.x2)) + st.blur_sqrt;
    const float ymin = fminf(t.y0, fminf(t.y1, t.y2)) - st.blur_sqrt, ymax = fmaxf(t.y0, fmaxf(t.y1, t.y2)) + st.blur_sqrt;
    const bool outside = px > xmax || px < xmin || py > ymax || py < ymin;
    const float area = edge_fn(t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);
    const bool back = area < 0.0f;
    const bool zero_area = area <= kRasterEps && area >= -kRasterEps;
    if (zmax < 0.0f || (st.cull_backfaces && back) || outside || zero_area) return false;
    // barycentric coordinates
    const float a = __fadd_rn(edge_fn(t.x2, t.y2, t.x0, t.y0, t.x1, t.y1), kRasterEps);
    float w0 = __fdiv_rn(edge_fn(px, py, t.x1, t.y1, t.x2, t.y2), a);
    float w1 = __fdiv_rn(edge_fn(px, py, t.x2, t.y2, t.x0, t.y0), a);
    float w2 = __fdiv_rn(edge_fn(px, py, t.x0, t.y0, t.x1, t.y1), a);
    if (st.perspective_correct) {
        const float t0 = __fmul_rn(__fmul_rn(w0, t.z1), t.z2);
        const float t1 = __fmul_rn(__fmul_rn(t.z0, w1), t.z2);
        const float t2 = __fmul_rn(__fmul_rn(t.z0, t.z1), w2);
        const float den = fmaxf(__fadd_rn(__fadd_rn(t0, t1), t2), kRasterEps);
        w0 = __fdiv_rn(t0, den); w1 = __fdiv_rn(t1, den); w2 = __fdiv_rn(t2, den);
    }
    float c0 = w0, c1 = w1, c2 = w2;
    if (st.clip_barycentric) {
        c0 = fmaxf(0.0f, fminf(1.0f, w0)); c1 = fmaxf(0.0f, fminf(1.0f, w1)); c2 = fmaxf(0.0f, fminf(1.0f, w2));
        const float s = fmaxf(__fadd_rn(__fadd_rn(c0, c1), c2), 1e-5f);
        c0 = __fdiv_rn(c0, s); c1 = __fdiv_rn(c1, s); c2 = __fdiv_rn(c2, s);
    }
    pz = __fadd_rn(__fadd_rn(__fmul_rn(c0, t.z0), __fmul_rn(c1, t.z1)), __fmul_rn(c2, t.z2));
    if (pz < 0.0f) return false;
    const bool inside = w0 > 0.0f && w1 > 0.0f && w2 > 0.0f;
    if (!inside) {
        const float d = fminf(seg_dist2(px, py, t.x0, t.y0, t.x1, t.y1),
                              fminf(seg_dist2(px, py, t.x0, t.y0, t.x2, t.y2), seg_dist2(px, py, t.x1, t.y1, t.x2, t.y2)));
        if (d >= st.blur_radius) return false;
    }
    b0 = c0; b1 = c1; b2 = c2;
    return true;
}

__global__ void __launch_bounds__(256) raster_project_kernel(const float* __restrict__ verts, int V, RasterCamDev cam,
                                                             float* __restrict__ ndc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float X = verts[3 * (size_t)i], Y = verts[3 * (size_t)i + 1], Z = verts[3 * (size_t)i + 2];
    // view = X R + T, left-to-right fp32 accumulation (identity R / zero T for every reference caller -> exact)
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, cam.R[c]), __fmul_rn(Y, cam.R[3 + c])), __fmul_rn(Z, cam.R[6 + c])), cam.T[c]);
    // homogeneous divide with pytorch3d's eps clamp on the denominator
    float den = v[2];
    const float ad = fmaxf(fabsf(den), 1e-8f);      // (MeshRasterizer passes eps = None -> no clamp; kept tiny for z = 0)
    den = den < 0.0f ? -ad : ad;
    ndc[3 * (size_t)i] = __fdiv_rn(__fmul_rn(cam.sx, v[0]), den);
    ndc[3 * (size_t)i + 1] = __fdiv_rn(__fmul_rn(cam.sy, v[1]), den);
    ndc[3 * (size_t)i + 2] = v[2];
}

__device__ __forceinline__ Tri load_tri(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int f) {
    const int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
    Tri t;
    t.x0 = ndc[3 * (size_t)i0]; t.y0 = ndc[3 * (size_t)i0 + 1]; t.z0 = ndc[3 * (size_t)i0 + 2];
    t.x1 = ndc[3 * (size_t)i1]; t.y1 = ndc[3 * (size_t)i1 + 1]; t.z1 = ndc[3 * (size_t)i1 + 2];
    t.x2 = ndc[3 * (size_t)i2]; t.y2 = ndc[3 * (size_t)i2 + 1]; t.z2 = ndc[3 * (size_t)i2 + 2];
    return t;
}

__global__ void __launch_bounds__(128) raster_faces_kernel(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int F,
                                                           int H, int W, RasterSettings st, unsigned long long* keys) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const Tri t = load_tri(ndc, faces, f);
    const float zmax = fmaxf(t.z0, fmaxf(t.z1, t.z2));
    if (zmax < 0.0f) return;
    // candidate pixel range from the blur-grown bounding box (conservative; the exact test is repeated per pixel)
    const float xmin = fminf(t.x0, fminf(t.x1, t.x2)) - st.blur_sqrt, xmax = fmaxf(t.x0, fmaxf(t.x1, t.x2)) + st.blur_sqrt;
    const float ymin = fminf(t.y0, fminf(t.y1, t.y2)) - st.blur_sqrt, ymax = fmaxf(t.y0, fmaxf(t.y1, t.y2)) + st.blur_sqrt;
    if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) return;     // NaN vertices
    const float rx = ndc_range(W, H), ry = ndc_range(H, W);
    // ndc = -r/2 + (r i + r/2) / S   <=>   i = ((ndc + r/2) S - r/2) / r
    auto to_idx = [](float v, float r, int S) { return ((v + 0.5f * r) * (float)S - 0.5f * r) / r; };
    int xi0 = (int)floorf(fminf(fmaxf(to_idx(xmin, rx, W), -2.0f), (float)W + 1.0f)) - 1;
    int xi1 = (int)ceilf(fminf(fmaxf(to_idx(xmax, rx, W), -2.0f), (float)W + 1.0f)) + 1;
    int yi0 = (int)floorf(fminf(fmaxf(to_idx(ymin, ry, H), -2.0f), (float)H + 1.0f)) - 1;
    int yi1 = (int)ceilf(fminf(fmaxf(to_idx(ymax, ry, H), -2.0f), (float)H + 1.0f)) + 1;
    xi0 = max(xi0, 0); yi0 = max(yi0, 0); xi1 = min(xi1, W - 1); yi1 = min(yi1, H - 1);
    for (int yidx = yi0; yidx <= yi1; ++yidx) {
        const float py = pix_to_ndc(yidx, H, W);
        for (int xidx = xi0; xidx <= xi1; ++xidx) {
            const float px = pix_to_ndc(xidx, W, H);
            float pz, b0, b1, b2;
            if (!pixel_in_face(t, px, py, st, pz, b0, b1, b2)) continue;
            // image row / column: both axes are reversed (+x left, +y up)
            const int row = H - 1 - yidx, col = W - 1 - xidx;
            const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned int)f;
            unsigned long long* k = keys + (size_t)row * W + col;
            if (key < *(volatile unsigned long long*)k) atomicMin(k, key);
        }
    }
}

__global__ void __launch_bounds__(256) raster_resolve_kernel(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int H,
                                                             int W, RasterSettings st, const unsigned long long* __restrict__ keys,
                                                             int32_t* __restrict__ pix_to_face, float* __restrict__ zbuf,
                                                             float* __restrict__ bary) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H * W) return;
    const unsigned long long key = keys[q];
    int f = -1;
    float pz = -1.0f, b0 = -1.0f, b1 = -1.0f, b2 = -1.0f;       // pytorch3d fills empty pixels with -1
    if (key != 0xFFFFFFFFFFFFFFFFull) {
        f = (int)(key & 0xFFFFFFFFull);
        const int row = q / W, col = q - row * W;
        const float px = pix_to_ndc(W - 1 - col, W, H), py = pix_to_ndc(H - 1 - row, H, W);
        const Tri t = load_tri(ndc, faces, f);
        pixel_in_face(t, px, py, st, pz, b0, b1, b2);
    }
    pix_to_face[q] = f;
    zbuf[q] = pz;
    bary[3 * (size_t)q] = b0; bary[3 * (size_t)q + 1] = b1; bary[3 * (size_t)q + 2] = b2;
}

// hard blend of a barycentrically interpolated vertex attribute: out (H,W,D+1), alpha in the last channel,
// background 0 (pytorch3d_renderer.py:56-142, :487-537)
__global__ void __launch_bounds__(256) raster_interpolate_kernel(const float* __restrict__ attr, int D, const int32_t* __restrict__ faces,
                                                                 const int32_t* __restrict__ pix_to_face, const float* __restrict__ bary,
                                                                 int P, float* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P) return;
    const int f = pix_to_face[q];
    float* o = out + (size_t)q * (D + 1);
    if (f < 0) {
        for (int d = 0; d <= D; ++d) o[d] = 0.0f;
        return;
    }
    const int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
    const float b0 = bary[3 * (size_t)q], b1 = bary[3 * (size_t)q + 1], b2 = bary[3 * (size_t)q + 2];
    for (int d = 0; d < D; ++d)
        o[d] = __fadd_rn(__fadd_rn(__fmul_rn(attr[(size_t)i0 * D + d], b0), __fmul_rn(attr[(size_t)i1 * D + d], b1)),
                         __fmul_rn(attr[(size_t)i2 * D + d], b2));
    o[D] = 1.0f;
}

static size_t raster_layout(int V, int P, size_t* o_ndc, size_t* o_keys) {
    size_t o = 0;
    *o_ndc = o;  o = align_up(o + sizeof(float) * 3 * (size_t)V, 256);
    *o_keys = o; o = align_up(o + sizeof(unsigned long long) * (size_t)P, 256);
    return o;
}

}  // namespace dh

using namespace dh;

extern "C" {

size_t dh_raster_workspace_bytes(int V, int H, int W) {
    if (V < 1 || H < 1 || W < 1) return 0;
    size_t a, b;
    return raster_layout(V, H * W, &a, &b);
}

int dh_rasterize_meshes(const float* verts, int V, const int32_t* faces, int F, int H, int W, const float* R_host9,
                        const float* T_host3, float sx, float sy, float blur_radius, int cull_backfaces, int perspective_correct,
                        int clip_barycentric, int32_t* pix_to_face, float* zbuf, float* bary, void* ws, size_t ws_bytes,
                        void* stream) {
    DH_REQUIRE(verts && faces && R_host9 && T_host3 && pix_to_face && zbuf && bary && ws);
    DH_REQUIRE(V >= 1 && F >= 0 && H >= 1 && W >= 1 && blur_radius >= 0.0f);
    size_t o_ndc, o_keys;
    if (ws_bytes < raster_layout(V, H * W, &o_ndc, &o_keys)) return DH_ERR_WORKSPACE;
    float* ndc = reinterpret_cast<float*>(static_cast<char*>(ws) + o_ndc);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(static_cast<char*>(ws) + o_keys);
    RasterCamDev cam;
    memcpy(cam.R, R_host9, sizeof(cam.R));
    memcpy(cam.T, T_host3, sizeof(cam.T));
    cam.sx = sx; cam.sy = sy;
    RasterSettings st;
    st.blur_radius = blur_radius; st.blur_sqrt = sqrtf(blur_radius);
    st.cull_backfaces = cull_backfaces; st.perspective_correct = perspective_correct; st.clip_barycentric = clip_barycentric;
    cudaStream_t s = as_stream(stream);
    DH_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * (size_t)H * W, s));
    raster_project_kernel<<<(V + 255) / 256, 256, 0, s>>>(verts, V, cam, ndc);
    DH_LAUNCH_CHECK();
    if (F > 0) {
        raster_faces_kernel<<<(F + 127) / 128, 128, 0, s>>>(ndc, faces, F, H, W, st, keys);
        DH_LAUNCH_CHECK();
    }
    raster_resolve_kernel<<<(H * W + 255) / 256, 256, 0, s>>>(ndc, faces, H, W, st, keys, pix_to_face, zbuf, bary);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_interpolate_face_attributes(const float* attr, int D, const int32_t* faces, const int32_t* pix_to_face, const float* bary,
                                   int H, int W, float* out, void* stream) {
    DH_REQUIRE(attr && faces && pix_to_face && bary && out && D >= 1 && H >= 1 && W >= 1);
    raster_interpolate_kernel<<<(H * W + 255) / 256, 256, 0, as_stream(stream)>>>(attr, D, faces, pix_to_face, bary, H * W, out);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
