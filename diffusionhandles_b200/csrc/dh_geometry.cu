// K1: fused unproject -> rigid transform -> project, plus the stand-alone
// unprojection / projection / transform_points entry points.
//
// Bit-exactness notes: every fp32/fp64 operation below is written with the
// explicit round-to-nearest intrinsics so that nvcc can never contract a multiply-add into an FMA;
// the file is additionally compiled with -fmad=false.  The only fused operations are the two that the
// reference itself executes fused (the fp32 dot product, A.2).
#include "dh_common.cuh"

#include <math.h>
#include <string.h>

namespace dh {

thread_local int g_last_cuda_error = 0;

constexpr int kTile = 4096;          // pixels per compaction tile
constexpr int kTileThreads = 1024;   // 4 pixels per thread

struct CamDev {
    float k[9];
    float kinv[9];
    int diag;   // 1: off-diagonal entries of k and kinv are all zero (the bit-exact contract class)
};

static CamDev make_cam(const dh_camera* c) {
    CamDev d;
    memcpy(d.k, c->k, sizeof(d.k));
    memcpy(d.kinv, c->kinv, sizeof(d.kinv));
    d.diag = 1;
    for (int i = 0; i < 9; ++i)
        if (i % 4 != 0 && (c->k[i] != 0.f || c->kinv[i] != 0.f)) d.diag = 0;
    return d;
}

// depth_transform.py:634-639: p = M @ ((D * Kinv) @ [x, y, 1]^T), identity extrinsics.
__device__ __forceinline__ void unproject(const CamDev& cam, float d, float x, float y, float& X, float& Y, float& Z) {
    if (cam.diag) {
        X = -__fmul_rn(__fmul_rn(d, cam.kinv[0]), x);
        Y = -__fmul_rn(__fmul_rn(d, cam.kinv[4]), y);
        Z = __fmul_rn(__fmul_rn(d, cam.kinv[8]), 1.0f);
    } else {  // general intrinsics: left-to-right fp32 accumulation (outside the bit-exact contract)
        float r[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a = __fmul_rn(__fmul_rn(d, cam.kinv[3 * i + 0]), x);
            float b = __fmul_rn(__fmul_rn(d, cam.kinv[3 * i + 1]), y);
            float c = __fmul_rn(__fmul_rn(d, cam.kinv[3 * i + 2]), 1.0f);
            r[i] = __fadd_rn(__fadd_rn(a, b), c);
        }
        X = -r[0]; Y = -r[1]; Z = r[2];
    }
}

// depth_transform.py:666-687: (X,Y,Z) fp64 in pytorch3d axes -> integer pixel (u = col, v = row).
// clip to [0, size-1] first, then round half to even.  Returns false when z is NaN (never wins).
__device__ __forceinline__ bool project(const CamDev& cam, double X, double Y, double Z, int H, int W,
                                        int& u, int& v, uint64_t& key) {
    const double px = -X, py = -Y, pz = Z;
    double projx, projy, projz;
    if (cam.diag) {
        projx = __dmul_rn((double)cam.k[0], px);
        projy = __dmul_rn((double)cam.k[4], py);
        projz = __dmul_rn((double)cam.k[8], pz);
    } else {
        projx = __dadd_rn(__dadd_rn(__dmul_rn((double)cam.k[0], px), __dmul_rn((double)cam.k[1], py)), __dmul_rn((double)cam.k[2], pz));
        projy = __dadd_rn(__dadd_rn(__dmul_rn((double)cam.k[3], px), __dmul_rn((double)cam.k[4], py)), __dmul_rn((double)cam.k[5], pz));
        projz = __dadd_rn(__dadd_rn(__dmul_rn((double)cam.k[6], px), __dmul_rn((double)cam.k[7], py)), __dmul_rn((double)cam.k[8], pz));
    }
    double uu = __ddiv_rn(projx, projz);
    double vv = __ddiv_rn(projy, projz);
    const double m = (double)((H > W ? H : W) - 1);
    uu = __dmul_rn(__dadd_rn(__dmul_rn(uu, 0.5), 0.5), m);
    vv = __dmul_rn(__dadd_rn(__dmul_rn(vv, 0.5), 0.5), m);
    uu = fmin(fmax(uu, 0.0), (double)(W - 1));
    vv = fmin(fmax(vv, 0.0), (double)(H - 1));
    u = (int)rint(uu);
    v = (int)rint(vv);
    const double z = __dadd_rn(Z, 0.0);   // folds -0 into +0 (the reference compares with '<')
    key = z_to_key(z);
    return !(z != z);
}

// ------------------------------------------------------------------------------------------------
// stand-alone unprojection (depth_to_world_coords)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unproject_kernel(const float* __restrict__ depth, int P, int W, CamDev cam,
                                                        const float* __restrict__ xs, const float* __restrict__ ys,
                                                        float* __restrict__ points) {
    const int e = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int row = p / W, col = p - row * W;
    float X, Y, Z;
    unproject(cam, depth[(size_t)e * P + p], xs[col], ys[row], X, Y, Z);
    float* o = points + ((size_t)e * P + p) * 3;
    o[0] = X; o[1] = Y; o[2] = Z;
}

// ------------------------------------------------------------------------------------------------
// foreground compaction (raster order) + unprojection of the foreground points
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) fg_count_kernel(const float* __restrict__ mask, int P, int ntiles,
                                                                int32_t* __restrict__ tile_counts) {
    __shared__ int warp_sums[32];
    const int e = blockIdx.y, tile = blockIdx.x;
    const int p0 = tile * kTile + threadIdx.x * 4;
    const float* m = mask + (size_t)e * P;
    int c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (p0 + i < P && m[p0 + i] != 0.0f) ++c;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane_id() == 0) warp_sums[warp_id()] = c;
    __syncthreads();
    if (warp_id() == 0) {
        int s = warp_sums[lane_id()];
        s = __reduce_add_sync(0xFFFFFFFFu, s);
        if (lane_id() == 0) tile_counts[e * ntiles + tile] = s;
    }
}

// `points` != nullptr: the foreground points are taken from an (P,3) fp32 array instead of being unprojected
// from `depth` (transform_point_cloud entry point).
__global__ void __launch_bounds__(kTileThreads) fg_compact_kernel(
    const float* __restrict__ depth, const float* __restrict__ mask, int P, int W, int ntiles, CamDev cam,
    const float* __restrict__ xs, const float* __restrict__ ys, const int32_t* __restrict__ tile_counts,
    int32_t* __restrict__ fg_index, float* __restrict__ fgX, float* __restrict__ fgY, float* __restrict__ fgZ,
    int32_t* __restrict__ n_fg, const float* __restrict__ points) {
    __shared__ int scan_smem[33];
    __shared__ int base_smem;
    const int e = blockIdx.y, tile = blockIdx.x;
    // base = number of foreground pixels in the preceding tiles of this edit
    int part = 0;
    for (int t = threadIdx.x; t < tile; t += blockDim.x) part += tile_counts[e * ntiles + t];
    int tot;
    block_exclusive_scan(part, scan_smem, tot);
    if (threadIdx.x == 0) base_smem = tot;
    __syncthreads();
    const int base = base_smem;

    const int p0 = tile * kTile + threadIdx.x * 4;
    const float* m = mask + (size_t)e * P;
    bool f[4];
    int c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[i] = (p0 + i < P) && (m[p0 + i] != 0.0f);
        c += f[i];
    }
    int total;
    int pos = base + block_exclusive_scan(c, scan_smem, total);
    const size_t eo = (size_t)e * P;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!f[i]) continue;
        const int p = p0 + i;
        const int row = p / W, col = p - row * W;
        float X, Y, Z;
        if (points) {
            X = points[3 * (eo + p)]; Y = points[3 * (eo + p) + 1]; Z = points[3 * (eo + p) + 2];
        } else {
            unproject(cam, depth[eo + p], xs[col], ys[row], X, Y, Z);
        }
        fg_index[eo + pos] = p;
        fgX[eo + pos] = X; fgY[eo + pos] = Y; fgZ[eo + pos] = Z;
        ++pos;
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) n_fg[e] = base + total;
}

// Sequential fp32 sum in raster order, one warp per component (np.mean(points[mask], axis=0) adds row
// by row, depth_transform.py:509).  All 32 lanes carry the same running sum.  The addends are staged in
// shared memory 512 at a time with asynchronous copies (a ring of tiles per warp) and read back with 128-bit
// broadcast loads into a REGISTER double buffer of 32 values: while the dependent FADD chain (4 cycles per
// element) works through one group, the loads of the next group are already in flight.
__global__ void __launch_bounds__(96) fg_centroid_kernel(const float* __restrict__ fgX, const float* __restrict__ fgY,
                                                         const float* __restrict__ fgZ, const int32_t* __restrict__ n_fg,
                                                         int P, float* __restrict__ centroid) {
    constexpr int kTileF = 512, kBuf = 4, kGroup = 8;            // floats per tile, ring depth, float4 per register group
    __shared__ __align__(16) float ring[3][kBuf][kTileF];
    const int e = blockIdx.x, comp = warp_id(), lane = lane_id();
    const float* a = (comp == 0 ? fgX : comp == 1 ? fgY : fgZ) + (size_t)e * P;   // 256-byte aligned (workspace layout)
    const int n = n_fg[e];
    const unsigned full = 0xFFFFFFFFu;
    float s = 0.0f;
    const int n_tiles = (reinterpret_cast<uintptr_t>(a) & 15) == 0 ? n / kTileF : 0;
    auto issue = [&](int t) {
        if (t < n_tiles) {
            const float* src = a + (size_t)t * kTileF + lane * 4;
            float* dst = &ring[comp][t % kBuf][lane * 4];
#pragma unroll
            for (int q = 0; q < kTileF / 128; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 128 * q)),
                             "l"(src + 128 * q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");       // (empty groups keep the wait count uniform)
    };
#pragma unroll
    for (int t = 0; t < kBuf - 1; ++t) issue(t);
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + kBuf - 1);                                        // refills the slot consumed in iteration t - 1
        asm volatile("cp.async.wait_group %0;" ::"n"(kBuf - 1) : "memory");
        __syncwarp();
        const float4* tile = reinterpret_cast<const float4*>(ring[comp][t % kBuf]);
        float4 cur[kGroup], nxt[kGroup];
#pragma unroll
        for (int j = 0; j < kGroup; ++j) cur[j] = tile[j];          // same address in every lane: broadcast
#pragma unroll
        for (int g = 0; g < kTileF / 4 / kGroup; ++g) {
            if (g + 1 < kTileF / 4 / kGroup) {
#pragma unroll
                for (int j = 0; j < kGroup; ++j) nxt[j] = tile[(g + 1) * kGroup + j];
            }
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                s = __fadd_rn(s, cur[j].x);
                s = __fadd_rn(s, cur[j].y);
                s = __fadd_rn(s, cur[j].z);
                s = __fadd_rn(s, cur[j].w);
            }
#pragma unroll
            for (int j = 0; j < kGroup; ++j) cur[j] = nxt[j];
        }
        __syncwarp();                                               // every lane is done with the slot before it is refilled
    }
    for (int base = n_tiles * kTileF; base < n; base += 32) {
        const float v = base + lane < n ? a[base + lane] : 0.0f;
        const int cnt = min(32, n - base);
        for (int j = 0; j < cnt; ++j) s = __fadd_rn(s, __shfl_sync(full, v, j));
    }
    if (lane == 0) centroid[e * 3 + comp] = __fdiv_rn(s, (float)n);
}

// depth_transform.py:512-531 for one point: Rodrigues about the centroid + translation, fp32 -> fp64 in the
// reference's operation order.
__device__ __forceinline__ void rigid_point(float px, float py, float pz, float cx, float cy, float cz, const dh_rigid& rg,
                                            double& X, double& Y, double& Z) {
    const float qx = __fsub_rn(px, cx), qy = __fsub_rn(py, cy), qz = __fsub_rn(pz, cz);
    const float ax = rg.axis[0], ay = rg.axis[1], az = rg.axis[2];
    const float crx = __fsub_rn(__fmul_rn(ay, qz), __fmul_rn(az, qy));
    const float cry = __fsub_rn(__fmul_rn(az, qx), __fmul_rn(ax, qz));
    const float crz = __fsub_rn(__fmul_rn(ax, qy), __fmul_rn(ay, qx));
    float d = __fmul_rn(qy, ay);
    d = __fmaf_rn(qx, ax, d);
    d = __fmaf_rn(qz, az, d);
    const float t3x = __fmul_rn(ax, d), t3y = __fmul_rn(ay, d), t3z = __fmul_rn(az, d);
    const double c = rg.cos_t, s = rg.sin_t, omc = __dsub_rn(1.0, c);
    X = __dadd_rn(__dadd_rn(__dmul_rn((double)qx, c), __dmul_rn((double)crx, s)), __dmul_rn((double)t3x, omc));
    Y = __dadd_rn(__dadd_rn(__dmul_rn((double)qy, c), __dmul_rn((double)cry, s)), __dmul_rn((double)t3y, omc));
    Z = __dadd_rn(__dadd_rn(__dmul_rn((double)qz, c), __dmul_rn((double)crz, s)), __dmul_rn((double)t3z, omc));
    X = __dadd_rn(__dadd_rn(X, (double)cx), rg.t[0]);
    Y = __dadd_rn(__dadd_rn(Y, (double)cy), rg.t[1]);
    Z = __dadd_rn(__dadd_rn(Z, (double)cz), rg.t[2]);
}

__global__ void __launch_bounds__(256) rigid_all_kernel(const float* __restrict__ points, int N, const dh_rigid* __restrict__ rigid,
                                                        const float* __restrict__ centroid, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double X, Y, Z;
    rigid_point(points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2], centroid[0], centroid[1], centroid[2],
                rigid[0], X, Y, Z);
    out[3 * (size_t)i] = X; out[3 * (size_t)i + 1] = Y; out[3 * (size_t)i + 2] = Z;
}

// ------------------------------------------------------------------------------------------------
// K1 main kernel: slot s < P -> background point s; slot s >= P -> foreground point j = s - P.
// ------------------------------------------------------------------------------------------------
// With kSplat the kernel also performs pass 1 of the z-buffer splat (dh_splat.cu): the 64-bit atomicMin on the
// order-preserving z key, pre-reduced inside a warp when all of its lanes hit one pixel and skipped when a plain load already
// shows the point cannot win.  The kernel is bound by its fp64 arithmetic, so the atomics ride along for free and the
// separate pass - one more read of pix / zkey - disappears.
template <bool kSplat>
__global__ void __launch_bounds__(256) transform_project_kernel(
    const float* __restrict__ bg_depth, int P, int H, int W, CamDev cam, const dh_rigid* __restrict__ rigid,
    const float* __restrict__ xs, const float* __restrict__ ys,
    const float* __restrict__ fgX, const float* __restrict__ fgY, const float* __restrict__ fgZ,
    const int32_t* __restrict__ n_fg, const float* __restrict__ centroid,
    int32_t* __restrict__ pix, uint64_t* __restrict__ zkey, double* __restrict__ points_out, uint64_t* zbuf) {
    const int e = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t po = (size_t)e * 2 * P + slot;
    bool live = slot < 2 * P;
    double X = 0.0, Y = 0.0, Z = 0.0;
    if (live && slot < P) {
        const int row = slot / W, col = slot - row * W;
        float x, y, z;
        unproject(cam, bg_depth[(size_t)e * P + slot], xs[col], ys[row], x, y, z);
        X = (double)x; Y = (double)y; Z = (double)z;
    } else if (live) {
        const int j = slot - P;
        live = j < n_fg[e];
        if (live) {
            const size_t eo = (size_t)e * P + j;
            rigid_point(fgX[eo], fgY[eo], fgZ[eo], centroid[e * 3 + 0], centroid[e * 3 + 1], centroid[e * 3 + 2], rigid[e], X, Y, Z);
        }
    }
    int q = -1;
    uint64_t key = kEmptyZ;
    if (live) {
        int u, v;
        const bool ok = project(cam, X, Y, Z, H, W, u, v, key);
        q = ok ? v * W + u : -1;
        pix[po] = q;
        zkey[po] = key;
        if (points_out) {
            double* o = points_out + po * 3;
            o[0] = X; o[1] = Y; o[2] = Z;
        }
    }
    if (!kSplat) return;
    uint64_t* zb = zbuf + (size_t)e * P;
    const unsigned full = 0xFFFFFFFFu;
    const int q0 = __shfl_sync(full, q, 0);
    if (__all_sync(full, q == q0)) {                 // clamped off-screen points pile onto border pixels: one atomic per warp
        if (q0 < 0) return;
        const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
        const uint32_t mhi = __reduce_min_sync(full, hi);
        const uint32_t mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xFFFFFFFFu);
        if (lane_id() == 0) {
            const uint64_t m = ((uint64_t)mhi << 32) | mlo;
            if (m < *(volatile uint64_t*)(zb + q0)) atomicMin((unsigned long long*)(zb + q0), (unsigned long long)m);
        }
        return;
    }
    if (q >= 0 && key < *(volatile uint64_t*)(zb + q)) atomicMin((unsigned long long*)(zb + q), (unsigned long long)key);
}

__global__ void __launch_bounds__(256) project_points_kernel(const double* __restrict__ points, int N, int H, int W,
                                                             CamDev cam, int32_t* __restrict__ pix, uint64_t* __restrict__ zkey,
                                                             int32_t* __restrict__ uo, int32_t* __restrict__ vo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int u, v;
    uint64_t key;
    const bool ok = project(cam, points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2], H, W, u, v, key);
    pix[i] = ok ? v * W + u : -1;
    zkey[i] = key;
    if (uo) uo[i] = u;
    if (vo) vo[i] = v;
}

// ------------------------------------------------------------------------------------------------
// transform_points (torch fp32 variant, depth_transform.py:439-459): centroid = mean over ALL points.
// Pairwise (deterministic) fp32 reduction; parity with torch is tolerance based.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tp_partial_sum_kernel(const float* __restrict__ pts, int N, double* __restrict__ partial) {
    __shared__ double sm[3][8];
    double s[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        s[0] += pts[3 * (size_t)i]; s[1] += pts[3 * (size_t)i + 1]; s[2] += pts[3 * (size_t)i + 2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_down_sync(0xFFFFFFFFu, s[c], o);
        if (lane_id() == 0) sm[c][warp_id()] = s[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w];
        partial[blockIdx.x * 3 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) tp_apply_kernel(const float* __restrict__ pts, int N, const double* __restrict__ partial,
                                                       int nblocks, float ax, float ay, float az, float c, float s,
                                                       float tx, float ty, float tz, float* __restrict__ out) {
    __shared__ float cen[3];
    if (threadIdx.x < 3) {
        double t = 0;
        for (int b = 0; b < nblocks; ++b) t += partial[b * 3 + threadIdx.x];
        cen[threadIdx.x] = (float)(t / (double)N);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float qx = pts[3 * (size_t)i] - cen[0], qy = pts[3 * (size_t)i + 1] - cen[1], qz = pts[3 * (size_t)i + 2] - cen[2];
    const float crx = ay * qz - az * qy, cry = az * qx - ax * qz, crz = ax * qy - ay * qx;
    const float d = qx * ax + qy * ay + qz * az;
    const float omc = 1.0f - c;
    out[3 * (size_t)i + 0] = (qx * c + crx * s + ax * d * omc) + cen[0] + tx;
    out[3 * (size_t)i + 1] = (qy * c + cry * s + ay * d * omc) + cen[1] + ty;
    out[3 * (size_t)i + 2] = (qz * c + crz * s + az * d * omc) + cen[2] + tz;
}

}  // namespace dh

using namespace dh;

extern "C" {

const char* dh_status_string(int status) {
    switch (status) {
        case DH_OK: return "ok";
        case DH_ERR_INVALID_ARGUMENT: return "invalid argument";
        case DH_ERR_UNSUPPORTED: return "unsupported";
        case DH_ERR_CUDA: return "CUDA runtime error";
        case DH_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown status";
    }
}

int dh_abi_version(void) { return DH_B200_ABI_VERSION; }

int dh_last_cuda_error(void) { return g_last_cuda_error; }

int dh_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    DH_CUDA_CHECK(cudaGetDevice(&dev));
    if (sm_count) DH_CUDA_CHECK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) DH_CUDA_CHECK(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) DH_CUDA_CHECK(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return DH_OK;
}

int dh_linspace_f32_host(float start, float end, int steps, float* out) {
    DH_REQUIRE(out && steps >= 1);
    if (steps == 1) { out[0] = start; return DH_OK; }
    const float step = (end - start) / (float)(steps - 1);
    const int half = steps / 2;
    for (int i = 0; i < steps; ++i)
        out[i] = i < half ? fmaf(step, (float)i, start) : fmaf(-step, (float)(steps - 1 - i), end);
    return DH_OK;
}

int dh_unproject(const float* depth, int B, int H, int W, const dh_camera* cam_host, const float* xs, const float* ys,
                 float* points, void* stream) {
    DH_REQUIRE(depth && cam_host && xs && ys && points && B >= 1);
    DH_REQUIRE(H >= 2 && W >= 2);
    const int P = H * W;
    dim3 grid((P + 255) / 256, B);
    unproject_kernel<<<grid, 256, 0, as_stream(stream)>>>(depth, P, W, make_cam(cam_host), xs, ys, points);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

size_t dh_transform_points_workspace_bytes(int N) {
    (void)N;
    return 256 * 3 * sizeof(double);
}

int dh_transform_points(const float* points, int N, float angle_degrees, const float* axis_host3,
                        const float* translation_host3, float* out, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(points && out && axis_host3 && translation_host3 && ws && N >= 1);
    if (ws_bytes < dh_transform_points_workspace_bytes(N)) return DH_ERR_WORKSPACE;
    const int nblocks = N >= 256 * 256 ? 256 : (N + 255) / 256;
    float a[3] = {axis_host3[0], axis_host3[1], axis_host3[2]};
    const float nrm = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (float& v : a) v /= nrm;
    const float ang = angle_degrees * (float)(M_PI / 180.0);
    tp_partial_sum_kernel<<<nblocks, 256, 0, as_stream(stream)>>>(points, N, (double*)ws);
    DH_LAUNCH_CHECK();
    tp_apply_kernel<<<(N + 255) / 256, 256, 0, as_stream(stream)>>>(points, N, (const double*)ws, nblocks, a[0], a[1], a[2],
                                                                     cosf(ang), sinf(ang), translation_host3[0],
                                                                     translation_host3[1], translation_host3[2], out);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

// workspace of the fused pc path: [tile_counts int32 B*ntiles][rigid B][fgX][fgY][fgZ float B*P each]
static size_t k1_ws_layout(int B, int P, size_t* off_counts, size_t* off_rigid, size_t* off_x, size_t* off_y, size_t* off_z) {
    const int ntiles = (P + kTile - 1) / kTile;
    size_t o = 0;
    *off_counts = o; o = align_up(o + sizeof(int32_t) * (size_t)B * ntiles, 256);
    *off_rigid = o;  o = align_up(o + sizeof(dh_rigid) * (size_t)B, 256);
    *off_x = o;      o = align_up(o + sizeof(float) * (size_t)B * P, 256);
    *off_y = o;      o = align_up(o + sizeof(float) * (size_t)B * P, 256);
    *off_z = o;      o = align_up(o + sizeof(float) * (size_t)B * P, 256);
    return o;
}

size_t dh_edit_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    size_t a, b, c, d, e;
    return k1_ws_layout(B, H * W, &a, &b, &c, &d, &e);
}

static int k1_launch(const float* depth, const float* bg_depth, const float* fg_mask, int B, int H, int W,
                     const dh_camera* cam_host, const dh_rigid* rigid_host, const float* xs, const float* ys,
                     int32_t* pix, uint64_t* zkey, int32_t* fg_index, int32_t* n_fg, float* centroid,
                     double* points_out, void* ws, size_t ws_bytes, uint64_t* zbuf, void* stream) {
    DH_REQUIRE(depth && bg_depth && fg_mask && cam_host && rigid_host && xs && ys && pix && zkey && fg_index && n_fg && centroid && ws);
    DH_REQUIRE(B >= 1 && H >= 2 && W >= 2);
    DH_REQUIRE((long long)H * W <= (1ll << 29));
    const int P = H * W;
    size_t oc, orr, ox, oy, oz;
    if (ws_bytes < k1_ws_layout(B, P, &oc, &orr, &ox, &oy, &oz)) return DH_ERR_WORKSPACE;
    char* w = static_cast<char*>(ws);
    int32_t* tile_counts = reinterpret_cast<int32_t*>(w + oc);
    dh_rigid* rigid_dev = reinterpret_cast<dh_rigid*>(w + orr);
    float* fgX = reinterpret_cast<float*>(w + ox);
    float* fgY = reinterpret_cast<float*>(w + oy);
    float* fgZ = reinterpret_cast<float*>(w + oz);
    cudaStream_t st = as_stream(stream);
    const CamDev cam = make_cam(cam_host);
    const int ntiles = (P + kTile - 1) / kTile;
    DH_CUDA_CHECK(cudaMemcpyAsync(rigid_dev, rigid_host, sizeof(dh_rigid) * (size_t)B, cudaMemcpyHostToDevice, st));
    dim3 tgrid(ntiles, B);
    fg_count_kernel<<<tgrid, kTileThreads, 0, st>>>(fg_mask, P, ntiles, tile_counts);
    DH_LAUNCH_CHECK();
    fg_compact_kernel<<<tgrid, kTileThreads, 0, st>>>(depth, fg_mask, P, W, ntiles, cam, xs, ys, tile_counts, fg_index, fgX, fgY, fgZ, n_fg,
                                                      nullptr);
    DH_LAUNCH_CHECK();
    fg_centroid_kernel<<<B, 96, 0, st>>>(fgX, fgY, fgZ, n_fg, P, centroid);
    DH_LAUNCH_CHECK();
    dim3 grid((2 * P + 255) / 256, B);
    if (zbuf)
        transform_project_kernel<true><<<grid, 256, 0, st>>>(bg_depth, P, H, W, cam, rigid_dev, xs, ys, fgX, fgY, fgZ, n_fg, centroid,
                                                             pix, zkey, points_out, zbuf);
    else
        transform_project_kernel<false><<<grid, 256, 0, st>>>(bg_depth, P, H, W, cam, rigid_dev, xs, ys, fgX, fgY, fgZ, n_fg, centroid,
                                                              pix, zkey, points_out, nullptr);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_unproject_transform_project(const float* depth, const float* bg_depth, const float* fg_mask, int B, int H, int W,
                                   const dh_camera* cam_host, const dh_rigid* rigid_host, const float* xs, const float* ys,
                                   int32_t* pix, uint64_t* zkey, int32_t* fg_index, int32_t* n_fg, float* centroid,
                                   double* points_out, void* ws, size_t ws_bytes, void* stream) {
    return k1_launch(depth, bg_depth, fg_mask, B, H, W, cam_host, rigid_host, xs, ys, pix, zkey, fg_index, n_fg, centroid, points_out,
                     ws, ws_bytes, nullptr, stream);
}

int dh_unproject_transform_project_splat(const float* depth, const float* bg_depth, const float* fg_mask, int B, int H, int W,
                                         const dh_camera* cam_host, const dh_rigid* rigid_host, const float* xs, const float* ys,
                                         int32_t* pix, uint64_t* zkey, int32_t* fg_index, int32_t* n_fg, float* centroid,
                                         double* points_out, uint64_t* zbuf, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(zbuf && B >= 1 && H >= 1 && W >= 1);
    DH_CUDA_CHECK(cudaMemsetAsync(zbuf, 0xFF, sizeof(uint64_t) * (size_t)B * H * W, as_stream(stream)));
    return k1_launch(depth, bg_depth, fg_mask, B, H, W, cam_host, rigid_host, xs, ys, pix, zkey, fg_index, n_fg, centroid, points_out,
                     ws, ws_bytes, zbuf, stream);
}

int dh_transform_point_cloud(const float* points, const float* mask, int N, const dh_rigid* rigid_host, double* out,
                             float* centroid, int32_t* n_masked, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(points && mask && rigid_host && out && centroid && n_masked && ws && N >= 1);
    // workspace = the K1 layout for one "image" of N pixels (+ the compacted index list)
    size_t oc, orr, ox, oy, oz;
    const size_t k1 = k1_ws_layout(1, N, &oc, &orr, &ox, &oy, &oz);
    if (ws_bytes < k1 + sizeof(int32_t) * (size_t)N) return DH_ERR_WORKSPACE;
    char* w = static_cast<char*>(ws);
    int32_t* tile_counts = reinterpret_cast<int32_t*>(w + oc);
    dh_rigid* rigid_dev = reinterpret_cast<dh_rigid*>(w + orr);
    float* fgX = reinterpret_cast<float*>(w + ox);
    float* fgY = reinterpret_cast<float*>(w + oy);
    float* fgZ = reinterpret_cast<float*>(w + oz);
    int32_t* index = reinterpret_cast<int32_t*>(w + k1);
    cudaStream_t st = as_stream(stream);
    const int ntiles = (N + kTile - 1) / kTile;
    CamDev cam;
    memset(&cam, 0, sizeof(cam));
    DH_CUDA_CHECK(cudaMemcpyAsync(rigid_dev, rigid_host, sizeof(dh_rigid), cudaMemcpyHostToDevice, st));
    fg_count_kernel<<<dim3(ntiles, 1), kTileThreads, 0, st>>>(mask, N, ntiles, tile_counts);
    DH_LAUNCH_CHECK();
    fg_compact_kernel<<<dim3(ntiles, 1), kTileThreads, 0, st>>>(nullptr, mask, N, N, ntiles, cam, nullptr, nullptr, tile_counts, index,
                                                               fgX, fgY, fgZ, n_masked, points);
    DH_LAUNCH_CHECK();
    fg_centroid_kernel<<<1, 96, 0, st>>>(fgX, fgY, fgZ, n_masked, N, centroid);
    DH_LAUNCH_CHECK();
    rigid_all_kernel<<<(N + 255) / 256, 256, 0, st>>>(points, N, rigid_dev, centroid, out);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

size_t dh_transform_point_cloud_workspace_bytes(int N) {
    if (N < 1) return 0;
    size_t a, b, c, d, e;
    return k1_ws_layout(1, N, &a, &b, &c, &d, &e) + sizeof(int32_t) * (size_t)N;
}

int dh_project_points(const double* points, int N, int H, int W, const dh_camera* cam_host, int32_t* pix, uint64_t* zkey,
                      int32_t* u, int32_t* v, void* stream) {
    DH_REQUIRE(points && cam_host && pix && zkey && N >= 0 && H >= 1 && W >= 1);
    if (N == 0) return DH_OK;
    project_points_kernel<<<(N + 255) / 256, 256, 0, as_stream(stream)>>>(points, N, H, W, make_cam(cam_host), pix, zkey, u, v);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
