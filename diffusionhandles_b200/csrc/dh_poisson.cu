// Poisson hole fill of the edited disparity (depth_transform.py:346-363, :535-587).
//
// The reference assembles the masked 5-point Laplacian (diag 4, -1 towards unknown neighbours, known
// neighbours moved to the right-hand side, image-border neighbours simply absent) and solves it with
// SuperLU in fp64.  Here: conjugate gradients in fp64 on the same SPD system.  One thread-block CLUSTER of
// 8 CTAs (8192 threads) works on one edit: the unknown pixels are compacted into a list (cluster-wide prefix
// over distributed shared memory) and split into eight contiguous runs; every solver vector of a run lives in the
// shared memory of its CTA and the residuals of neighbours owned by another CTA are read through DSMEM.  The
// iteration is the Chronopoulos-Gear arrangement of CG: the stencil is applied to the residual and s = A p follows by
// recurrence, so both inner products come out of one cluster-wide reduction (fixed order: warp tree, CTA tree, then
// the eight CTA partials read through DSMEM by every CTA - deterministic) and an iteration costs two cluster barriers.
// Systems with more than 32k unknowns use the same iteration with L2-resident vectors.
// Parity is tolerance based (|x - x_ref| <= 1e-3 on a 0..255 disparity after the fp32 cast; typically 0-1 ulp).
#include "dh_common.cuh"

#include <cooperative_groups.h>

namespace dh {

namespace cg = cooperative_groups;

constexpr int kPoissonThreads = 1024;
constexpr int kPoissonCluster = 8;
// Shared-memory path: systems of up to kPoissonCluster * kSmemUnknowns unknowns keep every vector of the solver in the
// shared memory of the cluster (56 bytes per unknown); the residuals of neighbours owned by another CTA are read through
// distributed shared memory.  Larger systems use the global-memory (L2) path.
constexpr int kSmemUnknowns = 4096;
constexpr size_t kPoissonSmemBytes = (size_t)kSmemUnknowns * (5 * sizeof(double) + 4 * sizeof(uint32_t));

__device__ __forceinline__ bool bit_at(const uint32_t* bits, int wpr, int row, int col) {
    return (bits[row * wpr + (col >> 5)] >> (col & 31)) & 1u;
}

__device__ __forceinline__ uint32_t cluster_addr(const void* smem_ptr, int rank) {      // shared::cluster address of a CTA's smem
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ double ld_cluster_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_f64x2(uint32_t addr, double a, double b) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}

// Two sums over the whole cluster in one reduction (one cluster barrier), identical in every thread of every CTA.
// All-to-all: every CTA pushes its pair of partials into the slot array of EVERY CTA before the barrier (remote stores are
// fire and forget), so that after the barrier each CTA sums eight LOCAL entries in rank order.
__device__ __forceinline__ void cluster_sum2(cg::cluster_group& cluster, double& a, double& b, double (*warp_sm2)[2],
                                             double (*part_all)[kPoissonCluster][2], int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    }
    if (lane_id() == 0) { warp_sm2[warp_id()][0] = a; warp_sm2[warp_id()][1] = b; }
    __syncthreads();
    if (warp_id() == 0) {
        double ta = warp_sm2[lane_id()][0], tb = warp_sm2[lane_id()][1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ta += __shfl_xor_sync(0xFFFFFFFFu, ta, o);
            tb += __shfl_xor_sync(0xFFFFFFFFu, tb, o);
        }
        if (lane_id() < kPoissonCluster)
            st_cluster_f64x2(cluster_addr(&part_all[slot][cluster.block_rank()][0], lane_id()), ta, tb);
    }
    cluster.sync();
    a = 0.0; b = 0.0;
#pragma unroll
    for (int r = 0; r < kPoissonCluster; ++r) { a += part_all[slot][r][0]; b += part_all[slot][r][1]; }
}

// Sum over the whole cluster, identical in every thread of every CTA.  `slot` alternates between calls so that one
// cluster barrier per reduction is enough.
__device__ __forceinline__ double cluster_sum(cg::cluster_group& cluster, double v, double* warp_sm, double* part_sm, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) warp_sm[warp_id()] = v;
    __syncthreads();
    if (warp_id() == 0) {
        double t = warp_sm[lane_id()];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
        if (lane_id() == 0) part_sm[slot] = t;
    }
    cluster.sync();
    double total = 0.0;
#pragma unroll
    for (int b = 0; b < kPoissonCluster; ++b) total += cluster.map_shared_rank(part_sm, b)[slot];
    return total;
}

__global__ void __cluster_dims__(kPoissonCluster, 1, 1) __launch_bounds__(kPoissonThreads) poisson_cg_kernel(
    const float* __restrict__ image, const uint32_t* __restrict__ mask_a, const uint32_t* __restrict__ mask_b,
    const float* __restrict__ lap_source, int H, int W,
    int wpr, float* __restrict__ out, int max_iter, double rel_tol, int32_t* __restrict__ iters_out,
    uint32_t* __restrict__ ws_mask, int32_t* __restrict__ ws_list, int4* __restrict__ ws_nbr, double* __restrict__ ws_x,
    double* __restrict__ ws_r, double* __restrict__ ws_p, double* __restrict__ ws_ap, double* __restrict__ ws_s) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double warp_sm[32];
    __shared__ double part_sm[4];
    __shared__ double warp_sm2[32][2];
    __shared__ __align__(16) double part_sm2[2][kPoissonCluster][2];
    __shared__ int scan_smem[33];
    __shared__ int count_sm;
    const int e = blockIdx.y, tid = threadIdx.x, P = H * W, nwords = H * wpr;
    const int rank = (int)cluster.block_rank();
    const int gtid = rank * kPoissonThreads + tid, gthreads = kPoissonCluster * kPoissonThreads;
    const float* img = image + (size_t)e * P;
    float* o = out + (size_t)e * P;
    uint32_t* mask = ws_mask + (size_t)e * nwords;
    int32_t* list = ws_list + (size_t)e * P;
    int4* nbr = ws_nbr + (size_t)e * P;
    double* x = ws_x + (size_t)e * P;          // compact (indexed by unknown)
    double* r = ws_r + (size_t)e * P;          // indexed by PIXEL (the stencil reads the neighbours' residuals)
    double* ap = ws_ap + (size_t)e * P;        // compact: w = A r
    double* p = ws_p + (size_t)e * P;          // compact: search direction
    double* sv = ws_s + (size_t)e * P;         // compact: s = A p, maintained by recurrence

    for (int q = gtid; q < P; q += gthreads) o[q] = img[q];
    // inpaint mask = a XOR b, compacted to a row-major list of unknown pixels: every thread owns a contiguous run of words
    const int wpt = (nwords + gthreads - 1) / gthreads;
    const int w0 = min(nwords, gtid * wpt), w1 = min(nwords, w0 + wpt);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) {
        const uint32_t m = mask_a[(size_t)e * nwords + w] ^ (mask_b ? mask_b[(size_t)e * nwords + w] : 0u);
        mask[w] = m;
        cnt += __popc(m);
    }
    int cta_total;
    int pos = block_exclusive_scan(cnt, scan_smem, cta_total);
    if (tid == 0) count_sm = cta_total;
    cluster.sync();
    int n = 0;
    for (int b = 0; b < kPoissonCluster; ++b) {
        const int cb = *cluster.map_shared_rank(&count_sm, b);
        if (b < rank) pos += cb;
        n += cb;
    }
    int32_t* word_prefix = reinterpret_cast<int32_t*>(nbr);      // nwords ints; the smem path does not use the int4 table
    const bool smem_path = n <= kPoissonCluster * kSmemUnknowns;
    for (int w = w0; w < w1; ++w) {
        uint32_t m = mask[w];
        if (smem_path) word_prefix[w] = pos;                     // compact index of the first unknown of this word
        const int row = w / wpr, cb = (w - row * wpr) * 32;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            list[pos++] = row * W + cb + b;
        }
    }
    cluster.sync();          // list and mask complete (and count_sm no longer needed remotely)
    if (n == 0) {
        if (gtid == 0 && iters_out) iters_out[e] = 0;
        return;
    }
    if (smem_path) {
        extern __shared__ __align__(16) unsigned char psm_raw[];
        double* const rs = reinterpret_cast<double*>(psm_raw);            // residual (read by neighbours, also remotely)
        double* const xs_ = rs + kSmemUnknowns;
        double* const ps = xs_ + kSmemUnknowns;
        double* const ss = ps + kSmemUnknowns;
        double* const wv = ss + kSmemUnknowns;
        uint32_t* const nb_s = reinterpret_cast<uint32_t*>(wv + kSmemUnknowns);     // [4][kSmemUnknowns]: rank << 16 | local index
        const int m_own = (n + kPoissonCluster - 1) / kPoissonCluster;              // unknowns per CTA (contiguous, raster order)
        const int k0 = min(n, rank * m_own), k1 = min(n, k0 + m_own), mine = k1 - k0;
        __shared__ double zero_sm;
        if (tid == 0) zero_sm = 0.0;
        const uint32_t zero_addr = cluster_addr(&zero_sm, rank);
        // shared::cluster address of the residual of an unknown pixel (in whichever CTA owns it): one load per neighbour
        // in the iteration, no owner test, no pointer arithmetic
        auto compact_of = [&](int row, int col) -> uint32_t {
            const int w = row * wpr + (col >> 5);
            const int k = word_prefix[w] + __popc(mask[w] & ((1u << (col & 31)) - 1u));
            const int rk = k / m_own;
            return cluster_addr(rs + (k - rk * m_own), rk);
        };
        double bb_part = 0.0;
        for (int j = tid; j < mine; j += kPoissonThreads) {
            const int q = list[k0 + j], row = q / W, col = q - row * W;
            double b = 0.0;
            uint32_t nb4[4] = {zero_addr, zero_addr, zero_addr, zero_addr};   // absent / known neighbours read a local zero
            if (row > 0) { if (bit_at(mask, wpr, row - 1, col)) nb4[0] = compact_of(row - 1, col); else b += (double)img[q - W]; }
            if (row < H - 1) { if (bit_at(mask, wpr, row + 1, col)) nb4[1] = compact_of(row + 1, col); else b += (double)img[q + W]; }
            if (col > 0) { if (bit_at(mask, wpr, row, col - 1)) nb4[2] = compact_of(row, col - 1); else b += (double)img[q - 1]; }
            if (col < W - 1) { if (bit_at(mask, wpr, row, col + 1)) nb4[3] = compact_of(row, col + 1); else b += (double)img[q + 1]; }
            if (lap_source) {
                const float* sdat = lap_source + (size_t)e * P;
                double lap = -4.0 * (double)sdat[q];
                if (row > 0) lap += (double)sdat[q - W];
                if (row < H - 1) lap += (double)sdat[q + W];
                if (col > 0) lap += (double)sdat[q - 1];
                if (col < W - 1) lap += (double)sdat[q + 1];
                b -= (double)(float)lap;
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) nb_s[d * kSmemUnknowns + j] = nb4[d];
            xs_[j] = 0.0; rs[j] = b;
            bb_part += b * b;
        }
        double rr0 = cluster_sum(cluster, bb_part, warp_sm, part_sm, 0);    // (its barrier also publishes the residuals)
        const double stop = rel_tol * rel_tol * rr0;
        if (max_iter <= 0) max_iter = 20000;
        auto spmv_dots = [&](double& g_part, double& d_part) {
            for (int j = tid; j < mine; j += kPoissonThreads) {
                const double rq = rs[j];
                const double n0 = ld_cluster_f64(nb_s[j]), n1 = ld_cluster_f64(nb_s[kSmemUnknowns + j]);
                const double n2 = ld_cluster_f64(nb_s[2 * kSmemUnknowns + j]), n3 = ld_cluster_f64(nb_s[3 * kSmemUnknowns + j]);
                const double a = 4.0 * rq - n0 - n1 - n2 - n3;
                wv[j] = a;
                g_part += rq * rq;
                d_part += a * rq;
            }
        };
        // the same Chronopoulos-Gear iteration as the global-memory path below
        double gam = 0.0, del = 0.0;
        spmv_dots(gam, del);
        cluster_sum2(cluster, gam, del, warp_sm2, part_sm2, 0);
        int slot2 = 1;
        double alpha = del > 0.0 ? gam / del : 0.0;
        for (int j = tid; j < mine; j += kPoissonThreads) { ps[j] = rs[j]; ss[j] = wv[j]; }
        int it = 0;
        while (it < max_iter && gam > stop && gam > 0.0 && alpha > 0.0) {
            // (every CTA has finished reading the old residuals: the reduction barrier of the previous step lies in between)
            for (int j = tid; j < mine; j += kPoissonThreads) {
                xs_[j] += alpha * ps[j];
                rs[j] -= alpha * ss[j];
            }
            cluster.sync();          // the new residual is visible to the neighbours in other CTAs
            double gn = 0.0, dn = 0.0;
            spmv_dots(gn, dn);
            cluster_sum2(cluster, gn, dn, warp_sm2, part_sm2, slot2);
            slot2 ^= 1;
            ++it;
            if (!(gn > stop)) { gam = gn; break; }
            const double beta = gn / gam;
            const double denom = dn - beta * gn / alpha;
            gam = gn;
            if (!(denom > 0.0)) break;
            alpha = gn / denom;
            for (int j = tid; j < mine; j += kPoissonThreads) {
                ps[j] = rs[j] + beta * ps[j];
                ss[j] = wv[j] + beta * ss[j];
            }
        }
        for (int j = tid; j < mine; j += kPoissonThreads) o[list[k0 + j]] = (float)xs_[j];
        if (gtid == 0 && iters_out) iters_out[e] = !(gam > stop) ? it : -it - 1;      // negative: stopped WITHOUT reaching the tolerance
        cluster.sync();          // no CTA exits while another may still read its shared memory
        return;
    }
    // neighbours (pixel index of an unknown neighbour, or -1) and the right-hand side: sum of the known in-image
    // neighbours (fp64 accumulation of fp32 values)
    double bb_part = 0.0;
    for (int k = gtid; k < n; k += gthreads) {
        const int q = list[k], row = q / W, col = q - row * W;
        double b = 0.0;
        int4 nb = make_int4(-1, -1, -1, -1);
        if (row > 0) { if (bit_at(mask, wpr, row - 1, col)) nb.x = q - W; else b += (double)img[q - W]; }
        if (row < H - 1) { if (bit_at(mask, wpr, row + 1, col)) nb.y = q + W; else b += (double)img[q + W]; }
        if (col > 0) { if (bit_at(mask, wpr, row, col - 1)) nb.z = q - 1; else b += (double)img[q - 1]; }
        if (col < W - 1) { if (bit_at(mask, wpr, row, col + 1)) nb.w = q + 1; else b += (double)img[q + 1]; }
        if (lap_source) {
            // utils.py:49-94 (solve_laplacian_depth): b -= laplacian(source)[q], the 5-point stencil with zero padding,
            // evaluated by scipy.ndimage.convolve in fp64 and stored as fp32
            const float* sdat = lap_source + (size_t)e * P;
            double lap = -4.0 * (double)sdat[q];
            if (row > 0) lap += (double)sdat[q - W];
            if (row < H - 1) lap += (double)sdat[q + W];
            if (col > 0) lap += (double)sdat[q - 1];
            if (col < W - 1) lap += (double)sdat[q + 1];
            b -= (double)(float)lap;
        }
        nbr[k] = nb;
        x[k] = 0.0; r[q] = b;
        bb_part += b * b;
    }
    int slot = 0;
    double rr = cluster_sum(cluster, bb_part, warp_sm, part_sm, slot);      // (also publishes r to the other CTAs)
    const double stop = rel_tol * rel_tol * rr;
    if (max_iter <= 0) max_iter = 20000;
    // Conjugate gradients in the Chronopoulos-Gear arrangement: the stencil is applied to the RESIDUAL (w = A r) and
    // s = A p follows by recurrence, so both inner products of an iteration - (r,r) and (w,r) - come out of ONE
    // cluster-wide reduction.  Two cluster barriers per iteration (residual published, reduction) instead of three.
    auto spmv_dots = [&](double& g_part, double& d_part) {
        for (int k = gtid; k < n; k += gthreads) {
            const int q = list[k];
            const int4 nb = nbr[k];
            const double rq = __ldcg(r + q);        // L2 loads: the neighbours may have been written by another CTA
            double a = 4.0 * rq;
            if (nb.x >= 0) a -= __ldcg(r + nb.x);
            if (nb.y >= 0) a -= __ldcg(r + nb.y);
            if (nb.z >= 0) a -= __ldcg(r + nb.z);
            if (nb.w >= 0) a -= __ldcg(r + nb.w);
            ap[k] = a;
            g_part += rq * rq;
            d_part += a * rq;
        }
    };
    double gam = 0.0, del = 0.0;
    spmv_dots(gam, del);
    cluster_sum2(cluster, gam, del, warp_sm2, part_sm2, 0);
    int slot2 = 1;
    double alpha = del > 0.0 ? gam / del : 0.0;
    for (int k = gtid; k < n; k += gthreads) { p[k] = __ldcg(r + list[k]); sv[k] = ap[k]; }
    int it = 0;
    while (it < max_iter && gam > stop && gam > 0.0 && alpha > 0.0) {
        for (int k = gtid; k < n; k += gthreads) {
            const int q = list[k];
            x[k] += alpha * p[k];
            r[q] = __ldcg(r + q) - alpha * sv[k];
        }
        cluster.sync();          // the new residual is visible to the neighbours in other CTAs
        double gn = 0.0, dn = 0.0;
        spmv_dots(gn, dn);
        cluster_sum2(cluster, gn, dn, warp_sm2, part_sm2, slot2);
        slot2 ^= 1;
        ++it;
        if (!(gn > stop)) { gam = gn; break; }
        const double beta = gn / gam;
        const double denom = dn - beta * gn / alpha;
        gam = gn;
        if (!(denom > 0.0)) break;
        alpha = gn / denom;
        for (int k = gtid; k < n; k += gthreads) {
            p[k] = __ldcg(r + list[k]) + beta * p[k];
            sv[k] = ap[k] + beta * sv[k];
        }
    }
    for (int k = gtid; k < n; k += gthreads) o[list[k]] = (float)x[k];
    if (gtid == 0 && iters_out) iters_out[e] = !(gam > stop) ? it : -it - 1;      // negative: stopped WITHOUT reaching the tolerance
    cluster.sync();          // no CTA exits while another may still read its shared memory
}

static size_t poisson_layout(int B, int H, int W, size_t off[8]) {
    const size_t P = (size_t)H * W, nwords = (size_t)H * ((W + 31) / 32);
    size_t o = 0;
    off[0] = o; o = align_up(o + sizeof(uint32_t) * B * nwords, 256);
    off[1] = o; o = align_up(o + sizeof(int32_t) * B * P, 256);
    off[2] = o; o = align_up(o + sizeof(int4) * B * P, 256);
    for (int i = 3; i < 8; ++i) { off[i] = o; o = align_up(o + sizeof(double) * B * P, 256); }
    return o;
}

}  // namespace dh

using namespace dh;

extern "C" {

size_t dh_poisson_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    size_t off[8];
    return poisson_layout(B, H, W, off);
}

int dh_poisson_fill(const float* image, const uint32_t* mask_a_bits, const uint32_t* mask_b_bits, int B, int H, int W,
                    float* out, int max_iter, double rel_tol, int32_t* iters_out, void* ws, size_t ws_bytes, void* stream) {
    return dh_poisson_fill_source(image, mask_a_bits, mask_b_bits, nullptr, B, H, W, out, max_iter, rel_tol, iters_out, ws, ws_bytes,
                                  stream);
}

int dh_poisson_fill_source(const float* image, const uint32_t* mask_a_bits, const uint32_t* mask_b_bits, const float* lap_source,
                           int B, int H, int W, float* out, int max_iter, double rel_tol, int32_t* iters_out, void* ws,
                           size_t ws_bytes, void* stream) {
    DH_REQUIRE(image && mask_a_bits && out && ws && B >= 1 && H >= 1 && W >= 1 && image != out);
    size_t off[8];
    if (ws_bytes < poisson_layout(B, H, W, off)) return DH_ERR_WORKSPACE;
    char* w = static_cast<char*>(ws);
    static bool attr_set[64];          // function attributes are per device
    int dev = 0;
    DH_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DH_ERR_UNSUPPORTED;
    if (!attr_set[dev]) {
        DH_CUDA_CHECK(cudaFuncSetAttribute(poisson_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoissonSmemBytes));
        attr_set[dev] = true;
    }
    poisson_cg_kernel<<<dim3(kPoissonCluster, B), kPoissonThreads, kPoissonSmemBytes, as_stream(stream)>>>(
        image, mask_a_bits, mask_b_bits, lap_source, H, W, (W + 31) / 32, out, max_iter, rel_tol > 0 ? rel_tol : 1e-13, iters_out,
        reinterpret_cast<uint32_t*>(w + off[0]), reinterpret_cast<int32_t*>(w + off[1]), reinterpret_cast<int4*>(w + off[2]),
        reinterpret_cast<double*>(w + off[3]), reinterpret_cast<double*>(w + off[4]), reinterpret_cast<double*>(w + off[5]),
        reinterpret_cast<double*>(w + off[6]), reinterpret_cast<double*>(w + off[7]));
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
