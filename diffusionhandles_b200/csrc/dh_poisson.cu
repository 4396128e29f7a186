// Poisson hole fill of the edited disparity (depth_transform.py:346-363, :535-587).
//
// The reference assembles the masked 5-point Laplacian (diag 4, -1 towards unknown neighbours, known
// neighbours moved to the right-hand side, image-border neighbours simply absent) and solves it with
// SuperLU in fp64.  Here: conjugate gradients in fp64 on the same SPD system, one CTA per edit, with
// deterministic (fixed-tree) reductions.  Parity is tolerance based (|x - x_ref| <= 1e-3 on a 0..255
// disparity after the fp32 cast; typically 0-1 ulp).
#include "dh_common.cuh"

namespace dh {

constexpr int kPoissonThreads = 1024;

__device__ __forceinline__ double block_sum(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane_id() == 0) sm[warp_id()] = v;
    __syncthreads();
    double t = sm[lane_id()];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
    __syncthreads();
    return t;
}

__device__ __forceinline__ bool bit_at(const uint32_t* bits, int wpr, int row, int col) {
    return (bits[row * wpr + (col >> 5)] >> (col & 31)) & 1u;
}

__global__ void __launch_bounds__(kPoissonThreads) poisson_cg_kernel(
    const float* __restrict__ image, const uint32_t* __restrict__ mask_a, const uint32_t* __restrict__ mask_b, int H, int W,
    int wpr, float* __restrict__ out, int max_iter, double rel_tol, int32_t* __restrict__ iters_out,
    uint32_t* __restrict__ ws_mask, int32_t* __restrict__ ws_list, double* __restrict__ ws_x, double* __restrict__ ws_r,
    double* __restrict__ ws_p, double* __restrict__ ws_ap) {
    __shared__ double red[32];
    __shared__ int scan_smem[33];
    const int e = blockIdx.x, tid = threadIdx.x, P = H * W, nwords = H * wpr;
    const float* img = image + (size_t)e * P;
    float* o = out + (size_t)e * P;
    uint32_t* mask = ws_mask + (size_t)e * nwords;
    int32_t* list = ws_list + (size_t)e * P;
    double* x = ws_x + (size_t)e * P;
    double* r = ws_r + (size_t)e * P;
    double* p = ws_p + (size_t)e * P;
    double* ap = ws_ap + (size_t)e * P;

    for (int q = tid; q < P; q += blockDim.x) o[q] = img[q];
    // inpaint mask = a XOR b, compacted to a list of unknown pixels (row-major)
    const int wpt = (nwords + blockDim.x - 1) / blockDim.x;
    const int w0 = tid * wpt, w1 = min(nwords, w0 + wpt);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) {
        const uint32_t m = mask_a[(size_t)e * nwords + w] ^ (mask_b ? mask_b[(size_t)e * nwords + w] : 0u);
        mask[w] = m;
        cnt += __popc(m);
    }
    int n;
    int pos = block_exclusive_scan(cnt, scan_smem, n);
    for (int w = w0; w < w1; ++w) {
        uint32_t m = mask[w];
        const int row = w / wpr, cb = (w - row * wpr) * 32;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            list[pos++] = row * W + cb + b;
        }
    }
    __syncthreads();
    if (n == 0) {
        if (tid == 0 && iters_out) iters_out[e] = 0;
        return;
    }
    // right-hand side: sum of the known in-image neighbours (fp64 accumulation of fp32 values)
    double bb_part = 0.0;
    for (int k = tid; k < n; k += blockDim.x) {
        const int q = list[k], row = q / W, col = q - row * W;
        double b = 0.0;
        if (row > 0 && !bit_at(mask, wpr, row - 1, col)) b += (double)img[q - W];
        if (row < H - 1 && !bit_at(mask, wpr, row + 1, col)) b += (double)img[q + W];
        if (col > 0 && !bit_at(mask, wpr, row, col - 1)) b += (double)img[q - 1];
        if (col < W - 1 && !bit_at(mask, wpr, row, col + 1)) b += (double)img[q + 1];
        x[q] = 0.0; r[q] = b; p[q] = b;
        bb_part += b * b;
    }
    double rr = block_sum(bb_part, red);
    const double stop = rel_tol * rel_tol * rr;
    if (max_iter <= 0) max_iter = 20000;
    int it = 0;
    while (it < max_iter && rr > stop && rr > 0.0) {
        double pap_part = 0.0;
        for (int k = tid; k < n; k += blockDim.x) {
            const int q = list[k], row = q / W, col = q - row * W;
            double a = 4.0 * p[q];
            if (row > 0 && bit_at(mask, wpr, row - 1, col)) a -= p[q - W];
            if (row < H - 1 && bit_at(mask, wpr, row + 1, col)) a -= p[q + W];
            if (col > 0 && bit_at(mask, wpr, row, col - 1)) a -= p[q - 1];
            if (col < W - 1 && bit_at(mask, wpr, row, col + 1)) a -= p[q + 1];
            ap[q] = a;
            pap_part += p[q] * a;
        }
        const double pap = block_sum(pap_part, red);
        if (!(pap > 0.0)) break;
        const double alpha = rr / pap;
        double rr_part = 0.0;
        for (int k = tid; k < n; k += blockDim.x) {
            const int q = list[k];
            x[q] += alpha * p[q];
            const double rn = r[q] - alpha * ap[q];
            r[q] = rn;
            rr_part += rn * rn;
        }
        const double rr_new = block_sum(rr_part, red);
        const double beta = rr_new / rr;
        rr = rr_new;
        for (int k = tid; k < n; k += blockDim.x) {
            const int q = list[k];
            p[q] = r[q] + beta * p[q];
        }
        __syncthreads();
        ++it;
    }
    for (int k = tid; k < n; k += blockDim.x) {
        const int q = list[k];
        o[q] = (float)x[q];
    }
    if (tid == 0 && iters_out) iters_out[e] = it;
}

static size_t poisson_layout(int B, int H, int W, size_t off[6]) {
    const size_t P = (size_t)H * W, nwords = (size_t)H * ((W + 31) / 32);
    size_t o = 0;
    off[0] = o; o = align_up(o + sizeof(uint32_t) * B * nwords, 256);
    off[1] = o; o = align_up(o + sizeof(int32_t) * B * P, 256);
    for (int i = 2; i < 6; ++i) { off[i] = o; o = align_up(o + sizeof(double) * B * P, 256); }
    return o;
}

}  // namespace dh

using namespace dh;

extern "C" {

size_t dh_poisson_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    size_t off[6];
    return poisson_layout(B, H, W, off);
}

int dh_poisson_fill(const float* image, const uint32_t* mask_a_bits, const uint32_t* mask_b_bits, int B, int H, int W,
                    float* out, int max_iter, double rel_tol, int32_t* iters_out, void* ws, size_t ws_bytes, void* stream) {
    DH_REQUIRE(image && mask_a_bits && out && ws && B >= 1 && H >= 1 && W >= 1 && image != out);
    size_t off[6];
    if (ws_bytes < poisson_layout(B, H, W, off)) return DH_ERR_WORKSPACE;
    char* w = static_cast<char*>(ws);
    poisson_cg_kernel<<<B, kPoissonThreads, 0, as_stream(stream)>>>(
        image, mask_a_bits, mask_b_bits, H, W, (W + 31) / 32, out, max_iter, rel_tol > 0 ? rel_tol : 1e-13, iters_out,
        reinterpret_cast<uint32_t*>(w + off[0]), reinterpret_cast<int32_t*>(w + off[1]), reinterpret_cast<double*>(w + off[2]),
        reinterpret_cast<double*>(w + off[3]), reinterpret_cast<double*>(w + off[4]), reinterpret_cast<double*>(w + off[5]));
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
