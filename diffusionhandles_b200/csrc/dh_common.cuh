// Shared device/host helpers of libdiffhandles_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/dh_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdiffhandles_b200 is written for sm_100a (B200) only"
#endif

namespace dh {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = static_cast<int>(e);
    return DH_ERR_CUDA;
}

#define DH_CUDA_CHECK(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::dh::cuda_fail(_e);    \
    } while (0)

#define DH_LAUNCH_CHECK() DH_CUDA_CHECK(cudaPeekAtLastError())

#define DH_REQUIRE(cond)                                      \
    do {                                                      \
        if (!(cond)) return DH_ERR_INVALID_ARGUMENT;          \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr uint32_t kNoWinner = 0xFFFFFFFFu;
constexpr uint64_t kEmptyZ = 0xFFFFFFFFFFFFFFFFull;

// Order-preserving map fp64 -> uint64 (negative values sort below positive ones; -0 is folded into +0
// by the caller).  Inverse below.
__host__ __device__ inline uint64_t z_to_key(double z) {
#ifdef __CUDA_ARCH__
    uint64_t b = static_cast<uint64_t>(__double_as_longlong(z));
#else
    uint64_t b;
    memcpy(&b, &z, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ inline double key_to_z(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double(static_cast<long long>(b));
}

// Order-preserving map fp32 -> uint32 for atomicMin/atomicMax on floats.
__device__ inline uint32_t f_to_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ inline float key_to_f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

__device__ inline int warp_id() { return threadIdx.x >> 5; }
__device__ inline int lane_id() { return threadIdx.x & 31; }

// Exclusive scan of one int per thread over a CTA of up to 1024 threads; returns the exclusive prefix,
// writes the CTA total to `total`.  `smem` needs 33 ints.
__device__ inline int block_exclusive_scan(int v, int* smem, int& total) {
    const int lane = lane_id(), wid = warp_id();
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int w = lane < nw ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += n;
        }
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = smem[wid] + inc - v;
    total = smem[32];
    __syncthreads();
    return res;
}

}  // namespace dh
