// the elementwise steps guided_inference runs around the U-Net, one launch each instead of
// two (latent update) and nine (classifier-free guidance + DDIM update) eager launches:
//   guided_stable_diffuser.py:434       latents = latents - grad_cond * 0.1
//   guided_stable_diffuser.py:470-471   noise_pred = uncond + 7.5 * (text - uncond)
//   guided_stable_diffuser.py:474       latents = scheduler.step(noise_pred, t, latents)   (diffusers 0.23 DDIMScheduler.step,
//                                       epsilon prediction, eta = 0, clip_sample = False - the reference's constructor, :31-32)
// The latents are 4 x 64 x 64 floats: the kernels are launch-latency bound, so the point is the launch count.  Every product and
// sum is rounded to fp32 on its own (this file is compiled with -fmad=false), i.e. the value sequence of the separate torch ops.
#include "dh_common.cuh"

namespace dh {

struct DdimParams {
    const float* uncond;
    const float* text;
    const float* sample;
    float* out;
    float* eps_out;
    size_t n;
    float scale, sqrt_beta_t, sqrt_alpha_t, inv_sqrt_alpha_t, sqrt_alpha_prev, sqrt_beta_prev;
    int reciprocal;
};

__device__ __forceinline__ float ddim_one(const DdimParams& p, float u, float t, float x, bool cfg, float& eps) {
    eps = cfg ? u + p.scale * (t - u) : u;
    const float num = x - p.sqrt_beta_t * eps;
    const float x0 = p.reciprocal ? num * p.inv_sqrt_alpha_t : num / p.sqrt_alpha_t;
    return p.sqrt_alpha_prev * x0 + p.sqrt_beta_prev * eps;
}

template <bool kVec>
__global__ void __launch_bounds__(256) cfg_ddim_kernel(const DdimParams p) {
    const bool cfg = p.text != nullptr;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kVec) {
        const size_t n4 = p.n >> 2;
        if (i < n4) {
            const float4 u = reinterpret_cast<const float4*>(p.uncond)[i];
            const float4 t = cfg ? reinterpret_cast<const float4*>(p.text)[i] : u;
            const float4 x = reinterpret_cast<const float4*>(p.sample)[i];
            float4 e, o;
            o.x = ddim_one(p, u.x, t.x, x.x, cfg, e.x);
            o.y = ddim_one(p, u.y, t.y, x.y, cfg, e.y);
            o.z = ddim_one(p, u.z, t.z, x.z, cfg, e.z);
            o.w = ddim_one(p, u.w, t.w, x.w, cfg, e.w);
            reinterpret_cast<float4*>(p.out)[i] = o;
            if (p.eps_out) reinterpret_cast<float4*>(p.eps_out)[i] = e;
        }
        const size_t tail = (n4 << 2) + i;            // at most three trailing elements
        if (i < (p.n & 3)) {
            float e;
            p.out[tail] = ddim_one(p, p.uncond[tail], cfg ? p.text[tail] : 0.f, p.sample[tail], cfg, e);
            if (p.eps_out) p.eps_out[tail] = e;
        }
    } else if (i < p.n) {
        float e;
        p.out[i] = ddim_one(p, p.uncond[i], cfg ? p.text[i] : 0.f, p.sample[i], cfg, e);
        if (p.eps_out) p.eps_out[i] = e;
    }
}

template <bool kVec>
__global__ void __launch_bounds__(256) latent_step_kernel(const float* lat, const float* grad, float step, float* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kVec) {
        const size_t n4 = n >> 2;
        if (i < n4) {
            const float4 a = reinterpret_cast<const float4*>(lat)[i];
            const float4 g = reinterpret_cast<const float4*>(grad)[i];
            float4 o;
            o.x = a.x - g.x * step; o.y = a.y - g.y * step; o.z = a.z - g.z * step; o.w = a.w - g.w * step;
            reinterpret_cast<float4*>(out)[i] = o;
        }
        const size_t tail = (n4 << 2) + i;
        if (i < (n & 3)) out[tail] = lat[tail] - grad[tail] * step;
    } else if (i < n) {
        out[i] = lat[i] - grad[i] * step;
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline unsigned blocks_for(size_t n, bool vec) {
    size_t threads = vec ? (n >> 2) : n;
    if (threads < 4) threads = 4;                    // the vector kernel's tail needs up to three threads
    return (unsigned)((threads + 255) / 256);
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_latent_step(const float* latents, const float* grad, float step_size, float* out, size_t n, void* stream) {
    if (n == 0) return DH_OK;                        // (an empty tensor has a null data pointer)
    DH_REQUIRE(latents && grad && out);
    DH_REQUIRE(n < ((size_t)1 << 40));
    const bool vec = aligned16(latents) && aligned16(grad) && aligned16(out);
    // `out` may be `latents` itself (every element is read and written by the same thread); a partial overlap is rejected
    if (out != latents) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(latents), o = reinterpret_cast<uintptr_t>(out), bytes = n * sizeof(float);
        DH_REQUIRE(o + bytes <= a || a + bytes <= o);
    }
    if (vec) latent_step_kernel<true><<<blocks_for(n, true), 256, 0, as_stream(stream)>>>(latents, grad, step_size, out, n);
    else latent_step_kernel<false><<<blocks_for(n, false), 256, 0, as_stream(stream)>>>(latents, grad, step_size, out, n);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

int dh_cfg_ddim_step(const float* noise_uncond, const float* noise_text, const float* sample, const dh_ddim_coeffs* coeffs_host,
                     float* out, float* eps_out, size_t n, void* stream) {
    DH_REQUIRE(coeffs_host && coeffs_host->sqrt_alpha_t > 0.f);
    if (n == 0) return DH_OK;
    DH_REQUIRE(noise_uncond && sample && out);
    DH_REQUIRE(n < ((size_t)1 << 40));
    DdimParams p;
    p.uncond = noise_uncond; p.text = noise_text; p.sample = sample; p.out = out; p.eps_out = eps_out; p.n = n;
    p.scale = coeffs_host->guidance_scale;
    p.sqrt_beta_t = coeffs_host->sqrt_beta_t;
    p.sqrt_alpha_t = coeffs_host->sqrt_alpha_t;
    p.inv_sqrt_alpha_t = 1.0f / coeffs_host->sqrt_alpha_t;      // fp32 reciprocal, as ATen computes it for a host-scalar divisor
    p.sqrt_alpha_prev = coeffs_host->sqrt_alpha_prev;
    p.sqrt_beta_prev = coeffs_host->sqrt_beta_prev;
    p.reciprocal = coeffs_host->divide_by_reciprocal;
    const bool vec = aligned16(noise_uncond) && (!noise_text || aligned16(noise_text)) && aligned16(sample) && aligned16(out) &&
                     (!eps_out || aligned16(eps_out));
    if (vec) cfg_ddim_kernel<true><<<blocks_for(n, true), 256, 0, as_stream(stream)>>>(p);
    else cfg_ddim_kernel<false><<<blocks_for(n, false), 256, 0, as_stream(stream)>>>(p);
    DH_LAUNCH_CHECK();
    return DH_OK;
}

}  // extern "C"
