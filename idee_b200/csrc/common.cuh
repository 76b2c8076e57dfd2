// Shared device/host helpers for the idee_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "idee_b200 kernels are written for sm_100a (B200) only"
#endif

// ---- error reporting (C-ABI: every entry point returns int, message via idee_last_error) ----
void idee_set_error(const char* fmt, ...);

#define IDEE_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            idee_set_error(__VA_ARGS__);                          \
            return 1;                                             \
        }                                                         \
    } while (0)

#define IDEE_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            idee_set_error("%s: kernel launch failed: %s", name, cudaGetErrorString(e__)); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

#define IDEE_CUDA(call, name)                                                        \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            idee_set_error("%s: %s failed: %s", name, #call, cudaGetErrorString(e__)); \
            return 2;                                                                \
        }                                                                            \
    } while (0)

int idee_num_sms();  // cached SM count of the current device

// n / d and n % d for n < 2^31 by multiply-high with m = ceil(2^32 / d) (0xFFFFFFFF for d == 1) and one correction step each way
struct FastDiv {
    uint32_t d, m;
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        uint32_t qq = __umulhi(n, m);
        int rr = (int)(n - qq * d);
        if (rr < 0) { --qq; rr += (int)d; }
        if (rr >= (int)d) { ++qq; rr -= (int)d; }
        q = qq; r = (uint32_t)rr;
    }
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)d;
    f.m = d > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d) : 0xFFFFFFFFu;
    return f;
}

// ---- device helpers ----
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// 256-bit global accesses (sm_100: LDG/STG.256): one full 32-byte sector per thread and instruction; p must be 32-byte aligned
__device__ __forceinline__ void st8u(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
                 "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void st8f(float* p, const float* r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]),
                 "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
__device__ __forceinline__ void ldg8f(float* r, const float* p) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]),
                 "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
__device__ __forceinline__ void load16(float* r, const float* p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 v = ldg4(p + 4 * i);
        r[4 * i + 0] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}
// same with two 256-bit loads when the row is 32-byte aligned (a32 is uniform: derived from the base pointer)
__device__ __forceinline__ void load16_a(float* r, const float* p, bool a32) {
    if (a32) { ldg8f(r, p); ldg8f(r + 8, p + 8); } else load16(r, p);
}
__device__ __forceinline__ void store16(float* p, const float* r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) st4(p + 4 * i, make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
}
// Coalesced variant for a warp whose lanes hold 16 consecutive floats of 32 CONSECUTIVE rows (64 B each): the rows are staged
// in the warp's 2 KB shared-memory region (XOR-swizzled 16-byte chunks, conflict free both ways) and written back as four
// fully coalesced 512-byte stores instead of four stores that each touch 32 half-filled sectors.
// gwarp: address of lane 0's row; nrows: rows of this warp that exist (rows >= nrows are not written).
__device__ __forceinline__ void store16_warp(float* gwarp, const float* r, float* smw, int lane, int nrows) {
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float4*>(smw + lane * 16 + ((k ^ sw) << 2)) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int row = j * 8 + (lane >> 2), part = lane & 3;
        const float4 v = *reinterpret_cast<const float4*>(smw + row * 16 + ((part ^ ((row >> 1) & 3)) << 2));
        if (row < nrows) st4(gwarp + (j * 32 + lane) * 4, v);
    }
    __syncwarp();
}
__device__ __forceinline__ void zero16(float* r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exact (erf) GELU and its derivative, as torch.nn.GELU() default (Swin_3D.py:27,32)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// bf16 path: GELU via the hardware tanh (MUFU.TANH): Phi(x) ~= 0.5 (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))).
// |gelu_tanh - gelu_erf| <= 5e-4 absolute, an order of magnitude below the bf16 rounding the value receives as the next
// MMA operand; the fp32 path keeps the exact erf form.  gelu_fast_grad also returns the derivative Phi(x) + x phi(x).
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
    const float u = x * (0.7978845608028654f + 0.0356774081363001f * x * x);
    const float hx = 0.5f * x;
    return hx + hx * tanh_approx(u);
}
__device__ __forceinline__ void gelu_fast_grad(float x, float& y, float& dy) {
    const float x2 = x * x;
    const float u = x * (0.7978845608028654f + 0.0356774081363001f * x2);
    const float cdf = 0.5f + 0.5f * tanh_approx(u);
    y = x * cdf;
    dy = cdf + x * (0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * x2));   // x * phi(x), exp(-x^2/2) = 2^(-x^2 log2(e)/2)
}

// LayerNorm over 16 channels, eps 1e-5, no affine (Swin_3D.py:214,220,469): returns rstd, writes xn
__device__ __forceinline__ float ln16(const float* x, float* xn) {
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) mu += x[i];
    mu *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = x[i] - mu; xn[i] = d; var += d * d; }
    const float rstd = 1.f / sqrtf(var * (1.f / 16.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) xn[i] *= rstd;
    return rstd;
}
// backward of the above: g_x = rstd * (g - mean(g) - xn * mean(g*xn))
__device__ __forceinline__ void ln16_bwd(const float* g, const float* xn, float rstd, float* gx) {
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { m1 += g[i]; m2 += g[i] * xn[i]; }
    m1 *= (1.f / 16.f); m2 *= (1.f / 16.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) gx[i] = rstd * (g[i] - m1 - xn[i] * m2);
}
