// CNN_3D encoder glue (sm_100a): the per-token tail of a residual conv block,
//     out = shortcut + ReLU(LayerNorm_16(y) * gamma + beta)                        models/encoder/CNN_3D.py:129-147
// on channel-last tokens [N, V, THW, 16] (y = the 3x3x3 replicate conv output, evaluated by the conv kernels of conv_tc.cu /
// conv16_umma.cu), one launch for all V variables (gamma / beta: [V][16]).  HBM-bound streaming kernels: a thread owns one token
// (16 channels = 64 bytes fp32 in registers, LayerNorm is thread-local), a warp writes its 32 tokens through the coalescing
// helper of common.cuh.  Forward optionally writes a bf16 copy of `out` (the next conv reads bf16, idee_conv_desc.x_dtype).
// Backward: g_y = LN_bwd(gamma * g_out * [act > 0]), d_gamma = sum g_act * xn, d_beta = sum g_act; the gradient w.r.t. the
// shortcut is g_out itself (the caller adds it).  Per-CTA partials in double, deterministic finalize.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace {

constexpr int C = 16, NT = 256, NB_MAX = 1024;

__global__ void __launch_bounds__(NT)
ln_act_res_fwd_kernel(const float* __restrict__ y, const float* __restrict__ shortcut, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ out, __nv_bfloat16* __restrict__ out16, int N, int V, int64_t thw) {
    const int v = blockIdx.y;
    __shared__ float gb[2 * C];
    if (threadIdx.x < C) { gb[threadIdx.x] = gamma[v * C + threadIdx.x]; gb[C + threadIdx.x] = beta[v * C + threadIdx.x]; }
    __syncthreads();
    const int64_t ntok = (int64_t)N * thw;
    for (int64_t tok = (int64_t)blockIdx.x * NT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * NT) {
        const int64_t n = tok / thw;
        const int64_t off = ((n * V + v) * thw + (tok - n * thw)) * C;
        float yr[C], xn[C], sc[C];
        load16(yr, y + off);
        load16(sc, shortcut + off);
        ln16(yr, xn);
#pragma unroll
        for (int c = 0; c < C; ++c) sc[c] += fmaxf(xn[c] * gb[c] + gb[C + c], 0.f);
        store16(out + off, sc);
        if (out16 != nullptr) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 t = __floats2bfloat162_rn(sc[2 * i], sc[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&t);
            }
            st8u(out16 + off, pk);
        }
    }
}

// partial layout per CTA: d_gamma[16] | d_beta[16] (double)
__global__ void __launch_bounds__(NT)
ln_act_res_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ gout, float* __restrict__ gy, double* __restrict__ partials, int N, int V, int64_t thw) {
    const int v = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float gb[2 * C];
    __shared__ double red[NT / 32][2 * C];
    if (threadIdx.x < C) { gb[threadIdx.x] = gamma[v * C + threadIdx.x]; gb[C + threadIdx.x] = beta[v * C + threadIdx.x]; }
    __syncthreads();
    float dg[C], db[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { dg[c] = 0.f; db[c] = 0.f; }
    const int64_t ntok = (int64_t)N * thw;
    for (int64_t tok = (int64_t)blockIdx.x * NT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * NT) {
        const int64_t n = tok / thw;
        const int64_t off = ((n * V + v) * thw + (tok - n * thw)) * C;
        float yr[C], xn[C], go[C], gxn[C], gx[C];
        load16(yr, y + off);
        load16(go, gout + off);
        const float rstd = ln16(yr, xn);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float ga = (xn[c] * gb[c] + gb[C + c] > 0.f) ? go[c] : 0.f;       // ReLU backward
            dg[c] += ga * xn[c];
            db[c] += ga;
            gxn[c] = ga * gb[c];
        }
        ln16_bwd(gxn, xn, rstd, gx);
        store16(gy + off, gx);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double a = warp_sum_d((double)dg[c]), b = warp_sum_d((double)db[c]);
        if (lane == 0) { red[warp][c] = a; red[warp][C + c] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[w][threadIdx.x];
        partials[((int64_t)v * gridDim.x + blockIdx.x) * 2 * C + threadIdx.x] = s;
    }
}

__global__ void ln_act_res_finalize_kernel(const double* __restrict__ partials, int nblocks, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int v = blockIdx.x, k = threadIdx.x;
    double a = 0.0;
    for (int b = 0; b < nblocks; ++b) a += partials[((int64_t)v * nblocks + b) * 2 * C + k];
    if (k < C) dgamma[v * C + k] = (float)a; else dbeta[v * C + k - C] = (float)a;
}

int grid_for(int64_t ntok, int V) {
    int64_t nb = (ntok + NT - 1) / NT;
    const int64_t cap = (int64_t)idee_num_sms() * 8 / (V > 0 ? V : 1);
    if (nb > cap) nb = cap;
    if (nb > NB_MAX) nb = NB_MAX;
    return nb < 1 ? 1 : (int)nb;
}

}  // namespace

extern "C" int idee_ln_act_res_fwd(const float* y, const float* shortcut, const float* gamma, const float* beta, float* out, void* out_bf16,
                                   int N, int V, int64_t THW, int Cch, void* stream) {
    IDEE_REQUIRE(Cch == C, "ln_act_res_fwd: only 16 channels are built (got %d)", Cch);
    IDEE_REQUIRE(N > 0 && V > 0 && THW > 0, "ln_act_res_fwd: empty tensor");
    ln_act_res_fwd_kernel<<<dim3(grid_for((int64_t)N * THW, V), V), NT, 0, (cudaStream_t)stream>>>(y, shortcut, gamma, beta, out,
                                                                                                   (__nv_bfloat16*)out_bf16, N, V, THW);
    IDEE_LAUNCH_CHECK("ln_act_res_fwd");
    return 0;
}

extern "C" size_t idee_ln_act_res_bwd_workspace_bytes(int V) { return sizeof(double) * (size_t)V * NB_MAX * 2 * C; }

extern "C" int idee_ln_act_res_bwd(const float* y, const float* gamma, const float* beta, const float* gout, float* gy, float* dgamma,
                                   float* dbeta, int N, int V, int64_t THW, int Cch, void* workspace, size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(Cch == C, "ln_act_res_bwd: only 16 channels are built (got %d)", Cch);
    IDEE_REQUIRE(workspace_bytes >= idee_ln_act_res_bwd_workspace_bytes(V), "ln_act_res_bwd: workspace too small");
    const int nb = grid_for((int64_t)N * THW, V);
    ln_act_res_bwd_kernel<<<dim3(nb, V), NT, 0, (cudaStream_t)stream>>>(y, gamma, beta, gout, gy, (double*)workspace, N, V, THW);
    IDEE_LAUNCH_CHECK("ln_act_res_bwd");
    ln_act_res_finalize_kernel<<<V, 2 * C, 0, (cudaStream_t)stream>>>((const double*)workspace, nb, dgamma, dbeta);
    IDEE_LAUNCH_CHECK("ln_act_res_finalize");
    return 0;
}
