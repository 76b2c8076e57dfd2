// Lookup-free 1-bit quantiser ("binary driver codebook"), fused forward + aux losses + backward (sm_100a).
//
// Replaces LFQ.forward (models/codebook/LFQ.py:183-307) for dim=16, codebook_size=2 (-> codebook_dim 1,
// project_in Linear(16,1), project_out Linear(1,16), codebook {-1,+1}); always fp32 like the reference
// (LFQ.py:183,199).  Pure HBM-bound streaming kernel: 64 B/token in, 64 B (z_q) + 8 B (int64 index) out.
//
// forward (per token): s = w_in.z + b_in ; q = s>0 ? +1 : -1 (tie -> -1 -> index 0) ; x = q (the reference's s + (q - s).detach() equals q up to one ulp; the exact value keeps z_q of a code identical for every token)
//   index = x>0 ; z_q = x*w_out + b_out ; train: p = softmax([-200 s, +200 s]), sums of per-token entropy,
//   of p (for the codebook entropy) and of (s-q)^2, reduced in double per CTA then by a 1-thread finalize.
// backward: single pass, g_s = <w_out, g_zq> (straight-through) + g_aux * d(aux)/ds, using the saved mean prob.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace {

constexpr int C = 16;
constexpr int LFQ_THREADS = 256;
constexpr float LOG_EPS = 1e-5f;   // LFQ.py:52

__device__ __forceinline__ void probs(float s, float inv_temp, float& p0, float& p1) {
    // softmax over logits (-2*inv_temp*s*c) for codes c = {-1,+1}   (LFQ.py:239-240)
    const float l1 = 2.f * inv_temp * s, l0 = -l1;
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv = 1.f / (e0 + e1);
    p0 = e0 * inv; p1 = e1 * inv;
}
__device__ __forceinline__ float ent_term(float p) { return -p * logf(fmaxf(p, LOG_EPS)); }
// d/dp [-p log(clamp(p, eps))]
__device__ __forceinline__ float ent_grad(float p) { return -logf(fmaxf(p, LOG_EPS)) - (p >= LOG_EPS ? 1.f : 0.f); }

template <int NACC>
__device__ __forceinline__ void block_reduce_store(double* acc, double* out) {
    __shared__ double red[LFQ_THREADS / 32][NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double v = warp_sum_d(acc[a]);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double v = 0.0;
        for (int w = 0; w < LFQ_THREADS / 32; ++w) v += red[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
}

// ZS: z holds the projected scalar s itself ([ntok], project_in already applied by the producer)
template <bool ZS>
__global__ void __launch_bounds__(LFQ_THREADS)
lfq_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w_in, const float* __restrict__ b_in,
               const float* __restrict__ w_out, const float* __restrict__ b_out, float* __restrict__ zq,
               long long* __restrict__ indices, float* __restrict__ xq, double* __restrict__ partials, int64_t ntok, int training,
               float inv_temp, __nv_bfloat16* __restrict__ zq16) {
    float wi[C], wo[C], bo[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { wi[c] = ZS ? 0.f : __ldg(w_in + c); wo[c] = __ldg(w_out + c); bo[c] = __ldg(b_out + c); }
    const float bi = ZS ? 0.f : __ldg(b_in);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // sum entropy, sum p0, sum p1, sum (s-q)^2
    __shared__ __align__(16) float zq_stage[LFQ_THREADS / 32][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // whole warps stay in the loop together (the z_q rows of a warp are written cooperatively); lanes past the end are masked
    for (int64_t tok0 = (int64_t)blockIdx.x * LFQ_THREADS + warp * 32; tok0 < ntok; tok0 += (int64_t)gridDim.x * LFQ_THREADS) {
        const int64_t tok = tok0 + lane;
        const bool live = tok < ntok;
        float s = 0.f;
        if (!live) {}
        else if (ZS) s = __ldg(z + tok);
        else {
            float zr[C];
            load16(zr, z + tok * C);
            s = bi;
#pragma unroll
            for (int c = 0; c < C; ++c) s += wi[c] * zr[c];
        }
        const float q = s > 0.f ? 1.f : -1.f;
        const float x = q;     // == s + (q - s) up to one ulp (LFQ.py:226); the straight-through gradient is applied analytically in the backward pass
        if (live) { indices[tok] = x > 0.f ? 1 : 0; if (xq) xq[tok] = x; }
        float r[C];
#pragma unroll
        for (int c = 0; c < C; ++c) r[c] = x * wo[c] + bo[c];
        store16_warp(zq + tok0 * C, r, zq_stage[warp], lane, (int)min((int64_t)32, ntok - tok0));
        if (zq16) {        // bf16 copy for consumers that round to bf16 anyway: re-read the staged rows, two coalesced 512-byte stores
            const int nrows = (int)min((int64_t)32, ntok - tok0);
            const float* smw = zq_stage[warp];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int f = j * 32 + lane, row = f >> 1, half = f & 1, sw = (row >> 1) & 3;
                const float4 a = *reinterpret_cast<const float4*>(smw + row * 16 + (((2 * half) ^ sw) << 2));
                const float4 b = *reinterpret_cast<const float4*>(smw + row * 16 + (((2 * half + 1) ^ sw) << 2));
                __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
                if (row < nrows)
                    *reinterpret_cast<uint4*>(zq16 + (tok0 + row) * C + half * 8) =
                        make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                                   *reinterpret_cast<uint32_t*>(&h3));
            }
            __syncwarp();
        }
        if (training && live) {
            float p0, p1;
            probs(s, inv_temp, p0, p1);
            acc[0] += (double)(ent_term(p0) + ent_term(p1));
            acc[1] += (double)p0; acc[2] += (double)p1;
            acc[3] += (double)((s - q) * (s - q));
        }
    }
    if (training) block_reduce_store<4>(acc, partials + (int64_t)blockIdx.x * 4);
}

// stats (float[8]): aux, per_sample_entropy, codebook_entropy, commit, mean p0, mean p1, ntok, 0
__global__ void lfq_finalize_kernel(const double* __restrict__ partials, int nblocks, int64_t ntok, float lam_commit,
                                    float lam_ent, float gamma, float* __restrict__ stats) {
    __shared__ double red[8][4];
    double a[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < nblocks; b += 256)
        for (int k = 0; k < 4; ++k) a[k] += partials[(int64_t)b * 4 + k];
    for (int k = 0; k < 4; ++k) a[k] = warp_sum_d(a[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 4; ++k) red[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) { a[k] = 0.0; for (int w = 0; w < 8; ++w) a[k] += red[w][k]; }
        const double n = (double)ntok;
        const float h_tok = (float)(a[0] / n), p0 = (float)(a[1] / n), p1 = (float)(a[2] / n), commit = (float)(a[3] / n);
        const float h_cb = ent_term(p0) + ent_term(p1);
        stats[0] = commit * lam_commit + (lam_ent * h_tok - gamma * h_cb);   // LFQ.py:262,300
        stats[1] = h_tok; stats[2] = h_cb; stats[3] = commit; stats[4] = p0; stats[5] = p1; stats[6] = (float)n; stats[7] = 0.f;
    }
}

constexpr int LFQ_NG = 3 * C + 1;   // g_w_in[16], g_b_in, g_w_out[16], g_b_out[16]

template <bool ZS>
__global__ void __launch_bounds__(LFQ_THREADS)
lfq_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gzq, const float* __restrict__ gxq, const float* __restrict__ g_aux,
               const float* __restrict__ stats, const float* __restrict__ w_in, const float* __restrict__ b_in,
               const float* __restrict__ w_out, float* __restrict__ gz, double* __restrict__ partials, int64_t ntok,
               float lam_commit, float lam_ent, float gamma, float inv_temp) {
    float wi[C], wo[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { wi[c] = ZS ? 0.f : __ldg(w_in + c); wo[c] = __ldg(w_out + c); }
    const float bi = ZS ? 0.f : __ldg(b_in);
    const float ga = g_aux ? __ldg(g_aux) : 0.f;
    const bool gzq_a32 = (reinterpret_cast<uintptr_t>(gzq) & 31) == 0;
    const float invn = 1.f / (float)ntok;
    const float cb_diff = ent_grad(__ldg(stats + 5)) - ent_grad(__ldg(stats + 4));   // f(pbar1) - f(pbar0)
    float a_wi[C], a_wo[C], a_bo[C], a_bi = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { a_wi[c] = 0.f; a_wo[c] = 0.f; a_bo[c] = 0.f; }
    // a thread sees at most 64 tokens (lfq_blocks sizes the grid accordingly), so plain fp32 running sums are exact enough;
    // they are widened to double only for the block / grid reduction (98 fewer live registers than double accumulators)
    for (int64_t tok = (int64_t)blockIdx.x * LFQ_THREADS + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * LFQ_THREADS) {
        float zr[C], gr[C];
        load16_a(gr, gzq + tok * C, gzq_a32);
        float s;
        if (ZS) s = __ldg(z + tok);
        else {
            load16(zr, z + tok * C);
            s = bi;
#pragma unroll
            for (int c = 0; c < C; ++c) s += wi[c] * zr[c];
        }
        const float q = s > 0.f ? 1.f : -1.f;
        const float x = q;
        float gs = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { gs += wo[c] * gr[c]; a_wo[c] += x * gr[c]; a_bo[c] += gr[c]; }
        if (gxq) gs += __ldg(gxq + tok);        // gradient that reached the quantised scalar x directly (rank-1 consumers of z_q)
        float p0, p1;
        probs(s, inv_temp, p0, p1);
        const float dp1 = 2.f * inv_temp * 2.f * p0 * p1;     // d p1 / d s  (= 400 p0 p1 at inv_temp 100)
        const float d_ent = dp1 * (ent_grad(p1) - ent_grad(p0));
        gs += ga * invn * (lam_commit * 2.f * (s - q) + lam_ent * d_ent - gamma * dp1 * cb_diff);
        if (ZS) gz[tok] = gs;                      // gradient w.r.t. the scalar; project_in's gradients come from the producer
        else {
            float r[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { r[c] = gs * wi[c]; a_wi[c] += gs * zr[c]; }
            a_bi += gs;
            store16(gz + tok * C, r);
        }
    }
    double acc[LFQ_NG];
#pragma unroll
    for (int c = 0; c < C; ++c) { acc[c] = a_wi[c]; acc[C + 1 + c] = a_wo[c]; acc[2 * C + 1 + c] = a_bo[c]; }
    acc[C] = a_bi;
    block_reduce_store<LFQ_NG>(acc, partials + (int64_t)blockIdx.x * LFQ_NG);
}

// grads (float[49]): g_w_in[16] | g_b_in | g_w_out[16] | g_b_out[16]
// one CTA per gradient element: 128 threads sum strided subsets of the per-block partials (double), fixed-order tree after
__global__ void __launch_bounds__(128) lfq_bwd_finalize_kernel(const double* __restrict__ partials, int nblocks, float* __restrict__ grads) {
    __shared__ double red[128];
    const int k = blockIdx.x, t = threadIdx.x;
    double a = 0.0;
    for (int b = t; b < nblocks; b += 128) a += partials[(int64_t)b * LFQ_NG + k];
    red[t] = a;
    __syncthreads();
#pragma unroll
    for (int o = 64; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
    }
    if (t == 0) grads[k] = (float)red[0];
}

// Grid-stride kernels with equal work per CTA: the grid is a whole number of resident waves (80 registers x 256 threads: 3 CTAs per
// SM), otherwise the last, partly filled wave costs a full wave time (1184 CTAs = 2.67 waves ran as 3).
int lfq_blocks(int64_t ntok) {
    int64_t nb = (ntok + LFQ_THREADS - 1) / LFQ_THREADS;
    const int64_t wave = (int64_t)idee_num_sms() * 3;
    int64_t cap = wave * 2;
    const int64_t per64 = (ntok + (int64_t)LFQ_THREADS * 64 - 1) / ((int64_t)LFQ_THREADS * 64);   // <= 64 tokens per thread
    if (cap < per64) cap = (per64 + wave - 1) / wave * wave;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (int)nb;
}

}  // namespace

extern "C" size_t idee_lfq_workspace_bytes(int64_t ntok) { return sizeof(double) * (size_t)lfq_blocks(ntok) * LFQ_NG; }

extern "C" int idee_lfq_fwd(const float* z, const float* w_in, const float* b_in, const float* w_out, const float* b_out,
                            float* zq, int64_t* indices, float* xq, float* stats, int64_t ntok, int dim, int codebook_size, int training,
                            float inv_temperature, float lambda_commit, float lambda_entropy, float diversity_gamma,
                            void* workspace, size_t workspace_bytes, void* zq_bf16, void* stream) {
    IDEE_REQUIRE((dim == C || dim == 1) && codebook_size == 2, "lfq_fwd: only dim=16 (or 1: pre-projected scalar), codebook_size=2 is built (got %d, %d)", dim, codebook_size);
    IDEE_REQUIRE(workspace_bytes >= idee_lfq_workspace_bytes(ntok), "lfq_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = lfq_blocks(ntok);
    if (dim == 1) lfq_fwd_kernel<true><<<nb, LFQ_THREADS, 0, st>>>(z, w_in, b_in, w_out, b_out, zq, (long long*)indices, xq, (double*)workspace,
                                                                  ntok, training, inv_temperature, (__nv_bfloat16*)zq_bf16);
    else lfq_fwd_kernel<false><<<nb, LFQ_THREADS, 0, st>>>(z, w_in, b_in, w_out, b_out, zq, (long long*)indices, xq, (double*)workspace, ntok,
                                                            training, inv_temperature, (__nv_bfloat16*)zq_bf16);
    IDEE_LAUNCH_CHECK("lfq_fwd");
    if (training) {
        lfq_finalize_kernel<<<1, 256, 0, st>>>((const double*)workspace, nb, ntok, lambda_commit, lambda_entropy, diversity_gamma, stats);
        IDEE_LAUNCH_CHECK("lfq_finalize");
    }
    return 0;
}

extern "C" int idee_lfq_bwd(const float* z, const float* gzq, const float* gxq, const float* g_aux, const float* stats, const float* w_in,
                            const float* b_in, const float* w_out, float* gz, float* grads, int64_t ntok, float inv_temperature,
                            float lambda_commit, float lambda_entropy, float diversity_gamma, void* workspace,
                            size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(workspace_bytes >= idee_lfq_workspace_bytes(ntok), "lfq_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = lfq_blocks(ntok);
    if (w_in == nullptr) lfq_bwd_kernel<true><<<nb, LFQ_THREADS, 0, st>>>(z, gzq, gxq, g_aux, stats, w_in, b_in, w_out, gz, (double*)workspace, ntok,
                                                                          lambda_commit, lambda_entropy, diversity_gamma, inv_temperature);
    else lfq_bwd_kernel<false><<<nb, LFQ_THREADS, 0, st>>>(z, gzq, gxq, g_aux, stats, w_in, b_in, w_out, gz, (double*)workspace, ntok, lambda_commit,
                                                            lambda_entropy, diversity_gamma, inv_temperature);
    IDEE_LAUNCH_CHECK("lfq_bwd");
    lfq_bwd_finalize_kernel<<<LFQ_NG, 128, 0, st>>>((const double*)workspace, nb, grads);
    IDEE_LAUNCH_CHECK("lfq_bwd_finalize");
    return 0;
}

// ---- scalar planes of the rank-1 form -> one 16-channel channel-last image (input of the joint head's folded first conv) ----
// planes[n, p, c] = xq[n, c, p] for c < V, 1 for c == V, 0 above: one pixel (64 B, two 256-bit stores) per thread, the V scalar reads
// are coalesced across the warp.  The backward gathers the first V channels of the image gradient back into scalar planes.
__global__ void __launch_bounds__(256) rank1_planes_fwd_kernel(const float* __restrict__ xq, float* __restrict__ planes, int V, uint32_t THW) {
    const uint32_t p = blockIdx.x * 256u + threadIdx.x;
    if (p >= THW) return;
    const float* src = xq + (size_t)blockIdx.y * V * THW + p;
    float r[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) r[c] = c < V ? __ldg(src + (size_t)c * THW) : (c == V ? 1.f : 0.f);
    float* dst = planes + ((size_t)blockIdx.y * THW + p) * 16;
    st8f(dst, r);
    st8f(dst + 8, r + 8);
}
__global__ void __launch_bounds__(256) rank1_planes_bwd_kernel(const float* __restrict__ gplanes, float* __restrict__ gxq, int V, uint32_t THW) {
    const uint32_t p = blockIdx.x * 256u + threadIdx.x;
    if (p >= THW) return;
    const float* src = gplanes + ((size_t)blockIdx.y * THW + p) * 16;
    float r[16];
    ldg8f(r, src);
    if (V > 8) ldg8f(r + 8, src + 8);
    float* dst = gxq + (size_t)blockIdx.y * V * THW + p;
#pragma unroll
    for (int c = 0; c < 15; ++c)
        if (c < V) dst[(size_t)c * THW] = r[c];
}

extern "C" int idee_rank1_planes_fwd(const float* xq, float* planes, int N, int V, int64_t THW, void* stream) {
    IDEE_REQUIRE(V >= 1 && V <= 15 && THW > 0 && THW < (1ll << 31) && N >= 1 && N <= 65535, "rank1_planes_fwd: need 1 <= V <= 15, THW < 2^31, N <= 65535");
    rank1_planes_fwd_kernel<<<dim3((unsigned)((THW + 255) / 256), N), 256, 0, (cudaStream_t)stream>>>(xq, planes, V, (uint32_t)THW);
    IDEE_LAUNCH_CHECK("rank1_planes_fwd");
    return 0;
}
extern "C" int idee_rank1_planes_bwd(const float* gplanes, float* gxq, int N, int V, int64_t THW, void* stream) {
    IDEE_REQUIRE(V >= 1 && V <= 15 && THW > 0 && THW < (1ll << 31) && N >= 1 && N <= 65535, "rank1_planes_bwd: need 1 <= V <= 15, THW < 2^31, N <= 65535");
    rank1_planes_bwd_kernel<<<dim3((unsigned)((THW + 255) / 256), N), 256, 0, (cudaStream_t)stream>>>(gplanes, gxq, V, (uint32_t)THW);
    IDEE_LAUNCH_CHECK("rank1_planes_bwd");
    return 0;
}
