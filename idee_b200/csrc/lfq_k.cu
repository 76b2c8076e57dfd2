// Lookup-free quantiser with a 2^K-entry codebook of K-bit sign codes, K = 2..4 (sm_100a): the general form of
// LFQ.forward (models/codebook/LFQ.py:183-307) for dim = 16 -- project_in Linear(16, K), codes {-1,+1}^K, project_out Linear(K, 16).
// (K = 1, the IDEE configuration, has its own scalar-plane kernels in lfq.cu.)  Always fp32 (LFQ.py:183,199).
//
// forward, one thread per token: s = W_in z + b_in (K dot products) ; distances d_j = -2 <s, c_j> to the 2^K codes ; the code is
//   the argmin (ties -> lowest index, i.e. bit i = [s_i > 0]: every s_i == 0 keeps bit 0, LFQ.py:221-222) ; driver index = sum of
//   bit_i << (K-1-i) ; straight-through value x = q (the reference's s + (q - s).detach(), exact here) ; z_q = b_out + sum_i W_out[:, i] x_i (FMA chain).
//   train: p = softmax_j(-inv_temp d_j) ; sums of the per-token entropy, of p (codebook entropy) and of |s - q|^2 in double per
//   CTA, then a one-thread finalize: aux = lambda_c commit + lambda_e H_tok - gamma H(mean p).
// backward, one pass: g_s = W_out^T g_zq (straight-through) + g_aux d(aux)/ds from the saved mean probabilities;
//   g_z = W_in^T g_s ; per-thread running sums of the four parameter gradients, reduced per CTA in double, deterministic finalize.
#include "common.cuh"
#include "idee_b200.h"

namespace {

constexpr int C = 16, NT = 128, NB_MAX = 2048;
constexpr float LOG_EPS = 1e-5f;   // LFQ.py:52

__device__ __forceinline__ float ent_grad(float p) { return -logf(fmaxf(p, LOG_EPS)) - (p >= LOG_EPS ? 1.f : 0.f); }

template <int NACC>
__device__ __forceinline__ void block_reduce_store(const double* acc, double* out) {
    __shared__ double red[NT / 32][NACC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double v = warp_sum_d(acc[a]);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < NACC; a += NT) {
        double v = 0.0;
        for (int w = 0; w < NT / 32; ++w) v += red[w][a];
        out[a] = v;
    }
}

// s, code bits and the softmax over the 2^K codes of one token
template <int K>
__device__ __forceinline__ void token_probs(const float* s, float inv_temp, float* p) {
    constexpr int J = 1 << K;
    float lg[J], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) dot += ((j >> (K - 1 - i)) & 1) ? s[i] : -s[i];      // <s, c_j>, bit i of code j at position K-1-i (LFQ.py:134,139-146)
        lg[j] = 2.f * inv_temp * dot;                                                    // -inv_temp * d_j, d_j = -2 <s, c_j> (LFQ.py:239-240)
        m = fmaxf(m, lg[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) { p[j] = expf(lg[j] - m); sum += p[j]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < J; ++j) p[j] *= inv;
}

// partial layout per CTA (double): entropy sum | commit sum | p sums [2^K]
template <int K>
__global__ void __launch_bounds__(NT)
lfqk_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w_in, const float* __restrict__ b_in, const float* __restrict__ w_out,
                const float* __restrict__ b_out, float* __restrict__ zq, long long* __restrict__ indices, double* __restrict__ partials,
                int64_t ntok, int training, float inv_temp) {
    constexpr int J = 1 << K;
    __shared__ float wi[K * C], bi[K], wo[C * K], bo[C];
    for (int e = threadIdx.x; e < K * C; e += NT) { wi[e] = w_in[e]; wo[e] = w_out[e]; }
    if (threadIdx.x < K) bi[threadIdx.x] = b_in[threadIdx.x];
    if (threadIdx.x < C) bo[threadIdx.x] = b_out[threadIdx.x];
    __syncthreads();
    double acc[2 + J];
#pragma unroll
    for (int a = 0; a < 2 + J; ++a) acc[a] = 0.0;
    for (int64_t tok = (int64_t)blockIdx.x * NT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * NT) {
        float zr[C], s[K], x[K];
        load16(zr, z + tok * C);
        long long idx = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            float a = bi[i];
#pragma unroll
            for (int c = 0; c < C; ++c) a += wi[i * C + c] * zr[c];
            s[i] = a;
            const float q = a > 0.f ? 1.f : -1.f;
            x[i] = q;                    // LFQ.py:226-230: s + (q - s).detach() == q up to one ulp; the STE gradient is analytic (backward)
            idx |= (long long)(x[i] > 0.f ? 1 : 0) << (K - 1 - i);                       // :234
            if (training) acc[1] += (double)((a - q) * (a - q));
        }
        float out[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float a = bo[c];
#pragma unroll
            for (int i = 0; i < K; ++i) a += wo[c * K + i] * x[i];
            out[c] = a;
        }
        store16(zq + tok * C, out);
        indices[tok] = idx;
        if (training) {
            float p[J];
            token_probs<K>(s, inv_temp, p);
#pragma unroll
            for (int j = 0; j < J; ++j) { acc[0] += (double)(-p[j] * logf(fmaxf(p[j], LOG_EPS))); acc[2 + j] += (double)p[j]; }
        }
    }
    block_reduce_store<2 + J>(acc, partials + (int64_t)blockIdx.x * (2 + J));
}

// stats: aux | H_tok | H_cb | commit | mean p [2^K]
template <int K>
__global__ void lfqk_finalize_kernel(const double* __restrict__ partials, int nblocks, int64_t ntok, float lam_c, float lam_e, float gamma,
                                     float* __restrict__ stats) {
    constexpr int J = 1 << K;
    if (threadIdx.x != 0) return;
    double tot[2 + J];
    for (int a = 0; a < 2 + J; ++a) tot[a] = 0.0;
    for (int b = 0; b < nblocks; ++b)
        for (int a = 0; a < 2 + J; ++a) tot[a] += partials[(int64_t)b * (2 + J) + a];
    const float htok = (float)(tot[0] / (double)ntok), commit = (float)(tot[1] / ((double)ntok * K));
    float hcb = 0.f;
    for (int j = 0; j < J; ++j) {
        const float pm = (float)(tot[2 + j] / (double)ntok);
        stats[4 + j] = pm;
        hcb += -pm * logf(fmaxf(pm, LOG_EPS));
    }
    stats[0] = lam_c * commit + lam_e * htok - gamma * hcb;                              // LFQ.py:262,300
    stats[1] = htok; stats[2] = hcb; stats[3] = commit;
}

// partial layout per CTA (double): g_w_in [K][16] | g_b_in [K] | g_w_out [16][K] | g_b_out [16]
template <int K>
__global__ void __launch_bounds__(NT)
lfqk_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gzq, const float* __restrict__ g_aux, const float* __restrict__ stats,
                const float* __restrict__ w_in, const float* __restrict__ b_in, const float* __restrict__ w_out, float* __restrict__ gz,
                double* __restrict__ partials, int64_t ntok, int training, float inv_temp, float lam_c, float lam_e, float gamma) {
    constexpr int J = 1 << K, NACC = K * C + K + C * K + C;
    __shared__ float wi[K * C], bi[K], wo[C * K], gpm[J];
    for (int e = threadIdx.x; e < K * C; e += NT) { wi[e] = w_in[e]; wo[e] = w_out[e]; }
    if (threadIdx.x < K) bi[threadIdx.x] = b_in[threadIdx.x];
    if (threadIdx.x < J) gpm[threadIdx.x] = ent_grad(stats[4 + threadIdx.x]);            // d H(mean p) / d mean p_j
    __syncthreads();
    const float ga = (training && g_aux != nullptr) ? g_aux[0] : 0.f;
    const float inv_n = 1.f / (float)ntok;
    float acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
    for (int64_t tok = (int64_t)blockIdx.x * NT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * NT) {
        float zr[C], gq[C], s[K], x[K], gs[K];
        load16(zr, z + tok * C);
        load16(gq, gzq + tok * C);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            float a = bi[i];
#pragma unroll
            for (int c = 0; c < C; ++c) a += wi[i * C + c] * zr[c];
            s[i] = a;
            const float q = a > 0.f ? 1.f : -1.f;
            x[i] = q;
            float g = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) g += wo[c * K + i] * gq[c];                      // straight-through: d z_q / d x_i
            gs[i] = training ? g + ga * lam_c * 2.f * (a - q) * inv_n * (1.f / K) : 0.f; // eval: x = q has no path to s (LFQ.py:229-230)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[K * C + K + c * K + i] += x[i] * gq[c];      // g_w_out[c][i]
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[K * C + K + C * K + c] += gq[c];                 // g_b_out
        if (training && ga != 0.f) {
            float p[J];
            token_probs<K>(s, inv_temp, p);
            // d aux / d p_j = lam_e / n * dH(p)/dp_j - gamma / n * dH(pm)/dpm_j ;  d p_j / d s_i = 2 inv_temp p_j (c_ji - m_i)
            float wp[J], wsum = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) { wp[j] = (lam_e * ent_grad(p[j]) - gamma * gpm[j]) * inv_n * p[j]; wsum += wp[j]; }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                float wc = 0.f, m = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float cji = ((j >> (K - 1 - i)) & 1) ? 1.f : -1.f;
                    wc += wp[j] * cji;
                    m += p[j] * cji;
                }
                gs[i] += ga * 2.f * inv_temp * (wc - m * wsum);
            }
        }
        float gzr[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < K; ++i) a += wi[i * C + c] * gs[i];
            gzr[c] = a;
        }
        store16(gz + tok * C, gzr);
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int c = 0; c < C; ++c) acc[i * C + c] += gs[i] * zr[c];                 // g_w_in[i][c]
            acc[K * C + i] += gs[i];                                                     // g_b_in[i]
        }
    }
    double dacc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) dacc[a] = (double)acc[a];
    block_reduce_store<NACC>(dacc, partials + (int64_t)blockIdx.x * NACC);
}

__global__ void lfqk_grad_finalize_kernel(const double* __restrict__ partials, int nblocks, int nacc, float* __restrict__ grads) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nacc) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partials[(int64_t)b * nacc + a];
    grads[a] = (float)v;
}

int blocks_for(int64_t ntok) {
    int64_t nb = (ntok + NT - 1) / NT;
    const int64_t cap = (int64_t)idee_num_sms() * 8;
    if (nb > cap) nb = cap;
    if (nb > NB_MAX) nb = NB_MAX;
    return nb < 1 ? 1 : (int)nb;
}

template <int K>
int run_fwd(const float* z, const float* w_in, const float* b_in, const float* w_out, const float* b_out, float* zq, int64_t* indices,
            float* stats, int64_t ntok, int training, float inv_temp, float lam_c, float lam_e, float gamma, double* ws, cudaStream_t st) {
    const int nb = blocks_for(ntok);
    lfqk_fwd_kernel<K><<<nb, NT, 0, st>>>(z, w_in, b_in, w_out, b_out, zq, (long long*)indices, ws, ntok, training, inv_temp);
    IDEE_LAUNCH_CHECK("lfqk_fwd");
    if (training) {
        lfqk_finalize_kernel<K><<<1, 32, 0, st>>>(ws, nb, ntok, lam_c, lam_e, gamma, stats);
        IDEE_LAUNCH_CHECK("lfqk_finalize");
    }
    return 0;
}
template <int K>
int run_bwd(const float* z, const float* gzq, const float* g_aux, const float* stats, const float* w_in, const float* b_in,
            const float* w_out, float* gz, float* grads, int64_t ntok, int training, float inv_temp, float lam_c, float lam_e, float gamma,
            double* ws, cudaStream_t st) {
    constexpr int NACC = K * C + K + C * K + C;
    const int nb = blocks_for(ntok);
    lfqk_bwd_kernel<K><<<nb, NT, 0, st>>>(z, gzq, g_aux, stats, w_in, b_in, w_out, gz, ws, ntok, training, inv_temp, lam_c, lam_e, gamma);
    IDEE_LAUNCH_CHECK("lfqk_bwd");
    lfqk_grad_finalize_kernel<<<(NACC + 127) / 128, 128, 0, st>>>(ws, nb, NACC, grads);
    IDEE_LAUNCH_CHECK("lfqk_grad_finalize");
    return 0;
}

}  // namespace

extern "C" size_t idee_lfqk_workspace_bytes(int codebook_bits) {
    const int nacc = codebook_bits * C + codebook_bits + C * codebook_bits + C, nf = 2 + (1 << codebook_bits);
    return sizeof(double) * (size_t)NB_MAX * (nacc > nf ? nacc : nf);
}

extern "C" int idee_lfqk_fwd(const float* z, const float* w_in, const float* b_in, const float* w_out, const float* b_out, float* zq,
                             int64_t* indices, float* stats, int64_t ntok, int dim, int codebook_bits, int training, float inv_temperature,
                             float lambda_commit, float lambda_entropy, float diversity_gamma, void* workspace, size_t workspace_bytes,
                             void* stream) {
    IDEE_REQUIRE(dim == C, "lfqk_fwd: only dim 16 is built (got %d)", dim);
    IDEE_REQUIRE(codebook_bits >= 2 && codebook_bits <= 4, "lfqk_fwd: codebook sizes 4, 8, 16 are built here (bits = %d); size 2 is idee_lfq_fwd", codebook_bits);
    IDEE_REQUIRE(workspace_bytes >= idee_lfqk_workspace_bytes(codebook_bits), "lfqk_fwd: workspace too small");
    IDEE_REQUIRE(ntok > 0, "lfqk_fwd: no tokens");
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = (double*)workspace;
    switch (codebook_bits) {
        case 2: return run_fwd<2>(z, w_in, b_in, w_out, b_out, zq, indices, stats, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
        case 3: return run_fwd<3>(z, w_in, b_in, w_out, b_out, zq, indices, stats, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
        default: return run_fwd<4>(z, w_in, b_in, w_out, b_out, zq, indices, stats, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
    }
}

/* grads: g_w_in [K][16] | g_b_in [K] | g_w_out [16][K] | g_b_out [16] */
extern "C" int idee_lfqk_bwd(const float* z, const float* gzq, const float* g_aux, const float* stats, const float* w_in, const float* b_in,
                             const float* w_out, float* gz, float* grads, int64_t ntok, int codebook_bits, int training,
                             float inv_temperature, float lambda_commit, float lambda_entropy, float diversity_gamma, void* workspace,
                             size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(codebook_bits >= 2 && codebook_bits <= 4, "lfqk_bwd: codebook bits must be 2..4 (got %d)", codebook_bits);
    IDEE_REQUIRE(workspace_bytes >= idee_lfqk_workspace_bytes(codebook_bits), "lfqk_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = (double*)workspace;
    switch (codebook_bits) {
        case 2: return run_bwd<2>(z, gzq, g_aux, stats, w_in, b_in, w_out, gz, grads, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
        case 3: return run_bwd<3>(z, gzq, g_aux, stats, w_in, b_in, w_out, gz, grads, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
        default: return run_bwd<4>(z, gzq, g_aux, stats, w_in, b_in, w_out, gz, grads, ntok, training, inv_temperature, lambda_commit, lambda_entropy, diversity_gamma, ws, st);
    }
}
