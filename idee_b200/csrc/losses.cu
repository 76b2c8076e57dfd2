// Fused loss kernels of the IDEE training step (sm_100a, fp32 with double block reductions).
//   BCE_loss_synthetic        losses.py:105-124   class-frequency weighted BCE-with-logits, K logit maps per launch
//   Anomaly_L1_loss_synthetic losses.py:147-168   masked L1 between z_q and the code of index 0, without the three
//                                                 full-size broadcast temporaries the reference materialises
#include "common.cuh"
#include "idee_b200.h"

namespace {


__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (warp == 0) {
        s = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
        s = warp_sum_d(s);
        if (lane == 0) red[0] = s;
    }
    __syncthreads();
    s = red[0];
    return s;
}

// ---- BCE_loss_synthetic: 4 small multi-block launches (range, counts, weighted loss + gradient, finalize) ----
constexpr int BNB = 64;      // blocks per pass
constexpr int BT = 256;
// workspace (doubles): [0,BNB) min | [BNB,2BNB) max | [2BNB,3BNB) count of bin 1 | [3BNB, 3BNB + K*BNB) loss partials
__global__ void __launch_bounds__(BT)
bce_range_kernel(const float* __restrict__ target, int64_t M, double* __restrict__ ws) {
    __shared__ float rl[BT / 32], rh[BT / 32];
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < M; i += (int64_t)BNB * BT) { const float t = target[i]; lo = fminf(lo, t); hi = fmaxf(hi, t); }
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { rl[threadIdx.x >> 5] = lo; rh[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BT / 32; ++w) { lo = fminf(lo, rl[w]); hi = fmaxf(hi, rh[w]); }
        ws[blockIdx.x] = lo; ws[BNB + blockIdx.x] = hi;
    }
}
__device__ __forceinline__ void bce_range(const double* ws, float& lo, float& hi) {
    lo = INFINITY; hi = -INFINITY;
    for (int b = 0; b < BNB; ++b) { lo = fminf(lo, (float)ws[b]); hi = fmaxf(hi, (float)ws[BNB + b]); }
    if (lo == hi) { lo -= 1.f; hi += 1.f; }     // torch.histc widens a degenerate range
}
// class counts of torch.histc(target, bins=2) over [min,max] (losses.py:115)
__global__ void __launch_bounds__(BT)
bce_count_kernel(const float* __restrict__ target, int64_t M, double* __restrict__ ws) {
    __shared__ double red[32];
    float lo, hi;
    bce_range(ws, lo, hi);
    double c1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < M; i += (int64_t)BNB * BT) {
        int bin = (int)((target[i] - lo) / (hi - lo) * 2.f);
        c1 += bin > 1 ? 1 : bin;
    }
    c1 = block_sum_d(c1, red);
    if (threadIdx.x == 0) ws[2 * BNB + blockIdx.x] = c1;
}
// w_k = log((cnt_k / total)^-0.5 + 1.1)  (losses.py:117-119)
__device__ __forceinline__ void bce_class_weights(const double* ws, int64_t M, float& w0, float& w1) {
    double c1 = 0.0;
    for (int b = 0; b < BNB; ++b) c1 += ws[2 * BNB + b];
    const float n1 = (float)c1, n0 = (float)((double)M - c1), tot = n0 + n1;
    w0 = logf(powf(n0 / tot, -0.5f) + 1.1f);
    w1 = logf(powf(n1 / tot, -0.5f) + 1.1f);
}
// grid (BNB, K): partial sums of w[target] * bce_with_logits(pred, target); dpred = w * (sigmoid(x) - t) / M
__global__ void __launch_bounds__(BT)
bce_loss_kernel(const float* __restrict__ pred, int64_t sk, int64_t sn, int N, int64_t HW, const float* __restrict__ target,
                double* __restrict__ ws, float* __restrict__ dpred) {
    __shared__ double red[32];
    const int k = blockIdx.y;
    const int64_t M = (int64_t)N * HW;
    float w0, w1;
    bce_class_weights(ws, M, w0, w1);
    const float invM = 1.f / (float)M;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * BT + threadIdx.x; i < M; i += (int64_t)BNB * BT) {
        const int64_t n = i / HW, r = i - n * HW;
        const int64_t o = k * sk + n * sn + r;
        const float x = pred[o], t = target[i];
        const float w = ((int)t >= 1) ? w1 : w0;
        const float l = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        acc += (double)(w * l);
        if (dpred) dpred[o] = w * (1.f / (1.f + expf(-x)) - t) * invM;
    }
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0) ws[3 * BNB + k * BNB + blockIdx.x] = acc;
}
__global__ void bce_finalize_kernel(const double* __restrict__ ws, int K, int64_t M, float* __restrict__ wts, float* __restrict__ loss) {
    const int k = threadIdx.x;
    if (k == 0) { float w0, w1; bce_class_weights(ws, M, w0, w1); wts[0] = w0; wts[1] = w1; }
    if (k < K) {
        double a = 0.0;
        for (int b = 0; b < BNB; ++b) a += ws[3 * BNB + k * BNB + b];
        loss[k] = (float)(a / (double)M);
    }
}

// ---- anomaly L1 ----
constexpr int AT = 256;
struct AnomP { const float* zq; const float* mask; const float* vq0; int N, V, T; int64_t HW; };

// partial sums: {sum |zq - vq0| * (1-m) over m != 1,  sum (1-m) over (n,h,w)}
__global__ void __launch_bounds__(AT)
anomaly_l1_fwd_kernel(AnomP p, double* __restrict__ partials) {
    __shared__ double red[32];
    float v0[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v0[c] = __ldg(p.vq0 + c);
    const int64_t per_n = (int64_t)p.V * p.T * p.HW, ntok = (int64_t)p.N * per_n;
    double a = 0.0, wsum = 0.0;
    for (int64_t tok = (int64_t)blockIdx.x * AT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * AT) {
        const int64_t n = tok / per_n, r = tok - n * per_n, hw = r % p.HW;
        const float m = __ldg(p.mask + n * p.HW + hw);
        if (r < p.HW) wsum += (double)(1.f - m);      // v == 0 && t == 0: count each (n,h,w) once
        if (m == 1.f) continue;
        float z[16];
        load16(z, p.zq + tok * 16);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) s += fabsf(z[c] - v0[c]);
        a += (double)(s * (1.f - m));
    }
    a = block_sum_d(a, red);
    wsum = block_sum_d(wsum, red);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = wsum; }
}
// out: {loss, total weight = sum(1-m) * V*C*T}
__global__ void anomaly_l1_finalize_kernel(const double* __restrict__ partials, int nblocks, int V, int T, float* __restrict__ out) {
    double a = 0.0, w = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) { a += partials[2 * b]; w += partials[2 * b + 1]; }
    a = warp_sum_d(a); w = warp_sum_d(w);
    if (threadIdx.x == 0) { const double tw = w * V * 16 * T; out[0] = (float)(a / tw); out[1] = (float)tw; }
}
// g_zq = g_loss * sign(zq - vq0) * (1-m) / total_weight  (0 where m == 1)
__global__ void __launch_bounds__(AT)
anomaly_l1_bwd_kernel(AnomP p, const float* __restrict__ out, const float* __restrict__ g_loss, float* __restrict__ gzq) {
    float v0[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v0[c] = __ldg(p.vq0 + c);
    const float scale = __ldg(g_loss) / __ldg(out + 1);
    const int64_t per_n = (int64_t)p.V * p.T * p.HW, ntok = (int64_t)p.N * per_n;
    for (int64_t tok = (int64_t)blockIdx.x * AT + threadIdx.x; tok < ntok; tok += (int64_t)gridDim.x * AT) {
        const int64_t n = tok / per_n, r = tok - n * per_n, hw = r % p.HW;
        const float m = __ldg(p.mask + n * p.HW + hw);
        float g[16];
        if (m == 1.f) zero16(g);
        else {
            float z[16];
            load16(z, p.zq + tok * 16);
            const float sc = scale * (1.f - m);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                // sign(zq - vq0), with differences at rounding level taken as the exact zero they stand for: a token quantised to
                // code 0 has zq == vq0 mathematically (both are project_out of the all -1 code, LFQ.py:152-181,284), but the two
                // are evaluated by different kernels (different FMA order) and may differ in the last bit
                const float d = z[c] - v0[c], tiny = 4e-7f * fmaxf(fabsf(z[c]), fabsf(v0[c]));
                g[c] = d > tiny ? sc : (d < -tiny ? -sc : 0.f);
            }
        }
        store16(gzq + tok * 16, g);
    }
}

// ---- anomaly L1 on the rank-1 form of z_q (z_q[c] = x * w_out[c] + b_out[c], x = +-1): the per-token loss takes one of two
// values A(+1), A(-1), so the forward pass is a mask-weighted count of the two signs over the scalar plane x (1/16 of z_q) ----
// partials per block: {sum (1-m) [x > 0], sum (1-m) [x <= 0] over tokens with m != 1, sum (1-m) over (n,h,w)}
__global__ void __launch_bounds__(AT)
anomaly_rank1_fwd_kernel(const float* __restrict__ xq, const float* __restrict__ mask, int N, int V, int T, int64_t HW,
                         double* __restrict__ partials) {
    // grid = (blocks, N): a block strides over the V*T*HW tokens of one sample with 32-bit indices (no 64-bit divisions)
    __shared__ double red[32];
    const int n = blockIdx.y;
    const uint32_t hw = (uint32_t)HW, per_n = (uint32_t)(V * T) * hw;
    const float* xn = xq + (int64_t)n * per_n;
    const float* mn = mask + (int64_t)n * HW;
    double cp = 0.0, cm = 0.0, wsum = 0.0;
    for (uint32_t r = blockIdx.x * AT + threadIdx.x; r < per_n; r += gridDim.x * AT) {
        const float m = __ldg(mn + r % hw);
        if (r < hw) wsum += (double)(1.f - m);
        if (m == 1.f) continue;
        if (__ldg(xn + r) > 0.f) cp += (double)(1.f - m); else cm += (double)(1.f - m);
    }
    cp = block_sum_d(cp, red); cm = block_sum_d(cm, red); wsum = block_sum_d(wsum, red);
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) { partials[3 * b] = cp; partials[3 * b + 1] = cm; partials[3 * b + 2] = wsum; }
}
// out: {loss, total weight = sum(1-m) * V*C*T, count(+1), count(-1)}
__global__ void anomaly_rank1_finalize_kernel(const double* __restrict__ partials, int nblocks, int V, int T, const float* __restrict__ w_out,
                                              const float* __restrict__ b_out, const float* __restrict__ vq0, float* __restrict__ out) {
    double cp = 0.0, cm = 0.0, w = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) { cp += partials[3 * b]; cm += partials[3 * b + 1]; w += partials[3 * b + 2]; }
    cp = warp_sum_d(cp); cm = warp_sum_d(cm); w = warp_sum_d(w);
    if (threadIdx.x == 0) {
        float ap = 0.f, am = 0.f;                       // same fp32 arithmetic as z_q = fma(x, w, b) followed by |z_q - vq0|
        for (int c = 0; c < 16; ++c) { ap += fabsf((w_out[c] + b_out[c]) - vq0[c]); am += fabsf((-w_out[c] + b_out[c]) - vq0[c]); }
        const double tw = w * V * 16 * T;
        out[0] = (float)((cp * (double)ap + cm * (double)am) / tw); out[1] = (float)tw; out[2] = (float)cp; out[3] = (float)cm;
    }
}
// g_x = g_loss * (1-m) / total_weight * sum_c sign(x w_c + b_c - v_c) w_c;  block 0 also writes g_w_out[16], g_b_out[16]
__global__ void __launch_bounds__(AT)
anomaly_rank1_bwd_kernel(const float* __restrict__ xq, const float* __restrict__ mask, int N, int V, int T, int64_t HW,
                         const float* __restrict__ w_out, const float* __restrict__ b_out, const float* __restrict__ vq0,
                         const float* __restrict__ out, const float* __restrict__ g_loss, float* __restrict__ gxq,
                         float* __restrict__ gw, float* __restrict__ gb) {
    const float scale = __ldg(g_loss) / __ldg(out + 1);
    float gp = 0.f, gm = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float w = __ldg(w_out + c), b = __ldg(b_out + c), v = __ldg(vq0 + c);
        const float dp = (w + b) - v, dm = (-w + b) - v;
        gp += dp > 0.f ? w : (dp < 0.f ? -w : 0.f);
        gm += dm > 0.f ? w : (dm < 0.f ? -w : 0.f);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 16) {
        const int c = threadIdx.x;
        const float w = w_out[c], b = b_out[c], v = vq0[c], cp = out[2], cm = out[3];
        const float dp = (w + b) - v, dm = (-w + b) - v;
        const float sp = dp > 0.f ? 1.f : (dp < 0.f ? -1.f : 0.f), sm = dm > 0.f ? 1.f : (dm < 0.f ? -1.f : 0.f);
        gw[c] = scale * (cp * sp - cm * sm);            // d z_q[c] / d w_out[c] = x
        gb[c] = scale * (cp * sp + cm * sm);
    }
    const int n = blockIdx.y;
    const uint32_t hw = (uint32_t)HW, per_n = (uint32_t)(V * T) * hw;
    const float* xn = xq + (int64_t)n * per_n;
    const float* mn = mask + (int64_t)n * HW;
    float* gn = gxq + (int64_t)n * per_n;
    for (uint32_t r = blockIdx.x * AT + threadIdx.x; r < per_n; r += gridDim.x * AT) {
        const float m = __ldg(mn + r % hw);
        gn[r] = m == 1.f ? 0.f : scale * (1.f - m) * (__ldg(xn + r) > 0.f ? gp : gm);
    }
}

int anom_blocks(int64_t ntok) {
    int64_t nb = (ntok + AT - 1) / AT;
    const int cap = idee_num_sms() * 8;
    return (int)(nb > cap ? cap : (nb < 1 ? 1 : nb));
}

// ---- Adam (torch.optim.Adam semantics: L2 weight decay folded into the gradient) ----
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] + wd * p[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// CUDA-graph form: the step counter and the learning rate live in device memory (state[0] = step as float, state[1] = lr), so one
// captured launch sequence serves every replay; the first kernel advances the counter, the second reads it
__global__ void adam_tick_kernel(float* state) { state[0] += 1.f; }
__global__ void adam_state_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                  const float* __restrict__ state, float b1, float b2, float eps, float wd) {
    const float step = state[0], lr = state[1];
    const float bc1 = 1.f - powf(b1, step), bc2_sqrt = sqrtf(1.f - powf(b2, step));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] + wd * p[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

}  // namespace

extern "C" size_t idee_bce_loss_workspace_bytes(int K) { return sizeof(double) * (size_t)(3 + K) * BNB; }

extern "C" int idee_bce_loss_fwd(const float* pred, int64_t stride_k, int64_t stride_n, int K, int N, int64_t HW, const float* target,
                                 float* wts, float* loss, float* dpred, void* workspace, size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(K >= 1 && K <= 1024, "bce_loss_fwd: K must be in [1,1024]");
    IDEE_REQUIRE(workspace_bytes >= idee_bce_loss_workspace_bytes(K), "bce_loss_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = (double*)workspace;
    const int64_t M = (int64_t)N * HW;
    bce_range_kernel<<<BNB, BT, 0, st>>>(target, M, ws);
    IDEE_LAUNCH_CHECK("bce_range");
    bce_count_kernel<<<BNB, BT, 0, st>>>(target, M, ws);
    IDEE_LAUNCH_CHECK("bce_count");
    bce_loss_kernel<<<dim3(BNB, K), BT, 0, st>>>(pred, stride_k, stride_n, N, HW, target, ws, dpred);
    IDEE_LAUNCH_CHECK("bce_loss");
    bce_finalize_kernel<<<1, K < 32 ? 32 : K, 0, st>>>(ws, K, M, wts, loss);
    IDEE_LAUNCH_CHECK("bce_finalize");
    return 0;
}

extern "C" size_t idee_anomaly_l1_workspace_bytes(int64_t ntok) { return sizeof(double) * 2 * (size_t)anom_blocks(ntok); }

extern "C" int idee_anomaly_l1_fwd(const float* zq, const float* mask, const float* vq0, int N, int V, int T, int64_t HW, int C, float* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(C == 16, "anomaly_l1: only C=16 is built (got %d)", C);
    const int64_t ntok = (int64_t)N * V * T * HW;
    IDEE_REQUIRE(workspace_bytes >= idee_anomaly_l1_workspace_bytes(ntok), "anomaly_l1_fwd: workspace too small");
    AnomP p{zq, mask, vq0, N, V, T, HW};
    const int nb = anom_blocks(ntok);
    anomaly_l1_fwd_kernel<<<nb, AT, 0, (cudaStream_t)stream>>>(p, (double*)workspace);
    IDEE_LAUNCH_CHECK("anomaly_l1_fwd");
    anomaly_l1_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, nb, V, T, out);
    IDEE_LAUNCH_CHECK("anomaly_l1_finalize");
    return 0;
}

extern "C" int idee_anomaly_l1_bwd(const float* zq, const float* mask, const float* vq0, int N, int V, int T, int64_t HW, int C,
                                   const float* out, const float* g_loss, float* gzq, void* stream) {
    IDEE_REQUIRE(C == 16, "anomaly_l1: only C=16 is built (got %d)", C);
    const int64_t ntok = (int64_t)N * V * T * HW;
    AnomP p{zq, mask, vq0, N, V, T, HW};
    anomaly_l1_bwd_kernel<<<anom_blocks(ntok), AT, 0, (cudaStream_t)stream>>>(p, out, g_loss, gzq);
    IDEE_LAUNCH_CHECK("anomaly_l1_bwd");
    return 0;
}

// one partial triple per (block, sample); a few thousand spare slots cover tiny inputs whose sample count exceeds the block cap
extern "C" size_t idee_anomaly_rank1_workspace_bytes(int64_t ntok) { return sizeof(double) * 3 * ((size_t)anom_blocks(ntok) + 4096); }

extern "C" int idee_anomaly_rank1_fwd(const float* xq, const float* mask, const float* w_out, const float* b_out, const float* vq0, int N,
                                      int V, int T, int64_t HW, int C, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    IDEE_REQUIRE(C == 16, "anomaly_rank1: only C=16 is built (got %d)", C);
    const int64_t ntok = (int64_t)N * V * T * HW;
    IDEE_REQUIRE(workspace_bytes >= idee_anomaly_rank1_workspace_bytes(ntok), "anomaly_rank1_fwd: workspace too small");
    IDEE_REQUIRE((int64_t)V * T * HW < (1ll << 31) && N <= 65535, "anomaly_rank1: sample too large for 32-bit token indices");
    int bx = anom_blocks(ntok) / N;
    if (bx < 1) bx = 1;
    IDEE_REQUIRE((size_t)bx * N <= (size_t)anom_blocks(ntok) + 4096, "anomaly_rank1_fwd: too many samples for the partial buffer");
    anomaly_rank1_fwd_kernel<<<dim3(bx, N), AT, 0, (cudaStream_t)stream>>>(xq, mask, N, V, T, HW, (double*)workspace);
    IDEE_LAUNCH_CHECK("anomaly_rank1_fwd");
    anomaly_rank1_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, bx * N, V, T, w_out, b_out, vq0, out);
    IDEE_LAUNCH_CHECK("anomaly_rank1_finalize");
    return 0;
}

extern "C" int idee_anomaly_rank1_bwd(const float* xq, const float* mask, const float* w_out, const float* b_out, const float* vq0, int N,
                                      int V, int T, int64_t HW, int C, const float* out, const float* g_loss, float* gxq, float* gw,
                                      float* gb, void* stream) {
    IDEE_REQUIRE(C == 16, "anomaly_rank1: only C=16 is built (got %d)", C);
    const int64_t ntok = (int64_t)N * V * T * HW;
    IDEE_REQUIRE((int64_t)V * T * HW < (1ll << 31) && N <= 65535, "anomaly_rank1: sample too large for 32-bit token indices");
    int bx = anom_blocks(ntok) / N;
    if (bx < 1) bx = 1;
    anomaly_rank1_bwd_kernel<<<dim3(bx, N), AT, 0, (cudaStream_t)stream>>>(xq, mask, N, V, T, HW, w_out, b_out, vq0, out, g_loss, gxq,
                                                                        gw, gb);
    IDEE_LAUNCH_CHECK("anomaly_rank1_bwd");
    return 0;
}

extern "C" int idee_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int step, void* stream) {
    IDEE_REQUIRE(step >= 1, "adam_step: step must be >= 1");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = sqrtf(1.f - powf(beta2, (float)step));
    int nb = (int)((n + 255) / 256);
    const int cap = idee_num_sms() * 8;
    if (nb > cap) nb = cap;
    adam_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2);
    IDEE_LAUNCH_CHECK("adam_step");
    return 0;
}

extern "C" int idee_adam_step_state(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2,
                                    float eps, float weight_decay, void* stream) {
    IDEE_REQUIRE(state != nullptr, "adam_step_state: state (step, lr) must be a device pointer");
    int nb = (int)((n + 255) / 256);
    const int cap = idee_num_sms() * 8;
    if (nb > cap) nb = cap;
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    IDEE_LAUNCH_CHECK("adam_tick");
    adam_state_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, state, beta1, beta2, eps, weight_decay);
    IDEE_LAUNCH_CHECK("adam_step_state");
    return 0;
}
