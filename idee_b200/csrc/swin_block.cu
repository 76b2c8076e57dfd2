// Fused Video-Swin block kernels, fp32 exact path (sm_100a).
//
// Replaces SwinTransformerBlock3D.forward (Swin_3D.py:224-287) + WindowAttention3D.forward (:145-178) +
// Mlp.forward (:36-42) + window_partition/reverse (:45-74) + torch.roll cyclic shift (:241-258) +
// compute_mask (:340-352) for C=16, 2 heads of 8, hidden 64 (config.py:51-66).
//
// Design: one warp owns 32 tokens = 32/G whole windows (G = Wd*Wh*Ww in {8,16,32}); a lane owns one token and
// keeps its 16 channels in registers.  Window partition, cyclic shift, zero padding and crop are pure index
// math on the lane's token coordinate; the shift mask is computed analytically from per-axis region ids;
// the relative-position-bias gather happens once per CTA into shared memory.  K/V rows are exchanged through
// shared memory (broadcast LDS.128), the softmax row lives in registers.
//
// Backward is recompute-based and split in two kernels around the saved mid-block residual y:
//   swin_mlp_bwd  : (y, g_out)  -> g_y,  d{fc1,fc2}
//   swin_attn_bwd : (x, g_y)    -> g_x,  d{qkv,proj,rpb}
// Weight gradients are accumulated in registers across the CTA's persistent loop via a shared-memory
// transposed outer-product phase, written as per-CTA partials and summed by a deterministic second stage.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace {

constexpr int C = 16, NH = 2, HD = 8, HID = 64;
constexpr int RS = 20;  // padded row stride (floats) for 16-float rows in smem: conflict-free LDS/STS.128
__host__ __device__ constexpr int bsz(int G) { return (NH * G * G + 3) / 4 * 4; }  // bias table floats, 16 B aligned

// packed per-variable block parameters, in the reference state_dict order of one block
// (attn.relative_position_bias_table, attn.qkv.{weight,bias}, attn.proj.{weight,bias}, mlp.fc1.*, mlp.fc2.*)
struct POff {
    int qkv_w, qkv_b, proj_w, proj_b, fc1_w, fc1_b, fc2_w, fc2_b, total;
    __host__ __device__ explicit POff(int tbl) {
        qkv_w = tbl * NH; qkv_b = qkv_w + 3 * C * C; proj_w = qkv_b + 3 * C; proj_b = proj_w + C * C;
        fc1_w = proj_b + C; fc1_b = fc1_w + HID * C; fc2_w = fc1_b + HID; fc2_b = fc2_w + C * HID;
        total = fc2_b + C;
    }
};

struct Geom {
    int N, V, T, H, W;
    int Tp, Hp, Wp;
    int nwt, nwh, nww, nwin_img;
    int st, sh, sw, masked;
    int n_wg;     // window groups (warps' worth of windows) per variable
    int tbl;      // rows of the rpb table
    float scale;
    int thwc;     // T*H*W*C (elements of one (n, v) image; < 2^31)
    void* out16;  // forward, bf16 path: optional bf16 copy of the block output (for the proj conv that consumes it)
    const float* emb_x; const float* emb_w; const float* emb_b;   // bf16 path: fused patch embedding (see idee_swin_desc)
    float* emb_gpart;                                             // backward: per-CTA partials [V][CTAs][32] of the embedding gradients
    int x32, out32;                                               // tcgen05 path: the block input / output tokens are fp32 (else bf16)
    FastDiv fd_img, fd_hw, fd_w;                                  // tcgen05 path: window index -> (n, dw, hw, ww) without integer division
};

__device__ __forceinline__ int region_id(int p, int S, int ws, int ss) {
    if (ss == 0) return 0;
    return p < S - ws ? 0 : (p < S - ss ? 1 : 2);
}

template <int WD, int WH, int WW>
struct TokenMap {
    static constexpr int G = WD * WH * WW;
    bool valid;     // token exists in the unpadded tensor
    int64_t off;    // element offset of its 16 channels
    int code;       // shift-mask region code
    __device__ __forceinline__ TokenMap(const Geom& g, int v, int wg, int lane) {
        const int win = wg * (32 / G) + lane / G;
        const int i = lane % G;
        const bool active = win < g.N * g.nwin_img;
        const int n = win / g.nwin_img;
        int r = win - n * g.nwin_img;
        const int dw = r / (g.nwh * g.nww);
        r -= dw * g.nwh * g.nww;
        const int hw = r / g.nww, ww = r - hw * g.nww;
        const int dl = i / (WH * WW), hl = (i / WW) % WH, wl = i % WW;
        const int pt = dw * WD + dl, ph = hw * WH + hl, pw = ww * WW + wl;   // shifted (rolled) frame
        int s_t = pt + g.st; if (s_t >= g.Tp) s_t -= g.Tp;                   // torch.roll(x, -shift)
        int s_h = ph + g.sh; if (s_h >= g.Hp) s_h -= g.Hp;
        int s_w = pw + g.sw; if (s_w >= g.Wp) s_w -= g.Wp;
        valid = active && s_t < g.T && s_h < g.H && s_w < g.W;
        off = ((((int64_t)(n * g.V + v) * g.T + s_t) * g.H + s_h) * g.W + s_w) * C;
        code = g.masked ? (region_id(pt, g.Tp, WD, g.st) * 9 + region_id(ph, g.Hp, WH, g.sh) * 3 +
                           region_id(pw, g.Wp, WW, g.sw)) : 0;
    }
};

// ---- shared-memory weight block common to forward and attention backward ----
struct AttnW {
    float qkv_w[3 * C * C];
    float qkv_b[3 * C];
    float proj_w[C * C];    // [out][in]
    float proj_b[C];
};

template <int G>
__device__ __forceinline__ void stage_bias(float* Bt, float* Bn, const float* tbl, const int* __restrict__ rel_index) {
    // Bt[h][j][i] (row pass: lane i reads over j)  and  Bn[h][i][j] (column pass: lane j reads over i)
    for (int e = threadIdx.x; e < NH * G * G; e += blockDim.x) {
        const int h = e / (G * G), i = (e / G) % G, j = e % G;
        const float b = tbl[rel_index[i * G + j] * NH + h];
        Bt[(h * G + j) * G + i] = b;
        if (Bn) Bn[(h * G + i) * G + j] = b;
    }
}

// q,k,v = Wqkv xn + b  (q scaled), lane-private
__device__ __forceinline__ void qkv_project(const AttnW& w, const float* xn, float scale, float* q, float* k, float* v) {
#pragma unroll
    for (int o = 0; o < 3 * C; ++o) {
        float acc = w.qkv_b[o];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 ww = ld4(&w.qkv_w[o * C + 4 * c4]);
            acc += ww.x * xn[4 * c4] + ww.y * xn[4 * c4 + 1] + ww.z * xn[4 * c4 + 2] + ww.w * xn[4 * c4 + 3];
        }
        if (o < C) q[o] = acc * scale; else if (o < 2 * C) k[o - C] = acc; else v[o - 2 * C] = acc;
    }
}

// one attention head, row pass: returns softmax stats (m, l) and o_h[8] (normalised)
template <int G>
__device__ __forceinline__ void attn_row(const float* q_h, const float* SK, const float* SV, const float* Bt_h,
                                         int base, int i, int h, int code, bool masked, float* o_h, float& m_out, float& l_out) {
    float s[G];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const float* kr = SK + (base + j) * RS + h * HD;
        const float4 k0 = ld4(kr), k1 = ld4(kr + 4);
        float acc = q_h[0] * k0.x + q_h[1] * k0.y + q_h[2] * k0.z + q_h[3] * k0.w +
                    q_h[4] * k1.x + q_h[5] * k1.y + q_h[6] * k1.z + q_h[7] * k1.w;
        acc += Bt_h[j * G + i];
        if (masked) {
            const int cj = __shfl_sync(0xffffffffu, code, base + j);
            if (cj != code) acc += -100.0f;
        }
        s[j] = acc;
        m = fmaxf(m, acc);
    }
    float l = 0.f;
#pragma unroll
    for (int e = 0; e < HD; ++e) o_h[e] = 0.f;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const float p = expf(s[j] - m);
        l += p;
        const float* vr = SV + (base + j) * RS + h * HD;
        const float4 v0 = ld4(vr), v1 = ld4(vr + 4);
        o_h[0] += p * v0.x; o_h[1] += p * v0.y; o_h[2] += p * v0.z; o_h[3] += p * v0.w;
        o_h[4] += p * v1.x; o_h[5] += p * v1.y; o_h[6] += p * v1.z; o_h[7] += p * v1.w;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < HD; ++e) o_h[e] *= inv;
    m_out = m; l_out = l;
}

// =====================================================================================================
// forward
// =====================================================================================================
struct FwdSmem {
    AttnW aw;
    float fc1_w[HID * C];   // [k][c]
    float fc1_b[HID];
    float fc2_wT[HID * C];  // [k][c] = fc2.weight[c][k]
    float fc2_b[C];
};

template <int WD, int WH, int WW, int NWARP>
__global__ void __launch_bounds__(NWARP * 32)
swin_block_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ ymid,
                      const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index, Geom g) {
    constexpr int G = WD * WH * WW;
    extern __shared__ __align__(16) float smem[];
    FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem);
    float* Bt = smem + sizeof(FwdSmem) / 4;             // [NH][G][G]
    float* SKV = Bt + bsz(G);                           // per warp: K[32][RS], V[32][RS]
    const int v = blockIdx.y;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);
    for (int e = threadIdx.x; e < 3 * C * C; e += blockDim.x) S.aw.qkv_w[e] = P[po.qkv_w + e];
    for (int e = threadIdx.x; e < 3 * C; e += blockDim.x) S.aw.qkv_b[e] = P[po.qkv_b + e];
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) S.aw.proj_w[e] = P[po.proj_w + e];
    for (int e = threadIdx.x; e < C; e += blockDim.x) { S.aw.proj_b[e] = P[po.proj_b + e]; S.fc2_b[e] = P[po.fc2_b + e]; }
    for (int e = threadIdx.x; e < HID * C; e += blockDim.x) {
        S.fc1_w[e] = P[po.fc1_w + e];
        const int k = e / C, c = e % C;
        S.fc2_wT[e] = P[po.fc2_w + c * HID + k];
    }
    for (int e = threadIdx.x; e < HID; e += blockDim.x) S.fc1_b[e] = P[po.fc1_b + e];
    stage_bias<G>(Bt, nullptr, P, rel_index);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* SK = SKV + warp * (2 * 32 * RS);
    float* SV = SK + 32 * RS;
    const int base = lane & ~(G - 1), i = lane & (G - 1);

    for (int wg = blockIdx.x * NWARP + warp; wg < g.n_wg; wg += gridDim.x * NWARP) {
        const TokenMap<WD, WH, WW> tm(g, v, wg, lane);
        float xr[C], xn[C];
        if (tm.valid) { load16(xr, x + tm.off); ln16(xr, xn); } else { zero16(xr); zero16(xn); }
        float q[C], o[C];
        {
            float k[C], vv[C];
            qkv_project(S.aw, xn, g.scale, q, k, vv);
            __syncwarp();
            store16(SK + lane * RS, k);
            store16(SV + lane * RS, vv);
            __syncwarp();
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float m, l;
            attn_row<G>(q + h * HD, SK, SV, Bt + h * G * G, base, i, h, tm.code, g.masked != 0, o + h * HD, m, l);
        }
        // y = x + proj(o)
        float y[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = S.aw.proj_b[c];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                const float4 ww = ld4(&S.aw.proj_w[c * C + 4 * e4]);
                acc += ww.x * o[4 * e4] + ww.y * o[4 * e4 + 1] + ww.z * o[4 * e4 + 2] + ww.w * o[4 * e4 + 3];
            }
            y[c] = xr[c] + acc;
        }
        if (tm.valid && ymid) store16(ymid + tm.off, y);
        // out = y + fc2(gelu(fc1(LN(y))))
        float yn[C], acc[C];
        ln16(y, yn);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = y[c] + S.fc2_b[c];
#pragma unroll 8
        for (int k = 0; k < HID; ++k) {
            float pre = S.fc1_b[k];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 ww = ld4(&S.fc1_w[k * C + 4 * c4]);
                pre += ww.x * yn[4 * c4] + ww.y * yn[4 * c4 + 1] + ww.z * yn[4 * c4 + 2] + ww.w * yn[4 * c4 + 3];
            }
            const float hk = gelu_erf(pre);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 ww = ld4(&S.fc2_wT[k * C + 4 * c4]);
                acc[4 * c4] += ww.x * hk; acc[4 * c4 + 1] += ww.y * hk; acc[4 * c4 + 2] += ww.z * hk; acc[4 * c4 + 3] += ww.w * hk;
            }
        }
        if (tm.valid) store16(out + tm.off, acc);
    }
}

// =====================================================================================================
// backward, MLP half:  out = y + fc2(gelu(fc1(LN(y))))
//   inputs y, g_out;  outputs g_y (may alias g_out), per-CTA partial d{fc1_w, fc1_b, fc2_w, fc2_b}
// =====================================================================================================
constexpr int MLP_TOK = 128;       // tokens per CTA batch (= threads)
constexpr int HS = HID + 1;        // odd stride -> conflict-free scalar access both ways
struct MlpSmem {
    float fc1_w[HID * C];
    float fc1_b[HID];
    float fc2_wT[HID * C];
    float s_h[MLP_TOK * HS];       // gelu(pre)
    float s_gp[MLP_TOK * HS];      // g_pre
    float s_yn[MLP_TOK * RS];
    float s_go[MLP_TOK * RS];
};
// partial layout per CTA: fc1_w[64*16] | fc1_b[64] | fc2_w[16*64] | fc2_b[16]
constexpr int MLP_PART = HID * C + HID + C * HID + C;

__global__ void __launch_bounds__(MLP_TOK)
swin_mlp_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gout, float* __restrict__ gy,
                    const float* __restrict__ params, int64_t pstride, int tbl, float* __restrict__ partials,
                    int N, int V, int64_t thw) {
    extern __shared__ __align__(16) float smem[];
    MlpSmem& S = *reinterpret_cast<MlpSmem*>(smem);
    const int v = blockIdx.y, tid = threadIdx.x;
    const float* P = params + (int64_t)v * pstride;
    const POff po(tbl);
    for (int e = tid; e < HID * C; e += MLP_TOK) {
        S.fc1_w[e] = P[po.fc1_w + e];
        const int k = e / C, c = e % C;
        S.fc2_wT[e] = P[po.fc2_w + c * HID + k];
    }
    for (int e = tid; e < HID; e += MLP_TOK) S.fc1_b[e] = P[po.fc1_b + e];
    __syncthreads();

    // phase-2 ownership: k = tid%64, channel group cg = tid/64 -> channels cg*8..cg*8+7
    const int pk = tid & 63, pcg = tid >> 6;
    float a_w1[8], a_w2[8], a_b1 = 0.f, a_b2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { a_w1[e] = 0.f; a_w2[e] = 0.f; }

    const int64_t ntok = (int64_t)N * thw;
    for (int64_t t0 = (int64_t)blockIdx.x * MLP_TOK; t0 < ntok; t0 += (int64_t)gridDim.x * MLP_TOK) {
        const int64_t tok = t0 + tid;
        const bool valid = tok < ntok;
        int64_t off = 0;
        if (valid) { const int64_t n = tok / thw; off = ((n * V + v) * thw + (tok - n * thw)) * C; }
        float yr[C], yn[C], go[C], gyn[C];
        float rstd = 0.f;
        if (valid) { load16(yr, y + off); load16(go, gout + off); rstd = ln16(yr, yn); }
        else { zero16(yn); zero16(go); }
        zero16(gyn);
#pragma unroll 4
        for (int k = 0; k < HID; ++k) {
            float w1[C];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 ww = ld4(&S.fc1_w[k * C + 4 * c4]);
                w1[4 * c4] = ww.x; w1[4 * c4 + 1] = ww.y; w1[4 * c4 + 2] = ww.z; w1[4 * c4 + 3] = ww.w;
            }
            float pre = S.fc1_b[k], gh = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) pre += w1[c] * yn[c];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 ww = ld4(&S.fc2_wT[k * C + 4 * c4]);
                gh += ww.x * go[4 * c4] + ww.y * go[4 * c4 + 1] + ww.z * go[4 * c4 + 2] + ww.w * go[4 * c4 + 3];
            }
            const float gp = valid ? gh * gelu_erf_grad(pre) : 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) gyn[c] += w1[c] * gp;
            S.s_h[tid * HS + k] = valid ? gelu_erf(pre) : 0.f;
            S.s_gp[tid * HS + k] = gp;
        }
        store16(&S.s_yn[tid * RS], yn);
        store16(&S.s_go[tid * RS], go);
        if (valid) {
            float gx[C];
            ln16_bwd(gyn, yn, rstd, gx);
#pragma unroll
            for (int c = 0; c < C; ++c) gx[c] += go[c];
            store16(gy + off, gx);
        }
        __syncthreads();
        // phase 2: transposed outer-product accumulation over the 128 staged tokens
#pragma unroll 4
        for (int t = 0; t < MLP_TOK; ++t) {
            const float hk = S.s_h[t * HS + pk], gpk = S.s_gp[t * HS + pk];
            const float4 g0 = ld4(&S.s_go[t * RS + pcg * 8]), g1 = ld4(&S.s_go[t * RS + pcg * 8 + 4]);
            const float4 n0 = ld4(&S.s_yn[t * RS + pcg * 8]), n1 = ld4(&S.s_yn[t * RS + pcg * 8 + 4]);
            a_w2[0] += g0.x * hk; a_w2[1] += g0.y * hk; a_w2[2] += g0.z * hk; a_w2[3] += g0.w * hk;
            a_w2[4] += g1.x * hk; a_w2[5] += g1.y * hk; a_w2[6] += g1.z * hk; a_w2[7] += g1.w * hk;
            a_w1[0] += n0.x * gpk; a_w1[1] += n0.y * gpk; a_w1[2] += n0.z * gpk; a_w1[3] += n0.w * gpk;
            a_w1[4] += n1.x * gpk; a_w1[5] += n1.y * gpk; a_w1[6] += n1.z * gpk; a_w1[7] += n1.w * gpk;
            if (pcg == 0) a_b1 += gpk;
            if (tid < C) a_b2 += S.s_go[t * RS + tid];
        }
        __syncthreads();
    }
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * MLP_PART;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        part[pk * C + pcg * 8 + e] = a_w1[e];                               // fc1.weight[k][c]
        part[HID * C + HID + (pcg * 8 + e) * HID + pk] = a_w2[e];           // fc2.weight[c][k]
    }
    if (pcg == 0) part[HID * C + pk] = a_b1;
    if (tid < C) part[HID * C + HID + C * HID + tid] = a_b2;
}

// =====================================================================================================
// backward, attention half:  y = x + proj(attn(LN(x)))
//   inputs x, g_y;  outputs g_x (may alias g_y), per-CTA partial d{qkv_w,qkv_b,proj_w,proj_b, bias[h][i][j]}
// =====================================================================================================
constexpr int AB_WARPS = 4;
// per-warp exchange/stage region, in floats.  Exchange: Q,K,V,GO rows [32][RS] + stats [32][8].
// Stage (aliases exchange after the window is done): xn[32][RS], o[32][RS], ga[32][RS], gqkv[32][52]
constexpr int GQS = 52;
constexpr int AB_EXCH = 4 * 32 * RS + 32 * 8;
constexpr int AB_STAGE = 3 * 32 * RS + 32 * GQS;
constexpr int AB_WREG = (AB_EXCH > AB_STAGE ? AB_EXCH : AB_STAGE);
// partial layout per CTA: qkv_w[48*16] | qkv_b[48] | proj_w[16*16] | proj_b[16] | dB[NH*G*G]
constexpr int ATT_PART_W = 3 * C * C + 3 * C + C * C + C;

template <int WD, int WH, int WW>
__global__ void __launch_bounds__(AB_WARPS * 32)
swin_attn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                     const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index,
                     float* __restrict__ partials, Geom g) {
    constexpr int G = WD * WH * WW;
    extern __shared__ __align__(16) float smem[];
    AttnW& AW = *reinterpret_cast<AttnW*>(smem);
    float* Bt = smem + sizeof(AttnW) / 4;      // [NH][G][G] : [h][j][i]
    float* Bn = Bt + bsz(G);                   // [NH][G][G] : [h][i][j]
    float* dBc = Bn + bsz(G);                  // [NH][G][G] CTA reduction target
    float* WREG = dBc + bsz(G);                // AB_WARPS * AB_WREG
    const int v = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);
    for (int e = tid; e < 3 * C * C; e += blockDim.x) AW.qkv_w[e] = P[po.qkv_w + e];
    for (int e = tid; e < 3 * C; e += blockDim.x) AW.qkv_b[e] = P[po.qkv_b + e];
    for (int e = tid; e < C * C; e += blockDim.x) AW.proj_w[e] = P[po.proj_w + e];
    for (int e = tid; e < C; e += blockDim.x) AW.proj_b[e] = P[po.proj_b + e];
    for (int e = tid; e < NH * G * G; e += blockDim.x) dBc[e] = 0.f;
    stage_bias<G>(Bt, Bn, P, rel_index);
    __syncthreads();

    float* R = WREG + warp * AB_WREG;
    float* SQ = R, *SK = R + 32 * RS, *SV = R + 2 * 32 * RS, *SGO = R + 3 * 32 * RS, *SST = R + 4 * 32 * RS;
    float* T_xn = R, *T_o = R + 32 * RS, *T_ga = R + 2 * 32 * RS, *T_gq = R + 3 * 32 * RS;
    const int base = lane & ~(G - 1), i = lane & (G - 1);
    const bool masked = g.masked != 0;

    // persistent accumulators
    float dB[NH][G];
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int j = 0; j < G; ++j) dB[h][j] = 0.f;
    // phase-2 ownership (warp-uniform roles):
    //   threads 0..95  : qkv_w[o][ch*8 .. ch*8+7], o = tid % 48, ch = tid / 48 ; threads 0..47 also own qkv_b[o]
    //   threads 96..127: proj_w[c][eh*8 .. eh*8+7], c = (tid-96) / 2, eh = (tid-96) % 2 ; eh==0 also owns proj_b[c]
    const bool own_qkv = tid < 96;
    const int p_row = own_qkv ? tid % 48 : (tid - 96) >> 1;
    const int p_half = own_qkv ? tid / 48 : (tid - 96) & 1;
    float a_w[8], a_b = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) a_w[r] = 0.f;

    const int n_iter = (g.n_wg + gridDim.x * AB_WARPS - 1) / (gridDim.x * AB_WARPS);
    for (int it = 0; it < n_iter; ++it) {
        const int wg = (it * gridDim.x + blockIdx.x) * AB_WARPS + warp;
        // windows beyond the end are processed as all-invalid tokens (zero contribution)
        const TokenMap<WD, WH, WW> tm(g, v, wg < g.n_wg ? wg : 0, lane);
        const bool valid = tm.valid && wg < g.n_wg;
        float xn[C];
        float rstd = 0.f;
        {
            float xr[C];
            if (valid) { load16(xr, x + tm.off); rstd = ln16(xr, xn); } else { zero16(xn); }
        }
        float q[C], k[C], vv[C];
        qkv_project(AW, xn, g.scale, q, k, vv);
        store16(SQ + lane * RS, q);
        store16(SK + lane * RS, k);
        store16(SV + lane * RS, vv);
        __syncwarp();
        float o[C], mh[NH], lh[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h)
            attn_row<G>(q + h * HD, SK, SV, Bt + h * G * G, base, i, h, tm.code, masked, o + h * HD, mh[h], lh[h]);
        // g_a = g_y ; g_o = Wproj^T g_a ; D_h = <g_o_h, o_h>
        float ga[C], go[C], Dh[NH];
        if (valid) load16(ga, gy + tm.off); else zero16(ga);
        zero16(go);
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                const float4 ww = ld4(&AW.proj_w[c * C + 4 * e4]);
                go[4 * e4] += ww.x * ga[c]; go[4 * e4 + 1] += ww.y * ga[c]; go[4 * e4 + 2] += ww.z * ga[c]; go[4 * e4 + 3] += ww.w * ga[c];
            }
        }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < HD; ++e) d += go[h * HD + e] * o[h * HD + e];
            Dh[h] = d;
        }
        store16(SGO + lane * RS, go);
        st4(SST + lane * 8, make_float4(mh[0], 1.f / lh[0], Dh[0], 0.f));
        st4(SST + lane * 8 + 4, make_float4(mh[1], 1.f / lh[1], Dh[1], 0.f));
        __syncwarp();

        float gq[C], gk[C], gv[C];
        zero16(gq); zero16(gk); zero16(gv);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float* qh = q + h * HD;
            const float* goh = go + h * HD;
            const float invl = 1.f / lh[h];
            // row pass (lane = query i): g_q_i, dB[h][i][j]
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const float* kr = SK + (base + j) * RS + h * HD;
                const float* vr = SV + (base + j) * RS + h * HD;
                const float4 k0 = ld4(kr), k1 = ld4(kr + 4), v0 = ld4(vr), v1 = ld4(vr + 4);
                float s = qh[0] * k0.x + qh[1] * k0.y + qh[2] * k0.z + qh[3] * k0.w +
                          qh[4] * k1.x + qh[5] * k1.y + qh[6] * k1.z + qh[7] * k1.w;
                s += Bt[(h * G + j) * G + i];
                if (masked) { const int cj = __shfl_sync(0xffffffffu, tm.code, base + j); if (cj != tm.code) s += -100.0f; }
                const float p = expf(s - mh[h]) * invl;
                const float gp = goh[0] * v0.x + goh[1] * v0.y + goh[2] * v0.z + goh[3] * v0.w +
                                 goh[4] * v1.x + goh[5] * v1.y + goh[6] * v1.z + goh[7] * v1.w;
                const float gs = p * (gp - Dh[h]);
                dB[h][j] += gs;
                gq[h * HD + 0] += gs * k0.x; gq[h * HD + 1] += gs * k0.y; gq[h * HD + 2] += gs * k0.z; gq[h * HD + 3] += gs * k0.w;
                gq[h * HD + 4] += gs * k1.x; gq[h * HD + 5] += gs * k1.y; gq[h * HD + 6] += gs * k1.z; gq[h * HD + 7] += gs * k1.w;
            }
            // column pass (lane = key j): g_k_j, g_v_j  -- recompute p_ij from the query rows' saved stats
            const float* kh = k + h * HD;
            const float* vh = vv + h * HD;
#pragma unroll
            for (int ii = 0; ii < G; ++ii) {
                const float* qr = SQ + (base + ii) * RS + h * HD;
                const float* gr = SGO + (base + ii) * RS + h * HD;
                const float4 q0 = ld4(qr), q1 = ld4(qr + 4), g0 = ld4(gr), g1 = ld4(gr + 4);
                const float4 stt = ld4(SST + (base + ii) * 8 + h * 4);   // m, 1/l, D
                float s = q0.x * kh[0] + q0.y * kh[1] + q0.z * kh[2] + q0.w * kh[3] +
                          q1.x * kh[4] + q1.y * kh[5] + q1.z * kh[6] + q1.w * kh[7];
                s += Bn[(h * G + ii) * G + i];
                if (masked) { const int ci = __shfl_sync(0xffffffffu, tm.code, base + ii); if (ci != tm.code) s += -100.0f; }
                const float p = expf(s - stt.x) * stt.y;
                const float gp = g0.x * vh[0] + g0.y * vh[1] + g0.z * vh[2] + g0.w * vh[3] +
                                 g1.x * vh[4] + g1.y * vh[5] + g1.z * vh[6] + g1.w * vh[7];
                const float gs = p * (gp - stt.z);
                gv[h * HD + 0] += p * g0.x; gv[h * HD + 1] += p * g0.y; gv[h * HD + 2] += p * g0.z; gv[h * HD + 3] += p * g0.w;
                gv[h * HD + 4] += p * g1.x; gv[h * HD + 5] += p * g1.y; gv[h * HD + 6] += p * g1.z; gv[h * HD + 7] += p * g1.w;
                gk[h * HD + 0] += gs * q0.x; gk[h * HD + 1] += gs * q0.y; gk[h * HD + 2] += gs * q0.z; gk[h * HD + 3] += gs * q0.w;
                gk[h * HD + 4] += gs * q1.x; gk[h * HD + 5] += gs * q1.y; gk[h * HD + 6] += gs * q1.z; gk[h * HD + 7] += gs * q1.w;
            }
        }
        // g wrt the unscaled q projection output
#pragma unroll
        for (int c = 0; c < C; ++c) gq[c] *= g.scale;
        // g_xn = Wqkv^T g_qkv ; g_x = g_y + LN_bwd(g_xn)
        float gxn[C];
        zero16(gxn);
#pragma unroll
        for (int oo = 0; oo < 3 * C; ++oo) {
            const float gqo = oo < C ? gq[oo] : (oo < 2 * C ? gk[oo - C] : gv[oo - 2 * C]);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 ww = ld4(&AW.qkv_w[oo * C + 4 * c4]);
                gxn[4 * c4] += ww.x * gqo; gxn[4 * c4 + 1] += ww.y * gqo; gxn[4 * c4 + 2] += ww.z * gqo; gxn[4 * c4 + 3] += ww.w * gqo;
            }
        }
        if (valid) {
            float gxr[C];
            ln16_bwd(gxn, xn, rstd, gxr);
#pragma unroll
            for (int c = 0; c < C; ++c) gxr[c] += ga[c];
            store16(gx + tm.off, gxr);
        }
        // stage per-token vectors for the CTA-wide weight-gradient phase (aliases the exchange region)
        __syncwarp();
        store16(T_xn + lane * RS, xn);
        store16(T_o + lane * RS, o);
        store16(T_ga + lane * RS, ga);
        store16(T_gq + lane * GQS, gq);
        store16(T_gq + lane * GQS + C, gk);
        store16(T_gq + lane * GQS + 2 * C, gv);
        __syncthreads();
        // phase 2: d qkv_w[o][c] += sum_t gqkv[t][o] * xn[t][c] ; d proj_w[c][e] += sum_t ga[t][c] * o[t][e]
#pragma unroll 4
        for (int t = 0; t < AB_WARPS * 32; ++t) {
            const float* Rt = WREG + (t >> 5) * AB_WREG;
            const int tl = t & 31;
            const float gsc = own_qkv ? Rt[3 * 32 * RS + tl * GQS + p_row] : Rt[2 * 32 * RS + tl * RS + p_row];
            const float* vec = (own_qkv ? Rt : Rt + 32 * RS) + tl * RS + p_half * 8;
            const float4 u0 = ld4(vec), u1 = ld4(vec + 4);
            a_w[0] += gsc * u0.x; a_w[1] += gsc * u0.y; a_w[2] += gsc * u0.z; a_w[3] += gsc * u0.w;
            a_w[4] += gsc * u1.x; a_w[5] += gsc * u1.y; a_w[6] += gsc * u1.z; a_w[7] += gsc * u1.w;
            a_b += gsc;
        }
        __syncthreads();
    }
    // reduce dB over the lanes that share (i) and over warps
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int j = 0; j < G; ++j) atomicAdd(&dBc[(h * G + i) * G + j], dB[h][j]);
    __syncthreads();
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * (ATT_PART_W + NH * G * G);
    if (own_qkv) {
#pragma unroll
        for (int r = 0; r < 8; ++r) part[p_row * C + p_half * 8 + r] = a_w[r];
        if (p_half == 0) part[3 * C * C + p_row] = a_b;
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) part[3 * C * C + 3 * C + p_row * C + p_half * 8 + r] = a_w[r];
        if (p_half == 0) part[3 * C * C + 3 * C + C * C + p_row] = a_b;
    }
    for (int e = tid; e < NH * G * G; e += blockDim.x) part[ATT_PART_W + e] = dBc[e];
}

// ---- second stage: deterministic sum of per-CTA partials into the packed gradient block -------------
// (1) swin_grad_sum_kernel: one thread per partial element sums it over the CTAs in a fixed order (coalesced across threads) and
//     leaves the sum in CTA 0's slot; (2) swin_grad_finalize_kernel, one CTA per variable: writes the packed gradient and gathers
//     the bias-table rows from the summed [h][i][j] gradient through rel_index.
template <int G>
__global__ void __launch_bounds__(256)
swin_grad_sum_kernel(float* __restrict__ part_attn, float* __restrict__ part_mlp, int ncta_attn, int ncta_mlp) {
    constexpr int APS = ATT_PART_W + NH * G * G;
    const int v = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x;
    // four interleaved accumulators: four loads in flight per step instead of a dependent add behind every load latency
    auto sum4 = [](const float* p, int n, int64_t stride) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int c = 0;
        for (; c + 4 <= n; c += 4) {
            a0 += p[(int64_t)c * stride]; a1 += p[(int64_t)(c + 1) * stride]; a2 += p[(int64_t)(c + 2) * stride]; a3 += p[(int64_t)(c + 3) * stride];
        }
        for (; c < n; ++c) a0 += p[(int64_t)c * stride];
        return (a0 + a1) + (a2 + a3);
    };
    if (e < APS) {
        float* p = part_attn + (int64_t)v * ncta_attn * APS + e;
        p[0] = sum4(p, ncta_attn, APS);
    } else if (e < APS + MLP_PART) {
        float* p = part_mlp + (int64_t)v * ncta_mlp * MLP_PART + (e - APS);
        p[0] = sum4(p, ncta_mlp, MLP_PART);
    }
}

template <int G>
__global__ void __launch_bounds__(1024)
swin_grad_finalize_kernel(const float* __restrict__ part_attn, const float* __restrict__ part_mlp,
                          int ncta_attn, int ncta_mlp, const int* __restrict__ rel_index,
                          float* __restrict__ gparams, int64_t pstride, int tbl) {
    constexpr int APS = ATT_PART_W + NH * G * G;
    __shared__ float sum_attn[APS];
    __shared__ float sum_mlp[MLP_PART];
    const int v = blockIdx.x;
    const POff po(tbl);
    for (int e = threadIdx.x; e < APS; e += blockDim.x) sum_attn[e] = part_attn[(int64_t)v * ncta_attn * APS + e];
    for (int e = threadIdx.x; e < MLP_PART; e += blockDim.x) sum_mlp[e] = part_mlp[(int64_t)v * ncta_mlp * MLP_PART + e];
    __syncthreads();
    float* gp = gparams + (int64_t)v * pstride;
    // bias-table entry (row, head) = sum over every (i, j) that indexes it: one warp per entry scans the G*G pairs, fixed-order
    // shuffle tree (the table is the reference's relative_position_index buffer, read as data)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int e = warp; e < po.qkv_w; e += nwarps) {
        const int row = e / NH, h = e % NH;
        float val = 0.f;
        for (int ij = lane; ij < G * G; ij += 32)
            if (rel_index[ij] == row) val += sum_attn[ATT_PART_W + h * G * G + ij];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) gp[e] = val;
    }
    for (int e = po.qkv_w + threadIdx.x; e < po.total; e += blockDim.x)
        gp[e] = e < po.fc1_w ? sum_attn[e - po.qkv_w] : sum_mlp[e - po.fc1_w];
}

template <int G>
int launch_grad_finalize(float* part_attn, float* part_mlp, int ncta_attn, int ncta_mlp, const int* rel_index, float* gparams,
                         int64_t pstride, int tbl, int V, cudaStream_t st) {
    constexpr int APS = ATT_PART_W + NH * G * G;
    swin_grad_sum_kernel<G><<<dim3((APS + MLP_PART + 255) / 256, V), 256, 0, st>>>(part_attn, part_mlp, ncta_attn, ncta_mlp);
    IDEE_LAUNCH_CHECK("swin_grad_sum");
    swin_grad_finalize_kernel<G><<<V, 1024, 0, st>>>(part_attn, part_mlp, ncta_attn, ncta_mlp, rel_index, gparams, pstride, tbl);
    IDEE_LAUNCH_CHECK("swin_grad_finalize");
    return 0;
}

#include "swin_tc.cuh"
#include "swin_umma.cuh"

// ---- host side -----------------------------------------------------------------------------------------
int make_geom(Geom& g, const idee_swin_desc* d, const char* who) {
    IDEE_REQUIRE(d->C == C && d->heads == NH && d->hidden == HID,
                 "%s: only C=16, heads=2, hidden=64 are built (got C=%d heads=%d hidden=%d)", who, d->C, d->heads, d->hidden);
    g.N = d->N; g.V = d->V; g.T = d->T; g.H = d->H; g.W = d->W;
    const int wd = d->wd, wh = d->wh, ww = d->ww;
    g.Tp = (d->T + wd - 1) / wd * wd; g.Hp = (d->H + wh - 1) / wh * wh; g.Wp = (d->W + ww - 1) / ww * ww;
    g.nwt = g.Tp / wd; g.nwh = g.Hp / wh; g.nww = g.Wp / ww;
    g.nwin_img = g.nwt * g.nwh * g.nww;
    g.st = d->st; g.sh = d->sh; g.sw = d->sw;
    g.masked = (d->st | d->sh | d->sw) != 0;
    IDEE_REQUIRE(d->st >= 0 && d->st < wd && d->sh >= 0 && d->sh < wh && d->sw >= 0 && d->sw < ww,
                 "%s: shift must be in [0, window)", who);
    const int Gt = wd * wh * ww;
    const int64_t nwin = (int64_t)d->N * g.nwin_img;
    g.n_wg = (int)((nwin + (32 / Gt) - 1) / (32 / Gt));
    g.tbl = d->rpb_rows;
    g.scale = d->scale;
    IDEE_REQUIRE((int64_t)d->T * d->H * d->W * C < (1ll << 31), "%s: one (n, v) volume must hold fewer than 2^31 elements", who);
    g.thwc = d->T * d->H * d->W * C;
    g.out16 = nullptr;
    IDEE_REQUIRE(d->embed_x == nullptr || (d->precision == 1 && d->embed_w && d->embed_b),
                 "%s: the fused patch embedding needs precision 1 and embed_w / embed_b", who);
    g.emb_x = d->embed_x; g.emb_w = d->embed_w; g.emb_b = d->embed_b; g.emb_gpart = nullptr;
    g.x32 = d->x_dtype == 0; g.out32 = d->out_dtype == 0;
    IDEE_REQUIRE(nwin < (1ll << 31), "%s: too many windows", who);
    g.fd_img = make_fastdiv(g.nwin_img); g.fd_hw = make_fastdiv(g.nwh * g.nww); g.fd_w = make_fastdiv(g.nww);
    IDEE_REQUIRE((d->embed_gw == nullptr) == (d->embed_gb == nullptr) && (d->embed_gw == nullptr || d->embed_x != nullptr),
                 "%s: embed_gw / embed_gb come together and need the fused patch embedding", who);
    return 0;
}

template <int G> constexpr size_t fwd_smem_bytes(int nwarp) {
    return sizeof(FwdSmem) + sizeof(float) * (bsz(G) + (size_t)nwarp * 2 * 32 * RS);
}
template <int G> constexpr size_t attn_bwd_smem_bytes() {
    return sizeof(AttnW) + sizeof(float) * (3 * bsz(G) + (size_t)AB_WARPS * AB_WREG);
}

constexpr int FWD_WARPS = 4;

template <int WD, int WH, int WW>
int launch_fwd(const idee_swin_desc* d, const Geom& g, const float* x, float* out, float* ymid, const float* params,
               const int* rel_index, cudaStream_t st) {
    constexpr int G = WD * WH * WW;
    if (d->precision == 1 && d->act_dtype == 1) {
        if (G < 8) { idee_set_error("swin_block_fwd(bf16 tokens): windows with fewer than 8 tokens are only built for the fp32 path"); return 1; }
        const bool emb = g.emb_x != nullptr;
        const size_t smem = sizeof(swu::FwdSm) + sizeof(float) * NH * G * swu::bns(G);
        const int n_tiles = (g.n_wg + 3) / 4;
        int per_v = idee_num_sms() * 4 / d->V;
        if (per_v > n_tiles) per_v = n_tiles;
        if (per_v < 1) per_v = 1;
        const void* bx = x;
        void* bo = out;
        auto by = reinterpret_cast<__nv_bfloat16*>(ymid);
        if (emb) {
            IDEE_CUDA(cudaFuncSetAttribute(swu::swin_fwd_umma_kernel<WD, WH, WW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_block_fwd(umma)");
            swu::swin_fwd_umma_kernel<WD, WH, WW, true><<<dim3(per_v, d->V), swu::NT, smem, st>>>(bx, bo, by, params, d->param_stride, rel_index, g);
        } else {
            IDEE_CUDA(cudaFuncSetAttribute(swu::swin_fwd_umma_kernel<WD, WH, WW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_block_fwd(umma)");
            swu::swin_fwd_umma_kernel<WD, WH, WW, false><<<dim3(per_v, d->V), swu::NT, smem, st>>>(bx, bo, by, params, d->param_stride, rel_index, g);
        }
        IDEE_LAUNCH_CHECK("swin_block_fwd(umma)");
        return 0;
    }
    if (d->precision == 1) {
        if (G < 8) { idee_set_error("swin_block_fwd(bf16): windows with fewer than 8 tokens are only built for the fp32 path"); return 1; }
        // persistent grid of exactly one resident wave (a partial second wave would run at a fraction of the occupancy)
        const bool emb = g.emb_x != nullptr;
        int per_sm = 1;
        if (emb) IDEE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, swin_fwd_tc_kernel<WD, WH, WW, true>, TCW * 32, 0), "swin_block_fwd(bf16)");
        else IDEE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, swin_fwd_tc_kernel<WD, WH, WW, false>, TCW * 32, 0), "swin_block_fwd(bf16)");
        if (per_sm < 1) per_sm = 1;
        int per_v = (g.n_wg + TCW - 1) / TCW;
        const int cap = idee_num_sms() * per_sm / d->V;
        if (per_v > cap) per_v = cap;
        if (per_v < 1) per_v = 1;
        if (emb) swin_fwd_tc_kernel<WD, WH, WW, true><<<dim3(per_v, d->V), TCW * 32, 0, st>>>(x, out, ymid, params, d->param_stride, rel_index, g);
        else swin_fwd_tc_kernel<WD, WH, WW, false><<<dim3(per_v, d->V), TCW * 32, 0, st>>>(x, out, ymid, params, d->param_stride, rel_index, g);
        IDEE_LAUNCH_CHECK("swin_block_fwd(bf16)");
        return 0;
    }
    auto kern = swin_block_fwd_kernel<WD, WH, WW, FWD_WARPS>;
    const size_t smem = fwd_smem_bytes<G>(FWD_WARPS);
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_block_fwd");
    int per_v = (g.n_wg + FWD_WARPS - 1) / FWD_WARPS;
    const int cap = (idee_num_sms() * 4 + d->V - 1) / d->V;
    if (per_v > cap) per_v = cap;
    if (per_v < 1) per_v = 1;
    kern<<<dim3(per_v, d->V), FWD_WARPS * 32, smem, st>>>(x, out, ymid, params, d->param_stride, rel_index, g);
    IDEE_LAUNCH_CHECK("swin_block_fwd");
    return 0;
}

int bwd_ctas_per_var(int V, int per_sm = 3) {
    int per_v = (idee_num_sms() * per_sm + V - 1) / V;
    return per_v < 1 ? 1 : per_v;
}

template <int WD, int WH, int WW>
int launch_bwd(const idee_swin_desc* d, const Geom& g, const float* x, const float* ymid, const float* gout, float* gx,
               const float* params, const int* rel_index, float* gparams, float* ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int G = WD * WH * WW;
    const int per_v = bwd_ctas_per_var(d->V, d->act_dtype == 1 ? 4 : 3);
    const int APS = ATT_PART_W + NH * G * G;
    const size_t need = sizeof(float) * (size_t)d->V * per_v * (APS + MLP_PART + (d->embed_gw ? 32 : 0));
    IDEE_REQUIRE(ws_bytes >= need, "swin_block_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
    float* part_attn = ws;
    float* part_mlp = ws + (size_t)d->V * per_v * APS;
    float* part_emb = part_mlp + (size_t)d->V * per_v * MLP_PART;
    const int64_t thw = (int64_t)d->T * d->H * d->W;
    if (d->precision == 1 && d->act_dtype == 1) {
        if (G < 8) { idee_set_error("swin_block_bwd(bf16 tokens): windows with fewer than 8 tokens are only built for the fp32 path"); return 1; }
        IDEE_REQUIRE((int64_t)d->N * thw < (1ll << 31), "swin_block_bwd(umma): N*T*H*W must be below 2^31");
        IDEE_REQUIRE(g.emb_x == nullptr || d->embed_gw != nullptr, "swin_block_bwd(umma): the fused embedding needs embed_gw / embed_gb");
        IDEE_REQUIRE(gx != gout, "swin_block_bwd(umma): gx must not alias gout");
        // persistent grids of one resident wave: 4 CTAs / SM (MLP half), 3 CTAs / SM (attention half)
        int pv_mlp = idee_num_sms() * 4 / d->V, pv_att = idee_num_sms() * 3 / d->V;
        if (pv_mlp < 1) pv_mlp = 1;
        if (pv_att < 1) pv_att = 1;
        float* u_attn = ws;
        float* u_mlp = u_attn + (size_t)d->V * pv_att * APS;
        float* u_emb = u_mlp + (size_t)d->V * pv_mlp * MLP_PART;
        const void* bx = x;
        auto bym = reinterpret_cast<const __nv_bfloat16*>(ymid);
        auto bgo = reinterpret_cast<const __nv_bfloat16*>(gout);
        auto bgx = reinterpret_cast<__nv_bfloat16*>(gx);
        {
            const size_t smem = sizeof(swu::MlpBwdSm);
            IDEE_CUDA(cudaFuncSetAttribute(swu::swin_mlp_bwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_mlp_bwd(umma)");
            swu::swin_mlp_bwd_umma_kernel<<<dim3(pv_mlp, d->V), swu::NT, smem, st>>>(bym, bgo, bgx, params, d->param_stride, g.tbl, u_mlp, d->N, d->V,
                                                                                     (int)thw, make_fastdiv((int)thw));
            IDEE_LAUNCH_CHECK("swin_mlp_bwd(umma)");
        }
        return swu::launch_attn_bwd<WD, WH, WW>(d, g, bx, bgx, params, rel_index, gparams, u_attn, u_mlp, u_emb, pv_att, pv_mlp, st);
    }
    if (d->precision == 1) {
        if (G < 8) { idee_set_error("swin_block_bwd(bf16): windows with fewer than 8 tokens are only built for the fp32 path"); return 1; }
        swin_mlp_bwd_tc_kernel<<<dim3(per_v, d->V), TCW * 32, 0, st>>>(ymid, gout, gx, params, d->param_stride, g.tbl, part_mlp, d->N, d->V, thw);
        IDEE_LAUNCH_CHECK("swin_mlp_bwd(bf16)");
        const size_t dyn = sizeof(float) * TCW * NH * G * dbs(G);
        if (g.emb_x) {
            Geom ge = g;
            ge.emb_gpart = d->embed_gw ? part_emb : nullptr;
            IDEE_CUDA(cudaFuncSetAttribute(swin_attn_bwd_tc_kernel<WD, WH, WW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "swin_attn_bwd(bf16)");
            swin_attn_bwd_tc_kernel<WD, WH, WW, true><<<dim3(per_v, d->V), TCW * 32, dyn, st>>>(x, gx, gx, params, d->param_stride, rel_index, part_attn, ge);
            IDEE_LAUNCH_CHECK("swin_attn_bwd(bf16,embed)");
            if (d->embed_gw) {
                embed_grad_finalize_kernel<<<d->V, 32, 0, st>>>(part_emb, per_v, d->embed_gw, d->embed_gb);
                IDEE_LAUNCH_CHECK("embed_grad_finalize");
            }
        } else {
            IDEE_CUDA(cudaFuncSetAttribute(swin_attn_bwd_tc_kernel<WD, WH, WW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn), "swin_attn_bwd(bf16)");
            swin_attn_bwd_tc_kernel<WD, WH, WW, false><<<dim3(per_v, d->V), TCW * 32, dyn, st>>>(x, gx, gx, params, d->param_stride, rel_index, part_attn, g);
            IDEE_LAUNCH_CHECK("swin_attn_bwd(bf16)");
        }
        if (launch_grad_finalize<G>(part_attn, part_mlp, per_v, per_v, rel_index, gparams, d->param_stride, g.tbl, d->V, st)) return 2;
        return 0;
    }
    IDEE_CUDA(cudaFuncSetAttribute(swin_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem)), "swin_mlp_bwd");
    // g_y is written into gx, then the attention half updates it in place
    swin_mlp_bwd_kernel<<<dim3(per_v, d->V), MLP_TOK, sizeof(MlpSmem), st>>>(ymid, gout, gx, params, d->param_stride, g.tbl,
                                                                              part_mlp, d->N, d->V, thw);
    IDEE_LAUNCH_CHECK("swin_mlp_bwd");
    auto kern = swin_attn_bwd_kernel<WD, WH, WW>;
    const size_t smem = attn_bwd_smem_bytes<G>();
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_attn_bwd");
    kern<<<dim3(per_v, d->V), AB_WARPS * 32, smem, st>>>(x, gx, gx, params, d->param_stride, rel_index, part_attn, g);
    IDEE_LAUNCH_CHECK("swin_attn_bwd");
    if (launch_grad_finalize<G>(part_attn, part_mlp, per_v, per_v, rel_index, gparams, d->param_stride, g.tbl, d->V, st)) return 2;
    return 0;
}

#define SWIN_DISPATCH(FN, ...)                                                                  \
    if (d->wd == 2 && d->wh == 4 && d->ww == 4) return FN<2, 4, 4>(__VA_ARGS__);               \
    if (d->wd == 8 && d->wh == 1 && d->ww == 1) return FN<8, 1, 1>(__VA_ARGS__);               \
    if (d->wd == 2 && d->wh == 2 && d->ww == 2) return FN<2, 2, 2>(__VA_ARGS__);               \
    if (d->wd == 4 && d->wh == 1 && d->ww == 1) return FN<4, 1, 1>(__VA_ARGS__);               \
    idee_set_error("swin_block: window (%d,%d,%d) is not built (built: (2,4,4) (8,1,1) (2,2,2) (4,1,1))", d->wd, d->wh, d->ww); \
    return 1;

}  // namespace

extern "C" int idee_swin_block_packed_floats(int rpb_rows) { return POff(rpb_rows).total; }

extern "C" size_t idee_swin_block_bwd_workspace_bytes(const idee_swin_desc* d) {
    const int Gt = d->wd * d->wh * d->ww;
    const int per_v = bwd_ctas_per_var(d->V, d->act_dtype == 1 ? 4 : 3);     // upper bound of either kernel's CTAs per variable
    return sizeof(float) * (size_t)d->V * per_v * (ATT_PART_W + NH * Gt * Gt + MLP_PART + (d->embed_gw ? 32 : 0));
}

extern "C" int idee_swin_block_fwd(const idee_swin_desc* d, const float* x, float* out, float* ymid, void* out_bf16,
                                   const float* params, const int32_t* rel_index, void* stream) {
    Geom g;
    if (make_geom(g, d, "swin_block_fwd")) return 1;
    IDEE_REQUIRE(out_bf16 == nullptr || d->precision == 1, "swin_block_fwd: the bf16 output copy is only produced by the bf16 path");
    IDEE_REQUIRE(d->act_dtype == 0 || (d->act_dtype == 1 && d->precision == 1 && out_bf16 == nullptr && out != nullptr),
                 "swin_block_fwd: bf16 tokens (act_dtype 1) need precision 1, out != NULL and out_bf16 == NULL");
    IDEE_REQUIRE(out != nullptr || out_bf16 != nullptr, "swin_block_fwd: out may only be NULL when the bf16 copy is requested");
    g.out16 = out_bf16;
    cudaStream_t st = (cudaStream_t)stream;
    SWIN_DISPATCH(launch_fwd, d, g, x, out, ymid, params, rel_index, st)
}

extern "C" int idee_swin_block_bwd(const idee_swin_desc* d, const float* x, const float* ymid, const float* gout, float* gx,
                                   const float* params, const int32_t* rel_index, float* gparams, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    Geom g;
    if (make_geom(g, d, "swin_block_bwd")) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SWIN_DISPATCH(launch_bwd, d, g, x, ymid, gout, gx, params, rel_index, gparams, (float*)workspace, workspace_bytes, st)
}
