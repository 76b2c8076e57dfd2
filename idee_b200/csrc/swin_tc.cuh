// bf16 tensor-core Swin block kernels (sm_100a): same math, partial layouts and C-ABI as the fp32 kernels in swin_block.cu,
// selected by idee_swin_desc.precision == 1.  Included by swin_block.cu inside its anonymous namespace.
//
// One warp owns 32 tokens (32/G whole windows) as two m16 tiles.  Every operand stays in registers in mma fragment layout:
//   * a [32 x 16] fp32 tile is float t[4][4]: row g+8r (g = lane/4), columns {c0, c0+1, c0+8, c0+9} (c0 = 2*(lane%4));
//     LayerNorm statistics are 4-lane (quad) shuffle reductions;
//   * GEMM outputs (D fragments) are re-packed to bf16 and fed straight back as A fragments (QKV -> QK^T -> softmax -> PV ->
//     proj -> fc1 -> GELU -> fc2) with no shared-memory round trip; transposed operands (V for PV, dS^T, P^T, and every
//     token-reduction for weight gradients) come from movmatrix.trans on 8x8 bf16 blocks;
//   * windows smaller than 32 tokens make the score matrix block diagonal: only the tiles on the diagonal are issued and
//     cross-window entries are forced to -inf / 0;
//   * weights live in shared memory pre-packed in B-fragment order (one conflict-free LDS.64 per fragment).
// Backward is recompute-based and split like the fp32 path (MLP half / attention half around the saved mid residual);
// weight-gradient accumulators stay in registers across the CTA's persistent loop and are reduced once per CTA.

constexpr float LOG2E = 1.4426950408889634f;   // attention scores are kept in the log2 domain (one ex2 per probability)

__device__ __forceinline__ uint32_t pk(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t movm(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

// ---- weight fragments in shared memory: frag f occupies 32 uint2 (b0,b1 of every lane) ----
// B(k,n) = W[(n0+n)*ld + k0+k]   ("T": Y = X W^T, W row-major [out][in])
__device__ __forceinline__ void pack_fragT(uint2* dst, const float* W, int ld, int k0, int n0, int lane) {
    const int n = n0 + lane / 4, c0 = 2 * (lane % 4);
    const float* r = W + n * ld + k0 + c0;
    dst[lane] = make_uint2(pk(r[0], r[1]), pk(r[8], r[9]));
}
// B(k,n) = W[(k0+k)*ld + n0+n]   ("N": Y = X W, W row-major [in][out])
__device__ __forceinline__ void pack_fragN(uint2* dst, const float* W, int ld, int k0, int n0, int lane) {
    const int n = n0 + lane / 4, c0 = 2 * (lane % 4);
    const float* r = W + (k0 + c0) * ld + n;
    dst[lane] = make_uint2(pk(r[0], r[ld]), pk(r[8 * ld], r[9 * ld]));
}

// A fragment (m16 x k16) of m-tile mi from a [32x16] tile
#define TILE_A(t, mi) pk(t[2 * (mi)][0], t[2 * (mi)][1]), pk(t[2 * (mi) + 1][0], t[2 * (mi) + 1][1]), \
                      pk(t[2 * (mi)][2], t[2 * (mi)][3]), pk(t[2 * (mi) + 1][2], t[2 * (mi) + 1][3])
// packed 8x8 block (token group ib = row index r, column n-tile nt) of a [32x16] tile, i.e. its ldmatrix-style fragment
#define TILE_BLK(t, ib, nt) pk(t[ib][2 * (nt)], t[ib][2 * (nt) + 1])

struct TokRows {
    bool valid[4];
    int64_t off[4];
    int code[4];
    bool masked;      // warp-uniform: some window of this warp touches the shifted border (its region codes differ)
};
// Token coordinates of the 4 rows (tokens g + 8r) of this lane.  The window index is decoded ONCE per warp iteration (three
// integer divisions) and advanced incrementally for the following windows of the warp; the in-window part uses compile-time
// divisors.  Same mapping as TokenMap (cyclic shift, padding, region codes).
template <int WD, int WH, int WW>
__device__ __forceinline__ void map_rows(TokRows& tr, const Geom& g, int v, int wg, bool wg_ok, int lane) {
    constexpr int G = WD * WH * WW;
    const int gq = lane / 4;
    const int win0 = wg * (32 / G);
    int n = win0 / g.nwin_img;
    int rem = win0 - n * g.nwin_img;
    int dw = rem / (g.nwh * g.nww);
    rem -= dw * g.nwh * g.nww;
    int hw = rem / g.nww, ww = rem - hw * g.nww;
    int cur = 0;
    bool any_border = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int wl = (8 * r) / G;                  // window of this row inside the warp (compile time)
        while (cur < wl) {                           // advance to the next window (row-major over ww, hw, dw, n)
            ++cur;
            if (++ww == g.nww) { ww = 0; if (++hw == g.nwh) { hw = 0; if (++dw == g.nwt) { dw = 0; ++n; } } }
        }
        const int i = (8 * r) % G + gq;
        const int dl = i / (WH * WW), hl = (i / WW) % WH, wl_ = i % WW;
        const int pt = dw * WD + dl, ph = hw * WH + hl, pw = ww * WW + wl_;
        int s_t = pt + g.st; if (s_t >= g.Tp) s_t -= g.Tp;
        int s_h = ph + g.sh; if (s_h >= g.Hp) s_h -= g.Hp;
        int s_w = pw + g.sw; if (s_w >= g.Wp) s_w -= g.Wp;
        tr.valid[r] = wg_ok && n < g.N && s_t < g.T && s_h < g.H && s_w < g.W;
        // host guarantees T*H*W*C < 2^31: one widening multiply for the image base, 32-bit math inside the image
        tr.off[r] = (int64_t)(n * g.V + v) * g.thwc + (((s_t * g.H + s_h) * g.W + s_w) * C);
        // Only the last window along a shifted axis mixes regions; everywhere else every token has region code 0, which is
        // also what the formula gives, so the codes (and the mask pass of the softmax) are skipped for interior windows.
        const bool border = g.masked && ((g.st && dw == g.nwt - 1) || (g.sh && hw == g.nwh - 1) || (g.sw && ww == g.nww - 1));
        tr.code[r] = 0;
        if (border) {
            tr.code[r] = region_id(pt, g.Tp, WD, g.st) * 9 + region_id(ph, g.Hp, WH, g.sh) * 3 + region_id(pw, g.Wp, WW, g.sw);
            any_border = true;
        }
    }
    tr.masked = any_border;
}
__device__ __forceinline__ void load_tile(float (&t)[4][4], const float* base, const TokRows& tr, int c0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (tr.valid[r]) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(base + tr.off[r] + c0));
            const float2 b = __ldg(reinterpret_cast<const float2*>(base + tr.off[r] + c0 + 8));
            t[r][0] = a.x; t[r][1] = a.y; t[r][2] = b.x; t[r][3] = b.y;
        } else { t[r][0] = t[r][1] = t[r][2] = t[r][3] = 0.f; }
    }
}
__device__ __forceinline__ void store_tile(float* base, const float (&t)[4][4], const TokRows& tr, int c0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (tr.valid[r]) {
            *reinterpret_cast<float2*>(base + tr.off[r] + c0) = make_float2(t[r][0], t[r][1]);
            *reinterpret_cast<float2*>(base + tr.off[r] + c0 + 8) = make_float2(t[r][2], t[r][3]);
        }
}
__device__ __forceinline__ void store_tile_bf16(__nv_bfloat16* base, const float (&t)[4][4], const TokRows& tr, int c0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (tr.valid[r]) {
            *reinterpret_cast<__nv_bfloat162*>(base + tr.off[r] + c0) = __floats2bfloat162_rn(t[r][0], t[r][1]);
            *reinterpret_cast<__nv_bfloat162*>(base + tr.off[r] + c0 + 8) = __floats2bfloat162_rn(t[r][2], t[r][3]);
        }
}
// LayerNorm(16, eps 1e-5, no affine) of every row; rows flagged invalid become exact zeros (padding after LN1)
__device__ __forceinline__ void ln_tile(const float (&x)[4][4], float (&xn)[4][4], float (&rstd)[4], const bool* valid) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float mu = quad_sum(x[r][0] + x[r][1] + x[r][2] + x[r][3]) * (1.f / 16.f);
        float d[4], s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { d[q] = x[r][q] - mu; s += d[q] * d[q]; }
        const float rs = rsqrtf(quad_sum(s) * (1.f / 16.f) + 1e-5f);
        rstd[r] = rs;
#pragma unroll
        for (int q = 0; q < 4; ++q) xn[r][q] = (valid == nullptr || valid[r]) ? d[q] * rs : 0.f;
    }
}
// Block input tile: either the stored tokens, or (fused patch embedding, in_chans == 1) LayerNorm(w[c] * x + b[c]) evaluated from
// the raw scalar input -- 4 bytes instead of 64 per token, and the embedded tokens never exist in HBM.
__device__ __forceinline__ void load_x_tile(float (&t)[4][4], const float* x, const TokRows& tr, int c0, const Geom& g, int v) {
    if (g.emb_x == nullptr) { load_tile(t, x, tr, c0); return; }
    const float* w = g.emb_w + v * C;
    const float* b = g.emb_b + v * C;
    const float w0 = __ldg(w + c0), w1 = __ldg(w + c0 + 1), w2 = __ldg(w + c0 + 8), w3 = __ldg(w + c0 + 9);
    const float b0 = __ldg(b + c0), b1 = __ldg(b + c0 + 1), b2 = __ldg(b + c0 + 8), b3 = __ldg(b + c0 + 9);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float xin = tr.valid[r] ? __ldg(g.emb_x + (tr.off[r] >> 4)) : 0.f;     // raw index = token index (C == 16)
        t[r][0] = w0 * xin + b0; t[r][1] = w1 * xin + b1; t[r][2] = w2 * xin + b2; t[r][3] = w3 * xin + b3;
    }
    float rs[4];
    ln_tile(t, t, rs, tr.valid);
}
// g_x = rstd * (g - mean(g) - xn * mean(g * xn))
__device__ __forceinline__ void ln_bwd_tile(const float (&g)[4][4], const float (&xn)[4][4], const float (&rstd)[4], float (&gx)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { s1 += g[r][q]; s2 += g[r][q] * xn[r][q]; }
        const float m1 = quad_sum(s1) * (1.f / 16.f), m2 = quad_sum(s2) * (1.f / 16.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) gx[r][q] = rstd[r] * (g[r][q] - m1 - xn[r][q] * m2);
    }
}
// out[32 x 16] (+)= in[32 x 16] * B, B given as the two n-tile fragments wf[0], wf[1] (each 32 uint2)
__device__ __forceinline__ void gemm16(float (&out)[4][4], const float (&in)[4][4], const uint2* wf, int lane) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const uint2 b = wf[nt * 32 + lane];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            float c[4] = {out[2 * mi][2 * nt], out[2 * mi][2 * nt + 1], out[2 * mi + 1][2 * nt], out[2 * mi + 1][2 * nt + 1]};
            mma16816(c, TILE_A(in, mi), b.x, b.y);
            out[2 * mi][2 * nt] = c[0]; out[2 * mi][2 * nt + 1] = c[1]; out[2 * mi + 1][2 * nt] = c[2]; out[2 * mi + 1][2 * nt + 1] = c[3];
        }
    }
}

// attention for both heads from the q/k/v tiles; produces the O tile and (optionally) keeps P, used by forward and backward
template <int G>
struct AttnTC {
    // packed operands
    uint32_t qa[2][2][2];   // [mi][h][half]   A (k8) fragments of scaled q
    uint32_t kb[2][4];      // [h][jn]         B (k8) fragments of k (token group jn)
    uint32_t vb[2][4];      // [h][jn]         packed v blocks (row j, cols e)
    static __device__ __forceinline__ bool tile_needed(int r, int nj) { return (8 * r) / G == (8 * nj) / G; }

    __device__ __forceinline__ void pack(const float (&q)[4][4], const float (&k)[4][4], const float (&v)[4][4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                qa[r / 2][h][r % 2] = TILE_BLK(q, r, h);
                kb[h][r] = TILE_BLK(k, r, h);
                vb[h][r] = TILE_BLK(v, r, h);
            }
        }
    }
    // scores + bias + mask + softmax for head h: p[mi][nj][4] (probabilities, 0 outside the window).  Everything is in
    // the log2 domain (q, the bias table and the mask constant carry a factor log2(e)), so the exponential is one ex2.
    // NORMALISE=false leaves p unnormalised and returns 1/rowsum in rinv (the forward pass scales O instead of P).
    template <bool NORMALISE>
    __device__ __forceinline__ void probs(int h, float (&p)[2][4][4], float (&rinv)[4], const float* Bn, const TokRows& tr, bool masked, int lane) {
        const int g = lane / 4, c0 = 2 * (lane % 4);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                p[mi][nj][0] = p[mi][nj][1] = p[mi][nj][2] = p[mi][nj][3] = 0.f;
                if (tile_needed(2 * mi, nj) || tile_needed(2 * mi + 1, nj)) mma1688(p[mi][nj], qa[mi][h][0], qa[mi][h][1], kb[h][nj]);
            }
        int cj[4][2];
        if (masked) {
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                cj[nj][0] = __shfl_sync(0xffffffffu, tr.code[nj], 4 * c0);
                cj[nj][1] = __shfl_sync(0xffffffffu, tr.code[nj], 4 * (c0 + 1));
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int mi = r / 2, hf = r % 2;
            const int il = (g + 8 * r) % G;
            float mx = -INFINITY;
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                if (tile_needed(r, nj)) {
                    const int jl = (8 * nj + c0) % G;
                    const float2 b = *reinterpret_cast<const float2*>(Bn + (h * G + il) * G + jl);
                    p[mi][nj][2 * hf] += b.x; p[mi][nj][2 * hf + 1] += b.y;
                }
            }
            if (masked) {                                  // warp-uniform: only warps holding a border window take this
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    if (tile_needed(r, nj)) {
                        if (cj[nj][0] != tr.code[r]) p[mi][nj][2 * hf] += -100.0f * LOG2E;
                        if (cj[nj][1] != tr.code[r]) p[mi][nj][2 * hf + 1] += -100.0f * LOG2E;
                    }
                }
            }
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
                if (tile_needed(r, nj)) mx = fmaxf(mx, fmaxf(p[mi][nj][2 * hf], p[mi][nj][2 * hf + 1]));
            mx = quad_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                if (tile_needed(r, nj)) {
                    const float e0 = ex2_approx(p[mi][nj][2 * hf] - mx), e1 = ex2_approx(p[mi][nj][2 * hf + 1] - mx);
                    p[mi][nj][2 * hf] = e0; p[mi][nj][2 * hf + 1] = e1; sum += e0 + e1;
                } else { p[mi][nj][2 * hf] = 0.f; p[mi][nj][2 * hf + 1] = 0.f; }
            }
            const float inv = __fdividef(1.f, quad_sum(sum));
            if (NORMALISE) {
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) { p[mi][nj][2 * hf] *= inv; p[mi][nj][2 * hf + 1] *= inv; }
            }
            rinv[r] = inv;
        }
    }
    // o (+)= P V_h into columns of head h of the O tile
    __device__ __forceinline__ void pv(int h, const float (&p)[2][4][4], float (&o)[4][4]) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const uint32_t b0 = movm(vb[h][2 * kk]), b1 = movm(vb[h][2 * kk + 1]);
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                if (!(tile_needed(2 * mi, 2 * kk) || tile_needed(2 * mi, 2 * kk + 1) || tile_needed(2 * mi + 1, 2 * kk) || tile_needed(2 * mi + 1, 2 * kk + 1))) continue;
                float c[4] = {o[2 * mi][2 * h], o[2 * mi][2 * h + 1], o[2 * mi + 1][2 * h], o[2 * mi + 1][2 * h + 1]};
                mma16816(c, pk(p[mi][2 * kk][0], p[mi][2 * kk][1]), pk(p[mi][2 * kk][2], p[mi][2 * kk][3]),
                         pk(p[mi][2 * kk + 1][0], p[mi][2 * kk + 1][1]), pk(p[mi][2 * kk + 1][2], p[mi][2 * kk + 1][3]), b0, b1);
                o[2 * mi][2 * h] = c[0]; o[2 * mi][2 * h + 1] = c[1]; o[2 * mi + 1][2 * h] = c[2]; o[2 * mi + 1][2 * h + 1] = c[3];
            }
        }
    }
};

// t = bias broadcast over rows (accumulator initialisation: the following gemm16 adds the product onto it)
__device__ __forceinline__ void init_bias_tile(float (&t)[4][4], const float* b, int c0) {
    const float b0 = b[c0], b1 = b[c0 + 1], b2 = b[c0 + 8], b3 = b[c0 + 9];
#pragma unroll
    for (int r = 0; r < 4; ++r) { t[r][0] = b0; t[r][1] = b1; t[r][2] = b2; t[r][3] = b3; }
}
__device__ __forceinline__ void add_bias_tile(float (&t)[4][4], const float* b, int c0) {
    const float b0 = b[c0], b1 = b[c0 + 1], b2 = b[c0 + 8], b3 = b[c0 + 9];
#pragma unroll
    for (int r = 0; r < 4; ++r) { t[r][0] += b0; t[r][1] += b1; t[r][2] += b2; t[r][3] += b3; }
}

template <int G>
__device__ __forceinline__ void stage_bias_n(float* Bn, const float* tbl, const int* __restrict__ rel_index) {
    for (int e = threadIdx.x; e < NH * G * G; e += blockDim.x) {
        const int h = e / (G * G), ij = e % (G * G);
        Bn[e] = tbl[rel_index[ij] * NH + h] * 1.4426950408889634f;      // log2 domain
    }
}

// q/k/v tiles from the normalised input tile (q scaled)
__device__ __forceinline__ void qkv_tiles(const float (&xn)[4][4], const uint2* wq, const float* bq, float scale,
                                          float (&q)[4][4], float (&k)[4][4], float (&v)[4][4], int lane) {
    const int c0 = 2 * (lane % 4);
    init_bias_tile(q, bq, c0);      gemm16(q, xn, wq, lane);
    init_bias_tile(k, bq + 16, c0); gemm16(k, xn, wq + 64, lane);
    init_bias_tile(v, bq + 32, c0); gemm16(v, xn, wq + 128, lane);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int x = 0; x < 4; ++x) q[r][x] *= scale;
}

// =====================================================================================================
// forward
// =====================================================================================================
constexpr int TCW = 4;   // warps per CTA
// fragment table (uint2[32] each): qkv^T 6 | proj^T 2 | fc1^T 8 | fc2^T (4 k-steps x 2 n-tiles) 8
constexpr int FW_FRAGS = 24;

template <int WD, int WH, int WW, bool EMB>
__global__ void __launch_bounds__(TCW * 32)
swin_fwd_tc_kernel(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ ymid, const float* __restrict__ params,
                   int64_t pstride, const int* __restrict__ rel_index, Geom g) {
    constexpr int G = WD * WH * WW;
    __shared__ __align__(16) uint2 wf[FW_FRAGS * 32];
    __shared__ __align__(16) float bias_s[3 * C + C + HID + C];
    __shared__ __align__(16) float Bn[bsz(G)];
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);
    for (int f = warp; f < FW_FRAGS; f += TCW) {
        if (f < 6) pack_fragT(wf + f * 32, P + po.qkv_w, C, 0, 8 * f, lane);
        else if (f < 8) pack_fragT(wf + f * 32, P + po.proj_w, C, 0, 8 * (f - 6), lane);
        else if (f < 16) pack_fragT(wf + f * 32, P + po.fc1_w, C, 0, 8 * (f - 8), lane);
        else pack_fragT(wf + f * 32, P + po.fc2_w, HID, 16 * ((f - 16) / 2), 8 * ((f - 16) % 2), lane);
    }
    for (int e = tid; e < 3 * C; e += blockDim.x) bias_s[e] = P[po.qkv_b + e];
    for (int e = tid; e < C; e += blockDim.x) { bias_s[3 * C + e] = P[po.proj_b + e]; bias_s[4 * C + HID + e] = P[po.fc2_b + e]; }
    for (int e = tid; e < HID; e += blockDim.x) bias_s[4 * C + e] = P[po.fc1_b + e];
    stage_bias_n<G>(Bn, P, rel_index);
    __syncthreads();
    const int c0 = 2 * (lane % 4);

    for (int wg = blockIdx.x * TCW + warp; wg < g.n_wg; wg += gridDim.x * TCW) {
        TokRows tr;
        map_rows<WD, WH, WW>(tr, g, v, wg, true, lane);
        float xt[4][4], y[4][4];
        if (EMB) load_x_tile(xt, x, tr, c0, g, v); else load_tile(xt, x, tr, c0);
        {
            float xn[4][4], rstd[4], q[4][4], k[4][4], vv[4][4], o[4][4];
            ln_tile(xt, xn, rstd, tr.valid);
            qkv_tiles(xn, wf, bias_s, g.scale * LOG2E, q, k, vv, lane);
            AttnTC<G> at;
            at.pack(q, k, vv);
#pragma unroll
            for (int r = 0; r < 4; ++r) o[r][0] = o[r][1] = o[r][2] = o[r][3] = 0.f;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float p[2][4][4], rinv[4];
                at.template probs<false>(h, p, rinv, Bn, tr, tr.masked, lane);
                at.pv(h, p, o);
#pragma unroll
                for (int r = 0; r < 4; ++r) { o[r][2 * h] *= rinv[r]; o[r][2 * h + 1] *= rinv[r]; }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) y[r][qd] = xt[r][qd];
            add_bias_tile(y, bias_s + 3 * C, c0);
            gemm16(y, o, wf + 6 * 32, lane);          // y = x + proj(o) + b
        }
        if (ymid) store_tile(ymid, y, tr, c0);
        float yn[4][4], rstd2[4], acc[4][4];
        ln_tile(y, yn, rstd2, nullptr);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) acc[r][qd] = y[r][qd];
        add_bias_tile(acc, bias_s + 4 * C + HID, c0);
        // hidden units in 4 chunks of 16: fc1 -> gelu -> fc2 k-step
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float h[4][4];
            init_bias_tile(h, bias_s + 4 * C + 16 * kk, c0);
            gemm16(h, yn, wf + (8 + 2 * kk) * 32, lane);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) h[r][qd] = gelu_fast(h[r][qd]);
            gemm16(acc, h, wf + (16 + 2 * kk) * 32, lane);
        }
        if (out) store_tile(out, acc, tr, c0);            // NULL: the consumer only reads the bf16 copy
        if (g.out16) store_tile_bf16(reinterpret_cast<__nv_bfloat16*>(g.out16), acc, tr, c0);
    }
}

// =====================================================================================================
// backward, MLP half
// =====================================================================================================
// fragment table: fc1^T 8 | fc2 "N" (k = c, n = hidden) 8 | fc1 "N" (k = hidden 4 k-steps, n = c 2 n-tiles) 8
constexpr int MB_FRAGS = 24;

__global__ void __launch_bounds__(TCW * 32, 3)
swin_mlp_bwd_tc_kernel(const float* __restrict__ y, const float* __restrict__ gout, float* __restrict__ gy,
                       const float* __restrict__ params, int64_t pstride, int tbl, float* __restrict__ partials,
                       int N, int V, int64_t thw) {
    __shared__ __align__(16) uint2 wf[MB_FRAGS * 32];
    __shared__ float b1_s[HID];
    __shared__ float red[MLP_PART];
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(tbl);
    for (int f = warp; f < MB_FRAGS; f += TCW) {
        if (f < 8) pack_fragT(wf + f * 32, P + po.fc1_w, C, 0, 8 * f, lane);
        else if (f < 16) pack_fragN(wf + f * 32, P + po.fc2_w, HID, 0, 8 * (f - 8), lane);
        else pack_fragN(wf + f * 32, P + po.fc1_w, C, 16 * ((f - 16) / 2), 8 * ((f - 16) % 2), lane);
    }
    for (int e = tid; e < HID; e += blockDim.x) b1_s[e] = P[po.fc1_b + e];
    for (int e = tid; e < MLP_PART; e += blockDim.x) red[e] = 0.f;
    __syncthreads();
    const int g = lane / 4, c0 = 2 * (lane % 4);
    // persistent accumulators (D fragments): dW2[c][k]: m-tile c(16) x 8 n-tiles ; dW1[k][c]: 4 m-tiles x 2 n-tiles
    float aW2[8][4], aW1[4][2][4], ab1[4][4], ab2[4];
#pragma unroll
    for (int a = 0; a < 8; ++a) aW2[a][0] = aW2[a][1] = aW2[a][2] = aW2[a][3] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        ab2[a] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) ab1[a][b] = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) aW1[a][b][0] = aW1[a][b][1] = aW1[a][b][2] = aW1[a][b][3] = 0.f;
    }
    const int64_t ntok = (int64_t)N * thw;
    const int64_t n_grp = (ntok + 31) / 32;
    // (sample, in-sample token) of the warp's first token, advanced incrementally: no division inside the loop
    const int64_t grp0 = (int64_t)blockIdx.x * TCW + warp, gstep = (int64_t)gridDim.x * TCW;
    int64_t n0 = (grp0 * 32) / thw, rem0 = grp0 * 32 - n0 * thw;
    for (int64_t grp = grp0; grp < n_grp; grp += gstep) {
        TokRows tr;
        tr.masked = false;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t tok = grp * 32 + g + 8 * r;
            tr.valid[r] = tok < ntok;
            int64_t n = n0, rem = rem0 + g + 8 * r;
            while (rem >= thw) { rem -= thw; ++n; }
            tr.off[r] = tr.valid[r] ? ((n * V + v) * thw + rem) * C : 0;
            tr.code[r] = 0;
        }
        rem0 += gstep * 32;
        while (rem0 >= thw) { rem0 -= thw; ++n0; }
        float yt[4][4], go[4][4], yn[4][4], rstd[4], dyn[4][4];
        load_tile(yt, y, tr, c0);
        load_tile(go, gout, tr, c0);
        ln_tile(yt, yn, rstd, tr.valid);
#pragma unroll
        for (int r = 0; r < 4; ++r) dyn[r][0] = dyn[r][1] = dyn[r][2] = dyn[r][3] = 0.f;
        // token-transposed blocks of go / yn for the weight-gradient GEMMs (K = tokens)
        uint32_t goT[4][2], ynT[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { goT[r][nt] = movm(TILE_BLK(go, r, nt)); ynT[r][nt] = movm(TILE_BLK(yn, r, nt)); }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) ab2[qd] += go[r][qd];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {          // 16 hidden units per chunk
            float pre[4][4], dh[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) dh[r][0] = dh[r][1] = dh[r][2] = dh[r][3] = 0.f;
            init_bias_tile(pre, b1_s + 16 * kk, c0);
            gemm16(pre, yn, wf + (2 * kk) * 32, lane);
            gemm16(dh, go, wf + (8 + 2 * kk) * 32, lane);        // dHid = dOut W2
            float hid[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    // rows past the end load go = 0, so dHid, dPre and their weight-gradient terms vanish without a mask
                    float dg;
                    gelu_fast_grad(pre[r][qd], hid[r][qd], dg);
                    dh[r][qd] *= dg;                                  // dPre
                    ab1[kk][qd] += dh[r][qd];
                }
            gemm16(dyn, dh, wf + (16 + 2 * kk) * 32, lane);      // dYn += dPre W1[chunk]
            // dW2[c][k] += go^T hid   (M = c, N = 16 hidden of this chunk as 2 n-tiles, K = tokens in 2 steps)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int ik = 0; ik < 2; ++ik)
                    mma16816(aW2[2 * kk + nt], goT[2 * ik][0], goT[2 * ik][1], goT[2 * ik + 1][0], goT[2 * ik + 1][1],
                             movm(TILE_BLK(hid, 2 * ik, nt)), movm(TILE_BLK(hid, 2 * ik + 1, nt)));
            // dW1[k][c] += dPre^T yn  (M = 16 hidden of this chunk, N = c as 2 n-tiles)
#pragma unroll
            for (int ik = 0; ik < 2; ++ik) {
                const uint32_t a0 = movm(TILE_BLK(dh, 2 * ik, 0)), a1 = movm(TILE_BLK(dh, 2 * ik, 1));
                const uint32_t a2 = movm(TILE_BLK(dh, 2 * ik + 1, 0)), a3 = movm(TILE_BLK(dh, 2 * ik + 1, 1));
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma16816(aW1[kk][nt], a0, a1, a2, a3, ynT[2 * ik][nt], ynT[2 * ik + 1][nt]);
            }
        }
        float gx[4][4];
        ln_bwd_tile(dyn, yn, rstd, gx);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) gx[r][qd] += go[r][qd];
        store_tile(gy, gx, tr, c0);
    }
    // CTA reduction of the per-warp accumulators, layout MLP_PART: fc1_w[k][c] | fc1_b[k] | fc2_w[c][k] | fc2_b[c]
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    atomicAdd(&red[(16 * kk + g + 8 * hf) * C + 8 * nt + c0 + b], aW1[kk][nt][2 * hf + b]);                       // dW1[k][c]
                    atomicAdd(&red[HID * C + HID + (g + 8 * hf) * HID + 16 * kk + 8 * nt + c0 + b], aW2[2 * kk + nt][2 * hf + b]);  // dW2[c][k]
                }
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) atomicAdd(&red[HID * C + 16 * kk + c0 + (qd & 1) + 8 * (qd >> 1)], ab1[kk][qd]);
    }
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) atomicAdd(&red[HID * C + HID + C * HID + c0 + (qd & 1) + 8 * (qd >> 1)], ab2[qd]);
    __syncthreads();
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * MLP_PART;
    for (int e = tid; e < MLP_PART; e += blockDim.x) part[e] = red[e];
}

// =====================================================================================================
// backward, attention half
// =====================================================================================================
// fragment table: qkv^T 6 | proj "N" (k = c, n = e) 2 | qkv "N" (k = o in 3 k-steps, n = c 2 n-tiles) 6
constexpr int AB_FRAGS = 14;
// row stride of the warp-private bias-gradient table: conflict-free 8-byte read-modify-write per half warp
__host__ __device__ constexpr int dbs(int G) { return G == 8 ? 8 : G + 8; }

// sums the per-CTA partials [V][nblocks][32] = (g_w[16] | g_b[16]) of the fused embedding backward
__global__ void embed_grad_finalize_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ gw, float* __restrict__ gb) {
    const int v = blockIdx.x, k = threadIdx.x;
    double a = 0.0;
    for (int b = 0; b < nblocks; ++b) a += part[((int64_t)v * nblocks + b) * 32 + k];
    if (k < 16) gw[v * 16 + k] = (float)a; else gb[v * 16 + k - 16] = (float)a;
}

template <int WD, int WH, int WW, bool EMB>
__global__ void __launch_bounds__(TCW * 32, 3)
swin_attn_bwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                        const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index,
                        float* __restrict__ partials, Geom g) {
    constexpr int G = WD * WH * WW;
    constexpr int PART = ATT_PART_W + NH * G * G;
    constexpr int DBS = dbs(G), DBW = NH * G * DBS;           // per-warp bias-gradient table
    __shared__ __align__(16) uint2 wf[AB_FRAGS * 32];
    __shared__ __align__(16) float bq_s[3 * C];
    __shared__ __align__(16) float Bn[bsz(G)];
    __shared__ float red[ATT_PART_W];
    extern __shared__ __align__(16) float dBw_all[];          // [TCW][DBW]
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);
    for (int f = warp; f < AB_FRAGS; f += TCW) {
        if (f < 6) pack_fragT(wf + f * 32, P + po.qkv_w, C, 0, 8 * f, lane);
        else if (f < 8) pack_fragN(wf + f * 32, P + po.proj_w, C, 0, 8 * (f - 6), lane);
        else pack_fragN(wf + f * 32, P + po.qkv_w, C, 16 * ((f - 8) / 2), 8 * ((f - 8) % 2), lane);
    }
    for (int e = tid; e < 3 * C; e += blockDim.x) bq_s[e] = P[po.qkv_b + e];
    for (int e = tid; e < ATT_PART_W; e += blockDim.x) red[e] = 0.f;
    for (int e = tid; e < TCW * DBW; e += blockDim.x) dBw_all[e] = 0.f;
    stage_bias_n<G>(Bn, P, rel_index);
    __syncthreads();
    const int gq = lane / 4, c0 = 2 * (lane % 4);
    float* dB = dBw_all + warp * DBW;

    // persistent accumulators: dWqkv[o][c] 3 m-tiles x 2 n-tiles ; dWproj[c][e] 1 x 2 ; biases per-lane column sums
    // bias-gradient column sums (4 float4 per lane) live in shared memory: 8 128-bit accesses per tile instead of 16 registers
    // held across the whole loop of a kernel that already spills
    __shared__ __align__(16) float4 bias_acc[TCW * 4 * 32];
    float aWq[3][2][4], aWp[2][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) bias_acc[(warp * 4 + a) * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) aWq[a][b][0] = aWq[a][b][1] = aWq[a][b][2] = aWq[a][b][3] = 0.f;
#pragma unroll
    for (int b = 0; b < 2; ++b) aWp[b][0] = aWp[b][1] = aWp[b][2] = aWp[b][3] = 0.f;
#pragma unroll
    // fused embedding backward (EMB with emb_gpart): its 8 running sums per lane live in shared memory (two float4 per lane and
    // tile), not in registers -- this kernel already spills
    __shared__ __align__(16) float4 emb_acc[EMB ? TCW * 2 * 32 : 1];
    const bool emb_bwd = EMB && g.emb_gpart != nullptr;
    if (EMB) { emb_acc[(warp * 2) * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); emb_acc[(warp * 2 + 1) * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); }

    const int n_iter = (g.n_wg + gridDim.x * TCW - 1) / (gridDim.x * TCW);
    for (int it = 0; it < n_iter; ++it) {
        const int wg = (it * gridDim.x + blockIdx.x) * TCW + warp;
        const bool wg_ok = wg < g.n_wg;
        TokRows tr;
        map_rows<WD, WH, WW>(tr, g, v, wg_ok ? wg : 0, wg_ok, lane);
        float xt[4][4], xn[4][4], rstd[4], ga[4][4];
        if (EMB) load_x_tile(xt, x, tr, c0, g, v); else load_tile(xt, x, tr, c0);
        load_tile(ga, gy, tr, c0);
        ln_tile(xt, xn, rstd, tr.valid);
        float q[4][4], k[4][4], vv[4][4];
        qkv_tiles(xn, wf, bq_s, g.scale * LOG2E, q, k, vv, lane);
        AttnTC<G> at;
        at.pack(q, k, vv);
        float dot[4][4];                                  // dO = dY Wproj
#pragma unroll
        for (int r = 0; r < 4; ++r) dot[r][0] = dot[r][1] = dot[r][2] = dot[r][3] = 0.f;
        gemm16(dot, ga, wf + 6 * 32, lane);
        float o[4][4], dq[4][4], dk[4][4], dv[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { o[r][qd] = 0.f; dq[r][qd] = 0.f; dk[r][qd] = 0.f; dv[r][qd] = 0.f; }
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float p[2][4][4], rinv[4];
            at.template probs<true>(h, p, rinv, Bn, tr, tr.masked, lane);
            at.pv(h, p, o);
            // D_r = <dO_r, O_r> over the 8 dims of head h
            float Dr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) Dr[r] = quad_sum(dot[r][2 * h] * o[r][2 * h] + dot[r][2 * h + 1] * o[r][2 * h + 1]);
            // dP = dO_h V_h^T, dS = P o (dP - D)
            float ds[2][4][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    ds[mi][nj][0] = ds[mi][nj][1] = ds[mi][nj][2] = ds[mi][nj][3] = 0.f;
                    if (AttnTC<G>::tile_needed(2 * mi, nj) || AttnTC<G>::tile_needed(2 * mi + 1, nj)) {
                        mma1688(ds[mi][nj], TILE_BLK(dot, 2 * mi, h), TILE_BLK(dot, 2 * mi + 1, h), at.vb[h][nj]);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                            for (int b = 0; b < 2; ++b) ds[mi][nj][2 * hf + b] = p[mi][nj][2 * hf + b] * (ds[mi][nj][2 * hf + b] - Dr[2 * mi + hf]);
                    }
                }
            // relative-position-bias gradient: dB[h][i_local][j_local] += dS.  The table is private to the warp and every
            // (i_local, j_local) pair is owned by exactly one lane, so a plain 8-byte read-modify-write is race-free.
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj)
                    if (AttnTC<G>::tile_needed(r, nj)) {
                        const int il = (gq + 8 * r) % G, jl = (8 * nj + c0) % G;
                        float2* pd = reinterpret_cast<float2*>(&dB[(h * G + il) * DBS + jl]);
                        float2 acc2 = *pd;
                        acc2.x += ds[r / 2][nj][2 * (r % 2)]; acc2.y += ds[r / 2][nj][2 * (r % 2) + 1];
                        *pd = acc2;
                    }
            // packed 8x8 blocks of dS and P: blk[ib][jb]
            uint32_t dsb[4][4], pb[4][4];
#pragma unroll
            for (int ib = 0; ib < 4; ++ib)
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) {
                    dsb[ib][jb] = pk(ds[ib / 2][jb][2 * (ib % 2)], ds[ib / 2][jb][2 * (ib % 2) + 1]);
                    pb[ib][jb] = pk(p[ib / 2][jb][2 * (ib % 2)], p[ib / 2][jb][2 * (ib % 2) + 1]);
                }
            // dQ_h = dS K_h  (K = tokens j)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint32_t b0 = movm(at.kb[h][2 * kk]), b1 = movm(at.kb[h][2 * kk + 1]);
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    float c[4] = {dq[2 * mi][2 * h], dq[2 * mi][2 * h + 1], dq[2 * mi + 1][2 * h], dq[2 * mi + 1][2 * h + 1]};
                    mma16816(c, dsb[2 * mi][2 * kk], dsb[2 * mi + 1][2 * kk], dsb[2 * mi][2 * kk + 1], dsb[2 * mi + 1][2 * kk + 1], b0, b1);
                    dq[2 * mi][2 * h] = c[0]; dq[2 * mi][2 * h + 1] = c[1]; dq[2 * mi + 1][2 * h] = c[2]; dq[2 * mi + 1][2 * h + 1] = c[3];
                }
            }
            // dK_h = dS^T Q_h, dV_h = P^T dO_h  (rows j, K = tokens i)
#pragma unroll
            for (int ik = 0; ik < 2; ++ik) {
                const uint32_t bq0 = movm(at.qa[ik][h][0]), bq1 = movm(at.qa[ik][h][1]);
                const uint32_t bo0 = movm(TILE_BLK(dot, 2 * ik, h)), bo1 = movm(TILE_BLK(dot, 2 * ik + 1, h));
#pragma unroll
                for (int jm = 0; jm < 2; ++jm) {
                    float ck[4] = {dk[2 * jm][2 * h], dk[2 * jm][2 * h + 1], dk[2 * jm + 1][2 * h], dk[2 * jm + 1][2 * h + 1]};
                    mma16816(ck, movm(dsb[2 * ik][2 * jm]), movm(dsb[2 * ik][2 * jm + 1]), movm(dsb[2 * ik + 1][2 * jm]), movm(dsb[2 * ik + 1][2 * jm + 1]), bq0, bq1);
                    dk[2 * jm][2 * h] = ck[0]; dk[2 * jm][2 * h + 1] = ck[1]; dk[2 * jm + 1][2 * h] = ck[2]; dk[2 * jm + 1][2 * h + 1] = ck[3];
                    float cv[4] = {dv[2 * jm][2 * h], dv[2 * jm][2 * h + 1], dv[2 * jm + 1][2 * h], dv[2 * jm + 1][2 * h + 1]};
                    mma16816(cv, movm(pb[2 * ik][2 * jm]), movm(pb[2 * ik][2 * jm + 1]), movm(pb[2 * ik + 1][2 * jm]), movm(pb[2 * ik + 1][2 * jm + 1]), bo0, bo1);
                    dv[2 * jm][2 * h] = cv[0]; dv[2 * jm][2 * h + 1] = cv[1]; dv[2 * jm + 1][2 * h] = cv[2]; dv[2 * jm + 1][2 * h + 1] = cv[3];
                }
            }
        }
        // gradient wrt the unscaled q projection
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { dq[r][qd] *= g.scale; dk[r][qd] *= (1.f / LOG2E); }   // dK was formed with q * scale * log2(e)
        // dXn = dQ Wq + dK Wk + dV Wv ; dX = dY + LN_bwd
        float dxn[4][4], gxt[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) dxn[r][0] = dxn[r][1] = dxn[r][2] = dxn[r][3] = 0.f;
        gemm16(dxn, dq, wf + 8 * 32, lane);
        gemm16(dxn, dk, wf + 10 * 32, lane);
        gemm16(dxn, dv, wf + 12 * 32, lane);
        ln_bwd_tile(dxn, xn, rstd, gxt);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) gxt[r][qd] += ga[r][qd];
        if (emb_bwd) {
            // the block input IS the embedding output: run its backward here (LN backward, then d/dw = sum ge * x, d/db = sum ge)
            // instead of writing the token gradient and reading it back in a separate kernel
            const float* w = g.emb_w + v * C;
            const float* b = g.emb_b + v * C;
            const float wq[4] = {__ldg(w + c0), __ldg(w + c0 + 1), __ldg(w + c0 + 8), __ldg(w + c0 + 9)};
            const float bq[4] = {__ldg(b + c0), __ldg(b + c0 + 1), __ldg(b + c0 + 8), __ldg(b + c0 + 9)};
            float xin[4], e[4][4], en[4][4], rs[4], ge[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                xin[r] = tr.valid[r] ? __ldg(g.emb_x + (tr.off[r] >> 4)) : 0.f;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) e[r][qd] = wq[qd] * xin[r] + bq[qd];
            }
            ln_tile(e, en, rs, nullptr);
            ln_bwd_tile(gxt, en, rs, ge);
            float4 sw4 = emb_acc[(warp * 2) * 32 + lane], sb4 = emb_acc[(warp * 2 + 1) * 32 + lane];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (tr.valid[r]) {
                    sw4.x += ge[r][0] * xin[r]; sw4.y += ge[r][1] * xin[r]; sw4.z += ge[r][2] * xin[r]; sw4.w += ge[r][3] * xin[r];
                    sb4.x += ge[r][0]; sb4.y += ge[r][1]; sb4.z += ge[r][2]; sb4.w += ge[r][3];
                }
            emb_acc[(warp * 2) * 32 + lane] = sw4; emb_acc[(warp * 2 + 1) * 32 + lane] = sb4;
        } else store_tile(gx, gxt, tr, c0);
        // weight gradients (K = the 32 tokens of this warp)
        uint32_t xnT[4][2], oT[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { xnT[r][nt] = movm(TILE_BLK(xn, r, nt)); oT[r][nt] = movm(TILE_BLK(o, r, nt)); }
#pragma unroll
        for (int ik = 0; ik < 2; ++ik) {
#pragma unroll
            for (int w3 = 0; w3 < 3; ++w3) {
                const float (&T)[4][4] = (w3 == 0) ? dq : (w3 == 1 ? dk : dv);
                const uint32_t a0 = movm(TILE_BLK(T, 2 * ik, 0)), a1 = movm(TILE_BLK(T, 2 * ik, 1));
                const uint32_t a2 = movm(TILE_BLK(T, 2 * ik + 1, 0)), a3 = movm(TILE_BLK(T, 2 * ik + 1, 1));
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma16816(aWq[w3][nt], a0, a1, a2, a3, xnT[2 * ik][nt], xnT[2 * ik + 1][nt]);
            }
            const uint32_t a0 = movm(TILE_BLK(ga, 2 * ik, 0)), a1 = movm(TILE_BLK(ga, 2 * ik, 1));
            const uint32_t a2 = movm(TILE_BLK(ga, 2 * ik + 1, 0)), a3 = movm(TILE_BLK(ga, 2 * ik + 1, 1));
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma16816(aWp[nt], a0, a1, a2, a3, oT[2 * ik][nt], oT[2 * ik + 1][nt]);
        }
        {
            float4 b4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) b4[a] = bias_acc[(warp * 4 + a) * 32 + lane];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                b4[0].x += dq[r][0]; b4[0].y += dq[r][1]; b4[0].z += dq[r][2]; b4[0].w += dq[r][3];
                b4[1].x += dk[r][0]; b4[1].y += dk[r][1]; b4[1].z += dk[r][2]; b4[1].w += dk[r][3];
                b4[2].x += dv[r][0]; b4[2].y += dv[r][1]; b4[2].z += dv[r][2]; b4[2].w += dv[r][3];
                b4[3].x += ga[r][0]; b4[3].y += ga[r][1]; b4[3].z += ga[r][2]; b4[3].w += ga[r][3];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) bias_acc[(warp * 4 + a) * 32 + lane] = b4[a];
        }
    }
    // CTA reduction, layout: qkv_w[48*16] | qkv_b[48] | proj_w[16*16] | proj_b[16] | dB
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int w3 = 0; w3 < 3; ++w3) atomicAdd(&red[(16 * w3 + gq + 8 * hf) * C + 8 * nt + c0 + b], aWq[w3][nt][2 * hf + b]);
                atomicAdd(&red[3 * C * C + 3 * C + (gq + 8 * hf) * C + 8 * nt + c0 + b], aWp[nt][2 * hf + b]);
            }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float4 b4 = bias_acc[(warp * 4 + a) * 32 + lane];
        const float vals[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const int col = c0 + (qd & 1) + 8 * (qd >> 1);
            atomicAdd(&red[a < 3 ? 3 * C * C + 16 * a + col : 3 * C * C + 3 * C + C * C + col], vals[qd]);
        }
    }
    if (emb_bwd) {                                        // rows live in lanes with equal lane % 4: reduce over lane / 4, then over the warps
        __shared__ float red_e[32];
        if (tid < 32) red_e[tid] = 0.f;
        __syncthreads();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 sw4 = emb_acc[(warp * 2) * 32 + lane], sb4 = emb_acc[(warp * 2 + 1) * 32 + lane];
            float sw_ = qd == 0 ? sw4.x : (qd == 1 ? sw4.y : (qd == 2 ? sw4.z : sw4.w));
            float sb_ = qd == 0 ? sb4.x : (qd == 1 ? sb4.y : (qd == 2 ? sb4.z : sb4.w));
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { sw_ += __shfl_xor_sync(0xffffffffu, sw_, o); sb_ += __shfl_xor_sync(0xffffffffu, sb_, o); }
            if (lane < 4) {
                const int col = c0 + (qd & 1) + 8 * (qd >> 1);
                atomicAdd(&red_e[col], sw_); atomicAdd(&red_e[16 + col], sb_);
            }
        }
        __syncthreads();
        if (tid < 32) g.emb_gpart[((int64_t)v * gridDim.x + blockIdx.x) * 32 + tid] = red_e[tid];
    }
    __syncthreads();
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * PART;
    for (int e = tid; e < ATT_PART_W; e += blockDim.x) part[e] = red[e];
    for (int e = tid; e < NH * G * G; e += blockDim.x) {
        const int hi = e / G, jl = e % G;     // hi = h * G + i_local
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < TCW; ++w) sum += dBw_all[w * DBW + hi * DBS + jl];
        part[ATT_PART_W + e] = sum;
    }
}
