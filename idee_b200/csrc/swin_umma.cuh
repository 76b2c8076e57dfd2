// Blackwell-native Swin block kernels (sm_100a): tcgen05.mma with TMEM accumulators for every linear layer and every weight
// gradient, bf16 token storage in HBM, TMA bulk copies (cp.async.bulk + mbarrier) for the token tiles.
// Selected by idee_swin_desc.precision == 1 && act_dtype == 1.  Included by swin_block.cu inside its anonymous namespace.
//
// A CTA (128 threads) owns tiles of 128 tokens = 4 warps x 32 tokens (32 / G whole windows per warp, window-major order as in
// the other paths).  Thread t owns token t of the tile = row t of every M = 128 MMA = TMEM lane t, so LayerNorm, bias, GELU and
// the residuals are thread-local (no shuffles) and the accumulators come back with one tcgen05.ld per 16 / 32 columns.
//   forward:  LN1 -> A operand written to TMEM (tcgen05.st) -> QKV = tcgen05.mma (A from TMEM, W^T from smem, N = 48)
//             -> q,k,v rows to a per-warp bf16 staging tile -> attention core per warp on mma.sync fragments loaded by ldmatrix
//                (32x32x8 per head and window: too small and block-diagonal for an M = 128 tile)
//             -> O to smem (canonical K-major core matrices) -> proj MMA -> +x -> LN2 -> fc1 MMA (N = 64) -> GELU -> fc2 MMA
//             (A from TMEM, K = 64 in four steps) -> +y -> bf16 store.
//   backward: same structure; activations (yn, g_out, g_pre, gelu(h) / xn, dq|dk|dv, g_y, o) live in smem as 8-channel chunk
//             planes [chunk][token][16 B].  One plane set is BOTH the K-major A operand of the data-gradient GEMMs (rows = tokens)
//             and, read through an MN-major descriptor, the transposed operand of the weight-gradient GEMMs (K = 128 tokens):
//             dW accumulates in TMEM across the CTA's whole persistent loop -- no registers, no per-warp HMMA, no movmatrix.
// Descriptor / instruction-descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp; the operand forms used here (K-major,
// MN-major SWIZZLE_NONE with LBO = K-group stride and SBO = MN-group stride, A from TMEM, M = 64 lane mapping) were checked on
// a B200 against a host reference by tools/umma_probe.cu.

namespace swu {

constexpr int NT = 128;                 // threads = tokens per tile
constexpr int STG = 112;                // staging row stride (bytes): q | k | v | pad -> conflict-free ldmatrix and STS.128
constexpr int PLANE = 2048;             // one 8-channel chunk plane of a 128-token tile (16 B per token)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D = F32, A = B = BF16; a_mn / b_mn: operand is MN-major (transposed view)
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                 "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// one row of an A operand into TMEM: 8 columns = 16 bf16 (column j = elements 2j, 2j+1 of the row)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ void unpack16(const uint32_t* p, float* x) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[2 * i] = bf_lo(p[i]); x[2 * i + 1] = bf_hi(p[i]); }
}
__device__ __forceinline__ void pack16(const float* x, uint32_t* p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
}
__device__ __forceinline__ void ldg_row16(const __nv_bfloat16* p, uint32_t* r) {      // 32 B, one 256-bit load
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]),
                 "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

// LayerNorm(16, eps 1e-5, no affine) of one token held by the thread; returns rstd
__device__ __forceinline__ float ln_row(const float* x, float* xn) {
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) mu += x[i];
    mu *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = x[i] - mu; xn[i] = d; var += d * d; }
    const float rs = rsqrtf(var * (1.f / 16.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) xn[i] *= rs;
    return rs;
}

// token of this thread: window-major order inside the warp (32 / G windows of G tokens), cyclic shift, padding, mask region code
struct Tok {
    bool valid;      // exists in the unpadded tensor
    int64_t off;     // element offset of its 16 channels
    int code;        // shift-mask region code (0 unless its window touches the shifted border)
    bool border;
};
template <int WD, int WH, int WW>
__device__ __forceinline__ Tok map_token(const Geom& g, int v, int wg, int lane) {
    constexpr int G = WD * WH * WW;
    Tok t;
    const int win = wg * (32 / G) + lane / G;
    const int i = lane % G;
    const bool active = wg < g.n_wg && win < g.N * g.nwin_img;
    const int n = win / g.nwin_img;
    int r = win - n * g.nwin_img;
    const int dw = r / (g.nwh * g.nww);
    r -= dw * g.nwh * g.nww;
    const int hw = r / g.nww, ww = r - hw * g.nww;
    const int dl = i / (WH * WW), hl = (i / WW) % WH, wl = i % WW;
    const int pt = dw * WD + dl, ph = hw * WH + hl, pw = ww * WW + wl;   // rolled frame
    int s_t = pt + g.st; if (s_t >= g.Tp) s_t -= g.Tp;                   // torch.roll(x, -shift)
    int s_h = ph + g.sh; if (s_h >= g.Hp) s_h -= g.Hp;
    int s_w = pw + g.sw; if (s_w >= g.Wp) s_w -= g.Wp;
    t.valid = active && s_t < g.T && s_h < g.H && s_w < g.W;
    t.off = t.valid ? (int64_t)(n * g.V + v) * g.thwc + (((s_t * g.H + s_h) * g.W + s_w) * C) : 0;
    t.border = active && g.masked && ((g.st && dw == g.nwt - 1) || (g.sh && hw == g.nwh - 1) || (g.sw && ww == g.nww - 1));
    t.code = t.border ? region_id(pt, g.Tp, WD, g.st) * 9 + region_id(ph, g.Hp, WH, g.sh) * 3 + region_id(pw, g.Wp, WW, g.sw) : 0;
    return t;
}

// attention core of one warp (32 tokens = 32 / G windows) on mma.sync fragments, operands from the warp's staging tile
template <int G>
struct AttnU {
    uint32_t qa[2][2][2];   // [mi][h][half]  A (k8) fragments of q (scaled, log2 domain)
    uint32_t kb[2][4];      // [h][key group] B (k8) fragments of k
    uint32_t vt[2][4];      // [h][key group] transposed 8x8 blocks of v (B fragments of P V)
    static __device__ __forceinline__ bool tile_needed(int r, int nj) { return (8 * r) / G == (8 * nj) / G; }

    // stg: shared address of the warp's [32][STG] staging tile; qoff / koff / voff: byte offsets of q, k, v inside a row
    __device__ __forceinline__ void load(uint32_t stg, int lane, int qoff, int koff, int voff) {
        const uint32_t a = stg + (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * STG + ((lane >> 3) & 1) * 16);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t ai = a + i * 16 * STG;
            ldsm_x4(ai + qoff, qa[i][0][0], qa[i][1][0], qa[i][0][1], qa[i][1][1]);
            ldsm_x4(ai + koff, kb[0][2 * i], kb[1][2 * i], kb[0][2 * i + 1], kb[1][2 * i + 1]);
            ldsm_x4_t(ai + voff, vt[0][2 * i], vt[1][2 * i], vt[0][2 * i + 1], vt[1][2 * i + 1]);
        }
    }
    // scores + bias + mask + softmax of head h (log2 domain).  code: region code of THIS lane's token; masked: warp-uniform
    template <bool NORMALISE>
    __device__ __forceinline__ void probs(int h, float (&p)[2][4][4], float (&rinv)[4], const float* Bn, int code, bool masked, int lane) {
        const int g = lane / 4, c0 = 2 * (lane % 4);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                p[mi][nj][0] = p[mi][nj][1] = p[mi][nj][2] = p[mi][nj][3] = 0.f;
                if (tile_needed(2 * mi, nj) || tile_needed(2 * mi + 1, nj)) mma1688(p[mi][nj], qa[mi][h][0], qa[mi][h][1], kb[h][nj]);
            }
        int cj[4][2], cr[4];
        if (masked) {
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                cj[nj][0] = __shfl_sync(0xffffffffu, code, 8 * nj + c0);
                cj[nj][1] = __shfl_sync(0xffffffffu, code, 8 * nj + c0 + 1);
                cr[nj] = __shfl_sync(0xffffffffu, code, g + 8 * nj);
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
            if (masked) {
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    if (tile_needed(r, nj)) {
                        if (cj[nj][0] != cr[r]) p[mi][nj][2 * hf] += -100.0f * LOG2E;
                        if (cj[nj][1] != cr[r]) p[mi][nj][2 * hf + 1] += -100.0f * LOG2E;
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
    // o (+)= P V_h into the columns of head h of the O tile (fragment layout: row g + 8r, columns 8h + c0, c0 + 1)
    __device__ __forceinline__ void pv(int h, const float (&p)[2][4][4], float (&o)[4][4]) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                if (!(tile_needed(2 * mi, 2 * kk) || tile_needed(2 * mi, 2 * kk + 1) || tile_needed(2 * mi + 1, 2 * kk) || tile_needed(2 * mi + 1, 2 * kk + 1))) continue;
                float c[4] = {o[2 * mi][2 * h], o[2 * mi][2 * h + 1], o[2 * mi + 1][2 * h], o[2 * mi + 1][2 * h + 1]};
                mma16816(c, pk(p[mi][2 * kk][0], p[mi][2 * kk][1]), pk(p[mi][2 * kk][2], p[mi][2 * kk][3]),
                         pk(p[mi][2 * kk + 1][0], p[mi][2 * kk + 1][1]), pk(p[mi][2 * kk + 1][2], p[mi][2 * kk + 1][3]), vt[h][2 * kk], vt[h][2 * kk + 1]);
                o[2 * mi][2 * h] = c[0]; o[2 * mi][2 * h + 1] = c[1]; o[2 * mi + 1][2 * h] = c[2]; o[2 * mi + 1][2 * h + 1] = c[3];
            }
        }
    }
};

// fp32 reference-layout weight [N][ld] (row n, K columns starting at k0) -> bf16 canonical K-major B operand:
// element (n, k) at (k / 8) * N * 16 + n * 16 + (k % 8) * 2 bytes; rows n < nscale are multiplied by `scale`
__device__ __forceinline__ void stage_b_kmajor(unsigned char* dst, const float* W, int ld, int N, int K, int nscale, float scale) {
    for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
        const int n = e / K, k = e % K;
        float w = W[n * ld + k];
        if (n < nscale) w *= scale;
        *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * N * 16 + n * 16 + (k & 7) * 2) = __float2bfloat16(w);
    }
}
// same with the roles of the weight's two indices swapped: B(n, k) = W[k][n]
__device__ __forceinline__ void stage_b_kmajor_t(unsigned char* dst, const float* W, int ld, int N, int K) {
    for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
        const int n = e / K, k = e % K;
        *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * N * 16 + n * 16 + (k & 7) * 2) = __float2bfloat16(W[k * ld + n]);
    }
}

// TMEM column map of the forward kernel (128 columns)
constexpr int F_DQKV = 0, F_DH = 0, F_DP = 64, F_DO = 64, F_AX = 80, F_AH = 96, F_COLS = 128;

struct FwdSm {          // dynamic shared memory of the forward kernel
    unsigned char xbuf[2][NT * 32];                  // bf16 token tiles (TMA bulk destination), double buffered
    unsigned char stg[4][32 * STG];                  // per-warp q | k | v staging
    unsigned char a2[2 * PLANE];                     // attention output as the proj GEMM's A operand
    unsigned char wqkv[48 * 32], wproj[16 * 32], w1[64 * 32], w2[16 * 128];
    float bqkv[48], bproj[16], b1[64], b2[16];
    uint64_t mma_bar;
    uint64_t load_bar[2][4];
    uint32_t tmem_slot;
};

template <int WD, int WH, int WW, bool EMB>
__global__ void __launch_bounds__(NT, 4)
swin_fwd_umma_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ ymid,
                     const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index, Geom g) {
    constexpr int G = WD * WH * WW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSm& S = *reinterpret_cast<FwdSm*>(smem_raw);
    float* Bn = reinterpret_cast<float*>(smem_raw + sizeof(FwdSm));          // [NH][G][G], log2 domain
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "n"(F_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&S.mma_bar), 1);
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int w = 0; w < 4; ++w) mbar_init(smem_u32(&S.load_bar[b][w]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const float qs = g.scale * LOG2E;                                        // q carries scale * log2(e): one ex2 per probability
    stage_b_kmajor(S.wqkv, P + po.qkv_w, C, 48, 16, 16, qs);
    stage_b_kmajor(S.wproj, P + po.proj_w, C, 16, 16, 0, 1.f);
    stage_b_kmajor(S.w1, P + po.fc1_w, C, 64, 16, 0, 1.f);
    stage_b_kmajor(S.w2, P + po.fc2_w, HID, 16, 64, 0, 1.f);
    for (int e = tid; e < 48; e += NT) S.bqkv[e] = P[po.qkv_b + e] * (e < 16 ? qs : 1.f);
    for (int e = tid; e < 16; e += NT) { S.bproj[e] = P[po.proj_b + e]; S.b2[e] = P[po.fc2_b + e]; }
    for (int e = tid; e < 64; e += NT) S.b1[e] = P[po.fc1_b + e];
    stage_bias_n<G>(Bn, P, rel_index);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);             // this warp's lane quadrant
    const uint32_t mma_bar = smem_u32(&S.mma_bar);
    uint32_t n_commit = 0;                                                   // commits so far -> wait parity

    const int n_tiles = (g.n_wg + 3) / 4;
    const uint32_t my_slot = (uint32_t)(tid * 32);
    // token tile prefetch (TMA bulk copy of this thread's 32-byte token row into its slot of the double buffer)
    auto prefetch = [&](int tile, int buf) {
        if (EMB) return;
        const Tok t = map_token<WD, WH, WW>(g, v, tile * 4 + warp, lane);
        const uint32_t bar = smem_u32(&S.load_bar[buf][warp]);
        const unsigned m = __ballot_sync(0xffffffffu, t.valid);
        if (lane == 0) mbar_expect_tx(bar, 32u * __popc(m));
        __syncwarp();
        const uint32_t dst = smem_u32(S.xbuf[buf]) + my_slot;
        if (t.valid) bulk_g2s(dst, x + t.off, 32, bar);
        else { sts128(dst, 0, 0, 0, 0); sts128(dst + 16, 0, 0, 0, 0); }
    };
    if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x, 0);

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const Tok tk = map_token<WD, WH, WW>(g, v, tile * 4 + warp, lane);
        const bool masked = __any_sync(0xffffffffu, tk.border);
        // ---- P0: token row -> LN1 -> A operand (TMEM) -> QKV GEMM ----
        float xr[16];
        if (EMB) {
            const float xin = tk.valid ? __ldg(g.emb_x + (tk.off >> 4)) : 0.f;
            const float* w = g.emb_w + v * C;
            const float* b = g.emb_b + v * C;
            float e[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) e[c] = __ldg(w + c) * xin + __ldg(b + c);
            ln_row(e, xr);
        } else {
            mbar_wait(smem_u32(&S.load_bar[buf][warp]), (uint32_t)(it >> 1) & 1u);
            const uint32_t src = smem_u32(S.xbuf[buf]) + my_slot;
            const uint4 lo = lds128(src), hi = lds128(src + 16);
            const uint32_t pr[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            unpack16(pr, xr);
        }
        if (tile + (int)gridDim.x < n_tiles) prefetch(tile + gridDim.x, buf ^ 1);
        {
            float xn[16];
            ln_row(xr, xn);
            uint32_t pa[8];
            if (tk.valid) pack16(xn, pa);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) pa[i] = 0u;                      // zero padding AFTER LN1 (Swin_3D.py:233-238)
            }
            tmem_st8(tlane + F_AX, pa);
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_ts(tmem + F_DQKV, tmem + F_AX, make_desc(smem_u32(S.wqkv), 48 * 16, 128), idesc(128, 48), 0);
            umma_commit(mma_bar);
        }
        // ---- P1: q | k | v rows (+ bias) -> bf16 staging tile of the warp ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        const uint32_t stg = smem_u32(S.stg[warp]);
        {
            float qkv[16];
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                tmem_ld16(tlane + F_DQKV + 16 * part, qkv);
#pragma unroll
                for (int c = 0; c < 16; ++c) qkv[c] += S.bqkv[16 * part + c];
                uint32_t pq[8];
                pack16(qkv, pq);
                sts128(stg + lane * STG + 32 * part, pq[0], pq[1], pq[2], pq[3]);
                sts128(stg + lane * STG + 32 * part + 16, pq[4], pq[5], pq[6], pq[7]);
            }
        }
        __syncwarp();
        // ---- P2: attention core of the warp -> O as the proj GEMM's A operand (smem, K-major core matrices) ----
        {
            AttnU<G> at;
            at.load(stg, lane, 0, 32, 64);
            float o[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) o[r][0] = o[r][1] = o[r][2] = o[r][3] = 0.f;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float p[2][4][4], rinv[4];
                at.template probs<false>(h, p, rinv, Bn, tk.code, masked, lane);
                at.pv(h, p, o);
#pragma unroll
                for (int r = 0; r < 4; ++r) { o[r][2 * h] *= rinv[r]; o[r][2 * h + 1] *= rinv[r]; }
            }
            const uint32_t a2 = smem_u32(S.a2) + (uint32_t)((warp * 32 + lane / 4) * 16 + (lane % 4) * 4);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int h = 0; h < NH; ++h) sts32(a2 + h * PLANE + r * 8 * 16, pk(o[r][2 * h], o[r][2 * h + 1]));
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_ss(tmem + F_DP, make_desc(smem_u32(S.a2), PLANE, 128), make_desc(smem_u32(S.wproj), 16 * 16, 128), idesc(128, 16), 0);
            umma_commit(mma_bar);
        }
        // ---- P3: y = x + proj(o) + b -> (ymid) -> LN2 -> A operand -> fc1 GEMM ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        float y[16];
        {
            tmem_ld16(tlane + F_DP, y);
#pragma unroll
            for (int c = 0; c < 16; ++c) y[c] += xr[c] + S.bproj[c];
            if (ymid != nullptr && tk.valid) {
                uint32_t py[8];
                pack16(y, py);
                st8u(ymid + tk.off, py);
            }
            float yn[16];
            ln_row(y, yn);
            uint32_t pa[8];
            pack16(yn, pa);
            tmem_st8(tlane + F_AX, pa);
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_ts(tmem + F_DH, tmem + F_AX, make_desc(smem_u32(S.w1), 64 * 16, 128), idesc(128, 64), 0);
            umma_commit(mma_bar);
        }
        // ---- P4: hidden = GELU(fc1 + b1) -> A operand (TMEM, 32 columns) -> fc2 GEMM in four K steps ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float hd[32];
            tmem_ld32(tlane + F_DH + 32 * half, hd);
            uint32_t ph[16];
#pragma unroll
            for (int c = 0; c < 16; ++c)
                ph[c] = pk(gelu_fast(hd[2 * c] + S.b1[32 * half + 2 * c]), gelu_fast(hd[2 * c + 1] + S.b1[32 * half + 2 * c + 1]));
            tmem_st16(tlane + F_AH + 16 * half, ph);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                umma_ts(tmem + F_DO, tmem + F_AH + 8 * s, make_desc(smem_u32(S.w2) + s * 512, 16 * 16, 128), idesc(128, 16), s > 0 ? 1u : 0u);
            umma_commit(mma_bar);
        }
        // ---- P5: out = y + fc2(hidden) + b2 ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        {
            float d[16];
            tmem_ld16(tlane + F_DO, d);
#pragma unroll
            for (int c = 0; c < 16; ++c) d[c] += y[c] + S.b2[c];
            if (tk.valid) {
                uint32_t po_[8];
                pack16(d, po_);
                st8u(out + tk.off, po_);
            }
        }
        tc_fence_before();      // the next tile's first MMA overwrites columns read above: ordered by its __syncthreads
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(F_COLS) : "memory");
}

}  // namespace swu
