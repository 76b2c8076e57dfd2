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
// D = F32, A = B = F16 (the fc2 GEMM of the forward kernel: its A operand is the packed-f16 GELU output)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
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

// GELU (tanh form) on two values packed as f16x2: six packed instructions and ONE MUFU for the pair.  The result feeds the fc2
// GEMM as an f16 operand (11 mantissa bits; the bf16 operand it replaces has 8).  y = 0.5 x (1 + tanh(x (a + b x^2)))
__device__ __forceinline__ uint32_t gelu_f16x2(float x0, float x1) {
    const __half2 x = __floats2half2_rn(x0, x1);
    const __half2 x2 = __hmul2(x, x);
    const __half2 u = __hmul2(x, __hfma2(x2, __float2half2_rn(0.0356774081363001f), __float2half2_rn(0.7978845608028654f)));
    uint32_t ub = *reinterpret_cast<const uint32_t*>(&u), tb;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(ub));
    const __half2 t = *reinterpret_cast<const __half2*>(&tb);
    const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
    const __half2 y = __hfma2(hx, t, hx);
    return *reinterpret_cast<const uint32_t*>(&y);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// LayerNorm(16, eps 1e-5, no affine) of one token held by the thread; returns rstd
// (sums in four interleaved partial accumulators: a 16-long dependent FADD / FFMA chain is 64 cycles of latency that four warps
// per scheduler cannot hide)
__device__ __forceinline__ float sum16(const float* x) {
    float a = x[0] + x[4], b = x[1] + x[5], c = x[2] + x[6], d = x[3] + x[7];
    a += x[8]; b += x[9]; c += x[10]; d += x[11];
    a += x[12]; b += x[13]; c += x[14]; d += x[15];
    return (a + b) + (c + d);
}
__device__ __forceinline__ float dot16(const float* x, const float* y) {
    float a = x[0] * y[0], b = x[1] * y[1], c = x[2] * y[2], d = x[3] * y[3];
#pragma unroll
    for (int i = 4; i < 16; i += 4) { a = fmaf(x[i], y[i], a); b = fmaf(x[i + 1], y[i + 1], b); c = fmaf(x[i + 2], y[i + 2], c); d = fmaf(x[i + 3], y[i + 3], d); }
    return (a + b) + (c + d);
}
__device__ __forceinline__ float ln_row(const float* x, float* xn) {
    const float mu = sum16(x) * (1.f / 16.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) xn[i] = x[i] - mu;
    const float var = dot16(xn, xn);
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
    uint32_t un, ur, udw, uhw, uww;
    g.fd_img.divmod(active ? (uint32_t)win : 0u, un, ur);
    g.fd_hw.divmod(ur, udw, ur);
    g.fd_w.divmod(ur, uhw, uww);
    const int n = (int)un, dw = (int)udw, hw = (int)uhw, ww = (int)uww;
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

// [h][i][j] float tables in shared memory (the staged relative-position bias Bn and the per-warp bias-gradient tables), accessed
// as float2 by lane (g = lane / 4 -> row i = g + 8 r, c0 = 2 (lane % 4) -> columns 8 nj + c0): the four rows a half warp touches
// must land in four different 32-byte bank groups.  G = 32: row stride 32 with the 8-column group index XOR-ed by (row & 3);
// G = 16: padded stride 24; G = 8: stride 8.
__host__ __device__ constexpr int bns(int G) { return G == 32 ? 32 : (G == 16 ? 24 : G); }
template <int G>
__device__ __forceinline__ int tcol(int row, int col) { return G == 32 ? (col ^ ((row & 3) << 3)) : col; }
template <int G>
__device__ __forceinline__ void stage_bias_pad(float* Bn, const float* tbl, const int* __restrict__ rel_index) {
    for (int e = threadIdx.x; e < NH * G * G; e += blockDim.x) {
        const int h = e / (G * G), ij = e % (G * G), i = ij / G, j = ij % G;
        Bn[(h * G + i) * bns(G) + tcol<G>(i, j)] = tbl[rel_index[ij] * NH + h] * LOG2E;      // log2 domain
    }
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
        // the accumulators start from the relative-position bias (log2 domain): s = q k^T + B comes out of the MMA
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) {
                p[mi][nj][0] = p[mi][nj][1] = p[mi][nj][2] = p[mi][nj][3] = 0.f;
                const int jl = (8 * nj + c0) % G;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (tile_needed(2 * mi + hf, nj)) {
                        const float2 b = *reinterpret_cast<const float2*>(Bn + (h * G + (g + 8 * (2 * mi + hf)) % G) * bns(G) + tcol<G>(g, jl));
                        p[mi][nj][2 * hf] = b.x; p[mi][nj][2 * hf + 1] = b.y;
                    }
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
            float mx = -INFINITY;
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
    unsigned char xbuf[2][NT * 64];                  // token tiles (TMA bulk destination; 32 B bf16 or 64 B fp32 rows), double buffered
    unsigned char stg[4][32 * STG];                  // per-warp q | k | v staging
    unsigned char a2[2 * PLANE];                     // attention output as the proj GEMM's A operand
    unsigned char wqkv[48 * 32], wproj[16 * 32], w1[64 * 32], w2[16 * 128];
    // biases ride on the tensor core: one more K step against a constant A operand whose elements 0 and 1 are 1 (chunk 0 = ones0,
    // chunk 1 = zero) and B rows (hi(b[n]), lo(b[n]), 0, ...) -- the bf16 hi + lo split keeps the fp32 bias to 2^-17
    unsigned char ones0[128];                         // one 8-row core matrix, shared by all 16 row groups (SBO = 0)
    unsigned char bias_c[(48 + 16 + 64 + 16) * 16];   // chunk 0 of the bias operands: qkv | proj | fc1 | fc2
    unsigned char zero[128];                          // chunk 1 of every bias-step operand: one zero core matrix (SBO = 0; after them: LBO > 0)
    uint64_t mma_bar;
    uint64_t load_bar[2][4];
    uint32_t tmem_slot;
};

template <int WD, int WH, int WW, bool EMB>
__global__ void __launch_bounds__(NT, 4)
swin_fwd_umma_kernel(const void* __restrict__ x, void* __restrict__ out, __nv_bfloat16* __restrict__ ymid,
                     const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index, Geom g) {
    constexpr int G = WD * WH * WW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSm& S = *reinterpret_cast<FwdSm*>(smem_raw);
    float* Bn = reinterpret_cast<float*>(smem_raw + sizeof(FwdSm));          // [NH][G][bns(G)], log2 domain
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
    for (int e = tid; e < 16 * 64; e += NT) {                                // fc2.weight as an f16 operand (see gelu_f16x2)
        const int n = e / 64, k = e % 64;
        *reinterpret_cast<__half*>(S.w2 + (k >> 3) * 16 * 16 + n * 16 + (k & 7) * 2) = __float2half_rn(P[po.fc2_w + n * HID + k]);
    }
    for (int e = tid; e < 144 * 8; e += NT) {
        const int n = e >> 3, k = e & 7;
        const float b = n < 48 ? P[po.qkv_b + n] * (n < 16 ? qs : 1.f) : (n < 64 ? P[po.proj_b + n - 48] : (n < 128 ? P[po.fc1_b + n - 64] : P[po.fc2_b + n - 128]));
        const __nv_bfloat16 hi = __float2bfloat16(b);
        const __nv_bfloat16 lo = __float2bfloat16(b - __bfloat162float(hi));
        reinterpret_cast<__nv_bfloat16*>(S.bias_c)[e] = k == 0 ? hi : (k == 1 ? lo : __float2bfloat16(0.f));
    }
    if (tid < 8) { sts128(smem_u32(S.ones0) + tid * 16, 0x3F803F80u, 0u, 0u, 0u); sts128(smem_u32(S.zero) + tid * 16, 0u, 0u, 0u, 0u); }
    stage_bias_pad<G>(Bn, P, rel_index);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);             // this warp's lane quadrant
    const uint32_t mma_bar = smem_u32(&S.mma_bar);
    uint32_t n_commit = 0;                                                   // commits so far -> wait parity
    // D[128 x N] += 1 * bias^T: A = (ones0, zero), B = (bias_c rows n0 .. n0 + N, zero)
    auto bias_step = [&](uint32_t dcol, int n0, int N) {
        const uint32_t a0 = smem_u32(S.ones0), b0 = smem_u32(S.bias_c) + n0 * 16, z = smem_u32(S.zero);
        // K-major operands as (chunk 0, chunk 1) pairs: A = (ones0, zero) with SBO = 0 (every 8-row group reads the same core
        // matrix); B chunk 0 = the bias rows (SBO = 128), chunk 1 = zero -- the B descriptor's SBO also applies to chunk 1, so
        // the zero chunk is read at z + 128 * group: `zero` is followed by >= 1 KB of bytes whose value times A's zero chunk
        // cannot matter only if they are finite; therefore B's chunk 1 points at the bias rows themselves (finite) and it is A's
        // chunk 1 (exact zeros) that cancels the second half of the K step
        umma_ss(tmem + dcol, make_desc(a0, z - a0, 0), make_desc(b0, 0, 128), idesc(128, N), 1);
    };

    const int n_tiles = (g.n_wg + 3) / 4;
    const uint32_t my_slot = (uint32_t)(tid * 64);
    const uint32_t row_bytes = g.x32 ? 64u : 32u;
    // token tile prefetch: TMA bulk copy of this thread's token row into its slot of the double buffer (fused embedding: the
    // raw scalar into a register).  The token map of a tile is computed once, here, and carried into the tile's iteration.
    float xin_next = 0.f;
    auto prefetch = [&](const Tok& t, int buf) {
        if (EMB) { xin_next = t.valid ? __ldg(g.emb_x + (t.off >> 4)) : 0.f; return; }
        const uint32_t bar = smem_u32(&S.load_bar[buf][warp]);
        const unsigned m = __ballot_sync(0xffffffffu, t.valid);
        if (lane == 0) mbar_expect_tx(bar, row_bytes * __popc(m));
        __syncwarp();
        const uint32_t dst = smem_u32(S.xbuf[buf]) + my_slot;
        if (t.valid) bulk_g2s(dst, reinterpret_cast<const unsigned char*>(x) + t.off * (g.x32 ? 4 : 2), row_bytes, bar);
        else { sts128(dst, 0, 0, 0, 0); sts128(dst + 16, 0, 0, 0, 0); sts128(dst + 32, 0, 0, 0, 0); sts128(dst + 48, 0, 0, 0, 0); }
    };
    Tok tk_next = map_token<WD, WH, WW>(g, v, (int)blockIdx.x * 4 + warp, lane);
    if ((int)blockIdx.x < n_tiles) prefetch(tk_next, 0);

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const Tok tk = tk_next;
        const float xin = xin_next;
        const bool masked = __any_sync(0xffffffffu, tk.border);
        // ---- P0: token row -> LN1 -> A operand (TMEM) -> QKV GEMM ----
        float xr[16];
        if (EMB) {
            const float* w = g.emb_w + v * C;
            const float* b = g.emb_b + v * C;
            float e[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) e[c] = __ldg(w + c) * xin + __ldg(b + c);
            ln_row(e, xr);
        } else {
            mbar_wait(smem_u32(&S.load_bar[buf][warp]), (uint32_t)(it >> 1) & 1u);
            const uint32_t src = smem_u32(S.xbuf[buf]) + my_slot;
            if (g.x32) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 q4 = lds128(src + 16 * i);
                    xr[4 * i] = __uint_as_float(q4.x); xr[4 * i + 1] = __uint_as_float(q4.y); xr[4 * i + 2] = __uint_as_float(q4.z); xr[4 * i + 3] = __uint_as_float(q4.w);
                }
            } else {
                const uint4 lo = lds128(src), hi = lds128(src + 16);
                const uint32_t pr[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                unpack16(pr, xr);
            }
        }
        if (tile + (int)gridDim.x < n_tiles) {
            tk_next = map_token<WD, WH, WW>(g, v, (tile + (int)gridDim.x) * 4 + warp, lane);
            prefetch(tk_next, buf ^ 1);
        }
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
            // the residual rides in the accumulator: x goes into the columns the proj GEMM accumulates onto (y = x + proj(o) + b), and
            // the fc2 GEMM accumulates onto y in the same columns (out = y + fc2(h) + b)
            uint32_t px[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) px[c] = __float_as_uint(xr[c]);
            tmem_st16(tlane + F_DP, px);
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_ts(tmem + F_DQKV, tmem + F_AX, make_desc(smem_u32(S.wqkv), 48 * 16, 128), idesc(128, 48), 0);
            bias_step(F_DQKV, 0, 48);
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
            umma_ss(tmem + F_DP, make_desc(smem_u32(S.a2), PLANE, 128), make_desc(smem_u32(S.wproj), 16 * 16, 128), idesc(128, 16), 1);
            bias_step(F_DP, 48, 16);
            umma_commit(mma_bar);
        }
        // ---- P3: y = x + proj(o) + b -> (ymid) -> LN2 -> A operand -> fc1 GEMM ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        {
            float y[16];
            tmem_ld16(tlane + F_DP, y);
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
            bias_step(F_DH, 64, 64);
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
                ph[c] = gelu_f16x2(hd[2 * c], hd[2 * c + 1]);
            tmem_st16(tlane + F_AH + 16 * half, ph);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                umma_ts(tmem + F_DO, tmem + F_AH + 8 * s, make_desc(smem_u32(S.w2) + s * 512, 16 * 16, 128), idesc_f16(128, 16), 1);
            bias_step(F_DO, 128, 16);
            umma_commit(mma_bar);
        }
        // ---- P5: out = y + fc2(hidden) + b2 ----
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        {
            float d[16];
            tmem_ld16(tlane + F_DO, d);
            if (tk.valid) {
                if (g.out32) { float* o32 = reinterpret_cast<float*>(out) + tk.off; st8f(o32, d); st8f(o32 + 8, d + 8); }
                else {
                    uint32_t po_[8];
                    pack16(d, po_);
                    st8u(reinterpret_cast<__nv_bfloat16*>(out) + tk.off, po_);
                }
            }
        }
        tc_fence_before();      // the next tile's first MMA overwrites columns read above: ordered by its __syncthreads
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(F_COLS) : "memory");
}


// =====================================================================================================
// backward, MLP half (tcgen05):  out = y + fc2(gelu(fc1(LN(y))))
//   inputs y, g_out (bf16 tokens);  outputs g_y (bf16), per-CTA partial d{fc1_w, fc1_b, fc2_w, fc2_b} (MLP_PART layout)
// =====================================================================================================
// GELU (tanh form, the forward's gelu_fast) and ITS derivative: y = 0.5 x (1 + t), t = tanh(u), u = x (a + b x^2);
// dy/dx = 0.5 (1 + t) + 0.5 x (1 - t^2) (a + 3 b x^2)
#ifndef IDEE_GELU_GRAD_PACKED
#define IDEE_GELU_GRAD_PACKED 1
#endif
__device__ __forceinline__ void gelu_tanh_grad(float x, float& y, float& dy) {
    const float x2 = x * x;
    const float t = tanh_approx(x * (0.7978845608028654f + 0.0356774081363001f * x2));
    const float cdf = 0.5f + 0.5f * t;
    y = x * cdf;
    dy = (0.5f * x) * (1.f - t * t) * (0.7978845608028654f + 0.1070322244089003f * x2) + cdf;
}
// The same pair of functions on two hidden units packed as bf16x2 (HFMA2.BF16 + one MUFU.TANH per pair): both results are
// MMA operands in bf16 anyway (gelu(h) for dW2, g_pre = g_h * dy for dW1 and g_yn), so the packed form costs about twice the
// rounding error of the fp32 form followed by one rounding (3.8e-3 against 2.0e-3 relative L2 on N(0, 1.5) inputs) and half
// the instructions.  g_h keeps its fp32 exponent range in bf16.  Returns g_pre; h receives gelu(x).
__device__ __forceinline__ uint32_t gelu_tanh_grad_bf16x2(float x0, float x1, float g0, float g1, uint32_t& h) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(x0, x1);
    const __nv_bfloat162 ka = __float2bfloat162_rn(0.7978845608028654f), half = __float2bfloat162_rn(0.5f);
    const __nv_bfloat162 x2 = __hmul2(x, x);
    const __nv_bfloat162 u = __hmul2(x, __hfma2(x2, __float2bfloat162_rn(0.0356774081363001f), ka));
    uint32_t ub = *reinterpret_cast<const uint32_t*>(&u), tb;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(tb) : "r"(ub));
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&tb);
    const __nv_bfloat162 cdf = __hfma2(t, half, half);
    const __nv_bfloat162 y = __hmul2(x, cdf);
    const __nv_bfloat162 q = __hfma2(x2, __float2bfloat162_rn(0.1070322244089003f), ka);
    const __nv_bfloat162 omt = __hfma2(__hneg2(t), t, __float2bfloat162_rn(1.f));
    const __nv_bfloat162 w = __hmul2(__hmul2(x, half), omt);
    const __nv_bfloat162 dy = __hfma2(w, q, cdf);
    const __nv_bfloat162 gp = __hmul2(__floats2bfloat162_rn(g0, g1), dy);
    h = *reinterpret_cast<const uint32_t*>(&y);
    return *reinterpret_cast<const uint32_t*>(&gp);
}

// TMEM column map of the MLP backward kernel (128 columns): pre-activation and hidden gradient of one 32-unit half, the token
// gradient, and the persistent weight-gradient accumulator D[128 x 48] = [g_pre | gelu(h)]^T [yn | g_out | 1 | 0]
constexpr int M_PRE = 0, M_DH = 32, M_DYN = 64, M_WACC = 80, M_COLS = 128;

struct MlpBwdSm {
    unsigned char pa[16 * PLANE];        // chunk planes: g_pre (8 chunks of 8 hidden units) | gelu(h) (8 chunks)
    unsigned char ones0[PLANE];          // A operand of the bias GEMM step, chunk 0: elements 0 and 1 of every row = 1 (chunk 1 = pb[5])
    unsigned char b1[64 * 16];           // B(n = hidden, k) chunk 0 of the bias step: k = 0 -> hi(b1[n]), k = 1 -> lo(b1[n]) (chunk 1 = pb[5])
    unsigned char w1[64 * 32];           // B(n = hidden, k = c) = fc1.weight[n][k]            (pre = yn W1^T)
    unsigned char w2t[64 * 32];          // B(n = hidden, k = c) = fc2.weight[k][n]            (g_h = g_out W2)
    unsigned char w1t[16 * 128];         // B(n = c, k = hidden) = fc1.weight[k][n]            (g_yn = g_pre W1)
    unsigned char pb[6 * PLANE];         // chunk planes: yn (2) | g_out (2) | ones (1) | zeros (1); after ones0 / b1: their LBO > 0
    float red[4][16];
    uint64_t mma_bar;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(NT, 4)
swin_mlp_bwd_umma_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ gout, __nv_bfloat16* __restrict__ gy,
                         const float* __restrict__ params, int64_t pstride, int tbl, float* __restrict__ partials,
                         int N, int V, int thw, FastDiv fd_thw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MlpBwdSm& S = *reinterpret_cast<MlpBwdSm*>(smem_raw);
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(tbl);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "n"(M_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { mbar_init(smem_u32(&S.mma_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    stage_b_kmajor(S.w1, P + po.fc1_w, C, 64, 16, 0, 1.f);
    stage_b_kmajor_t(S.w2t, P + po.fc2_w, HID, 64, 16);
    stage_b_kmajor_t(S.w1t, P + po.fc1_w, C, 16, 64);
    for (int e = tid; e < 64 * 8; e += NT) {                       // bias rows: hi + lo bf16 split keeps the fp32 bias to 2^-17
        const int n = e >> 3, k = e & 7;
        const float b = P[po.fc1_b + n];
        const __nv_bfloat16 hi = __float2bfloat16(b);
        const __nv_bfloat16 lo = __float2bfloat16(b - __bfloat162float(hi));
        reinterpret_cast<__nv_bfloat16*>(S.b1)[e] = k == 0 ? hi : (k == 1 ? lo : __float2bfloat16(0.f));
    }
    // constant planes: ones (B operand column block of the bias-gradient sums), zeros, and the A operand of the bias step
    sts128(smem_u32(S.pb) + 4 * PLANE + tid * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    sts128(smem_u32(S.pb) + 5 * PLANE + tid * 16, 0u, 0u, 0u, 0u);
    sts128(smem_u32(S.ones0) + tid * 16, 0x3F803F80u, 0u, 0u, 0u);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t mma_bar = smem_u32(&S.mma_bar);
    const uint32_t pa = smem_u32(S.pa), pb = smem_u32(S.pb);
    uint32_t n_commit = 0;
    float db2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) db2[c] = 0.f;

    const int64_t ntok = (int64_t)N * thw;
    const int n_tiles = (int)((ntok + NT - 1) / NT);
    bool first = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tok = (uint32_t)tile * NT + tid;
        const bool valid = tok < ntok;
        uint32_t n, rem;
        fd_thw.divmod(valid ? tok : 0u, n, rem);
        const int64_t off = ((int64_t)(n * V + v) * thw + rem) * C;
        {   // next tile's rows towards L2 while this tile computes
            const uint32_t tok2 = tok + gridDim.x * NT;
            if (tok2 < ntok) {
                uint32_t n2, rem2;
                fd_thw.divmod(tok2, n2, rem2);
                const int64_t off2 = ((int64_t)(n2 * V + v) * thw + rem2) * C;
                prefetch_l2(y + off2); prefetch_l2(gout + off2);
            }
        }
        // ---- R0: rows -> LN -> planes yn | g_out -> GEMMs of half 0 ----
        uint32_t pg[8];
        float yn[16], rstd;
        {
            uint32_t py[8];
            if (valid) { ldg_row16(y + off, py); ldg_row16(gout + off, pg); }
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { py[i] = 0u; pg[i] = 0u; }
            }
            float yr[16];
            unpack16(py, yr);
            rstd = ln_row(yr, yn);
            uint32_t pn[8];
            pack16(yn, pn);
            sts128(pb + tid * 16, pn[0], pn[1], pn[2], pn[3]);
            sts128(pb + PLANE + tid * 16, pn[4], pn[5], pn[6], pn[7]);
            sts128(pb + 2 * PLANE + tid * 16, pg[0], pg[1], pg[2], pg[3]);
            sts128(pb + 3 * PLANE + tid * 16, pg[4], pg[5], pg[6], pg[7]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { db2[2 * i] += bf_lo(pg[i]); db2[2 * i + 1] += bf_hi(pg[i]); }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            proxy_fence();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                // pre = yn W1^T + b1 (bias as a second K step against the constant ones operand);  g_h = g_out W2
                umma_ss(tmem + M_PRE, make_desc(pb, PLANE, 128), make_desc(smem_u32(S.w1) + half * 512, 64 * 16, 128), idesc(128, 32), 0);
                umma_ss(tmem + M_PRE, make_desc(smem_u32(S.ones0), pb + 5 * PLANE - smem_u32(S.ones0), 128),
                        make_desc(smem_u32(S.b1) + half * 512, pb + 5 * PLANE - (smem_u32(S.b1) + half * 512), 128), idesc(128, 32), 1);
                umma_ss(tmem + M_DH, make_desc(pb + 2 * PLANE, PLANE, 128), make_desc(smem_u32(S.w2t) + half * 512, 64 * 16, 128), idesc(128, 32), 0);
                umma_commit(mma_bar);
            }
            mbar_wait(mma_bar, n_commit++ & 1u);
            tc_fence_after();
            // ---- R1 / R2: GELU and its derivative on this half's 32 hidden units -> planes g_pre | gelu(h) ----
            float pre[32], dh[32];
            tmem_ld32(tlane + M_PRE, pre);
            tmem_ld32(tlane + M_DH, dh);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t pp[4], ph[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#if IDEE_GELU_GRAD_PACKED
                    pp[j] = gelu_tanh_grad_bf16x2(pre[8 * c4 + 2 * j], pre[8 * c4 + 2 * j + 1], dh[8 * c4 + 2 * j], dh[8 * c4 + 2 * j + 1], ph[j]);
#else
                    float h0, d0, h1, d1;
                    gelu_tanh_grad(pre[8 * c4 + 2 * j], h0, d0);
                    gelu_tanh_grad(pre[8 * c4 + 2 * j + 1], h1, d1);
                    pp[j] = pk(dh[8 * c4 + 2 * j] * d0, dh[8 * c4 + 2 * j + 1] * d1);
                    ph[j] = pk(h0, h1);
#endif
                }
                sts128(pa + (4 * half + c4) * PLANE + tid * 16, pp[0], pp[1], pp[2], pp[3]);
                sts128(pa + (8 + 4 * half + c4) * PLANE + tid * 16, ph[0], ph[1], ph[2], ph[3]);
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // weight gradients: D[128 x 48] += [g_pre | gelu(h)]^T [yn | g_out | 1 | 0], K = the tile's 128 tokens in 8 steps;
            // both operands are MN-major views of the token-major chunk planes (LBO = 8-token group, SBO = chunk plane)
#pragma unroll
            for (int s = 0; s < 8; ++s)
                umma_ss(tmem + M_WACC, make_desc(pa + s * 256, 128, PLANE), make_desc(pb + s * 256, 128, PLANE), idesc(128, 48, 1, 1),
                        (first && s == 0) ? 0u : 1u);
            // g_yn = g_pre W1  (K = 64 hidden units in four steps)
#pragma unroll
            for (int s = 0; s < 4; ++s)
                umma_ss(tmem + M_DYN, make_desc(pa + 2 * s * PLANE, PLANE, 128), make_desc(smem_u32(S.w1t) + s * 512, 16 * 16, 128), idesc(128, 16), s > 0 ? 1u : 0u);
            umma_commit(mma_bar);
        }
        first = false;
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        // ---- R3: LayerNorm backward + residual -> g_y ----
        {
            float d[16];
            tmem_ld16(tlane + M_DYN, d);
            const float m1 = sum16(d) * (1.f / 16.f), m2 = dot16(d, yn) * (1.f / 16.f);
            float go[16];
            unpack16(pg, go);
#pragma unroll
            for (int c = 0; c < 16; ++c) d[c] = go[c] + rstd * (d[c] - m1 - yn[c] * m2);
            if (valid) {
                uint32_t po_[8];
                pack16(d, po_);
                st8u(gy + off, po_);
            }
        }
        tc_fence_before();
    }
    // ---- per-CTA partials: fc1_w[k][c] | fc1_b[k] | fc2_w[c][k] | fc2_b[c] ----
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        float s = db2[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) S.red[warp][c] = s;          // fixed summation order: bit-identical from run to run
    }
    __syncthreads();
    tc_fence_after();
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * MLP_PART;
    const bool any = (int)blockIdx.x < n_tiles;                     // a CTA without tiles never wrote its accumulator
    if (tid < 64) {                                                 // rows 0..63 of D: g_pre units
        float w[16], b[16];
        tmem_ld16(tlane + M_WACC, w);
        tmem_ld16(tlane + M_WACC + 32, b);
#pragma unroll
        for (int c = 0; c < 16; ++c) part[tid * C + c] = any ? w[c] : 0.f;
        part[HID * C + tid] = any ? b[0] : 0.f;
    } else {                                                        // rows 64..127: gelu(h) units against g_out
        float w[16];
        tmem_ld16(tlane + M_WACC + 16, w);
        const int k = tid - 64;
#pragma unroll
        for (int c = 0; c < 16; ++c) part[HID * C + HID + c * HID + k] = any ? w[c] : 0.f;
    }
    if (tid < 16) part[HID * C + HID + C * HID + tid] = (S.red[0][tid] + S.red[1][tid]) + (S.red[2][tid] + S.red[3][tid]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(M_COLS) : "memory");
}


// =====================================================================================================
// backward, attention half (tcgen05):  y = x + proj(attn(LN(x)))
//   inputs x (or the raw embedding input), g_y (bf16);  outputs g_x (bf16, may alias g_y), per-CTA partial
//   d{qkv_w, qkv_b, proj_w, proj_b, bias[h][i][j]} (ATT_PART_W + NH*G*G layout) and, with the fused embedding, d{embed w, b}
// =====================================================================================================
// chunk planes [plane][token][16 B] of a tile (see the file header):
//   0-1 q -> dq | 2-3 k -> dk | 4-5 v -> dv | 6-7 g_y | 8-9 xn | 10-11 dO -> o | 12 ones | 13 zeros
// weight gradients: D[64 x 48] += [dq | dk | dv | g_y]^T [xn | o | 1 | 0]
// EMB: the block's input tokens are recomputed from the raw input (fused patch embedding, idee_swin_desc.embed_x); the gradient
// w.r.t. those tokens is written to gx like for any other block and embed_bwd_tokens_kernel turns it into d{embed w, b} (the
// LayerNorm backward of the embedding inside this kernel cost 49 registers, one CTA per SM and an extra MMA round per tile).
constexpr int A_Q = 0, A_K = 2, A_V = 4, A_GY = 6;
struct APl {
    static constexpr int XN = 8, DO = XN + 2, ONE = XN + 4, X1 = XN + 5, PLANES = XN + 6;
};
constexpr int B_QKV = 0, B_DO = 48, B_DXN = 64, B_WACC = 80, B_COLS = 128;       // TMEM columns

struct AttnBwdSm {
    unsigned char pl[APl::PLANES * PLANE];
    unsigned char ones0[128];            // A operand of the bias GEMM step: one core matrix (1, 1, 0, ...) for every row group (SBO = 0)
    unsigned char bq[48 * 16];           // B(n, k) chunk 0 of the bias step: k = 0 -> hi(bqkv[n]), k = 1 -> lo (q rows pre-scaled)
    unsigned char wqkv[48 * 32];         // B(n = o, k = c) = qkv.weight[n][k] (q rows pre-scaled)       qkv = xn Wqkv^T
    unsigned char wpt[16 * 32];          // B(n = e, k = c) = proj.weight[k][n]                          dO = g_y Wproj
    unsigned char wqt[16 * 96];          // B(n = c, k = o) = qkv.weight[k][n]                           dxn = [dq|dk|dv] Wqkv
    unsigned char zq[128];               // zeros: chunk 1 of the bias step's A operand (after ones0: LBO > 0)
    uint64_t mma_bar;
    uint32_t tmem_slot;
};

template <int WD, int WH, int WW, bool EMB>
__global__ void __launch_bounds__(NT, 3)
swin_attn_bwd_umma_kernel(const void* __restrict__ x, const __nv_bfloat16* __restrict__ gy, __nv_bfloat16* __restrict__ gx,
                          const float* __restrict__ params, int64_t pstride, const int* __restrict__ rel_index,
                          float* __restrict__ partials, Geom g) {
    constexpr int G = WD * WH * WW;
    constexpr int PART = ATT_PART_W + NH * G * G;
    constexpr int DBS = bns(G), DBW = NH * G * DBS;            // per-warp bias-gradient table (conflict-free 8-byte RMW, see bns / tcol)
    constexpr int WM = 64;                                     // rows of the weight-gradient GEMM
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int A_XN = APl::XN, A_DO = APl::DO, A_ONE = APl::ONE, A_X1 = APl::X1;
    AttnBwdSm& S = *reinterpret_cast<AttnBwdSm*>(smem_raw);
    float* Bn = reinterpret_cast<float*>(smem_raw + sizeof(AttnBwdSm));  // [NH][G][bns(G)]
    float* dBw_all = Bn + NH * G * bns(G);                                  // [4 warps][DBW]
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = params + (int64_t)v * pstride;
    const POff po(g.tbl);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_slot)), "n"(B_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { mbar_init(smem_u32(&S.mma_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const float qs = g.scale * LOG2E;
    stage_b_kmajor(S.wqkv, P + po.qkv_w, C, 48, 16, 16, qs);
    stage_b_kmajor_t(S.wpt, P + po.proj_w, C, 16, 16);
    stage_b_kmajor_t(S.wqt, P + po.qkv_w, C, 16, 48);
    for (int e = tid; e < 48 * 8; e += NT) {
        const int n = e >> 3, k = e & 7;
        const float b = P[po.qkv_b + n] * (n < 16 ? qs : 1.f);
        const __nv_bfloat16 hi = __float2bfloat16(b);
        const __nv_bfloat16 lo = __float2bfloat16(b - __bfloat162float(hi));
        reinterpret_cast<__nv_bfloat16*>(S.bq)[e] = k == 0 ? hi : (k == 1 ? lo : __float2bfloat16(0.f));
    }
    const uint32_t pl = smem_u32(S.pl);
    sts128(pl + A_ONE * PLANE + tid * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    sts128(pl + A_X1 * PLANE + tid * 16, 0u, 0u, 0u, 0u);
    if (tid < 8) { sts128(smem_u32(S.ones0) + tid * 16, 0x3F803F80u, 0u, 0u, 0u); sts128(smem_u32(S.zq) + tid * 16, 0u, 0u, 0u, 0u); }
    for (int e = tid; e < 4 * DBW; e += NT) dBw_all[e] = 0.f;
    stage_bias_pad<G>(Bn, P, rel_index);
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t mma_bar = smem_u32(&S.mma_bar);
    uint32_t n_commit = 0;
    float* dB = dBw_all + warp * DBW;
    const int gq = lane / 4, c0 = 2 * (lane % 4);
    const uint32_t my_row = (uint32_t)(tid * 16);                                            // this token's 16 bytes inside a plane
    // fragment-layout store address (row g, column pair c0 of chunk h): + h * PLANE + r * 128
    const uint32_t fr = (uint32_t)((warp * 32 + gq) * 16 + (lane % 4) * 4);
    const float* ew = g.emb_w + v * C;
    const float* eb = g.emb_b + v * C;

    const int n_tiles = (g.n_wg + 3) / 4;
    bool first = true;
    Tok tk_next = map_token<WD, WH, WW>(g, v, (int)blockIdx.x * 4 + warp, lane);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const Tok tk = tk_next;
        const bool masked = __any_sync(0xffffffffu, tk.border);
        if (tile + (int)gridDim.x < n_tiles) {                 // next tile's token map, and its rows towards L2 while this tile computes
            tk_next = map_token<WD, WH, WW>(g, v, (tile + (int)gridDim.x) * 4 + warp, lane);
            if (tk_next.valid) {
                prefetch_l2(gy + tk_next.off);
                if (EMB) prefetch_l2(g.emb_x + (tk_next.off >> 4));
                else prefetch_l2(reinterpret_cast<const unsigned char*>(x) + tk_next.off * (g.x32 ? 4 : 2));
            }
        }
        // ---- R0: rows -> LN1 -> planes xn | g_y -> qkv and dO GEMMs ----
        float rstd, xin = 0.f;
        {
            float xr[16];
            if (EMB) {
                xin = tk.valid ? __ldg(g.emb_x + (tk.off >> 4)) : 0.f;
                float e[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) e[c] = __ldg(ew + c) * xin + __ldg(eb + c);
                ln_row(e, xr);
            } else if (g.x32) {
                if (tk.valid) { const float* xp = reinterpret_cast<const float*>(x) + tk.off; ldg8f(xr, xp); ldg8f(xr + 8, xp + 8); }
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) xr[i] = 0.f;
                }
            } else {
                uint32_t px[8];
                if (tk.valid) ldg_row16(reinterpret_cast<const __nv_bfloat16*>(x) + tk.off, px);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) px[i] = 0u;
                }
                unpack16(px, xr);
            }
            uint32_t pg[8];
            if (tk.valid) ldg_row16(gy + tk.off, pg);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) pg[i] = 0u;
            }
            float xn[16];
            rstd = ln_row(xr, xn);
            uint32_t pn[8];
            if (tk.valid) pack16(xn, pn);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) pn[i] = 0u;
            }
            sts128(pl + A_XN * PLANE + my_row, pn[0], pn[1], pn[2], pn[3]);
            sts128(pl + (A_XN + 1) * PLANE + my_row, pn[4], pn[5], pn[6], pn[7]);
            sts128(pl + A_GY * PLANE + my_row, pg[0], pg[1], pg[2], pg[3]);
            sts128(pl + (A_GY + 1) * PLANE + my_row, pg[4], pg[5], pg[6], pg[7]);
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_ss(tmem + B_QKV, make_desc(pl + A_XN * PLANE, PLANE, 128), make_desc(smem_u32(S.wqkv), 48 * 16, 128), idesc(128, 48), 0);
            // bias step: A = (ones0, zq) with SBO = 0; B chunk 1 = chunk 0 (finite values against A's exact zeros)
            umma_ss(tmem + B_QKV, make_desc(smem_u32(S.ones0), smem_u32(S.zq) - smem_u32(S.ones0), 0), make_desc(smem_u32(S.bq), 0, 128), idesc(128, 48), 1);
            umma_ss(tmem + B_DO, make_desc(pl + A_GY * PLANE, PLANE, 128), make_desc(smem_u32(S.wpt), 16 * 16, 128), idesc(128, 16), 0);
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        // ---- R1: q | k | v | dO rows -> bf16 planes (ldmatrix sources of the warp's attention core) ----
        {
            float t16[16];
            uint32_t pq[8];
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                tmem_ld16(tlane + B_QKV + 16 * part, t16);
                pack16(t16, pq);
                const uint32_t dst = pl + (part < 3 ? 2 * part : A_DO) * PLANE + my_row;
                sts128(dst, pq[0], pq[1], pq[2], pq[3]);
                sts128(dst + PLANE, pq[4], pq[5], pq[6], pq[7]);
            }
        }
        __syncwarp();
        // ---- attention forward recompute + backward, one head at a time; results go back into the planes ----
        uint32_t st_dq[2][4], st_dk[2][4], st_dv[2][4], st_o[2][4];          // packed bf16 pairs [h][r] awaiting the plane stores
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            // fragments of head h (chunk h of each plane pair): plain and transposed 8x8 blocks of q, k, v, dO (row groups 0..3)
            uint32_t qa[4], qt[4], kb[4], kt[4], vb[4], vt[4], da[4], dt[4];
            {
                const uint32_t a0 = pl + (uint32_t)(h * PLANE + (warp * 32 + (lane & 7) + 8 * (lane >> 3)) * 16);   // matrices = row groups 0..3
                ldsm_x4(a0 + A_Q * PLANE, qa[0], qa[1], qa[2], qa[3]);
                ldsm_x4_t(a0 + A_Q * PLANE, qt[0], qt[1], qt[2], qt[3]);
                ldsm_x4(a0 + A_K * PLANE, kb[0], kb[1], kb[2], kb[3]);
                ldsm_x4_t(a0 + A_K * PLANE, kt[0], kt[1], kt[2], kt[3]);
                ldsm_x4(a0 + A_V * PLANE, vb[0], vb[1], vb[2], vb[3]);
                ldsm_x4_t(a0 + A_V * PLANE, vt[0], vt[1], vt[2], vt[3]);
                ldsm_x4(a0 + A_DO * PLANE, da[0], da[1], da[2], da[3]);
                ldsm_x4_t(a0 + A_DO * PLANE, dt[0], dt[1], dt[2], dt[3]);
            }
            // scores (log2 domain) + bias + mask + softmax (normalised)
            float p[2][4][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    p[mi][nj][0] = p[mi][nj][1] = p[mi][nj][2] = p[mi][nj][3] = 0.f;
                    const int jl = (8 * nj + c0) % G;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
                        if (AttnU<G>::tile_needed(2 * mi + hf, nj)) {
                            const float2 b = *reinterpret_cast<const float2*>(Bn + (h * G + (gq + 8 * (2 * mi + hf)) % G) * bns(G) + tcol<G>(gq, jl));
                            p[mi][nj][2 * hf] = b.x; p[mi][nj][2 * hf + 1] = b.y;
                        }
                    if (AttnU<G>::tile_needed(2 * mi, nj) || AttnU<G>::tile_needed(2 * mi + 1, nj)) mma1688(p[mi][nj], qa[2 * mi], qa[2 * mi + 1], kb[nj]);
                }
            int cj[4][2], cr[4];
            if (masked) {
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    cj[nj][0] = __shfl_sync(0xffffffffu, tk.code, 8 * nj + c0);
                    cj[nj][1] = __shfl_sync(0xffffffffu, tk.code, 8 * nj + c0 + 1);
                    cr[nj] = __shfl_sync(0xffffffffu, tk.code, gq + 8 * nj);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int mi = r / 2, hf = r % 2;
                float mx = -INFINITY;
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    if (AttnU<G>::tile_needed(r, nj)) {
                        if (masked) {
                            if (cj[nj][0] != cr[r]) p[mi][nj][2 * hf] += -100.0f * LOG2E;
                            if (cj[nj][1] != cr[r]) p[mi][nj][2 * hf + 1] += -100.0f * LOG2E;
                        }
                        mx = fmaxf(mx, fmaxf(p[mi][nj][2 * hf], p[mi][nj][2 * hf + 1]));
                    }
                }
                mx = quad_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    if (AttnU<G>::tile_needed(r, nj)) {
                        const float e0 = ex2_approx(p[mi][nj][2 * hf] - mx), e1 = ex2_approx(p[mi][nj][2 * hf + 1] - mx);
                        p[mi][nj][2 * hf] = e0; p[mi][nj][2 * hf + 1] = e1; sum += e0 + e1;
                    } else { p[mi][nj][2 * hf] = 0.f; p[mi][nj][2 * hf + 1] = 0.f; }
                }
                const float inv = __fdividef(1.f, quad_sum(sum));
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) { p[mi][nj][2 * hf] *= inv; p[mi][nj][2 * hf + 1] *= inv; }
            }
            // packed P blocks pb[ib][jb] (row group ib, key group jb); o_h = P V_h
            uint32_t pb[4][4];
#pragma unroll
            for (int ib = 0; ib < 4; ++ib)
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) pb[ib][jb] = pk(p[ib / 2][jb][2 * (ib % 2)], p[ib / 2][jb][2 * (ib % 2) + 1]);
            float oh[2][4];                                                   // [mi][c-fragment]: rows g + 8(2mi), g + 8(2mi+1)
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                oh[mi][0] = oh[mi][1] = oh[mi][2] = oh[mi][3] = 0.f;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    if (!(AttnU<G>::tile_needed(2 * mi, 2 * kk) || AttnU<G>::tile_needed(2 * mi, 2 * kk + 1) || AttnU<G>::tile_needed(2 * mi + 1, 2 * kk) ||
                          AttnU<G>::tile_needed(2 * mi + 1, 2 * kk + 1))) continue;
                    mma16816(oh[mi], pb[2 * mi][2 * kk], pb[2 * mi + 1][2 * kk], pb[2 * mi][2 * kk + 1], pb[2 * mi + 1][2 * kk + 1], vt[2 * kk], vt[2 * kk + 1]);
                }
            }
            // D_r = <dO_r, O_r> over the 8 dims of head h;  dP = dO_h V_h^T;  dS = P o (dP - D)
            float Dr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                Dr[r] = quad_sum(bf_lo(da[r]) * oh[r / 2][2 * (r % 2)] + bf_hi(da[r]) * oh[r / 2][2 * (r % 2) + 1]);
            float ds[2][4][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) {
                    ds[mi][nj][0] = ds[mi][nj][1] = ds[mi][nj][2] = ds[mi][nj][3] = 0.f;
                    if (AttnU<G>::tile_needed(2 * mi, nj) || AttnU<G>::tile_needed(2 * mi + 1, nj)) {
                        mma1688(ds[mi][nj], da[2 * mi], da[2 * mi + 1], vb[nj]);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                            for (int b = 0; b < 2; ++b) ds[mi][nj][2 * hf + b] = p[mi][nj][2 * hf + b] * (ds[mi][nj][2 * hf + b] - Dr[2 * mi + hf]);
                    }
                }
            // relative-position-bias gradient: warp-private table, every (i, j) pair owned by exactly one lane
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj)
                    if (AttnU<G>::tile_needed(r, nj)) {
                        const int il = (gq + 8 * r) % G, jl = (8 * nj + c0) % G;
                        float2* pd = reinterpret_cast<float2*>(&dB[(h * G + il) * DBS + tcol<G>(gq, jl)]);
                        float2 acc2 = *pd;
                        acc2.x += ds[r / 2][nj][2 * (r % 2)]; acc2.y += ds[r / 2][nj][2 * (r % 2) + 1];
                        *pd = acc2;
                    }
            uint32_t dsb[4][4];
#pragma unroll
            for (int ib = 0; ib < 4; ++ib)
#pragma unroll
                for (int jb = 0; jb < 4; ++jb) dsb[ib][jb] = pk(ds[ib / 2][jb][2 * (ib % 2)], ds[ib / 2][jb][2 * (ib % 2) + 1]);
            // dQ_h = dS K_h (K = keys j);  dK_h = dS^T Q_h, dV_h = P^T dO_h (rows j, K = queries i)
            float dq[2][4], dk[2][4], dv[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                dq[mi][0] = dq[mi][1] = dq[mi][2] = dq[mi][3] = 0.f;
                dk[mi][0] = dk[mi][1] = dk[mi][2] = dk[mi][3] = 0.f;
                dv[mi][0] = dv[mi][1] = dv[mi][2] = dv[mi][3] = 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
                    mma16816(dq[mi], dsb[2 * mi][2 * kk], dsb[2 * mi + 1][2 * kk], dsb[2 * mi][2 * kk + 1], dsb[2 * mi + 1][2 * kk + 1], kt[2 * kk], kt[2 * kk + 1]);
#pragma unroll
            for (int ik = 0; ik < 2; ++ik)
#pragma unroll
                for (int jm = 0; jm < 2; ++jm) {
                    mma16816(dk[jm], movm(dsb[2 * ik][2 * jm]), movm(dsb[2 * ik][2 * jm + 1]), movm(dsb[2 * ik + 1][2 * jm]), movm(dsb[2 * ik + 1][2 * jm + 1]),
                             qt[2 * ik], qt[2 * ik + 1]);
                    mma16816(dv[jm], movm(pb[2 * ik][2 * jm]), movm(pb[2 * ik][2 * jm + 1]), movm(pb[2 * ik + 1][2 * jm]), movm(pb[2 * ik + 1][2 * jm + 1]),
                             dt[2 * ik], dt[2 * ik + 1]);
                }
            // gradient w.r.t. the unscaled projections: dq carries `scale`; dK was formed with q * scale * log2(e)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                st_dq[h][r] = pk(dq[r / 2][2 * (r % 2)] * g.scale, dq[r / 2][2 * (r % 2) + 1] * g.scale);
                st_dk[h][r] = pk(dk[r / 2][2 * (r % 2)] * (1.f / LOG2E), dk[r / 2][2 * (r % 2) + 1] * (1.f / LOG2E));
                st_dv[h][r] = pk(dv[r / 2][2 * (r % 2)], dv[r / 2][2 * (r % 2) + 1]);
                st_o[h][r] = pk(oh[r / 2][2 * (r % 2)], oh[r / 2][2 * (r % 2) + 1]);
            }
        }
        __syncwarp();                 // every lane of the warp has read its q / k / v / dO fragments: the rows can be overwritten
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t a = pl + fr + h * PLANE + r * 128;
                sts32(a + A_Q * PLANE, st_dq[h][r]);
                sts32(a + A_K * PLANE, st_dk[h][r]);
                sts32(a + A_V * PLANE, st_dv[h][r]);
                sts32(a + A_DO * PLANE, st_o[h][r]);
            }
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // dxn = [dq | dk | dv] Wqkv  (K = 48 in three steps)
#pragma unroll
            for (int s3 = 0; s3 < 3; ++s3)
                umma_ss(tmem + B_DXN, make_desc(pl + 2 * s3 * PLANE, PLANE, 128), make_desc(smem_u32(S.wqt) + s3 * 512, 16 * 16, 128), idesc(128, 16), s3 > 0 ? 1u : 0u);
            // weight gradients: D[64 x 48] += [dq | dk | dv | g_y]^T [xn | o | 1 | 0] over the tile's 128 tokens (MN-major views)
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8)
                umma_ss(tmem + B_WACC, make_desc(pl + s8 * 256, 128, PLANE), make_desc(pl + A_XN * PLANE + s8 * 256, 128, PLANE),
                        idesc(WM, 48, 1, 1), (first && s8 == 0) ? 0u : 1u);
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, n_commit++ & 1u);
        tc_fence_after();
        // ---- R2: LayerNorm backward + residual -> g_x (or the fused embedding backward) ----
        {
            float d[16], xn[16], gyr[16];
            tmem_ld16(tlane + B_DXN, d);
            {
                const uint4 a = lds128(pl + A_XN * PLANE + my_row), b = lds128(pl + (A_XN + 1) * PLANE + my_row);
                const uint32_t pn[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                unpack16(pn, xn);
                const uint4 c = lds128(pl + A_GY * PLANE + my_row), e = lds128(pl + (A_GY + 1) * PLANE + my_row);
                const uint32_t pg[8] = {c.x, c.y, c.z, c.w, e.x, e.y, e.z, e.w};
                unpack16(pg, gyr);
            }
            const float m1 = sum16(d) * (1.f / 16.f), m2 = dot16(d, xn) * (1.f / 16.f);
#pragma unroll
            for (int c = 0; c < 16; ++c) d[c] = gyr[c] + rstd * (d[c] - m1 - xn[c] * m2);
            if (tk.valid) {
                uint32_t po_[8];
                pack16(d, po_);
                st8u(gx + tk.off, po_);
            }
        }
        first = false;
        tc_fence_before();
    }
    // ---- per-CTA partials: qkv_w[48*16] | qkv_b[48] | proj_w[16*16] | proj_b[16] | dB[h][i][j] ----
    __syncthreads();
    tc_fence_after();
    float* part = partials + ((int64_t)v * gridDim.x + blockIdx.x) * PART;
    const bool any = (int)blockIdx.x < n_tiles;
    {
        // M = 64: row m in lane (m % 16) + 32 (m / 16): warp w holds rows 16 w .. 16 w + 15 in its lanes 0..15
        float w[16], o2[16], b[16];
        tmem_ld16(tlane + B_WACC, w);
        tmem_ld16(tlane + B_WACC + 16, o2);
        tmem_ld16(tlane + B_WACC + 32, b);
        if (lane < 16) {
            const int m = 16 * warp + lane;
            if (m < 48) {
#pragma unroll
                for (int c = 0; c < 16; ++c) part[m * C + c] = any ? w[c] : 0.f;
                part[3 * C * C + m] = any ? b[0] : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) part[3 * C * C + 3 * C + (m - 48) * C + c] = any ? o2[c] : 0.f;
                part[3 * C * C + 3 * C + C * C + m - 48] = any ? b[0] : 0.f;
            }
        }
    }
    for (int e = tid; e < NH * G * G; e += NT) {
        const int hi = e / G, jl = e % G;     // hi = h * G + i_local
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) sum += dBw_all[w * DBW + hi * DBS + tcol<G>(hi % G, jl)];
        part[ATT_PART_W + e] = sum;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(B_COLS) : "memory");
}

// Backward of the patch embedding e[c] = w[c] x + b[c] -> LayerNorm (PatchEmbed3D, Swin_3D.py:473-491, in_chans == 1) from the bf16
// gradient of its tokens: per token g_e = LN-backward(g_tok), then d w[c] += g_e[c] x, d b[c] += g_e[c].  One streaming pass over
// g_tok (32 B / token) and the raw input (4 B / token); per-thread fp32 sums, fixed-order CTA reduction, partials [V][nblk][32]
// for embed_grad_finalize_kernel.  grid = (nblk, V).
__global__ void __launch_bounds__(256)
embed_bwd_tokens_kernel(const __nv_bfloat16* __restrict__ gtok, const float* __restrict__ xraw, const float* __restrict__ emb_w,
                        const float* __restrict__ emb_b, float* __restrict__ part, int N, int V, int thw) {
    __shared__ float red[8][32];
    const int v = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float w[16], b[16], gw[16], gb[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { w[c] = __ldg(emb_w + v * C + c); b[c] = __ldg(emb_b + v * C + c); gw[c] = 0.f; gb[c] = 0.f; }
    for (int n = 0; n < N; ++n) {
        const int64_t img = (int64_t)(n * V + v) * thw;
        for (int t = blockIdx.x * 256 + tid; t < thw; t += gridDim.x * 256) {
            const float x = __ldg(xraw + img + t);
            uint32_t pg[8];
            ldg_row16(gtok + (img + t) * C, pg);
            float d[16], e[16], en[16];
            unpack16(pg, d);
#pragma unroll
            for (int c = 0; c < 16; ++c) e[c] = w[c] * x + b[c];
            const float rs = ln_row(e, en);
            const float s1 = sum16(d) * (1.f / 16.f), s2 = dot16(d, en) * (1.f / 16.f);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float ge = rs * (d[c] - s1 - en[c] * s2);
                gw[c] += ge * x; gb[c] += ge;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        float a = gw[c], q = gb[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (lane == 0) { red[warp][c] = a; red[warp][16 + c] = q; }
    }
    __syncthreads();
    if (tid < 32) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += red[k][tid];
        part[((int64_t)v * gridDim.x + blockIdx.x) * 32 + tid] = a;
    }
}

template <int WD, int WH, int WW>
int launch_attn_bwd(const idee_swin_desc* d, const Geom& g, const void* x, __nv_bfloat16* gx, const float* params,
                    const int* rel_index, float* gparams, float* part_attn, float* part_mlp, float* part_emb, int per_v, int per_v_mlp,
                    cudaStream_t st) {
    constexpr int G = WD * WH * WW;
    const size_t smem = sizeof(AttnBwdSm) + sizeof(float) * 5 * NH * G * bns(G);
    if (g.emb_x) {
        IDEE_CUDA(cudaFuncSetAttribute(swin_attn_bwd_umma_kernel<WD, WH, WW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_attn_bwd(umma)");
        swin_attn_bwd_umma_kernel<WD, WH, WW, true><<<dim3(per_v, d->V), NT, smem, st>>>(x, gx, gx, params, d->param_stride, rel_index, part_attn, g);
        IDEE_LAUNCH_CHECK("swin_attn_bwd(umma,embed)");
        if (d->embed_gw) {
            // gx now holds the gradient w.r.t. the embedded tokens: one streaming pass turns it into d{embed w, b}
            int occ = 0;                                        // one resident wave (122 registers x 256 threads: 2 CTAs per SM)
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, embed_bwd_tokens_kernel, 256, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
            int nb_e = idee_num_sms() * occ / d->V;
            if (nb_e > per_v) nb_e = per_v;                     // part_emb holds per_v slots per variable
            if (nb_e < 1) nb_e = 1;
            embed_bwd_tokens_kernel<<<dim3(nb_e, d->V), 256, 0, st>>>(gx, g.emb_x, g.emb_w, g.emb_b, part_emb, d->N, d->V, d->T * d->H * d->W);
            IDEE_LAUNCH_CHECK("embed_bwd_tokens");
            embed_grad_finalize_kernel<<<d->V, 32, 0, st>>>(part_emb, nb_e, d->embed_gw, d->embed_gb);
            IDEE_LAUNCH_CHECK("embed_grad_finalize");
        }
    } else {
        IDEE_CUDA(cudaFuncSetAttribute(swin_attn_bwd_umma_kernel<WD, WH, WW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "swin_attn_bwd(umma)");
        swin_attn_bwd_umma_kernel<WD, WH, WW, false><<<dim3(per_v, d->V), NT, smem, st>>>(x, gx, gx, params, d->param_stride, rel_index, part_attn, g);
        IDEE_LAUNCH_CHECK("swin_attn_bwd(umma)");
    }
    if (launch_grad_finalize<G>(part_attn, part_mlp, per_v, per_v_mlp, rel_index, gparams, d->param_stride, g.tbl, d->V, st)) return 2;
    return 0;
}

}  // namespace swu
