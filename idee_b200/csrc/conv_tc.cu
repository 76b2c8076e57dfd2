// bf16 tensor-core implicit-GEMM convolutions of the IDEE hot path (sm_100a), fp32 activations in HBM, fp32 accumulate.
//
// Same geometries and C-ABI as conv.cu (PROJ: 3x3x3 replicate, CLS: (2,3,3)/(2,1,1) zero pad), selected by
// idee_conv_desc.precision == 1.  Layout of the computation:
//   * a CTA owns an 8 x 16 tile of output pixels at a fixed (image, t); its input halo [kt][10][18][Cin] is read from HBM
//     ONCE, converted to bf16 and parked in shared memory (pixel stride Cin+8 halves -> conflict-free ldmatrix);
//   * every tap is a [16 pixels x Cin] x [Cin x Cout] product: A fragments come from the halo with one ldmatrix.x4 per
//     (m-tile, k-step), B fragments from weights that a prep kernel re-ordered into mma fragment order (one LDS.64 each);
//     96-channel weights are streamed per tap through a cp.async double buffer;
//   * the data gradient is the same kernel on transposed/flipped weights.  For the replicate-padded conv it is evaluated on
//     the padded domain (uniform taps, zero outside) and a fold kernel adds the padding ring back onto the border (the
//     adjoint of clamping), which keeps every m-tile on a single weight matrix and the result deterministic;
//   * the weight gradient is a GEMM over pixels: A = halo^T (ldmatrix.trans), B = gout tile (ldmatrix.trans), taps are
//     distributed over the 4 warps, accumulators stay in registers across the CTA's persistent loop over tiles, the bias
//     gradient rides along as one extra mma against a ones row; per-CTA partials are reduced by conv.cu's second stage.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace convtc {

enum { CLS_FWD = 0, PROJ_FWD = 1, CLS_DGRAD = 2, PROJ_DGRAD_PAD = 3 };
constexpr int TH = 8, TW = 16, HH = TH + 2, HW_ = TW + 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem));
}
// 16-byte async copy; src_bytes == 0 writes zeros (the source address must still be valid)
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct P {
    const float* in; float* out; const float* bias; const float* relu_src;
    const uint2* wfrag;          // [wset][oc][ftap][kstep][ntile][lane] (b0b1, b2b3)
    int N, V, Vw;
    int Ti, Hi, Wi, To, Ho, Wo;  // gather-input / output extents (output may be the padded domain)
    int64_t in_sn, in_sv, in_st, in_sh, in_sw, in_sg;
    int64_t out_sn, out_sv, out_st, out_sh, out_sw, out_sg;
    int in_cpg, out_cpg;
    int CO;                      // output channels in total
    int CIr;                     // real gather-input channels (may be < 16*KS: conv3 data gradient has 1)
    int NTf;                     // forward taps (KT*9)
    int relu;
    int tiles_w;
    int KS;                      // k-steps (gather-in channels / 16) for the generic (template KS == 0) kernels
};

// fp32 PyTorch weights [FCO][FCI][NTf] -> bf16 mma B fragments.  B(k, n) = W[o=n][c=k] (forward) or W[o=k][c=n] (dgrad).
__global__ void prep_weights_kernel(const float* __restrict__ w, uint2* __restrict__ wfrag, int FCI, int FCO, int NTf, int dgrad,
                                    int KS, int NTL, int n_oc, int64_t w_set_stride) {
    const int wset = blockIdx.y;
    const int64_t total = (int64_t)n_oc * NTf * KS * NTL * 32;
    const int GI = dgrad ? FCO : FCI, GO = dgrad ? FCI : FCO;   // gather-in / out channel totals
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(e % 32);
        int64_t r = e / 32;
        const int nt = (int)(r % NTL); r /= NTL;
        const int ks = (int)(r % KS); r /= KS;
        const int ft = (int)(r % NTf);
        const int oc = (int)(r / NTf);
        const int n = oc * NTL * 8 + nt * 8 + lane / 4;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = ks * 16 + (lane % 4) * 2 + (q & 1) + (q >> 1) * 8;
            float val = 0.f;
            if (k < GI && n < GO) {
                const int fo = dgrad ? k : n, fc = dgrad ? n : k;
                val = w[wset * w_set_stride + ((int64_t)fo * FCI + fc) * NTf + ft];
            }
            v[q] = val;
        }
        wfrag[(int64_t)wset * total + e] = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
}

// KS = gather-in channels / 16 (0: runtime p.KS), NTL = n-tiles (of 8 output channels) per CTA, STREAM = weights streamed per tap
template <int MODE, int KS_T, int NTL, bool STREAM>
__global__ void __launch_bounds__(128)
conv_tc_kernel(P p) {
    const int KS = KS_T ? KS_T : p.KS;
    const int CI = KS * 16, CP = CI + 8;                       // halo pixel stride in halves
    constexpr int KTIN = (MODE == CLS_FWD) ? 2 : (MODE == CLS_DGRAD ? 1 : 3);
    constexpr int NJ = (MODE == CLS_FWD) ? 18 : (MODE == CLS_DGRAD ? 9 : 27);
    const int WTAP = KS * NTL * 32;                            // uint2 per tap
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(smem_raw);                      // [KTIN][HH][HW_][CP]
    uint2* wsm = reinterpret_cast<uint2*>(smem_raw + (size_t)KTIN * HH * HW_ * CP * 2);    // resident: [NJ'][WTAP]; stream: [2][WTAP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_h = blockIdx.x / p.tiles_w, tile_w = blockIdx.x % p.tiles_w;
    const int h0 = tile_h * TH, w0 = tile_w * TW;
    int it = blockIdx.y;
    const int t = it % p.To; it /= p.To;
    const int v = it % p.V, n = it / p.V;
    const int oc = blockIdx.z, n_oc = gridDim.z;
    const int wset = p.Vw == 1 ? 0 : v;
    const uint2* wf = p.wfrag + ((int64_t)wset * n_oc + oc) * p.NTf * WTAP;

    // forward tap of gather tap j (uniform over the CTA)
    auto ftap_of = [&](int j) -> int {
        if (MODE == CLS_FWD || MODE == PROJ_FWD) return j;
        if (MODE == CLS_DGRAD) return (t & 1) * 9 + (2 - j / 3) * 3 + (2 - j % 3);
        return (2 - j / 9) * 9 + (2 - (j / 3) % 3) * 3 + (2 - j % 3);
    };
    if (!STREAM) {
        for (int e = tid; e < NJ * WTAP; e += 128) wsm[e] = wf[(int64_t)ftap_of(e / WTAP) * WTAP + e % WTAP];
    } else {
        for (int e = tid; e < WTAP / 2; e += 128) cp_async16(&wsm[2 * e], &wf[(int64_t)ftap_of(0) * WTAP + 2 * e]);
        cp_async_commit();
    }
    // ---- halo: fp32 HBM -> bf16 smem (zero / clamp handled here, so the MMA loop is branch-free) ----
    const float* in_img = p.in + n * p.in_sn + v * p.in_sv;
    const int V4 = CI / 4;
    // Halo staging: element e = row * NCOL + col (row = (kt, hh) halo row, col = (pixel column, float4 channel group)).
    // All 128 threads stride through e with an incrementally updated (row, col) pair (no per-element division by the
    // halo extents) and keep 4 independent 16-byte loads in flight before converting/storing.
    {
        const int NCOL = HW_ * V4, NROW = KTIN * HH;
        int row = tid / NCOL, col = tid - row * NCOL;
        const int drow = 128 / NCOL, dcol = 128 - drow * NCOL;
        while (row < NROW) {
            const float* src[4];
            int dst[4], chs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                src[u] = nullptr; dst[u] = -1; chs[u] = 0;
                if (row < NROW) {
                    const int kt = row / HH, hh = row - kt * HH;
                    const int ww = col / V4, c4 = col - ww * V4;
                    int ti, hi = h0 + hh - 1, wi = w0 + ww - 1;
                    bool ok = true;
                    if (MODE == CLS_FWD) { ti = 2 * t + kt; ok = hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi; }
                    else if (MODE == PROJ_FWD) { ti = min(max(t + kt - 1, 0), p.Ti - 1); hi = min(max(hi, 0), p.Hi - 1); wi = min(max(wi, 0), p.Wi - 1); }
                    else if (MODE == CLS_DGRAD) { ti = t >> 1; ok = ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi; }
                    else { ti = t + kt - 2; hi -= 1; wi -= 1; ok = ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi; }
                    const int ch = c4 * 4, chunk = ch >> 4;
                    dst[u] = (row * HW_ + ww) * CP + ch;
                    chs[u] = ch;
                    if (ok) src[u] = in_img + ti * p.in_st + hi * p.in_sh + wi * p.in_sw + (chunk / p.in_cpg) * p.in_sg + (chunk % p.in_cpg) * 16 + (ch & 15);
                    row += drow; col += dcol;
                    if (col >= NCOL) { col -= NCOL; ++row; }
                }
            }
            float4 f[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src[u]) {
                    if (p.CIr >= CI) f[u] = ldg4(src[u]);
                    else {
                        if (chs[u] < p.CIr) f[u].x = __ldg(src[u]);
                        if (chs[u] + 1 < p.CIr) f[u].y = __ldg(src[u] + 1);
                        if (chs[u] + 2 < p.CIr) f[u].z = __ldg(src[u] + 2);
                        if (chs[u] + 3 < p.CIr) f[u].w = __ldg(src[u] + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dst[u] >= 0) *reinterpret_cast<uint2*>(halo + dst[u]) = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
        }
    }
    float acc[2][NTL][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][nt][q] = 0.f;
    __syncthreads();

    // ldmatrix row address of this lane inside an m-tile (16 consecutive w of one tile row)
    const int a_pix = (lane & 7) + ((lane >> 3) & 1) * 8, a_koff = (lane >> 4) * 8;
    for (int j = 0; j < NJ; ++j) {
        const uint2* wt;
        if (STREAM) {
            if (j + 1 < NJ) {
                uint2* dst = wsm + ((j + 1) & 1) * WTAP;
                const uint2* src = wf + (int64_t)ftap_of(j + 1) * WTAP;
                for (int e = tid; e < WTAP / 2; e += 128) cp_async16(&dst[2 * e], &src[2 * e]);
                cp_async_commit();
                cp_async_wait<1>();
            } else cp_async_wait<0>();
            __syncthreads();
            wt = wsm + (j & 1) * WTAP;
        } else wt = wsm + j * WTAP;
        int kt, kh, kw;
        if (MODE == CLS_DGRAD) { kt = 0; kh = j / 3; kw = j % 3; }
        else { kt = j / 9; kh = (j / 3) % 3; kw = j % 3; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int row = warp * 2 + m;
                ldsm_x4(a[m], halo + ((size_t)(kt * HH + row + kh) * HW_ + kw + a_pix) * CP + ks * 16 + a_koff);
            }
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const uint2 b = wt[(ks * NTL + nt) * 32 + lane];
                mma_bf16(acc[0][nt], a[0], b.x, b.y);
                mma_bf16(acc[1][nt], a[1], b.x, b.y);
            }
        }
        if (STREAM) __syncthreads();
    }
    // ---- epilogue: bias, ReLU (forward) or ReLU mask (dgrad), fp32 stores ----
    const float* B = p.bias ? p.bias + (int64_t)wset * p.CO : nullptr;
    float* out_img = p.out + n * p.out_sn + v * p.out_sv + t * p.out_st;
    const float* rs_img = p.relu_src ? p.relu_src + n * p.out_sn + v * p.out_sv + t * p.out_st : nullptr;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int h = h0 + warp * 2 + m;
        if (h >= p.Ho) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int w = w0 + lane / 4 + half * 8;
            if (w >= p.Wo) continue;
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const int co = oc * NTL * 8 + nt * 8 + (lane % 4) * 2;
                if (co >= p.CO) continue;
                const int chunk = co / 16;
                const int64_t o = h * p.out_sh + w * p.out_sw + (chunk / p.out_cpg) * p.out_sg + (chunk % p.out_cpg) * 16 + (co % 16);
                float v0 = acc[m][nt][half * 2], v1 = acc[m][nt][half * 2 + 1];
                if (B) { v0 += B[co]; if (co + 1 < p.CO) v1 += B[co + 1]; }
                if (p.relu && MODE <= PROJ_FWD) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                if (rs_img) { if (!(rs_img[o] > 0.f)) v0 = 0.f; if (co + 1 < p.CO && !(rs_img[o + 1] > 0.f)) v1 = 0.f; }
                if (co + 1 < p.CO) *reinterpret_cast<float2*>(out_img + o) = make_float2(v0, v1);
                else out_img[o] = v0;
            }
        }
    }
}

constexpr int C16_THREADS = 128;     // 4 warps per CTA, two tile rows (m-tiles) per warp (8 warps measured slower)

// ------------------------------------------------------------------------------------------------------------------
// 16-input-channel convs (proj_var, per-variable classifier heads, and their data gradients): persistent, pipelined.
// A CTA loads the weights once, then walks tiles:  wait(stage i) -> convert fp32 stage -> bf16 halo -> issue cp.async for
// tile i+1 (zero-fill for padding, clamped addresses for replicate) -> MMA + epilogue of tile i.  The global-memory latency
// of the next tile is hidden behind the tensor-core phase of the current one; taps are fully unrolled.
// ------------------------------------------------------------------------------------------------------------------
template <int MODE, int NTL>
__global__ void __launch_bounds__(C16_THREADS)
conv_tc16_kernel(P p, int64_t total_tiles64, int tiles_h) {
    constexpr int CP = 24, NTH = C16_THREADS, MT = 8 / (NTH / 32);   // m-tiles (tile rows of 16 pixels) per warp
    constexpr int KTIN = (MODE == CLS_FWD) ? 2 : (MODE == CLS_DGRAD ? 1 : 3);
    constexpr int NJ = (MODE == CLS_FWD) ? 18 : (MODE == CLS_DGRAD ? 9 : 27);
    constexpr int NTF = (MODE == CLS_FWD || MODE == CLS_DGRAD) ? 18 : 27;
    constexpr int WTAP = NTL * 32, NPIX = KTIN * HH * HW_;
    constexpr int NCOL = HW_ * 4, TOTAL = KTIN * HH * NCOL, NEL = (TOTAL + NTH - 1) / NTH;   // 16-byte halo elements
    constexpr int OT = (MODE == PROJ_FWD) ? -1 : (MODE == PROJ_DGRAD_PAD ? -2 : 0);          // halo origin relative to the tile
    constexpr int OHW = (MODE == PROJ_DGRAD_PAD) ? -2 : -1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* wsm = reinterpret_cast<uint2*>(smem_raw);                                     // [NTF][WTAP]
    float* stage = reinterpret_cast<float*>(smem_raw + sizeof(uint2) * NTF * WTAP);      // [NPIX][16] fp32
    __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(stage + NPIX * 16);           // [NPIX][CP] bf16
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total_tiles = (uint32_t)total_tiles64;

    // Per-thread constants: the halo elements a thread fetches are the same for every tile, so their source offsets
    // relative to the halo origin are computed once (32-bit), leaving one 64-bit add per element on interior tiles.
    int rel_src[NEL];
#pragma unroll
    for (int i = 0; i < NEL; ++i) {
        const int e = tid + i * NTH, row = e / NCOL, col = e - row * NCOL;
        const int kt = row / HH, hh = row - kt * HH, ww = col >> 2, c4 = col & 3;
        rel_src[i] = (int)(kt * p.in_st + hh * p.in_sh + ww * p.in_sw) + c4 * 4;
    }
    // epilogue offsets of this thread's 2*MT (row, half) output pixels relative to the tile origin
    int rel_out[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) rel_out[m][hf] = (int)((warp * MT + m) * p.out_sh + (lane / 4 + hf * 8) * p.out_sw) + (lane % 4) * 2;

    auto decode = [&](uint32_t tile, int& n, int& v, int& t, int& h0, int& w0) {
        uint32_t r = tile;                                       // 32-bit: tw fastest, then th, t, v, n
        uint32_t q = r / (uint32_t)p.tiles_w; w0 = (int)(r - q * p.tiles_w) * TW; r = q;
        q = r / (uint32_t)tiles_h; h0 = (int)(r - q * tiles_h) * TH; r = q;
        q = r / (uint32_t)p.To; t = (int)(r - q * p.To); r = q;
        q = r / (uint32_t)p.V; v = (int)(r - q * p.V); n = (int)q;
    };
    auto issue = [&](uint32_t tile) {
        int n, v, t, h0, w0;
        decode(tile, n, v, t, h0, w0);
        const float* in_img = p.in + n * p.in_sn + v * p.in_sv;
        const int tb = (MODE == CLS_FWD) ? 2 * t : (MODE == CLS_DGRAD ? (t >> 1) : t);
        const int t_lo = tb + OT, h_lo = h0 + OHW, w_lo = w0 + OHW;      // halo origin
        const bool interior = t_lo >= 0 && t_lo + KTIN - 1 < p.Ti && h_lo >= 0 && h_lo + HH - 1 < p.Hi && w_lo >= 0 && w_lo + HW_ - 1 < p.Wi;
        if (interior) {
            const float* base = in_img + t_lo * p.in_st + h_lo * p.in_sh + w_lo * p.in_sw;
#pragma unroll
            for (int i = 0; i < NEL; ++i) {
                const int e = tid + i * NTH;
                if (e < TOTAL) cp_async16_zfill(stage + e * 4, base + rel_src[i], 16);
            }
        } else {
#pragma unroll 1
            for (int e = tid; e < TOTAL; e += NTH) {
                const int row = e / NCOL, col = e - row * NCOL;
                const int kt = row / HH, hh = row - kt * HH, ww = col >> 2, c4 = col & 3;
                int ti = t_lo + kt, hi = h_lo + hh, wi = w_lo + ww;
                bool ok = true;
                if (MODE == PROJ_FWD) { ti = min(max(ti, 0), p.Ti - 1); hi = min(max(hi, 0), p.Hi - 1); wi = min(max(wi, 0), p.Wi - 1); }
                else ok = ti >= 0 && ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
                const float* src = ok ? in_img + ti * p.in_st + hi * p.in_sh + wi * p.in_sw + c4 * 4 : p.in;
                cp_async16_zfill(stage + e * 4, src, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };

    const uint32_t first = blockIdx.x;
    if (first < total_tiles) issue(first);
    int cur_wset = -1;
    const int a_pix = (lane & 7) + ((lane >> 3) & 1) * 8, a_koff = (lane >> 4) * 8;
    for (uint32_t tile = first; tile < total_tiles; tile += gridDim.x) {
        int n, v, t, h0, w0;
        decode(tile, n, v, t, h0, w0);
        const int wset = p.Vw == 1 ? 0 : v;
        cp_async_wait<0>();
        __syncthreads();                                   // stage(tile) landed; every warp is done with the previous halo/weights
        if (wset != cur_wset) {
            const uint2* wf = p.wfrag + (int64_t)wset * NTF * WTAP;
            for (int e = tid; e < NTF * WTAP; e += NTH) wsm[e] = wf[e];
            cur_wset = wset;
        }
#pragma unroll
        for (int i = 0; i < (NPIX * 4 + NTH - 1) / NTH; ++i) {   // fp32 stage -> bf16 halo (padded pixel stride)
            const int e = tid + i * NTH;
            if (e < NPIX * 4) {
                const float4 f = ld4(stage + e * 4);
                *reinterpret_cast<uint2*>(halo + (e >> 2) * CP + (e & 3) * 4) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
            }
        }
        __syncthreads();
        if (tile + gridDim.x < total_tiles) issue(tile + gridDim.x);

        float acc[MT][NTL][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
        const __nv_bfloat16* abase = halo + ((warp * MT) * HW_ + a_pix) * CP + a_koff;
        const uint2* wpar = wsm + ((MODE == CLS_DGRAD) ? (t & 1) * 9 * WTAP : 0) + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int kt = (MODE == CLS_DGRAD) ? 0 : j / 9, kh = (MODE == CLS_DGRAD) ? j / 3 : (j / 3) % 3, kw = j % 3;
            const int ft = (MODE == CLS_FWD || MODE == PROJ_FWD) ? j
                         : (MODE == CLS_DGRAD ? (2 - kh) * 3 + (2 - kw) : (2 - kt) * 9 + (2 - kh) * 3 + (2 - kw));
            uint32_t af[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) ldsm_x4(af[m], abase + ((kt * HH + kh + m) * HW_ + kw) * CP);
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const uint2 b = wpar[ft * WTAP + nt * 32];
#pragma unroll
                for (int m = 0; m < MT; ++m) mma_bf16(acc[m][nt], af[m], b.x, b.y);
            }
        }
        // epilogue
        const float* B = p.bias ? p.bias + (int64_t)wset * p.CO : nullptr;
        const int64_t tile_off = n * p.out_sn + v * p.out_sv + t * p.out_st + h0 * p.out_sh + w0 * p.out_sw;
        float* out_tile = p.out + tile_off;
        const float* rs_tile = p.relu_src ? p.relu_src + tile_off : nullptr;
        float bias0[NTL], bias1[NTL];
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            const int co = nt * 8 + (lane % 4) * 2;
            bias0[nt] = (B && co < p.CO) ? B[co] : 0.f;
            bias1[nt] = (B && co + 1 < p.CO) ? B[co + 1] : 0.f;
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (h0 + warp * MT + m >= p.Ho) continue;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (w0 + lane / 4 + half * 8 >= p.Wo) continue;
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int co = nt * 8 + (lane % 4) * 2;
                    if (co >= p.CO) continue;
                    const int o = rel_out[m][half] + nt * 8;
                    float v0 = acc[m][nt][half * 2] + bias0[nt], v1 = acc[m][nt][half * 2 + 1] + bias1[nt];
                    if (p.relu && MODE <= PROJ_FWD) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    if (rs_tile) { if (!(rs_tile[o] > 0.f)) v0 = 0.f; if (co + 1 < p.CO && !(rs_tile[o + 1] > 0.f)) v1 = 0.f; }
                    if (co + 1 < p.CO) *reinterpret_cast<float2*>(out_tile + o) = make_float2(v0, v1);
                    else out_tile[o] = v0;
                }
            }
        }
    }
}

// gin[r] = sum of gpad over the padded positions that clamp to r (adjoint of replicate padding), optional ReLU mask
__global__ void fold_pad_kernel(const float* __restrict__ gpad, float* __restrict__ gin, const float* __restrict__ relu_src,
                                int NV, int T, int H, int W) {
    const int64_t total = (int64_t)NV * T * H * W * 4;   // float4 units of the 16 channels
    const int64_t Wp = W + 2, Hp = H + 2, Tp = T + 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e & 3);
        int64_t r = e >> 2;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H); r /= H;
        const int t = (int)(r % T);
        const int64_t img = r / T;
        const int t0 = t == 0 ? 0 : t + 1, t1 = t == T - 1 ? T + 1 : t + 1;
        const int h0 = h == 0 ? 0 : h + 1, h1 = h == H - 1 ? H + 1 : h + 1;
        const int w0 = w == 0 ? 0 : w + 1, w1 = w == W - 1 ? W + 1 : w + 1;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tt = t0; tt <= t1; ++tt)
            for (int hh = h0; hh <= h1; ++hh)
                for (int ww = w0; ww <= w1; ++ww) {
                    const float4 g = ldg4(gpad + ((((img * Tp + tt) * Hp + hh) * Wp + ww) * 16 + c4 * 4));
                    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
                }
        if (relu_src) {
            const float4 a = ldg4(relu_src + e * 4);
            if (!(a.x > 0.f)) s.x = 0.f; if (!(a.y > 0.f)) s.y = 0.f; if (!(a.z > 0.f)) s.z = 0.f; if (!(a.w > 0.f)) s.w = 0.f;
        }
        st4(gin + e * 4, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------------
struct WP {
    const float* in; const float* gout; float* partials;
    int N, V, Vw;
    int Ti, Hi, Wi, To, Ho, Wo;
    int64_t in_sn, in_sv, in_st, in_sh, in_sw, in_sg;
    int64_t go_sn, go_sv, go_st, go_sh, go_sw;
    int in_cpg, FCO, proj;
    int n_ic, n_oc16, S;        // n_oc16: 16-wide output chunks in total (partials layout of conv.cu)
    int tiles_h, tiles_w;
    int64_t tiles_per_set;      // imgs_per_set * To * tiles_h * tiles_w
};

// NT taps, NTL n-tiles (8 output channels each) per CTA; cin chunk 16
template <int NT, int NTL>
__global__ void __launch_bounds__(128)
wgrad_tc_kernel(WP p) {
    constexpr int KTIN = NT / 9, CPA = 24, NC = NTL * 8, CPG = NC + 8, TPW = (NT + 3) / 4, NPIX = KTIN * HH * HW_;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* stageA = reinterpret_cast<float*>(smem_raw);                                  // [NPIX][16] fp32
    float* stageG = stageA + NPIX * 16;                                                  // [128][NC] fp32
    __nv_bfloat16* tileA = reinterpret_cast<__nv_bfloat16*>(stageG + TH * TW * NC);      // [NPIX][CPA]
    __nv_bfloat16* tileG = tileA + NPIX * CPA;                                           // [128][CPG]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x, wset = blockIdx.y;
    const int n_occ = (p.n_oc16 * 16 + NC - 1) / NC;          // output chunks of NC channels
    const int ic = blockIdx.z / n_occ, occ = blockIdx.z % n_occ;
    float acc[TPW][NTL][4], accb[NTL][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][nt][q] = 0.f;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) accb[nt][q] = 0.f;
    const bool do_bias = (ic == 0 && warp == 3);              // the lightest-loaded warp carries the bias row
    uint32_t ones[4];
    ones[0] = ones[2] = (lane < 4) ? 0x3F803F80u : 0u;        // A row 0 = 1.0 (bf16), all other rows 0
    ones[1] = ones[3] = 0u;

    const int imgs_per_n = p.Vw == 1 ? p.V : 1;
    const int64_t t_begin = p.tiles_per_set * s / p.S, t_end = p.tiles_per_set * (s + 1) / p.S;
    const int a_pix = (lane & 7) + (lane >> 4) * 8, a_coff = ((lane >> 3) & 1) * 8;     // A (trans) lane address
    const int b_pix = (lane & 7) + ((lane >> 3) & 1) * 8, b_coff = (lane >> 4) * 8;     // B (trans) lane address
    const int64_t coff = (ic / p.in_cpg) * p.in_sg + (ic % p.in_cpg) * 16;
    const bool g_vec = (p.FCO % 4) == 0;                      // gout rows can be fetched in 16-byte pieces
    // smem offsets (halves) of this warp's taps inside the halo tile
    int a_off[TPW];
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int ft = warp + 4 * i, kt = ft / 9, kh = (ft / 3) % 3, kw = ft % 3;
        a_off[i] = (((kt * HH + kh) * HW_ + kw) + a_pix) * CPA + a_coff;
    }
    constexpr int NCOLA = HW_ * 4, TOTALA = KTIN * HH * NCOLA, NELA = (TOTALA + 127) / 128;
    constexpr int TOTALG = TH * TW * (NC / 4), NELG = (TOTALG + 127) / 128;
    // tile-independent relative offsets of the elements this thread fetches (32-bit), see conv_tc16_kernel
    int rel_a[NELA], rel_g[NELG];
#pragma unroll
    for (int i = 0; i < NELA; ++i) {
        const int e = tid + i * 128, row = e / NCOLA, col = e - row * NCOLA;
        const int kt = row / HH, hh = row - kt * HH, ww = col >> 2, c4 = col & 3;
        rel_a[i] = (int)(kt * p.in_st + hh * p.in_sh + ww * p.in_sw) + c4 * 4;
    }
#pragma unroll
    for (int i = 0; i < NELG; ++i) {
        const int e = tid + i * 128, c4 = e % (NC / 4), pix = e / (NC / 4);
        rel_g[i] = (int)((pix / TW) * p.go_sh + (pix % TW) * p.go_sw) + occ * NC + c4 * 4;
    }
    const bool g_full = occ * NC + NC <= p.FCO;             // every channel of this output chunk exists
    auto decode = [&](int64_t tile64, int& n, int& v, int& t, int& h0, int& w0) {
        uint32_t r = (uint32_t)tile64;             // t fastest: consecutive tiles of a CTA share input t-slices (L1/L2 hits)
        uint32_t q = r / (uint32_t)p.To; t = (int)(r - q * p.To); r = q;
        q = r / (uint32_t)p.tiles_w; w0 = (int)(r - q * p.tiles_w) * TW; r = q;
        q = r / (uint32_t)p.tiles_h; h0 = (int)(r - q * p.tiles_h) * TH;
        const uint32_t img = q;
        n = (int)(img / (uint32_t)imgs_per_n); v = p.Vw == 1 ? (int)(img % (uint32_t)imgs_per_n) : wset;
    };
    auto issue = [&](int64_t tile) {
        int n, v, t, h0, w0;
        decode(tile, n, v, t, h0, w0);
        const float* in_img = p.in + n * p.in_sn + v * p.in_sv + coff;
        const int t_lo = p.proj ? t - 1 : 2 * t, h_lo = h0 - 1, w_lo = w0 - 1;
        const bool interior = t_lo >= 0 && t_lo + KTIN - 1 < p.Ti && h_lo >= 0 && h_lo + HH - 1 < p.Hi && w_lo >= 0 && w_lo + HW_ - 1 < p.Wi;
        if (interior) {
            const float* base = in_img + t_lo * p.in_st + h_lo * p.in_sh + w_lo * p.in_sw;
#pragma unroll
            for (int i = 0; i < NELA; ++i) {
                const int e = tid + i * 128;
                if (e < TOTALA) cp_async16_zfill(stageA + e * 4, base + rel_a[i], 16);
            }
        } else {
#pragma unroll 1
            for (int e = tid; e < TOTALA; e += 128) {
                const int row = e / NCOLA, col = e - row * NCOLA;
                const int kt = row / HH, hh = row - kt * HH, ww = col >> 2, c4 = col & 3;
                int ti = t_lo + kt, hi = h_lo + hh, wi = w_lo + ww;
                bool ok = true;
                if (p.proj) { ti = min(max(ti, 0), p.Ti - 1); hi = min(max(hi, 0), p.Hi - 1); wi = min(max(wi, 0), p.Wi - 1); }
                else ok = hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
                const float* src = ok ? in_img + ti * p.in_st + hi * p.in_sh + wi * p.in_sw + c4 * 4 : p.in;
                cp_async16_zfill(stageA + e * 4, src, ok ? 16 : 0);
            }
        }
        if (g_vec) {
            const float* go_tile = p.gout + n * p.go_sn + v * p.go_sv + t * p.go_st + h0 * p.go_sh + w0 * p.go_sw;
            if (g_full && h0 + TH <= p.Ho && w0 + TW <= p.Wo) {
#pragma unroll
                for (int i = 0; i < NELG; ++i) {
                    const int e = tid + i * 128;
                    if (e < TOTALG) cp_async16_zfill(stageG + e * 4, go_tile + rel_g[i], 16);
                }
            } else {
#pragma unroll 1
                for (int e = tid; e < TOTALG; e += 128) {
                    const int c4 = e % (NC / 4), pix = e / (NC / 4);
                    const int h = h0 + pix / TW, w = w0 + pix % TW;
                    const int co = occ * NC + c4 * 4;
                    const bool ok = h < p.Ho && w < p.Wo && co + 3 < p.FCO;
                    const float* src = ok ? go_tile + (pix / TW) * p.go_sh + (pix % TW) * p.go_sw + co : p.gout;
                    cp_async16_zfill(stageG + e * 4, src, ok ? 16 : 0);
                }
            }
        }
        cp_async_commit();
    };
    if (t_begin < t_end) issue(t_begin);
    for (int64_t tile = t_begin; tile < t_end; ++tile) {
        cp_async_wait<0>();
        __syncthreads();
        for (int e = tid; e < NPIX * 4; e += 128) {
            const float4 f = ld4(stageA + e * 4);
            *reinterpret_cast<uint2*>(tileA + (e >> 2) * CPA + (e & 3) * 4) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
        }
        if (g_vec) {
            for (int e = tid; e < TH * TW * (NC / 4); e += 128) {
                const int c4 = e % (NC / 4), pix = e / (NC / 4);
                const float4 f = ld4(stageG + e * 4);
                *reinterpret_cast<uint2*>(tileG + pix * CPG + c4 * 4) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
            }
        } else {                                   // Cout not a multiple of 4 (the 1-channel logit convs): scalar fetch
            int n, v, t, h0, w0;
            decode(tile, n, v, t, h0, w0);
            const float* go_img = p.gout + n * p.go_sn + v * p.go_sv + t * p.go_st;
            for (int e = tid; e < TH * TW * (NC / 4); e += 128) {
                const int c4 = e % (NC / 4), pix = e / (NC / 4);
                const int h = h0 + pix / TW, w = w0 + pix % TW;
                const int co = occ * NC + c4 * 4;
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (h < p.Ho && w < p.Wo) {
                    const float* g = go_img + h * p.go_sh + w * p.go_sw + co;
                    if (co < p.FCO) f.x = __ldg(g); if (co + 1 < p.FCO) f.y = __ldg(g + 1);
                    if (co + 2 < p.FCO) f.z = __ldg(g + 2); if (co + 3 < p.FCO) f.w = __ldg(g + 3);
                }
                *reinterpret_cast<uint2*>(tileG + pix * CPG + c4 * 4) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
            }
        }
        __syncthreads();
        if (tile + 1 < t_end) issue(tile + 1);      // next tile's loads fly while this tile's MMAs run
#pragma unroll
        for (int ks = 0; ks < TH; ++ks) {            // k-step = one tile row of 16 pixels
            uint32_t b[NTL / 2 > 0 ? NTL / 2 : 1][4];
#pragma unroll
            for (int np = 0; np < (NTL + 1) / 2; ++np) ldsm_x4_t(b[np], tileG + (size_t)(ks * TW + b_pix) * CPG + np * 16 + b_coff);
            if (do_bias) {
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) mma_bf16(accb[nt], ones, b[nt / 2][(nt & 1) * 2], b[nt / 2][(nt & 1) * 2 + 1]);
            }
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (warp + 4 * i < NT) {
                    uint32_t a[4];
                    ldsm_x4_t(a, tileA + a_off[i] + ks * HW_ * CPA);
#pragma unroll
                    for (int nt = 0; nt < NTL; ++nt) mma_bf16(acc[i][nt], a, b[nt / 2][(nt & 1) * 2], b[nt / 2][(nt & 1) * 2 + 1]);
                }
            }
        }
    }
    // partials in conv.cu's layout: [wset][ic][oc16][s][NT*256 + 16], entry (ft, c, o) at ft*256 + c*16 + o
    constexpr int PS = NT * 256 + 16;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
        const int co = occ * NC + nt * 8 + (lane % 4) * 2;     // two consecutive output channels co, co+1
        const int oc16 = co / 16, ol = co % 16;
        if (oc16 >= p.n_oc16) continue;
        float* part = p.partials + ((((int64_t)wset * p.n_ic + ic) * p.n_oc16 + oc16) * p.S + s) * PS;
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const int ft = warp + 4 * i;
            if (ft < NT) {
                const int c = lane / 4;
                part[ft * 256 + c * 16 + ol] = acc[i][nt][0];
                part[ft * 256 + c * 16 + ol + 1] = acc[i][nt][1];
                part[ft * 256 + (c + 8) * 16 + ol] = acc[i][nt][2];
                part[ft * 256 + (c + 8) * 16 + ol + 1] = acc[i][nt][3];
            }
        }
        if (do_bias && lane < 4) { part[NT * 256 + ol] = accb[nt][0]; part[NT * 256 + ol + 1] = accb[nt][1]; }
    }
}

}  // namespace convtc

// ------------------------------------------------------------------------------------------------------------------
// host side (called from conv.cu's entry points when desc.precision == 1)
// ------------------------------------------------------------------------------------------------------------------
using namespace convtc;

namespace {

struct Plan { int KS, NTL, n_oc, NTf, NJ, KTIN; bool stream; size_t wfrag_bytes, smem; };

// gi/go: gather-in / output channel totals of this launch
Plan make_plan(int mode, int gi, int go, int Vw) {
    Plan pl{};
    pl.KS = (gi + 15) / 16;
    pl.NTf = (mode == PROJ_FWD || mode == PROJ_DGRAD_PAD) ? 27 : 18;
    pl.NJ = mode == CLS_DGRAD ? 9 : pl.NTf;
    pl.KTIN = mode == CLS_FWD ? 2 : (mode == CLS_DGRAD ? 1 : 3);
    pl.stream = pl.KS > 1;
    if (go == 96 && (pl.KS == 1 || pl.KS == 6)) pl.NTL = 12; else if (go >= 16) pl.NTL = 2; else pl.NTL = 1;
    if (pl.NTL == 2 && pl.KS != 1) pl.stream = true;
    pl.n_oc = (go + pl.NTL * 8 - 1) / (pl.NTL * 8);
    pl.wfrag_bytes = sizeof(uint2) * (size_t)Vw * pl.n_oc * pl.NTf * pl.KS * pl.NTL * 32;
    const size_t halo = (size_t)pl.KTIN * HH * HW_ * (pl.KS * 16 + 8) * 2;
    const size_t wtap = sizeof(uint2) * (size_t)pl.KS * pl.NTL * 32;
    pl.smem = halo + (pl.stream ? 2 * wtap : pl.NJ * wtap);
    return pl;
}

template <int MODE, int KS, int NTL, bool STREAM>
int launch_tc(P p, const Plan& pl, int n_img_t, cudaStream_t st, const char* who) {
    p.KS = pl.KS;
    auto kern = conv_tc_kernel<MODE, KS, NTL, STREAM>;
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem), who);
    const int tiles_h = (p.Ho + TH - 1) / TH;
    dim3 grid(tiles_h * p.tiles_w, n_img_t, pl.n_oc);
    kern<<<grid, 128, pl.smem, st>>>(p);
    IDEE_LAUNCH_CHECK(who);
    return 0;
}

template <int MODE, int NTL>
int launch_tc16(const P& p, int n_img_t, cudaStream_t st, const char* who) {
    constexpr int KTIN = (MODE == CLS_FWD) ? 2 : (MODE == CLS_DGRAD ? 1 : 3);
    constexpr int NTF = (MODE == CLS_FWD || MODE == CLS_DGRAD) ? 18 : 27;
    const size_t smem = sizeof(uint2) * NTF * NTL * 32 + (size_t)KTIN * HH * HW_ * (16 * 4 + 24 * 2);
    auto kern = conv_tc16_kernel<MODE, NTL>;
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), who);
    const int tiles_h = (p.Ho + TH - 1) / TH;
    const int64_t total = (int64_t)n_img_t * tiles_h * p.tiles_w;
    IDEE_REQUIRE(total < (1ll << 31) && 3 * p.in_st + HH * p.in_sh + HW_ * p.in_sw < (1ll << 31) &&
                 TH * p.out_sh + TW * p.out_sw < (1ll << 31), "%s: tensor too large for 32-bit tile-relative offsets", who);
    int per_sm = 1;
    IDEE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C16_THREADS, smem), who);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)idee_num_sms() * per_sm;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, C16_THREADS, smem, st>>>(p, total, tiles_h);
    IDEE_LAUNCH_CHECK(who);
    return 0;
}

template <int MODE>
int dispatch_tc(const P& p, const Plan& pl, int n_img_t, cudaStream_t st, const char* who) {
    if (pl.KS == 1 && p.CIr == 16 && pl.n_oc == 1 && pl.NTL == 2) return launch_tc16<MODE, 2>(p, n_img_t, st, who);
    if (pl.KS == 1 && p.CIr == 16 && pl.n_oc == 1 && pl.NTL == 1) return launch_tc16<MODE, 1>(p, n_img_t, st, who);
    if (pl.KS == 1 && pl.NTL == 2) return launch_tc<MODE, 1, 2, false>(p, pl, n_img_t, st, who);
    if (pl.KS == 1 && pl.NTL == 1) return launch_tc<MODE, 1, 1, false>(p, pl, n_img_t, st, who);
    if (pl.KS == 1 && pl.NTL == 12) return launch_tc<MODE, 1, 12, false>(p, pl, n_img_t, st, who);
    if (pl.KS == 6 && pl.NTL == 12) return launch_tc<MODE, 6, 12, true>(p, pl, n_img_t, st, who);
    if (pl.KS == 6 && pl.NTL == 1) return launch_tc<MODE, 6, 1, true>(p, pl, n_img_t, st, who);
    // any other multiple-of-16 channel count (e.g. the joint head of a model with in_vars != 6): runtime k-step count
    if (pl.stream && pl.NTL == 2) return launch_tc<MODE, 0, 2, true>(p, pl, n_img_t, st, who);
    if (pl.stream && pl.NTL == 1) return launch_tc<MODE, 0, 1, true>(p, pl, n_img_t, st, who);
    idee_set_error("%s: channel configuration (%d k-steps, %d n-tiles) is not built for the bf16 path", who, pl.KS, pl.NTL);
    return 1;
}

int prep(const idee_conv_desc* d, const float* w, uint2* wfrag, const Plan& pl, int dgrad, cudaStream_t st) {
    const int64_t total = (int64_t)pl.n_oc * pl.NTf * pl.KS * pl.NTL * 32;
    int nb = (int)((total + 255) / 256);
    if (nb > 1024) nb = 1024;
    prep_weights_kernel<<<dim3(nb, d->Vw), 256, 0, st>>>(w, wfrag, d->Cin, d->Cout, pl.NTf, dgrad, pl.KS, pl.NTL, pl.n_oc,
                                                           (int64_t)d->Cin * d->Cout * pl.NTf);
    IDEE_LAUNCH_CHECK("conv3d prep_weights");
    return 0;
}

}  // namespace

// tcgen05 / TMEM path for the 96 -> 96 classifier conv (conv_umma.cu)
bool conv_umma_eligible(const idee_conv_desc* d);
size_t conv_umma_workspace_bytes();
int conv_umma_run(const idee_conv_desc* d, int dgrad, const float* in, const float* w, const float* bias, const float* relu_src,
                  float* out, void* ws, cudaStream_t st);

size_t conv_tc_fwd_workspace_bytes(const idee_conv_desc* d) {
    if (conv_umma_eligible(d)) return conv_umma_workspace_bytes();
    return make_plan(d->proj ? PROJ_FWD : CLS_FWD, d->Cin, d->Cout, d->Vw).wfrag_bytes;
}

size_t conv_tc_dgrad_workspace_bytes(const idee_conv_desc* d) {
    if (conv_umma_eligible(d)) return conv_umma_workspace_bytes();
    size_t b = make_plan(d->proj ? PROJ_DGRAD_PAD : CLS_DGRAD, d->Cout, d->Cin, d->Vw).wfrag_bytes;
    b = (b + 255) / 256 * 256;
    if (d->proj) b += sizeof(float) * (size_t)d->N * d->V * (d->Ti + 2) * (d->Hi + 2) * (d->Wi + 2) * 16;
    return b;
}

int conv_tc_fwd(const idee_conv_desc* d, const float* x, const float* w, const float* b, float* y, void* ws, cudaStream_t st) {
    if (conv_umma_eligible(d)) return conv_umma_run(d, 0, x, w, b, nullptr, y, ws, st);
    const int mode = d->proj ? PROJ_FWD : CLS_FWD;
    const Plan pl = make_plan(mode, d->Cin, d->Cout, d->Vw);
    if (prep(d, w, (uint2*)ws, pl, 0, st)) return 2;
    P p{};
    p.in = x; p.out = y; p.bias = b; p.relu_src = nullptr; p.wfrag = (const uint2*)ws;
    p.N = d->N; p.V = d->V; p.Vw = d->Vw;
    p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
    p.in_sn = d->x_sn; p.in_sv = d->x_sv; p.in_st = d->x_st; p.in_sh = d->x_sh; p.in_sw = d->x_sw; p.in_sg = d->x_sg; p.in_cpg = d->in_cpg;
    p.out_sn = d->y_sn; p.out_sv = d->y_sv; p.out_st = d->y_st; p.out_sh = d->y_sh; p.out_sw = d->y_sw; p.out_sg = d->y_sg; p.out_cpg = d->out_cpg;
    p.CO = d->Cout; p.CIr = d->Cin; p.NTf = pl.NTf; p.relu = d->relu; p.tiles_w = (p.Wo + TW - 1) / TW;
    const int nit = d->N * d->V * d->To;
    return d->proj ? dispatch_tc<PROJ_FWD>(p, pl, nit, st, "conv3d_fwd(proj,bf16)") : dispatch_tc<CLS_FWD>(p, pl, nit, st, "conv3d_fwd(cls,bf16)");
}

int conv_tc_dgrad(const idee_conv_desc* d, const float* gy, const float* w, const float* relu_src, float* gx, void* ws, cudaStream_t st) {
    if (conv_umma_eligible(d)) return conv_umma_run(d, 1, gy, w, nullptr, relu_src, gx, ws, st);
    const int mode = d->proj ? PROJ_DGRAD_PAD : CLS_DGRAD;
    const Plan pl = make_plan(mode, d->Cout, d->Cin, d->Vw);
    if (prep(d, w, (uint2*)ws, pl, 1, st)) return 2;
    P p{};
    p.in = gy; p.bias = nullptr; p.wfrag = (const uint2*)ws;
    p.N = d->N; p.V = d->V; p.Vw = d->Vw;
    p.Ti = d->To; p.Hi = d->Ho; p.Wi = d->Wo;
    p.in_sn = d->y_sn; p.in_sv = d->y_sv; p.in_st = d->y_st; p.in_sh = d->y_sh; p.in_sw = d->y_sw; p.in_sg = d->y_sg; p.in_cpg = d->out_cpg;
    p.CO = d->Cin; p.CIr = d->Cout; p.NTf = pl.NTf; p.relu = 0;
    if (!d->proj) {
        p.out = gx; p.relu_src = relu_src;
        p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
        p.out_sn = d->x_sn; p.out_sv = d->x_sv; p.out_st = d->x_st; p.out_sh = d->x_sh; p.out_sw = d->x_sw; p.out_sg = d->x_sg; p.out_cpg = d->in_cpg;
        p.tiles_w = (p.Wo + TW - 1) / TW;
        return dispatch_tc<CLS_DGRAD>(p, pl, d->N * d->V * p.To, st, "conv3d_dgrad(cls,bf16)");
    }
    // replicate conv: gradient on the padded domain, then fold the ring onto the border
    IDEE_REQUIRE(d->x_sw == 16 && d->x_sh == (int64_t)d->Wi * 16 && d->x_st == (int64_t)d->Hi * d->Wi * 16 &&
                 d->x_sv == (int64_t)d->Ti * d->Hi * d->Wi * 16 && d->x_sn == d->x_sv * d->V,
                 "conv3d_dgrad(proj,bf16): the input gradient must be a contiguous [N,V,T,H,W,16] tensor");
    float* gpad = (float*)((char*)ws + (pl.wfrag_bytes + 255) / 256 * 256);
    const int Tp = d->Ti + 2, Hp = d->Hi + 2, Wp = d->Wi + 2;
    p.out = gpad; p.relu_src = nullptr;
    p.To = Tp; p.Ho = Hp; p.Wo = Wp;
    p.out_sw = 16; p.out_sh = (int64_t)Wp * 16; p.out_st = (int64_t)Hp * Wp * 16; p.out_sv = (int64_t)Tp * Hp * Wp * 16;
    p.out_sn = p.out_sv * d->V; p.out_sg = 0; p.out_cpg = 1;
    p.tiles_w = (p.Wo + TW - 1) / TW;
    if (dispatch_tc<PROJ_DGRAD_PAD>(p, pl, d->N * d->V * Tp, st, "conv3d_dgrad(proj,bf16)")) return 2;
    const int64_t total = (int64_t)d->N * d->V * d->Ti * d->Hi * d->Wi * 4;
    int nb = (int)((total + 255) / 256);
    const int cap = idee_num_sms() * 16;
    if (nb > cap) nb = cap;
    fold_pad_kernel<<<nb, 256, 0, st>>>(gpad, gx, relu_src, d->N * d->V, d->Ti, d->Hi, d->Wi);
    IDEE_LAUNCH_CHECK("conv3d_dgrad fold");
    return 0;
}

int conv_tc_wgrad_ncout(const idee_conv_desc* d) { return d->Cout >= 32 ? 32 : (d->Cout >= 16 ? 16 : 8); }

int conv_tc_wgrad_splits(const idee_conv_desc* d) {
    const int n_ic = (d->Cin + 15) / 16, NC = conv_tc_wgrad_ncout(d);
    const int n_occ = (((d->Cout + 15) / 16) * 16 + NC - 1) / NC;
    int S = (idee_num_sms() * 4 + d->Vw * n_ic * n_occ - 1) / (d->Vw * n_ic * n_occ);
    if (S < 1) S = 1;
    if (S > 512) S = 512;
    return S;
}

size_t conv_tc_wgrad_workspace_bytes(const idee_conv_desc* d) {
    const int n_ic = (d->Cin + 15) / 16, n_oc16 = (d->Cout + 15) / 16, NT = (d->proj ? 3 : 2) * 9;
    return sizeof(float) * (size_t)d->Vw * n_ic * n_oc16 * conv_tc_wgrad_splits(d) * (NT * 256 + 16);
}

// launches the tensor-core partial kernel; the caller (conv.cu) runs the shared reduce stage with S = conv_tc_wgrad_splits
int conv_tc_wgrad_partials(const idee_conv_desc* d, const float* x, const float* gy, float* partials, cudaStream_t st) {
    WP p{};
    p.in = x; p.gout = gy; p.partials = partials;
    p.N = d->N; p.V = d->V; p.Vw = d->Vw;
    p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
    p.in_sn = d->x_sn; p.in_sv = d->x_sv; p.in_st = d->x_st; p.in_sh = d->x_sh; p.in_sw = d->x_sw; p.in_sg = d->x_sg; p.in_cpg = d->in_cpg;
    p.go_sn = d->y_sn; p.go_sv = d->y_sv; p.go_st = d->y_st; p.go_sh = d->y_sh; p.go_sw = d->y_sw;
    p.FCO = d->Cout; p.proj = d->proj;
    p.n_ic = (d->Cin + 15) / 16; p.n_oc16 = (d->Cout + 15) / 16; p.S = conv_tc_wgrad_splits(d);
    p.tiles_h = (d->Ho + TH - 1) / TH; p.tiles_w = (d->Wo + TW - 1) / TW;
    p.tiles_per_set = (int64_t)d->N * (d->Vw == 1 ? d->V : 1) * d->To * p.tiles_h * p.tiles_w;
    const int NC = conv_tc_wgrad_ncout(d);
    const int n_occ = (p.n_oc16 * 16 + NC - 1) / NC;
    // the partial buffer is only partly written when Cout is not a multiple of 16 (Cout == 1): clear it first
    if (d->Cout % 16) IDEE_CUDA(cudaMemsetAsync(partials, 0, conv_tc_wgrad_workspace_bytes(d), st), "conv3d_wgrad(bf16)");
    dim3 grid(p.S, d->Vw, p.n_ic * n_occ);
    const int KTIN = d->proj ? 3 : 2;
    const size_t smem = (size_t)KTIN * HH * HW_ * (16 * 4 + 24 * 2) + (size_t)TH * TW * (NC * 4 + (NC + 8) * 2);
#define IDEE_WGRAD_LAUNCH(NT_, NTL_)                                                                                       \
    do {                                                                                                                   \
        IDEE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<NT_, NTL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_wgrad(bf16)"); \
        wgrad_tc_kernel<NT_, NTL_><<<grid, 128, smem, st>>>(p);                                                            \
    } while (0)
    if (d->proj) { if (NC == 16) IDEE_WGRAD_LAUNCH(27, 2); else { idee_set_error("conv3d_wgrad(proj,bf16): Cout must be 16"); return 1; } }
    else if (NC == 32) IDEE_WGRAD_LAUNCH(18, 4);
    else if (NC == 16) IDEE_WGRAD_LAUNCH(18, 2);
    else IDEE_WGRAD_LAUNCH(18, 1);
#undef IDEE_WGRAD_LAUNCH
    IDEE_LAUNCH_CHECK("conv3d_wgrad(bf16)");
    return 0;
}
