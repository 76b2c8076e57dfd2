// tcgen05 / TMEM implicit-GEMM kernel for the 16 -> 16 channel 3x3x3 replicate convolution of the encoder (proj_var,
// Swin_3D.py:586-592): forward and (padded-domain) data gradient, bf16 activations in HBM, fp32 accumulate (sm_100a).
//
// A CTA owns tiles of 128 output pixels = 16 rows x 8 columns = the 128 TMEM lanes; the accumulator D[128 x 16] fp32 takes
// 16 TMEM columns and there are two of them, so the MMAs of tile i overlap the epilogue of tile i-1 and the loads of i+1.
//
// Zero-copy operands.  The halo [3 t][18 rows][10 columns] is copied by cp.async straight from HBM (bf16) into two 8-channel
// chunk planes, plane[kc][t][row][col] with 16 bytes per pixel.  Because a tile row is exactly 8 pixels, the 8 x 16-byte
// core matrices of the canonical K-major SWIZZLE_NONE layout are the tile rows themselves: for tap (kt,kh,kw) the A operand
// [128 pixels x 16 channels] is the descriptor {start = plane + ((kt*18+kh)*10+kw)*16, SBO = 160 B (next tile row),
// LBO = plane size (next 8-channel chunk)} -- no per-tap copies, no shifted duplicates, no ldmatrix, no per-warp HMMA.
// One elected thread issues the 27 tcgen05.mma.cta_group::1.kind::f16 (M=128, N=16, K=16) of a tile and one tcgen05.commit;
// every warp then reads its 32 lanes x 16 columns with tcgen05.ld (one pixel with all 16 channels per thread), adds the
// bias / ReLU and stores 32 (bf16) or 64 (fp32) contiguous bytes.  Replicate padding is resolved by the loader (clamped
// addresses), zero padding of the data gradient by cp.async zero-fill.
// Descriptor and instruction-descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (see conv_umma.cu).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "idee_b200.h"

namespace conv16u {

constexpr int TR = 16, TC = 8, HR = TR + 2, HC = TC + 2, KTIN = 3, NTAP = 27, NTH = 128;
constexpr int NPX = KTIN * HR * HC;             // 540 halo pixels
constexpr int CHUNK = NPX * 16;                 // bytes of one 8-channel chunk plane
constexpr int HALO = 2 * CHUNK;                 // one halo buffer
constexpr int B_TAP = 512, B_BYTES = NTAP * B_TAP;
constexpr int TMEM_COLS = 32;
#ifndef IDEE_U16_NBUF
#define IDEE_U16_NBUF 2
#endif
constexpr int NBUF = IDEE_U16_NBUF;             // halo buffers / accumulators per CTA (1: overlap comes from co-resident CTAs)
// U16_DGRAD_T: data gradient on the domain [T][H+2][W+2] (time unpadded: the replicate adjoint along t is 9 extra MMAs on the
// first / last slice, see the kernel; h / w padded).  Pixels that need no folding along h / w are final: ReLU mask, bf16, straight
// into gx.  Only the two-pixel ring around the image goes to the fp32 padded buffer for fold_ring_kernel (conv_tc.cu).
enum { U16_FWD = 0, U16_DGRAD_PAD = 1, U16_DGRAD_T = 2 };


struct UP16 {
    const __nv_bfloat16* in; void* out; const float* bias; const __nv_bfloat16* wB;   // wB: [wset][tap][512 B]
    int V, Vw, Ti, Hi, Wi, To, Ho, Wo, relu;
    int64_t in_sn, in_sv, out_sn, out_sv;
    int in_st, in_sh, in_sw, out_st, out_sh, out_sw;
    uint32_t total_tiles;
    FastDiv fd_tw, fd_th, fd_to, fd_v;
    __nv_bfloat16* gx; const __nv_bfloat16* relu_src;      // U16_DGRAD_T: contiguous [N,V,T,H,W,16] gradient / ReLU source (or null)
    int H, W;                                              // U16_DGRAD_T: unpadded image extent (Ho = H + 2, Wo = W + 2)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 16, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void cp_async16(uint32_t smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// fp32 reference-layout weights [Vw][16][16][27] -> bf16 canonical K-major B tiles per gather tap j = (kt*3+kh)*3+kw:
//   wB[wset][j][kc][ng][r][e] = B(n = ng*8 + r, k = kc*8 + e);  forward: B(n,k) = W[n][k][j];  dgrad: B(n,k) = W[k][n][26 - j]
__global__ void prep_umma16_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wB, int Vw, int dgrad) {
    const int total = Vw * NTAP * 256;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int el = e & 7, r = (e >> 3) & 7, ng = (e >> 6) & 1, kc = (e >> 7) & 1;
        const int j = (e >> 8) % NTAP, ws = (e >> 8) / NTAP;
        const int n = ng * 8 + r, k = kc * 8 + el;
        const int fo = dgrad ? k : n, fc = dgrad ? n : k, ft = dgrad ? 26 - j : j;
        wB[e] = __float2bfloat16(w[(((int64_t)ws * 16 + fo) * 16 + fc) * NTAP + ft]);
    }
}

template <int MODE, bool OUT16>
__global__ void __launch_bounds__(NTH)
conv16_umma_kernel(UP16 p) {
    constexpr int OT = MODE == U16_DGRAD_PAD ? -2 : -1, OHW = MODE == U16_FWD ? -1 : -2;   // halo origin relative to the tile
    constexpr int TOTAL = NPX * 2, NEL = (TOTAL + NTH - 1) / NTH;      // 16-byte elements: (pixel, chunk), chunk fastest
    constexpr int PLANE = HR * HC * 2;                                 // elements per input time slice
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* halo = smem_raw;                                    // [2 buffers][2 chunks][NPX][16 B]
    unsigned char* Bw = smem_raw + NBUF * HALO;                        // [27 taps][512 B]
    float* bias_s = reinterpret_cast<float*>(Bw + B_BYTES);            // [2 tile parities][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 32);         // [2] MMAs of accumulator a complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // per-thread loader constants (see conv_tc16_kernel): element tid + i*NTH = (pixel q = tid/2 + 64 i, chunk tid & 1)
    int rel_src[NEL], hw[NEL];
#pragma unroll
    for (int i = 0; i < NEL; ++i) {
        const int q = (tid >> 1) + 64 * i, kt = q / (HR * HC), rem = q - kt * (HR * HC), hh = rem / HC, px = rem - hh * HC;
        rel_src[i] = kt * p.in_st + hh * p.in_sh + px * p.in_sw + (tid & 1) * 8;
        hw[i] = hh | (px << 5);
    }
    const int csub = (tid & 1) * 8;
    const uint32_t dst_base = smem_u32(halo) + (tid & 1) * CHUNK + (tid >> 1) * 16;
    constexpr int DSTEP = 64 * 16;

    struct Tile { int n, v, t, h0, w0; };
    auto decode = [&](uint32_t tile) {
        Tile c;
        uint32_t q, r;
        // t fastest: consecutive tiles of a CTA share two of their three input time slices (L2 hits instead of DRAM re-reads)
        p.fd_to.divmod(tile, q, r); c.t = (int)r;
        p.fd_tw.divmod(q, q, r); c.w0 = (int)r * TC;
        p.fd_th.divmod(q, q, r); c.h0 = (int)r * TR;
        p.fd_v.divmod(q, q, r); c.v = (int)r; c.n = (int)q;
        return c;
    };
    auto issue = [&](const Tile& c, int buf) {
        const __nv_bfloat16* in_img = p.in + c.n * p.in_sn + c.v * p.in_sv;
        const int t_lo = c.t + OT, h_lo = c.h0 + OHW, w_lo = c.w0 + OHW;
        const uint32_t dst0 = dst_base + buf * HALO;
        int dT[KTIN]; bool okT[KTIN];
#pragma unroll
        for (int k = 0; k < KTIN; ++k) {
            const int ti = t_lo + k;
            if (MODE == U16_FWD) { dT[k] = (min(max(ti, 0), p.Ti - 1) - ti) * p.in_st; okT[k] = true; }
            else { okT[k] = (unsigned)ti < (unsigned)p.Ti; dT[k] = okT[k] ? 0 : -(ti * p.in_st); }
        }
        if (h_lo >= 0 && h_lo + HR - 1 < p.Hi && w_lo >= 0 && w_lo + HC - 1 < p.Wi) {
            const __nv_bfloat16* base = in_img + (int64_t)t_lo * p.in_st + h_lo * p.in_sh + w_lo * p.in_sw;
#pragma unroll
            for (int i = 0; i < NEL; ++i) {
                const int k0 = (i * NTH) / PLANE, k1 = (i * NTH + NTH - 1) / PLANE;
                if (tid + i * NTH < TOTAL) {
                    int d = dT[k0]; bool ok = okT[k0];
                    if (k0 != k1 && k1 < KTIN) { const bool up = tid >= k1 * PLANE - i * NTH; d = up ? dT[k1 < KTIN ? k1 : k0] : d; ok = up ? okT[k1 < KTIN ? k1 : k0] : ok; }
                    cp_async16(dst0 + i * DSTEP, base + (rel_src[i] + d), ok ? 16 : 0);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NEL; ++i) {
                const int k0 = (i * NTH) / PLANE, k1 = (i * NTH + NTH - 1) / PLANE;
                if (tid + i * NTH < TOTAL) {
                    int kt = k0;
                    if (k0 != k1 && k1 < KTIN) kt = tid >= k1 * PLANE - i * NTH ? k1 : k0;
                    int ti = t_lo + kt, hi = h_lo + (hw[i] & 31), wi = w_lo + (hw[i] >> 5);
                    bool ok = true;
                    if (MODE == U16_FWD) { ti = min(max(ti, 0), p.Ti - 1); hi = min(max(hi, 0), p.Hi - 1); wi = min(max(wi, 0), p.Wi - 1); }
                    else ok = (unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                    const int off = ok ? ti * p.in_st + hi * p.in_sh + wi * p.in_sw + csub : 0;
                    cp_async16(dst0 + i * DSTEP, in_img + off, ok ? 16 : 0);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // epilogue of one finished tile: TMEM lane = pixel (row = lane / 8 of the warp's 4 tile rows), 16 columns = channels
    auto epilogue = [&](const Tile& c, int acc) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 16, v);
        const int r = warp * 4 + (lane >> 3), cc = lane & 7;
        if (MODE == U16_DGRAD_T) {
            const int h = c.h0 + r - 1, w = c.w0 + cc - 1;             // unpadded pixel of this padded position
            if (h >= 1 && h <= p.H - 2 && w >= 1 && w <= p.W - 2) {    // no folding along h / w: final value
                const int64_t o = ((((int64_t)c.n * p.V + c.v) * p.To + c.t) * p.H + h) * (int64_t)p.W * 16 + w * 16;
                if (p.relu_src) {
                    uint32_t m[8];
                    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]),
                                 "=r"(m[4]), "=r"(m[5]), "=r"(m[6]), "=r"(m[7]) : "l"(p.relu_src + o));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!(__uint_as_float(m[i] << 16) > 0.f)) v[2 * i] = 0.f;
                        if (!(__uint_as_float(m[i] & 0xFFFF0000u) > 0.f)) v[2 * i + 1] = 0.f;
                    }
                }
                const uint32_t pk[8] = {pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]),
                                        pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15])};
                st8u(p.gx + o, pk);
            } else if (c.h0 + r < p.Ho && c.w0 + cc < p.Wo) {          // ring: fp32 into the padded buffer, folded later
                const int64_t o = c.n * p.out_sn + c.v * p.out_sv + (int64_t)(c.t * p.out_st + (c.h0 + r) * p.out_sh + (c.w0 + cc) * p.out_sw);
                float* dst = reinterpret_cast<float*>(p.out) + o;
                st8f(dst, v); st8f(dst + 8, v + 8);
            }
        } else
        if (c.h0 + r < p.Ho && c.w0 + cc < p.Wo) {
            const int64_t o = c.n * p.out_sn + c.v * p.out_sv + (int64_t)(c.t * p.out_st + (c.h0 + r) * p.out_sh + (c.w0 + cc) * p.out_sw);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(bias_s + acc * 16 + i);
                v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
            }
            if (p.relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (OUT16) {                               // 16 bf16 = one 32-byte sector in a single 256-bit store
                const uint32_t pk[8] = {pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]),
                                        pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15])};
                st8u(reinterpret_cast<__nv_bfloat16*>(p.out) + o, pk);
            } else {
                float* dst = reinterpret_cast<float*>(p.out) + o;
                st8f(dst, v); st8f(dst + 8, v + 8);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    // contiguous tile range per CTA: one or two weight sets per CTA, shared halo rows of neighbouring tiles hit L2
    const uint32_t per_cta = (p.total_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t first = blockIdx.x * per_cta, last = min(p.total_tiles, first + per_cta);
    Tile cur{}, prev{}, nxt{};
    if (first < last) { nxt = decode(first); issue(nxt, 0); }
    int cur_wset = -1;
    uint32_t it = 0;
    for (uint32_t tile = first; tile < last; ++tile, ++it) {
        prev = cur; cur = nxt;
        const int buf = NBUF == 2 ? (it & 1) : 0;
        const int wset = p.Vw == 1 ? 0 : cur.v;
        bool prev_waited = false;
        if (wset != cur_wset) {                           // (rare) new weight set: the previous tile's MMAs still read the old one
            if (NBUF == 2 && it > 0) { mbar_wait(smem_u32(&bars[(it - 1) & 1]), ((it - 1) >> 1) & 1u); prev_waited = true; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncthreads();
            const uint4* src = reinterpret_cast<const uint4*>(p.wB + (size_t)wset * NTAP * 256);
            for (int e = tid; e < B_BYTES / 16; e += NTH) reinterpret_cast<uint4*>(Bw)[e] = __ldg(src + e);
            cur_wset = wset;
        }
        if (tid < 16) bias_s[buf * 16 + tid] = p.bias ? p.bias[wset * 16 + tid] : 0.f;   // read by this tile's (deferred) epilogue
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // halo (and weights) -> visible to the tensor core
        __syncthreads();                                                  // ... and every warp has drained accumulator `buf`
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors differ from tap to tap only in the 14-bit start-address field: one add on the low word each
            const uint64_t adesc0 = make_desc(smem_u32(halo) + buf * HALO, CHUNK, HC * 16);
            const uint64_t bdesc0 = make_desc(smem_u32(Bw), 256, 128);
            const uint32_t dcol = tmem_base + buf * 16;
#pragma unroll
            for (int j = 0; j < NTAP; ++j) {
                const int kt = j / 9, kh = (j / 3) % 3, kw = j % 3;
                umma_bf16(dcol, adesc0 + (uint64_t)((kt * HR + kh) * HC + kw), bdesc0 + (uint64_t)(j * (B_TAP / 16)), j > 0 ? 1u : 0u);
            }
            if (MODE == U16_DGRAD_T) {
                // adjoint of the replicate padding along t: slice 0 also receives gy[0] W[kt = 0] (the clamped read of x[-1]) and
                // slice T-1 gy[T-1] W[kt = 2]: the centre plane of the halo against the weight tiles of the far / near time tap
                const bool lo = cur.t == 0, hi = cur.t == p.To - 1;
                if (lo || hi) {
#pragma unroll
                    for (int j9 = 0; j9 < 9; ++j9) {
                        const uint64_t ad = adesc0 + (uint64_t)((HR + j9 / 3) * HC + j9 % 3);
                        if (lo) umma_bf16(dcol, ad, bdesc0 + (uint64_t)((18 + j9) * (B_TAP / 16)), 1u);
                        if (hi) umma_bf16(dcol, ad, bdesc0 + (uint64_t)(j9 * (B_TAP / 16)), 1u);
                    }
                }
            }
            umma_commit(smem_u32(&bars[buf]));
        }
        if (NBUF == 2) {
            // while the tensor core works on this tile: retire the previous one and request the next
            if (it > 0 && !prev_waited) mbar_wait(smem_u32(&bars[(it - 1) & 1]), ((it - 1) >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tile + 1 < last) { nxt = decode(tile + 1); issue(nxt, buf ^ 1); }   // buffer of tile it-1: its MMAs are done
            if (it > 0) epilogue(prev, (it - 1) & 1);
        } else {
            // one buffer: wait for this tile's MMAs, request the next tile into the freed halo, then drain the accumulator
            mbar_wait(smem_u32(&bars[0]), it & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tile + 1 < last) { nxt = decode(tile + 1); issue(nxt, 0); }
            epilogue(cur, 0);
        }
    }
    if (NBUF == 2 && it > 0) {
        mbar_wait(smem_u32(&bars[(it - 1) & 1]), ((it - 1) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        epilogue(cur, (it - 1) & 1);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace conv16u

using namespace conv16u;

size_t conv16_umma_workspace_bytes(int Vw) { return sizeof(__nv_bfloat16) * (size_t)Vw * NTAP * 256; }

// mode 0: y = conv(x) (+bias, ReLU);  mode 1: padded-domain data gradient (out = [N,V,T+2,H+2,W+2,16] fp32, no bias);
// mode 2: data gradient with final pixels written to gx (bf16, optional ReLU mask) and the h / w ring to out = [N,V,T,H+2,W+2,16] fp32
int conv16_umma_run(int mode, int out16, const void* in, const float* w, const float* bias, void* out, void* ws, int N, int V, int Vw,
                    int Ti, int Hi, int Wi, int To, int Ho, int Wo, const int64_t* in_s, const int64_t* out_s, int relu,
                    cudaStream_t st, void* gx, const void* relu_src) {
    __nv_bfloat16* wB = (__nv_bfloat16*)ws;
    prep_umma16_weights_kernel<<<(Vw * NTAP * 256 + 255) / 256, 256, 0, st>>>(w, wB, Vw, mode != U16_FWD);
    IDEE_LAUNCH_CHECK("conv3d(umma16) prep");
    UP16 p{};
    p.in = (const __nv_bfloat16*)in; p.out = out; p.bias = bias; p.wB = wB;
    p.V = V; p.Vw = Vw; p.Ti = Ti; p.Hi = Hi; p.Wi = Wi; p.To = To; p.Ho = Ho; p.Wo = Wo; p.relu = relu;
    p.gx = (__nv_bfloat16*)gx; p.relu_src = (const __nv_bfloat16*)relu_src; p.H = Ho - 2; p.W = Wo - 2;
    p.in_sn = in_s[0]; p.in_sv = in_s[1]; p.in_st = (int)in_s[2]; p.in_sh = (int)in_s[3]; p.in_sw = (int)in_s[4];
    p.out_sn = out_s[0]; p.out_sv = out_s[1]; p.out_st = (int)out_s[2]; p.out_sh = (int)out_s[3]; p.out_sw = (int)out_s[4];
    IDEE_REQUIRE(in_s[4] == 16 && out_s[4] == 16, "conv3d(umma16): pixels must hold 16 contiguous channels");
    IDEE_REQUIRE((int64_t)(Ti + 3) * in_s[2] + (int64_t)(Hi + HR) * in_s[3] + (int64_t)(Wi + HC) * in_s[4] < (1ll << 31) &&
                 (int64_t)(To + 1) * out_s[2] + (int64_t)(Ho + TR) * out_s[3] + (int64_t)(Wo + TC) * out_s[4] < (1ll << 31),
                 "conv3d(umma16): tensor too large for 32-bit image-relative offsets");
    const int tiles_w = (Wo + TC - 1) / TC, tiles_h = (Ho + TR - 1) / TR;
    const int64_t total = (int64_t)N * V * To * tiles_h * tiles_w;
    IDEE_REQUIRE(total < (1ll << 31), "conv3d(umma16): too many tiles");
    p.total_tiles = (uint32_t)total;
    p.fd_tw = make_fastdiv(tiles_w); p.fd_th = make_fastdiv(tiles_h); p.fd_to = make_fastdiv(To); p.fd_v = make_fastdiv(V);
    const size_t smem = (size_t)NBUF * HALO + B_BYTES + 128 + 16 + 16;
#define IDEE_U16_LAUNCH(M_, O_)                                                                                              \
    do {                                                                                                                     \
        auto kern = conv16_umma_kernel<M_, O_>;                                                                              \
        IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d(umma16)");     \
        /* the occupancy API reports 1 CTA/SM for kernels that allocate TMEM; CTAs that each take 32 of the 512 columns do   \
           co-reside, so the resident count is derived from shared memory, registers and TMEM columns directly */            \
        cudaFuncAttributes fa;                                                                                               \
        IDEE_CUDA(cudaFuncGetAttributes(&fa, kern), "conv3d(umma16)");                                                       \
        int dev = 0, smem_sm = 0, regs_sm = 0;                                                                               \
        IDEE_CUDA(cudaGetDevice(&dev), "conv3d(umma16)");                                                                    \
        IDEE_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev), "conv3d(umma16)");     \
        IDEE_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev), "conv3d(umma16)");        \
        int per_sm = (int)(smem_sm / (smem + fa.sharedSizeBytes + 1024));                                                    \
        const int by_regs = regs_sm / (((fa.numRegs + 7) / 8 * 8) * NTH);                                                    \
        if (per_sm > by_regs) per_sm = by_regs;                                                                              \
        if (getenv("IDEE_B200_DEBUG")) fprintf(stderr, "conv16_umma: %d CTAs/SM (smem %zu, %d regs)\n", per_sm, smem, fa.numRegs); \
        if (per_sm < 1) per_sm = 1;                                                                                          \
        if (per_sm > 512 / TMEM_COLS) per_sm = 512 / TMEM_COLS;                                                              \
        int64_t grid = (int64_t)idee_num_sms() * per_sm;                                                                     \
        if (grid > total) grid = total;                                                                                      \
        kern<<<(unsigned)grid, NTH, smem, st>>>(p);                                                                          \
    } while (0)
    if (mode == U16_FWD) { if (out16) IDEE_U16_LAUNCH(U16_FWD, true); else IDEE_U16_LAUNCH(U16_FWD, false); }
    else if (mode == U16_DGRAD_T) IDEE_U16_LAUNCH(U16_DGRAD_T, false);
    else IDEE_U16_LAUNCH(U16_DGRAD_PAD, false);
#undef IDEE_U16_LAUNCH
    IDEE_LAUNCH_CHECK("conv3d(umma16)");
    return 0;
}
