// Warp-specialised tcgen05 / TMEM implicit-GEMM kernel for the dense contractions of the joint classifier head: the 96 -> 96
// convolution Conv3d(96, 96, (2,3,3), stride (2,1,1), pad (0,1,1)) (classifier/CNN_3D.py:84), forward and data gradient, and the
// head's first convolution on the 16-channel plane image of the rank-1 form of z_q (16 -> 96 forward, 96 -> 16 data gradient;
// template parameters GI / GO = gather-input / output channels; their 18 weight taps stay resident in shared memory).
//
// Roles (416 threads):  warps 0-7 load halos (fp32 HBM -> bf16 shared memory) up to two loads ahead; warps 8-11 drain the
// accumulator of tile i (tcgen05.ld -> bias / ReLU / mask -> fp32 stores; warp w owns TMEM lanes 32 (w & 3)..); warp 12 streams
// the per-tap weight matrices through a cp.async ring (or keeps all taps resident) and issues the tcgen05.mma of tile i.
// Two halo buffers and two TMEM accumulators (2 x 96 columns) let the three stages run concurrently; all hand-offs are mbarriers (halo full, accumulator done via tcgen05.commit, accumulator drained, weight slot
// free) -- there is no CTA-wide barrier in the steady state.
//
// Zero-copy operands (same construction as conv16_umma.cu): a tile is 16 rows x 8 columns of output pixels = 128 TMEM lanes,
// the halo [kt][18][10] is stored as twelve 8-channel chunk planes with 16 bytes per pixel, so a tile row is one 8 x 16-byte
// core matrix of the canonical K-major SWIZZLE_NONE layout and tap (kt,kh,kw), k-step ks is the descriptor
// {start = plane(2 ks) + ((kt*18+kh)*10+kw)*16, LBO = plane stride, SBO = 160 B}.  K = taps x 96 = 1728 (forward) or 864 (data
// gradient, taps of the output time parity) is accumulated in TMEM by 108 / 54 MMAs of shape M=128, N=96, K=16.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace conv96u {

constexpr int TR = 16, TC = 8, HR = TR + 2, HC = TC + 2;
constexpr int NB_MAX = 8;                              // weight-tap ring depth (96 -> 96 only): 4 (forward, two-slice halos) or 8 (data gradient)
constexpr int B_SBO = 128;
constexpr int NLOAD = 256;                             // loader threads (warps 0-7)
constexpr int NEPI = 128;                              // epilogue threads (warps 8-11: warp & 3 = TMEM lane quadrant)
constexpr int MMA_WARP = (NLOAD + NEPI) / 32, NTHREADS = NLOAD + NEPI + 32;     // warp 12 = weights + MMA issue
enum { U_FWD = 0, U_DGRAD = 1 };

struct UP {
    const float* in; float* out; const float* bias; const float* relu_src; const __nv_bfloat16* wB;
    int N, Ti, Hi, Wi, To, Ho, Wo;
    int64_t in_sn, out_sn;
    int in_st, in_sh, in_sw, out_st, out_sh, out_sw;
    int relu, tiles_w, tiles_h;
    uint32_t total_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_n(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    while (!ok) {
        __nanosleep(32);
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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 reference-layout weights [Co][Ci][18] -> bf16 canonical K-major tiles, one per forward tap (GI = k extent, GO = n extent):
//   wB[ft][kc][ng][r][e] = B(n = ng*8 + r, k = kc*8 + e);  forward: B(n,k) = W[n][k][ft];  dgrad: B(n,k) = W[k][n][ft]
__global__ void prep96_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wB, int dgrad, int GI, int GO) {
    const int total = 18 * GO * GI, KC = GI / 8, ci_ref = dgrad ? GO : GI;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int el = e & 7, r = (e >> 3) & 7;
        int q = e >> 6;
        const int ng = q % (GO / 8); q /= (GO / 8);
        const int kc = q % KC;
        const int ft = q / KC;
        const int n = ng * 8 + r, k = kc * 8 + el;
        const int fo = dgrad ? k : n, fc = dgrad ? n : k;
        wB[e] = __float2bfloat16(w[((int64_t)fo * ci_ref + fc) * 18 + ft]);
    }
}

// GI: gather-input channels (K per tap), GO: output channels (N).  RESIDENT (GI * GO small): all 18 weight taps are copied to
// shared memory once per CTA; otherwise they stream through an NB-deep cp.async ring.
template <int MODE, int GI, int GO>
__global__ void __launch_bounds__(NTHREADS, 1)
conv96_umma_kernel(UP p) {
    constexpr int CI = GI, CO = GO, KC = CI / 8;
    constexpr int B_TAP = CO * CI * 2;                     // bytes per tap, canonical [kc][ng][8][16 B]
    constexpr int B_LBO = (CO / 8) * 128;
    constexpr bool RESIDENT = 18 * B_TAP <= 64 * 1024;
    constexpr int NB = MODE == U_DGRAD ? 8 : 4;            // ring depth: the one-slice halos of the data gradient leave room for 8 taps in flight
    constexpr int B_SLOTS = RESIDENT ? 18 : NB;
    constexpr int TMEM_COLS = 2 * CO > 128 ? 256 : (2 * CO > 64 ? 128 : (2 * CO > 32 ? 64 : 32));
    constexpr uint32_t IDESC = idesc_n(CO);
    constexpr int KTIN = MODE == U_FWD ? 2 : 1, NT = MODE == U_FWD ? 18 : 9;
    constexpr int NPX = KTIN * HR * HC;                   // halo pixels
    constexpr int CHUNK = NPX * 16 + 16;                  // plane stride (+16: a pixel's 12 chunks land in different banks)
    constexpr int HALO = KC * CHUNK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* halo = smem_raw;                                        // [2][KC][NPX][16 B]
    unsigned char* Bring = smem_raw + 2 * ((HALO + 127) / 128 * 128);      // [B_SLOTS][B_TAP]
    float* bias_s = reinterpret_cast<float*>(Bring + B_SLOTS * B_TAP);     // [CO]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + CO);             // full[2] | accdone[2] | accfree[2] | bslot[NB] | halofree[2] | wfull[NB]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * NB_MAX);
    constexpr int HALO_PAD = (HALO + 127) / 128 * 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[2]), bar_free = smem_u32(&bars[4]), bar_slot = smem_u32(&bars[6]),
                   bar_hfree = smem_u32(&bars[6 + NB_MAX]), bar_wfull = smem_u32(&bars[8 + NB_MAX]);

    if (tid < CO) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, NLOAD); mbar_init(bar_done + 8 * i, 1); mbar_init(bar_free + 8 * i, NEPI); mbar_init(bar_hfree + 8 * i, 1);
        }
        for (int i = 0; i < NB; ++i) { mbar_init(bar_slot + 8 * i, 1); mbar_init(bar_wfull + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // contiguous tile range of this CTA
    // data gradient with an even number of output slices: tiles 2k and 2k+1 of a CTA (t fastest) are the two time parities of one
    // (n, t >> 1, h0, w0) and read the SAME halo -> one halo load per pair (pshift = 1), CTA ranges start on even tiles
    const uint32_t pshift = (MODE == U_DGRAD && (p.To & 1) == 0) ? 1u : 0u;
    uint32_t per_cta = (p.total_tiles + gridDim.x - 1) / gridDim.x;
    if (pshift) per_cta = (per_cta + 1) & ~1u;
    const uint32_t first = min(p.total_tiles, blockIdx.x * per_cta), last = min(p.total_tiles, first + per_cta);
    const uint32_t ntile = last > first ? last - first : 0;
    struct Tile { int n, t, h0, w0; };
    auto decode = [&](uint32_t tile) {
        Tile c;
        uint32_t r = tile;
        c.t = (int)(r % (uint32_t)p.To); r /= (uint32_t)p.To;      // t fastest: the forward kernel's tiles at t, t+1 are disjoint in
        c.w0 = (int)(r % (uint32_t)p.tiles_w) * TC; r /= (uint32_t)p.tiles_w;   // input, the data gradient's share their slice
        c.h0 = (int)(r % (uint32_t)p.tiles_h) * TR; c.n = (int)(r / (uint32_t)p.tiles_h);
        return c;
    };

    if (warp == MMA_WARP) {
        // ================= weights (resident or ring) + MMA issue =================
        auto tile_ft = [&](uint32_t it, int j) -> int {                 // forward-tap index of tap j of this CTA's tile it
            if (MODE == U_FWD) return j;
            const Tile c = decode(first + it);
            return (c.t & 1) * 9 + (2 - j / 3) * 3 + (2 - j % 3);
        };
        if (RESIDENT) {
            const uint4* src = reinterpret_cast<const uint4*>(p.wB);
            for (int i = lane; i < 18 * B_TAP / 16; i += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(Bring) + i * 16), "l"(src + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            for (uint32_t it = 0; it < ntile; ++it) {
                const uint32_t b = it & 1;
                const uint32_t q = it >> pshift, hb = q & 1;              // halo load index / buffer of this tile
                mbar_wait(bar_full + 8 * hb, (q >> 1) & 1u);
                if (it >= 2) mbar_wait(bar_free + 8 * b, ((it >> 1) - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t adesc0 = make_desc(smem_u32(halo) + hb * HALO_PAD, CHUNK, HC * 16);
                    const uint32_t dcol = tmem_base + b * CO;
#pragma unroll 1
                    for (int j = 0; j < NT; ++j) {
                        const int kt = MODE == U_FWD ? j / 9 : 0, kh = (j / 3) % 3, kw = j % 3;
                        const uint64_t ad = adesc0 + (uint64_t)((kt * HR + kh) * HC + kw);
                        const uint64_t bd = make_desc(smem_u32(Bring) + tile_ft(it, j) * B_TAP, B_LBO, B_SBO);
#pragma unroll
                        for (int ks = 0; ks < CI / 16; ++ks)
                            umma_bf16(dcol, ad + (uint64_t)(ks * 2 * (CHUNK / 16)), bd + (uint64_t)(ks * 2 * (B_LBO / 16)), IDESC, (j > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_done + 8 * b);
                    if (((it + 1) & pshift) == 0 || it + 1 == ntile) umma_commit(bar_hfree + 8 * hb);   // last tile that reads this halo
                }
                __syncwarp();
            }
        } else {
        const uint32_t total_taps = ntile * NT;
        auto tap_ft = [&](uint32_t g) -> int {                          // forward-tap index of global tap g
            const uint32_t it = g / NT;
            return tile_ft(it, (int)(g - it * NT));
        };
        // one TMA bulk copy per weight tap (18 KB), completion on the slot's mbarrier
        auto load_B = [&](uint32_t g) {
            if (lane == 0) {
                const uint32_t bar = bar_wfull + 8 * (g % NB);
                mbar_expect_tx(bar, B_TAP);
                bulk_g2s(smem_u32(Bring) + (g % NB) * B_TAP, p.wB + (size_t)tap_ft(g) * CO * CI, B_TAP, bar);
            }
        };
        for (uint32_t g = 0; g < NB; ++g)                                // fill the ring
            if (g < total_taps) load_B(g);
        uint32_t g = 0;
        for (uint32_t it = 0; it < ntile; ++it) {
            const uint32_t b = it & 1;
            const uint32_t q = it >> pshift, hb = q & 1;                                   // halo load index / buffer of this tile
            mbar_wait(bar_full + 8 * hb, (q >> 1) & 1u);                                   // halo of this tile is in shared memory
            if (it >= 2) mbar_wait(bar_free + 8 * b, ((it >> 1) - 1) & 1u);                // accumulator b has been drained
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc0 = make_desc(smem_u32(halo) + hb * HALO_PAD, CHUNK, HC * 16);
            const uint32_t dcol = tmem_base + b * CO;
#pragma unroll 1
            for (int j = 0; j < NT; ++j, ++g) {
                if (lane == 0) {
                    mbar_wait(bar_wfull + 8 * (g % NB), (g / NB) & 1u);                     // tap g has landed in its slot
                    const int kt = MODE == U_FWD ? j / 9 : 0, kh = (j / 3) % 3, kw = j % 3;
                    const uint64_t ad = adesc0 + (uint64_t)((kt * HR + kh) * HC + kw);
                    const uint64_t bd = make_desc(smem_u32(Bring) + (g % NB) * B_TAP, B_LBO, B_SBO);
#pragma unroll
                    for (int ks = 0; ks < CI / 16; ++ks)
                        umma_bf16(dcol, ad + (uint64_t)(ks * 2 * (CHUNK / 16)), bd + (uint64_t)(ks * 2 * (B_LBO / 16)), IDESC, (j > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(bar_slot + 8 * (g % NB));                                   // slot is free when these MMAs have read it
                    if (j == NT - 1) {
                        umma_commit(bar_done + 8 * b);                                      // ... and the accumulator is complete
                        if (((it + 1) & pshift) == 0 || it + 1 == ntile) umma_commit(bar_hfree + 8 * hb);   // last reader of this halo
                    }
                }
                __syncwarp();
                // refill the slot released by tap g-1 (its MMAs precede tap g's in the pipe, so this wait overlaps tap g)
                if (g >= 1) {
                    const uint32_t gp = g - 1;
                    if (gp + NB < total_taps) {
                        if (lane == 0) mbar_wait(bar_slot + 8 * (gp % NB), (gp / NB) & 1u);
                        load_B(gp + NB);
                    }
                }
            }
        }
        }
    } else if (warp < NLOAD / 32) {
        // ================= halo loads: load q = tiles (q << pshift).., buffer q & 1, as soon as the MMAs of load q-2 are done =================
        auto load_halo = [&](const Tile& c, int buf) {
            unsigned char* dst = halo + buf * HALO_PAD;
            const float* in_n = p.in + c.n * p.in_sn;
            constexpr int V4 = CI / 4, TOTAL = NPX * V4;                   // float4 units: (pixel, c4), c4 fastest
            constexpr int DQ = NLOAD / V4, DC = NLOAD - DQ * V4;
            int q = tid / V4, c4 = tid - q * V4;
            constexpr int DEPTH = 12;
            for (int e0 = tid; e0 < TOTAL; e0 += NLOAD * DEPTH) {
                const float* src[DEPTH]; int dsto[DEPTH];
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) {
                    src[u] = nullptr; dsto[u] = -1;
                    if (e0 + u * NLOAD < TOTAL) {
                        const int kt = q / (HR * HC), rem = q - kt * (HR * HC), hh = rem / HC, ww = rem - hh * HC;
                        const int ti = MODE == U_FWD ? 2 * c.t + kt : (c.t >> 1);
                        const int hi = c.h0 + hh - 1, wi = c.w0 + ww - 1;
                        const bool ok = ti < p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                        dsto[u] = (c4 >> 1) * CHUNK + q * 16 + (c4 & 1) * 8;
                        if (ok) src[u] = in_n + (int64_t)(ti * p.in_st + hi * p.in_sh + wi * p.in_sw) + c4 * 4;
                        q += DQ; c4 += DC;
                        if (c4 >= V4) { c4 -= V4; ++q; }
                    }
                }
                float4 f[DEPTH];
#pragma unroll
                for (int u = 0; u < DEPTH; ++u) f[u] = src[u] ? ldg4(src[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < DEPTH; ++u)
                    if (dsto[u] >= 0) *reinterpret_cast<uint2*>(dst + dsto[u]) = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            mbar_arrive(bar_full + 8 * buf);
        };
        const uint32_t nload = (ntile + pshift) >> pshift;
        for (uint32_t q = 0; q < nload; ++q) {
            if (q >= 2) mbar_wait(bar_hfree + 8 * (q & 1), ((q >> 1) - 1) & 1u);     // the MMAs that read this buffer (load q-2) are done
            load_halo(decode(first + (q << pshift)), (int)(q & 1));
        }
    } else {
        // ================= epilogue: warp w drains TMEM lanes 32 (w & 3).. (its 4 tile rows) x all CO columns =================
        // For the data gradient the ReLU mask of the pixel's channels is fetched as a bit mask BEFORE waiting for the accumulator,
        // so its latency hides behind the MMAs.
        const int quad = warp & 3;
        auto out_offset = [&](const Tile& c, bool& pix_ok) -> int64_t {
            const int r = quad * 4 + (lane >> 3), cc = lane & 7;
            const int h = c.h0 + r, w = c.w0 + cc;
            pix_ok = h < p.Ho && w < p.Wo;
            return c.n * p.out_sn + (int64_t)(c.t * p.out_st + h * p.out_sh + w * p.out_sw);
        };
        constexpr int NW = (CO + 31) / 32;
        auto relu_bits = [&](const Tile& c, uint32_t (&bits)[NW]) {
#pragma unroll
            for (int k = 0; k < NW; ++k) bits[k] = 0xFFFFFFFFu;
            if (!p.relu_src) return;
            bool pix_ok;
            const int64_t o = out_offset(c, pix_ok);
            if (!pix_ok) return;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                constexpr int full = 32;
                const int nch = CO - 32 * k < full ? CO - 32 * k : full;
                float a[32];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (8 * i < nch) ldg8f(a + 8 * i, p.relu_src + o + 32 * k + 8 * i);
                uint32_t b0 = 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nch) b0 |= (a[i] > 0.f ? 1u : 0u) << i;
                bits[k] = b0;
            }
        };
        auto epilogue = [&](const Tile& c, int acc, const uint32_t (&bits)[NW]) {
            bool pix_ok;
            const int64_t o = out_offset(c, pix_ok);
            float* orow = p.out + o;
#pragma unroll
            for (int k = 0; k < CO / 16; ++k) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * CO + 16 * k, v);
                if (pix_ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int ch = 16 * k + i;
                        float ov = v[i] + bias_s[ch];
                        if (p.relu) ov = fmaxf(ov, 0.f);
                        if (!((bits[ch >> 5] >> (ch & 31)) & 1u)) ov = 0.f;
                        v[i] = ov;
                    }
                    st8f(orow + 16 * k, v); st8f(orow + 16 * k + 8, v + 8);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_free + 8 * acc);
        };
        for (uint32_t it = 0; it < ntile; ++it) {
            const Tile cur = decode(first + it);
            uint32_t bits[NW];
            relu_bits(cur, bits);
            mbar_wait(bar_done + 8 * (it & 1), (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue(cur, it & 1, bits);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace conv96u

using namespace conv96u;

bool conv96_umma_eligible(const idee_conv_desc* d) {
    return d->umma96 && d->precision >= 1 && !d->proj && d->Cin == 96 && d->Cout == 96 && d->V == 1 && d->Vw == 1 && d->in_cpg == 6 &&
           d->out_cpg == 6 && d->x_sw == 96 && d->y_sw == 96 && !d->x_dtype && !d->y_dtype && !d->gx_dtype;
}

// the joint head's first conv on the 16-channel plane image (16 -> 96): forward and data gradient with resident weights
bool conv16to96_umma_eligible(const idee_conv_desc* d) {
    return d->umma96 && d->precision >= 1 && !d->proj && d->Cin == 16 && d->Cout == 96 && d->V == 1 && d->Vw == 1 && d->in_cpg == 1 &&
           d->out_cpg == 6 && d->x_sw == 16 && d->y_sw == 96 && !d->x_dtype && !d->y_dtype && !d->gx_dtype;
}

size_t conv96_umma_workspace_bytes() { return sizeof(__nv_bfloat16) * 18 * 96 * 96; }

template <int MODE, int GI, int GO>
static int launch96(const UP& p, int64_t total, cudaStream_t st) {
    constexpr int ktin = MODE == U_FWD ? 2 : 1;
    constexpr size_t halo_pad = ((size_t)(GI / 8) * (ktin * HR * HC * 16 + 16) + 127) / 128 * 128;
    constexpr size_t b_tap = (size_t)GO * GI * 2;
    constexpr size_t slots = 18 * b_tap <= 64 * 1024 ? 18 : (MODE == U_DGRAD ? 8 : 4);
    constexpr size_t smem = 2 * halo_pad + slots * b_tap + GO * 4 + (8 + 2 * NB_MAX) * 8 + 16;
    constexpr int tmem_cols = 2 * GO > 128 ? 256 : (2 * GO > 64 ? 128 : (2 * GO > 32 ? 64 : 32));
    auto kern = conv96_umma_kernel<MODE, GI, GO>;
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d(umma96)");
    (void)tmem_cols;
    int64_t grid = (int64_t)idee_num_sms();               // one persistent CTA per SM (416 threads x 118 registers)
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, NTHREADS, smem, st>>>(p);
    IDEE_LAUNCH_CHECK("conv3d(umma96)");
    return 0;
}

// dgrad == 0: y = conv(x) (+bias, ReLU);  dgrad == 1: gx = conv^T(gy) (optional ReLU mask)
int conv96_umma_run(const idee_conv_desc* d, int dgrad, const float* in, const float* w, const float* bias, const float* relu_src,
                    float* out, void* ws, cudaStream_t st) {
    __nv_bfloat16* wB = (__nv_bfloat16*)ws;
    const int GI = dgrad ? d->Cout : d->Cin, GO = dgrad ? d->Cin : d->Cout;      // gather-input / output channels of this pass
    prep96_weights_kernel<<<(18 * GI * GO + 255) / 256, 256, 0, st>>>(w, wB, dgrad, GI, GO);
    IDEE_LAUNCH_CHECK("conv3d(umma96) prep");
    UP p{};
    p.in = in; p.out = out; p.bias = bias; p.relu_src = relu_src; p.wB = wB; p.N = d->N;
    int64_t is[4], os[4];
    if (!dgrad) {
        p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
        is[0] = d->x_sn; is[1] = d->x_st; is[2] = d->x_sh; is[3] = d->x_sw; os[0] = d->y_sn; os[1] = d->y_st; os[2] = d->y_sh; os[3] = d->y_sw;
        p.relu = d->relu;
    } else {
        p.Ti = d->To; p.Hi = d->Ho; p.Wi = d->Wo; p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
        is[0] = d->y_sn; is[1] = d->y_st; is[2] = d->y_sh; is[3] = d->y_sw; os[0] = d->x_sn; os[1] = d->x_st; os[2] = d->x_sh; os[3] = d->x_sw;
        p.relu = 0;
    }
    IDEE_REQUIRE((int64_t)(p.Ti + 1) * is[1] + (int64_t)(p.Hi + HR) * is[2] + (int64_t)(p.Wi + HC) * is[3] < (1ll << 31) &&
                 (int64_t)(p.To + 1) * os[1] + (int64_t)(p.Ho + TR) * os[2] + (int64_t)(p.Wo + TC) * os[3] < (1ll << 31),
                 "conv3d(umma96): tensor too large for 32-bit image-relative offsets");
    p.in_sn = is[0]; p.in_st = (int)is[1]; p.in_sh = (int)is[2]; p.in_sw = (int)is[3];
    p.out_sn = os[0]; p.out_st = (int)os[1]; p.out_sh = (int)os[2]; p.out_sw = (int)os[3];
    p.tiles_w = (p.Wo + TC - 1) / TC; p.tiles_h = (p.Ho + TR - 1) / TR;
    const int64_t total = (int64_t)d->N * p.To * p.tiles_h * p.tiles_w;
    IDEE_REQUIRE(total < (1ll << 31), "conv3d(umma96): too many tiles");
    p.total_tiles = (uint32_t)total;
    if (d->Cin == 96) return dgrad ? launch96<U_DGRAD, 96, 96>(p, total, st) : launch96<U_FWD, 96, 96>(p, total, st);
    return dgrad ? launch96<U_DGRAD, 96, 16>(p, total, st) : launch96<U_FWD, 16, 96>(p, total, st);
}
