// tcgen05 / TMEM weight gradients of the joint classifier head: the 96 -> 96 convolution Conv3d(96, 96, (2,3,3), stride (2,1,1),
// pad (0,1,1)) (classifier/CNN_3D.py:84; first kernel) and the head's first conv on the 16-channel plane image (second kernel):   dW[co][ci][tap] = sum over output pixels p of gy[p][co] * x[p + tap][ci],   db[co] = sum_p gy[p][co].
//
// The contraction runs over PIXELS, so both operands are read as MN-major views of channel-chunk planes (the construction of the
// Swin weight-gradient GEMMs, swin_umma.cuh): the gy tile [128 pixels x 96] and one input time slice of the halo [18 x 10 pixels
// x 96] are stored as twelve 8-channel planes with 16 bytes per pixel; a K step of 16 pixels = two tile rows of 8 pixels = two
// 8 x 16-byte core matrices (LBO = next tile row, SBO = next channel chunk), and tap (kh, kw) only moves the start address of
// the halo operand.  One MMA (M = 128: 96 co + 32 idle lanes, N = 96 ci, K = 16 pixels) per tap and K step accumulates
// D_tap[co][ci] in TMEM over the CTA's whole persistent loop: the accumulators never visit registers or shared memory.
// 18 taps x 96 columns do not fit the 512 TMEM columns, so a CTA owns ONE of four tap groups (time slice kt, taps 0-4 or 5-8 of
// its 3x3) and a contiguous range of tiles; the four CTAs of a range run side by side and share their reads in L2.  The bias
// gradient is one more N = 16 MMA per K step against a constant ones operand (LBO = SBO = 0).
// Roles (544 threads): warps 0-15 convert the fp32 halo slice and gy tile of tile i+1 into the bf16 planes while warp 16 issues
// the MMAs of tile i; hand-offs are mbarriers (planes full / MMAs done), no CTA-wide barrier in the steady state.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace conv96w {

constexpr int TR = 16, TC = 8, HR = TR + 2, HC = TC + 2;
constexpr int CO = 96;
constexpr int NPX = HR * HC;                          // 180 halo pixels of one time slice
constexpr int GCHUNK = TR * TC * 16 + 16;             // gy plane stride (+16: a pixel's chunks land in different banks)
constexpr int GBUF = (12 * GCHUNK + 127) / 128 * 128;
constexpr int NLOAD = 512, MMA_WARP = NLOAD / 32, NTHREADS = NLOAD + 32;
constexpr int TMEM_COLS = 512;

// CI = 96 (the 96 -> 96 conv, the instantiated case): four tap groups, one input time slice per CTA.  The CI = 16 parameters (all 18
// taps as N = 16 MMAs in one CTA) describe the first tcgen05 form of the joint head's 16 -> 96 weight gradient; it re-read gy once
// per tap and is superseded by conv16_wgrad_stack_kernel below.
template <int CI_> struct Cfg {
    static constexpr int CI = CI_, KC = CI / 8;
    static constexpr int NGROUP = CI == 96 ? 4 : 1, MAXTAPS = CI == 96 ? 5 : 18, NPL = CI == 96 ? 1 : 2;   // halo time slices per CTA
    static constexpr int XCHUNK = NPL * NPX * 16 + 16;                    // halo plane stride
    static constexpr int XBUF = (KC * XCHUNK + 127) / 128 * 128;
    static constexpr int BIAS_COL = MAXTAPS * CI;
    static constexpr int PART = MAXTAPS * CO * CI + CO;                   // floats per CTA: [tap slot][co][ci] | db[co]
};

struct WPU {
    const float* x; const float* gy; float* partials;
    int N, Ti, Hi, Wi, To;
    int64_t x_sn, gy_sn;
    int x_st, x_sh, x_sw, gy_st, gy_sh, gy_sw;
    int tiles_w, tiles_h, S;
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
// D = F32, A = B = BF16, both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
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
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// grid = NGROUP * S: CTA b owns tap group b % NGROUP (CI = 96: time slice kt = group >> 1, taps [0,5) or [5,9) of its 3x3) and tile
// slice b / NGROUP
template <int CI_>
__global__ void __launch_bounds__(NTHREADS, 1)
conv96_wgrad_umma_kernel(WPU p) {
    using C_ = Cfg<CI_>;
    constexpr int CI = C_::CI, XCHUNK = C_::XCHUNK, XBUF = C_::XBUF, MAXTAPS = C_::MAXTAPS, BIAS_COL = C_::BIAS_COL,
                  PART = C_::PART, NGROUP = C_::NGROUP, NPL = C_::NPL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* gbuf = smem_raw;                                        // [2][KC][128 px][16 B]  gy tile planes
    unsigned char* xbuf = smem_raw + 2 * GBUF;                             // [2][KC][180 px][16 B]  halo slice planes
    unsigned char* ones = xbuf + 2 * XBUF;                                 // one 8 x 16-byte core matrix of bf16 ones
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 128);              // full[2] | done[2] | final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[2]), bar_final = smem_u32(&bars[4]);
    const int group = blockIdx.x % NGROUP, slice = blockIdx.x / NGROUP;
    // taps of this CTA: CI = 96: j9 = tap0 + j of time slice kt0;  CI = 16: tap j = kt * 9 + j9 over both slices
    const int kt0 = NGROUP == 4 ? group >> 1 : 0, tap0 = (NGROUP == 4 && (group & 1)) ? 5 : 0;
    const int ntap = NGROUP == 4 ? ((group & 1) ? 4 : 5) : 18;

    if (tid < 32) reinterpret_cast<uint32_t*>(ones)[tid] = 0x3F803F80u;
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, NLOAD); mbar_init(bar_done + 8 * i, 1); }
        mbar_init(bar_final, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t per = (p.total_tiles + p.S - 1) / p.S;
    const uint32_t first = min(p.total_tiles, (uint32_t)slice * per), last = min(p.total_tiles, first + per);
    const uint32_t ntile = last - first;
    struct Tile { int n, t, h0, w0; };
    auto decode = [&](uint32_t tile) {
        Tile c;
        uint32_t r = tile;
        c.w0 = (int)(r % (uint32_t)p.tiles_w) * TC; r /= (uint32_t)p.tiles_w;
        c.h0 = (int)(r % (uint32_t)p.tiles_h) * TR; r /= (uint32_t)p.tiles_h;
        c.t = (int)(r % (uint32_t)p.To); c.n = (int)(r / (uint32_t)p.To);
        return c;
    };

    if (warp == MMA_WARP) {
        // ================= MMA issue =================
        constexpr uint32_t ID_W = idesc_mn(128, CI), ID_B = idesc_mn(128, 16);
        const uint64_t ones_desc = make_desc(smem_u32(ones), 0, 0);
        for (uint32_t it = 0; it < ntile; ++it) {
            const uint32_t b = it & 1;
            mbar_wait(bar_full + 8 * b, (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint64_t adesc0 = make_desc(smem_u32(gbuf) + b * GBUF, TC * 16, GCHUNK);      // K: next tile row, M: next chunk
                const uint64_t bdesc0 = make_desc(smem_u32(xbuf) + b * XBUF, HC * 16, XCHUNK);
#pragma unroll 1
                for (int ks = 0; ks < TR / 2; ++ks) {                  // 16 pixels = tile rows 2 ks, 2 ks + 1
                    const uint64_t ad = adesc0 + (uint64_t)(ks * 2 * TC);
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    for (int j = 0; j < ntap; ++j) {
                        const int jj = tap0 + j, pl = jj / 9, j9 = jj - pl * 9, kh = j9 / 3, kw = j9 - kh * 3;   // pl: slice inside the halo buffer
                        umma(tmem_base + j * CI, ad, bdesc0 + (uint64_t)((pl * HR + 2 * ks + kh) * HC + kw), ID_W, acc);
                    }
                    umma(tmem_base + BIAS_COL, ad, ones_desc, ID_B, acc);
                }
                umma_commit(bar_done + 8 * b);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(bar_final);
        __syncwarp();
    } else {
        // ================= loaders: fp32 HBM -> bf16 chunk planes, one tile ahead of the MMAs =================
        constexpr int V4 = CI / 4;                                     // float4 units per pixel
        constexpr int G4 = CO / 4;                                     // float4 units per gy pixel
        constexpr int XU = NPL * NPX * V4, GU = TR * TC * G4, TOTAL = XU + GU;
        constexpr int DEPTH = (TOTAL + NLOAD - 1) / NLOAD;             // 15: the whole tile in ONE round of loads per thread (one memory latency per tile)
        // unit e -> (source pointer or null, destination byte offset inside the tile's x / gy planes; < 0: no such unit)
        auto unit = [&](int e, const Tile& c, bool t_ok, const float* x_n, const float* g_n, const float*& src) -> int {
            src = nullptr;
            if (e < XU) {
                const int q = e / V4, c4 = e - q * V4, pl = q / NPX, q1 = q - pl * NPX, hh = q1 / HC, ww = q1 - hh * HC;
                const int hi = c.h0 + hh - 1, wi = c.w0 + ww - 1;
                const bool ok = (NPL == 1 ? t_ok : 2 * c.t + pl < p.Ti) && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi;
                if (ok) src = x_n + (pl * p.x_st + hi * p.x_sh + wi * p.x_sw) + c4 * 4;
                return (c4 >> 1) * XCHUNK + q * 16 + (c4 & 1) * 8;
            }
            if (e < TOTAL) {
                const int eg = e - XU, q = eg / G4, c4 = eg - q * G4, r = q / TC, cc = q - r * TC;
                const int h = c.h0 + r, w = c.w0 + cc;
                if (h < p.Hi && w < p.Wi) src = g_n + (h * p.gy_sh + w * p.gy_sw) + c4 * 4;
                return 2 * XBUF + (c4 >> 1) * GCHUNK + q * 16 + (c4 & 1) * 8;      // marks a gy unit: offset beyond the x buffers
            }
            return -1;
        };
        for (uint32_t it = 0; it < ntile; ++it) {
            const uint32_t b = it & 1;
            const Tile c = decode(first + it);
            unsigned char* xdst = xbuf + b * XBUF;
            unsigned char* gdst = gbuf + b * GBUF;
            const float* x_n = p.x + c.n * p.x_sn + (int64_t)(2 * c.t + kt0) * p.x_st;
            const float* g_n = p.gy + c.n * p.gy_sn + (int64_t)c.t * p.gy_st;
            const bool t_ok = 2 * c.t + kt0 < p.Ti;
            float4 f[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {                          // loads of tile it fly while the MMAs of tile it-2 drain
                const float* src;
                unit(tid + u * NLOAD, c, t_ok, x_n, g_n, src);
                f[u] = src ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (it >= 2) mbar_wait(bar_done + 8 * b, ((it >> 1) - 1) & 1u);      // the MMAs of tile it-2 have read buffer b
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const float* src;
                const int o = unit(tid + u * NLOAD, c, t_ok, x_n, g_n, src);
                if (o >= 0) {
                    unsigned char* dst = o >= 2 * XBUF ? gdst + (o - 2 * XBUF) : xdst + o;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            mbar_arrive(bar_full + 8 * b);
        }
    }
    // ================= accumulators -> per-CTA partials =================
    float* part = p.partials + (size_t)blockIdx.x * PART;
    if (warp < 3) {                                                    // TMEM lanes 0..95 = co
        if (ntile > 0) {
            mbar_wait(bar_final, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int co = warp * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int col0 = 0; col0 < MAXTAPS * CI; col0 += 32) {          // 32 accumulator columns = (tap slot, ci) pairs, ci fastest
            float v[32];
            if (ntile > 0) tmem_ld32(tl + col0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const int col = col0 + i, j = col / CI, ci = col - j * CI;
                const bool live = ntile > 0 && j < ntap;                   // unused tap slots and CTAs without tiles contribute zeros
                *reinterpret_cast<float4*>(part + ((size_t)j * CO + co) * CI + ci) =
                    live ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        {
            float v[32];
            if (ntile > 0) tmem_ld32(tl + BIAS_COL - 16, v);           // columns BIAS_COL-16 .. BIAS_COL+15: element 16 = db[co]
            part[(size_t)MAXTAPS * CO * CI + co] = ntile > 0 ? v[16] : 0.f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// gw[co][ci][kt][kh][kw] = sum over the S tile slices of the owning group's partial; gb[co] from the group-0 CTAs
template <int CI_>
__global__ void conv96_wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ gw, float* __restrict__ gb, int S) {
    using C_ = Cfg<CI_>;
    constexpr int CI = C_::CI, PART = C_::PART, NGROUP = C_::NGROUP, MAXTAPS = C_::MAXTAPS;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;                // (tap, co, ci), ci fastest: coalesced partial reads
    if (e < 18 * CO * CI) {
        const int ci = e % CI, co = (e / CI) % CO, tap = e / (CI * CO);
        int group = 0, j = tap;
        if (NGROUP == 4) { const int kt = tap / 9, j9 = tap % 9; group = kt * 2 + (j9 >= 5 ? 1 : 0); j = j9 >= 5 ? j9 - 5 : j9; }
        const float* src = partials + (size_t)group * PART + ((size_t)j * CO + co) * CI + ci;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += src[(size_t)s * NGROUP * PART];
        gw[((size_t)co * CI + ci) * 18 + tap] = acc;
    } else if (e < 18 * CO * CI + CO && gb) {
        const int co = e - 18 * CO * CI;
        const float* src = partials + (size_t)MAXTAPS * CO * CI + co;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += src[(size_t)s * NGROUP * PART];
        gb[co] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// 16 -> 96 (the joint head's first conv on the 16-channel plane image): with N = 16 input channels per tap the kernel above re-reads
// the 96-channel gy operand for every one of its 18 taps (85 KB of shared-memory operand reads per 16 pixels).  Here the roles are
// swapped and the three kh taps are STACKED along M:  D_kw[(kh, ci)][co] += x[row + kh][col + kw][ci] * gy[row][col][co]
//   A = halo slice, MN-major, M = 64 = 4 row slots x 16 ci: the slice is stored [row][chunk][col][16 B] (chunk stride 160 B, row
//       stride 320 B), so the eight M core matrices (row slot s, chunk) sit at the uniform stride 160 B; slot 3 is idle;
//   B = gy tile, MN-major, N = 96 co;   three MMAs (kw) + one bias MMA (A = ones) per 16 pixels: 15 KB of operand reads.
// A CTA owns one input time slice kt (grid = 2 x S) and the 3 x 96 + 96 accumulator columns of its nine taps + bias.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int X2ROWS = HR + 1;                                    // + one slack row read by the idle row slot
constexpr int X2BUF = (X2ROWS * 2 * HC * 16 + 127) / 128 * 128;
constexpr int PART2 = 9 * 16 * CO + CO;                           // floats per CTA: [kh][kw][ci][co] | db[co]

__global__ void __launch_bounds__(NTHREADS, 1)
conv16_wgrad_stack_kernel(WPU p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* gbuf = smem_raw;                                        // [2][12 chunks][128 px][16 B]
    unsigned char* xbuf = smem_raw + 2 * GBUF;                             // [2][19 rows][2 chunks][10 cols][16 B]
    unsigned char* ones = xbuf + 2 * X2BUF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 128);              // full[2] | done[2] | final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[2]), bar_final = smem_u32(&bars[4]);
    const int kt = blockIdx.x & 1, slice = blockIdx.x >> 1;

    if (tid < 32) reinterpret_cast<uint32_t*>(ones)[tid] = 0x3F803F80u;
    for (int i = tid; i < 2 * X2BUF / 16; i += NTHREADS) reinterpret_cast<uint4*>(xbuf)[i] = make_uint4(0u, 0u, 0u, 0u);   // slack rows: finite
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, NLOAD); mbar_init(bar_done + 8 * i, 1); }
        mbar_init(bar_final, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t per = (p.total_tiles + p.S - 1) / p.S;
    const uint32_t first = min(p.total_tiles, (uint32_t)slice * per), last = min(p.total_tiles, first + per);
    const uint32_t ntile = last - first;
    struct Tile { int n, t, h0, w0; };
    auto decode = [&](uint32_t tile) {
        Tile c;
        uint32_t r = tile;
        c.w0 = (int)(r % (uint32_t)p.tiles_w) * TC; r /= (uint32_t)p.tiles_w;
        c.h0 = (int)(r % (uint32_t)p.tiles_h) * TR; r /= (uint32_t)p.tiles_h;
        c.t = (int)(r % (uint32_t)p.To); c.n = (int)(r / (uint32_t)p.To);
        return c;
    };

    if (warp == MMA_WARP) {
        constexpr uint32_t ID = idesc_mn(64, CO);
        const uint64_t ones_desc = make_desc(smem_u32(ones), 0, 0);
        for (uint32_t it = 0; it < ntile; ++it) {
            const uint32_t b = it & 1;
            mbar_wait(bar_full + 8 * b, (it >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint64_t bdesc0 = make_desc(smem_u32(gbuf) + b * GBUF, TC * 16, GCHUNK);          // gy: K next tile row, N next chunk
                const uint64_t adesc0 = make_desc(smem_u32(xbuf) + b * X2BUF, 2 * HC * 16, HC * 16);   // x: K next halo row, M next (slot, chunk)
#pragma unroll 1
                for (int ks = 0; ks < TR / 2; ++ks) {
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    const uint64_t bd = bdesc0 + (uint64_t)(ks * 2 * TC);
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
                        umma(tmem_base + kw * CO, adesc0 + (uint64_t)(2 * ks * 2 * HC + kw), bd, ID, acc);
                    umma(tmem_base + 3 * CO, ones_desc, bd, ID, acc);
                }
                umma_commit(bar_done + 8 * b);
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(bar_final);
        __syncwarp();
    } else {
        constexpr int XU = NPX * 4, G4 = CO / 4, GU = TR * TC * G4, TOTAL = XU + GU;
        constexpr int DEPTH = (TOTAL + NLOAD - 1) / NLOAD;
        auto unit = [&](int e, const Tile& c, bool t_ok, const float* x_n, const float* g_n, const float*& src) -> int {
            src = nullptr;
            if (e < XU) {
                const int q = e >> 2, c4 = e & 3, hh = q / HC, ww = q - hh * HC;
                const int hi = c.h0 + hh - 1, wi = c.w0 + ww - 1;
                if (t_ok && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi) src = x_n + (hi * p.x_sh + wi * p.x_sw) + c4 * 4;
                return (hh * 2 + (c4 >> 1)) * (HC * 16) + ww * 16 + (c4 & 1) * 8;
            }
            if (e < TOTAL) {
                const int eg = e - XU, q = eg / G4, c4 = eg - q * G4, r = q / TC, cc = q - r * TC;
                const int h = c.h0 + r, w = c.w0 + cc;
                if (h < p.Hi && w < p.Wi) src = g_n + (h * p.gy_sh + w * p.gy_sw) + c4 * 4;
                return 2 * X2BUF + (c4 >> 1) * GCHUNK + q * 16 + (c4 & 1) * 8;             // marks a gy unit
            }
            return -1;
        };
        for (uint32_t it = 0; it < ntile; ++it) {
            const uint32_t b = it & 1;
            const Tile c = decode(first + it);
            unsigned char* xdst = xbuf + b * X2BUF;
            unsigned char* gdst = gbuf + b * GBUF;
            const float* x_n = p.x + c.n * p.x_sn + (int64_t)(2 * c.t + kt) * p.x_st;
            const float* g_n = p.gy + c.n * p.gy_sn + (int64_t)c.t * p.gy_st;
            const bool t_ok = 2 * c.t + kt < p.Ti;
            float4 f[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const float* src;
                unit(tid + u * NLOAD, c, t_ok, x_n, g_n, src);
                f[u] = src ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (it >= 2) mbar_wait(bar_done + 8 * b, ((it >> 1) - 1) & 1u);
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const float* src;
                const int o = unit(tid + u * NLOAD, c, t_ok, x_n, g_n, src);
                if (o >= 0) {
                    unsigned char* dst = o >= 2 * X2BUF ? gdst + (o - 2 * X2BUF) : xdst + o;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_full + 8 * b);
        }
    }
    // accumulators -> partials: M = 64 rows (kh slot s, ci) live in lane ci of warp s
    float* part = p.partials + (size_t)blockIdx.x * PART2;
    if (warp < 3) {
        if (ntile > 0) {
            mbar_wait(bar_final, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int col0 = 0; col0 < 4 * CO; col0 += 32) {
            float v[32];
            if (ntile > 0) tmem_ld32(tl + col0, v);
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            const int kw = col0 / CO, co0 = col0 - kw * CO;
            if (kw < 3) {
                if (lane < 16) {
                    float* dst = part + ((size_t)((warp * 3 + kw) * 16 + lane)) * CO + co0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            } else if (warp == 0 && lane == 0) {
                float* dst = part + (size_t)9 * 16 * CO + co0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// gw[co][ci][kt][kh][kw] = sum over the S tile slices of part[slice * 2 + kt][kh][kw][ci][co]; gb from the kt = 0 CTAs
__global__ void conv16_wgrad_stack_reduce_kernel(const float* __restrict__ partials, float* __restrict__ gw, float* __restrict__ gb, int S) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;                // (kt, j9, ci, co), co fastest: coalesced partial reads
    if (e < 18 * 16 * CO) {
        const int co = e % CO, ci = (e / CO) % 16, j9 = (e / (CO * 16)) % 9, kt = e / (CO * 16 * 9);
        const float* src = partials + (size_t)kt * PART2 + ((size_t)j9 * 16 + ci) * CO + co;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += src[(size_t)s * 2 * PART2];
        gw[((size_t)co * 16 + ci) * 18 + kt * 9 + j9] = acc;
    } else if (e < 18 * 16 * CO + CO && gb) {
        const int co = e - 18 * 16 * CO;
        const float* src = partials + (size_t)9 * 16 * CO + co;
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += src[(size_t)s * 2 * PART2];
        gb[co] = acc;
    }
}

}  // namespace conv96w

using namespace conv96w;

template <int CI_> static int wgrad_slices() {
    int S = idee_num_sms() / Cfg<CI_>::NGROUP;
    return S < 1 ? 1 : S;
}

bool conv16to96_wgrad_umma_eligible(const idee_conv_desc* d) {
    return d->umma96 && d->precision >= 1 && !d->proj && d->Cin == 16 && d->Cout == 96 && d->V == 1 && d->Vw == 1 && d->in_cpg == 1 &&
           d->out_cpg == 6 && d->x_sw == 16 && d->y_sw == 96 && !d->x_dtype && !d->y_dtype;
}

static int stack_slices() { const int S = idee_num_sms() / 2; return S < 1 ? 1 : S; }

size_t conv96_wgrad_umma_workspace_bytes(int Cin) {
    return Cin == 96 ? sizeof(float) * (size_t)4 * wgrad_slices<96>() * Cfg<96>::PART : sizeof(float) * (size_t)2 * stack_slices() * PART2;
}

template <int CI_>
static int wgrad_run(const idee_conv_desc* d, const float* x, const float* gy, float* gw, float* gb, void* ws, cudaStream_t st) {
    using C_ = Cfg<CI_>;
    WPU p{};
    p.x = x; p.gy = gy; p.partials = (float*)ws;
    p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To;
    IDEE_REQUIRE((int64_t)(d->Ti + 1) * d->x_st + (int64_t)(d->Hi + HR) * d->x_sh + (int64_t)(d->Wi + HC) * d->x_sw < (1ll << 31) &&
                 (int64_t)(d->To + 1) * d->y_st + (int64_t)(d->Ho + TR) * d->y_sh + (int64_t)(d->Wo + TC) * d->y_sw < (1ll << 31),
                 "conv3d_wgrad(umma96): tensor too large for 32-bit image-relative offsets");
    p.x_sn = d->x_sn; p.x_st = (int)d->x_st; p.x_sh = (int)d->x_sh; p.x_sw = (int)d->x_sw;
    p.gy_sn = d->y_sn; p.gy_st = (int)d->y_st; p.gy_sh = (int)d->y_sh; p.gy_sw = (int)d->y_sw;
    p.tiles_w = (d->Wo + TC - 1) / TC; p.tiles_h = (d->Ho + TR - 1) / TR;
    const int64_t total = (int64_t)d->N * d->To * p.tiles_h * p.tiles_w;
    IDEE_REQUIRE(total < (1ll << 31), "conv3d_wgrad(umma96): too many tiles");
    p.total_tiles = (uint32_t)total;
    p.S = wgrad_slices<CI_>();
    const size_t smem = 2 * (size_t)GBUF + 2 * (size_t)C_::XBUF + 128 + 5 * 8 + 16;
    IDEE_CUDA(cudaFuncSetAttribute(conv96_wgrad_umma_kernel<CI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_wgrad(umma96)");
    conv96_wgrad_umma_kernel<CI_><<<C_::NGROUP * p.S, NTHREADS, smem, st>>>(p);
    IDEE_LAUNCH_CHECK("conv3d_wgrad(umma96)");
    const int nel = 18 * CO * CI_ + CO;
    conv96_wgrad_reduce_kernel<CI_><<<(nel + 255) / 256, 256, 0, st>>>(p.partials, gw, gb, p.S);
    IDEE_LAUNCH_CHECK("conv3d_wgrad(umma96) reduce");
    return 0;
}

static int stack_run(const idee_conv_desc* d, const float* x, const float* gy, float* gw, float* gb, void* ws, cudaStream_t st) {
    WPU p{};
    p.x = x; p.gy = gy; p.partials = (float*)ws;
    p.N = d->N; p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To;
    IDEE_REQUIRE((int64_t)(d->Ti + 1) * d->x_st + (int64_t)(d->Hi + HR) * d->x_sh + (int64_t)(d->Wi + HC) * d->x_sw < (1ll << 31) &&
                 (int64_t)(d->To + 1) * d->y_st + (int64_t)(d->Ho + TR) * d->y_sh + (int64_t)(d->Wo + TC) * d->y_sw < (1ll << 31),
                 "conv3d_wgrad(umma 16->96): tensor too large for 32-bit image-relative offsets");
    p.x_sn = d->x_sn; p.x_st = (int)d->x_st; p.x_sh = (int)d->x_sh; p.x_sw = (int)d->x_sw;
    p.gy_sn = d->y_sn; p.gy_st = (int)d->y_st; p.gy_sh = (int)d->y_sh; p.gy_sw = (int)d->y_sw;
    p.tiles_w = (d->Wo + TC - 1) / TC; p.tiles_h = (d->Ho + TR - 1) / TR;
    const int64_t total = (int64_t)d->N * d->To * p.tiles_h * p.tiles_w;
    IDEE_REQUIRE(total < (1ll << 31), "conv3d_wgrad(umma 16->96): too many tiles");
    p.total_tiles = (uint32_t)total;
    p.S = stack_slices();
    const size_t smem = 2 * (size_t)GBUF + 2 * (size_t)X2BUF + 128 + 5 * 8 + 16;
    IDEE_CUDA(cudaFuncSetAttribute(conv16_wgrad_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_wgrad(umma 16->96)");
    conv16_wgrad_stack_kernel<<<2 * p.S, NTHREADS, smem, st>>>(p);
    IDEE_LAUNCH_CHECK("conv3d_wgrad(umma 16->96)");
    const int nel = 18 * 16 * CO + CO;
    conv16_wgrad_stack_reduce_kernel<<<(nel + 255) / 256, 256, 0, st>>>(p.partials, gw, gb, p.S);
    IDEE_LAUNCH_CHECK("conv3d_wgrad(umma 16->96) reduce");
    return 0;
}

int conv96_wgrad_umma_run(const idee_conv_desc* d, const float* x, const float* gy, float* gw, float* gb, void* ws, cudaStream_t st) {
    return d->Cin == 96 ? wgrad_run<96>(d, x, gy, gw, gb, ws, st) : stack_run(d, x, gy, gw, gb, ws, st);
}
