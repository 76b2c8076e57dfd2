// tcgen05 / TMEM implicit-GEMM kernel for the one dense contraction of the path: the 96 -> 96 classifier convolution
// Conv3d(96, 96, (2,3,3), stride (2,1,1), pad (0,1,1)) (classifier/CNN_3D.py:84), forward and data gradient (sm_100a).
//
// A CTA owns a tile of 128 output pixels (8 x 16) = the 128 TMEM lanes; the accumulator D[128 x 96] fp32 lives in 96 TMEM
// columns for the whole K loop (K = taps x 96 = 1728 forward, 864 per t-parity in the data gradient).  Per tap:
//   * A_tap [128 pixels x 96 channels] bf16 is copied from the shared-memory halo into the canonical K-major
//     SWIZZLE_NONE layout (8x8 core matrices, SBO = 128 B between 8-row groups, LBO = 2048 B between 8-channel chunks),
//   * B_tap [96 x 96] bf16 (pre-arranged in the same canonical layout by a prep kernel) is copied from L2,
//   * one elected thread issues 6 tcgen05.mma.cta_group::1.kind::f16 (M=128, N=96, K=16) from shared-memory descriptors and a
//     tcgen05.commit that arrives on the mbarrier guarding the double-buffered A/B slots.
// The epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 columns per warp and instruction), adds bias / ReLU and
// stores fp32.  Descriptor and instruction-descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp.
#include <cuda_bf16.h>

#include "common.cuh"
#include "idee_b200.h"

namespace convumma {

constexpr int TH = 8, TW = 16, HH = TH + 2, HWp = TW + 2;
constexpr int CI = 96, CO = 96, CPH = CI + 8;          // halo pixel stride in halves (conflict-free 16-byte reads)
constexpr int A_BYTES = 128 * CI * 2, B_BYTES = CO * CI * 2;
constexpr int A_LBO = 16 * 128, B_LBO = (CO / 8) * 128, SBO = 128;
constexpr int TMEM_COLS = 128;

enum { U_FWD = 0, U_DGRAD = 1 };

struct UP {
    const float* in; float* out; const float* bias; const float* relu_src; const __nv_bfloat16* wB;
    int N, Ti, Hi, Wi, To, Ho, Wo;
    int64_t in_sn, in_st, in_sh, in_sw, out_sn, out_st, out_sh, out_sw;
    int relu, tiles_w, tiles_h;
    int64_t total_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 96, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    while (!ok) {
        __nanosleep(64);                      // idle warps must not steal issue slots from the MMA / copy warp
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

// fp32 reference-layout weights [Co][Ci][18] -> bf16 canonical K-major tiles, one per forward tap:
//   wB[ft][kc][ng][r][e] = B(n = ng*8 + r, k = kc*8 + e);  forward: B(n,k) = W[n][k][ft];  dgrad: B(n,k) = W[k][n][ft]
__global__ void prep_umma_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wB, int dgrad) {
    const int total = 18 * CO * CI;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int el = e & 7, r = (e >> 3) & 7;
        int q = e >> 6;
        const int ng = q % (CO / 8); q /= (CO / 8);
        const int kc = q % (CI / 8);
        const int ft = q / (CI / 8);
        const int n = ng * 8 + r, k = kc * 8 + el;
        const int fo = dgrad ? k : n, fc = dgrad ? n : k;
        wB[e] = __float2bfloat16(w[((int64_t)fo * CI + fc) * 18 + ft]);
    }
}

// Zero-copy operand layout.  The halo of a tile is stored as 3 column-shifted copies ("planes"), one per kw, each with a row
// pitch of exactly 16 pixels and split into 8-channel (16-byte) chunks:
//     plane[kw][kt][kc][hh][px]  (16 B each)  =  channels kc*8..kc*8+7 of halo pixel (kt, hh, px + kw)
// Output pixel m = r*16 + c of the tile reads, for tap (kt,kh,kw), plane[kw][kt][kc][r+kh][c] = linear row kh*16 + m of that
// plane: exactly the canonical K-major SWIZZLE_NONE layout (8 rows x 16 B core matrices, SBO = 128 B, LBO = HH*16*16 B), so a
// tap is nothing but a descriptor start address.  No per-tap copies, no CTA-wide barriers inside the K loop.
template <int MODE>
__global__ void __launch_bounds__(128, 1)
conv_umma_kernel(UP p) {
    constexpr int NPH = MODE == U_FWD ? 2 : 1;                      // phases = forward kt taps (one input time slice each)
    constexpr int NB = 6;                                           // weight-tap ring depth
    constexpr int P_LBO = HH * 16 * 16 + 16;                        // bytes between 8-channel chunks of a plane (+16: the 12
                                                                    // chunks of a pixel fall into different banks when stored)
    constexpr int P_BYTES = (CI / 8) * P_LBO;                       // one kw plane
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* planes = smem_raw;                                                          // [3 kw] planes
    unsigned char* Bbuf = planes + 3 * P_BYTES;                                                // NB x B_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bbuf + NB * B_BYTES);                         // [NB] slot free, [NB] = phase done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NB + 1);
    __shared__ __align__(16) float bias_s[CO];           // parameters are only 4-byte aligned (views of a flat buffer)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < CO) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i <= NB; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    uint32_t slot_par = 0u;            // bit i = phase parity of slot barrier i (warp 0 only)
    uint32_t done_par = 0u;

    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int64_t rr = tile;
        const int w0 = (int)(rr % p.tiles_w) * TW; rr /= p.tiles_w;
        const int h0 = (int)(rr % p.tiles_h) * TH; rr /= p.tiles_h;
        const int t = (int)(rr % p.To);
        const int n = (int)(rr / p.To);
#pragma unroll 1
        for (int ph = 0; ph < NPH; ++ph) {
            // ---- halo slice: fp32 HBM -> bf16, written into the three shifted planes (zero padding) ----
            {
                const int ti = MODE == U_FWD ? 2 * t + ph : (t >> 1);
                const float* in_img = p.in + n * p.in_sn + ti * p.in_st;
                const bool tok = ti < p.Ti;
                constexpr int V4 = CI / 4, NCOL = HWp * V4, DROW = 128 / NCOL, DCOL = 128 - DROW * NCOL;
                int row = tid / NCOL, col = tid - row * NCOL;
                constexpr int DEPTH = 16;          // independent 16-byte loads in flight per thread (1 CTA/SM: registers are free)
                while (row < HH) {
                    const float* src[DEPTH];
                    int dsts[DEPTH];               // packed (row, ww, c4); -1 = none
#pragma unroll
                    for (int u = 0; u < DEPTH; ++u) {
                        src[u] = nullptr; dsts[u] = -1;
                        if (row < HH) {
                            const int ww = col / V4, c4 = col - ww * V4;
                            const int hi = h0 + row - 1, wi = w0 + ww - 1;
                            const bool ok = tok && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
                            dsts[u] = (row << 16) | (ww << 8) | c4;
                            if (ok) src[u] = in_img + hi * p.in_sh + wi * p.in_sw + c4 * 4;
                            row += DROW; col += DCOL;
                            if (col >= NCOL) { col -= NCOL; ++row; }
                        }
                    }
                    float4 f[DEPTH];
#pragma unroll
                    for (int u = 0; u < DEPTH; ++u) f[u] = src[u] ? ldg4(src[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < DEPTH; ++u) {
                        if (dsts[u] < 0) continue;
                        const uint2 v = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
                        const int rw = dsts[u] >> 16, ww = (dsts[u] >> 8) & 0xFF, c4 = dsts[u] & 0xFF;
                        const int kc = c4 >> 1, half = c4 & 1;                  // 8-channel chunk and which 8-byte half of it
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const int px = ww - kw;
                            if (px >= 0 && px < 16)
                                *reinterpret_cast<uint2*>(planes + kw * P_BYTES + kc * P_LBO + (rw * 16 + px) * 16 + half * 8) = v;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            __syncthreads();
            // ---- 9 taps: warp 0 streams the weight taps (cp.async ring) and issues the MMAs; no CTA-wide barriers ----
            if (warp == 0) {
                auto ftap_of = [&](int j) -> int {
                    if (MODE == U_FWD) return ph * 9 + j;
                    return (t & 1) * 9 + (2 - j / 3) * 3 + (2 - j % 3);
                };
                auto load_B = [&](int j) {
                    const uint4* src = reinterpret_cast<const uint4*>(p.wB + (size_t)ftap_of(j) * CO * CI);
                    unsigned char* dst = Bbuf + (j % NB) * B_BYTES;
#pragma unroll 4
                    for (int i = lane; i < B_BYTES / 16; i += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i * 16)), "l"(src + i) : "memory");
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
#pragma unroll
                for (int j = 0; j < NB; ++j) load_B(j);                    // every slot is free: the previous phase was drained
#pragma unroll 1
                for (int j = 0; j < 9; ++j) {
                    const int b = j % NB;
                    // B(j) has landed once at most the younger groups are still pending.  Issued so far: taps 0..NB-1 up front plus
                    // one refill at the end of each earlier iteration >= 1, i.e. up to tap min(8, max(NB-1, j+NB-2)).
                    const int newest = min(8, max(NB - 1, j + NB - 2));
                    const int younger = newest - j;
                    if (younger >= 5) asm volatile("cp.async.wait_group 5;" ::: "memory");
                    else if (younger == 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
                    else if (younger == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
                    else if (younger == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
                    else if (younger == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const int kh = j / 3, kw = j % 3;
                        const uint32_t a0 = smem_u32(planes + kw * P_BYTES + kh * 16 * 16);
                        const uint32_t b0 = smem_u32(Bbuf + b * B_BYTES);
#pragma unroll
                        for (int ks = 0; ks < CI / 16; ++ks)
                            umma_bf16(tmem_base, make_desc(a0 + ks * 2 * P_LBO, P_LBO, SBO), make_desc(b0 + ks * 2 * B_LBO, B_LBO, SBO),
                                      (ph > 0 || j > 0 || ks > 0) ? 1u : 0u);
                        umma_commit(smem_u32(&bars[b]));                    // slot b is free when these MMAs have read it
                        if (j == 8) umma_commit(smem_u32(&bars[NB]));       // ... and this phase's MMAs are complete
                    }
                    __syncwarp();
                    if (j >= 1 && j - 1 + NB < 9) {                         // refill the slot released by tap j-1 (overlaps MMA j)
                        const int bb = (j - 1) % NB;
                        mbar_wait(smem_u32(&bars[bb]), (slot_par >> bb) & 1u);
                        slot_par ^= 1u << bb;
                        load_B(j - 1 + NB);
                    }
                }
                // consume the slot-barrier phases not waited on inside the loop (taps whose slot was not refilled)
#pragma unroll 1
                for (int j = 0; j < 9; ++j) {
                    if (j + NB < 9) continue;
                    const int bb = j % NB;
                    mbar_wait(smem_u32(&bars[bb]), (slot_par >> bb) & 1u);
                    slot_par ^= 1u << bb;
                }
            }
            mbar_wait(smem_u32(&bars[NB]), done_par);                       // every thread: this phase's MMAs are complete
            done_par ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ph + 1 < NPH) __syncthreads();                              // planes may be refilled for the next time slice
        }
        // ---- epilogue: TMEM lane = pixel, 96 columns = output channels ----
        const int pr = tid >> 4, pc = tid & 15;
        const int h = h0 + pr, w = w0 + pc;
        const bool pix_ok = h < p.Ho && w < p.Wo;
        float* orow = p.out + n * p.out_sn + t * p.out_st + h * p.out_sh + w * p.out_sw;
        const float* rrow = p.relu_src ? p.relu_src + n * p.out_sn + t * p.out_st + h * p.out_sh + w * p.out_sw : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < CO; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            if (pix_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    { const float4 bb = ld4(bias_s + c0 + i); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (rrow) {
                        const float4 a = ldg4(rrow + c0 + i);
                        if (!(a.x > 0.f)) o.x = 0.f; if (!(a.y > 0.f)) o.y = 0.f; if (!(a.z > 0.f)) o.z = 0.f; if (!(a.w > 0.f)) o.w = 0.f;
                    }
                    st4(orow + c0 + i, o);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                  // every warp has drained the accumulator; planes may be overwritten
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace convumma

using namespace convumma;

bool conv_umma_eligible(const idee_conv_desc* d) {
    return d->precision == 2 && !d->proj && d->Cin == 96 && d->Cout == 96 && d->V == 1 && d->Vw == 1 && d->in_cpg == 6 &&
           d->out_cpg == 6 && d->x_sw == 96 && d->y_sw == 96;
}

size_t conv_umma_workspace_bytes() { return sizeof(__nv_bfloat16) * 18 * CO * CI; }

// dgrad == 0: y = conv(x) (+bias, ReLU);  dgrad == 1: gx = conv^T(gy) (optional ReLU mask)
int conv_umma_run(const idee_conv_desc* d, int dgrad, const float* in, const float* w, const float* bias, const float* relu_src,
                  float* out, void* ws, cudaStream_t st) {
    __nv_bfloat16* wB = (__nv_bfloat16*)ws;
    prep_umma_weights_kernel<<<256, 256, 0, st>>>(w, wB, dgrad);
    IDEE_LAUNCH_CHECK("conv3d(umma) prep");
    UP p{};
    p.in = in; p.out = out; p.bias = bias; p.relu_src = relu_src; p.wB = wB; p.N = d->N;
    if (!dgrad) {
        p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
        p.in_sn = d->x_sn; p.in_st = d->x_st; p.in_sh = d->x_sh; p.in_sw = d->x_sw;
        p.out_sn = d->y_sn; p.out_st = d->y_st; p.out_sh = d->y_sh; p.out_sw = d->y_sw;
        p.relu = d->relu;
    } else {
        p.Ti = d->To; p.Hi = d->Ho; p.Wi = d->Wo; p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
        p.in_sn = d->y_sn; p.in_st = d->y_st; p.in_sh = d->y_sh; p.in_sw = d->y_sw;
        p.out_sn = d->x_sn; p.out_st = d->x_st; p.out_sh = d->x_sh; p.out_sw = d->x_sw;
        p.relu = 0;
    }
    p.tiles_w = (p.Wo + TW - 1) / TW; p.tiles_h = (p.Ho + TH - 1) / TH;
    p.total_tiles = (int64_t)d->N * p.To * p.tiles_h * p.tiles_w;
    const size_t smem = (size_t)3 * (CI / 8) * (HH * 16 * 16 + 16) + (size_t)6 * B_BYTES + 128;
    int64_t grid = idee_num_sms();
    if (grid > p.total_tiles) grid = p.total_tiles;
    if (!dgrad) {
        IDEE_CUDA(cudaFuncSetAttribute(conv_umma_kernel<U_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d(umma)");
        conv_umma_kernel<U_FWD><<<(unsigned)grid, 128, smem, st>>>(p);
    } else {
        IDEE_CUDA(cudaFuncSetAttribute(conv_umma_kernel<U_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d(umma)");
        conv_umma_kernel<U_DGRAD><<<(unsigned)grid, 128, smem, st>>>(p);
    }
    IDEE_LAUNCH_CHECK("conv3d(umma)");
    return 0;
}
