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
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
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

template <int MODE>
__global__ void __launch_bounds__(128, 1)
conv_umma_kernel(UP p) {
    constexpr int KTIN = MODE == U_FWD ? 2 : 1;
    constexpr int NJ = MODE == U_FWD ? 18 : 9;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(smem_raw);                          // [KTIN*HH*HWp][CPH]
    unsigned char* Abuf = smem_raw + ((KTIN * HH * HWp * CPH * 2 + 1023) / 1024) * 1024;       // 2 x A_BYTES
    unsigned char* Bbuf = Abuf + 2 * A_BYTES;                                                  // 2 x B_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bbuf + 2 * B_BYTES);                          // [2]
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
    uint32_t par0 = 0u, par1 = 0u;     // mbarrier phase parity of the two A/B slots

    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int64_t rr = tile;
        const int w0 = (int)(rr % p.tiles_w) * TW; rr /= p.tiles_w;
        const int h0 = (int)(rr % p.tiles_h) * TH; rr /= p.tiles_h;
        const int t = (int)(rr % p.To);
        const int n = (int)(rr / p.To);
        // ---- halo: fp32 HBM -> bf16 smem, zero padding ----
        {
            const float* in_img = p.in + n * p.in_sn;
            constexpr int V4 = CI / 4, NCOL = HWp * V4, NROW = KTIN * HH, DROW = 128 / NCOL, DCOL = 128 - DROW * NCOL;
            int row = tid / NCOL, col = tid - row * NCOL;
            while (row < NROW) {
                const float* src[4];
                int dst[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    src[u] = nullptr; dst[u] = -1;
                    if (row < NROW) {
                        const int kt = row / HH, hh = row - kt * HH;
                        const int ww = col / V4, c4 = col - ww * V4;
                        const int ti = MODE == U_FWD ? 2 * t + kt : (t >> 1);
                        const int hi = h0 + hh - 1, wi = w0 + ww - 1;
                        const bool ok = ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
                        dst[u] = (row * HWp + ww) * CPH + c4 * 4;
                        if (ok) src[u] = in_img + ti * p.in_st + hi * p.in_sh + wi * p.in_sw + c4 * 4;
                        row += DROW; col += DCOL;
                        if (col >= NCOL) { col -= NCOL; ++row; }
                    }
                }
                float4 f[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) f[u] = src[u] ? ldg4(src[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (dst[u] >= 0) *reinterpret_cast<uint2*>(halo + dst[u]) = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
            }
        }
        __syncthreads();
        // ---- K loop over taps ----
        const int pr = tid >> 4, pc = tid & 15;              // this thread's pixel (row, col) inside the tile = TMEM lane tid
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
            const int b = j & 1;
            if (j >= 2) {                                   // MMAs of tap j-2 released slot b
                mbar_wait(smem_u32(&bars[b]), b ? par1 : par0);
                if (b) par1 ^= 1u; else par0 ^= 1u;
            }
            int kt, kh, kw, ft;
            if (MODE == U_FWD) { kt = j / 9; kh = (j / 3) % 3; kw = j % 3; ft = j; }
            else { kt = 0; kh = j / 3; kw = j % 3; ft = (t & 1) * 9 + (2 - kh) * 3 + (2 - kw); }
            // B tap (18 KB, L2 resident) -> slot b
            {
                const uint4* src = reinterpret_cast<const uint4*>(p.wB + (size_t)ft * CO * CI);
                uint4* dst = reinterpret_cast<uint4*>(Bbuf + b * B_BYTES);
#pragma unroll
                for (int i = 0; i < B_BYTES / 16 / 128; ++i) dst[tid + i * 128] = __ldg(src + tid + i * 128);
            }
            // A tap: this thread's pixel row (96 channels = 12 x 16 B) -> canonical layout
            {
                const uint4* src = reinterpret_cast<const uint4*>(halo + ((kt * HH + pr + kh) * HWp + pc + kw) * CPH);
                unsigned char* dst = Abuf + b * A_BYTES + (tid >> 3) * SBO + (tid & 7) * 16;
#pragma unroll
                for (int kc = 0; kc < CI / 8; ++kc) *reinterpret_cast<uint4*>(dst + kc * A_LBO) = src[kc];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(Abuf + b * A_BYTES), b0 = smem_u32(Bbuf + b * B_BYTES);
#pragma unroll
                for (int ks = 0; ks < CI / 16; ++ks)
                    umma_bf16(tmem_base, make_desc(a0 + ks * 2 * A_LBO, A_LBO, SBO), make_desc(b0 + ks * 2 * B_LBO, B_LBO, SBO),
                              (j > 0 || ks > 0) ? 1u : 0u);
                umma_commit(smem_u32(&bars[b]));
            }
        }
        // drain: the last two taps' commits
#pragma unroll
        for (int jj = NJ - 2; jj < NJ; ++jj) {
            const int b = jj & 1;
            mbar_wait(smem_u32(&bars[b]), b ? par1 : par0);
            if (b) par1 ^= 1u; else par0 ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: TMEM lane = pixel, 96 columns = output channels ----
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
                    if (p.bias) { const float4 bb = ldg4(p.bias + c0 + i); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
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
        __syncthreads();                                  // every warp has drained the accumulator and the halo
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
    const int KTIN = dgrad ? 1 : 2;
    const size_t smem = ((size_t)(KTIN * HH * HWp * CPH * 2 + 1023) / 1024) * 1024 + 2 * A_BYTES + 2 * B_BYTES + 64;
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
