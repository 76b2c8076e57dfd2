// PatchEmbed3D with patch (1,1,1): Conv3d(Cin->16, k=1, bias) + LayerNorm(16, no affine), written straight into the
// channel-last token layout [N,V,T,H,W,16] that the Swin block kernels consume (sm_100a, fp32).
// Replaces PatchEmbed3D.forward (Swin_3D.py:473-491) and the NCDHW->NDHWC rearrange of BasicLayer.forward (:434).
#include "common.cuh"
#include "idee_b200.h"

namespace {

constexpr int C = 16;
constexpr int MAXCIN = 4;
constexpr int EMB_THREADS = 256;

struct EmbP {
    const float* x; int64_t xs_n, xs_v, xs_c, xs_t, xs_h, xs_w;
    const float* w; const float* b;   // [V][16][Cin], [V][16]
    int N, V, Cin, T, H, W;
};

// Offset of in-image token r (< T*H*W, 32-bit) inside image (n, v) of x.  Dense (t,h,w) strides need no division at all.
__device__ __forceinline__ int64_t x_offset(const EmbP& p, uint32_t r, bool dense) {
    if (dense) return (int64_t)r * p.xs_w;
    const uint32_t hw = (uint32_t)(p.H * p.W);
    const uint32_t t = r / hw, r2 = r - t * hw, h = r2 / (uint32_t)p.W, w = r2 - h * (uint32_t)p.W;
    return t * p.xs_t + h * p.xs_h + w * p.xs_w;
}
__device__ __forceinline__ bool x_dense(const EmbP& p) { return p.xs_h == p.W * p.xs_w && p.xs_t == (int64_t)p.H * p.W * p.xs_w; }

__global__ void __launch_bounds__(EMB_THREADS)
embed_ln_fwd_kernel(EmbP p, float* __restrict__ y) {
    const int v = blockIdx.y;
    float wr[C][MAXCIN], br[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        br[c] = __ldg(p.b + v * C + c);
#pragma unroll
        for (int ci = 0; ci < MAXCIN; ++ci) wr[c][ci] = ci < p.Cin ? __ldg(p.w + (v * C + c) * p.Cin + ci) : 0.f;
    }
    // grid = (blocks, V, N): a block strides over the T*H*W tokens of one (n, v) image with 32-bit indices
    const int n = blockIdx.z;
    const uint32_t thw = (uint32_t)(p.T * p.H * p.W);
    const bool dense = x_dense(p);
    const float* ximg = p.x + n * p.xs_n + v * p.xs_v;
    float* yimg = y + ((int64_t)n * p.V + v) * thw * C;
    __shared__ __align__(16) float y_stage[EMB_THREADS / 32][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // whole warps iterate together: the 32 consecutive token rows of a warp are written with coalesced 512-byte stores
    for (uint32_t r0 = blockIdx.x * EMB_THREADS + warp * 32; r0 < thw; r0 += gridDim.x * EMB_THREADS) {
        const uint32_t r = r0 + lane;
        const bool live = r < thw;
        float xin[MAXCIN];
        if (live) {
            const int64_t xo = x_offset(p, r, dense);
#pragma unroll
            for (int ci = 0; ci < MAXCIN; ++ci) xin[ci] = ci < p.Cin ? __ldg(ximg + xo + ci * p.xs_c) : 0.f;
        } else {
#pragma unroll
            for (int ci = 0; ci < MAXCIN; ++ci) xin[ci] = 0.f;
        }
        float e[C], en[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float a = br[c];
#pragma unroll
            for (int ci = 0; ci < MAXCIN; ++ci) a += wr[c][ci] * xin[ci];
            e[c] = a;
        }
        ln16(e, en);
        store16_warp(yimg + (int64_t)r0 * C, en, y_stage[warp], lane, (int)min(32u, thw - r0));
    }
}

constexpr int EMB_NG = C * MAXCIN + C;

template <int CIN>
__global__ void __launch_bounds__(EMB_THREADS)
embed_ln_bwd_kernel(EmbP p, const float* __restrict__ gy, float* __restrict__ partials) {
    const int v = blockIdx.y;
    float wr[C][CIN], br[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        br[c] = __ldg(p.b + v * C + c);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) wr[c][ci] = __ldg(p.w + (v * C + c) * CIN + ci);
    }
    constexpr int NG = C * CIN + C;
    float acc[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) acc[k] = 0.f;
    // blocks of one variable stride over the (n, in-image token) pairs: n advances only when a block wraps an image
    const uint32_t thw = (uint32_t)(p.T * p.H * p.W);
    const bool dense = x_dense(p);
    const uint32_t per_img = (thw + EMB_THREADS - 1) / EMB_THREADS;            // chunks of EMB_THREADS tokens per image
    // software pipeline: the loads of the next chunk (4 B of x, 64 B of gy per thread) are in flight while this one is reduced
    const bool gy_a32 = (reinterpret_cast<uintptr_t>(gy) & 31) == 0;
    const uint32_t nchunk = per_img * (uint32_t)p.N;
    auto fetch = [&](uint32_t chunk, float (&xin)[CIN], float (&g)[C]) -> bool {
        const uint32_t n = chunk / per_img, r = (chunk - n * per_img) * EMB_THREADS + threadIdx.x;
        if (chunk >= nchunk || r >= thw) return false;
        const float* ximg = p.x + n * p.xs_n + v * p.xs_v;
        const int64_t xo = x_offset(p, r, dense);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xin[ci] = __ldg(ximg + xo + ci * p.xs_c);
        load16_a(g, gy + (((int64_t)n * p.V + v) * thw + r) * C, gy_a32);
        return true;
    };
    float xin[CIN], g[C];
    bool have = fetch(blockIdx.x, xin, g);
    for (uint32_t chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        float xin_n[CIN], g_n[C];
        const bool have_n = fetch(chunk + gridDim.x, xin_n, g_n);
        if (have) {
            float e[C], en[C], ge[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float a = br[c];
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) a += wr[c][ci] * xin[ci];
                e[c] = a;
            }
            const float rstd = ln16(e, en);
            ln16_bwd(g, en, rstd, ge);
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) acc[c * CIN + ci] += ge[c] * xin[ci];
                acc[C * CIN + c] += ge[c];
            }
        }
        have = have_n;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xin[ci] = xin_n[ci];
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = g_n[c];
    }
    __shared__ float red[EMB_THREADS / 32][NG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < NG) {
        float s = 0.f;
        for (int w = 0; w < EMB_THREADS / 32; ++w) s += red[w][threadIdx.x];
        // partial layout is the MAXCIN one: weight (c, ci) at c*MAXCIN + ci, bias c at C*MAXCIN + c
        const int k = threadIdx.x;
        const int dst = k < C * CIN ? (k / CIN) * MAXCIN + (k % CIN) : C * MAXCIN + (k - C * CIN);
        partials[((int64_t)v * gridDim.x + blockIdx.x) * EMB_NG + dst] = s;
    }
}

__global__ void embed_ln_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int Cin, float* __restrict__ gw,
                                             float* __restrict__ gb) {
    const int v = blockIdx.x, k = threadIdx.x;
    if (k >= EMB_NG) return;
    if (k < C * MAXCIN && (k % MAXCIN) >= Cin) return;      // slot not produced for this Cin
    double a = 0.0;
    for (int b = 0; b < nblocks; ++b) a += partials[((int64_t)v * nblocks + b) * EMB_NG + k];
    if (k < C * MAXCIN) {
        const int c = k / MAXCIN, ci = k % MAXCIN;
        gw[(v * C + c) * Cin + ci] = (float)a;
    } else gb[v * C + (k - C * MAXCIN)] = (float)a;
}

int emb_blocks(int V) { int b = (idee_num_sms() * 4 + V - 1) / V; return b < 1 ? 1 : b; }

int fill(EmbP& p, const float* x, const int64_t* xs, const float* w, const float* b, int N, int V, int Cin, int T, int H, int W, int E) {
    IDEE_REQUIRE(E == C, "embed_ln: only embed_dim=16 is built (got %d)", E);
    IDEE_REQUIRE(Cin >= 1 && Cin <= MAXCIN, "embed_ln: in_chans must be in [1,%d] (got %d)", MAXCIN, Cin);
    p.x = x; p.xs_n = xs[0]; p.xs_v = xs[1]; p.xs_c = xs[2]; p.xs_t = xs[3]; p.xs_h = xs[4]; p.xs_w = xs[5];
    p.w = w; p.b = b; p.N = N; p.V = V; p.Cin = Cin; p.T = T; p.H = H; p.W = W;
    return 0;
}

}  // namespace

extern "C" size_t idee_embed_ln_bwd_workspace_bytes(int V) { return sizeof(float) * (size_t)V * emb_blocks(V) * EMB_NG; }

extern "C" int idee_embed_ln_fwd(const float* x, const int64_t* x_strides, const float* w, const float* b, float* y, int N, int V,
                                 int Cin, int T, int H, int W, int E, void* stream) {
    EmbP p;
    if (fill(p, x, x_strides, w, b, N, V, Cin, T, H, W, E)) return 1;
    IDEE_REQUIRE((int64_t)T * H * W < (1ll << 31) / C && N <= 65535, "embed_ln: image too large for 32-bit token indices");
    int bx = (emb_blocks(V) + N - 1) / N;
    const int need = (int)(((int64_t)T * H * W + EMB_THREADS - 1) / EMB_THREADS);
    if (bx > need) bx = need;
    if (bx < 1) bx = 1;
    embed_ln_fwd_kernel<<<dim3(bx, V, N), EMB_THREADS, 0, (cudaStream_t)stream>>>(p, y);
    IDEE_LAUNCH_CHECK("embed_ln_fwd");
    return 0;
}

extern "C" int idee_embed_ln_bwd(const float* x, const int64_t* x_strides, const float* w, const float* b, const float* gy, float* gw,
                                 float* gb, int N, int V, int Cin, int T, int H, int W, int E, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    EmbP p;
    if (fill(p, x, x_strides, w, b, N, V, Cin, T, H, W, E)) return 1;
    IDEE_REQUIRE(workspace_bytes >= idee_embed_ln_bwd_workspace_bytes(V), "embed_ln_bwd: workspace too small");
    IDEE_REQUIRE((int64_t)T * H * W < (1ll << 31) / C, "embed_ln: image too large for 32-bit token indices");
    const int nb = emb_blocks(V);
    if (Cin == 1) embed_ln_bwd_kernel<1><<<dim3(nb, V), EMB_THREADS, 0, (cudaStream_t)stream>>>(p, gy, (float*)workspace);
    else if (Cin == 2) embed_ln_bwd_kernel<2><<<dim3(nb, V), EMB_THREADS, 0, (cudaStream_t)stream>>>(p, gy, (float*)workspace);
    else if (Cin == 3) embed_ln_bwd_kernel<3><<<dim3(nb, V), EMB_THREADS, 0, (cudaStream_t)stream>>>(p, gy, (float*)workspace);
    else embed_ln_bwd_kernel<4><<<dim3(nb, V), EMB_THREADS, 0, (cudaStream_t)stream>>>(p, gy, (float*)workspace);
    IDEE_LAUNCH_CHECK("embed_ln_bwd");
    embed_ln_bwd_finalize_kernel<<<V, 128, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, Cin, gw, gb);
    IDEE_LAUNCH_CHECK("embed_ln_bwd_finalize");
    return 0;
}
