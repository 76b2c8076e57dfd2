// Channel-last 3-D convolutions of the IDEE hot path, fp32 exact path (sm_100a).
//
//   PROJ : Conv3d(16,16,k=3,s=1,p=1,padding_mode='replicate')            Swin_3D.py:586-592   (proj_var)
//   CLS  : Conv3d(Ci,Co,k=(2,3,3),s=(2,1,1),p=(0,1,1)) zero padding       classifier/CNN_3D.py:36-38,83-85
//
// One gather kernel serves forward and data-gradient of both geometries: every output pixel gathers NJ taps,
// each tap contributing a [16 x 16] (in-chunk x out-chunk) product with weights staged in shared memory as
// ws[forward_tap][ci][co].  The data gradient of the replicate-padded conv is written in gather form by
// reflecting out-of-range (pixel, tap) pairs onto the border (the adjoint of clamping), see conv_tap().
// Weights stay in the reference (PyTorch) layout [Co][Ci][kt][kh][kw]; the transposition happens while staging.
// The weight gradient kernel keeps a [4 x 4] register block per (tap, ci-block, co-block) thread, reads
// activations through L1 (every pixel vector is reused by all taps), runs persistent over a pixel range and
// writes per-CTA partials that a second deterministic stage sums into the reference layout.
#include "common.cuh"
#include "idee_b200.h"

// bf16 tensor-core path (conv_tc.cu)
size_t conv_tc_fwd_workspace_bytes(const idee_conv_desc* d);
size_t conv_tc_dgrad_workspace_bytes(const idee_conv_desc* d);
size_t conv_tc_wgrad_workspace_bytes(const idee_conv_desc* d);
int conv_tc_wgrad_splits(const idee_conv_desc* d);
int conv_tc_fwd(const idee_conv_desc* d, const void* x, const float* w, const float* b, void* y, void* ws, cudaStream_t st);
int conv_tc_dgrad(const idee_conv_desc* d, const void* gy, const float* w, const void* relu_src, void* gx, void* ws, cudaStream_t st);
int conv_tc_wgrad_partials(const idee_conv_desc* d, const void* x, const void* gy, float* partials, cudaStream_t st);
bool conv96_umma_eligible(const idee_conv_desc* d);
size_t conv96_wgrad_umma_workspace_bytes(int Cin);
bool conv16to96_wgrad_umma_eligible(const idee_conv_desc* d);
int conv96_wgrad_umma_run(const idee_conv_desc* d, const float* x, const float* gy, float* gw, float* gb, void* ws, cudaStream_t st);
bool conv_tc_bwd_fused_eligible(const idee_conv_desc* d, const void* x, const void* relu_src);
int conv_tc_bwd_fused_splits(const idee_conv_desc* d);
size_t conv_tc_bwd_fused_workspace_bytes(const idee_conv_desc* d);
int conv_tc_bwd_fused_partials(const idee_conv_desc* d, const void* x, const void* gy, const float* w, const void* relu_src, void* gx,
                               float* partials, cudaStream_t st);

namespace {

enum { CLS_FWD = 0, PROJ_FWD = 1, CLS_DGRAD = 2, PROJ_DGRAD = 3 };

struct ConvP {
    const float* in; float* out; const float* w; const float* bias; const float* relu_src;
    int N, V, Vw;
    int Ti, Hi, Wi, To, Ho, Wo;
    int64_t in_sn, in_sv, in_st, in_sh, in_sw, in_sg;
    int64_t out_sn, out_sv, out_st, out_sh, out_sw, out_sg;
    int in_cpg, out_cpg;      // 16-channel chunks per channel group
    int CI, CO;               // gather-input / output channel totals
    int FCI, FCO, KT;         // forward conv geometry: weights [FCO][FCI][KT][3][3]
    int relu;
    int64_t w_set_stride, b_set_stride;
};

struct Tap { bool valid; int64_t off; int ftap; };

// geometry of gather tap j for output pixel (t,h,w)
template <int MODE>
__device__ __forceinline__ Tap conv_tap(const ConvP& p, int j, int t, int h, int w) {
    Tap r;
    if (MODE == CLS_FWD) {
        const int kt = j / 9, kh = (j / 3) % 3, kw = j % 3;
        const int ti = 2 * t + kt, hi = h + kh - 1, wi = w + kw - 1;
        r.valid = ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
        r.off = ti * p.in_st + hi * p.in_sh + wi * p.in_sw;
        r.ftap = j;
    } else if (MODE == PROJ_FWD) {
        const int kt = j / 9, kh = (j / 3) % 3, kw = j % 3;
        const int ti = min(max(t + kt - 1, 0), p.Ti - 1), hi = min(max(h + kh - 1, 0), p.Hi - 1), wi = min(max(w + kw - 1, 0), p.Wi - 1);
        r.valid = true;
        r.off = ti * p.in_st + hi * p.in_sh + wi * p.in_sw;
        r.ftap = j;
    } else if (MODE == CLS_DGRAD) {
        // out = grad wrt conv input at (t,h,w); in = grad wrt conv output (t/2, h', w'); forward kt = t % 2
        const int kh = j / 3, kw = j % 3;
        const int ti = t >> 1, hi = h + kh - 1, wi = w + kw - 1;
        r.valid = ti < p.Ti && hi >= 0 && hi < p.Hi && wi >= 0 && wi < p.Wi;
        r.off = ti * p.in_st + hi * p.in_sh + wi * p.in_sw;
        r.ftap = (t & 1) * 9 + (2 - kh) * 3 + (2 - kw);
    } else {
        // adjoint of replicate padding: pair (pixel p = r + k' - 1, forward tap 2 - k'); an out-of-range pixel is
        // reflected onto the border and then pairs with forward tap k' (the clamped read it stood for).
        int kk[3] = {j / 9, (j / 3) % 3, j % 3};
        const int rr[3] = {t, h, w};
        const int SS[3] = {p.Ti, p.Hi, p.Wi};
        int pp[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int q = rr[a] + kk[a] - 1, kf = 2 - kk[a];
            if (q < 0) { q = 0; kf = kk[a]; }
            else if (q >= SS[a]) { q = SS[a] - 1; kf = kk[a]; }
            pp[a] = q; kk[a] = kf;
        }
        r.valid = true;
        r.off = pp[0] * p.in_st + pp[1] * p.in_sh + pp[2] * p.in_sw;
        r.ftap = kk[0] * 9 + kk[1] * 3 + kk[2];
    }
    return r;
}

template <int MODE> struct ModeInfo {
    static constexpr bool dgrad = (MODE == CLS_DGRAD || MODE == PROJ_DGRAD);
};

constexpr int CONV_THREADS = 128;
constexpr int CONV_PIX = 2;                         // pixels per thread
constexpr int CONV_TILE = CONV_THREADS * CONV_PIX;  // output pixels per CTA

// out[pix][oc*16 + co] = bias + sum_{ic} sum_{j} sum_{ci} in[tap_j(pix)][ic*16 + ci] * ws[ftap][ci][co]
template <int MODE, int COT>
__global__ void __launch_bounds__(CONV_THREADS)
conv_gather_kernel(ConvP p) {
    extern __shared__ __align__(16) float ws[];   // [NT][16][COT]
    const int NT = p.KT * 9;
    const int NJ = (MODE == CLS_DGRAD) ? 9 : NT;
    const int img = blockIdx.y, n = img / p.V, v = img % p.V;
    const int oc = blockIdx.z;
    const int wset = p.Vw == 1 ? 0 : v;
    const float* W = p.w + wset * p.w_set_stride;
    const int64_t npix = (int64_t)p.To * p.Ho * p.Wo;
    const int n_ic = (p.CI + 15) / 16;

    int pt[CONV_PIX], ph[CONV_PIX], pw[CONV_PIX];
    bool pv[CONV_PIX];
    float acc[CONV_PIX][COT];
#pragma unroll
    for (int q = 0; q < CONV_PIX; ++q) {
        const int64_t pix = (int64_t)blockIdx.x * CONV_TILE + q * CONV_THREADS + threadIdx.x;
        pv[q] = pix < npix;
        const int64_t pc = pv[q] ? pix : 0;
        pt[q] = (int)(pc / ((int64_t)p.Ho * p.Wo));
        const int rem = (int)(pc - (int64_t)pt[q] * p.Ho * p.Wo);
        ph[q] = rem / p.Wo; pw[q] = rem - ph[q] * p.Wo;
#pragma unroll
        for (int co = 0; co < COT; ++co) acc[q][co] = 0.f;
    }
    const float* in_img = p.in + n * p.in_sn + v * p.in_sv;

    for (int ic = 0; ic < n_ic; ++ic) {
        __syncthreads();
        // stage weights: forward layout W[o][c][ftap]
        for (int e = threadIdx.x; e < NT * 16 * COT; e += CONV_THREADS) {
            const int co = e % COT, ci = (e / COT) % 16, ft = e / (COT * 16);
            const int gi = ic * 16 + ci, go = oc * 16 + co;     // gather-in / out channel
            float val = 0.f;
            if (gi < p.CI && go < p.CO) {
                const int fo = ModeInfo<MODE>::dgrad ? gi : go, fc = ModeInfo<MODE>::dgrad ? go : gi;
                val = W[((int64_t)fo * p.FCI + fc) * NT + ft];
            }
            ws[e] = val;
        }
        __syncthreads();
        const int64_t coff = (ic / p.in_cpg) * p.in_sg + (ic % p.in_cpg) * 16;
        const int nci = min(16, p.CI - ic * 16);
        for (int j = 0; j < NJ; ++j) {
            float a[CONV_PIX][16];
            const float* wr[CONV_PIX];
#pragma unroll
            for (int q = 0; q < CONV_PIX; ++q) {
                const Tap tp = conv_tap<MODE>(p, j, pt[q], ph[q], pw[q]);
                wr[q] = ws + tp.ftap * 16 * COT;
                if (tp.valid && pv[q]) {
                    const float* src = in_img + tp.off + coff;
                    if (nci == 16) load16(a[q], src);
                    else {
#pragma unroll
                        for (int ci = 0; ci < 16; ++ci) a[q][ci] = ci < nci ? __ldg(src + ci) : 0.f;
                    }
                } else zero16(a[q]);
            }
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) {
                if (COT >= 4) {
#pragma unroll
                    for (int c4 = 0; c4 < COT / 4; ++c4) {
                        if (ModeInfo<MODE>::dgrad) {   // forward tap differs per pixel (border reflection / t parity)
#pragma unroll
                            for (int q = 0; q < CONV_PIX; ++q) {
                                const float4 ww = ld4(wr[q] + ci * COT + 4 * c4);
                                acc[q][4 * c4] += a[q][ci] * ww.x; acc[q][4 * c4 + 1] += a[q][ci] * ww.y;
                                acc[q][4 * c4 + 2] += a[q][ci] * ww.z; acc[q][4 * c4 + 3] += a[q][ci] * ww.w;
                            }
                        } else {
                            const float4 ww = ld4(wr[0] + ci * COT + 4 * c4);
#pragma unroll
                            for (int q = 0; q < CONV_PIX; ++q) {
                                acc[q][4 * c4] += a[q][ci] * ww.x; acc[q][4 * c4 + 1] += a[q][ci] * ww.y;
                                acc[q][4 * c4 + 2] += a[q][ci] * ww.z; acc[q][4 * c4 + 3] += a[q][ci] * ww.w;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < CONV_PIX; ++q) acc[q][0] += a[q][ci] * wr[q][ci * COT];
                }
            }
        }
    }
    // epilogue
    const float* B = p.bias ? p.bias + wset * p.b_set_stride : nullptr;
    const int64_t ooff = (oc / p.out_cpg) * p.out_sg + (oc % p.out_cpg) * 16;
#pragma unroll
    for (int q = 0; q < CONV_PIX; ++q) {
        if (!pv[q]) continue;
        const int64_t o = n * p.out_sn + v * p.out_sv + pt[q] * p.out_st + ph[q] * p.out_sh + pw[q] * p.out_sw + ooff;
        float r[COT];
#pragma unroll
        for (int co = 0; co < COT; ++co) {
            float val = acc[q][co];
            if (B && oc * 16 + co < p.CO) val += B[oc * 16 + co];
            if (p.relu && !ModeInfo<MODE>::dgrad) val = fmaxf(val, 0.f);
            r[co] = val;
        }
        if (p.relu && ModeInfo<MODE>::dgrad) {
#pragma unroll
            for (int co = 0; co < COT; ++co) if (!(p.relu_src[o + co] > 0.f)) r[co] = 0.f;
        }
        if (COT == 16) store16(p.out + o, r);
        else p.out[o] = r[0];
    }
}

// ------------------------------------------------------------------------------------------------------
// weight gradient:  dW[o][c][ftap] = sum_pix gout[pix][o] * in[src(pix, ftap)][c] ;  db[o] = sum_pix gout[pix][o]
// ------------------------------------------------------------------------------------------------------
struct WgradP {
    const float* in; const float* gout; float* partials;
    int N, V, Vw;
    int Ti, Hi, Wi, To, Ho, Wo;
    int64_t in_sn, in_sv, in_st, in_sh, in_sw, in_sg;
    int64_t go_sn, go_sv, go_st, go_sh, go_sw;
    int in_cpg;
    int FCI, FCO, KT, proj;
    int n_ic, n_oc, S;      // chunk counts and pixel-range splits
    int64_t rows_per_set;   // output rows (img,t,h) per weight set
};
// partial layout: [wset][ic][oc][s][NT*256 + 16]   (entry (ft, c, o) at ft*256 + c*16 + o, bias at NT*256 + o)

template <int NT>
__global__ void __launch_bounds__(NT * 16)
conv_wgrad_kernel(WgradP p) {
    const int tid = threadIdx.x;
    const int ft = tid >> 4, cb = (tid >> 2) & 3, ob = tid & 3;
    const int kt = ft / 9, kh = (ft / 3) % 3, kw = ft % 3;
    const int s = blockIdx.x, wset = blockIdx.y;
    const int ic = blockIdx.z / p.n_oc, oc = blockIdx.z % p.n_oc;
    float acc[4][4], accb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { accb[a] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f; }
    const int64_t coff = (ic / p.in_cpg) * p.in_sg + (ic % p.in_cpg) * 16 + cb * 4;
    const int o0 = oc * 16 + ob * 4;
    const bool o_vec = (p.FCO % 4) == 0;
    const bool own_bias = (ft == 0 && cb == 0 && ic == 0);
    const int64_t r0 = p.rows_per_set * s / p.S, r1 = p.rows_per_set * (s + 1) / p.S;
    const int rows_per_img = p.To * p.Ho;
    const int imgs_per_n = p.Vw == 1 ? p.V : 1;     // weight set shared by all variables, or one variable per set
    for (int64_t row = r0; row < r1; ++row) {
        const int64_t img = row / rows_per_img;
        const int rr = (int)(row - img * rows_per_img);
        const int t = rr / p.Ho, h = rr - t * p.Ho;
        const int n = (int)(img / imgs_per_n), v = p.Vw == 1 ? (int)(img % imgs_per_n) : wset;
        int ti, hi; bool rv;
        if (p.proj) { ti = min(max(t + kt - 1, 0), p.Ti - 1); hi = min(max(h + kh - 1, 0), p.Hi - 1); rv = true; }
        else { ti = 2 * t + kt; hi = h + kh - 1; rv = ti < p.Ti && hi >= 0 && hi < p.Hi; }
        const float* in_row = p.in + n * p.in_sn + v * p.in_sv + ti * p.in_st + hi * p.in_sh + coff;
        const float* go_row = p.gout + n * p.go_sn + v * p.go_sv + t * p.go_st + h * p.go_sh;
        for (int w = 0; w < p.Wo; ++w) {
            int wi; bool ok = rv;
            if (p.proj) wi = min(max(w + kw - 1, 0), p.Wi - 1);
            else { wi = w + kw - 1; ok = ok && wi >= 0 && wi < p.Wi; }
            float4 g;
            const float* gp = go_row + w * p.go_sw;
            if (o_vec) g = ldg4(gp + o0);
            else { g.x = (o0 < p.FCO) ? __ldg(gp + o0) : 0.f; g.y = (o0 + 1 < p.FCO) ? __ldg(gp + o0 + 1) : 0.f;
                   g.z = (o0 + 2 < p.FCO) ? __ldg(gp + o0 + 2) : 0.f; g.w = (o0 + 3 < p.FCO) ? __ldg(gp + o0 + 3) : 0.f; }
            if (own_bias) { accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w; }
            if (!ok) continue;
            const float4 a = ldg4(in_row + wi * p.in_sw);
            acc[0][0] += a.x * g.x; acc[0][1] += a.x * g.y; acc[0][2] += a.x * g.z; acc[0][3] += a.x * g.w;
            acc[1][0] += a.y * g.x; acc[1][1] += a.y * g.y; acc[1][2] += a.y * g.z; acc[1][3] += a.y * g.w;
            acc[2][0] += a.z * g.x; acc[2][1] += a.z * g.y; acc[2][2] += a.z * g.z; acc[2][3] += a.z * g.w;
            acc[3][0] += a.w * g.x; acc[3][1] += a.w * g.y; acc[3][2] += a.w * g.z; acc[3][3] += a.w * g.w;
        }
    }
    constexpr int PS = NT * 256 + 16;
    float* part = p.partials + ((((int64_t)wset * p.n_ic + ic) * p.n_oc + oc) * p.S + s) * PS;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) part[ft * 256 + (cb * 4 + a) * 16 + ob * 4 + b] = acc[a][b];
    if (own_bias) {
#pragma unroll
        for (int b = 0; b < 4; ++b) part[NT * 256 + ob * 4 + b] = accb[b];
    }
}

// second stage: a CTA of 8 warps owns 32 consecutive entries of the PARTIAL layout (tap, c, o): lane = entry (consecutive lanes read
// consecutive floats of every partial block), warp w sums the splits s = w, w + 8, ... in order, and the eight warp sums are added
// in a fixed order: deterministic, and the S dependent load latencies of the one-thread-per-entry form become S / 8
__global__ void __launch_bounds__(256)
conv_wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ gw, float* __restrict__ gb,
                         int Vw, int FCI, int FCO, int NT, int n_ic, int n_oc, int S,
                         int64_t w_set_stride, int64_t b_set_stride) {
    __shared__ float red[8][32];
    const int wset = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PS = NT * 256 + 16;
    const int64_t per_set = (int64_t)n_ic * n_oc * PS;
    const int64_t q = (int64_t)blockIdx.x * 32 + lane;                    // (ic, oc, slot)
    int slot = 0, oc = 0, ic = 0;
    bool live = q < per_set;
    int c = 0, o = 0, ft = 0;
    bool is_bias = false;
    if (live) {
        slot = (int)(q % PS); oc = (int)((q / PS) % n_oc); ic = (int)(q / ((int64_t)PS * n_oc));
        if (slot < NT * 256) { ft = slot >> 8; c = ic * 16 + ((slot >> 4) & 15); o = oc * 16 + (slot & 15); live = c < FCI && o < FCO; }
        else { is_bias = true; o = oc * 16 + (slot - NT * 256); live = gb != nullptr && ic == 0 && o < FCO; }
    }
    float acc = 0.f;
    if (live) {
        const float* p = partials + ((((int64_t)wset * n_ic + ic) * n_oc + oc) * S) * PS + slot;
        for (int s = warp; s < S; s += 8) acc += p[(int64_t)s * PS];
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && live) {
        const float v = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
        if (is_bias) gb[wset * b_set_stride + o] = v;
        else gw[wset * w_set_stride + ((int64_t)o * FCI + c) * NT + ft] = v;
    }
}

int fill_common(ConvP& p, const idee_conv_desc* d) {
    p.N = d->N; p.V = d->V; p.Vw = d->Vw;
    p.FCI = d->Cin; p.FCO = d->Cout; p.KT = d->proj ? 3 : 2;
    p.w_set_stride = (int64_t)d->Cin * d->Cout * p.KT * 9; p.b_set_stride = d->Cout;
    return 0;
}

int wgrad_splits(const idee_conv_desc* d) {
    const int n_ic = (d->Cin + 15) / 16, n_oc = (d->Cout + 15) / 16;
    int S = (idee_num_sms() * 3 + d->Vw * n_ic * n_oc - 1) / (d->Vw * n_ic * n_oc);
    if (S < 1) S = 1;
    if (S > 256) S = 256;
    return S;
}

int check_desc(const idee_conv_desc* d, const char* who) {
    IDEE_REQUIRE(d->Cin == 16 || d->Cin == 96 || d->Cin % 16 == 0, "%s: Cin must be a multiple of 16 (got %d)", who, d->Cin);
    IDEE_REQUIRE(d->Cout == 1 || d->Cout % 16 == 0, "%s: Cout must be 1 or a multiple of 16 (got %d)", who, d->Cout);
    IDEE_REQUIRE(d->Vw == 1 || d->Vw == d->V, "%s: Vw must be 1 or V", who);
    IDEE_REQUIRE(d->in_cpg >= 1 && d->out_cpg >= 1, "%s: chunks-per-group must be >= 1", who);
    if (d->proj) IDEE_REQUIRE(d->To == d->Ti && d->Ho == d->Hi && d->Wo == d->Wi, "%s: proj conv keeps the shape", who);
    else IDEE_REQUIRE(d->To == (d->Ti - 2) / 2 + 1 && d->Ti >= 2 && d->Ho == d->Hi && d->Wo == d->Wi, "%s: cls conv output shape mismatch", who);
    IDEE_REQUIRE((unsigned)d->x_dtype <= 1u && (unsigned)d->y_dtype <= 1u && (unsigned)d->gx_dtype <= 1u, "%s: dtype fields must be 0 (float) or 1 (bf16)", who);
    if (d->x_dtype || d->y_dtype || d->gx_dtype)
        IDEE_REQUIRE(d->precision >= 1 && d->Cin == 16 && (d->Cout == 16 || (d->proj && d->Cout == 1)) && d->in_cpg == 1 && d->out_cpg == 1,
                     "%s: bf16 activation storage is only built for the precision >= 1 16 -> 16 convs and the 16 -> 1 proj conv", who);
    return 0;
}

template <int MODE, int COT>
int launch_gather(const ConvP& p, cudaStream_t st, const char* who) {
    const size_t smem = sizeof(float) * p.KT * 9 * 16 * COT;
    auto kern = conv_gather_kernel<MODE, COT>;
    IDEE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), who);
    const int64_t npix = (int64_t)p.To * p.Ho * p.Wo;
    dim3 grid((unsigned)((npix + CONV_TILE - 1) / CONV_TILE), p.N * p.V, (p.CO + 15) / 16);
    kern<<<grid, CONV_THREADS, smem, st>>>(p);
    IDEE_LAUNCH_CHECK(who);
    return 0;
}

}  // namespace

extern "C" size_t idee_conv3d_fwd_workspace_bytes(const idee_conv_desc* d) { return d->precision >= 1 ? conv_tc_fwd_workspace_bytes(d) : 0; }
extern "C" size_t idee_conv3d_dgrad_workspace_bytes(const idee_conv_desc* d) { return d->precision >= 1 ? conv_tc_dgrad_workspace_bytes(d) : 0; }

extern "C" int idee_conv3d_fwd(const idee_conv_desc* d, const void* x, const float* w, const float* b, void* y,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d, "conv3d_fwd")) return 1;
    IDEE_REQUIRE(workspace_bytes >= idee_conv3d_fwd_workspace_bytes(d), "conv3d_fwd: workspace too small");
    if (d->precision >= 1) return conv_tc_fwd(d, x, w, b, y, workspace, (cudaStream_t)stream);
    ConvP p{};
    fill_common(p, d);
    p.in = (const float*)x; p.out = (float*)y; p.w = w; p.bias = b; p.relu_src = nullptr; p.relu = d->relu;
    p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
    p.in_sn = d->x_sn; p.in_sv = d->x_sv; p.in_st = d->x_st; p.in_sh = d->x_sh; p.in_sw = d->x_sw; p.in_sg = d->x_sg; p.in_cpg = d->in_cpg;
    p.out_sn = d->y_sn; p.out_sv = d->y_sv; p.out_st = d->y_st; p.out_sh = d->y_sh; p.out_sw = d->y_sw; p.out_sg = d->y_sg; p.out_cpg = d->out_cpg;
    p.CI = d->Cin; p.CO = d->Cout;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->proj) return launch_gather<PROJ_FWD, 16>(p, st, "conv3d_fwd(proj)");
    if (d->Cout == 1) return launch_gather<CLS_FWD, 1>(p, st, "conv3d_fwd(cls,Co=1)");
    return launch_gather<CLS_FWD, 16>(p, st, "conv3d_fwd(cls)");
}

// gx = conv^T(gy); if relu_src != NULL the result is multiplied by (relu_src > 0) (relu_src has gx's layout)
extern "C" int idee_conv3d_dgrad(const idee_conv_desc* d, const void* gy, const float* w, const void* relu_src, void* gx,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d, "conv3d_dgrad")) return 1;
    IDEE_REQUIRE(workspace_bytes >= idee_conv3d_dgrad_workspace_bytes(d), "conv3d_dgrad: workspace too small");
    if (d->precision >= 1) return conv_tc_dgrad(d, gy, w, relu_src, gx, workspace, (cudaStream_t)stream);
    ConvP p{};
    fill_common(p, d);
    p.in = (const float*)gy; p.out = (float*)gx; p.w = w; p.bias = nullptr; p.relu_src = (const float*)relu_src; p.relu = relu_src != nullptr;
    // gather-input = gradient wrt the forward output, output = gradient wrt the forward input
    p.Ti = d->To; p.Hi = d->Ho; p.Wi = d->Wo; p.To = d->Ti; p.Ho = d->Hi; p.Wo = d->Wi;
    p.in_sn = d->y_sn; p.in_sv = d->y_sv; p.in_st = d->y_st; p.in_sh = d->y_sh; p.in_sw = d->y_sw; p.in_sg = d->y_sg; p.in_cpg = d->out_cpg;
    p.out_sn = d->x_sn; p.out_sv = d->x_sv; p.out_st = d->x_st; p.out_sh = d->x_sh; p.out_sw = d->x_sw; p.out_sg = d->x_sg; p.out_cpg = d->in_cpg;
    p.CI = d->Cout; p.CO = d->Cin;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->proj) return launch_gather<PROJ_DGRAD, 16>(p, st, "conv3d_dgrad(proj)");
    return launch_gather<CLS_DGRAD, 16>(p, st, "conv3d_dgrad(cls)");
}

extern "C" size_t idee_conv3d_wgrad_workspace_bytes(const idee_conv_desc* d) {
    if (d->precision >= 1 && (conv96_umma_eligible(d) || conv16to96_wgrad_umma_eligible(d))) return conv96_wgrad_umma_workspace_bytes(d->Cin);
    if (d->precision >= 1) return conv_tc_wgrad_workspace_bytes(d);
    const int n_ic = (d->Cin + 15) / 16, n_oc = (d->Cout + 15) / 16, NT = (d->proj ? 3 : 2) * 9;
    return sizeof(float) * (size_t)d->Vw * n_ic * n_oc * wgrad_splits(d) * (NT * 256 + 16);
}

extern "C" int idee_conv3d_wgrad(const idee_conv_desc* d, const void* x, const void* gy, float* gw, float* gb,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d, "conv3d_wgrad")) return 1;
    IDEE_REQUIRE(workspace_bytes >= idee_conv3d_wgrad_workspace_bytes(d), "conv3d_wgrad: workspace too small");
    IDEE_REQUIRE(d->out_cpg * 16 >= d->Cout || d->Cout == 1, "conv3d_wgrad: grouped output layout is not supported");
    if (d->precision >= 1 && (conv96_umma_eligible(d) || conv16to96_wgrad_umma_eligible(d)))   // 96 / 16 -> 96 classifier convs: tcgen05 / TMEM accumulation over pixels
        return conv96_wgrad_umma_run(d, (const float*)x, (const float*)gy, gw, gb, workspace, (cudaStream_t)stream);
    if (d->precision >= 1) {
        cudaStream_t st = (cudaStream_t)stream;
        if (conv_tc_wgrad_partials(d, x, gy, (float*)workspace, st)) return 2;
        const int NT = (d->proj ? 3 : 2) * 9;
        const int64_t nel = (int64_t)((d->Cin + 15) / 16) * ((d->Cout + 15) / 16) * (NT * 256 + 16);
        conv_wgrad_reduce_kernel<<<dim3((unsigned)((nel + 31) / 32), d->Vw), 256, 0, st>>>(
            (const float*)workspace, gw, gb, d->Vw, d->Cin, d->Cout, NT, (d->Cin + 15) / 16, (d->Cout + 15) / 16, conv_tc_wgrad_splits(d),
            (int64_t)d->Cin * d->Cout * NT, d->Cout);
        IDEE_LAUNCH_CHECK("conv3d_wgrad_reduce");
        return 0;
    }
    WgradP p{};
    p.in = (const float*)x; p.gout = (const float*)gy; p.partials = (float*)workspace;
    p.N = d->N; p.V = d->V; p.Vw = d->Vw;
    p.Ti = d->Ti; p.Hi = d->Hi; p.Wi = d->Wi; p.To = d->To; p.Ho = d->Ho; p.Wo = d->Wo;
    p.in_sn = d->x_sn; p.in_sv = d->x_sv; p.in_st = d->x_st; p.in_sh = d->x_sh; p.in_sw = d->x_sw; p.in_sg = d->x_sg; p.in_cpg = d->in_cpg;
    p.go_sn = d->y_sn; p.go_sv = d->y_sv; p.go_st = d->y_st; p.go_sh = d->y_sh; p.go_sw = d->y_sw;
    p.FCI = d->Cin; p.FCO = d->Cout; p.KT = d->proj ? 3 : 2; p.proj = d->proj;
    p.n_ic = (d->Cin + 15) / 16; p.n_oc = (d->Cout + 15) / 16; p.S = wgrad_splits(d);
    p.rows_per_set = (int64_t)d->N * (d->Vw == 1 ? d->V : 1) * d->To * d->Ho;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(p.S, d->Vw, p.n_ic * p.n_oc);
    const int NT = p.KT * 9;
    if (d->proj) conv_wgrad_kernel<27><<<grid, 27 * 16, 0, st>>>(p);
    else conv_wgrad_kernel<18><<<grid, 18 * 16, 0, st>>>(p);
    IDEE_LAUNCH_CHECK("conv3d_wgrad");
    const int64_t nel = (int64_t)p.n_ic * p.n_oc * (NT * 256 + 16);
    conv_wgrad_reduce_kernel<<<dim3((unsigned)((nel + 31) / 32), d->Vw), 256, 0, st>>>(
        p.partials, gw, gb, d->Vw, d->Cin, d->Cout, NT, p.n_ic, p.n_oc, p.S, (int64_t)d->Cin * d->Cout * NT, d->Cout);
    IDEE_LAUNCH_CHECK("conv3d_wgrad_reduce");
    return 0;
}

// Whole backward of one conv in one call.  The 16 -> 1 proj conv on bf16 storage runs ONE kernel for both gradients (conv_tc.cu,
// proj_bwd_scalar_kernel); every other geometry runs the weight-gradient and data-gradient entry points back to back on the
// same workspace.
extern "C" size_t idee_conv3d_bwd_workspace_bytes(const idee_conv_desc* d) {
    size_t a = idee_conv3d_wgrad_workspace_bytes(d), b = idee_conv3d_dgrad_workspace_bytes(d);
    if (b > a) a = b;
    if (d->precision >= 1 && d->proj && d->Cin == 16 && d->Cout == 1) { b = conv_tc_bwd_fused_workspace_bytes(d); if (b > a) a = b; }
    return a;
}

extern "C" int idee_conv3d_bwd(const idee_conv_desc* d, const void* x, const void* gy, const float* w, const void* relu_src, void* gx,
                               float* gw, float* gb, void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d, "conv3d_bwd")) return 1;
    IDEE_REQUIRE(workspace_bytes >= idee_conv3d_bwd_workspace_bytes(d), "conv3d_bwd: workspace too small");
    if (conv_tc_bwd_fused_eligible(d, x, relu_src)) {
        cudaStream_t st = (cudaStream_t)stream;
        if (conv_tc_bwd_fused_partials(d, x, gy, w, relu_src, gx, (float*)workspace, st)) return 2;
        const int64_t nel = 27 * 256 + 16;
        conv_wgrad_reduce_kernel<<<dim3((unsigned)((nel + 31) / 32), d->Vw), 256, 0, st>>>(
            (const float*)workspace, gw, gb, d->Vw, d->Cin, d->Cout, 27, 1, 1, conv_tc_bwd_fused_splits(d), (int64_t)d->Cin * d->Cout * 27, d->Cout);
        IDEE_LAUNCH_CHECK("conv3d_bwd reduce");
        return 0;
    }
    if (idee_conv3d_wgrad(d, x, gy, gw, gb, workspace, workspace_bytes, stream)) return 2;
    return idee_conv3d_dgrad(d, gy, w, relu_src, gx, workspace, workspace_bytes, stream);
}
