// tcgen05 feature probe (sm_100a): checks, against a host reference, the operand forms the Swin kernels rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe <test>
// tests:  k      K-major SS, M=128, N=48, K=16 and K=64 (known-good form, canonical SWIZZLE_NONE layout)
//         mn0/1  MN-major A and B (transposed views of token-major chunk planes), M=128, N=48, K=128; variant 0: SBO = MN-group
//                stride, LBO = K-group stride; variant 1: swapped
//         m64    MN-major, M=64: dumps which TMEM lane holds which row
//         ts     A operand from TMEM (tcgen05.st 32x32b.x8 packed bf16 pairs), B from smem
// Each test runs in its own process (a bad descriptor kills the context).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// mode 0: K-major (A [128 x K] row r chunk c at c*2048 + r*16 ; B [N x K] row n chunk c at c*N*16 + n*16), K = 16*ksteps
// mode 1/2: MN-major variants; A = X^T with X [128 tok x MA feat] stored chunk c at c*2048 + tok*16, B = Y [128 tok x 48]
// mode 3: TS
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* X, const __nv_bfloat16* Y, float* D, int mode, int M, int N, int ksteps,
                                             int XC, int YC) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* sa = sm;                       // X planes
    unsigned char* sb = sm + 16 * 2048;           // Y planes
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    // stage operands: X [128][XC*8] row-major in gmem -> chunk planes; Y [rows][YC*8]
    for (int c = 0; c < XC; ++c) *reinterpret_cast<uint4*>(sa + c * 2048 + tid * 16) = *reinterpret_cast<const uint4*>(X + (size_t)tid * XC * 8 + c * 8);
    if (mode == 0 || mode == 3) {
        for (int c = 0; c < YC; ++c)
            if (tid < N) *reinterpret_cast<uint4*>(sb + c * N * 16 + tid * 16) = *reinterpret_cast<const uint4*>(Y + (size_t)tid * YC * 8 + c * 8);
    } else {
        for (int c = 0; c < YC; ++c) *reinterpret_cast<uint4*>(sb + c * 2048 + tid * 16) = *reinterpret_cast<const uint4*>(Y + (size_t)tid * YC * 8 + c * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    if (mode == 3) {
        // A rows into TMEM columns [64, 64 + 8*ksteps): column j of lane r = elements (2j, 2j+1) of row r
        for (int s = 0; s < ksteps; ++s) {
            uint32_t r[8];
            const uint4 lo = *reinterpret_cast<const uint4*>(sa + (2 * s) * 2048 + tid * 16), hi = *reinterpret_cast<const uint4*>(sa + (2 * s + 1) * 2048 + tid * 16);
            r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w; r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + 64 + 8 * s),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        if (mode == 0) {
            for (int s = 0; s < ksteps; ++s)
                umma_ss(tm, make_desc(smem_u32(sa) + 2 * s * 2048, 2048, 128), make_desc(smem_u32(sb) + 2 * s * N * 16, N * 16, 128), idesc(M, N, 0, 0), s > 0);
        } else if (mode == 1 || mode == 2) {
            const uint32_t lbo = mode == 1 ? 128 : 2048, sbo = mode == 1 ? 2048 : 128;
            for (int s = 0; s < ksteps; ++s)
                umma_ss(tm, make_desc(smem_u32(sa) + s * 256, lbo, sbo), make_desc(smem_u32(sb) + s * 256, lbo, sbo), idesc(M, N, 1, 1), s > 0);
        } else {
            for (int s = 0; s < ksteps; ++s)
                umma_ts(tm, tm + 64 + 8 * s, make_desc(smem_u32(sb) + 2 * s * N * 16, N * 16, 128), idesc(M, N, 0, 0), s > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(128) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
    const char* t = argc > 1 ? argv[1] : "k";
    int mode = 0, M = 128, N = 48, ksteps = 1, XC = 2, YC = 2;
    if (!strcmp(t, "k")) { mode = 0; ksteps = 1; XC = 2; YC = 2; }
    else if (!strcmp(t, "k64")) { mode = 0; ksteps = 4; XC = 8; YC = 8; N = 16; }
    else if (!strcmp(t, "mn0")) { mode = 1; ksteps = 8; XC = 16; YC = 6; }
    else if (!strcmp(t, "mn1")) { mode = 2; ksteps = 8; XC = 16; YC = 6; }
    else if (!strcmp(t, "m64")) { mode = 1; M = 64; ksteps = 8; XC = 8; YC = 6; }
    else if (!strcmp(t, "m64b")) { mode = 2; M = 64; ksteps = 8; XC = 8; YC = 6; }
    else if (!strcmp(t, "ts")) { mode = 3; ksteps = 1; XC = 2; YC = 2; }
    else if (!strcmp(t, "ts64")) { mode = 3; ksteps = 4; XC = 8; YC = 8; N = 16; }
    else { printf("unknown test %s\n", t); return 2; }
    const int yrows = (mode == 0 || mode == 3) ? N : 128;
    std::vector<float> X(128 * XC * 8), Y((size_t)yrows * YC * 8);
    srand(7);
    for (auto& v : X) v = bf((rand() % 2001 - 1000) / 500.f);
    for (auto& v : Y) v = bf((rand() % 2001 - 1000) / 500.f);
    std::vector<__nv_bfloat16> Xb(X.size()), Yb(Y.size());
    for (size_t i = 0; i < X.size(); ++i) Xb[i] = __float2bfloat16(X[i]);
    for (size_t i = 0; i < Y.size(); ++i) Yb[i] = __float2bfloat16(Y[i]);
    __nv_bfloat16 *dX, *dY;
    float* dD;
    cudaMalloc(&dX, Xb.size() * 2); cudaMalloc(&dY, Yb.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dX, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, Yb.data(), Yb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * N * 4);
    const int smem = 16 * 2048 + 8 * 2048 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dX, dY, dD, mode, M, N, ksteps, XC, YC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    // reference
    const int K = (mode == 0 || mode == 3) ? ksteps * 16 : 128;
    const int Mrows = (mode == 0 || mode == 3) ? 128 : M;
    std::vector<double> R((size_t)Mrows * N, 0.0);
    for (int m = 0; m < Mrows; ++m)
        for (int n = 0; n < N; ++n) {
            double a = 0;
            for (int k = 0; k < K; ++k)
                a += (mode == 0 || mode == 3) ? (double)X[(size_t)m * XC * 8 + k] * Y[(size_t)n * YC * 8 + k]
                                              : (double)X[(size_t)k * XC * 8 + m] * Y[(size_t)k * YC * 8 + n];
            R[(size_t)m * N + n] = a;
        }
    if (M == 64) {   // find the lane of every row
        int found = 0;
        for (int m = 0; m < 64; ++m) {
            int best = -1; double be = 1e30;
            for (int l = 0; l < 128; ++l) {
                double err = 0;
                for (int n = 0; n < N; ++n) err = fmax(err, fabs(D[(size_t)l * N + n] - R[(size_t)m * N + n]));
                if (err < be) { be = err; best = l; }
            }
            if (be < 1e-2) ++found;
            if (m % 8 == 0 || be >= 1e-2) printf("row %2d -> lane %3d (err %.3g)\n", m, best, be);
        }
        printf("%s: %d of 64 rows located\n", t, found);
        return found == 64 ? 0 : 1;
    }
    double worst = 0, ref = 0;
    for (size_t i = 0; i < R.size(); ++i) { worst = fmax(worst, fabs(D[i] - R[i])); ref = fmax(ref, fabs(R[i])); }
    printf("%s: max |err| %.4g (max |ref| %.4g) -> %s\n", t, worst, ref, worst < 1e-3 * ref + 1e-3 ? "OK" : "MISMATCH");
    return worst < 1e-3 * ref + 1e-3 ? 0 : 1;
}
