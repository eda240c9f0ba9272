// fp32 SIMT GEMM with implicit-im2col A loader and fused epilogue.
// Numerical cross-check path (cotr_set_gemm_path(m, 1)) and the producer of the constant
// position-bias matrices at model creation.  The product path is gemm_tc.cu (tcgen05).
#include "a_loader.cuh"

namespace cotr {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const int lrow = tid >> 2;          // 0..63
    const int lk = (tid & 3) * 4;       // 0,4,8,12
    const ARow arow = decode_a_row(p, m0 + lrow);
    const int wn = n0 + lrow;
    const float* wrow = p.Wt + (size_t)wn * p.K;
    const bool w_vec = (p.K & 3) == 0;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
        const int k = k0 + lk;
        const float4 a = load_a4(p, arow, k);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wn < p.N && k < p.K) {
            if (w_vec) {
                b = __ldg(reinterpret_cast<const float4*>(wrow + k));
            } else {
                b.x = __ldg(wrow + k);
                if (k + 1 < p.K) b.y = __ldg(wrow + k + 1);
                if (k + 2 < p.K) b.z = __ldg(wrow + k + 2);
                if (k + 3 < p.K) b.w = __ldg(wrow + k + 3);
            }
        }
        __syncthreads();
        As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
        Bs[lk + 0][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        const float* add_row = p.addmat ? p.addmat + (size_t)(m % p.add_period) * p.ld_add : nullptr;
        const float* res_row = p.residual ? p.residual + (size_t)m * p.ldr : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + n);
            if (add_row) v += __ldg(add_row + n);
            if (res_row) v += __ldg(res_row + n);
            if (p.relu) v = fmaxf(v, 0.f);
            p.out[(size_t)m * p.ldc + n] = v;
        }
    }
}

}  // namespace

int launch_gemm_simt(const GemmParams& p, cudaStream_t s) {
    COTR_CHECK(p.M > 0 && p.N > 0 && p.K > 0, "gemm_simt: empty problem %d x %d x %d", p.M, p.N, p.K);
    COTR_CHECK(p.a_mode != A_CONV_NHWC || (p.C & 3) == 0, "gemm_simt: NHWC conv needs C %% 4 == 0 (C=%d)", p.C);
    COTR_CHECK(p.ln_gamma == nullptr, "gemm_simt: fused LayerNorm is a tensor-core-path feature");
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
    gemm_simt_kernel<<<grid, 256, 0, s>>>(p);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cotr
