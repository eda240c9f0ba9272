// Small HBM-bound kernels of the hot path: stem max-pool, LayerNorm, lin_sine query encoding.
#include "common.cuh"

namespace cotr {

namespace {

// torchvision resnet stem: MaxPool2d(kernel 3, stride 2, padding 1) on NHWC, float4 over channels.
__global__ void maxpool_3x3s2_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                          int N, int H, int W, int C4) {
    const int OH = H / 2, OW = W / 2;
    const size_t total = (size_t)N * OH * OW * C4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c4 = idx % C4;
        size_t t = idx / C4;
        const int ow = t % OW; t /= OW;
        const int oh = t % OH;
        const int n = t / OH;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int ih = oh * 2 - 1 + dh;
            if (ih < 0 || ih >= H) continue;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int iw = ow * 2 - 1 + dw;
                if (iw < 0 || iw >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(in) + (((size_t)n * H + ih) * W + iw) * C4 + c4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        reinterpret_cast<float4*>(out)[idx] = m;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out = LayerNorm(x (+ residual)) over 256 channels, eps 1e-5, biased variance.  One warp per row.
__global__ void __launch_bounds__(256) layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ residual,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ out, int rows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * kDModel);
    float4 a = xr[lane], b = xr[32 + lane];
    if (residual) {
        const float4* rr = reinterpret_cast<const float4*>(residual + (size_t)warp * kDModel);
        const float4 ra = rr[lane], rb = rr[32 + lane];
        a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
        b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
    }
    const float mean = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.f / kDModel);
    a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
    b.x -= mean; b.y -= mean; b.z -= mean; b.w -= mean;
    const float var = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w +
                               b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w) * (1.f / kDModel);
    const float rstd = 1.f / sqrtf(var + 1e-5f);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + lane), gb = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
    const float4 ba = __ldg(reinterpret_cast<const float4*>(beta) + lane), bb = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
    float4 oa, ob;
    oa.x = a.x * rstd * ga.x + ba.x; oa.y = a.y * rstd * ga.y + ba.y; oa.z = a.z * rstd * ga.z + ba.z; oa.w = a.w * rstd * ga.w + ba.w;
    ob.x = b.x * rstd * gb.x + bb.x; ob.y = b.y * rstd * gb.y + bb.y; ob.z = b.z * rstd * gb.z + bb.z; ob.w = b.w * rstd * gb.w + bb.w;
    float4* orow = reinterpret_cast<float4*>(out + (size_t)warp * kDModel);
    orow[lane] = oa;
    orow[32 + lane] = ob;
}

// position_encoding.py:41-45 with bases 1..64 on (x, y):
// channel 2(k-1)+a = sin(fp32(k*pi) * p_a), channel 128 + 2(k-1)+a = cos(...).  Accurate sincosf: |angle| <= 64*pi.
__global__ void __launch_bounds__(128) query_encode_kernel(const float* __restrict__ queries, float* __restrict__ qpos, int rows) {
    const int row = blockIdx.x;
    if (row >= rows) return;
    const int t = threadIdx.x;          // 0..127 = 2*(k-1) + axis
    const int k = (t >> 1) + 1;
    const float p = __ldg(queries + (size_t)row * 2 + (t & 1));
    const float kpi = (float)((double)k * 3.14159265358979323846);
    const float angle = __fmul_rn(kpi, p);
    float s, c;
    sincosf(angle, &s, &c);
    qpos[(size_t)row * kDModel + t] = s;
    qpos[(size_t)row * kDModel + 128 + t] = c;
}

}  // namespace

int launch_maxpool_3x3s2_nhwc(const float* in, float* out, int N, int H, int W, int C, cudaStream_t s) {
    COTR_CHECK((C & 3) == 0 && (H & 1) == 0 && (W & 1) == 0, "maxpool: unsupported shape");
    const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 4);
    const int block = 256;
    const int grid = (int)((total + block - 1) / block < 148 * 16 ? (total + block - 1) / block : 148 * 16);
    maxpool_3x3s2_nhwc_kernel<<<grid, block, 0, s>>>(in, out, N, H, W, C / 4);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, float* out,
                     int rows, cudaStream_t s) {
    if (rows <= 0) return 0;
    layernorm256_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, residual, gamma, beta, out, rows);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_query_encode(const float* queries, float* qpos, int rows, cudaStream_t s) {
    if (rows <= 0) return 0;
    query_encode_kernel<<<rows, 128, 0, s>>>(queries, qpos, rows);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cotr
