// Device-side post-processing of the dense pass (COTR/inference/inference_helper.py:131-145): the network answers the
// 131 072 grid queries (j/512, i/256) of a 256 x 512 canvas; this kernel turns the (256,512,2) predictions into the
// (256,512,3) map the reference builds on the host -
//   * in [-1,1] coordinates (v * 2 - 1),
//   * cycle consistency: the prediction field sampled AT the predicted location with torch's grid_sample semantics
//     (bilinear, zero padding, align_corners = False: inference_helper.py:139), confidence = | cycle - query |,
//   * x re-expressed in the OTHER image: left half x * 2 - 1, right half x * 2 + 1 (:141-143).
// HBM-bound elementwise work: 1 MB read (plus 4 gathered taps per pixel, L2 hits), 1.5 MB written per canvas.
#include "common.cuh"

namespace cotr {

namespace {

constexpr int kH = 256, kW = 512;

__device__ __forceinline__ float2 field(const float* __restrict__ pred, int iy, int ix) {
    // the sampled field is out_grid = pred * 2 - 1; taps outside the canvas contribute zero
    if (ix < 0 || ix >= kW || iy < 0 || iy >= kH) return make_float2(0.f, 0.f);
    const float2 v = *reinterpret_cast<const float2*>(pred + ((size_t)iy * kW + ix) * 2);
    return make_float2(__fsub_rn(__fmul_rn(v.x, 2.f), 1.f), __fsub_rn(__fmul_rn(v.y, 2.f), 1.f));
}

__global__ void __launch_bounds__(256) dense_post_kernel(const float* __restrict__ pred_all, float* __restrict__ out_all) {
    pdl_wait();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;            // i * 512 + j
    const float* pred = pred_all + (size_t)blockIdx.y * kH * kW * 2;
    float* out = out_all + (size_t)blockIdx.y * kH * kW * 3;
    const int i = pix >> 9, j = pix & (kW - 1);
    const float2 p = *reinterpret_cast<const float2*>(pred + (size_t)pix * 2);
    const float ox = __fsub_rn(__fmul_rn(p.x, 2.f), 1.f), oy = __fsub_rn(__fmul_rn(p.y, 2.f), 1.f);
    // grid_sample(out_grid, out_grid): unnormalise with align_corners = False, 4 bilinear taps, zeros outside
    const float fx = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(ox, 1.f), (float)kW), 1.f), 2.f);
    const float fy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(oy, 1.f), (float)kH), 1.f), 2.f);
    float cx = 0.f, cy = 0.f;
    if (isfinite(fx) && isfinite(fy) && fabsf(fx) < 1e7f && fabsf(fy) < 1e7f) {
        const float x0 = floorf(fx), y0 = floorf(fy);
        const float x1 = x0 + 1.f, y1 = y0 + 1.f;
        const float w_nw = __fmul_rn(__fsub_rn(x1, fx), __fsub_rn(y1, fy));
        const float w_ne = __fmul_rn(__fsub_rn(fx, x0), __fsub_rn(y1, fy));
        const float w_sw = __fmul_rn(__fsub_rn(x1, fx), __fsub_rn(fy, y0));
        const float w_se = __fmul_rn(__fsub_rn(fx, x0), __fsub_rn(fy, y0));
        const int ix0 = (int)x0, iy0 = (int)y0;
        const float2 nw = field(pred, iy0, ix0), ne = field(pred, iy0, ix0 + 1);
        const float2 sw = field(pred, iy0 + 1, ix0), se = field(pred, iy0 + 1, ix0 + 1);
        cx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw.x, w_nw), __fmul_rn(ne.x, w_ne)), __fmul_rn(sw.x, w_sw)), __fmul_rn(se.x, w_se));
        cy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw.y, w_nw), __fmul_rn(ne.y, w_ne)), __fmul_rn(sw.y, w_sw)), __fmul_rn(se.y, w_se));
    }
    // the query this pixel asked: (j / 512, i / 256) -> [-1,1]
    const float qx = __fsub_rn((float)j * (1.f / 256.f), 1.f), qy = __fsub_rn((float)i * (1.f / 128.f), 1.f);
    const float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy);
    const float conf = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float x_other = j < kW / 2 ? __fsub_rn(__fmul_rn(ox, 2.f), 1.f) : __fadd_rn(__fmul_rn(ox, 2.f), 1.f);
    float* o = out + (size_t)pix * 3;
    o[0] = x_other;
    o[1] = oy;
    o[2] = conf;
}

}  // namespace

// pred: (n, 256*512, 2) fp32 predictions of the grid queries; out: (n, 256, 512, 3) fp32 [x in the other image, y, confidence]
int dense_post_launch(const float* pred, float* out, int n, cudaStream_t s) {
    COTR_CHECK(pred && out && n >= 1 && n <= 65535, "cotr_dense_postprocess: bad arguments");
    COTR_CHECK_CUDA(launch_kernel(dense_post_kernel, dim3(kH * kW / 256, (unsigned)n), dim3(256), 0, s, pred, out));
    return 0;
}

}  // namespace cotr
