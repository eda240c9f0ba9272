// Shared declarations of the cotr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace cotr {

void set_error(const char* fmt, ...);

#define COTR_CHECK_CUDA(expr)                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ::cotr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define COTR_CHECK(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::cotr::set_error(__VA_ARGS__);                                                        \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

constexpr int kDModel = 256;
constexpr int kHeads = 8;
constexpr int kHeadDim = 32;
constexpr int kTokens = 512;   // 16 x 32 context grid
constexpr int kFF = 1024;
constexpr int kEncLayers = 6;
constexpr int kDecLayers = 6;

// How a GEMM finds row m, column k of its A operand.
enum AMode : int {
    A_ROWMAJOR = 0,   // A[m * lda + k]
    A_CONV_NHWC = 1,  // implicit im2col over an NHWC activation: m -> (n, oh, ow), k -> (kh, kw, c)
    A_STEM_NCHW = 2,  // implicit im2col over the (B,3,256,512) NCHW canvas, halves as separate images
    A_TOKENS = 3,     // m = pair*512 + i*32 + j gathers row ((2*pair + (j>>4))*16 + i)*16 + (j&15)
};

// D[M,N] = epilogue( A[M,K] * W[N,K]^T ).  Everything fp32 in global memory.
struct GemmParams {
    int M, N, K;
    const float* A;
    int a_mode;
    int lda;
    // conv geometry (A_CONV_NHWC / A_STEM_NCHW)
    int H, W, C;      // input height / width / channels (per image)
    int OH, OW;       // output height / width
    int KH, KW, stride, pad;
    // weights, row-major [N, K] (K ordered (kh, kw, c) for convolutions)
    const float* Wt;
    // tensor-core path: the same weights pre-scaled by a power of two, pre-split into fp16 hi/lo and pre-tiled
    // (see gemm_tc.cu); may be null.  acc_scale undoes the power of two on the accumulator.
    const void* Wtc;
    float acc_scale;
    // epilogue: v = acc + bias[n] + addmat[(m % add_period) * ld_add + n] + residual[m * ldr + n]; relu; LN
    const float* bias;
    const float* addmat;
    int add_period, ld_add;
    const float* residual;
    int ldr;
    int relu;
    // optional fused LayerNorm over the N = 256 columns of each row (after residual), tensor-core path only
    const float* ln_gamma;
    const float* ln_beta;
    float* out;
    int ldc;
};

// softmax(q k^T) v per head; q already carries the head_dim^-0.5 scale.
struct AttnParams {
    const float* q; int ldq;     // rows: local row r = pair_local * nq + i
    const float* k; int ldk;     // rows: (pair0 + pair_local) * 512 + key
    const float* v; int ldv;
    float* out; int ldo;
    int nq;                      // query rows per pair in this launch
    int npairs;
    int pair0;
};

int launch_gemm_simt(const GemmParams& p, cudaStream_t s);
int launch_gemm_tc(const GemmParams& p, cudaStream_t s);
int launch_attention_simt(const AttnParams& p, cudaStream_t s);
int launch_attention_tc(const AttnParams& p, cudaStream_t s);
int launch_maxpool_3x3s2_nhwc(const float* in, float* out, int N, int H, int W, int C, cudaStream_t s);
int launch_layernorm(const float* x, const float* residual, const float* gamma, const float* beta, float* out,
                     int rows, cudaStream_t s);
int launch_query_encode(const float* queries, float* qpos, int rows, cudaStream_t s);

// Bytes of the pre-tiled fp16 hi/lo image of an [N,K] weight matrix, and the host-side packer (returns acc_scale).
size_t tc_weight_bytes(int N, int K);
float tc_pack_weight(const float* w, int N, int K, void* dst_host);

}  // namespace cotr
