// tcgen05 attention over the 512-token context (encoder self-attention and decoder cross-attention).
//
// One CTA per (128-query tile, head, image pair).  The whole score tile S = Q K^T (128 x 512 fp32) fits TMEM exactly
// (512 columns), so the softmax is exact (no online rescaling):
//   warps 0-3  stage Q, K and V^T of this head as fp16 hi/lo core-matrix tiles (fp32-faithful 3-product scheme,
//              see gemm_tc.cu), then run the softmax straight out of TMEM - each thread owns one query row, so the
//              row max / row sum need no cross-thread reduction - and hand P to the MMA in 64-key chunks (double buffered);
//   warp 4     (one lane) issues the tcgen05 MMAs: 2 x (128x256x32) for S, then 8 x (128x32x64) for O = P V.  The O
//              accumulators re-use TMEM columns of S chunks that have already been turned into P: because the tensor
//              core's fp32 accumulate truncates (profiles/r01_tc_precision.md) the hi*hi products alternate between
//              two accumulators ([0,32) and [64,96)) and the small correction products go to a third ([32,64)).
// q is expected pre-scaled by head_dim^-0.5 (folded into the projection weights).
#include "common.cuh"
#include "tc_common.cuh"

namespace cotr {

extern int g_tc_variant;

namespace {

using namespace tc;

constexpr int kTile = 128;
constexpr int kThreads = 160;
constexpr int kChunk = 64;                                // keys per P chunk
constexpr int kChunks = kTokens / kChunk;                 // 8
constexpr uint32_t kQLbo = kTile * 16;                    // Q tile  [4 K-groups][128 rows][16 B]
constexpr uint32_t kQPlane = 4 * kQLbo;                   // 8 KB
constexpr uint32_t kKLbo = kTokens * 16;                  // K tile  [4 K-groups][512 keys][16 B]
constexpr uint32_t kKPlane = 4 * kKLbo;                   // 32 KB
constexpr uint32_t kVLbo = kHeadDim * 16;                 // V^T tile [64 key-groups][32 d][16 B]
constexpr uint32_t kVPlane = (kTokens / 8) * kVLbo;       // 32 KB
constexpr uint32_t kPLbo = kTile * 16;                    // P chunk [8 key-groups][128 rows][16 B]
constexpr uint32_t kPPlane = (kChunk / 8) * kPLbo;        // 16 KB
constexpr uint32_t kSbo = 128;

constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + 2 * kQPlane;
constexpr uint32_t kOffV = kOffK + 2 * kKPlane;
constexpr uint32_t kOffP = kOffV + 2 * kVPlane;           // 2 buffers x (hi, lo)
constexpr uint32_t kOffBar = kOffP + 4 * kPPlane;
constexpr uint32_t kSmemBytes = kOffBar + 128;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int variant) {
    return (variant & 1) ? make_smem_desc(addr, sbo, lbo) : make_smem_desc(addr, lbo, sbo);
}

__global__ void __launch_bounds__(kThreads, 1) attention_tc_kernel(const AttnParams p, const int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* o_full = bars + 3;
    uint64_t* p_full = bars + 4;     // [2]
    uint64_t* p_empty = bars + 6;    // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int pair_local = blockIdx.z;
    const int row0 = blockIdx.x * kTile;

    if (threadIdx.x == 0) {
        mbar_init(qk_full, 128);
        mbar_init(v_full, 128);
        mbar_init(s_full, 1);
        mbar_init(o_full, 1);
        mbar_init(&p_full[0], 128);
        mbar_init(&p_full[1], 128);
        mbar_init(&p_empty[0], 1);
        mbar_init(&p_empty[1], 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        const int t = threadIdx.x;                       // == query row inside the tile == TMEM lane
        const int qi = row0 + t;
        const bool row_ok = qi < p.nq;
        const size_t grow = (size_t)pair_local * p.nq + (row_ok ? qi : 0);
        const size_t kv_row0 = (size_t)(p.pair0 + pair_local) * kTokens;

        // ---- stage Q (one row per thread) and K (4 keys per thread) ------------------------------------------
        {
            const float* src = p.q + grow * p.ldq + head * kHeadDim;
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (row_ok) {
                    v0 = __ldg(reinterpret_cast<const float4*>(src + kg * 8));
                    v1 = __ldg(reinterpret_cast<const float4*>(src + kg * 8 + 4));
                }
                uint4 hi, lo;
                split_f16x2(v0.x, v0.y, hi.x, lo.x);
                split_f16x2(v0.z, v0.w, hi.y, lo.y);
                split_f16x2(v1.x, v1.y, hi.z, lo.z);
                split_f16x2(v1.z, v1.w, hi.w, lo.w);
                const uint32_t off = kg * kQLbo + t * 16;
                *reinterpret_cast<uint4*>(smem + kOffQ + off) = hi;
                *reinterpret_cast<uint4*>(smem + kOffQ + kQPlane + off) = lo;
            }
        }
#pragma unroll 1
        for (int i = 0; i < kTokens / 128; ++i) {
            const int key = t + 128 * i;
            const float* src = p.k + (kv_row0 + key) * p.ldk + head * kHeadDim;
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + kg * 8));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + kg * 8 + 4));
                uint4 hi, lo;
                split_f16x2(v0.x, v0.y, hi.x, lo.x);
                split_f16x2(v0.z, v0.w, hi.y, lo.y);
                split_f16x2(v1.x, v1.y, hi.z, lo.z);
                split_f16x2(v1.z, v1.w, hi.w, lo.w);
                const uint32_t off = kg * kKLbo + key * 16;
                *reinterpret_cast<uint4*>(smem + kOffK + off) = hi;
                *reinterpret_cast<uint4*>(smem + kOffK + kKPlane + off) = lo;
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(qk_full);

        // ---- stage V^T: B operand of O = P V is [d][key] with keys contiguous (K-major) ------------------------
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int u = t + 128 * i;
            const int dq = u & 7, kg8 = u >> 3;
            const float* src = p.v + (kv_row0 + (size_t)kg8 * 8) * p.ldv + head * kHeadDim + dq;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = __ldg(src + (size_t)j * p.ldv + 8 * e);
                uint4 hi, lo;
                split_f16x2(x[0], x[1], hi.x, lo.x);
                split_f16x2(x[2], x[3], hi.y, lo.y);
                split_f16x2(x[4], x[5], hi.z, lo.z);
                split_f16x2(x[6], x[7], hi.w, lo.w);
                const uint32_t off = kg8 * kVLbo + (dq + 8 * e) * 16;
                *reinterpret_cast<uint4*>(smem + kOffV + off) = hi;
                *reinterpret_cast<uint4*>(smem + kOffV + kVPlane + off) = lo;
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(v_full);

        // ---- softmax out of TMEM ---------------------------------------------------------------------------
        mbar_wait(s_full, 0);
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kTokens; c += 16) {
            float v[16];
            __syncwarp();
            tmem_ld16(trow + c, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j]);
        }
        const float kLog2e = 1.4426950408889634f;
        const float mxs = mx * kLog2e;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
            const int buf = c & 1;
            if (c >= 2) mbar_wait(&p_empty[buf], (uint32_t)((c >> 1) - 1) & 1u);
            uint8_t* p_hi = smem + kOffP + buf * 2 * kPPlane;
            uint8_t* p_lo = p_hi + kPPlane;
#pragma unroll
            for (int h = 0; h < kChunk / 16; ++h) {
                float v[16];
                __syncwarp();
                tmem_ld16(trow + c * kChunk + h * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = fast_exp2(fmaf(v[j], kLog2e, -mxs));
                    sum += v[j];
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint4 hi, lo;
                    split_f16x2(v[g * 8 + 0], v[g * 8 + 1], hi.x, lo.x);
                    split_f16x2(v[g * 8 + 2], v[g * 8 + 3], hi.y, lo.y);
                    split_f16x2(v[g * 8 + 4], v[g * 8 + 5], hi.z, lo.z);
                    split_f16x2(v[g * 8 + 6], v[g * 8 + 7], hi.w, lo.w);
                    const uint32_t off = (h * 2 + g) * kPLbo + t * 16;
                    *reinterpret_cast<uint4*>(p_hi + off) = hi;
                    *reinterpret_cast<uint4*>(p_lo + off) = lo;
                }
            }
            tcgen05_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&p_full[buf]);
        }

        // ---- O / sum -> global -----------------------------------------------------------------------------
        mbar_wait(o_full, 0);
        tcgen05_fence_after();
        const float inv = 1.f / sum;
        float* dst = p.out + grow * p.ldo + head * kHeadDim;
#pragma unroll
        for (int c = 0; c < kHeadDim; c += 16) {
            float v[16], w1[16], w2[16];
            __syncwarp();
            tmem_ld16(trow + c, v);
            tmem_ld16(trow + 64 + c, w1);
            tmem_ld16(trow + 32 + c, w2);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] + w1[j]) + w2[j];
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + c + j) = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
            }
        }
    } else {
        // ================= MMA issuer =========================================================================
        if (lane == 0) {
            const uint32_t sbase = smem_u32(smem);
            constexpr uint32_t idesc_s = make_idesc_f16_f32(128, 256);
            constexpr uint32_t idesc_o = make_idesc_f16_f32(128, kHeadDim);
            mbar_wait(qk_full, 0);
            tcgen05_fence_after();
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t qa = sbase + kOffQ + ks * 2 * kQLbo;
                    const uint32_t ka = sbase + kOffK + nh * 256 * 16 + ks * 2 * kKLbo;
                    const uint64_t qh = desc(qa, kQLbo, kSbo, variant), ql = desc(qa + kQPlane, kQLbo, kSbo, variant);
                    const uint64_t kh = desc(ka, kKLbo, kSbo, variant), kl = desc(ka + kKPlane, kKLbo, kSbo, variant);
                    const uint32_t d = tmem_base + nh * 256;
                    umma_f16_ss(d, ql, kh, idesc_s, ks != 0);
                    umma_f16_ss(d, qh, kl, idesc_s, true);
                    umma_f16_ss(d, qh, kh, idesc_s, true);
                }
            }
            umma_commit(s_full);
            mbar_wait(v_full, 0);
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                const int buf = c & 1;
                mbar_wait(&p_full[buf], (uint32_t)(c >> 1) & 1u);
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < kChunk / 16; ++ks) {
                    const uint32_t pa = sbase + kOffP + buf * 2 * kPPlane + ks * 2 * kPLbo;
                    const uint32_t va = sbase + kOffV + (c * (kChunk / 8) + ks * 2) * kVLbo;
                    const uint64_t ph = desc(pa, kPLbo, kSbo, variant), pl = desc(pa + kPPlane, kPLbo, kSbo, variant);
                    const uint64_t vh = desc(va, kVLbo, kSbo, variant), vl = desc(va + kVPlane, kVLbo, kSbo, variant);
                    // chunk c may only touch TMEM columns of S chunks <= c (already consumed by the softmax warps)
                    const uint32_t o_main = tmem_base + ((c & 1) ? 64u : 0u);
                    umma_f16_ss(tmem_base + 32u, pl, vh, idesc_o, (c | ks) != 0);
                    umma_f16_ss(tmem_base + 32u, ph, vl, idesc_o, true);
                    umma_f16_ss(o_main, ph, vh, idesc_o, (c >= 2) || ks != 0);
                }
                umma_commit(&p_empty[buf]);
            }
            umma_commit(o_full);
        }
        __syncwarp();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int launch_attention_tc(const AttnParams& p, cudaStream_t s) {
    if (p.nq <= 0 || p.npairs <= 0) return 0;
    if (p.nq < 32) return launch_attention_simt(p, s);   // a 128-row MMA tile would be > 75% padding
    static bool configured = false;
    if (!configured) {
        COTR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured = true;
    }
    COTR_CHECK(p.npairs <= 65535, "attention: too many pairs in one launch (%d)", p.npairs);
    COTR_CHECK((p.ldq & 3) == 0 && (p.ldk & 3) == 0 && (p.ldo & 3) == 0, "attention_tc: leading dimensions must be multiples of 4");
    dim3 grid((p.nq + kTile - 1) / kTile, kHeads, p.npairs);
    attention_tc_kernel<<<grid, kThreads, kSmemBytes, s>>>(p, g_tc_variant);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cotr
