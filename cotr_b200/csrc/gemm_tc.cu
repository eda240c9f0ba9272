// tcgen05 GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * W[N,K]^T ),  fp32 in / fp32 out.
//
// Precision: the 1e-3 parity bar on predicted (x,y) rules out single-pass bf16 / tf32 / fp16 operands
// (SURVEY.md appendix E.3).  Each fp32 operand is split into two fp16 terms (x ~= hi + lo, ~22 mantissa bits)
// and the product is formed as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM - three kind::f16 MMAs
// per K step, ~2^-21 relative error per product.  Weights are pre-multiplied by a per-tensor power of two so
// that their lo terms stay in fp16's normal range; the epilogue multiplies the accumulator by the inverse (exact).
//
// Data movement per CTA (one 128 x BN output tile, K walked in chunks of 64):
//   * weights: pre-split, pre-tiled in HBM at model creation (tc_pack_weight) so that each pipeline stage is a
//     handful of contiguous bulk-TMA copies (cp.async.bulk -> UBLKCP) straight into the UMMA canonical layout;
//   * activations: warps 0-3 load fp32 (implicit im2col for the convolutions, a_loader.cuh), split to fp16
//     hi/lo in registers and store 16-byte core-matrix rows to shared memory (conflict-free thanks to a padded LBO);
//   * warp 4 (one lane) issues the TMA copies, warp 5 (one lane) issues tcgen05.mma and owns the TMEM allocation;
//   * warps 0-3 then run the epilogue out of TMEM: bias / constant add-matrix / residual / ReLU, or the fused
//     residual + LayerNorm over the full 256-wide row (each thread owns one row, so no cross-thread reduction).
#include <cmath>
#include <cstring>
#include <vector>

#include "a_loader.cuh"
#include "tc_common.cuh"

namespace cotr {

int g_tc_variant = 0;   // bring-up switch: bit0 swaps the LBO / SBO fields of the shared-memory descriptors

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr uint32_t kALbo = BM * 16 + 16;       // padded: 8 lanes writing the 8 K-groups of one row hit 32 distinct banks
constexpr uint32_t kAPlane = 8 * kALbo;        // one fp16 plane (hi or lo) of the 128 x 64 A tile
constexpr uint32_t kSbo = 128;                 // 8 rows x 16 bytes

template <int BN>
struct Cfg {
    static constexpr uint32_t kBPlane = BN * 128;                       // 8 K-groups x BN rows x 16 bytes
    static constexpr uint32_t kStage = 2 * kAPlane + 2 * kBPlane;
    static constexpr int kStagesRaw = (int)((227u * 1024u - 2048u) / kStage);
    static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
    // The tensor core adds into its fp32 accumulator with truncation (measured: relative error grows ~1e-7 per
    // chained MMA, profiles/r01_tc_precision.md).  K steps are therefore dealt round-robin onto kMainAcc TMEM
    // accumulators and the small hi*lo / lo*hi products onto a separate one; the epilogue adds them up in fp32 RN.
    static constexpr int kMainAcc = BN >= 256 ? 1 : (BN >= 128 ? 3 : 4);
    static constexpr uint32_t kAccCols = (kMainAcc + 1) * BN;
    static constexpr uint32_t kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
    static constexpr uint32_t kSmemBytes = kStages * kStage + 1024;
    static_assert(kStages >= 2, "pipeline needs at least two stages");
};

__host__ __device__ inline int tc_npad(int N) { return N >= 64 ? ((N + 63) / 64) * 64 : ((N + 15) / 16) * 16; }

template <int BN, bool LN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const GemmParams p, const int npad, const int variant) {
    using C = Cfg<BN>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStage);
    uint64_t* full_a = bars;
    uint64_t* full_b = bars + C::kStages;
    uint64_t* empty = bars + 2 * C::kStages;
    uint64_t* accum_full = bars + 3 * C::kStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int KC = (p.K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_a[s], 128);
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_full, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc(tmem_ptr, C::kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        // ================= A producer: fp32 global -> fp16 hi/lo core-matrix rows in shared memory ============
        const int t = threadIdx.x;
        const int kg = t & 7;          // 16-byte K group (8 halves) inside the 64-wide chunk
        const int rb = t >> 3;         // rows rb, rb+16, ..., rb+112
        ARow rows[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rows[i] = decode_a_row(p, m0 + rb + 16 * i);
        const bool gather = (p.a_mode == A_ROWMAJOR || p.a_mode == A_TOKENS);
        const bool fast = gather ? ((p.K & 7) == 0 && (p.lda & 3) == 0) : (p.a_mode == A_CONV_NHWC && (p.C & 63) == 0);

        // global loads of chunk `it` into registers (issued one chunk ahead of their use to keep loads in flight)
        auto fetch = [&](int it, float4 (&buf)[16]) {
            const int k0 = it * BK;
            int kh = 0, kw = 0, c0 = 0;
            if (!gather && fast) {     // a 64-wide K chunk lies inside one filter tap because C % 64 == 0
                const int tap = k0 / p.C;
                c0 = k0 - tap * p.C + kg * 8;
                kh = tap / p.KW;
                kw = tap - kh * p.KW;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                const ARow& r = rows[i];
                if (fast) {
                    if (gather) {
                        const int k = k0 + kg * 8;
                        if (r.valid && k < p.K) {
                            v0 = __ldg(reinterpret_cast<const float4*>(r.base + k));
                            v1 = __ldg(reinterpret_cast<const float4*>(r.base + k + 4));
                        }
                    } else {
                        const int ih = r.ih0 + kh, iw = r.iw0 + kw;
                        if (r.valid && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            const float* src = r.base + ((size_t)ih * p.W + iw) * p.C + c0;
                            v0 = __ldg(reinterpret_cast<const float4*>(src));
                            v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                        }
                    }
                } else {
                    v0 = load_a4(p, r, k0 + kg * 8);
                    v1 = load_a4(p, r, k0 + kg * 8 + 4);
                }
                buf[2 * i] = v0;
                buf[2 * i + 1] = v1;
            }
        };

        float4 cur[16], nxt[16];
        fetch(0, cur);
        for (int it = 0; it < KC; ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
            if (it + 1 < KC) fetch(it + 1, nxt);
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t* a_hi = stage_base + (size_t)s * C::kStage;
            uint8_t* a_lo = a_hi + kAPlane;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v0 = cur[2 * i], v1 = cur[2 * i + 1];
                uint4 hi, lo;
                split_f16x2(v0.x, v0.y, hi.x, lo.x);
                split_f16x2(v0.z, v0.w, hi.y, lo.y);
                split_f16x2(v1.x, v1.y, hi.z, lo.z);
                split_f16x2(v1.z, v1.w, hi.w, lo.w);
                const uint32_t off = (uint32_t)kg * kALbo + (uint32_t)(rb + 16 * i) * 16u;
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&full_a[s]);
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }

        // ================= epilogue: TMEM -> registers -> global =============================================
        mbar_wait(accum_full, 0);
        tcgen05_fence_after();
        const int n_ksteps = KC * (BK / 16);
        const int row = m0 + warp * 32 + lane;
        const bool row_ok = row < p.M;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float* add_row = (row_ok && p.addmat) ? p.addmat + (size_t)(row % p.add_period) * p.ld_add : nullptr;
        const float* res_row = (row_ok && p.residual) ? p.residual + (size_t)row * p.ldr : nullptr;
        float* out_row = p.out + (size_t)(row_ok ? row : 0) * p.ldc;
        if constexpr (!LN) {
            const bool vec_ok = (p.ldc & 3) == 0;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(trow + c, v);
#pragma unroll
                for (int a = 1; a <= C::kMainAcc; ++a) {
                    if (a < C::kMainAcc && a >= n_ksteps) continue;     // accumulator never written (tiny K)
                    float w2[16];
                    tmem_ld16(trow + a * BN + c, w2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += w2[j];
                }
                if (!row_ok) continue;
                const int nb = n0 + c;
                if (nb >= p.N) continue;
                if (nb + 15 < p.N) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x = v[j] * p.acc_scale;
                        if (p.bias) x += __ldg(p.bias + nb + j);
                        if (add_row) x += __ldg(add_row + nb + j);
                        if (res_row) x += __ldg(res_row + nb + j);
                        if (p.relu) x = fmaxf(x, 0.f);
                        v[j] = x;
                    }
                    if (vec_ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(out_row + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) out_row[nb + j] = v[j];
                    }
                } else {
                    for (int j = 0; j < 16 && nb + j < p.N; ++j) {
                        float x = v[j] * p.acc_scale;
                        if (p.bias) x += __ldg(p.bias + nb + j);
                        if (add_row) x += __ldg(add_row + nb + j);
                        if (res_row) x += __ldg(res_row + nb + j);
                        if (p.relu) x = fmaxf(x, 0.f);
                        out_row[nb + j] = x;
                    }
                }
            }
        } else {
            // fused residual + LayerNorm (eps 1e-5, biased variance) over the 256 columns this thread owns
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(trow + c, v);
#pragma unroll
                for (int a = 1; a <= C::kMainAcc; ++a) {
                    if (a < C::kMainAcc && a >= n_ksteps) continue;     // accumulator never written (tiny K)
                    float w2[16];
                    tmem_ld16(trow + a * BN + c, w2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += w2[j];
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = v[j] * p.acc_scale;
                    if (p.bias) x += __ldg(p.bias + c + j);
                    if (add_row) x += __ldg(add_row + c + j);
                    if (res_row) x += __ldg(res_row + c + j);
                    v[j] = x;
                    sum += x;
                }
                tmem_st16(trow + c, v);
            }
            tmem_st_wait();
            const float mean = sum * (1.f / 256.f);
            float sq = 0.f;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(trow + c, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float d = v[j] - mean;
                    sq = fmaf(d, d, sq);
                }
            }
            const float rstd = 1.f / sqrtf(sq * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(trow + c, v);
                if (!row_ok) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = (v[j] - mean) * rstd * __ldg(p.ln_gamma + c + j) + __ldg(p.ln_beta + c + j);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(out_row + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    } else if (warp == 4) {
        // ================= weight producer: bulk TMA of the pre-tiled fp16 hi/lo image ==========================
        if (lane == 0) {
            const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.Wtc);
            for (int it = 0; it < KC; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], 2u * C::kBPlane);
                uint8_t* b_dst = stage_base + (size_t)s * C::kStage + 2 * kAPlane;
                // image: [k chunk][plane][K group][npad rows][16 bytes]
                const uint8_t* src = wimg + ((size_t)it * 16) * (size_t)npad * 16 + (size_t)n0 * 16;
#pragma unroll 1
                for (int j = 0; j < 16; ++j)
                    tma_bulk_g2s(b_dst + (size_t)j * BN * 16, src + (size_t)j * npad * 16, BN * 16, &full_b[s]);
            }
        }
        __syncwarp();
    } else {
        // ================= MMA issuer ===========================================================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, BN);
            const uint32_t b_lbo = BN * 16;
            for (int it = 0; it < KC; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
                mbar_wait(&full_a[s], ph);
                mbar_wait(&full_b[s], ph);
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(stage_base + (size_t)s * C::kStage);
                const uint32_t a_lo = a_hi + kAPlane;
                const uint32_t b_hi = a_hi + 2 * kAPlane;
                const uint32_t b_lo = b_hi + C::kBPlane;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const uint32_t ao = ks * 2 * kALbo, bo = ks * 2 * b_lbo;
                    uint64_t dah, dal, dbh, dbl;
                    if (variant & 1) {
                        dah = make_smem_desc(a_hi + ao, kSbo, kALbo); dal = make_smem_desc(a_lo + ao, kSbo, kALbo);
                        dbh = make_smem_desc(b_hi + bo, kSbo, b_lbo); dbl = make_smem_desc(b_lo + bo, kSbo, b_lbo);
                    } else {
                        dah = make_smem_desc(a_hi + ao, kALbo, kSbo); dal = make_smem_desc(a_lo + ao, kALbo, kSbo);
                        dbh = make_smem_desc(b_hi + bo, b_lbo, kSbo); dbl = make_smem_desc(b_lo + bo, b_lbo, kSbo);
                    }
                    const int g = it * (BK / 16) + ks;                       // global K step
                    const uint32_t main_acc = tmem_base + (uint32_t)(g % C::kMainAcc) * BN;
                    const uint32_t corr_acc = tmem_base + (uint32_t)C::kMainAcc * BN;
                    umma_f16_ss(corr_acc, dal, dbh, idesc, g != 0);
                    umma_f16_ss(corr_acc, dah, dbl, idesc, true);
                    umma_f16_ss(main_acc, dah, dbh, idesc, g >= C::kMainAcc);
                }
                umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
            }
            umma_commit(accum_full);
        }
        __syncwarp();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN, bool LN>
int launch_one(const GemmParams& p, cudaStream_t s) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        COTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
        configured = true;
    }
    const int npad = tc_npad(p.N);
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
    gemm_tc_kernel<BN, LN><<<grid, kThreads, C::kSmemBytes, s>>>(p, npad, g_tc_variant);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

inline uint16_t f32_to_f16_rn(float f) {     // round-to-nearest-even, saturating, subnormals supported
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7FFFFFFFu;
    if (x >= 0x47800000u) return sign | 0x7BFFu;                     // >= 65536 (or NaN): clamp to max finite
    if (x < 0x38800000u) {                                           // < 2^-14: fp16 subnormal, spacing 2^-24
        float af;
        memcpy(&af, &x, 4);
        const uint32_t m = (uint32_t)nearbyintf(af * 16777216.0f);   // <= 0x400 (== smallest normal when it rounds up)
        return sign | (uint16_t)m;
    }
    const uint32_t mant = x & 0x7FFFFFu;
    uint32_t h = (((x >> 23) - 112u) << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    if (h >= 0x7C00u) h = 0x7BFFu;
    return sign | (uint16_t)h;
}
inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    float mag;
    if (e == 0) mag = (float)m * (1.0f / 16777216.0f);
    else {
        const uint32_t u = ((e + 112u) << 23) | (m << 13);
        memcpy(&mag, &u, 4);
    }
    uint32_t u;
    memcpy(&u, &mag, 4);
    u |= sign;
    float out;
    memcpy(&out, &u, 4);
    return out;
}

}  // namespace

size_t tc_weight_bytes(int N, int K) {
    const size_t kc = (K + BK - 1) / BK;
    return kc * 16 * (size_t)tc_npad(N) * 16;
}

// Image layout: [k chunk (64)][plane: hi, lo][K group (8 halves)][npad rows][8 halves]; zero padded in N and K.
// The matrix is multiplied by 2^e, e chosen so that max|w| * 2^e lies in [2^12, 2^13); returns 2^-e for the epilogue.
float tc_pack_weight(const float* w, int N, int K, void* dst_host) {
    const int npad = tc_npad(N);
    const int kc_n = (K + BK - 1) / BK;
    float amax = 0.f;
    for (size_t i = 0; i < (size_t)N * K; ++i) amax = fmaxf(amax, fabsf(w[i]));
    int e = 0;
    if (amax > 0.f && std::isfinite(amax)) {
        e = 12 - (int)floorf(log2f(amax));
        if (e > 24) e = 24;
        if (e < -24) e = -24;
    }
    const float scale = ldexpf(1.f, e);
    uint16_t* out = reinterpret_cast<uint16_t*>(dst_host);
    for (int kc = 0; kc < kc_n; ++kc)
        for (int kg = 0; kg < 8; ++kg)
            for (int n = 0; n < npad; ++n)
                for (int el = 0; el < 8; ++el) {
                    const int k = kc * BK + kg * 8 + el;
                    const float x = (n < N && k < K) ? w[(size_t)n * K + k] * scale : 0.f;
                    const uint16_t hi = f32_to_f16_rn(x);
                    const uint16_t lo = f32_to_f16_rn(x - f16_to_f32(hi));
                    out[((((size_t)kc * 2 + 0) * 8 + kg) * npad + n) * 8 + el] = hi;
                    out[((((size_t)kc * 2 + 1) * 8 + kg) * npad + n) * 8 + el] = lo;
                }
    return ldexpf(1.f, -e);
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t s) {
    COTR_CHECK(p.M > 0 && p.N > 0 && p.K > 0, "gemm_tc: empty problem %d x %d x %d", p.M, p.N, p.K);
    COTR_CHECK(p.Wtc != nullptr, "gemm_tc: weight has no tensor-core image");
    COTR_CHECK(p.a_mode != A_CONV_NHWC || (p.C & 3) == 0, "gemm_tc: NHWC conv needs C %% 4 == 0 (C=%d)", p.C);
    if (p.ln_gamma) {
        COTR_CHECK(p.N == 256 && p.ldc == 256 && p.relu == 0, "gemm_tc: LayerNorm epilogue needs N = ldc = 256");
        return launch_one<256, true>(p, s);
    }
    if (p.N <= 16) return launch_one<16, false>(p, s);
    const int mt = (p.M + BM - 1) / BM;
    if ((p.N % 128) == 0 && (long long)mt * (p.N / 128) >= 120) return launch_one<128, false>(p, s);
    return launch_one<64, false>(p, s);
}

}  // namespace cotr
