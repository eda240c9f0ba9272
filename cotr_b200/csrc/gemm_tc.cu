// tcgen05 GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * W[N,K]^T ),  fp32 in / fp32 out.
//
// Precision: the 1e-3 parity bar on predicted (x,y) rules out single-pass bf16 / tf32 / fp16 operands
// (SURVEY.md appendix E.3).  Each fp32 operand is split into two fp16 terms (x ~= hi + lo, ~22 mantissa bits)
// and the product is formed as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM - three kind::f16 MMAs
// per K step.  Weights are pre-multiplied by a per-tensor power of two so that their lo terms stay in fp16's
// normal range; the epilogue multiplies the accumulator by the inverse (exact).  The tensor core adds into its fp32
// accumulator with truncation (measured: ~1e-7 relative per chained MMA, profiles/r01_tc_precision.md), so the K steps
// are dealt round-robin onto several TMEM accumulators and the small hi*lo / lo*hi products onto a separate one; the
// epilogue adds them up with fp32 round-to-nearest.
//
// Data movement per CTA (one 128 x BN output tile, K walked in chunks of 64):
//   * weights: pre-split, pre-tiled in HBM at model creation (tc_pack_weight) into 64-row x 64-k blocks that ARE the
//     UMMA canonical shared-memory image, so a pipeline stage is BN/64 contiguous 16 KB bulk-TMA copies
//     (cp.async.bulk -> UBLKCP) completing on an mbarrier;
//   * activations: warps 0-3 load fp32 (implicit im2col for the convolutions), split to fp16 hi/lo in registers
//     (loads for chunk i+1 are in flight while chunk i is converted) and store 16-byte core-matrix rows to shared
//     memory (conflict-free thanks to a padded LBO);
//   * warp 4 (one lane) issues the TMA copies, warp 5 (one lane) issues tcgen05.mma (N = 64 atoms) and owns TMEM;
//   * warps 0-3 then run the epilogue out of TMEM: bias / constant add-matrix / residual / ReLU, or the fused
//     residual + LayerNorm over the full 256-wide row (each thread owns one row, so no cross-thread reduction).
// The kernel is templated on the A-operand addressing mode so that each instantiation carries exactly one loader
// (an earlier all-modes-in-one kernel was ~30k SASS instructions and instruction-cache bound, profiles/r01_*).
#include <cmath>
#include <cstring>
#include <vector>

#include "a_loader.cuh"
#include "tc_common.cuh"

namespace cotr {

int g_tc_variant = 0;   // bring-up switch: bit0 swaps the LBO / SBO fields of the shared-memory descriptors
long long* g_tc_timestamps = nullptr;   // debug: 64 clock64() stamps per CTA (cotr_debug_set_timestamps), else null

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr uint32_t kALbo = BM * 16 + 16;       // padded: 8 lanes writing the 8 K-groups of one row hit 32 distinct banks
constexpr uint32_t kAPlane = 8 * kALbo;        // one fp16 plane (hi or lo) of the 128 x 64 A tile
constexpr uint32_t kSbo = 128;                 // 8 rows x 16 bytes

enum LoaderMode : int { LD_GATHER = 0, LD_CONV = 1, LD_GENERIC = 2 };

__host__ __device__ inline int tc_npad(int N) { return N >= 64 ? ((N + 63) / 64) * 64 : ((N + 15) / 16) * 16; }
__host__ __device__ inline int tc_block_rows(int N) { return N >= 64 ? 64 : tc_npad(N); }

template <int BN>
struct Cfg {
    static constexpr int kNB = BN >= 64 ? 64 : BN;                      // rows of one weight block == MMA N
    static constexpr int kBlocks = BN / kNB;
    static constexpr uint32_t kBBlock = 2u * 8u * kNB * 16u;            // [plane][K group][kNB rows][16 B]
    static constexpr uint32_t kBStage = kBlocks * kBBlock;
    static constexpr uint32_t kStage = 2 * kAPlane + kBStage;
    static constexpr int kStagesRaw = (int)((227u * 1024u - 2048u) / kStage);
    static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
    static constexpr int kMainAcc = BN >= 256 ? 1 : (BN >= 128 ? 3 : 4);
    static constexpr uint32_t kAccCols = (kMainAcc + 1) * BN;
    static constexpr uint32_t kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
    static constexpr uint32_t kSmemBytes = kStages * kStage + 1024;
    static_assert(kStages >= 2, "pipeline needs at least two stages");
    static_assert(kAccCols <= 512, "TMEM has 512 columns");
};

// ---- A-operand fetch: 8 rows x 8 consecutive k (two float4) per thread and K chunk ------------------------------
template <int MODE>
__device__ __forceinline__ void fetch_a(const GemmParams& p, const ARow (&rows)[8], int k0, int kg, float4 (&buf)[16]) {
    if constexpr (MODE == LD_GATHER) {
        const int k = k0 + kg * 8;
        const bool k_ok = k < p.K;                 // K % 8 == 0 on this path
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (rows[i].valid && k_ok) {
                v0 = __ldg(reinterpret_cast<const float4*>(rows[i].base + k));
                v1 = __ldg(reinterpret_cast<const float4*>(rows[i].base + k + 4));
            }
            buf[2 * i] = v0;
            buf[2 * i + 1] = v1;
        }
    } else if constexpr (MODE == LD_CONV) {
        // C % 64 == 0: a 64-wide K chunk lies inside one filter tap
        const int tap = k0 / p.C;
        const int c0 = k0 - tap * p.C + kg * 8;
        const int kh = tap / p.KW;
        const int kw = tap - kh * p.KW;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            const int ih = rows[i].ih0 + kh, iw = rows[i].iw0 + kw;
            if (rows[i].valid && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                const float* src = rows[i].base + ((size_t)ih * p.W + iw) * p.C + c0;
                v0 = __ldg(reinterpret_cast<const float4*>(src));
                v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            }
            buf[2 * i] = v0;
            buf[2 * i + 1] = v1;
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            buf[2 * i] = load_a4(p, rows[i], k0 + kg * 8);
            buf[2 * i + 1] = load_a4(p, rows[i], k0 + kg * 8 + 4);
        }
    }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int BN, bool LN, int MODE>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const GemmParams p, const int npad, const int variant,
                                                              long long* __restrict__ ts) {
    using C = Cfg<BN>;
    // debug timeline (ts != null): slot layout documented in tools/bringup.py::gemm_timeline
    long long* my_ts = ts ? ts + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 64 : nullptr;
    const long long t_start = ts ? clock64() : 0;
#define COTR_TS(slot) do { if (my_ts) my_ts[(slot)] = clock64() - t_start; } while (0)
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
    if (threadIdx.x == 0) COTR_TS(1);

    if (warp < 4) {
        // ================= A producer: fp32 global -> fp16 hi/lo core-matrix rows in shared memory ============
        const int t = threadIdx.x;
        const int kg = t & 7;          // 16-byte K group (8 halves) inside the 64-wide chunk
        const int rb = t >> 3;         // rows rb, rb+16, ..., rb+112
        ARow rows[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rows[i] = decode_a_row(p, m0 + rb + 16 * i);

        float4 cur[16], nxt[16];
        fetch_a<MODE>(p, rows, 0, kg, cur);
        if (threadIdx.x == 0) COTR_TS(2);
#pragma unroll 1
        for (int it = 0; it < KC; ++it) {
            const int s = it % C::kStages;
            const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
            if (it + 1 < KC) fetch_a<MODE>(p, rows, (it + 1) * BK, kg, nxt);
            mbar_wait(&empty[s], ph ^ 1u);
            if (threadIdx.x == 0 && it < 8) COTR_TS(3 + 2 * it);
            uint8_t* a_hi = stage_base + (size_t)s * C::kStage + (uint32_t)kg * kALbo + (uint32_t)rb * 16u;
            uint8_t* a_lo = a_hi + kAPlane;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v0 = cur[2 * i], v1 = cur[2 * i + 1];
                uint4 hi, lo;
                split_f16x2(v0.x, v0.y, hi.x, lo.x);
                split_f16x2(v0.z, v0.w, hi.y, lo.y);
                split_f16x2(v1.x, v1.y, hi.z, lo.z);
                split_f16x2(v1.z, v1.w, hi.w, lo.w);
                *reinterpret_cast<uint4*>(a_hi + i * 256) = hi;      // 16 rows x 16 bytes further down
                *reinterpret_cast<uint4*>(a_lo + i * 256) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&full_a[s]);
            if (threadIdx.x == 0 && it < 8) COTR_TS(4 + 2 * it);
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }

        // ================= epilogue: TMEM -> registers -> global =============================================
        mbar_wait(accum_full, 0);
        tcgen05_fence_after();
        if (threadIdx.x == 0) COTR_TS(20);
        const int row = m0 + warp * 32 + lane;
        const bool row_ok = row < p.M;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float* add_row = (row_ok && p.addmat) ? p.addmat + (size_t)(row % p.add_period) * p.ld_add : nullptr;
        const float* res_row = (row_ok && p.residual) ? p.residual + (size_t)row * p.ldr : nullptr;
        float* out_row = p.out + (size_t)(row_ok ? row : 0) * p.ldc;
        const float acc_scale = p.acc_scale;

        // v[0..15] = sum over all accumulators of columns [c, c+16)
        auto load_acc = [&](int c, float (&v)[16]) {
            __syncwarp();
            tmem_ld16(trow + c, v);
#pragma unroll
            for (int a = 1; a <= C::kMainAcc; ++a) {
                float w2[16];
                tmem_ld16(trow + a * BN + c, w2);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += w2[j];
            }
        };
        // x += src[0..15] (vectorised; all row operands are 16-byte aligned on this path)
        auto add16 = [&](const float* src, float (&v)[16]) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 t4 = ldg4(src + j);
                v[j] += t4.x; v[j + 1] += t4.y; v[j + 2] += t4.z; v[j + 3] += t4.w;
            }
        };

        if constexpr (!LN) {
            const bool vec_ok = (p.ldc & 3) == 0 && (p.N & 15) == 0 && (p.ld_add & 3) == 0 && (p.ldr & 3) == 0;
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                load_acc(c, v);
                const int nb = n0 + c;
                if (!row_ok || nb >= p.N) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= acc_scale;
                if (vec_ok) {
                    if (p.bias) add16(p.bias + nb, v);
                    if (add_row) add16(add_row + nb, v);
                    if (res_row) add16(res_row + nb, v);
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(out_row + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll 1
                    for (int j = 0; j < 16; ++j) {
                        if (nb + j >= p.N) break;
                        float x = v[j];
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
                load_acc(c, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= acc_scale;
                if (p.bias) add16(p.bias + c, v);
                if (add_row) add16(add_row + c, v);
                if (res_row) add16(res_row + c, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sum += v[j];
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
                for (int j = 0; j < 16; j += 4) {
                    const float4 g4 = ldg4(p.ln_gamma + c + j), b4 = ldg4(p.ln_beta + c + j);
                    float4 o;
                    o.x = (v[j] - mean) * rstd * g4.x + b4.x;
                    o.y = (v[j + 1] - mean) * rstd * g4.y + b4.y;
                    o.z = (v[j + 2] - mean) * rstd * g4.z + b4.z;
                    o.w = (v[j + 3] - mean) * rstd * g4.w + b4.w;
                    *reinterpret_cast<float4*>(out_row + c + j) = o;
                }
            }
        }
        if (threadIdx.x == 0) COTR_TS(21);
    } else if (warp == 4) {
        // ================= weight producer: bulk TMA of the pre-tiled fp16 hi/lo image ==========================
        if (lane == 0) {
            const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.Wtc);
            const int blocks_total = npad / C::kNB;
            const int blk0 = n0 / C::kNB;
#pragma unroll 1
            for (int it = 0; it < KC; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], C::kBStage);
                uint8_t* b_dst = stage_base + (size_t)s * C::kStage + 2 * kAPlane;
                // image: [k chunk][64-row block][plane][K group][64 rows][16 bytes]; the blocks of one tile are adjacent
                const uint8_t* src = wimg + ((size_t)it * blocks_total + blk0) * C::kBBlock;
#pragma unroll
                for (int j = 0; j < C::kBlocks; ++j)
                    tma_bulk_g2s(b_dst + (size_t)j * C::kBBlock, src + (size_t)j * C::kBBlock, C::kBBlock, &full_b[s]);
                if (it < 8) COTR_TS(44 + it);
            }
        }
        __syncwarp();
    } else {
        // ================= MMA issuer ===========================================================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, C::kNB);
            constexpr uint32_t b_lbo = C::kNB * 16;
            constexpr uint32_t b_plane = 8 * b_lbo;
#pragma unroll 1
            for (int it = 0; it < KC; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = (uint32_t)(it / C::kStages) & 1u;
                mbar_wait(&full_a[s], ph);
                if (it < 8) COTR_TS(24 + 2 * it);
                mbar_wait(&full_b[s], ph);
                tcgen05_fence_after();
                const uint32_t a_hi = smem_u32(stage_base + (size_t)s * C::kStage);
                const uint32_t a_lo = a_hi + kAPlane;
                const uint32_t b_base = a_hi + 2 * kAPlane;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const int g = it * (BK / 16) + ks;                       // global K step
                    const uint32_t ao = ks * 2 * kALbo, bo = ks * 2 * b_lbo;
                    const uint64_t dah = (variant & 1) ? make_smem_desc(a_hi + ao, kSbo, kALbo) : make_smem_desc(a_hi + ao, kALbo, kSbo);
                    const uint64_t dal = (variant & 1) ? make_smem_desc(a_lo + ao, kSbo, kALbo) : make_smem_desc(a_lo + ao, kALbo, kSbo);
                    const uint32_t main_col = (uint32_t)(g % C::kMainAcc) * BN;
                    const uint32_t corr_col = (uint32_t)C::kMainAcc * BN;
#pragma unroll
                    for (int j = 0; j < C::kBlocks; ++j) {
                        const uint32_t bh = b_base + j * C::kBBlock + bo, bl = bh + b_plane;
                        const uint64_t dbh = (variant & 1) ? make_smem_desc(bh, kSbo, b_lbo) : make_smem_desc(bh, b_lbo, kSbo);
                        const uint64_t dbl = (variant & 1) ? make_smem_desc(bl, kSbo, b_lbo) : make_smem_desc(bl, b_lbo, kSbo);
                        const uint32_t col = tmem_base + j * C::kNB;
                        umma_f16_ss(col + corr_col, dal, dbh, idesc, g != 0);
                        umma_f16_ss(col + corr_col, dah, dbl, idesc, true);
                        umma_f16_ss(col + main_col, dah, dbh, idesc, g >= C::kMainAcc);
                    }
                }
                umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
                if (it < 8) COTR_TS(25 + 2 * it);
            }
            umma_commit(accum_full);
            COTR_TS(41);
        }
        __syncwarp();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, C::kTmemCols);
    if (threadIdx.x == 160) COTR_TS(60);
#undef COTR_TS
}

template <int BN, bool LN, int MODE>
int launch_one(const GemmParams& p, cudaStream_t s) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        COTR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, LN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
        configured = true;
    }
    const int npad = tc_npad(p.N);
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
    gemm_tc_kernel<BN, LN, MODE><<<grid, kThreads, C::kSmemBytes, s>>>(p, npad, g_tc_variant, g_tc_timestamps);
    COTR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN, bool LN>
int launch_mode(const GemmParams& p, cudaStream_t s) {
    const bool gather = (p.a_mode == A_ROWMAJOR || p.a_mode == A_TOKENS);
    if (gather && (p.K & 7) == 0 && (p.lda & 3) == 0) return launch_one<BN, LN, LD_GATHER>(p, s);
    if constexpr (!LN && BN >= 64) {
        if (p.a_mode == A_CONV_NHWC && (p.C & 63) == 0) return launch_one<BN, LN, LD_CONV>(p, s);
    }
    if constexpr (!LN && BN == 64) return launch_one<64, false, LD_GENERIC>(p, s);
    set_error("gemm_tc: no kernel instantiation for a_mode %d, K %d, lda %d with tile N %d", p.a_mode, p.K, p.lda, BN);
    return 1;
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

// Image layout: [k chunk (64)][row block (64 rows; 16 when N < 64)][plane: hi, lo][K group (8 halves)][row][8 halves],
// zero padded in N and K.  The matrix is multiplied by 2^e, e chosen so that max|w| * 2^e lies in [2^12, 2^13);
// returns 2^-e for the epilogue.
float tc_pack_weight(const float* w, int N, int K, void* dst_host) {
    const int npad = tc_npad(N);
    const int nb = tc_block_rows(N);
    const int blocks = npad / nb;
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
        for (int b = 0; b < blocks; ++b)
            for (int kg = 0; kg < 8; ++kg)
                for (int r = 0; r < nb; ++r)
                    for (int el = 0; el < 8; ++el) {
                        const int n = b * nb + r;
                        const int k = kc * BK + kg * 8 + el;
                        const float x = (n < N && k < K) ? w[(size_t)n * K + k] * scale : 0.f;
                        const uint16_t hi = f32_to_f16_rn(x);
                        const uint16_t lo = f32_to_f16_rn(x - f16_to_f32(hi));
                        const size_t blk = ((size_t)kc * blocks + b) * 2;
                        out[(((blk + 0) * 8 + kg) * nb + r) * 8 + el] = hi;
                        out[(((blk + 1) * 8 + kg) * nb + r) * 8 + el] = lo;
                    }
    return ldexpf(1.f, -e);
}

int launch_gemm_tc(const GemmParams& p, cudaStream_t s) {
    COTR_CHECK(p.M > 0 && p.N > 0 && p.K > 0, "gemm_tc: empty problem %d x %d x %d", p.M, p.N, p.K);
    COTR_CHECK(p.Wtc != nullptr, "gemm_tc: weight has no tensor-core image");
    COTR_CHECK(p.a_mode != A_CONV_NHWC || (p.C & 3) == 0, "gemm_tc: NHWC conv needs C %% 4 == 0 (C=%d)", p.C);
    if (p.ln_gamma) {
        COTR_CHECK(p.N == 256 && p.ldc == 256 && p.relu == 0, "gemm_tc: LayerNorm epilogue needs N = ldc = 256");
        COTR_CHECK((p.ldr & 3) == 0 && (p.ld_add & 3) == 0, "gemm_tc: LayerNorm epilogue needs 16-byte aligned row operands");
        return launch_mode<256, true>(p, s);
    }
    if (p.N < 64) {
        COTR_CHECK(p.N <= 16, "gemm_tc: N between 17 and 63 is not instantiated");
        return launch_mode<16, false>(p, s);
    }
    const int mt = (p.M + BM - 1) / BM;
    if ((p.N % 128) == 0 && (long long)mt * (p.N / 128) >= 120) return launch_mode<128, false>(p, s);
    return launch_mode<64, false>(p, s);
}

}  // namespace cotr
