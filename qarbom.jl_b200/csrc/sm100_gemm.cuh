// tcgen05 / TMEM / TMA fused Gibbs-step GEMM for sm_100a (QB_PREC_BF16 path).
//
//   D[m][n] = sum_k A[m][k] * B[n][k]  ( - sum_k A2[m][k] * B2[n][k] )      bf16 x bf16 -> fp32 in TMEM
//
// Both operands are K-major (row = m or n, contiguous k), staged by TMA into 128B-swizzled
// shared memory, multiplied by single-thread tcgen05.mma (cta_group::1, M=128, N=BN, K=16)
// into one of two TMEM accumulator stages, and drained by two epilogue warpgroups while the
// MMA warp already works on the next tile (persistent CTAs, static tile schedule).
//
// Orientation for the Gibbs half-steps: UNITS are on M (TMEM lanes, one thread per unit),
// batch ROWS are on N (TMEM columns).  So the weight matrix is the A operand and the state
// matrix the B operand; bias and bias-gradient sums are thread-local, row-major state
// stores are lane-coalesced and transposed state stores are per-thread contiguous.
//
// Epilogues (all fused, nothing but the requested outputs ever reaches HBM):
//   EPI_SAMPLE: x = acc + bias[unit];  p = sigmoid(x) | x (Gaussian mean / raw)
//               s = (u < p) | x + z;   u,z from Philox4x32-10 or injected arrays
//               optional stores: state row-major / transposed (bf16), prob row-major /
//               transposed (bf16), prob fp32; optional reductions: bias gradient
//               (thread-local), squared reconstruction error, softplus row sums.
//   EPI_GRAD:   fp32 store of the accumulator (the phase-2 product is negated in the MMA).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace qb {
namespace sm100 {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARP0 = 4;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4.. epilogue
constexpr int SMEM_AB_BUDGET = 192 * 1024;

enum { EPI_SAMPLE = 0, EPI_GRAD = 1 };
// kernel specialisations: generic epilogue, gradient store, and the two compile-time fast paths
// (numbering = QB_GEMM_KIND_* of include/qarbom_b200.h)
enum { K_GENERIC = 0, K_GRAD = 1, K_FAST_SAMPLE = 2, K_FAST_PROB = 3, K_FAST_SAMPLE_PROB = 4, K_FAST_GAUSS = 5, K_FAST_LABEL = 6,
       K_CHAIN = 7 /* histogram slot of the multi-half-step kernel (sm100_chain.cuh), not an epilogue */,
       K_GRAD_LONG = 9 /* gradient-store epilogue with a product list longer than two (QB_PREC_FP32X3's component products): the
                          only kernels that index the list at run time (slot 8 is the multi-mini-batch kernel's) */,
       K_FAST_SQERR = 10 /* reconstruction error: p = sigmoid(x), sum of (ref - p)^2 -- nothing is stored (evaluate /
                            _evaluate(MeanSquaredError), src/evaluation.jl:16-20, and the per-step error of qb_*_step_host) */ };
// Epilogue warpgroups per CTA: the lean epilogues are latency-bound with 2 warps per scheduler
// (ncu: IPC 0.3, epilogue time ~ MMA time), so they run 4 warpgroups (16 warps, 64 columns each at
// BN = 256); the register-hungry generic epilogue keeps 2.
template <int BN, int KIND>
struct Shape {
    static constexpr int WGS = (KIND != K_GENERIC && BN >= 64) ? 4 : 2;
    static constexpr int THREADS = 128 + 128 * WGS;
};
enum { ACT_SIGMOID = 0, ACT_GAUSSIAN = 1, ACT_GAUSS_THRESH = 2 };  // THRESH: 1[u < x + z] (FastPCD chains of a GRBM, src/gibbs.jl:22-25)

struct EpiParams {
    int mode;
    int M, N;  // valid extents of D (M: units or H; N: batch rows or VL)
    // ---- EPI_SAMPLE
    const float *bias;
    int act;
    int sample;  // draw a state (Bernoulli / Gaussian noise); 0 = probabilities / mean only
    int split;   // units >= split are label units: raw pre-activation goes to lab_pre only
    float *lab_pre; long long ld_lab;
    bf16 *state_rm; long long ld_srm;  // [rows][units]
    bf16 *state_T; long long ld_sT;    // [units][rows]
    bf16 *prob_rm; long long ld_prm;
    bf16 *prob_T; long long ld_pT;
    float *prob_f32; long long ld_pf;  // [rows][units]
    int f32_from_state;                // prob_f32 receives the sampled / noisy value instead of p
    const float *inj; long long ld_inj;  // [rows][units] injected uniforms (Bernoulli) or normals (Gaussian)
    RngKey key; long long row0; int use_rng;
    float *bias_grad; float bg_sign; int bg_what;  // 1: probabilities/mean, 2: sampled state
    const bf16 *ref_rm; long long ld_ref; double *sqerr;
    float *softplus_rows;
    // ---- EPI_GRAD
    float *out_f32; long long ld_out;
    int grad_sub;  // out = out - acc instead of out = acc (negative statistics subtracted from an already reduced positive part)
    // EPI_GRAD, integer form: the accumulators are exact small integers (a product of binary states) and are stored as uint16
    // [M][ld_out] instead of fp32 (the negative product of the split gradient exchange, p2p_kernels.cuh)
    unsigned short *out_u16;
    // affine form (QB_PREC_FP32X3's GEMM): out = g_alpha * acc + g_beta * out + g_bias[n]; active when g_affine != 0
    int g_affine; float g_alpha, g_beta; const float *g_bias;
    float pscale;  // probabilities are multiplied by this before being stored / summed (n_chains / batch; 1 normally)
};

// A launch accumulates a LIST of products into one TMEM tile: D = sum_p (+/-) A_p B_p^T.  One phase is a plain GEMM, two
// (the second negated) the CD / PCD gradient, and the fp32-grade mode (QB_PREC_FP32X3) expands every real-valued operand into
// bf16 components (x = hi + mid + lo, 24 mantissa bits) and lists the component products that matter (i + j <= 2).
constexpr int MAX_PHASES = 12;
struct Phases {
    int n;
    int kb_end[MAX_PHASES];   // cumulative 64-deep k-blocks at the end of each phase
    unsigned int neg_mask;    // bit p: phase p is subtracted (a_negate in the instruction descriptor)
    CUtensorMap A[MAX_PHASES], B[MAX_PHASES];
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// One lane of a CONVERGED warp.  The TMA and MMA warps run their loops with all 32 lanes (uniform control flow, operands in
// uniform registers) and predicate only the tcgen05 / TMA instructions on this: inside an `if (lane == 0)` region the
// compiler has to broadcast every descriptor into uniform registers with an ELECT / R2UR.BROADCAST loop per instruction,
// which costs ~100 clocks per tcgen05.mma (393 clocks per k-step of four, measured with tools/epoch_trace.py) -- more than
// the MMA itself for tiles narrower than 256.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Non-blocking look at a barrier phase.  The MMA warp tests the NEXT k-step's barrier before it issues the current k-step's MMAs,
// so the ~170-clock round trip of a wait overlaps the issue instead of preceding it (tools/mma_issue_probe.cu: 326 -> 256
// clocks per k-step of four 128 x N x 16 MMAs for N <= 128, against 432 with a single-lane issuer).
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Predicated forms for the warp-uniform issue loops: `issue` is 1 in the elected lane and 0 elsewhere, and the predicate sits
// INSIDE the asm, so there is no branch (no divergent region to enter and leave) around the tcgen05 / TMA instruction.  With a
// C++ `if (issuer)` around four MMAs a k-step costs 256 clocks whatever the tile width; predicated, it runs at the rate the
// MMAs execute (tools/mma_issue_probe.cu, patterns 8 / 11: N = 32: 256 -> 165 clocks, N = 64: 256 -> 191, N = 128: 256 -> 253).
// Used by the single-CTA kernels (this file's gibbs_gemm_kernel, sm100_epoch.cuh), whose tiles are narrow.  The CTA-pair
// kernels keep the `if (issuer)` form: their operands depend on the CTA's rank in the pair, which the compiler cannot prove
// warp-uniform, so every predicated instruction drags a row of R2UR.BROADCAST with it -- A/B on one box 4160 vs 3940 us per
// cfg5 step, and their 256-wide tiles have the clocks to hide a branch anyway.
__device__ __forceinline__ void mbar_expect_tx_if(uint32_t issue, uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_if(uint32_t issue, uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t issue, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_if(uint32_t issue, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16]) {
    // the registers are listed as in/out so no consumer can be scheduled above the wait
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused: one swizzle atom along K) | SBO>>4 [32,46) = 1024 B
// (8 rows x 128 B) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1,
// b=BF16 [10,13)=1, a_negate bit 13, a/b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool negate_a) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((negate_a ? 1u : 0u) << 13) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

template <int BN>
struct Cfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_AB_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_AB_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;  // two accumulator stages
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ----------------------------------------------------------------------------- epilogues
template <int BN, int WGS>
__device__ __forceinline__ void epilogue_sample(const EpiParams &ep, uint32_t tmem_acc, int m_blk, int n_blk, int quarter,
                                                int wg, int lane) {
    const int unit = m_blk * BM + quarter * 32 + lane;
    const bool uvalid = unit < ep.M;
    const bool is_label = uvalid && unit >= ep.split;
    const float bias = uvalid ? ep.bias[unit] : 0.f;
    float bg_acc = 0.f, sq_acc = 0.f;
    constexpr int COLS_PER_WG = BN / WGS;
#pragma unroll 1
    for (int c0 = wg * COLS_PER_WG; c0 < (wg + 1) * COLS_PER_WG; c0 += 16) {
        uint32_t r[16];
        __syncwarp();  // re-converge lanes that skipped stores before the .aligned TMEM load
        tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld(r);
        const int row_base = n_blk * BN + c0;
        if (row_base >= ep.N) continue;  // warp-uniform
        const bool full = (row_base + 16 <= ep.N);
        float pv[16], sv[16];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t rnd[4] = {0, 0, 0, 0};
            float zn[4] = {0.f, 0.f, 0.f, 0.f};
            if (ep.sample && !ep.inj && ep.use_rng) {
                const unsigned long long grow = (unsigned long long)(ep.row0 + row_base + 4 * g);
                if (ep.act == ACT_SIGMOID) {
                    Philox4 ph = rng_uniform4(ep.key, (uint32_t)unit, (uint32_t)(grow >> 2));
                    rnd[0] = ph.x; rnd[1] = ph.y; rnd[2] = ph.z; rnd[3] = ph.w;
                } else {
                    rng_normal4(ep.key, (uint32_t)unit, (uint32_t)(grow >> 2), zn);
                    if (ep.act == ACT_GAUSS_THRESH) {
                        Philox4 ph = rng_uniform4(ep.key, (uint32_t)unit, (uint32_t)(grow >> 2));
                        rnd[0] = ph.x; rnd[1] = ph.y; rnd[2] = ph.z; rnd[3] = ph.w;
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = 4 * g + jj;
                const int row = row_base + j;
                const float x = __uint_as_float(r[j]) + bias;
                float p, s;
                if (ep.act == ACT_SIGMOID && !is_label) {
                    p = sigmoid_fast(x);
                    s = p;
                    if (ep.sample) {
                        float u;
                        if (ep.inj) u = (uvalid && row < ep.N) ? ep.inj[(long long)row * ep.ld_inj + unit] : 1.f;
                        else u = u23(rnd[jj]);
                        s = (u < p) ? 1.f : 0.f;
                    }
                } else {
                    p = x;
                    s = x;
                    if (ep.sample && !is_label) {
                        float z = zn[jj];
                        if (ep.inj) z = (uvalid && row < ep.N) ? ep.inj[(long long)row * ep.ld_inj + unit] : 0.f;
                        s = x + z;
                        if (ep.act == ACT_GAUSS_THRESH) s = (u23(rnd[jj]) < s) ? 1.f : 0.f;
                    }
                }
                pv[j] = p;
                sv[j] = s;
            }
        }
        if (ep.softplus_rows) {  // free energy: sum over units of softplus(pre) per batch row
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float sp = (uvalid && !is_label) ? softplus_f(pv[j]) : 0.f;  // act must be raw (pv == pre)
                sp = warp_sum(sp);
                if (lane == 0 && row_base + j < ep.N) atomicAdd(ep.softplus_rows + row_base + j, sp);
            }
            continue;
        }
        if (!uvalid) continue;
        if (is_label) {  // label units: raw pre-activations, finished by label_kernel
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (row_base + j < ep.N) ep.lab_pre[(long long)(row_base + j) * ep.ld_lab + (unit - ep.split)] = pv[j];
            continue;
        }
        if (ep.bg_what) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (full || row_base + j < ep.N) bg_acc += (ep.bg_what == 1) ? pv[j] : sv[j];
        }
        if (ep.sqerr) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (full || row_base + j < ep.N) {
                    float d = __bfloat162float(ep.ref_rm[(long long)(row_base + j) * ep.ld_ref + unit]) - sv[j];
                    sq_acc = fmaf(d, d, sq_acc);
                }
        }
        if (ep.state_rm) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (full || row_base + j < ep.N)
                    ep.state_rm[(long long)(row_base + j) * ep.ld_srm + unit] = __float2bfloat16(sv[j]);
        }
        if (ep.prob_rm) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (full || row_base + j < ep.N)
                    ep.prob_rm[(long long)(row_base + j) * ep.ld_prm + unit] = __float2bfloat16(pv[j]);
        }
        if (ep.prob_f32) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (full || row_base + j < ep.N) ep.prob_f32[(long long)(row_base + j) * ep.ld_pf + unit] = ep.f32_from_state ? sv[j] : pv[j];
        }
        if (ep.state_T) {
            bf16 *dst = ep.state_T + (long long)unit * ep.ld_sT + row_base;
            if (full) {
                uint4 v0 = make_uint4(pack_bf16x2(sv[0], sv[1]), pack_bf16x2(sv[2], sv[3]), pack_bf16x2(sv[4], sv[5]),
                                      pack_bf16x2(sv[6], sv[7]));
                uint4 v1 = make_uint4(pack_bf16x2(sv[8], sv[9]), pack_bf16x2(sv[10], sv[11]), pack_bf16x2(sv[12], sv[13]),
                                      pack_bf16x2(sv[14], sv[15]));
                reinterpret_cast<uint4 *>(dst)[0] = v0;
                reinterpret_cast<uint4 *>(dst)[1] = v1;
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (row_base + j < ep.N) dst[j] = __float2bfloat16(sv[j]);
            }
        }
        if (ep.prob_T) {
            bf16 *dst = ep.prob_T + (long long)unit * ep.ld_pT + row_base;
            if (full) {
                uint4 v0 = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                                      pack_bf16x2(pv[6], pv[7]));
                uint4 v1 = make_uint4(pack_bf16x2(pv[8], pv[9]), pack_bf16x2(pv[10], pv[11]), pack_bf16x2(pv[12], pv[13]),
                                      pack_bf16x2(pv[14], pv[15]));
                reinterpret_cast<uint4 *>(dst)[0] = v0;
                reinterpret_cast<uint4 *>(dst)[1] = v1;
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (row_base + j < ep.N) dst[j] = __float2bfloat16(pv[j]);
            }
        }
    }
    if (ep.bg_what && uvalid && !is_label) atomicAdd(ep.bias_grad + unit, ep.bg_sign * bg_acc);
    if (ep.sqerr) {
        sq_acc = warp_sum(sq_acc);
        if (lane == 0) atomicAdd(ep.sqerr, (double)sq_acc);
    }
}

// Compile-time-specialised epilogue for the hot Gibbs half-steps (Bernoulli units, Philox or no
// sampling, no label units, no metric reductions): a fraction of the generic epilogue's
// instruction count and code size, so it hides completely behind the next tile's MMAs.
//   SAMPLE = true : s = (u < sigmoid(x)); stores state_rm (required), state_T (optional);
//                   optional prob_T; bias gradient of s (bg_what == 2) or p (bg_what == 1)
//   SAMPLE = false: p = sigmoid(x); stores prob_T / prob_rm (optional); bias gradient of p
// sigmoid with exactly two MUFU ops and no range fix-ups: x -> -inf gives ex2 = +inf, rcp(inf) = 0
// sigmoid(acc + bias) with the bias folded into the exponent's multiply: ex2(acc * -log2e + bias * -log2e), one FFMA instead
// of FADD + FMUL (bias_l2e = bias * -log2e, once per thread)
__device__ __forceinline__ float sigmoid_mufu_fma(float acc, float bias_l2e) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(acc, -1.4426950408889634f, bias_l2e)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ float sigmoid_mufu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// Box-Muller with four MUFU ops for two normals (same formula as box_muller2 in common.cuh; lg2 / sqrt / sin / cos are the
// approximate hardware functions, |error| < 1e-5 absolute on z): the Gaussian-visible epilogue costs the same two MUFU per
// element as the sigmoid one.  The angle is taken around zero, 2 pi (u2 - 1/2) in [-pi, pi), where sin/cos.approx are accurate.
__device__ __forceinline__ void box_muller2_fast(uint32_t a, uint32_t b, float &z0, float &z1) {
    const float u1 = 1.0f - u23(a);
    const float ang = (u23(b) - 0.5f) * 6.283185307179586f;
    float l, r, sn, cs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -1.3862943611198906f));  // -2 ln u1 = -2 ln2 lg2 u1
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
    z0 = -r * cs; z1 = -r * sn;   // cos(x + pi) = -cos x, sin(x + pi) = -sin x
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one 16-column chunk of the fast epilogue; FULL = all 16 rows and the unit are in range (the common
// case: no per-element predicates at all).  GAUSS: s = x + z (Gaussian visibles, src/rbm.jl:187) instead of a Bernoulli draw.
template <bool SAMPLE, bool PROB, bool FULL, bool GAUSS = false>
__device__ __forceinline__ void fast_chunk(const EpiParams &ep, const uint32_t (&r)[16], int unit, int row_base, int nv, float bias,
                                           bool want_sT, bool want_pT, bool want_prm, float &bg_p, int &bg_s) {
    uint32_t sb[16], pb[16];
    const float bias_l2e = bias * -1.4426950408889634f;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t rnd[4];
        if (SAMPLE) {
            const unsigned long long grow = (unsigned long long)(ep.row0 + row_base + 4 * g);
            Philox4 ph = philox4x32_10((uint32_t)unit, (uint32_t)(grow >> 2), ep.key.draw_lo, ep.key.draw_hi | (GAUSS ? 0x80000000u : 0u),
                                       ep.key.seed_lo, ep.key.seed_hi);
            rnd[0] = ph.x; rnd[1] = ph.y; rnd[2] = ph.z; rnd[3] = ph.w;
        }
        if (GAUSS) {
            float z[4];
            box_muller2_fast(rnd[0], rnd[1], z[0], z[1]);
            box_muller2_fast(rnd[2], rnd[3], z[2], z[3]);
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = 4 * g + jj;
                const float sv = __uint_as_float(r[j]) + bias + z[jj];
                if (FULL || j < nv) bg_p += sv;
                sb[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(sv));
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = 4 * g + jj;
                const float p = sigmoid_mufu_fma(__uint_as_float(r[j]), bias_l2e);
                if (PROB) {
                    const float ps = p * ep.pscale;
                    if (FULL || j < nv) bg_p += ps;
                    pb[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(ps));
                }
                if (SAMPLE) {
                    sb[j] = (u23(rnd[jj]) < p) ? 0x3F80u : 0u;  // bf16(1.0) / bf16(0.0)
                    // bias-gradient count: the state bits are summed as they are (one add, no second use of the predicate) and
                    // divided by 0x3F80 once per thread at the end
                    bg_s += (FULL || j < nv) ? (int)sb[j] : 0;
                }
            }
        }
    }
    if (SAMPLE) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(ep.state_rm) + (long long)row_base * ep.ld_srm + unit;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (FULL || j < nv) *dst = (unsigned short)sb[j];
            dst += ep.ld_srm;
        }
    }
    if (PROB && want_prm) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(ep.prob_rm) + (long long)row_base * ep.ld_prm + unit;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (FULL || j < nv) *dst = (unsigned short)pb[j];
            dst += ep.ld_prm;
        }
    }
    if (SAMPLE && want_sT) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(ep.state_T) + (long long)unit * ep.ld_sT + row_base;
        if (FULL) {
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(sb[0] | (sb[1] << 16), sb[2] | (sb[3] << 16), sb[4] | (sb[5] << 16), sb[6] | (sb[7] << 16));
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(sb[8] | (sb[9] << 16), sb[10] | (sb[11] << 16), sb[12] | (sb[13] << 16), sb[14] | (sb[15] << 16));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (j < nv) dst[j] = (unsigned short)sb[j];
        }
    }
    if (PROB && want_pT) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(ep.prob_T) + (long long)unit * ep.ld_pT + row_base;
        if (FULL) {
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(pb[0] | (pb[1] << 16), pb[2] | (pb[3] << 16), pb[4] | (pb[5] << 16), pb[6] | (pb[7] << 16));
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(pb[8] | (pb[9] << 16), pb[10] | (pb[11] << 16), pb[12] | (pb[13] << 16), pb[14] | (pb[15] << 16));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (j < nv) dst[j] = (unsigned short)pb[j];
        }
    }
}

// Compile-time-specialised epilogue for the hot Gibbs half-steps (Bernoulli units, Philox or no
// sampling, no label units, no metric reductions): a fraction of the generic epilogue's
// instruction count and code size, so it hides completely behind the next tile's MMAs.
//   SAMPLE: s = (u < sigmoid(x)); stores state_rm (required), state_T (optional); bias gradient of s
//   PROB  : p = sigmoid(x); stores prob_T / prob_rm (optional); bias gradient of p (bg_what == 1)
//   MODE 1: Gaussian visibles, s = x + z (bias gradient of s)
//   MODE 2: Bernoulli visibles followed by label units (RBMClassifier): the one warp whose 32 lanes hold the label units
//           [split, M) finishes them in place -- softmax over the label lanes by warp shuffles, then independent Bernoulli
//           draws (gibbs_sample_label, src/gibbs.jl:46-49 + src/rbm.jl:196-205) -- every other warp runs the plain path.
template <bool FULL>
__device__ __forceinline__ void label_chunk(const EpiParams &ep, const uint32_t (&r)[16], int unit, bool uvalid, int row_base, int nrows,
                                            float bias, bool want_sT, int &bg_s) {
    const bool is_label = uvalid && unit >= ep.split;
    uint32_t sb[16];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const unsigned long long grow = (unsigned long long)(ep.row0 + row_base + 4 * g);
        Philox4 ph = rng_uniform4(ep.key, (uint32_t)unit, (uint32_t)(grow >> 2));
        const uint32_t rnd[4] = {ph.x, ph.y, ph.z, ph.w};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j = 4 * g + jj;
            const float x = __uint_as_float(r[j]) + bias;
            const float mx = warp_max(is_label ? x : -INFINITY);
            const float e = is_label ? __expf(x - mx) : 0.f;
            const float den = warp_sum(e);
            const float p = is_label ? __fdividef(e, den) : sigmoid_mufu(x);
            sb[j] = (u23(rnd[jj]) < p) ? 0x3F80u : 0u;
            bg_s += (FULL || j < nrows) ? (int)sb[j] : 0;   // (summed state bits, / 0x3F80 at the end: see fast_chunk)
        }
    }
    if (!uvalid) return;
    unsigned short *dst = reinterpret_cast<unsigned short *>(ep.state_rm) + (long long)row_base * ep.ld_srm + unit;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (FULL || j < nrows) *dst = (unsigned short)sb[j];
        dst += ep.ld_srm;
    }
    if (want_sT) {
        unsigned short *dT = reinterpret_cast<unsigned short *>(ep.state_T) + (long long)unit * ep.ld_sT + row_base;
        if (FULL) {
            reinterpret_cast<uint4 *>(dT)[0] = make_uint4(sb[0] | (sb[1] << 16), sb[2] | (sb[3] << 16), sb[4] | (sb[5] << 16), sb[6] | (sb[7] << 16));
            reinterpret_cast<uint4 *>(dT)[1] = make_uint4(sb[8] | (sb[9] << 16), sb[10] | (sb[11] << 16), sb[12] | (sb[13] << 16), sb[14] | (sb[15] << 16));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (j < nrows) dT[j] = (unsigned short)sb[j];
        }
    }
}

//   MODE 3: reconstruction error only: p = sigmoid(x), acc += (ref - p)^2, no stores
template <bool FULL>
__device__ __forceinline__ void sqerr_chunk(const EpiParams &ep, const uint32_t (&r)[16], int unit, int row_base, int nv, float bias, float &acc) {
    const float bias_l2e = bias * -1.4426950408889634f;
    const unsigned short *ref = reinterpret_cast<const unsigned short *>(ep.ref_rm) + (long long)row_base * ep.ld_ref + unit;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float p = sigmoid_mufu_fma(__uint_as_float(r[j]), bias_l2e);
        if (FULL || j < nv) {
            const float d = __uint_as_float((uint32_t)ref[(long long)j * ep.ld_ref] << 16) - p;   // bf16 -> fp32
            acc = fmaf(d, d, acc);
        }
    }
}

template <int BN, int WGS, bool SAMPLE, bool PROB, int MODE = 0>
__device__ __forceinline__ void epilogue_fast(const EpiParams &ep, uint32_t tmem_acc, int m_blk, int n_blk, int quarter, int wg,
                                              int lane) {
    const int unit = m_blk * BM + quarter * 32 + lane;
    const bool uvalid = unit < ep.M;
    const float bias = uvalid ? __ldcg(ep.bias + unit) : 0.f;   // L2: sm100_epoch.cuh rewrites the biases between iterations
    const bool want_sT = SAMPLE && ep.state_T != nullptr, want_pT = PROB && ep.prob_T != nullptr, want_prm = PROB && ep.prob_rm != nullptr;
    // MODE 2: does this warp hold the label units? (warp-uniform; pick_kind guarantees they sit inside ONE 32-lane group)
    const bool label_warp = MODE == 2 && (unit - lane) + 32 > ep.split && (unit - lane) < ep.M;
    float bg_p = 0.f;
    int bg_s = 0;
    constexpr int COLS_PER_WG = BN / WGS;
#pragma unroll 1
    for (int c0 = wg * COLS_PER_WG; c0 < (wg + 1) * COLS_PER_WG; c0 += 16) {
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld(r);
        const int row_base = n_blk * BN + c0;
        int nvalid = ep.N - row_base;  // rows of this chunk inside the matrix (warp-uniform)
        if (nvalid <= 0) continue;
        nvalid = nvalid > 16 ? 16 : nvalid;
        const int nv = uvalid ? nvalid : 0;
        if (label_warp) {   // all 32 lanes take part in the shuffles, stores are predicated per lane
            if (nvalid == 16) label_chunk<true>(ep, r, unit, uvalid, row_base, nvalid, bias, want_sT, bg_s);
            else label_chunk<false>(ep, r, unit, uvalid, row_base, nvalid, bias, want_sT, bg_s);
            continue;
        }
        if (MODE == 3) {
            if (nv == 16) sqerr_chunk<true>(ep, r, unit, row_base, nv, bias, bg_p);
            else if (nv > 0) sqerr_chunk<false>(ep, r, unit, row_base, nv, bias, bg_p);
            continue;
        }
        if (nv == 16) fast_chunk<SAMPLE, PROB, true, MODE == 1>(ep, r, unit, row_base, nv, bias, want_sT, want_pT, want_prm, bg_p, bg_s);
        else if (nv > 0) fast_chunk<SAMPLE, PROB, false, MODE == 1>(ep, r, unit, row_base, nv, bias, want_sT, want_pT, want_prm, bg_p, bg_s);
    }
    if (MODE == 3) {   // one double atomic per warp and tile
        const float sq = warp_sum(bg_p);
        if (lane == 0) atomicAdd(ep.sqerr, (double)sq);
        return;
    }
    if (ep.bg_what && uvalid)
        atomicAdd(ep.bias_grad + unit, ep.bg_sign * (((PROB && ep.bg_what == 1) || MODE == 1) ? bg_p : (float)(bg_s / 0x3F80)));
}

template <int BN, int WGS>
__device__ __forceinline__ void epilogue_grad(const EpiParams &ep, uint32_t tmem_acc, int m_blk, int n_blk, int quarter, int wg,
                                              int lane) {
    const int m = m_blk * BM + quarter * 32 + lane;
    constexpr int COLS_PER_WG = BN / WGS;
#pragma unroll 1
    for (int c0 = wg * COLS_PER_WG; c0 < (wg + 1) * COLS_PER_WG; c0 += 16) {
        uint32_t r[16];
        __syncwarp();  // re-converge lanes that skipped stores before the .aligned TMEM load
        tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld(r);
        const int n0 = n_blk * BN + c0;
        if (n0 >= ep.N || m >= ep.M) continue;
        if (ep.out_u16) {
            unsigned short *d16 = ep.out_u16 + (long long)m * ep.ld_out + n0;
            if (n0 + 16 <= ep.N) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    pk[j] = __float2uint_rn(__uint_as_float(r[2 * j])) | (__float2uint_rn(__uint_as_float(r[2 * j + 1])) << 16);
                reinterpret_cast<uint4 *>(d16)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);   // (ld_out % 8 == 0, n0 % 16 == 0: 16-byte aligned)
                reinterpret_cast<uint4 *>(d16)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (n0 + j < ep.N) d16[j] = (unsigned short)__float2uint_rn(__uint_as_float(r[j]));
            }
            continue;
        }
        float *dst = ep.out_f32 + (long long)m * ep.ld_out + n0;
        if (ep.g_affine) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (n0 + j < ep.N) {
                    float v = ep.g_alpha * __uint_as_float(r[j]);
                    if (ep.g_beta != 0.f) v = fmaf(ep.g_beta, dst[j], v);
                    if (ep.g_bias) v += ep.g_bias[n0 + j];
                    r[j] = __float_as_uint(v);
                }
        }
        if (ep.grad_sub) {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (n0 + j < ep.N) r[j] = __float_as_uint(dst[j] - __uint_as_float(r[j]));
        }
        if (n0 + 16 <= ep.N && (ep.ld_out & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                reinterpret_cast<uint4 *>(dst)[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (n0 + j < ep.N) dst[j] = __uint_as_float(r[j]);
        }
    }
}

// ----------------------------------------------------------------------------- the kernel
template <int BN, int KIND>
__global__ void __launch_bounds__((Shape<BN, KIND>::THREADS), 1)
gibbs_gemm_kernel(const __grid_constant__ Phases ph, const EpiParams ep) {
    using C = Cfg<BN>;
    constexpr int WGS = Shape<BN, KIND>::WGS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
    // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (ep.M + BM - 1) / BM, tiles_n = (ep.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int kb_total = ph.kb_end[ph.n - 1];

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < ph.n; p++) { prefetch_tmap(&ph.A[p]); prefetch_tmap(&ph.B[p]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * WGS); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
    // prefetch) overlapped the previous kernel's tail; global memory is only touched below.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer: the whole warp runs the loop (uniform control flow), the instructions are predicated on one elected
        // lane (see elect_one, *_if); the next stage's `empty` barrier is tested before this stage's loads are issued
        {
            const uint32_t issue = elect_one() ? 1u : 0u;
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
                uint32_t eready = mbar_test(empty_bar(stage), phase ^ 1);
                if (KIND != K_GRAD_LONG) {   // constant tensor-map indices on the production paths (see gibbs_gemm_kernel_2cta)
                    const int kb1 = ph.kb_end[0];
                    for (int kb = 0; kb < kb_total; kb++) {
                        if (!eready) mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES, fb = full_bar(stage);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                        eready = (kb + 1 < kb_total) ? mbar_test(empty_bar(stage), phase ^ 1) : 0u;
                        mbar_expect_tx_if(issue, fb, C::STAGE_BYTES);
                        if (kb < kb1) {
                            tma_load_2d_if(issue, sa, &ph.A[0], fb, kb * BK, m_blk * BM);
                            tma_load_2d_if(issue, sb, &ph.B[0], fb, kb * BK, n_blk * BN);
                        } else {
                            tma_load_2d_if(issue, sa, &ph.A[1], fb, (kb - kb1) * BK, m_blk * BM);
                            tma_load_2d_if(issue, sb, &ph.B[1], fb, (kb - kb1) * BK, n_blk * BN);
                        }
                    }
                } else {
                    int p = 0, kb0 = 0;
                    for (int kb = 0; kb < kb_total; kb++) {
                        while (kb >= ph.kb_end[p]) kb0 = ph.kb_end[p++];
                        if (!eready) mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES, fb = full_bar(stage);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                        eready = (kb + 1 < kb_total) ? mbar_test(empty_bar(stage), phase ^ 1) : 0u;
                        mbar_expect_tx_if(issue, fb, C::STAGE_BYTES);
                        tma_load_2d_if(issue, sa, &ph.A[p], fb, (kb - kb0) * BK, m_blk * BM);
                        tma_load_2d_if(issue, sb, &ph.B[p], fb, (kb - kb0) * BK, n_blk * BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues; the next k-step's barrier is tested while this one's
        // MMAs are issued (see mbar_test)
        {
            const uint32_t issue = elect_one() ? 1u : 0u;
            constexpr uint32_t idesc_pos = make_idesc(BN, false), idesc_neg = make_idesc(BN, true);
            int stage = 0; uint32_t phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, iter++) {
                const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                int p = 0;
                constexpr bool short_list = KIND != K_GRAD_LONG;
                const int kb1m = ph.kb_end[0];
                const uint32_t idesc0 = (ph.neg_mask & 1u) ? idesc_neg : idesc_pos, idesc1 = (ph.neg_mask & 2u) ? idesc_neg : idesc_pos;
                uint32_t ready = mbar_test(full_bar(stage), phase);
                for (int kb = 0; kb < kb_total; kb++) {
                    if (!short_list) while (kb >= ph.kb_end[p]) p++;
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(sa), bdesc = make_kmajor_sw128_desc(sb);
                    const uint32_t idesc = short_list ? ((kb < kb1m) ? idesc0 : idesc1) : (((ph.neg_mask >> p) & 1u) ? idesc_neg : idesc_pos);
                    const int cur = stage;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    ready = (kb + 1 < kb_total) ? mbar_test(full_bar(stage), phase) : 0u;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        umma_bf16_if(issue, tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit_if(issue, empty_bar(cur));  // frees the smem slot once these MMAs retire
                    umma_commit_if(kb == kb_total - 1 ? issue : 0u, tfull_bar(as));
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue: 2 warpgroups x 4 warps; warp (w % 4) owns TMEM lanes [32*(w%4), +32)
        const int quarter = warp & 3, wg = (warp - EPI_WARP0) >> 2;
        int iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, iter++) {
            const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            if (KIND == K_GENERIC) epilogue_sample<BN, WGS>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_GRAD || KIND == K_GRAD_LONG) epilogue_grad<BN, WGS>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SAMPLE) epilogue_fast<BN, WGS, true, false>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SAMPLE_PROB) epilogue_fast<BN, WGS, true, true>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_GAUSS) epilogue_fast<BN, WGS, true, false, 1>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_LABEL) epilogue_fast<BN, WGS, true, false, 2>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SQERR) epilogue_fast<BN, WGS, false, true, 3>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else epilogue_fast<BN, WGS, false, true>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
        }
    }
    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// ----------------------------------------------------------------------------- 2-CTA (cta_group::2) variant
// A CTA pair (cluster of 2 on one TPC) computes one 256 x BN tile: each CTA stages its own 128 rows
// of A and HALF of the B tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads
// both CTAs' shared memory and writes each CTA's 128 accumulator rows into its own TMEM.  Per CTA
// and k-block this loads 16 KB + BN/2 x 128 B instead of 16 KB + BN x 128 B: a third less L2->SM
// traffic at BN = 256 and 6 pipeline stages instead of 4.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    // executed by both CTAs; the peer bit of the barrier address is cleared so the bytes are
    // accounted on the leader CTA's barrier (cute::SM100_TMA_2SM_LOAD_2D)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on the same barrier offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n, bool negate_a) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((negate_a ? 1u : 0u) << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN>
struct Cfg2 {
    static constexpr int A_BYTES = BM * BK * 2;            // this CTA's 128 rows of A
    static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the B tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_AB_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_AB_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// Tile number -> (pair row block, column block) of the persistent schedule.  Half-steps: unit tiles fastest (a wave shares few
// row blocks of states and every weight panel once).  Gradient GEMM (K = the whole batch, twice): a wave of the plain order
// touches all 16 A panels plus 5 B panels of 4-8 MiB each -- more than L2 holds, so every wave re-reads every A panel from HBM
// (ncu: 696 MB read for 268 MB of operands at cfg5).  Groups of 8 row blocks x all column blocks halve the A panels a wave
// keeps live at the price of reading B once per group.
template <int KIND>
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int &m_blk, int &n_blk) {
    if ((KIND == K_GRAD || KIND == K_GRAD_LONG) && tiles_m > 8) {
        constexpr int GM = 8;
        const int per_group = GM * tiles_n;
        const int g = tile / per_group, within = tile - g * per_group;
        const int gm = (tiles_m - g * GM) < GM ? (tiles_m - g * GM) : GM;
        m_blk = g * GM + within % gm;
        n_blk = within / gm;
    } else {
        m_blk = tile % tiles_m;
        n_blk = tile / tiles_m;
    }
}

template <int BN, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Shape<BN, KIND>::THREADS), 1)
gibbs_gemm_kernel_2cta(const __grid_constant__ Phases ph, const EpiParams ep) {
    using C = Cfg2<BN>;
    constexpr int WGS = Shape<BN, KIND>::WGS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int tiles_m = (ep.M + 2 * BM - 1) / (2 * BM), tiles_n = (ep.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int kb_total = ph.kb_end[ph.n - 1];

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < ph.n; p++) { prefetch_tmap(&ph.A[p]); prefetch_tmap(&ph.B[p]); }
    }
    if (warp == 1 && lane == 0) {
        // full: leader's arrive.expect_tx + the peer producer's remote arrive; tmem_empty: 8 epilogue warps per CTA
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * WGS); }
        fence_barrier_init();
    }
    cluster_sync_all();  // both CTAs are resident before the pair-wide TMEM allocation
    if (warp == 2) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();  // barriers initialised and TMEM allocated in both CTAs
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own A rows, own half of B; bytes land on the leader's barrier); warp-uniform loop, one
        // elected lane issues
        {
            const bool issuer = elect_one();
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int m_blk, n_blk;
                tile_coords<KIND>(tile, tiles_m, tiles_n, m_blk, n_blk);
                const int m0 = m_blk * 2 * BM + (int)rank * BM, n0 = n_blk * BN + (int)rank * (BN / 2);
                // One or two products (every production path): the tensor maps are addressed with CONSTANT indices.  A run-time
                // index into the kernel parameter makes the compiler keep a local copy of the whole product list, and TMA then
                // fetches its descriptors from local memory -- measured 6 % slower per GEMM.  Longer lists (QB_PREC_FP32X3's
                // component products) take the general loop.
                if (KIND != K_GRAD_LONG) {   // (lists longer than two products run the K_GRAD_LONG instantiation)
                    const int kb1 = ph.kb_end[0];
                    for (int kb = 0; kb < kb_total; kb++) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                        if (issuer) {
                            if (leader) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                            else mbar_arrive_remote(full_bar(stage), 0);
                            if (kb < kb1) {
                                tma_load_2d_2sm(sa, &ph.A[0], full_bar(stage), kb * BK, m0);
                                tma_load_2d_2sm(sb, &ph.B[0], full_bar(stage), kb * BK, n0);
                            } else {
                                tma_load_2d_2sm(sa, &ph.A[1], full_bar(stage), (kb - kb1) * BK, m0);
                                tma_load_2d_2sm(sb, &ph.B[1], full_bar(stage), (kb - kb1) * BK, n0);
                            }
                        }
                        __syncwarp();
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                } else {
                    int p = 0, kb0 = 0;
                    for (int kb = 0; kb < kb_total; kb++) {
                        while (kb >= ph.kb_end[p]) kb0 = ph.kb_end[p++];
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                        if (issuer) {
                            if (leader) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                            else mbar_arrive_remote(full_bar(stage), 0);
                            tma_load_2d_2sm(sa, &ph.A[p], full_bar(stage), (kb - kb0) * BK, m0);
                            tma_load_2d_2sm(sb, &ph.B[p], full_bar(stage), (kb - kb0) * BK, n0);
                        }
                        __syncwarp();
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected lane of the leader CTA's (warp-uniform) loop drives both tensor cores
        if (leader) {
            const bool issuer = elect_one();
            constexpr uint32_t idesc_pos = make_idesc_m256(BN, false), idesc_neg = make_idesc_m256(BN, true);
            int stage = 0; uint32_t phase = 0;
            int iter = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, iter++) {
                const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                int p = 0;
                constexpr bool short_list = KIND != K_GRAD_LONG;
                const int kb1m = ph.kb_end[0];
                const uint32_t idesc0 = (ph.neg_mask & 1u) ? idesc_neg : idesc_pos, idesc1 = (ph.neg_mask & 2u) ? idesc_neg : idesc_pos;
                uint32_t ready = mbar_test(full_bar(stage), phase);
                for (int kb = 0; kb < kb_total; kb++) {
                    if (!short_list) while (kb >= ph.kb_end[p]) p++;
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(sa), bdesc = make_kmajor_sw128_desc(sb);
                    const uint32_t idesc = short_list ? ((kb < kb1m) ? idesc0 : idesc1) : (((ph.neg_mask >> p) & 1u) ? idesc_neg : idesc_pos);
                    const int cur = stage;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    ready = (kb + 1 < kb_total) ? mbar_test(full_bar(stage), phase) : 0u;
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++)
                            umma_bf16_2sm(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        umma_commit_2sm(empty_bar(cur));
                        if (kb == kb_total - 1) umma_commit_2sm(tfull_bar(as));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue: each CTA drains its own 128 accumulator rows
        const int quarter = warp & 3, wg = (warp - EPI_WARP0) >> 2;
        int iter = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, iter++) {
            int m_pair, n_blk;
            tile_coords<KIND>(tile, tiles_m, tiles_n, m_pair, n_blk);
            const int m_blk = m_pair * 2 + (int)rank;  // in units of 128 rows
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            if (KIND == K_GENERIC) epilogue_sample<BN, WGS>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_GRAD || KIND == K_GRAD_LONG) epilogue_grad<BN, WGS>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SAMPLE) epilogue_fast<BN, WGS, true, false>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SAMPLE_PROB) epilogue_fast<BN, WGS, true, true>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_GAUSS) epilogue_fast<BN, WGS, true, false, 1>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_LABEL) epilogue_fast<BN, WGS, true, false, 2>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else if (KIND == K_FAST_SQERR) epilogue_fast<BN, WGS, false, true, 3>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            else epilogue_fast<BN, WGS, false, true>(ep, tmem_acc, m_blk, n_blk, quarter, wg, lane);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(tempty_bar(as));
                else mbar_arrive_remote(tempty_bar(as), 0);
            }
        }
    }
    __syncwarp();  // single-lane roles rejoin their warp before the .aligned cluster barrier
    tcgen05_fence_before();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal it
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------- host side
struct Operand {
    const bf16 *ptr;
    long long rows, k, ld;  // [rows][k] with row stride ld (elements); ld % 8 == 0, ptr 16 B aligned
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        QB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        QB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, -2, "cuTensorMapEncodeTiled not available from the driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

inline CUtensorMap encode_tmap(const Operand &o, int box_rows);

// Descriptors are pure functions of (pointer, extents, stride, box): cache them, the training loop
// re-launches the same few dozen operand views every step.
inline CUtensorMap make_tmap(const Operand &o, int box_rows) {
    struct Key {
        const void *p; long long rows, k, ld; int box;
        bool operator==(const Key &b) const { return p == b.p && rows == b.rows && k == b.k && ld == b.ld && box == b.box; }
    };
    struct Slot { Key key; CUtensorMap map; bool used = false; };
    static thread_local Slot cache[256];
    const Key key{o.ptr, o.rows, o.k, o.ld, box_rows};
    size_t hsh = ((uintptr_t)o.ptr >> 4) * 0x9E3779B97F4A7C15ull + (size_t)o.rows * 1315423911u + (size_t)o.k * 2654435761u + (size_t)box_rows;
    Slot &sl = cache[(hsh >> 20) & 255];
    if (sl.used && sl.key == key) return sl.map;
    sl.map = encode_tmap(o, box_rows);
    sl.key = key; sl.used = true;
    return sl.map;
}

inline CUtensorMap encode_tmap(const Operand &o, int box_rows) {
    QB_CHECK(o.ld % 8 == 0 && ((uintptr_t)o.ptr % 16) == 0, -1, "TMA operand must be 16-byte aligned with ld % 8 == 0");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)o.k, (cuuint64_t)o.rows};
    cuuint64_t strides[1] = {(cuuint64_t)o.ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)o.ptr, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QB_CHECK(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// which specialisation serves these epilogue parameters
inline int pick_kind(const EpiParams &ep) {
    if (ep.mode == EPI_GRAD) return K_GRAD;
    if (ep.sqerr && !ep.inj && !ep.softplus_rows && !ep.prob_f32 && ep.act == ACT_SIGMOID && ep.split >= ep.M && !ep.sample && !ep.state_rm &&
        !ep.state_T && !ep.prob_rm && !ep.prob_T && !ep.bg_what && ep.pscale == 1.f)
        return K_FAST_SQERR;
    const bool lean = !ep.inj && !ep.sqerr && !ep.softplus_rows && !ep.prob_f32;
    const bool plain = lean && ep.act == ACT_SIGMOID && ep.split >= ep.M;
    if (plain && ep.sample && ep.use_rng && ep.state_rm && !ep.prob_rm)
        return (ep.prob_T || ep.bg_what == 1) ? K_FAST_SAMPLE_PROB : K_FAST_SAMPLE;
    if (plain && !ep.sample && !ep.state_rm && !ep.state_T) return K_FAST_PROB;
    const bool sample_only = lean && ep.sample && ep.use_rng && ep.state_rm && !ep.prob_rm && !ep.prob_T && ep.bg_what != 1;
    if (sample_only && ep.act == ACT_GAUSSIAN && ep.split >= ep.M) return K_FAST_GAUSS;
    // label units fused when they all sit in one 32-lane group of a tile
    if (sample_only && ep.act == ACT_SIGMOID && ep.split < ep.M && (ep.split % 32) + (ep.M - ep.split) <= 32) return K_FAST_LABEL;
    return K_GENERIC;
}

// one product of a launch: D (+/-)= A B^T
struct Product {
    Operand A, B;
    bool negate;
};

template <int BN, bool PAIR>
inline void fill_phases(Phases &ph, const Product *prods, int n) {
    QB_CHECK(n >= 1 && n <= MAX_PHASES, -1, "a launch takes 1.." + std::to_string(MAX_PHASES) + " products");
    ph.n = n; ph.neg_mask = 0u;
    int kb = 0;
    for (int p = 0; p < n; p++) {
        QB_CHECK(prods[p].A.k == prods[p].B.k && prods[p].A.k > 0, -1, "GEMM operands disagree on K");
        ph.A[p] = make_tmap(prods[p].A, BM);
        ph.B[p] = make_tmap(prods[p].B, PAIR ? BN / 2 : BN);
        kb += (int)ceil_div(prods[p].A.k, BK);
        ph.kb_end[p] = kb;
        if (prods[p].negate) ph.neg_mask |= 1u << p;
    }
    for (int p = n; p < MAX_PHASES; p++) ph.kb_end[p] = kb;
}

template <int BN, int KIND>
inline void launch_bn_kind(cudaStream_t st, const Phases &ph, const EpiParams &ep, int grid) {
    using C = Cfg<BN>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Shape<BN, KIND>::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QB_CUDA(cudaLaunchKernelEx(&cfg, gibbs_gemm_kernel<BN, KIND>, ph, ep));
}

template <int BN, int KIND>
inline void launch_bn_kind_2cta(cudaStream_t st, const Phases &ph, const EpiParams &ep, int clusters) {
    using C = Cfg2<BN>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(Shape<BN, KIND>::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QB_CUDA(cudaLaunchKernelEx(&cfg, gibbs_gemm_kernel_2cta<BN, KIND>, ph, ep));
}

template <int BN>
inline void launch_bn_2cta(cudaStream_t st, const Product *prods, int n, const EpiParams &ep, int num_sms) {
    static thread_local Phases ph;
    fill_phases<BN, true>(ph, prods, n);
    const int tiles = (int)(ceil_div(ep.M, 2 * BM) * ceil_div(ep.N, BN));
    const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
    switch (pick_kind(ep)) {
        case K_GRAD:
            if (n > 2) launch_bn_kind_2cta<BN, K_GRAD_LONG>(st, ph, ep, clusters);
            else launch_bn_kind_2cta<BN, K_GRAD>(st, ph, ep, clusters);
            break;
        case K_FAST_SAMPLE: launch_bn_kind_2cta<BN, K_FAST_SAMPLE>(st, ph, ep, clusters); break;
        case K_FAST_PROB: launch_bn_kind_2cta<BN, K_FAST_PROB>(st, ph, ep, clusters); break;
        case K_FAST_SAMPLE_PROB: launch_bn_kind_2cta<BN, K_FAST_SAMPLE_PROB>(st, ph, ep, clusters); break;
        case K_FAST_GAUSS: launch_bn_kind_2cta<BN, K_FAST_GAUSS>(st, ph, ep, clusters); break;
        case K_FAST_LABEL: launch_bn_kind_2cta<BN, K_FAST_LABEL>(st, ph, ep, clusters); break;
        case K_FAST_SQERR: launch_bn_kind_2cta<BN, K_FAST_SQERR>(st, ph, ep, clusters); break;
        default: launch_bn_kind_2cta<BN, K_GENERIC>(st, ph, ep, clusters); break;
    }
    QB_CUDA(cudaGetLastError());
}

template <int BN>
inline void launch_bn(cudaStream_t st, const Product *prods, int n, const EpiParams &ep, int num_sms) {
    static thread_local Phases ph;
    fill_phases<BN, false>(ph, prods, n);
    const int tiles = (int)(ceil_div(ep.M, BM) * ceil_div(ep.N, BN));
    const int grid = tiles < num_sms ? tiles : num_sms;
    switch (pick_kind(ep)) {
        case K_GRAD:
            if (n > 2) launch_bn_kind<BN, K_GRAD_LONG>(st, ph, ep, grid);
            else launch_bn_kind<BN, K_GRAD>(st, ph, ep, grid);
            break;
        case K_FAST_SAMPLE: launch_bn_kind<BN, K_FAST_SAMPLE>(st, ph, ep, grid); break;
        case K_FAST_PROB: launch_bn_kind<BN, K_FAST_PROB>(st, ph, ep, grid); break;
        case K_FAST_SAMPLE_PROB: launch_bn_kind<BN, K_FAST_SAMPLE_PROB>(st, ph, ep, grid); break;
        case K_FAST_GAUSS: launch_bn_kind<BN, K_FAST_GAUSS>(st, ph, ep, grid); break;
        case K_FAST_LABEL: launch_bn_kind<BN, K_FAST_LABEL>(st, ph, ep, grid); break;
        case K_FAST_SQERR: launch_bn_kind<BN, K_FAST_SQERR>(st, ph, ep, grid); break;
        default: launch_bn_kind<BN, K_GENERIC>(st, ph, ep, grid); break;
    }
    QB_CUDA(cudaGetLastError());
}

template <int BN>
inline void init_bn() {
    const int b = Cfg<BN>::SMEM_BYTES;
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_GRAD_LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_SAMPLE_PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_LABEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel<BN, K_FAST_SQERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
}
template <int BN>
inline void init_bn_2cta() {
    const int b = Cfg2<BN>::SMEM_BYTES;
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_GRAD_LONG>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_SAMPLE_PROB>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_LABEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
    QB_CUDA(cudaFuncSetAttribute(gibbs_gemm_kernel_2cta<BN, K_FAST_SQERR>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
}
// once per device (qb_create): opt in to the large dynamic shared-memory carve-out
inline void init_device() {
    init_bn<256>(); init_bn<128>(); init_bn<64>(); init_bn<32>();
    init_bn_2cta<256>(); init_bn_2cta<128>();
}

// D[M][N]: picks the widest N tile that still yields at least one tile per SM.
// Tile shape selection: estimated time = waves x relative tile cost.  A 128 x BN tile on one SM and a
// 256 x BN tile on a CTA pair take the same time; narrower tiles cost more than proportionally
// (lower arithmetic intensity against L2).  Ties prefer the pair (a third less L2->SM traffic).
// use_2cta: 0 = never, 1 = automatic, 2 = always when BN >= 128 (tests)
inline void pick_shape(long long M, long long N, int num_sms, int force_bn, int use_2cta, int &bn, bool &pair) {
    const int bns[4] = {256, 128, 64, 32};
    const double cost[4] = {1.0, 0.85, 0.75, 0.70};  // measured: at K = 4096 a 256 x 128 pair tile takes ~as long as 256 x 256 (L2-bound)
    double best = 1e30;
    bn = 256; pair = false;
    for (int two = 1; two >= 0; two--) {
        if (two && use_2cta == 0) continue;
        for (int i = 0; i < 4; i++) {
            if (force_bn && bns[i] != force_bn) continue;
            if (two && bns[i] < 128) continue;
            if (!two && use_2cta == 2 && bns[i] >= 128) continue;
            const long long tiles = ceil_div(M, two ? 2 * BM : BM) * ceil_div(N, bns[i]);
            const long long slots = two ? num_sms / 2 : num_sms;
            const double est = (double)ceil_div(tiles, slots) * cost[i] * (two ? 0.97 : 1.0);
            if (est < best - 1e-9) { best = est; bn = bns[i]; pair = two != 0; }
        }
    }
}

// which (tile width, pair, specialisation) the last launch() of this thread picked: introspection for tests / bench
struct LaunchShape { int bn = 0, pair = 0, kind = -1; };
inline LaunchShape &last_shape() { static thread_local LaunchShape s; return s; }

// a list of products accumulated into one tile (see Phases)
inline void launch_products(cudaStream_t st, const Product *prods, int n, const EpiParams &ep, int num_sms, int force_bn = 0,
                            int use_2cta = 1) {
    int bn; bool pair;
    pick_shape(ep.M, ep.N, num_sms, force_bn, use_2cta, bn, pair);
    last_shape().bn = bn; last_shape().pair = pair ? 1 : 0; last_shape().kind = (pick_kind(ep) == K_GRAD && n > 2) ? K_GRAD_LONG : pick_kind(ep);
    if (pair) {
        if (bn == 256) launch_bn_2cta<256>(st, prods, n, ep, num_sms);
        else launch_bn_2cta<128>(st, prods, n, ep, num_sms);
        return;
    }
    switch (bn) {
        case 256: launch_bn<256>(st, prods, n, ep, num_sms); break;
        case 128: launch_bn<128>(st, prods, n, ep, num_sms); break;
        case 64: launch_bn<64>(st, prods, n, ep, num_sms); break;
        default: launch_bn<32>(st, prods, n, ep, num_sms); break;
    }
}

inline void launch(cudaStream_t st, const Operand &A, const Operand &B, const Operand *A2, const Operand *B2, const EpiParams &ep,
                   int num_sms, int force_bn = 0, int use_2cta = 1) {
    QB_CHECK((A2 == nullptr) == (B2 == nullptr), -1, "phase-2 operands must come in pairs");
    Product prods[2] = {{A, B, false}, {A, B, true}};
    if (A2) { prods[1].A = *A2; prods[1].B = *B2; }
    launch_products(st, prods, A2 ? 2 : 1, ep, num_sms, force_bn, use_2cta);
}

}  // namespace sm100
}  // namespace qb
