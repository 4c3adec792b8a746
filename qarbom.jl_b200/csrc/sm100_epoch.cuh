// Persistent multi-mini-batch training kernel for small models on sm_100a (QB_PREC_BF16 path).
//
// For the 784 x 500-sized models of BASELINE.json's configs 1, 2 and 4 a mini-batch step is a chain of 4-5 DEPENDENT GEMMs of
// a fraction of a microsecond each: launching them one by one costs 5-6 us per launch (host enqueue + launch latency + TMEM /
// barrier prologue + TMA ramp), twenty times the arithmetic.  This kernel runs MANY consecutive mini-batch steps of
// contrastive_divergence! / persistent_contrastive_divergence! (src/training/train_cd.jl:2-21, train_pcd.jl:2-43) in ONE launch:
//
//   per iteration (= mini-batch):  forward half-steps (the fused sampling GEMMs of sm100_gemm.cuh, 128 x BN single-CTA tiles)
//                                  -> gradient GEMM  G = P_h' X - H~' V~  whose epilogue IS update_rbm! (src/rbm.jl:152-163):
//                                     W += lr/B G on the fp32 master, both bf16 shadows rewritten, biases updated and the
//                                     bias-gradient accumulators left at zero
//
// All tiles of all iterations form one static schedule over persistent CTAs (one per SM).  Ordering is by arrival counters in
// global memory, as in sm100_chain.cuh: a half-step tile waits for the row block it contracts over, the gradient tiles wait for
// whole steps (every row of the batch contributes), and the first tiles of iteration i+1 wait for the update of iteration i.
// Data written through the generic proxy (states, shadows) and read by TMA is fenced with fence.proxy.async on both sides;
// data read by ordinary loads after another SM rewrote it (biases, the fp32 master) is loaded with ld.global.cg (L2).
//
// The mini-batch of iteration i is resident batch (batch0 + i) mod n_batches: its rows are a row offset into ONE tensor map
// of the whole dataset (and of its per-mini-batch transposed copies), the Philox draw ids advance by a fixed amount per
// iteration -- so the results are bit-identical to running the same steps one launch at a time.
#pragma once
#include "sm100_chain.cuh"

namespace qb {
namespace sm100 {

constexpr int EPOCH_MAX_STEPS = 26;   // half-steps + gradient per iteration (CD-k / PCD-k with k <= 12)
constexpr int EPOCH_MAX_MAPS = 12;
constexpr int K_GRAD_UPDATE = 8;

struct EpochStep {
    int kind;                      // K_FAST_* or K_GRAD_UPDATE
    int M, N;                      // extents of D: (units, batch rows) for half-steps, (hidden units, visible+label columns) for the gradient
    int tiles_m, tiles_n, tile_base;
    int n_prod;                    // products accumulated into the tile (2 for the gradient, the second negated)
    int a_map[2], b_map[2], K[2];
    int a_row_mb[2], b_row_mb[2];  // row-coordinate increment per mini-batch index (operands that are views of the dataset)
    int dep_rb, dep_rb_count;      // step of this iteration whose same row block must be complete (-1: none)
    int dep_all[2], dep_all_count[2];   // steps of this iteration that must be complete in full (-1: none)
    int dep_prev;                  // wait for the last step (the update) of the previous iteration
    int draw_off;                  // Philox draw id = draw0 + iteration * draws_per_iter + draw_off
    EpiParams ep;
};
struct EpochProgram {
    int n_steps, n_iters, tiles_per_iter, draws_per_iter, max_tiles_n, n_maps;
    int last_step_total;           // arrivals that complete the last step (the update) of an iteration
    unsigned long long draw0;
    int batch0, n_batches;
    unsigned int *rb_cnt;          // [n_iters][n_steps][max_tiles_n]
    unsigned int *tot_cnt;         // [n_iters][n_steps]
    unsigned int *error_flag, *abort_flag;
    // update_rbm! fused into the gradient epilogue
    float *W; bf16 *Wb, *WbT; long long ldv, ldh; float lr, llr; int V;
    float *bias_P, *bias_G; const float *datasum; long long datasum_ld; float dscale; int n_vis, n_bias;
#ifdef QB_EPOCH_TRACE
    unsigned long long *trace;     // [tile][32] %globaltimer stamps (tools/epoch_trace.py); development builds only
#endif
    EpochStep step[EPOCH_MAX_STEPS];
};
#ifdef QB_EPOCH_TRACE
__device__ __forceinline__ void epoch_stamp(const EpochProgram &P, long long tile, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (P.trace && (threadIdx.x & 31) == 0) P.trace[tile * 32 + slot] = t;
}
#define QB_STAMP(tile, slot) epoch_stamp(prog, tile, slot)
#else
#define QB_STAMP(tile, slot) ((void)0)
#endif
struct EpochMaps {
    CUtensorMap m[EPOCH_MAX_MAPS];
};

__device__ __forceinline__ bool epoch_wait(const unsigned int *cnt, unsigned int need, const EpochProgram &prog) {
    if (ld_acquire_gpu(cnt) >= need) return true;
    const long long t0 = clock64();
    unsigned int spins = 0;
    while (ld_acquire_gpu(cnt) < need) {
        __nanosleep(32);
        if ((++spins & 255u) == 0u) {
            if (*(volatile unsigned int *)prog.abort_flag) return false;
            if (clock64() - t0 > 1000000000ll) {
                *(volatile unsigned int *)prog.abort_flag = 1u;
                *(volatile unsigned int *)prog.error_flag = 1u;
                return false;
            }
        }
    }
    return true;
}

// gradient tile -> update_rbm!: W[h][c] += lr_c acc; bf16 shadows; biases by the first row / column of tiles
template <int BN, int WGS>
__device__ __forceinline__ void epilogue_grad_update(const EpochProgram &P, const EpochStep &S, uint32_t tmem_acc, int m_blk, int n_blk,
                                                     int quarter, int wg, int lane, int mb) {
    const int hrow = m_blk * BM + quarter * 32 + lane;
    constexpr int COLS_PER_WG = BN / WGS;
#pragma unroll 1
    for (int c0 = wg * COLS_PER_WG; c0 < (wg + 1) * COLS_PER_WG; c0 += 16) {
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld(r);
        const int n0 = n_blk * BN + c0;
        if (n0 >= S.N || hrow >= S.M) continue;
        float *wrow = P.W + (long long)hrow * P.ldv + n0;
        unsigned short *brow = reinterpret_cast<unsigned short *>(P.Wb) + (long long)hrow * P.ldv + n0;
        unsigned short *tcol = reinterpret_cast<unsigned short *>(P.WbT) + (long long)n0 * P.ldh + hrow;
        uint32_t wb[16];
        if (n0 + 16 <= S.N) {
            float w[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(wrow) + q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
                w[j] += ((n0 + j < P.V) ? P.lr : P.llr) * __uint_as_float(r[j]);
                wb[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(w[j]));
            }
#pragma unroll
            for (int q = 0; q < 4; q++) reinterpret_cast<float4 *>(wrow)[q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            reinterpret_cast<uint4 *>(brow)[0] = make_uint4(wb[0] | (wb[1] << 16), wb[2] | (wb[3] << 16), wb[4] | (wb[5] << 16), wb[6] | (wb[7] << 16));
            reinterpret_cast<uint4 *>(brow)[1] = make_uint4(wb[8] | (wb[9] << 16), wb[10] | (wb[11] << 16), wb[12] | (wb[13] << 16), wb[14] | (wb[15] << 16));
#pragma unroll
            for (int j = 0; j < 16; j++) { *tcol = (unsigned short)wb[j]; tcol += P.ldh; }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (n0 + j < S.N) {
                    const float w = __ldcg(wrow + j) + ((n0 + j < P.V) ? P.lr : P.llr) * __uint_as_float(r[j]);
                    const unsigned short b = __bfloat16_as_ushort(__float2bfloat16(w));
                    wrow[j] = w; brow[j] = b; tcol[(long long)j * P.ldh] = b;
                }
        }
    }
    // biases: b (hidden) by the tiles of the first column block, [a | c] (visible + label) by the tiles of the first row block
    if (n_blk == 0 && wg == 0 && hrow < S.M) {
        const int i = P.n_vis + hrow;
        const float g = __ldcg(P.bias_G + i);
        P.bias_P[i] = __ldcg(P.bias_P + i) + P.lr * g;
        P.bias_G[i] = 0.f;
    }
    if (m_blk == 0) {
        const int idx = (wg * 4 + quarter) * 32 + lane;
        const int c = n_blk * BN + idx;
        if (idx < BN && c < S.N) {
            float g = __ldcg(P.bias_G + c);
            if (P.datasum) g += P.dscale * P.datasum[(long long)mb * P.datasum_ld + c];
            P.bias_P[c] = __ldcg(P.bias_P + c) + ((c < P.V) ? P.lr : P.llr) * g;
            P.bias_G[c] = 0.f;
        }
    }
}

template <int BN>
__global__ void __launch_bounds__((Shape<BN, K_FAST_SAMPLE>::THREADS), 1)
gibbs_epoch_kernel(const __grid_constant__ EpochMaps maps, const __grid_constant__ EpochProgram prog) {
    using C = Cfg<BN>;
    constexpr int WGS = Shape<BN, K_FAST_SAMPLE>::WGS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_iter = prog.tiles_per_iter;
    const long long total_tiles = (long long)tiles_per_iter * prog.n_iters;

    if (warp == 0 && lane == 0)
        for (int i = 0; i < prog.n_maps; i++) prefetch_tmap(&maps.m[i]);
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
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // (iteration, step, local tile) of a global tile number
    auto locate = [&](long long tile, int &iter, int &si, int &local) {
        iter = (int)(tile / tiles_per_iter);
        local = (int)(tile - (long long)iter * tiles_per_iter);
        si = 0;
        while (si + 1 < prog.n_steps && local >= prog.step[si + 1].tile_base) si++;
        local -= prog.step[si].tile_base;
    };

    if (warp == 0) {
        // ===== TMA producer: the whole warp runs the loop, the instructions are predicated on one elected lane (elect_one, *_if)
        {
            const uint32_t issue = elect_one() ? 1u : 0u;
            int stage = 0; uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int iter, si, local;
                locate(tile, iter, si, local);
                const EpochStep &S = prog.step[si];
                const int m_blk = local % S.tiles_m, n_blk = local / S.tiles_m;
                const int mb = (prog.batch0 + iter) % prog.n_batches;
                const unsigned int *rb = prog.rb_cnt + ((long long)iter * prog.n_steps) * prog.max_tiles_n;
                const unsigned int *tot = prog.tot_cnt + (long long)iter * prog.n_steps;
                // whole-step dependencies gate BOTH operands (new weights / operands produced by every row of the batch)
                // EVERY tile of iteration iter waits for the update of iteration iter - 1 before its first load: the weights of
                // a row-block-dependent step are fetched ahead of its row-block wait, and this CTA need not have run any of
                // the iteration's earlier steps (dep_prev steps need it by construction).  tot - 1: last step of iteration iter - 1
                bool waited = false;
                QB_STAMP(tile, 0);
                if (iter > 0) { epoch_wait(tot - 1, (unsigned int)prog.last_step_total, prog); waited = true; }
                for (int d = 0; d < 2; d++)
                    if (S.dep_all[d] >= 0) { epoch_wait(tot + S.dep_all[d], (unsigned int)S.dep_all_count[d], prog); waited = true; }
                if (waited) fence_proxy_async_global();
                QB_STAMP(tile, 1);
                int kb_total = 0;
                for (int p = 0; p < S.n_prod; p++) kb_total += (S.K[p] + BK - 1) / BK;
                const int kb_first = (S.K[0] + BK - 1) / BK;
                int kb = 0;
                if (S.dep_rb >= 0) {
                    // weights of the first stages ahead of the row-block wait (single-product steps only)
                    const int pre = kb_total < C::STAGES ? kb_total : C::STAGES;
                    int st = stage; uint32_t ph = phase;
                    for (int i = 0; i < pre; i++) {
                        mbar_wait(empty_bar(st), ph ^ 1);
                        mbar_expect_tx_if(issue, full_bar(st), C::STAGE_BYTES);
                        tma_load_2d_if(issue, smem_base + st * C::STAGE_BYTES, &maps.m[S.a_map[0]], full_bar(st), i * BK, m_blk * BM + mb * S.a_row_mb[0]);
                        if (++st == C::STAGES) { st = 0; ph ^= 1; }
                    }
                    epoch_wait(rb + S.dep_rb * prog.max_tiles_n + n_blk, (unsigned int)S.dep_rb_count, prog);
                    fence_proxy_async_global();
                    QB_STAMP(tile, 2);
#ifdef QB_EPOCH_TRACE
                    const long long c_tma0 = clock64();
#endif
                    for (int i = 0; i < pre; i++) {
                        tma_load_2d_if(issue, smem_base + stage * C::STAGE_BYTES + C::A_BYTES, &maps.m[S.b_map[0]], full_bar(stage), i * BK,
                                       n_blk * BN + mb * S.b_row_mb[0]);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    }
#ifdef QB_EPOCH_TRACE
                    if (prog.trace && lane == 0) prog.trace[tile * 32 + 28] = (unsigned long long)(clock64() - c_tma0);
#endif
                    kb = pre;
                }
                uint32_t eready = kb < kb_total ? mbar_test(empty_bar(stage), phase ^ 1) : 0u;
                for (; kb < kb_total; kb++) {
                    const int p = kb < kb_first ? 0 : 1, kk = kb < kb_first ? kb : kb - kb_first;
                    if (!eready) mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES, fb = full_bar(stage);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    eready = (kb + 1 < kb_total) ? mbar_test(empty_bar(stage), phase ^ 1) : 0u;
                    mbar_expect_tx_if(issue, fb, C::STAGE_BYTES);
                    tma_load_2d_if(issue, sa, &maps.m[S.a_map[p]], fb, kk * BK, m_blk * BM + mb * S.a_row_mb[p]);
                    tma_load_2d_if(issue, sb, &maps.m[S.b_map[p]], fb, kk * BK, n_blk * BN + mb * S.b_row_mb[p]);
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
            long long it = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                int iter, si, local;
                locate(tile, iter, si, local);
                const EpochStep &S = prog.step[si];
                int kb_total = 0;
                for (int p = 0; p < S.n_prod; p++) kb_total += (S.K[p] + BK - 1) / BK;
                const int kb_first = (S.K[0] + BK - 1) / BK;
                const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                uint32_t ready = mbar_test(full_bar(stage), phase);
                for (int kb = 0; kb < kb_total; kb++) {
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    if (kb == 0) QB_STAMP(tile, 3);
                    if (kb == kb_total - 1) QB_STAMP(tile, 4);
#ifdef QB_EPOCH_TRACE
                    const long long c_issue0 = clock64();   // SM clock: k-step kb's operands are in shared memory
#endif
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(sa), bdesc = make_kmajor_sw128_desc(sb);
                    const uint32_t idesc = kb < kb_first ? idesc_pos : idesc_neg;
                    const int cur = stage;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    ready = (kb + 1 < kb_total) ? mbar_test(full_bar(stage), phase) : 0u;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++)
                        umma_bf16_if(issue, tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    umma_commit_if(issue, empty_bar(cur));
                    umma_commit_if(kb == kb_total - 1 ? issue : 0u, tfull_bar(as));
#ifdef QB_EPOCH_TRACE
                    if (prog.trace && lane == 0 && kb < 20)   // low words: (clock after the MMAs were issued, clock when the operands had landed)
                        prog.trace[tile * 32 + 8 + kb] = ((unsigned long long)(uint32_t)clock64() << 32) | (unsigned long long)(uint32_t)c_issue0;
#endif
                }
            }
        }
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4 * WGS) {
        // ===== epilogue
        const int quarter = warp & 3, wg = (warp - EPI_WARP0) >> 2;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            int iter, si, local;
            locate(tile, iter, si, local);
            const EpochStep &S = prog.step[si];
            const int m_blk = local % S.tiles_m, n_blk = local / S.tiles_m;
            const int mb = (prog.batch0 + iter) % prog.n_batches;
            const int as = (int)(it & 1); const uint32_t aphase = (uint32_t)((it >> 1) & 1);
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            if (warp == EPI_WARP0 && lane == 0) QB_STAMP(tile, 5);
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            if (S.kind == K_GRAD_UPDATE) {
                epilogue_grad_update<BN, WGS>(prog, S, tmem_acc, m_blk, n_blk, quarter, wg, lane, mb);
            } else {
                EpiParams e = S.ep;
                const unsigned long long draw = prog.draw0 + (unsigned long long)iter * (unsigned long long)prog.draws_per_iter + (unsigned long long)S.draw_off;
                e.key.draw_lo = (uint32_t)draw; e.key.draw_hi = (uint32_t)(draw >> 32) & 0x7FFFFFFFu;
                switch (S.kind) {   // warp-uniform
                    case K_FAST_SAMPLE: epilogue_fast<BN, WGS, true, false>(e, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                    case K_FAST_SAMPLE_PROB: epilogue_fast<BN, WGS, true, true>(e, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                    case K_FAST_GAUSS: epilogue_fast<BN, WGS, true, false, 1>(e, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                    case K_FAST_LABEL: epilogue_fast<BN, WGS, true, false, 2>(e, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                    default: epilogue_fast<BN, WGS, false, true>(e, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                }
            }
            if (warp == EPI_WARP0 && lane == 0) QB_STAMP(tile, 6);
            tcgen05_fence_before();
            fence_proxy_async_global();
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tempty_bar(as));
                // (every lane has run a gpu-scope fence above: the relaxed adds complete the release)
                red_relaxed_gpu_add(prog.rb_cnt + ((long long)iter * prog.n_steps + si) * prog.max_tiles_n + n_blk, 1u);
                red_relaxed_gpu_add(prog.tot_cnt + (long long)iter * prog.n_steps + si, 1u);
            }
            if (warp == EPI_WARP0 && lane == 0) QB_STAMP(tile, 7);
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

// ----------------------------------------------------------------------------- host side
struct EpochBuilder {
    EpochMaps maps;
    EpochProgram prog;
    Operand ops[EPOCH_MAX_MAPS];
    int op_box[EPOCH_MAX_MAPS];
    int bn = 64;
    double flops_per_iter = 0.0;
    EpochBuilder() { memset(&maps, 0, sizeof(maps)); memset(&prog, 0, sizeof(prog)); }

    void begin(int bn_) {
        bn = bn_;
        prog.n_steps = 0; prog.tiles_per_iter = 0; prog.n_maps = 0; prog.max_tiles_n = 1; flops_per_iter = 0.0;
    }
    int warps_per_tile() const { return 4 * (bn >= 64 ? 4 : 2); }
    int map_index(const Operand &o, int box_rows) {
        for (int i = 0; i < prog.n_maps; i++)
            if (ops[i].ptr == o.ptr && ops[i].rows == o.rows && ops[i].k == o.k && ops[i].ld == o.ld && op_box[i] == box_rows) return i;
        QB_CHECK(prog.n_maps < EPOCH_MAX_MAPS, -1, "epoch program uses too many distinct operands");
        ops[prog.n_maps] = o; op_box[prog.n_maps] = box_rows;
        maps.m[prog.n_maps] = make_tmap(o, box_rows);
        return prog.n_maps++;
    }
    EpochStep &new_step(int kind, int M, int N) {
        QB_CHECK(prog.n_steps < EPOCH_MAX_STEPS, -1, "epoch program too long");
        EpochStep &S = prog.step[prog.n_steps];
        memset(&S, 0, sizeof(S));
        S.kind = kind; S.M = M; S.N = N;
        S.tiles_m = (int)ceil_div(M, BM); S.tiles_n = (int)ceil_div(N, bn);
        S.tile_base = prog.tiles_per_iter;
        S.dep_rb = -1; S.dep_all[0] = S.dep_all[1] = -1;
        prog.tiles_per_iter += S.tiles_m * S.tiles_n;
        if (S.tiles_n > prog.max_tiles_n) prog.max_tiles_n = S.tiles_n;
        return S;
    }
    // a sampling / probability half-step; A = weights, B = states (b_row_mb: rows per mini-batch when B views the dataset);
    // dep: index of the step whose outputs it reads (-1: ready when the previous iteration's update is done)
    int add_half_step(const Operand &A, const Operand &B, int b_row_mb, const EpiParams &ep, int dep, int draw_off) {
        const int kind = pick_kind(ep);
        QB_CHECK(kind >= K_FAST_SAMPLE && kind <= K_FAST_LABEL, -5, "half-step has no compile-time epilogue");
        EpochStep &S = new_step(kind, ep.M, ep.N);
        S.n_prod = 1;
        S.a_map[0] = map_index(A, BM); S.b_map[0] = map_index(B, bn);
        S.K[0] = (int)A.k; S.b_row_mb[0] = b_row_mb;
        S.dep_rb = dep;
        if (dep >= 0) S.dep_rb_count = prog.step[dep].tiles_m * warps_per_tile();
        else S.dep_prev = 1;
        S.draw_off = draw_off;
        S.ep = ep;
        flops_per_iter += 2.0 * (double)ep.M * (double)ep.N * (double)A.k;
        return prog.n_steps++;
    }
    // G = A1 B1' - A2 B2' with the update fused; waits for steps d0, d1 (whole); must be the LAST step of the iteration
    int add_grad_update(const Operand &A1, const Operand &B1, int b1_row_mb, const Operand &A2, const Operand &B2, int M, int N, int d0, int d1) {
        EpochStep &S = new_step(K_GRAD_UPDATE, M, N);
        S.n_prod = 2;
        S.a_map[0] = map_index(A1, BM); S.b_map[0] = map_index(B1, bn); S.K[0] = (int)A1.k; S.b_row_mb[0] = b1_row_mb;
        S.a_map[1] = map_index(A2, BM); S.b_map[1] = map_index(B2, bn); S.K[1] = (int)A2.k;
        S.dep_all[0] = d0; S.dep_all[1] = d1;
        if (d0 >= 0) S.dep_all_count[0] = prog.step[d0].tiles_m * prog.step[d0].tiles_n * warps_per_tile();
        if (d1 >= 0) S.dep_all_count[1] = prog.step[d1].tiles_m * prog.step[d1].tiles_n * warps_per_tile();
        prog.last_step_total = S.tiles_m * S.tiles_n * warps_per_tile();
        flops_per_iter += 2.0 * (double)M * (double)N * (double)(A1.k + A2.k);
        return prog.n_steps++;
    }
};

template <int BN>
inline void launch_epoch_bn(cudaStream_t st, const EpochBuilder &eb, int num_sms) {
    using C = Cfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        QB_CUDA(cudaFuncSetAttribute(gibbs_epoch_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    int max_tiles = 1;
    for (int s = 0; s < eb.prog.n_steps; s++) max_tiles = std::max(max_tiles, eb.prog.step[s].tiles_m * eb.prog.step[s].tiles_n);
    const int grid = std::min(num_sms, max_tiles);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(Shape<BN, K_FAST_SAMPLE>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QB_CUDA(cudaLaunchKernelEx(&cfg, gibbs_epoch_kernel<BN>, eb.maps, eb.prog));
}

inline void launch_epoch(cudaStream_t st, const EpochBuilder &eb, int num_sms) {
    QB_CHECK(eb.prog.n_steps >= 2 && eb.prog.step[eb.prog.n_steps - 1].kind == K_GRAD_UPDATE, -1, "epoch program must end with the gradient/update step");
    if (eb.bn == 64) launch_epoch_bn<64>(st, eb, num_sms);
    else launch_epoch_bn<32>(st, eb, num_sms);
    QB_CUDA(cudaGetLastError());
}

}  // namespace sm100
}  // namespace qb
