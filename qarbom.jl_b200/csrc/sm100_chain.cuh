// Persistent multi-half-step Gibbs chain kernel for sm_100a (QB_PREC_BF16 path).
//
// One launch runs a whole PROGRAM of dependent half-steps -- e.g. the 2k half-steps of _update_fantasy_data!
// (src/fantasy_data.jl:12-19), or the 2k+1 forward contractions of a CD-k mini-batch (src/training/train_cd.jl:4-19) --
// instead of one launch per half-step.  Every half-step is the same fused GEMM + epilogue as gibbs_gemm_kernel_2cta
// (256 x BN tiles on CTA pairs, tcgen05.mma cta_group::2, TMA-fed 128B-swizzled smem ring, two TMEM accumulator stages,
// the compile-time epilogues of sm100_gemm.cuh), and the tiles of ALL half-steps form one static schedule.
//
// Dependencies are per block of BN batch rows, not per half-step: a tile of half-step t for rows [n BN, (n+1) BN) contracts
// over all units of those rows, i.e. it needs exactly the unit tiles (t-1, *, n).  Every epilogue warp that finishes a tile
// bumps the arrival counter of (step, row block) after a gpu-scope fence; the TMA producer of a dependent tile loads the
// WEIGHT operand (always ready) into the ring first, then polls the counter with an acquire load, issues a generic->async
// proxy fence and only then loads the STATE operand.  Row blocks therefore drift against each other: the epilogue of one
// row block hides under the MMAs of another, the tail of half-step t overlaps the head of half-step t+1 (no wave
// quantisation, no launch gaps, no TMEM/barrier prologue per half-step).
//
// Deadlock freedom: tiles are numbered (step, row block, unit tile) lexicographically, CTA pair c processes tiles
// c, c + P, c + 2P, ... in increasing order, every dependency of a tile has a smaller number, and all P <= 74 pairs are
// co-resident (one CTA per SM by shared-memory footprint).  The poll is bounded (about two seconds): on expiry the kernel
// raises a host-visible error flag and carries on instead of hanging the GPU.
#pragma once
#include "sm100_gemm.cuh"

namespace qb {
namespace sm100 {

constexpr int CHAIN_MAX_STEPS = 24;
constexpr int CHAIN_MAX_MAPS = 8;

struct ChainStep {
    int a_map, b_map;   // indices into ChainMaps::m (A: weights [units][K], box 64 x 128; B: states [rows][K], box 64 x BN/2)
    int M, K;           // units of this half-step (D rows) and contraction length
    int kind;           // K_FAST_SAMPLE / K_FAST_SAMPLE_PROB / K_FAST_PROB / K_FAST_GAUSS / K_FAST_LABEL
    int dep_step;       // -1: inputs are ready at launch; else the step whose outputs (same row block) this one contracts over
    int dep_count;      // arrivals that complete (dep_step, row block): 2 CTAs x epilogue warps x unit tiles of dep_step
    int tile_base;      // global number of this step's first tile
    int tiles_m;        // unit tiles (256 units each) of this step
    EpiParams ep;
};
struct ChainProgram {
    int n_steps, N, tiles_n, total_tiles, n_maps;
    unsigned int *counters;      // [n_steps][tiles_n], zero at launch
    unsigned int *error_flag;    // mapped host memory: set when a dependency poll expired
    unsigned int *abort_flag;    // device copy of the same: once set, no poll of this or a later launch waits again
    ChainStep step[CHAIN_MAX_STEPS];
};
struct ChainMaps {
    CUtensorMap m[CHAIN_MAX_MAPS];
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// after a gpu-scope fence by every contributing lane and a warp barrier, the relaxed add completes a release pattern
// (fence.acq_rel + relaxed atomic); red.release would run a second fence of ~0.7 us behind the first one
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned int *p, unsigned int v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Shape<BN, K_FAST_SAMPLE>::THREADS), 1)
gibbs_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainProgram prog) {
    using C = Cfg2<BN>;
#ifdef QB_EXP_BFRAC
    constexpr int EXP_TX = 2 * (C::A_BYTES + C::B_BYTES / QB_EXP_BFRAC);
#else
    constexpr int EXP_TX = 2 * C::STAGE_BYTES;
#endif
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
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int total_tiles = prog.total_tiles, tiles_n = prog.tiles_n;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < prog.n_maps; i++) prefetch_tmap(&maps.m[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * WGS); }
        fence_barrier_init();
    }
    cluster_sync_all();
    if (warp == 2) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // step of a global tile number (tiles are visited in increasing order: the cursor only moves forward)
    auto advance = [&](int tile, int &si) {
        while (si + 1 < prog.n_steps && tile >= prog.step[si + 1].tile_base) si++;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own 128 weight rows, own half of the state tile; bytes land on the leader's barrier);
        // warp-uniform loop, one elected lane issues (see elect_one)
        {
            const bool issuer = elect_one();
            int stage = 0; uint32_t phase = 0;
            int si = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                advance(tile, si);
                const ChainStep &S = prog.step[si];
                const int local = tile - S.tile_base;
                const int m_blk = local % S.tiles_m, n_blk = local / S.tiles_m;
                const int m0 = m_blk * 2 * BM + (int)rank * BM, n0 = n_blk * BN + (int)rank * (BN / 2);
                const int kb_total = (S.K + BK - 1) / BK;
                const CUtensorMap *mA = &maps.m[S.a_map], *mB = &maps.m[S.b_map];
                int kb = 0;
                if (S.dep_step >= 0) {
                    // weights first: they do not depend on the previous half-step
                    const int pre = kb_total < C::STAGES ? kb_total : C::STAGES;
                    int st = stage; uint32_t ph = phase;
                    for (int i = 0; i < pre; i++) {
                        mbar_wait(empty_bar(st), ph ^ 1);
                        if (issuer) {
                            if (leader) mbar_expect_tx(full_bar(st), EXP_TX);
                            else mbar_arrive_remote(full_bar(st), 0);
                            tma_load_2d_2sm(smem_base + st * C::STAGE_BYTES, mA, full_bar(st), i * BK, m0);
                        }
                        __syncwarp();
                        if (++st == C::STAGES) { st = 0; ph ^= 1; }
                    }
                    // ... then wait until every unit tile of (dep_step, this row block) has been stored
                    const unsigned int *cnt = prog.counters + S.dep_step * tiles_n + n_blk;
                    const unsigned int need = (unsigned int)S.dep_count;
                    if (ld_acquire_gpu(cnt) < need) {
                        const long long t0 = clock64();
                        unsigned int spins = 0;
                        while (ld_acquire_gpu(cnt) < need) {
                            __nanosleep(64);
                            if ((++spins & 255u) == 0u) {
                                if (*(volatile unsigned int *)prog.abort_flag) break;
                                if (clock64() - t0 > 1000000000ll) {   // ~0.5 s: a broken program must not wedge the GPU
                                    *(volatile unsigned int *)prog.abort_flag = 1u;
                                    *(volatile unsigned int *)prog.error_flag = 1u;
                                    break;
                                }
                            }
                        }
                    }
                    fence_proxy_async_global();   // the states were written through the generic proxy, TMA reads through the async one
                    if (issuer) {
                        int st2 = stage;
                        for (int i = 0; i < pre; i++) {
                            tma_load_2d_2sm(smem_base + st2 * C::STAGE_BYTES + C::A_BYTES, mB, full_bar(st2), i * BK, n0);
                            if (++st2 == C::STAGES) st2 = 0;
                        }
                    }
                    __syncwarp();
                    stage = st; phase = ph;
                    kb = pre;
                }
                for (; kb < kb_total; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    if (issuer) {
                        if (leader) mbar_expect_tx(full_bar(stage), EXP_TX);
                        else mbar_arrive_remote(full_bar(stage), 0);
                        tma_load_2d_2sm(sa, mA, full_bar(stage), kb * BK, m0);
                        tma_load_2d_2sm(sb, mB, full_bar(stage), kb * BK, n0);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected lane of the leader CTA's (warp-uniform) loop drives both tensor cores; the next k-step's
        // barrier is tested while this one's MMAs are issued (see mbar_test)
        if (leader) {
            const bool issuer = elect_one();
            constexpr uint32_t idesc = make_idesc_m256(BN, false);
            int stage = 0; uint32_t phase = 0;
            int iter = 0, si = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, iter++) {
                advance(tile, si);
                const int kb_total = (prog.step[si].K + BK - 1) / BK;
                const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                uint32_t ready = mbar_test(full_bar(stage), phase);
                for (int kb = 0; kb < kb_total; kb++) {
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(sa), bdesc = make_kmajor_sw128_desc(sb);
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
        // ===== epilogue: each CTA drains its own 128 accumulator rows, then publishes the tile
        const int quarter = warp & 3, wg = (warp - EPI_WARP0) >> 2;
        int iter = 0, si = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, iter++) {
            advance(tile, si);
            const ChainStep &S = prog.step[si];
            const int local = tile - S.tile_base;
            const int m_blk = (local % S.tiles_m) * 2 + (int)rank, n_blk = local / S.tiles_m;
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            switch (S.kind) {   // warp-uniform
                case K_FAST_SAMPLE: epilogue_fast<BN, WGS, true, false>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_SAMPLE_PROB: epilogue_fast<BN, WGS, true, true>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_GAUSS: epilogue_fast<BN, WGS, true, false, 1>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_LABEL: epilogue_fast<BN, WGS, true, false, 2>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                default: epilogue_fast<BN, WGS, false, true>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
            }
            tcgen05_fence_before();
            // publish: every lane's stores are ordered before the gpu-scope release of lane 0 (fence by all lanes, warp barrier)
            fence_proxy_async_global();
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(tempty_bar(as));
                else mbar_arrive_remote(tempty_bar(as), 0);
                red_relaxed_gpu_add(prog.counters + si * tiles_n + n_blk, 1u);
            }
        }
    }
    __syncwarp();
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------- quad variant: 4-CTA clusters, weights multicast
// When a half-step is about ONE wave of 256 x 256 pair tiles (e.g. cfg5 on 8 GPUs: 1024 rows per rank = 64 tiles on 74 pairs),
// every pair holds one tile per half-step, all tiles of a row block finish together and the epilogue of the only tile is fully
// exposed.  Halving the tile width gives every pair two independent tiles per half-step -- one's epilogue hides under the
// other's MMAs -- at the price of streaming the weights twice.  Here two CTA pairs form a cluster of FOUR CTAs that works on
// one 256-unit x 256-row tile: pair q takes the 128-row block 2R + q, and every 128 x 64 weight slab is loaded ONCE per cluster
// -- each CTA fetches 64 of its 128 weight rows and TMA-multicasts them to the CTA of the other pair that needs the same rows.
// L2 traffic is that of the 256 x 256 pair tile, dependencies and epilogues are per 128-row block.  Measured level with the
// plain 256 x 128 pair chain (the weights' second pass comes out of L2 at no visible cost), so it is an option, not the default.
//   barriers: full (per pair leader) = leader's arrive.expect_tx + peer's arrive + 48 KB of bytes (half of the A bytes come
//   from the other pair's multicasts); empty (every CTA) = 2 arrivals, the tcgen05.commit of BOTH pair leaders multicast to all
//   four CTAs -- a stage may be refilled only when both pairs have consumed it, because every load writes into two CTAs.
constexpr int QUAD_BN = 128;   // rows per pair

__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mask(uint32_t bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__((Shape<QUAD_BN, K_FAST_SAMPLE>::THREADS), 1)
gibbs_chain_kernel_quad(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainProgram prog) {
    constexpr int BN = QUAD_BN;
    using C = Cfg2<BN>;
    constexpr int WGS = Shape<BN, K_FAST_SAMPLE>::WGS;
    constexpr int A_HALF_BYTES = C::A_BYTES / 2;   // the 64 weight rows this CTA fetches for itself and its twin
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
    const uint32_t q = rank >> 1, pr = rank & 1u, lead_rank = rank & ~1u;   // pair, CTA within the pair, rank of the pair's leader
    const bool leader = pr == 0;
    const unsigned short twin_mask = (unsigned short)((1u << rank) | (1u << (rank ^ 2u)));
    const unsigned short pair_mask = (unsigned short)(3u << lead_rank);
    const int quad_id = blockIdx.x >> 2, num_quads = gridDim.x >> 2;
    const int total_tiles = prog.total_tiles, tiles_n = prog.tiles_n;   // tiles_n counts 128-row blocks (even)

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < prog.n_maps; i++) prefetch_tmap(&maps.m[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 2); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * WGS); }
        fence_barrier_init();
    }
    cluster_sync_all();
    if (warp == 2) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto advance = [&](int tile, int &si) {
        while (si + 1 < prog.n_steps && tile >= prog.step[si + 1].tile_base) si++;
    };

    if (warp == 0) {
        // ===== TMA producer (all four CTAs); warp-uniform loop, one elected lane issues
        {
            const bool issuer = elect_one();
            int stage = 0; uint32_t phase = 0;
            int si = 0;
            for (int tile = quad_id; tile < total_tiles; tile += num_quads) {
                advance(tile, si);
                const ChainStep &S = prog.step[si];
                const int local = tile - S.tile_base;
                const int m_blk = local % S.tiles_m, n_blk = 2 * (local / S.tiles_m) + (int)q;   // n_blk: this pair's 128-row block
                const int m0 = m_blk * 2 * BM + (int)pr * BM + (int)q * (BM / 2);                // my half of the weight rows both twins need
                const int n0 = n_blk * BN + (int)pr * (BN / 2);
                const int kb_total = (S.K + BK - 1) / BK;
                const CUtensorMap *mA = &maps.m[S.a_map], *mB = &maps.m[S.b_map];
                auto load_a = [&](int st, int kb) {
                    tma_load_2d_2sm_mc(smem_base + st * C::STAGE_BYTES + q * A_HALF_BYTES, mA, full_bar(st), kb * BK, m0, twin_mask);
                };
                int kb = 0;
                if (S.dep_step >= 0) {
                    const int pre = kb_total < C::STAGES ? kb_total : C::STAGES;
                    int st = stage; uint32_t ph = phase;
                    for (int i = 0; i < pre; i++) {
                        mbar_wait(empty_bar(st), ph ^ 1);
                        if (issuer) {
                            if (leader) mbar_expect_tx(full_bar(st), 2 * C::STAGE_BYTES);
                            else mbar_arrive_remote(full_bar(st), lead_rank);
                            load_a(st, i);
                        }
                        __syncwarp();
                        if (++st == C::STAGES) { st = 0; ph ^= 1; }
                    }
                    const unsigned int *cnt = prog.counters + S.dep_step * tiles_n + n_blk;
                    const unsigned int need = (unsigned int)S.dep_count;
                    if (ld_acquire_gpu(cnt) < need) {
                        const long long t0 = clock64();
                        unsigned int spins = 0;
                        while (ld_acquire_gpu(cnt) < need) {
                            __nanosleep(64);
                            if ((++spins & 255u) == 0u) {
                                if (*(volatile unsigned int *)prog.abort_flag) break;
                                if (clock64() - t0 > 1000000000ll) {
                                    *(volatile unsigned int *)prog.abort_flag = 1u;
                                    *(volatile unsigned int *)prog.error_flag = 1u;
                                    break;
                                }
                            }
                        }
                    }
                    fence_proxy_async_global();
                    if (issuer) {
                        int st2 = stage;
                        for (int i = 0; i < pre; i++) {
                            tma_load_2d_2sm(smem_base + st2 * C::STAGE_BYTES + C::A_BYTES, mB, full_bar(st2), i * BK, n0);
                            if (++st2 == C::STAGES) st2 = 0;
                        }
                    }
                    __syncwarp();
                    stage = st; phase = ph;
                    kb = pre;
                }
                for (; kb < kb_total; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (issuer) {
                        if (leader) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                        else mbar_arrive_remote(full_bar(stage), lead_rank);
                        load_a(stage, kb);
                        tma_load_2d_2sm(smem_base + stage * C::STAGE_BYTES + C::A_BYTES, mB, full_bar(stage), kb * BK, n0);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected lane of each pair's leader CTA
        if (leader) {
            const bool issuer = elect_one();
            constexpr uint32_t idesc = make_idesc_m256(BN, false);
            int stage = 0; uint32_t phase = 0;
            int iter = 0, si = 0;
            for (int tile = quad_id; tile < total_tiles; tile += num_quads, iter++) {
                advance(tile, si);
                const int kb_total = (prog.step[si].K + BK - 1) / BK;
                const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                uint32_t ready = mbar_test(full_bar(stage), phase);
                for (int kb = 0; kb < kb_total; kb++) {
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t adesc = make_kmajor_sw128_desc(sa), bdesc = make_kmajor_sw128_desc(sb);
                    const int cur = stage;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                    ready = (kb + 1 < kb_total) ? mbar_test(full_bar(stage), phase) : 0u;
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++)
                            umma_bf16_2sm(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        umma_commit_2sm_mask(empty_bar(cur), (unsigned short)0xF);   // this pair is done with the stage: tell all four producers
                        if (kb == kb_total - 1) umma_commit_2sm_mask(tfull_bar(as), pair_mask);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue: each CTA drains its own 128 accumulator rows x 128 batch rows, then publishes the tile
        const int quarter = warp & 3, wg = (warp - EPI_WARP0) >> 2;
        int iter = 0, si = 0;
        for (int tile = quad_id; tile < total_tiles; tile += num_quads, iter++) {
            advance(tile, si);
            const ChainStep &S = prog.step[si];
            const int local = tile - S.tile_base;
            const int m_blk = (local % S.tiles_m) * 2 + (int)pr, n_blk = 2 * (local / S.tiles_m) + (int)q;
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            switch (S.kind) {   // warp-uniform
                case K_FAST_SAMPLE: epilogue_fast<BN, WGS, true, false>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_SAMPLE_PROB: epilogue_fast<BN, WGS, true, true>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_GAUSS: epilogue_fast<BN, WGS, true, false, 1>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                case K_FAST_LABEL: epilogue_fast<BN, WGS, true, false, 2>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
                default: epilogue_fast<BN, WGS, false, true>(S.ep, tmem_acc, m_blk, n_blk, quarter, wg, lane); break;
            }
            tcgen05_fence_before();
            fence_proxy_async_global();
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(tempty_bar(as));
                else mbar_arrive_remote(tempty_bar(as), lead_rank);
                red_relaxed_gpu_add(prog.counters + si * tiles_n + n_blk, 1u);
            }
        }
    }
    __syncwarp();
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------- host side
// Collects the half-steps of one program; operands are registered as (pointer, extents) and deduplicated into the map table.
struct ChainBuilder {
    ChainMaps maps;
    ChainProgram prog;
    Operand ops[CHAIN_MAX_MAPS];
    int op_box[CHAIN_MAX_MAPS];
    int n_maps = 0, bn = 0;
    int pairs_limit = 0;   // CTA pairs the launch may use (0 = all): see chain_pairs_for in qb_api.cu
    bool quad = false;   // gibbs_chain_kernel_quad: bn = QUAD_BN rows per pair, tiles are 256 units x (2 x bn) rows
    double flops = 0.0;
    ChainBuilder() { memset(&maps, 0, sizeof(maps)); memset(&prog, 0, sizeof(prog)); }

    int map_index(const Operand &o, int box_rows) {
        for (int i = 0; i < n_maps; i++)
            if (ops[i].ptr == o.ptr && ops[i].rows == o.rows && ops[i].k == o.k && ops[i].ld == o.ld && op_box[i] == box_rows) return i;
        QB_CHECK(n_maps < CHAIN_MAX_MAPS, -1, "chain program uses too many distinct operands");
        ops[n_maps] = o; op_box[n_maps] = box_rows;
        maps.m[n_maps] = make_tmap(o, box_rows);
        return n_maps++;
    }
    bool full() const { return prog.n_steps >= CHAIN_MAX_STEPS; }
    // dep: index of the step whose outputs this one reads (-1 = none)
    int add(const Operand &A, const Operand &B, const EpiParams &ep, int dep) {
        QB_CHECK(!full(), -1, "chain program too long");
        QB_CHECK(A.k == B.k && A.k > 0, -1, "GEMM operands disagree on K");
        const int kind = pick_kind(ep);
        QB_CHECK(kind >= K_FAST_SAMPLE && kind <= K_FAST_LABEL, -5, "half-step has no compile-time epilogue: not chainable");
        ChainStep &S = prog.step[prog.n_steps];
#ifdef QB_EXP_BFRAC
        S.a_map = map_index(A, quad ? BM / 2 : BM); S.b_map = map_index(B, quad ? bn / 2 : bn / 2 / QB_EXP_BFRAC);
#else
        S.a_map = map_index(A, quad ? BM / 2 : BM); S.b_map = map_index(B, bn / 2);
#endif
        S.M = ep.M; S.K = (int)A.k; S.kind = kind; S.dep_step = dep;
        S.tiles_m = (int)ceil_div(ep.M, 2 * BM);
        S.tile_base = prog.total_tiles;
        S.ep = ep;
        S.dep_count = dep >= 0 ? prog.step[dep].tiles_m * 2 * 4 * Shape<256, K_FAST_SAMPLE>::WGS : 0;
        prog.total_tiles += S.tiles_m * (quad ? prog.tiles_n / 2 : prog.tiles_n);
        prog.n_maps = n_maps;
        flops += 2.0 * (double)ep.M * (double)ep.N * (double)A.k;
        return prog.n_steps++;
    }
    void begin(int N, int bn_, bool quad_ = false) {
        bn = bn_; quad = quad_;
        QB_CHECK(!quad || (bn == QUAD_BN && N % (2 * QUAD_BN) == 0), -1, "quad chain needs a multiple of 256 rows");
        prog.N = N; prog.tiles_n = (int)ceil_div(N, bn_); prog.n_steps = 0; prog.total_tiles = 0;
        n_maps = 0; flops = 0.0;
    }
};

// a program whose epilogues are all compile-time specialisations can run as one chain launch
inline bool chainable(const EpiParams &ep) {
    const int kind = pick_kind(ep);
    return kind >= K_FAST_SAMPLE && kind <= K_FAST_LABEL;
}

// tile width of a chain over N batch rows: as wide as possible while every CTA pair still gets work in a half-step
inline int chain_bn(long long N, long long M_min, int num_sms) {
    const long long pairs = num_sms / 2;
    for (int bn : {256, 128}) {
        if (ceil_div(N, bn) * ceil_div(M_min, 2 * BM) >= pairs || bn == 128) {
            if (N >= bn || bn == 128) return N <= 64 ? 64 : bn;
        }
    }
    return 128;
}

template <int BN>
inline void launch_chain_bn(cudaStream_t st, const ChainBuilder &cb, int num_sms) {
    using C = Cfg2<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        QB_CUDA(cudaFuncSetAttribute(gibbs_chain_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int pairs = (cb.pairs_limit > 0 && cb.pairs_limit < num_sms / 2) ? cb.pairs_limit : num_sms / 2;
    const int clusters = cb.prog.total_tiles < pairs ? cb.prog.total_tiles : pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(Shape<BN, K_FAST_SAMPLE>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QB_CUDA(cudaLaunchKernelEx(&cfg, gibbs_chain_kernel<BN>, cb.maps, cb.prog));
}

// clusters of four that can be co-resident (the static schedule needs every cluster of the grid on the machine at once; a
// GPC whose SM count is not a multiple of 4 strands some SMs: 33 clusters on a 148-SM B200)
inline int quad_clusters(int num_sms) {
    static int cached = -1;
    if (cached < 0) {
        using C = Cfg2<QUAD_BN>;
        QB_CUDA(cudaFuncSetAttribute(gibbs_chain_kernel_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(4 * 64); cfg.blockDim = dim3(Shape<QUAD_BN, K_FAST_SAMPLE>::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
        int n = 0;   // (the cluster shape is the kernel's compile-time __cluster_dims__)
        if (cudaOccupancyMaxActiveClusters(&n, gibbs_chain_kernel_quad, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&n, gibbs_chain_kernel_quad, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 0; }
        }
        cached = n;
    }
    return cached < num_sms / 4 ? cached : num_sms / 4;
}

inline void launch_chain_quad(cudaStream_t st, const ChainBuilder &cb, int num_sms) {
    using C = Cfg2<QUAD_BN>;
    const int quads = quad_clusters(num_sms);
    QB_CHECK(quads >= 1, -5, "no co-resident 4-CTA clusters available for the quad chain kernel");
    const int clusters = cb.prog.total_tiles < quads ? cb.prog.total_tiles : quads;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(4 * clusters)); cfg.blockDim = dim3(Shape<QUAD_BN, K_FAST_SAMPLE>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QB_CUDA(cudaLaunchKernelEx(&cfg, gibbs_chain_kernel_quad, cb.maps, cb.prog));
}

inline void launch_chain(cudaStream_t st, const ChainBuilder &cb, int num_sms) {
    QB_CHECK(cb.prog.n_steps >= 1 && cb.prog.counters != nullptr && cb.prog.error_flag != nullptr && cb.prog.abort_flag != nullptr, -1, "empty chain program");
    if (cb.quad) { launch_chain_quad(st, cb, num_sms); QB_CUDA(cudaGetLastError()); return; }
    switch (cb.bn) {
        case 256: launch_chain_bn<256>(st, cb, num_sms); break;
        case 128: launch_chain_bn<128>(st, cb, num_sms); break;
        default: launch_chain_bn<64>(st, cb, num_sms); break;
    }
    QB_CUDA(cudaGetLastError());
}

}  // namespace sm100
}  // namespace qb
