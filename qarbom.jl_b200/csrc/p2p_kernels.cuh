// Data-parallel exchange over NVLink peer memory, fused with the parameter update (bf16 mode, world > 1).
//
// After the gradient GEMM every rank holds its partial G.  Instead of ncclAllReduce(G) followed by an update
// kernel that every rank repeats on the full matrix, ONE kernel per rank does
//     reduce-scatter : for the hidden rows it OWNS, g = sum over ranks of G_r[h][c]   (P2P 128-bit loads)
//     update         : w += lr/B * g (+ momentum / weight decay) on its shard of the fp32 master
//     all-gather     : the new bf16 shadows Wb[h][c] and WbT[c][h] are stored straight into EVERY rank's
//                      shadow buffers (P2P stores)
// so the bytes that cross NVLink are 4 B/param in and 4 B/param out per shard (the all-gather moves bf16, not
// fp32), the update work is 1/world per rank, and no reduced fp32 gradient is ever written back.  Two
// system-scope flag barriers bracket the kernel: "every G is complete" and "every push has landed".
// The kernel is launched cooperatively (grid barrier between the flag barriers and the tile loop); each GPU
// runs its own kernel, ranks never share a GPU.
//
// NVLS variant (template parameter NVLS; addresses from nvls_setup.cu): the same kernel over NVSwitch MULTICAST addresses.
//     reduce-scatter : ONE multimem.ld_reduce.add.v4.f32 per 16 bytes -- the switch sums the ranks' gradients in flight, the
//                      owner receives 4 B/param instead of 4 B/param from every peer
//     all-gather     : ONE multimem.st per 8 bytes of the new row-major bf16 shadow -- the switch replicates it into every
//                      rank's buffer (2 B/param out per shard instead of 2 B/param to every peer)
// The transposed shadow is NOT pushed in this variant (it would double the inbound NVLink bytes of every GPU); each rank
// rebuilds it from the complete row-major shadow with a local transpose kernel, HBM-bound and 5x cheaper than the link.
// Biases are reduced over unicast peer addresses in rank order by every rank, so the replicas stay bit-identical.
//
// SPLIT variant (binary PCD, option "split_exchange"): the gradient is exchanged as its two products.  The positive product
// P_h' X depends only on W_t and the data, so it is computed BEFORE the chain advance and reduced behind it by a small kernel
// (nvls_reduce_shard_kernel: blocks of 128 threads at <= 32 registers run beside the persistent GEMM CTAs and on the SMs a
// one-wave half-step leaves idle).  The negative product H~' V~ of binary chains is a matrix of integer counts <= rows per
// rank: the gradient GEMM stores it as uint16, and ONE multimem.ld_reduce.add.u64 sums four packed counts of every rank
// without carry (global batch < 65536) -- exact, at half the bytes of the fp32 reduction.  What stays exposed after the chain
// advance is 2 B/param instead of 4.
#pragma once
#include "common.cuh"
#include "online_kernels.cuh"  // grid_barrier

namespace qb {

constexpr int QB_MAX_PEERS = 8;

struct P2pParams {
    int rank, world;
    const float *G[QB_MAX_PEERS];       // every rank's gradient buffer [W | a | b]
    const float *G_mc;                  // NVLS: multicast address of the gradient buffers (multimem.ld_reduce = sum over ranks)
    bf16 *Wb_mc;                        // NVLS: multicast address of the row-major shadows (multimem.st = store to all ranks)
    const unsigned short *Gn16_mc;      // SPLIT: multicast address of the uint16 negative products; G[rank] then holds the REDUCED positive product of the owned rows
    bf16 *Wb[QB_MAX_PEERS], *WbT[QB_MAX_PEERS];
    unsigned int *flags[QB_MAX_PEERS];  // flags[p][r]: written by rank r, polled by rank p
    float *W, *Vel;                     // local fp32 master / velocity ([H][ldv] + biases)
    float *bias_rd;                     // optional versioned bias copy (stale-chains mode), else nullptr
    int H, VL, V, h_begin, h_end;
    long long ldv, ldh, nW;
    float lr, llr, momentum, wd;
    unsigned int epoch;                 // flag value of the first barrier of this launch (second = epoch + 1)
    unsigned int *grid_bar;
    unsigned long long *phase_ns;       // [2] accumulated ns: waiting for the peers' gradients, reduce + update + push
    unsigned int *timeout_flag;         // mapped host memory: set when a peer did not show up within the deadline
};

__device__ __forceinline__ void xgpu_barrier(const P2pParams &p, unsigned int value) {
    if (threadIdx.x < (unsigned)p.world) {
        __threadfence_system();
        unsigned int *remote = p.flags[threadIdx.x] + p.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(value) : "memory");
        const unsigned int *mine = p.flags[p.rank] + threadIdx.x;
        unsigned int seen;
        // bounded wait: a peer that threw, ran a different number of steps or died must not wedge this GPU.  After the
        // deadline (~5 s) the rank raises a host-visible flag and carries on; p2p_update turns it into QB_E_CUDA.
        const unsigned long long t0 = gtimer();
        unsigned int spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int)(seen - value) >= 0) break;
            if ((++spins & 1023u) == 0u && gtimer() - t0 > 5000000000ull) {
                if (p.timeout_flag) *(volatile unsigned int *)p.timeout_flag = 1u;
                break;
            }
        } while (true);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_b32x2(void *mc, uint2 v) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}

__device__ __forceinline__ unsigned long long multimem_ld_reduce_add_u64(const void *mc) {
    unsigned long long v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
    return v;
}

// ---- SPLIT, part 1: "my positive product is complete" to every peer (one block, after the positive-product GEMM in stream order)
__global__ void nvls_signal_kernel(const P2pParams p) {
    if (threadIdx.x < (unsigned)p.world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[threadIdx.x] + p.rank), "r"(p.epoch) : "memory");
    }
}
// ---- SPLIT, part 2: G[rank][owned rows] <- sum over ranks of G_r[owned rows], in place, through the switch.  No block depends
// on another block of this grid (each waits for the peers' flags itself), so it is safe to launch beside a persistent kernel
// that leaves only a few block slots free.
__global__ void __launch_bounds__(128, 16) nvls_reduce_shard_kernel(const P2pParams p, float *G_local) {
    if (threadIdx.x < (unsigned)p.world) {
        const unsigned int *mine = p.flags[p.rank] + threadIdx.x;
        unsigned int seen;
        const unsigned long long t0 = gtimer();
        unsigned int spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int)(seen - p.epoch) >= 0) break;
            if ((++spins & 1023u) == 0u && gtimer() - t0 > 5000000000ull) {
                if (p.timeout_flag) *(volatile unsigned int *)p.timeout_flag = 1u;
                break;
            }
        } while (true);
    }
    __syncthreads();
    const long long base = (long long)p.h_begin * p.ldv;
    const long long n4 = ((long long)(p.h_end - p.h_begin) * p.ldv) >> 2;   // ldv % 8 == 0
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u * stride < n4) v[u] = multimem_ld_reduce_add_f32x4(p.G_mc + base + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u * stride < n4) *reinterpret_cast<float4 *>(G_local + base + 4 * (i + u * stride)) = v[u];
    }
}

template <bool VEL, bool NVLS = false, bool SPLIT = false>
__global__ void __launch_bounds__(256) p2p_reduce_update_kernel(const P2pParams p) {
    constexpr int TH = 64, TC = 128, PITCH = TC + 8, ROWS = NVLS ? 4 : 2;   // NVLS: one 16-byte load per row, so more rows in flight
    __shared__ __align__(16) unsigned short tile[TH][PITCH];
    unsigned int gepoch = 0;
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0 && p.phase_ns != nullptr;
    unsigned long long t0 = timer ? gtimer() : 0ull;
    if (blockIdx.x == 0) xgpu_barrier(p, p.epoch);  // every rank's G is complete
    grid_barrier(p.grid_bar, gepoch, gridDim.x);
    unsigned long long t1 = timer ? gtimer() : 0ull;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_c = (p.VL + TC - 1) / TC;
    const int tiles_h = (p.h_end - p.h_begin + TH - 1) / TH;
    for (int t = blockIdx.x; t < tiles_c * tiles_h; t += gridDim.x) {
        const int c = (t % tiles_c) * TC + 4 * tx;
        const int h0 = p.h_begin + (t / tiles_c) * TH;
        const bool vec = (c + 3 < p.VL);
        float lrc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) lrc[i] = (c + i < p.V) ? p.lr : p.llr;
#pragma unroll 1
        for (int pass = 0; pass < TH / (8 * ROWS); pass++) {
            float4 w[ROWS], g[ROWS], v[ROWS], q[ROWS][NVLS ? 1 : QB_MAX_PEERS];
            // issue every load first (fully unrolled, predicated on the rank count): a remote 128-bit load takes
            // ~2-3 us over NVLink, so ROWS x world of them must be in flight per thread, not one after the other
#pragma unroll
            for (int it = 0; it < ROWS; it++) {
                const int h = h0 + ty + 8 * (pass * ROWS + it);
                const bool ok = h < p.h_end && c < p.VL;
                const long long o = (long long)h * p.ldv + c;  // ldv % 8 == 0, padding columns are zero: 128-bit loads stay in bounds
                w[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (VEL) v[it] = w[it];
                if (NVLS && SPLIT) {
                    q[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) {
                        // reduced positive product (this rank's own buffer) minus the negative counts of every rank: four uint16
                        // per 8 bytes, summed by the switch as ONE u64 (no carry between the fields: global batch < 65536)
                        const float4 gp = *reinterpret_cast<const float4 *>(p.G[p.rank] + o);
                        const unsigned long long c = multimem_ld_reduce_add_u64(p.Gn16_mc + o);
                        q[it][0] = make_float4(gp.x - (float)(unsigned)(c & 0xFFFFull), gp.y - (float)(unsigned)((c >> 16) & 0xFFFFull),
                                               gp.z - (float)(unsigned)((c >> 32) & 0xFFFFull), gp.w - (float)(unsigned)(c >> 48));
                    }
                } else if (NVLS) {
                    q[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) q[it][0] = multimem_ld_reduce_add_f32x4(p.G_mc + o);   // summed over the ranks inside the switch
                } else {
#pragma unroll
                    for (int r = 0; r < (NVLS ? 1 : QB_MAX_PEERS); r++) {
                        q[it][r] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok && r < p.world) q[it][r] = __ldcg(reinterpret_cast<const float4 *>(p.G[r] + o));
                    }
                }
                if (ok) {
                    w[it] = *reinterpret_cast<const float4 *>(p.W + o);
                    if (VEL) v[it] = *reinterpret_cast<const float4 *>(p.Vel + o);
                }
            }
#pragma unroll
            for (int it = 0; it < ROWS; it++) {
                g[it] = q[it][0];
#pragma unroll
                for (int r = 1; r < (NVLS ? 1 : QB_MAX_PEERS); r++) { g[it].x += q[it][r].x; g[it].y += q[it][r].y; g[it].z += q[it][r].z; g[it].w += q[it][r].w; }
            }
#pragma unroll
            for (int it = 0; it < ROWS; it++) {
                const int hl = ty + 8 * (pass * ROWS + it), h = h0 + hl;
                float *wp = reinterpret_cast<float *>(&w[it]), *gp = reinterpret_cast<float *>(&g[it]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (VEL) {
                        float *vp = reinterpret_cast<float *>(&v[it]);
                        vp[i] = p.momentum * vp[i] + lrc[i] * gp[i] - lrc[i] * p.wd * wp[i];
                        wp[i] += vp[i];
                    } else {
                        wp[i] += lrc[i] * gp[i];
                    }
                    if (!vec && c + i >= p.VL) wp[i] = 0.f;
                }
                __nv_bfloat162 b01 = __floats2bfloat162_rn(wp[0], wp[1]), b23 = __floats2bfloat162_rn(wp[2], wp[3]);
                const uint2 packed = make_uint2(*reinterpret_cast<uint32_t *>(&b01), *reinterpret_cast<uint32_t *>(&b23));
                if (!NVLS) *reinterpret_cast<uint2 *>(&tile[hl][4 * tx]) = packed;
                if (h < p.h_end && c < p.VL) {
                    const long long o = (long long)h * p.ldv + c;
                    *reinterpret_cast<float4 *>(p.W + o) = w[it];
                    if (VEL) *reinterpret_cast<float4 *>(p.Vel + o) = v[it];
                    if (NVLS) multimem_st_b32x2(p.Wb_mc + o, packed);   // all-gather by the switch: one store, every rank's shadow
                    else for (int r = 0; r < p.world; r++) *reinterpret_cast<uint2 *>(p.Wb[r] + o) = packed;  // all-gather, row-major shadow
                }
            }
        }
        if (NVLS) continue;   // the transposed shadow is rebuilt locally after the kernel
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; q++) {  // all-gather, transposed shadow: 32 B per (column, 16-row chunk)
            const int task = threadIdx.x + 256 * q;
            const int cl = task & (TC - 1), hc = task >> 7;
            const int cc = (t % tiles_c) * TC + cl, hb = h0 + 16 * hc;
            if (cc >= p.VL || hb >= p.h_end) continue;
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; i++) pk[i] = (uint32_t)tile[16 * hc + 2 * i][cl] | ((uint32_t)tile[16 * hc + 2 * i + 1][cl] << 16);
            for (int r = 0; r < p.world; r++) {
                unsigned short *dst = reinterpret_cast<unsigned short *>(p.WbT[r]) + (long long)cc * p.ldh + hb;
                if (hb + 16 <= p.h_end) {
                    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
                    for (int i = 0; i < 16; i++)
                        if (hb + i < p.h_end) dst[i] = tile[16 * hc + i][cl];
                }
            }
        }
        __syncthreads();
    }
    // biases [acat | b]: tiny, every rank reduces and updates its own full copy (same order => identical everywhere)
    const int nbias = (int)(p.ldv + p.ldh);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbias; i += gridDim.x * blockDim.x) {
        float g = 0.f;
        for (int r = 0; r < p.world; r++) g += __ldcg(p.G[r] + p.nW + i);
        const float lr_c = (i < p.ldv && i >= p.V) ? p.llr : p.lr;
        float b = p.W[p.nW + i];
        if (VEL) {
            float vv = p.momentum * p.Vel[p.nW + i] + lr_c * g;
            p.Vel[p.nW + i] = vv;
            b += vv;
        } else {
            b += lr_c * g;
        }
        p.W[p.nW + i] = b;
        if (p.bias_rd) p.bias_rd[i] = b;
    }
    __threadfence_system();
    grid_barrier(p.grid_bar, gepoch, gridDim.x);
    if (timer) { const unsigned long long t2 = gtimer(); p.phase_ns[0] += t1 - t0; p.phase_ns[1] += t2 - t1; }
    if (blockIdx.x == 0) xgpu_barrier(p, p.epoch + 1);  // every push has landed; nobody still reads my G
}

}  // namespace qb
