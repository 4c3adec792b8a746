// What one thread pays to ISSUE tcgen05.mma / tcgen05.commit / an mbarrier wait on sm_100a (development probe, ONE CTA).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qarbom.jl_b200/csrc tools/mma_issue_probe.cu -I include -o tools/_bin/mma_issue_probe
//
// Each pattern is R "k-steps" of four 128 x N x 16 bf16 MMAs on zeroed shared memory; the issuing warp reads the SM clock
// before and after the loop (issue cost) and after a final commit has arrived (execution).  Patterns:
//   0  4 MMAs                                    warp-uniform loop, elected lane issues
//   1  4 MMAs + commit                           "
//   2  wait(completed mbarrier) + 4 MMAs + commit "
//   3  wait + tcgen05.fence::after + 4 MMAs + commit "
//   4  as 3 inside an `if (lane == 0)` region (what the kernels did before the elect_one change)
//   5  8 MMAs + commit per step (a 128-wide k-step), wait + no fence
//   6  as 2 with the wait software-pipelined: ONE try_wait on the next step's barrier is issued before this step's MMAs,
//      the blocking wait runs only if that try failed
//   7  as 2 with the issuing lane elected once, outside the loop
//   8  6 + 7
//   9  8 without the __syncwarp at the end of the step
//  10  9 with 8 MMAs per step
//  11  8 with the MMAs and the commit predicated inside the asm (no branch around them)
//  12  11 with the barrier of step i + 2 tested (two steps of look-ahead)
#include <cstdio>
#include <cuda_runtime.h>

#include "sm100_gemm.cuh"

using namespace qb::sm100;

constexpr int R = 64;

__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
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
__device__ __forceinline__ void umma_commit_if(uint32_t issue, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(issue)
        : "memory");
}

template <int N, int PATTERN>
__global__ void __launch_bounds__(128, 1) probe_kernel(long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sa = smem_base, sb = smem_base + 32768;             // 128 x 128 bf16 of A, 256 x 128 bf16 of B (zeros)
    const uint32_t bar = smem_base + 32768 + 65536, done_bar = bar + 8, ring = bar + 16, tmem_slot = bar + 16 + 8 * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)))[i] = 0u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(done_bar, 1);
        for (int s = 0; s < 16; s++) mbar_init(ring + 8 * s, 1);
        fence_barrier_init();
        mbar_arrive(bar);   // phase 0 of `bar` is complete for good: a wait on parity 0 returns at once
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    constexpr uint32_t idesc = make_idesc(N, false);
    if (warp == 1) {
        long long t0 = 0, t1 = 0, t2 = 0;
        if (PATTERN == 4) {
            if (lane == 0) {
                t0 = clock64();
                for (int it = 0; it < R; it++) {
                    mbar_wait(bar, 0);
                    tcgen05_fence_after();
                    const uint64_t adesc = make_kmajor_sw128_desc(sa + (it & 1) * 16384), bdesc = make_kmajor_sw128_desc(sb + (it & 1) * 32768);
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
                    umma_commit(ring + 8 * (it & 15));
                }
                t1 = clock64();
                umma_commit(done_bar);
                mbar_wait(done_bar, 0);
                t2 = clock64();
            }
        } else if (PATTERN >= 11) {
            const uint32_t issue = elect_one() ? 1u : 0u;
            t0 = clock64();
            uint32_t ready = mbar_try(bar, 0), ready2 = (PATTERN == 12) ? mbar_try(bar, 0) : 0u;
            for (int it = 0; it < R; it++) {
                if (!ready) mbar_wait(bar, 0);
                const uint64_t adesc = make_kmajor_sw128_desc(sa + (it & 1) * 16384), bdesc = make_kmajor_sw128_desc(sb + (it & 1) * 32768);
                if (PATTERN == 12) { ready = ready2; ready2 = mbar_try(bar, 0); }
                else ready = mbar_try(bar, 0);
#pragma unroll
                for (int k = 0; k < 4; k++) umma_bf16_if(issue, tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
                umma_commit_if(issue, ring + 8 * (it & 15));
            }
            t1 = clock64();
            if (elect_one()) umma_commit(done_bar);
            mbar_wait(done_bar, 0);
            t2 = clock64();
        } else if (PATTERN >= 6) {
            const bool leader = elect_one();
            t0 = clock64();
            uint32_t ready = (PATTERN == 7) ? 0u : mbar_try(bar, 0);
            for (int it = 0; it < R; it++) {
                if (!ready) mbar_wait(bar, 0);
                const uint64_t adesc = make_kmajor_sw128_desc(sa + (it & 1) * 16384), bdesc = make_kmajor_sw128_desc(sb + (it & 1) * 32768);
                if (PATTERN != 7) ready = mbar_try(bar, 0);   // next step's barrier, in flight while the MMAs below are issued
                if (PATTERN == 6 ? elect_one() : leader) {
                    if (PATTERN == 10) {
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
                    umma_commit(ring + 8 * (it & 15));
                }
                if (PATTERN < 9) __syncwarp();
            }
            t1 = clock64();
            if (elect_one()) umma_commit(done_bar);
            mbar_wait(done_bar, 0);
            t2 = clock64();
        } else {
            t0 = clock64();
            for (int it = 0; it < R; it++) {
                if (PATTERN >= 2) mbar_wait(bar, 0);
                if (PATTERN == 3) tcgen05_fence_after();
                const uint64_t adesc = make_kmajor_sw128_desc(sa + (it & 1) * 16384), bdesc = make_kmajor_sw128_desc(sb + (it & 1) * 32768);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; k++) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
                    if (PATTERN == 5) {
#pragma unroll
                        for (int k = 0; k < 4; k++) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1);
                    }
                    if (PATTERN >= 1) umma_commit(ring + 8 * (it & 15));
                }
            }
            t1 = clock64();
            if (elect_one()) umma_commit(done_bar);
            mbar_wait(done_bar, 0);
            t2 = clock64();
        }
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int N, int PATTERN>
void run(long long *d_out) {
    const int smem = 32768 + 65536 + 1024 + 512;
    cudaFuncSetAttribute(probe_kernel<N, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[2] = {0, 0}, best[2] = {1ll << 60, 1ll << 60};
    for (int rep = 0; rep < 5; rep++) {
        probe_kernel<N, PATTERN><<<1, 128, smem>>>(d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("N=%d pattern %d: %s\n", N, PATTERN, cudaGetErrorString(cudaGetLastError())); return; }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 2; i++) best[i] = h[i] < best[i] ? h[i] : best[i];
    }
    printf("N=%3d pattern %d: issue %6.1f clocks per k-step, issue + drain %6.1f clocks per k-step\n", N, PATTERN, (double)best[0] / R, (double)best[1] / R);
}

template <int N>
void run_all(long long *d_out) {
    run<N, 0>(d_out); run<N, 1>(d_out); run<N, 2>(d_out); run<N, 3>(d_out); run<N, 4>(d_out); run<N, 5>(d_out); run<N, 6>(d_out); run<N, 7>(d_out); run<N, 8>(d_out); run<N, 9>(d_out); run<N, 10>(d_out); run<N, 11>(d_out); run<N, 12>(d_out);
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    run_all<32>(d_out); run_all<64>(d_out); run_all<128>(d_out); run_all<256>(d_out);
    return 0;
}
