// fp32 CUDA-core kernels: the validation-precision path (QB_PREC_FP32) and the
// bandwidth-bound helper kernels shared with the tensor-core path (update, column sums,
// conversions, transposes, metric reductions).
#pragma once
#include "common.cuh"

namespace qb {

// ----------------------------------------------------------------------------- fp32 GEMM
// C[m*ldc + n] = alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + beta * C + bias_n[n]
// 64x64 tile, 16-deep k slices, 256 threads, 4x4 register micro-tile, sequential-k fp32 FMA.
__global__ void __launch_bounds__(256) simt_gemm_kernel(const float *__restrict__ A, int64_t sam, int64_t sak,
                                                        const float *__restrict__ B, int64_t sbk, int64_t sbn,
                                                        float *__restrict__ C, int64_t ldc, int M, int N, int K,
                                                        float alpha, float beta, const float *__restrict__ bias_n) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tm = (t / 16) * 4, tn = (t % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int idx = t + i * 256;
            int m, k;
            if (sak == 1) { m = idx / 16; k = idx % 16; } else { m = idx % 64; k = idx / 64; }
            float v = 0.f;
            if (m0 + m < M && k0 + k < K) v = A[(int64_t)(m0 + m) * sam + (int64_t)(k0 + k) * sak];
            As[k][m] = v;
            int n, kb;
            if (sbn == 1) { n = idx % 64; kb = idx / 64; } else { kb = idx % 16; n = idx / 16; }
            float w = 0.f;
            if (n0 + n < N && k0 + kb < K) w = B[(int64_t)(k0 + kb) * sbk + (int64_t)(n0 + n) * sbn];
            Bs[kb][n] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[k][tm + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[k][tn + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + tm + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tn + j;
            if (n >= N) continue;
            float v = alpha * acc[i][j];
            if (beta != 0.f) v += beta * C[(int64_t)m * ldc + n];
            if (bias_n) v += bias_n[n];
            C[(int64_t)m * ldc + n] = v;
        }
    }
}

// ----------------------------------------------------------------------------- fp32 -> three bf16 components (QB_PREC_FP32X3)
// x = hi + mid + lo with hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid): 3 x 8 significand bits cover fp32's 24.
// src element (r, k) at src[r * s_row + k * s_k]; components are written K-major, c[i][r * ldk + k], columns [K, ldk) zeroed.
// 32 x 32 tiles through shared memory so that both the strided read (s_k != 1: a transposed view) and the write coalesce.
__global__ void __launch_bounds__(256) split3_kernel(const float *__restrict__ src, int64_t s_row, int64_t s_k, int rows, int K,
                                                     bf16 *__restrict__ c0, bf16 *__restrict__ c1, bf16 *__restrict__ c2, int64_t ldk) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const bool k_fast = (s_k == 1);   // which source index is contiguous
    for (int i = ty; i < 32; i += 8) {
        const int r = k_fast ? r0 + i : r0 + tx, k = k_fast ? k0 + tx : k0 + i;
        float v = 0.f;
        if (r < rows && k < K) v = src[(int64_t)r * s_row + (int64_t)k * s_k];
        if (k_fast) tile[i][tx] = v; else tile[tx][i] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, k = k0 + tx;
        if (r >= rows || k >= ldk) continue;
        const float v = tile[i][tx];
        const bf16 h = __float2bfloat16(v);
        const float r1 = v - __bfloat162float(h);
        const bf16 m = __float2bfloat16(r1);
        const bf16 l = __float2bfloat16(r1 - __bfloat162float(m));
        const int64_t o = (int64_t)r * ldk + k;
        c0[o] = h;
        if (c1) c1[o] = m;
        if (c2) c2[o] = l;
    }
}

// ----------------------------------------------------------------------------- activation + sampling
// act: 0 = sigmoid (+ Bernoulli sample when `state`), 1 = Gaussian visible (value = pre [+ z])
// Units in [0, n_units) of every row; label units are finished by label_kernel instead.
__global__ void act_sample_kernel(const float *__restrict__ pre, int64_t ld, int rows, int n_units, int act,
                                  float *__restrict__ prob, int64_t ldp, float *__restrict__ state, int64_t lds,
                                  const float *__restrict__ inj, int64_t ldi, RngKey key, int64_t row0, int use_rng) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * n_units) return;
    int r = (int)(i / n_units), u = (int)(i % n_units);
    float x = pre[(int64_t)r * ld + u];
    if (act == 0) {
        float p = sigmoid_exact(x);
        if (prob) prob[(int64_t)r * ldp + u] = p;
        if (state) {
            float rn = inj ? inj[(int64_t)r * ldi + u] : rng_uniform(key, (uint32_t)u, (uint64_t)(row0 + r));
            state[(int64_t)r * lds + u] = (rn < p) ? 1.f : 0.f;
        }
    } else {
        if (prob) prob[(int64_t)r * ldp + u] = x;
        if (state) {
            float z = 0.f;
            if (inj) z = inj[(int64_t)r * ldi + u];
            else if (use_rng) z = rng_normal(key, (uint32_t)u, (uint64_t)(row0 + r));
            float s = x + z;
            // act 2 (FastPCD chains of a GRBM, src/gibbs.jl:22-25): the noisy value is thresholded like a probability
            if (act == 2) s = (rng_uniform(key, (uint32_t)u, (uint64_t)(row0 + r)) < s) ? 1.f : 0.f;
            state[(int64_t)r * lds + u] = s;
        }
    }
}

// gibbs_sample_label (src/gibbs.jl:46-49 + src/rbm.jl:196-205): softmax over the L label
// pre-activations of a row, then independent Bernoulli per label unit.  One thread per row.
__global__ void label_kernel(const float *__restrict__ pre, int64_t ld, int rows, int V, int L, float *__restrict__ prob,
                             int64_t ldp, float *__restrict__ state, int64_t lds, const float *__restrict__ inj,
                             int64_t ldi, RngKey key, int64_t row0) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *p = pre + (int64_t)r * ld + V;
    float mx = -INFINITY;
    for (int l = 0; l < L; l++) mx = fmaxf(mx, p[l]);
    float s = 0.f;
    for (int l = 0; l < L; l++) s += expf(p[l] - mx);
    float lden = mx + logf(s);
    for (int l = 0; l < L; l++) {
        float q = expf(p[l] - lden);
        if (prob) prob[(int64_t)r * ldp + V + l] = q;
        if (state) {
            float rn = inj ? inj[(int64_t)r * ldi + V + l] : rng_uniform(key, (uint32_t)(V + l), (uint64_t)(row0 + r));
            state[(int64_t)r * lds + V + l] = (rn < q) ? 1.f : 0.f;
        }
    }
}

// ----------------------------------------------------------------------------- column sums (bias gradients)
// out[c] (+)= sign * sum_r A[r*lda + c]; 32 columns x 8 row-lanes per block, rows strided by gridDim.y*8.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T *__restrict__ A, int64_t lda, int rows, int cols,
                                                     float sign, float *__restrict__ out) {
    __shared__ float red[8][33];
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int rl = threadIdx.x >> 5;
    float acc = 0.f;
    if (c < cols)
        for (int r = blockIdx.y * 8 + rl; r < rows; r += gridDim.y * 8) acc += (float)A[(int64_t)r * lda + c];
    red[rl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += red[i][threadIdx.x & 31];
        atomicAdd(out + c, sign * s);
    }
}

// ----------------------------------------------------------------------------- parameter update (K5)
// W region: tile [64 h][128 c] per block, 128-bit loads/stores (4 columns per thread, 8 rows in
// flight per thread); fp32 read-modify-write of theta (+ velocity), refresh of the bf16 shadows
// Wb[h][c] (8-byte stores) and, through a shared-memory tile, WbT[c][h] (32 B per thread).
//   vel = momentum*vel + lr_c*g - lr_c*wd*w ;  w += vel     (momentum = wd = 0: w += lr_c*g,
//   i.e. update_rbm!, src/rbm.jl:152-163 / :120-136 with lr_c = lr/B for c < V, llr/B for labels)
// Algorithmic bytes per parameter: 12 (read w, read g, write w) + 4 (two bf16 shadows); +8 with velocity.
// The bias region [acat | b] that follows W in the flat vectors is updated by one extra row of blocks
// (blockIdx.y == tiles over H) of the same launch: g = G + dscale * datasum (cached data-side visible sums,
// single-GPU path), optional versioned copy, and optionally G <- 0 so the next step needs no init kernel.
struct BiasTail {
    float *P, *G, *Vel; const float *datasum; float dscale; float *copy_out; int n_vis, n_total, zero_after;
};

template <bool VEL>
__global__ void __launch_bounds__(256, 3) update_w_kernel(float *__restrict__ W, const float *__restrict__ G,
                                                          float *__restrict__ Vel, int H, int VL, int V, int64_t ldv,
                                                          float lr, float llr, float momentum, float wd,
                                                          bf16 *__restrict__ Wb, bf16 *__restrict__ WbT, int64_t ldh,
                                                          const BiasTail bt) {
    constexpr int TH = 64, TC = 128, PITCH = TC + 8, ROWS = 4;  // ROWS rows in flight per thread
    // programmatic dependent launch: wait for the gradient GEMM, then let the next GEMM start its prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((int)blockIdx.y * TH >= H) {  // the bias blocks
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bt.n_total; i += gridDim.x * blockDim.x) {
            const float lr_c = (i < bt.n_vis && i >= V) ? llr : lr;  // label biases c use llr; a and b use lr
            float g = bt.G[i];
            if (bt.datasum && i < bt.n_vis) g += bt.dscale * bt.datasum[i];
            float pv = bt.P[i];
            if (VEL) {
                const float v = momentum * bt.Vel[i] + lr_c * g;
                bt.Vel[i] = v;
                pv += v;
            } else {
                pv += lr_c * g;
            }
            bt.P[i] = pv;
            if (bt.copy_out) bt.copy_out[i] = pv;
            if (bt.zero_after) bt.G[i] = 0.f;
        }
        return;
    }
    __shared__ __align__(16) unsigned short tile[TH][PITCH];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * TC + 4 * tx;
    const int h0 = blockIdx.y * TH;
    const bool vec = (c + 3 < VL);
    float lrc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) lrc[i] = (c + i < V) ? lr : llr;
#pragma unroll 1
    for (int pass = 0; pass < TH / (8 * ROWS); pass++) {
        float4 w[ROWS], g[ROWS], v[ROWS];
#pragma unroll
        for (int it = 0; it < ROWS; it++) {  // all loads first: independent 128-bit requests
            const int h = h0 + ty + 8 * (pass * ROWS + it);
            w[it] = make_float4(0.f, 0.f, 0.f, 0.f); g[it] = w[it];
            if (VEL) v[it] = w[it];
            if (h < H) {
                const int64_t o = (int64_t)h * ldv + c;
                if (vec) {
                    w[it] = *reinterpret_cast<const float4 *>(W + o);
                    g[it] = *reinterpret_cast<const float4 *>(G + o);
                    if (VEL) v[it] = *reinterpret_cast<const float4 *>(Vel + o);
                } else {
                    float *wp = reinterpret_cast<float *>(&w[it]), *gp = reinterpret_cast<float *>(&g[it]);
                    for (int i = 0; i < 4; i++)
                        if (c + i < VL) { wp[i] = W[o + i]; gp[i] = G[o + i]; if (VEL) reinterpret_cast<float *>(&v[it])[i] = Vel[o + i]; }
                }
            }
        }
#pragma unroll
        for (int it = 0; it < ROWS; it++) {
            const int hl = ty + 8 * (pass * ROWS + it), h = h0 + hl;
            float *wp = reinterpret_cast<float *>(&w[it]), *gp = reinterpret_cast<float *>(&g[it]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (VEL) {
                    float *vp = reinterpret_cast<float *>(&v[it]);
                    vp[i] = momentum * vp[i] + lrc[i] * gp[i] - lrc[i] * wd * wp[i];
                    wp[i] += vp[i];
                } else {
                    wp[i] += lrc[i] * gp[i];
                }
            }
            __nv_bfloat162 b01 = __floats2bfloat162_rn(wp[0], wp[1]), b23 = __floats2bfloat162_rn(wp[2], wp[3]);
            const uint2 packed = make_uint2(*reinterpret_cast<uint32_t *>(&b01), *reinterpret_cast<uint32_t *>(&b23));
            *reinterpret_cast<uint2 *>(&tile[hl][4 * tx]) = packed;
            if (h < H) {
                const int64_t o = (int64_t)h * ldv + c;
                if (vec) {
                    *reinterpret_cast<float4 *>(W + o) = w[it];
                    if (VEL) *reinterpret_cast<float4 *>(Vel + o) = v[it];
                    if (Wb) *reinterpret_cast<uint2 *>(Wb + o) = packed;
                } else {
                    for (int i = 0; i < 4; i++)
                        if (c + i < VL) {
                            W[o + i] = wp[i];
                            if (VEL) Vel[o + i] = reinterpret_cast<float *>(&v[it])[i];
                            if (Wb) Wb[o + i] = __float2bfloat16(wp[i]);
                        }
                }
            }
        }
    }
    if (!WbT) return;
    __syncthreads();
    // transposed shadow: task = (column, 16-row chunk); consecutive lanes take consecutive columns
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int task = threadIdx.x + 256 * q;
        const int cl = task & (TC - 1), hc = task >> 7;
        const int cc = blockIdx.x * TC + cl, hb = h0 + 16 * hc;
        if (cc >= VL || hb >= H) continue;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; i++) pk[i] = (uint32_t)tile[16 * hc + 2 * i][cl] | ((uint32_t)tile[16 * hc + 2 * i + 1][cl] << 16);
        unsigned short *dst = reinterpret_cast<unsigned short *>(WbT) + (int64_t)cc * ldh + hb;
        if (hb + 16 <= H) {
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
            for (int i = 0; i < 16; i++)
                if (hb + i < H) dst[i] = tile[16 * hc + i][cl];
        }
    }
}

// vectorised column sums of a bf16 matrix: out[c] += sign * sum_r A[r][c]; 8 columns (16 B) per thread
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const bf16 *__restrict__ A, int64_t lda, int rows, int cols,
                                                            float sign, float *__restrict__ out) {
    __shared__ float red[8][32][9];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < cols) {  // lda % 8 == 0 and padding columns are zero, so the 16-byte load is always in bounds
        for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) {
            const uint4 q = *reinterpret_cast<const uint4 *>(A + (int64_t)r * lda + c);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[2 * i] += __uint_as_float(w[i] << 16);
                acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) red[ty][tx][i] = acc[i];
    __syncthreads();
    if (ty == 0 && c < cols) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) s2 += red[j][tx][i];
            if (c + i < cols) atomicAdd(out + c + i, sign * s2);
        }
    }
}

// start of a step: G_a = cached data-side visible sums (or 0), G_b = 0 -- one launch instead of memcpy + memset
__global__ void init_bias_grads_kernel(float *__restrict__ Ga, const float *__restrict__ datasum, float scale, int n_vis,
                                       int n_total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) Ga[i] = (datasum && i < n_vis) ? scale * datasum[i] : 0.f;
}

// refresh both bf16 shadows from fp32 W (after qb_set_params)
__global__ void __launch_bounds__(256) shadow_kernel(const float *__restrict__ W, int H, int VL, int64_t ldv,
                                                     bf16 *__restrict__ Wb, bf16 *__restrict__ WbT, int64_t ldh) {
    __shared__ bf16 tile[64][64 + 2];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int c = blockIdx.x * 64 + tx, h0 = blockIdx.y * 64;
    for (int i = ty; i < 64; i += 4) {
        int h = h0 + i;
        float w = (h < H && c < VL) ? W[(int64_t)h * ldv + c] : 0.f;
        if (h < H && c < VL) Wb[(int64_t)h * ldv + c] = __float2bfloat16(w);
        tile[i][tx] = __float2bfloat16(w);
    }
    __syncthreads();
    const int h = h0 + tx;
    for (int i = ty; i < 64; i += 4) {
        int cc = blockIdx.x * 64 + i;
        if (cc < VL && h < H) WbT[(int64_t)cc * ldh + h] = tile[tx][i];
    }
}

// ----------------------------------------------------------------------------- conversions / transposes
template <typename T>
__global__ void convert_rows_kernel(const T *__restrict__ src, int64_t src_cols, int64_t rows, float *__restrict__ dstf,
                                    bf16 *__restrict__ dstb, int64_t ld, int64_t col0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * src_cols) return;
    int64_t r = i / src_cols, c = i % src_cols;
    float v = (float)src[i];
    if (dstf) dstf[r * ld + col0 + c] = v;
    if (dstb) dstb[r * ld + col0 + c] = __float2bfloat16(v);
}

// Same conversion, eight consecutive elements of a row per thread (src_cols % 8 == 0, col0 % 8 == 0, ld % 8 == 0, bases 16-byte
// aligned): one index division per eight elements, one 16-byte bf16 store (two float4 stores) -- the upload of a host mini-batch
// sits on the critical path of qb_*_step_host (its kernels cannot share an SM with the persistent GEMM CTAs).
// Blocks of 128 threads at <= 32 registers: ONE such block fits beside a persistent GEMM CTA (640 threads x 96 registers leave
// 4096 registers per SM), so the conversion of the next host batch proceeds on the copy stream WHILE the chain kernel runs --
// a fatter kernel would wait for it to end (measured: Int64 rows 5.5 -> 7.1 ms per step with 256-thread blocks).
// One-byte elements are the exception: their conversion is a few tens of microseconds when it has the machine to itself, and
// running beside the GEMM CTAs costs the chain kernel more than that (measured 4.75 vs 4.84 ms per end-to-end step), so they
// use 256-thread blocks that wait for free SMs.
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS) convert_rows8_kernel(const T *__restrict__ src, int64_t src_cols, int64_t rows, float *__restrict__ dstf,
                                                            bf16 *__restrict__ dstb, int64_t ld, int64_t col0) {
    const int64_t gpr = src_cols >> 3;   // groups of eight per row
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * gpr) return;
    const int64_t r = i / gpr, g = i - r * gpr;
    const T *s = src + r * src_cols + 8 * g;
    float v[8];
    if (sizeof(T) == 1) {
        const uint2 w = *reinterpret_cast<const uint2 *>(s);
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = (float)((w.x >> (8 * j)) & 0xFFu); v[4 + j] = (float)((w.y >> (8 * j)) & 0xFFu); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (float)s[j];
    }
    const int64_t o = r * ld + col0 + 8 * g;
    if (dstf) {
        reinterpret_cast<float4 *>(dstf + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(dstf + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (dstb) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<const uint32_t *>(&b);
        }
        *reinterpret_cast<uint4 *>(dstb + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

template <typename T>
__global__ void export_rows_kernel(const T *__restrict__ src, int64_t ld, int64_t col0, int64_t rows, int64_t cols,
                                   double *__restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols, c = i % cols;
    dst[i] = (double)(float)src[r * ld + col0 + c];
}

// dst[c*ldd + r] = src[r*lds + c]   (bf16; 32x32 tiles)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16 *__restrict__ src, int64_t lds, int rows, int cols,
                                                             bf16 *__restrict__ dst, int64_t ldd) {
    __shared__ bf16 tile[32][33];
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = ty; i < 32; i += 8) {
        int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[(int64_t)r * lds + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) dst[(int64_t)c * ldd + r] = tile[tx][i];
    }
}

// dst[c][r] = src[r][c] with 128-bit loads and 2 x 128-bit stores per task (64 x 128 tiles): the whole-matrix transpose of
// the sharded update path (32 MB in, 32 MB out at cfg5).  lds, ldd multiples of 8 elements, 16-byte aligned bases.
__global__ void __launch_bounds__(256) transpose_bf16_wide_kernel(const bf16 *__restrict__ src, int64_t lds, int rows, int cols,
                                                                  bf16 *__restrict__ dst, int64_t ldd) {
    constexpr int TR = 64, TC = 128, PITCH = TC + 8;
    __shared__ __align__(16) unsigned short tile[TR][PITCH];
    const int c0 = blockIdx.x * TC, r0 = blockIdx.y * TR;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int idx = threadIdx.x + 256 * q, r = idx >> 4, cv = (idx & 15) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < rows && c0 + cv < cols) v = *reinterpret_cast<const uint4 *>(src + (int64_t)(r0 + r) * lds + c0 + cv);  // row padding (ld % 8 == 0) is readable
        *reinterpret_cast<uint4 *>(&tile[r][cv]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int task = threadIdx.x + 256 * q;
        const int cl = task & (TC - 1), hc = task >> 7;
        const int cc = c0 + cl, hb = r0 + 16 * hc;
        if (cc >= cols || hb >= rows) continue;
        unsigned short *d = reinterpret_cast<unsigned short *>(dst) + (int64_t)cc * ldd + hb;
        if (hb + 16 <= rows) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; i++) pk[i] = (uint32_t)tile[16 * hc + 2 * i][cl] | ((uint32_t)tile[16 * hc + 2 * i + 1][cl] << 16);
            reinterpret_cast<uint4 *>(d)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4 *>(d)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
            for (int i = 0; i < 16; i++)
                if (hb + i < rows) d[i] = tile[16 * hc + i][cl];
        }
    }
}

// The same transpose as a template on the block size: THREADS = 128 at <= 32 registers (17 KB of shared memory) fits beside a
// persistent GEMM CTA, so the transposed copy of the NEXT host batch is built on the copy stream while the chain kernel runs.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS) transpose_bf16_wide_t_kernel(const bf16 *__restrict__ src, int64_t lds, int rows, int cols,
                                                                  bf16 *__restrict__ dst, int64_t ldd) {
    constexpr int TR = 64, TC = 128, PITCH = TC + 8;
    __shared__ __align__(16) unsigned short tile[TR][PITCH];
    const int c0 = blockIdx.x * TC, r0 = blockIdx.y * TR;
#pragma unroll
    for (int q = 0; q < 1024 / THREADS; q++) {
        const int idx = threadIdx.x + THREADS * q, r = idx >> 4, cv = (idx & 15) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < rows && c0 + cv < cols) v = *reinterpret_cast<const uint4 *>(src + (int64_t)(r0 + r) * lds + c0 + cv);  // row padding (ld % 8 == 0) is readable
        *reinterpret_cast<uint4 *>(&tile[r][cv]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 512 / THREADS; q++) {
        const int task = threadIdx.x + THREADS * q;
        const int cl = task & (TC - 1), hc = task >> 7;
        const int cc = c0 + cl, hb = r0 + 16 * hc;
        if (cc >= cols || hb >= rows) continue;
        unsigned short *d = reinterpret_cast<unsigned short *>(dst) + (int64_t)cc * ldd + hb;
        if (hb + 16 <= rows) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; i++) pk[i] = (uint32_t)tile[16 * hc + 2 * i][cl] | ((uint32_t)tile[16 * hc + 2 * i + 1][cl] << 16);
            reinterpret_cast<uint4 *>(d)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4 *>(d)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
            for (int i = 0; i < 16; i++)
                if (hb + i < rows) d[i] = tile[16 * hc + i][cl];
        }
    }
}

// Binary states as bits for the multi-GPU state exchange.  src: transposed bf16 states [units][ld] (0 / 1), `rows` a multiple
// of 8; bits[u][j] bit i = (src[u][8 j + i] != 0).
__global__ void pack_state_bits_kernel(const bf16 *__restrict__ src, int64_t ld, int units, int rows, uint8_t *__restrict__ bits) {
    const int bpr = rows >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)units * bpr) return;
    const int u = (int)(i / bpr), j = (int)(i % bpr);
    const uint4 v = *reinterpret_cast<const uint4 *>(src + (int64_t)u * ld + 8 * j);  // ld % 8 == 0: 16-byte aligned
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t b = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        b |= ((w[q] & 0xFFFFu) != 0u ? 1u : 0u) << (2 * q);
        b |= ((w[q] >> 16) != 0u ? 1u : 0u) << (2 * q + 1);
    }
    bits[i] = (uint8_t)b;
}

// bits_all: [world][units_total][rows_l / 8] (the all-gathered packs); dst[u][r * rows_l + 8 j + i] = bit ? 1 : 0 for the unit
// range [u0, u0 + units) -- the K-major gradient operand over the rows of ALL ranks
__global__ void unpack_state_bits_kernel(const uint8_t *__restrict__ bits_all, int world, int units_total, int u0, int units, int rows_l,
                                         bf16 *__restrict__ dst, int64_t ldd) {
    const int bpr = rows_l >> 3;
    const int64_t per_unit = (int64_t)world * bpr;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)units * per_unit) return;
    const int u = (int)(i / per_unit);
    const int rj = (int)(i % per_unit), r = rj / bpr, j = rj % bpr;
    const uint32_t b = bits_all[((int64_t)r * units_total + u0 + u) * bpr + j];
    const uint32_t one = 0x3F80u;  // bf16 1.0
    uint4 v;
    v.x = ((b & 1u) ? one : 0u) | ((b & 2u) ? one << 16 : 0u);
    v.y = ((b & 4u) ? one : 0u) | ((b & 8u) ? one << 16 : 0u);
    v.z = ((b & 16u) ? one : 0u) | ((b & 32u) ? one << 16 : 0u);
    v.w = ((b & 64u) ? one : 0u) | ((b & 128u) ? one << 16 : 0u);
    *reinterpret_cast<uint4 *>(dst + (int64_t)u * ldd + (int64_t)r * rows_l + 8 * j) = v;
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ src, bf16 *__restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(src[i]);
}
__global__ void bf16_to_f32_kernel(const bf16 *__restrict__ src, float *__restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}

// ----------------------------------------------------------------------------- metric reductions
// sum over rows x cols of (X - R)^2 -> double accumulator (block reduce, one atomic per block)
template <typename TX>
__global__ void __launch_bounds__(256) sqerr_kernel(const TX *__restrict__ X, int64_t ldx, const float *__restrict__ R,
                                                    int64_t ldr, int rows, int cols, double *__restrict__ out) {
    __shared__ float red[8];
    float acc = 0.f;
    int64_t n = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cols, c = i % cols;
        float d = (float)X[r * ldx + c] - R[r * ldr + c];
        acc = fmaf(d, d, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; i++) s += red[i];
        atomicAdd(out, (double)s);
    }
}

// free energy of each row: F = vis_term - sum_j softplus(pre_h[r][j]);  vis_term = -a.v - c.y (Bernoulli)
// or 0.5*||v - a||^2 - c.y (Gaussian).  The softplus sum comes either from fp32 pre-activations
// (fp32 path) or from per-row sums accumulated by the tensor-core epilogue.  One warp per row.
template <typename TX>
__global__ void __launch_bounds__(256) free_energy_kernel(const TX *__restrict__ X, int64_t ldx, const float *__restrict__ preh,
                                                          int64_t ldp, const float *__restrict__ softplus_rows,
                                                          const float *__restrict__ acat, int rows, int V, int ncols, int H,
                                                          int gaussian, double *__restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    float f = 0.f;
    for (int c = lane; c < ncols; c += 32) {
        float x = (float)X[(int64_t)warp * ldx + c], a = acat[c];
        if (gaussian && c < V) f += 0.5f * (x - a) * (x - a);
        else f -= a * x;
    }
    if (preh)
        for (int j = lane; j < H; j += 32) f -= softplus_f(preh[(int64_t)warp * ldp + j]);
    f = warp_sum(f);
    if (lane == 0) {
        if (softplus_rows) f -= softplus_rows[warp];
        atomicAdd(out, (double)f);
    }
}

// bf16-path label sampling: softmax over the L raw label pre-activations written by the visible
// GEMM epilogue, independent Bernoulli per unit, stores into both state orientations and folds
// the (negative) label bias gradient.
__global__ void label_bf16_kernel(const float *__restrict__ lab_pre, int64_t ld_lab, int rows, int V, int L,
                                  bf16 *__restrict__ state_rm, int64_t ld, bf16 *__restrict__ state_T, int64_t ldT,
                                  const float *__restrict__ inj, int64_t ldi, RngKey key, int64_t row0,
                                  float *__restrict__ bias_grad) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *p = lab_pre + (int64_t)r * ld_lab;
    float mx = -INFINITY;
    for (int l = 0; l < L; l++) mx = fmaxf(mx, p[l]);
    float s = 0.f;
    for (int l = 0; l < L; l++) s += expf(p[l] - mx);
    float lden = mx + logf(s);
    for (int l = 0; l < L; l++) {
        float q = expf(p[l] - lden);
        float rn = inj ? inj[(int64_t)r * ldi + V + l] : rng_uniform(key, (uint32_t)(V + l), (uint64_t)(row0 + r));
        float st = (rn < q) ? 1.f : 0.f;
        if (state_rm) state_rm[(int64_t)r * ld + V + l] = __float2bfloat16(st);
        if (state_T) state_T[(int64_t)(V + l) * ldT + r] = __float2bfloat16(st);
        if (bias_grad && st != 0.f) atomicAdd(bias_grad + V + l, -st);
    }
}

// conditional_prob_y_given_h (src/rbm.jl:196-205): row softmax of the L label pre-activations
__global__ void label_softmax_kernel(const float *__restrict__ pre, int64_t ld, int rows, int L, double *__restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *p = pre + (int64_t)r * ld;
    float mx = -INFINITY;
    for (int l = 0; l < L; l++) mx = fmaxf(mx, p[l]);
    float s = 0.f;
    for (int l = 0; l < L; l++) s += expf(p[l] - mx);
    float lden = mx + logf(s);
    for (int l = 0; l < L; l++) out[(int64_t)r * L + l] = (double)expf(p[l] - lden);
}

// _init_fantasy_data with the device stream: U[0,1) floats (src/fantasy_data.jl:66-86)
__global__ void uniform_fill_kernel(float *__restrict__ f, bf16 *__restrict__ b, int64_t ld, int64_t rows, int64_t units,
                                    RngKey key, int64_t row0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * units) return;
    int64_t r = i / units, u = i % units;
    float v = rng_uniform(key, (uint32_t)u, (uint64_t)(row0 + r));
    if (f) f[r * ld + u] = v;
    if (b) b[r * ld + u] = __float2bfloat16(v);
}

}  // namespace qb
