// Online (batch size 1) contrastive divergence: the reference's DEFAULT `train!(rbm, x, CD)` semantics
// (src/training/train_cd.jl:2-21): W changes after every sample, so samples are inherently sequential.
// Launching 2k+3 kernels per sample costs ~40 us/sample; here one cooperative persistent kernel walks
// the whole epoch with 2k+1 grid barriers per sample and everything else on chip:
//
//   * every CTA owns a fixed set of hidden rows of W1 = W[H][ldv] (for W'v) and of visible rows of the
//     fp32 transposed copy W2 = W^T[VL][ldh] (for W h): each matrix-vector product is computed entirely by
//     the owners (one warp per row, coalesced, shuffle reduction), no cross-CTA reduction;
//   * the four small vectors of a sample (h_data, h sample, v_model, h_model) are exchanged through
//     global memory, double-buffered by sample parity;
//   * the rank-2 update  W += lr (v h_d' - v~ h~')  (src/rbm.jl:138-150) is applied by the owners to both
//     layouts right away -- the next sample's products read only rows the CTA itself updated.
//
// Classifier models (src/training/train_cd.jl:23-43 -- the classifier DEFAULT): the label units are the last L of the V rows
// ([v | y] data rows, [W | U] columns, [a | c] biases).  The chain runs on the 3-argument conditional_prob_h, the visible
// half-step leaves the label PRE-activations in the exchange vector and every CTA finishes them identically (softmax over the
// L units, then independent Bernoulli draws: gibbs_sample_label, src/gibbs.jl:46-49), h_model comes from the 2-argument form
// that ignores the sampled labels (train_cd.jl:34), and label columns move with label_learning_rate (src/rbm.jl:101-118).
//
// fp32 arithmetic (a matrix-vector product is bandwidth/latency-bound: tensor cores do not apply).
// Random numbers follow the library contract with row = 0 and two draw ids per Gibbs step, i.e. exactly
// what a sequence of qb_cd_step(B = 1) calls consumes.
#pragma once
#include "common.cuh"

namespace qb {

struct OnlineCdParams {
    float *W1; long long ldv;     // [H][ldv]   (fp32 master, updated in place)
    float *W2; long long ldh;     // [V][ldh]   (fp32 transposed copy, updated in place)
    float *a, *b;                 // biases (updated in place)
    const float *Xf; const bf16 *Xb; long long ldx;  // dataset rows (one of the two is set)
    int V, H, n_samples, k, gaussian;   // V counts the label units too ([v | y] rows)
    int Vx;                             // visible units proper: rows / columns >= Vx are label units
    float lr, llr;
    float *vec;                   // scratch: 2 (parity) x [hd | hs | hm | vm]
    unsigned int *barrier;        // grid barrier counter (zeroed before launch)
    const float *injUh, *injUv;   // optional injected streams [k][n_samples][H] / [k][n_samples][V] (uniforms or normals)
    RngKey key0;                  // draw id of the first half-step
    unsigned long long *phase_ns; // [3] accumulated ns of (sample, gibbs, update) measured by block 0
};

__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int target = (++epoch) * nblocks;
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ RngKey key_plus(const RngKey &k0, unsigned long long add) {
    RngKey k = k0;
    unsigned long long d = ((unsigned long long)k0.draw_hi << 32 | k0.draw_lo) + add;
    k.draw_lo = (uint32_t)d; k.draw_hi = (uint32_t)(d >> 32) & 0x7FFFFFFFu;
    return k;
}

// out[j] = bias[j] + dot(Wrow_j, x) for the rows this CTA owns; one warp per row.  `cache` (optional): this
// warp's rows held in shared memory for the whole epoch, slot m = the m-th row the warp owns.
template <typename F>
__device__ __forceinline__ void owned_rows(int n_rows, const float *__restrict__ W, long long ld, int n_cols,
                                           const float *__restrict__ x_smem, const float *cache, F &&finish) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int m = 0;
    for (int r = blockIdx.x * wpb + warp; r < n_rows; r += gridDim.x * wpb, m++) {
        const float *w = cache ? cache + (long long)m * n_cols : W + (long long)r * ld;
        float acc = 0.f;
        for (int c = lane; c < n_cols; c += 32) acc = fmaf(w[c], x_smem[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) finish(r, acc);
    }
}

// same with a dot product over the first n_dot of the row's n_cols columns
template <typename F>
__device__ __forceinline__ void owned_rows_pitch(int n_rows, const float *__restrict__ W, long long ld, int n_cols, int n_dot,
                                                 const float *__restrict__ x_smem, const float *cache, F &&finish) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int m = 0;
    for (int r = blockIdx.x * wpb + warp; r < n_rows; r += gridDim.x * wpb, m++) {
        const float *w = cache ? cache + (long long)m * n_cols : W + (long long)r * ld;
        float acc = 0.f;
        for (int c = lane; c < n_dot; c += 32) acc = fmaf(w[c], x_smem[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) finish(r, acc);
    }
}

// CACHE: the rows a warp owns (of both layouts) live in shared memory for the whole epoch and are written back
// once at the end -- every product and update then runs at shared-memory latency (models up to a few MB).
template <bool CACHE>
__global__ void __launch_bounds__(1024) online_cd_kernel(const OnlineCdParams p) {
    extern __shared__ float sm[];
    float *sx = sm;              // data row            [V]
    float *sv = sx + p.V;        // current visible      [V]
    float *sh = sv + p.V;        // hidden vector in use [H]
    float *sh2 = sh + p.H;       // second hidden vector [H]
    const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31, wpb_ = blockDim.x >> 5;
    const int rpw1 = (p.H + gridDim.x * wpb_ - 1) / (gridDim.x * wpb_), rpw2 = (p.V + gridDim.x * wpb_ - 1) / (gridDim.x * wpb_);
    float *c1 = nullptr, *c2 = nullptr;  // this warp's cached rows of W1 ([rpw1][V]) and W2 ([rpw2][H])
    if (CACHE) {
        float *base = sh2 + p.H;
        c1 = base + (long long)warp_ * (rpw1 * p.V + rpw2 * p.H);
        c2 = c1 + rpw1 * p.V;
        int m = 0;
        for (int j = blockIdx.x * wpb_ + warp_; j < p.H; j += gridDim.x * wpb_, m++)
            for (int c = lane_; c < p.V; c += 32) c1[m * p.V + c] = p.W1[(long long)j * p.ldv + c];
        m = 0;
        for (int i = blockIdx.x * wpb_ + warp_; i < p.V; i += gridDim.x * wpb_, m++)
            for (int c = lane_; c < p.H; c += 32) c2[m * p.H + c] = p.W2[(long long)i * p.ldh + c];
        __syncwarp();
    }
    unsigned int epoch = 0;
    const unsigned int nb = gridDim.x;
    const int V = p.V, H = p.H;
    const int vstride = 3 * H + V;
    unsigned long long t0 = 0, acc_ns[3] = {0, 0, 0};
    const bool timer = (blockIdx.x == 0 && threadIdx.x == 0 && p.phase_ns != nullptr);

    for (int n = 0; n < p.n_samples; n++) {
        float *vec = p.vec + (n & 1) * vstride;
        float *g_hd = vec, *g_hs = vec + H, *g_hm = vec + 2 * H, *g_vm = vec + 3 * H;
        for (int i = threadIdx.x; i < V; i += blockDim.x)
            sx[i] = p.Xf ? p.Xf[(long long)n * p.ldx + i] : __bfloat162float(p.Xb[(long long)n * p.ldx + i]);
        __syncthreads();
        if (timer) t0 = gtimer();
        // ---- h_data = conditional_prob_h(v_data) and the first hidden sample (same probabilities)
        {
            const RngKey kh = key_plus(p.key0, (unsigned long long)n * 2ull * p.k);
            owned_rows(H, p.W1, p.ldv, V, sx, c1, [&](int j, float acc) {
                const float pr = sigmoid_exact(acc + p.b[j]);
                const float u = p.injUh ? p.injUh[(long long)n * H + j] : rng_uniform(kh, (uint32_t)j, 0ull);
                g_hd[j] = pr;
                g_hs[j] = (u < pr) ? 1.f : 0.f;
            });
        }
        grid_barrier(p.barrier, epoch, nb);
        if (timer) { unsigned long long t1 = gtimer(); acc_ns[0] += t1 - t0; t0 = t1; }
        for (int s = 0; s < p.k; s++) {
            for (int j = threadIdx.x; j < H; j += blockDim.x) sh[j] = g_hs[j];
            __syncthreads();
            // ---- gibbs_sample_visible: Bernoulli sample or Gaussian mean + N(0,1)
            const RngKey kv = key_plus(p.key0, (unsigned long long)n * 2ull * p.k + 2ull * s + 1ull);
            owned_rows(V, p.W2, p.ldh, H, sh, c2, [&](int i, float acc) {
                const float x = acc + p.a[i];
                const long long io = ((long long)s * p.n_samples + n) * V + i;
                if (i >= p.Vx) {
                    g_vm[i] = x;   // label unit: pre-activation c_l + U_l . h, finished after the barrier
                } else if (p.gaussian) {
                    const float z = p.injUv ? p.injUv[io] : rng_normal(kv, (uint32_t)i, 0ull);
                    g_vm[i] = x + z;
                } else {
                    const float pr = sigmoid_exact(x);
                    const float u = p.injUv ? p.injUv[io] : rng_uniform(kv, (uint32_t)i, 0ull);
                    g_vm[i] = (u < pr) ? 1.f : 0.f;
                }
            });
            grid_barrier(p.barrier, epoch, nb);
            for (int i = threadIdx.x; i < p.Vx; i += blockDim.x) sv[i] = g_vm[i];
            if (p.Vx < V) {
                // ---- gibbs_sample_label: every warp reduces the L pre-activations (same numbers in every CTA), then each thread
                // draws its label units -- p_l = exp(x_l - logsumexp(x)) (src/rbm.jl:196-205), y_l = 1[u_l < p_l]
                float mx = -INFINITY;
                for (int l = p.Vx + lane_; l < V; l += 32) mx = fmaxf(mx, g_vm[l]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float den = 0.f;
                for (int l = p.Vx + lane_; l < V; l += 32) den += expf(g_vm[l] - mx);
                den = warp_sum(den);
                const float lse = mx + logf(den);
                for (int l = p.Vx + threadIdx.x; l < V; l += blockDim.x) {
                    const float pr = expf(g_vm[l] - lse);
                    const float u = p.injUv ? p.injUv[((long long)s * p.n_samples + n) * V + l] : rng_uniform(kv, (uint32_t)l, 0ull);
                    sv[l] = (u < pr) ? 1.f : 0.f;
                }
            }
            __syncthreads();
            if (s + 1 < p.k) {
                // ---- next hidden sample of the chain
                const RngKey kh = key_plus(p.key0, (unsigned long long)n * 2ull * p.k + 2ull * (s + 1));
                owned_rows(H, p.W1, p.ldv, V, sv, c1, [&](int j, float acc) {
                    const float pr = sigmoid_exact(acc + p.b[j]);
                    const float u = p.injUh ? p.injUh[((long long)(s + 1) * p.n_samples + n) * H + j]
                                            : rng_uniform(kh, (uint32_t)j, 0ull);
                    g_hs[j] = (u < pr) ? 1.f : 0.f;
                });
                grid_barrier(p.barrier, epoch, nb);
            }
        }
        // ---- h_model = conditional_prob_h(v_model): probabilities
        // (2-argument form: only the Vx visible columns -- the sampled labels are ignored, train_cd.jl:34; a cached row keeps its
        // pitch V, the dot product just stops at Vx)
        owned_rows_pitch(H, p.W1, p.ldv, V, p.Vx, sv, c1, [&](int j, float acc) { g_hm[j] = sigmoid_exact(acc + p.b[j]); });
        grid_barrier(p.barrier, epoch, nb);
        if (timer) { unsigned long long t1 = gtimer(); acc_ns[1] += t1 - t0; t0 = t1; }
        // ---- update_rbm!(rbm, v_data, h_data, v_model, h_model, lr): owners update both layouts
        for (int j = threadIdx.x; j < H; j += blockDim.x) { sh[j] = g_hd[j]; sh2[j] = g_hm[j]; }
        __syncthreads();
        {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
            int m = 0;
            for (int j = blockIdx.x * wpb + warp; j < H; j += gridDim.x * wpb, m++) {
                float *w = CACHE ? c1 + (long long)m * V : p.W1 + (long long)j * p.ldv;
                const float hd = sh[j], hm = sh2[j];
                for (int c = lane; c < V; c += 32) w[c] += (c < p.Vx ? p.lr : p.llr) * (sx[c] * hd - sv[c] * hm);
                if (lane == 0) p.b[j] += p.lr * (hd - hm);
            }
            m = 0;
            for (int i = blockIdx.x * wpb + warp; i < V; i += gridDim.x * wpb, m++) {
                float *w = CACHE ? c2 + (long long)m * H : p.W2 + (long long)i * p.ldh;
                const float xd = sx[i], xm = sv[i];
                const float lr_i = i < p.Vx ? p.lr : p.llr;
                for (int c = lane; c < H; c += 32) w[c] += lr_i * (xd * sh[c] - xm * sh2[c]);
                if (lane == 0) p.a[i] += lr_i * (xd - xm);
            }
        }
        __syncthreads();  // sx / sv / sh are rewritten by the next sample
        if (timer) { unsigned long long t1 = gtimer(); acc_ns[2] += t1 - t0; }
    }
    if (CACHE) {  // write the cached rows back (W2 is scratch: only the master layout W1 matters afterwards)
        int m = 0;
        for (int j = blockIdx.x * wpb_ + warp_; j < p.H; j += gridDim.x * wpb_, m++)
            for (int c = lane_; c < p.V; c += 32) p.W1[(long long)j * p.ldv + c] = c1[m * p.V + c];
    }
    if (timer) { p.phase_ns[0] = acc_ns[0]; p.phase_ns[1] = acc_ns[1]; p.phase_ns[2] = acc_ns[2]; }
}

// ----------------------------------------------------------------------------- FastPCD (online part of a mini-batch)
// fast_persistent_contrastive_divergence! (src/training/train_fast_pcd.jl:16-49): for every sample of the
// mini-batch, in order: h_data from the CURRENT slow weights, slow update with lr/B against chain i, fast
// weights decayed by 19/20 and pushed with fast_lr/B -- one grid barrier per sample.  The chains advance
// afterwards (batched GEMMs on the effective weights), outside this kernel.
struct FastPcdParams {
    float *W1, *F1; long long ldv;   // slow / fast weights [H][ldv]
    float *W2, *F2; long long ldh;   // transposed copies [V][ldh]
    float *a, *b, *fa, *fb;          // slow / fast biases
    const float *Xf; const bf16 *Xb; long long ldx;      // the B data rows of this mini-batch
    const float *CVf; const bf16 *CVb; long long ldcv;   // chain visibles [B][ldcv]
    const float *CHf; const bf16 *CHb; long long ldch;   // chain hiddens  [B][ldch]
    int V, H, B;                     // V counts the label units too ([v | y] rows)
    int Vx;                          // visible units proper: columns >= Vx are label units (llr / fllr)
    int chain_stride;                // 1: sample n pairs with chain n; 0: every sample pairs with chain 0 (reference classifier)
    float lr, flr, llr, fllr;        // already divided by B
    float *vec; unsigned int *barrier;
};

template <bool CACHE>
__global__ void __launch_bounds__(1024) online_fastpcd_kernel(const FastPcdParams p) {
    extern __shared__ float sm[];
    const int V = p.V, H = p.H;
    float *sx = sm, *scv = sx + V, *shd = scv + V, *sch = shd + H;   // x_i, chain v_i, h_data, chain h_i
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int rpw1 = (H + gridDim.x * wpb - 1) / (gridDim.x * wpb), rpw2 = (V + gridDim.x * wpb - 1) / (gridDim.x * wpb);
    float *c1 = nullptr, *f1 = nullptr, *c2 = nullptr, *f2 = nullptr;
    if (CACHE) {
        float *base = sch + H + (long long)warp * 2 * (rpw1 * V + rpw2 * H);
        c1 = base; f1 = c1 + rpw1 * V; c2 = f1 + rpw1 * V; f2 = c2 + rpw2 * H;
        int m = 0;
        for (int j = blockIdx.x * wpb + warp; j < H; j += gridDim.x * wpb, m++)
            for (int c = lane; c < V; c += 32) { c1[m * V + c] = p.W1[(long long)j * p.ldv + c]; f1[m * V + c] = p.F1[(long long)j * p.ldv + c]; }
        m = 0;
        for (int i = blockIdx.x * wpb + warp; i < V; i += gridDim.x * wpb, m++)
            for (int c = lane; c < H; c += 32) { c2[m * H + c] = p.W2[(long long)i * p.ldh + c]; f2[m * H + c] = p.F2[(long long)i * p.ldh + c]; }
        __syncwarp();
    }
    unsigned int epoch = 0;
    const float decay = 19.0f / 20.0f;
    for (int n = 0; n < p.B; n++) {
        float *g_hd = p.vec + (n & 1) * H;
        const long long nc = (long long)n * p.chain_stride;
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
            sx[i] = p.Xf ? p.Xf[(long long)n * p.ldx + i] : __bfloat162float(p.Xb[(long long)n * p.ldx + i]);
            scv[i] = p.CVf ? p.CVf[nc * p.ldcv + i] : __bfloat162float(p.CVb[nc * p.ldcv + i]);
        }
        for (int j = threadIdx.x; j < H; j += blockDim.x)
            sch[j] = p.CHf ? p.CHf[nc * p.ldch + j] : __bfloat162float(p.CHb[nc * p.ldch + j]);
        __syncthreads();
        owned_rows(H, p.W1, p.ldv, V, sx, c1, [&](int j, float acc) { g_hd[j] = sigmoid_exact(acc + p.b[j]); });
        grid_barrier(p.barrier, epoch, gridDim.x);
        for (int j = threadIdx.x; j < H; j += blockDim.x) shd[j] = g_hd[j];
        __syncthreads();
        int m = 0;
        for (int j = blockIdx.x * wpb + warp; j < H; j += gridDim.x * wpb, m++) {
            float *w = CACHE ? c1 + (long long)m * V : p.W1 + (long long)j * p.ldv;
            float *f = CACHE ? f1 + (long long)m * V : p.F1 + (long long)j * p.ldv;
            const float hd = shd[j], hc = sch[j];
            for (int c = lane; c < V; c += 32) {
                const float d = sx[c] * hd - scv[c] * hc;
                w[c] += (c < p.Vx ? p.lr : p.llr) * d;
                f[c] = f[c] * decay + (c < p.Vx ? p.flr : p.fllr) * d;
            }
            if (lane == 0) { p.b[j] += p.lr * (hd - hc); p.fb[j] = p.fb[j] * decay + p.flr * (hd - hc); }
        }
        m = 0;
        for (int i = blockIdx.x * wpb + warp; i < V; i += gridDim.x * wpb, m++) {
            float *w = CACHE ? c2 + (long long)m * H : p.W2 + (long long)i * p.ldh;
            float *f = CACHE ? f2 + (long long)m * H : p.F2 + (long long)i * p.ldh;
            const float xd = sx[i], xc = scv[i];
            const float lr_i = i < p.Vx ? p.lr : p.llr, flr_i = i < p.Vx ? p.flr : p.fllr;
            for (int c = lane; c < H; c += 32) {
                const float d = xd * shd[c] - xc * sch[c];
                w[c] += lr_i * d;
                f[c] = f[c] * decay + flr_i * d;
            }
            if (lane == 0) { p.a[i] += lr_i * (xd - xc); p.fa[i] = p.fa[i] * decay + flr_i * (xd - xc); }
        }
        __syncthreads();
    }
    if (CACHE) {  // masters (and the transposed scratch copies, reused by the next mini-batch) go back to HBM
        int m = 0;
        for (int j = blockIdx.x * wpb + warp; j < H; j += gridDim.x * wpb, m++)
            for (int c = lane; c < V; c += 32) { p.W1[(long long)j * p.ldv + c] = c1[m * V + c]; p.F1[(long long)j * p.ldv + c] = f1[m * V + c]; }
        m = 0;
        for (int i = blockIdx.x * wpb + warp; i < V; i += gridDim.x * wpb, m++)
            for (int c = lane; c < H; c += 32) { p.W2[(long long)i * p.ldh + c] = c2[m * H + c]; p.F2[(long long)i * p.ldh + c] = f2[m * H + c]; }
    }
}

// effective parameters of the FastPCD chains: Pe = P + F over the flat vector [W | a | b]
__global__ void add_vectors_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// W2[v][h] = W1[h][v]  (fp32)
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float *__restrict__ src, long long lds, int rows, int cols,
                                                            float *__restrict__ dst, long long ldd) {
    __shared__ float tile[32][33];
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = ty; i < 32; i += 8) {
        int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[(long long)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) dst[(long long)c * ldd + r] = tile[tx][i];
    }
}

}  // namespace qb
