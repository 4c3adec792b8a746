/*
 * CPU oracle (C restatement) for the classical RBM training hot path of NITeQ/QARBoM.jl.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-check of the numpy restatement),
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Never
 * linked into, or called by, the product library.
 *
 * PARITY UNPINNED: Julia is not installable in this image and the reference's tests
 * hold no golden vectors (test/generative.jl:59-67), so this file restates the
 * reference's Float64, one-sample-at-a-time algorithm from its sources (paths relative
 * to /root/reference) and is cross-checked only against the independent numpy
 * restatement in qarbom_oracle.py.
 *
 * It deliberately keeps the reference's memory-traffic pattern (GEMV per conditional,
 * materialised V x H outer-product temporaries per sample, src/training/train_pcd.jl:30
 * and src/rbm.jl:146) because it doubles as the "reference CPU path" timing baseline.
 * `threads` > 1 parallelises the BLAS-like inner loops the way a threaded BLAS would;
 * the Julia code itself is single-threaded.
 *
 * Layout: W is column-major V x H (Julia Matrix{Float64}): W[i + V*j]; U is L x H:
 * U[l + L*j].  Datasets are N rows of V (resp. L) contiguous doubles (Vector{Vector}).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    const double *u; int64_t un, upos;
    const double *z; int64_t zn, zpos;
    uint64_t s[4]; int free_running; int err;
    int have_spare; double spare;
} qo_stream;

static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xo_next(qo_stream *st) {
    uint64_t *s = st->s;
    uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static void stream_init(qo_stream *st, const double *u, int64_t un, const double *z, int64_t zn, uint64_t seed) {
    memset(st, 0, sizeof(*st));
    st->u = u; st->un = un; st->z = z; st->zn = zn;
    st->free_running = (u == NULL && z == NULL);
    uint64_t x = seed + 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < 4; i++) { /* splitmix64 */
        uint64_t zz = (x += 0x9E3779B97F4A7C15ULL);
        zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ULL; zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBULL;
        st->s[i] = zz ^ (zz >> 31);
    }
}
static inline double s_rand(qo_stream *st) {
    if (st->free_running) return (double)(xo_next(st) >> 11) * (1.0 / 9007199254740992.0);
    if (st->upos >= st->un) { st->err = 1; return 0.5; }
    return st->u[st->upos++];
}
static inline double s_randn(qo_stream *st) {
    if (st->free_running) {
        if (st->have_spare) { st->have_spare = 0; return st->spare; }
        double u1 = 1.0 - s_rand(st), u2 = s_rand(st);
        double r = sqrt(-2.0 * log(u1)), t = 6.283185307179586 * u2;
        st->spare = r * sin(t); st->have_spare = 1;
        return r * cos(t);
    }
    if (st->zpos >= st->zn) { st->err = 1; return 0.0; }
    return st->z[st->zpos++];
}

typedef struct { int64_t V, H, L; int gaussian; double *W, *a, *b, *U, *c; int threads; } qo_model;

/* LogExpFunctions.logistic (Float64), used at src/rbm.jl:165,171,182 */
static inline double logistic(double x) {
    if (x < -744.4400719213812) return 0.0;
    if (x > 36.7368005696771) return 1.0;
    double e = exp(x);
    return e / (1.0 + e);
}

/* src/rbm.jl:165 / :170-171: logistic.(b .+ W' * v [.+ U' * y]) */
static void cond_prob_h(const qo_model *m, const double *v, const double *y, double *out) {
    const int64_t V = m->V, H = m->H, L = m->L;
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t j = 0; j < H; j++) {
        const double *w = m->W + V * j;
        double acc = 0.0;
        for (int64_t i = 0; i < V; i++) acc += w[i] * v[i];
        double pre = m->b[j] + acc;
        if (y) {
            const double *u = m->U + L * j;
            double accu = 0.0;
            for (int64_t l = 0; l < L; l++) accu += u[l] * y[l];
            pre += accu;
        }
        out[j] = logistic(pre);
    }
}

/* a .+ W * h  (shared by src/rbm.jl:182 and :187); column-major W*h = axpy over columns */
static void visible_mean(const qo_model *m, const double *h, double *out) {
    const int64_t V = m->V, H = m->H;
    const int64_t blk = 512; /* row blocks: each thread streams its slice of every column */
    const int64_t nblk = (V + blk - 1) / blk;
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t bi = 0; bi < nblk; bi++) {
        const int64_t i0 = bi * blk, i1 = (i0 + blk < V) ? i0 + blk : V;
        for (int64_t i = i0; i < i1; i++) out[i] = 0.0;
        for (int64_t j = 0; j < H; j++) {
            const double hj = h[j];
            const double *w = m->W + V * j;
            for (int64_t i = i0; i < i1; i++) out[i] += w[i] * hj;
        }
        for (int64_t i = i0; i < i1; i++) out[i] += m->a[i];
    }
}

/* src/rbm.jl:196-205: softmax(c + U h) through logsumexp */
static void cond_prob_y_given_h(const qo_model *m, const double *h, double *out) {
    const int64_t H = m->H, L = m->L;
    double mx = -INFINITY;
    for (int64_t l = 0; l < L; l++) {
        double acc = 0.0;
        for (int64_t j = 0; j < H; j++) acc += m->U[l + L * j] * h[j];
        out[l] = m->c[l] + acc;
        if (out[l] > mx) mx = out[l];
    }
    double s = 0.0;
    for (int64_t l = 0; l < L; l++) s += exp(out[l] - mx);
    double lden = mx + log(s);
    for (int64_t l = 0; l < L; l++) out[l] = exp(out[l] - lden);
}

/* src/gibbs.jl:3-6,29-32 */
static void sample_hidden(const qo_model *m, const double *v, const double *y, qo_stream *st, double *h, double *tmp) {
    cond_prob_h(m, v, y, tmp);
    for (int64_t j = 0; j < m->H; j++) h[j] = (s_rand(st) < tmp[j]) ? 1.0 : 0.0;
}
/* src/gibbs.jl:13-20 (+ src/rbm.jl:187 for Gaussian visibles: mean + 1.0*randn()) */
static void sample_visible(const qo_model *m, const double *h, qo_stream *st, double *v) {
    visible_mean(m, h, v);
    if (m->gaussian) { for (int64_t i = 0; i < m->V; i++) v[i] = v[i] + 1.0 * s_randn(st); }
    else { for (int64_t i = 0; i < m->V; i++) v[i] = (s_rand(st) < logistic(v[i])) ? 1.0 : 0.0; }
}
/* src/gibbs.jl:46-49: independent Bernoulli per label unit */
static void sample_label(const qo_model *m, const double *h, qo_stream *st, double *y) {
    cond_prob_y_given_h(m, h, y);
    for (int64_t l = 0; l < m->L; l++) y[l] = (s_rand(st) < y[l]) ? 1.0 : 0.0;
}

/* t = x*hx' .- w*hw'  with both products materialised first (src/rbm.jl:146, train_pcd.jl:30) */
static double *outer_diff(const qo_model *m, int64_t R, const double *x, const double *hx, const double *w, const double *hw) {
    const int64_t H = m->H;
    double *t1 = (double *)malloc(sizeof(double) * R * H), *t2 = (double *)malloc(sizeof(double) * R * H);
    double *t3 = (double *)malloc(sizeof(double) * R * H);
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t j = 0; j < H; j++) for (int64_t i = 0; i < R; i++) t1[i + R * j] = x[i] * hx[j];
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t j = 0; j < H; j++) for (int64_t i = 0; i < R; i++) t2[i + R * j] = w[i] * hw[j];
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t e = 0; e < R * H; e++) t3[e] = t1[e] - t2[e];
    free(t1); free(t2);
    return t3;
}
/* d = d + t  (non in-place `+=` on arrays allocates a fresh array, train_pcd.jl:30) */
static double *add_new(const qo_model *m, double *d, double *t, int64_t n) {
    double *r = (double *)malloc(sizeof(double) * n);
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t e = 0; e < n; e++) r[e] = d[e] + t[e];
    free(d); free(t);
    return r;
}
static void axpy_inplace(const qo_model *m, double *dst, const double *src, double s, int64_t n) {
    #pragma omp parallel for num_threads(m->threads) schedule(static) if (m->threads > 1)
    for (int64_t e = 0; e < n; e++) dst[e] += src[e] * s;
}

/* One chain, one Gibbs step: src/fantasy_data.jl:14-17 / :23-27 */
static void chain_step(const qo_model *m, double *cv, double *ch, double *cy, qo_stream *st, double *tmpH) {
    sample_hidden(m, cv, (m->L > 0) ? cy : NULL, st, ch, tmpH);
    sample_visible(m, ch, st, cv);
    if (m->L > 0) sample_label(m, ch, st, cy);
}

/*
 * persistent_contrastive_divergence!  (src/training/train_pcd.jl:2-43, classifier :46-95)
 * X: N x V, Y: N x L (or NULL), chains: B x V, B x H, B x L.  Processes mini-batches
 * [first_batch, first_batch + n_batches).  Returns 0, or 1 if an injected stream ran dry.
 */
int qo_pcd_epoch(int64_t V, int64_t H, int64_t L, int gaussian, double *W, double *a, double *b, double *U, double *c,
                 const double *X, const double *Y, int64_t N, int64_t B, int64_t Nc, int64_t k, double lr, double llr,
                 double *chv, double *chh, double *chy, const double *u, int64_t un, const double *z, int64_t zn,
                 uint64_t seed, int64_t first_batch, int64_t n_batches, int threads) {
    qo_model m = {V, H, L, gaussian, W, a, b, U, c, threads < 1 ? 1 : threads};
    qo_stream st; stream_init(&st, u, un, z, zn, seed);
    int64_t nb_total = (N - N % B) / B; /* src/utils.jl:13-26 */
    if (n_batches < 0 || first_batch + n_batches > nb_total) n_batches = nb_total - first_batch;
    if (Nc <= 0) Nc = B;
    /* Nc == B: the reference.  Nc != B: declared extension (see qarbom_oracle.py::_pcd_minibatch_more_chains):
     * dW = (Nc/B) sum_data - sum_chains, applied with lr / Nc. */
    const double scale = (double)Nc / (double)B;
    double *tmpH = (double *)malloc(sizeof(double) * H), *hd = (double *)malloc(sizeof(double) * H);
    double *zeroV = (double *)calloc(V > L ? V : L, sizeof(double)), *zeroH = (double *)calloc(H, sizeof(double));
    double *sx = (double *)malloc(sizeof(double) * (V > L ? V : L));
    for (int64_t mb = first_batch; mb < first_batch + n_batches; mb++) {
        if (L == 0) /* chains first: train_pcd.jl:14-16 */
            for (int64_t s = 0; s < k; s++)
                for (int64_t i = 0; i < Nc; i++) chain_step(&m, chv + i * V, chh + i * H, NULL, &st, tmpH);
        double *dW = (double *)calloc(V * H, sizeof(double)), *da = (double *)calloc(V, sizeof(double));
        double *db = (double *)calloc(H, sizeof(double));
        double *dU = L ? (double *)calloc(L * H, sizeof(double)) : NULL, *dc = L ? (double *)calloc(L, sizeof(double)) : NULL;
        if (Nc == B) {
            for (int64_t idx = 0; idx < B; idx++) {
                const double *vd = X + (mb * B + idx) * V;
                const double *yd = L ? Y + (mb * B + idx) * L : NULL;
                const double *cv = chv + idx * V, *ch = chh + idx * H, *cy = L ? chy + idx * L : NULL;
                cond_prob_h(&m, vd, yd, hd);
                dW = add_new(&m, dW, outer_diff(&m, V, vd, hd, cv, ch), V * H);
                if (L) dU = add_new(&m, dU, outer_diff(&m, L, yd, hd, cy, ch), L * H);
                for (int64_t i = 0; i < V; i++) da[i] += vd[i] - cv[i];
                for (int64_t j = 0; j < H; j++) db[j] += hd[j] - ch[j];
                for (int64_t l = 0; l < L; l++) dc[l] += yd[l] - cy[l];
            }
        } else {
            for (int64_t idx = 0; idx < B; idx++) { /* positive statistics, scaled by Nc / B */
                const double *vd = X + (mb * B + idx) * V;
                const double *yd = L ? Y + (mb * B + idx) * L : NULL;
                cond_prob_h(&m, vd, yd, hd);
                for (int64_t i = 0; i < V; i++) sx[i] = scale * vd[i];
                dW = add_new(&m, dW, outer_diff(&m, V, sx, hd, zeroV, zeroH), V * H);
                for (int64_t i = 0; i < V; i++) da[i] += sx[i];
                for (int64_t j = 0; j < H; j++) db[j] += scale * hd[j];
                if (L) {
                    for (int64_t l = 0; l < L; l++) sx[l] = scale * yd[l];
                    dU = add_new(&m, dU, outer_diff(&m, L, sx, hd, zeroV, zeroH), L * H);
                    for (int64_t l = 0; l < L; l++) dc[l] += sx[l];
                }
            }
            for (int64_t i2 = 0; i2 < Nc; i2++) { /* negative statistics over all chains */
                const double *cv = chv + i2 * V, *ch = chh + i2 * H, *cy = L ? chy + i2 * L : NULL;
                dW = add_new(&m, dW, outer_diff(&m, V, zeroV, zeroH, cv, ch), V * H);
                if (L) dU = add_new(&m, dU, outer_diff(&m, L, zeroV, zeroH, cy, ch), L * H);
                for (int64_t i = 0; i < V; i++) da[i] -= cv[i];
                for (int64_t j = 0; j < H; j++) db[j] -= ch[j];
                for (int64_t l = 0; l < L; l++) dc[l] -= cy[l];
            }
        }
        /* update_rbm!(rbm, dW, [dU,] da, db, [dc,] lr/B [, llr/B]): src/rbm.jl:120-136,152-163 */
        axpy_inplace(&m, W, dW, lr / (double)Nc, V * H);
        if (L) axpy_inplace(&m, U, dU, llr / (double)Nc, L * H);
        for (int64_t i = 0; i < V; i++) a[i] += da[i] * (lr / (double)Nc);
        for (int64_t j = 0; j < H; j++) b[j] += db[j] * (lr / (double)Nc);
        for (int64_t l = 0; l < L; l++) c[l] += dc[l] * (llr / (double)Nc);
        free(dW); free(da); free(db); free(dU); free(dc);
        if (L) /* classifier: chains AFTER the update, train_pcd.jl:89-92 */
            for (int64_t s = 0; s < k; s++)
                for (int64_t i = 0; i < Nc; i++) chain_step(&m, chv + i * V, chh + i * H, chy + i * L, &st, tmpH);
    }
    free(tmpH); free(hd); free(zeroV); free(zeroH); free(sx);
    return st.err;
}

/*
 * fast_persistent_contrastive_divergence!  (src/training/train_fast_pcd.jl:1-57, classifier :59-127): one epoch.
 * Inside a mini-batch the slow parameters are updated per SAMPLE with lr/B (update_rbm! per-sample form, src/rbm.jl:101-118,
 * 138-150) and the fast parameters are decayed by 19/20 and pushed by fast_lr/B times the same difference; they start at
 * zero every call.  The chains then advance with W + W_fast, ... (src/fantasy_data.jl:31-64); the fast gibbs_sample_visible
 * thresholds rand() < conditional_prob_v(...) for EVERY model (src/gibbs.jl:22-25), i.e. a Gaussian model's noisy value
 * mu + randn() too (src/rbm.jl:188-189).  pair_chains < 0: the reference (classifier: batch_index is never advanced, every
 * sample pairs with chain 1, :77-119; otherwise sample i with chain i); 0 / 1 force it.
 */
int qo_fast_pcd_epoch(int64_t V, int64_t H, int64_t L, int gaussian, double *W, double *a, double *b, double *U, double *c,
                      const double *X, const double *Y, int64_t N, int64_t B, int64_t k, double lr, double flr, double llr,
                      double fllr, double *chv, double *chh, double *chy, const double *u, int64_t un, const double *z,
                      int64_t zn, uint64_t seed, int pair_chains, int threads) {
    qo_model m = {V, H, L, gaussian, W, a, b, U, c, threads < 1 ? 1 : threads};
    qo_stream st; stream_init(&st, u, un, z, zn, seed);
    const int64_t nb = (N - N % B) / B;
    const int pair = pair_chains < 0 ? (L == 0) : (pair_chains != 0);
    const double decay = 19.0 / 20.0, s = lr / (double)B, fs = flr / (double)B, ls = llr / (double)B, fls = fllr / (double)B;
    double *Wf = (double *)calloc(V * H, sizeof(double)), *af = (double *)calloc(V, sizeof(double));
    double *bf = (double *)calloc(H, sizeof(double));
    double *Uf = L ? (double *)calloc(L * H, sizeof(double)) : NULL, *cf = L ? (double *)calloc(L, sizeof(double)) : NULL;
    double *We = (double *)malloc(sizeof(double) * V * H), *ae = (double *)malloc(sizeof(double) * V);
    double *be = (double *)malloc(sizeof(double) * H);
    double *Ue = L ? (double *)malloc(sizeof(double) * L * H) : NULL, *ce = L ? (double *)malloc(sizeof(double) * L) : NULL;
    double *hd = (double *)malloc(sizeof(double) * H), *tmpH = (double *)malloc(sizeof(double) * H);
    for (int64_t mb = 0; mb < nb; mb++) {
        for (int64_t idx = 0; idx < B; idx++) {
            const double *vd = X + (mb * B + idx) * V;
            const double *yd = L ? Y + (mb * B + idx) * L : NULL;
            const int64_t ci = pair ? idx : 0;
            const double *cv = chv + ci * V, *ch = chh + ci * H, *cy = L ? chy + ci * L : NULL;
            cond_prob_h(&m, vd, yd, hd);
            double *t = outer_diff(&m, V, vd, hd, cv, ch);
            for (int64_t e = 0; e < V * H; e++) { W[e] += s * t[e]; Wf[e] = Wf[e] * decay + fs * t[e]; }
            free(t);
            for (int64_t i = 0; i < V; i++) { a[i] += s * (vd[i] - cv[i]); af[i] = af[i] * decay + fs * (vd[i] - cv[i]); }
            if (L) {
                t = outer_diff(&m, L, yd, hd, cy, ch);
                for (int64_t e = 0; e < L * H; e++) { U[e] += ls * t[e]; Uf[e] = Uf[e] * decay + fls * t[e]; }
                free(t);
                for (int64_t l = 0; l < L; l++) { c[l] += ls * (yd[l] - cy[l]); cf[l] = cf[l] * decay + fls * (yd[l] - cy[l]); }
            }
            /* b last: hd was computed with the b of before this sample's update, like the reference */
            for (int64_t j = 0; j < H; j++) { b[j] += s * (hd[j] - ch[j]); bf[j] = bf[j] * decay + fs * (hd[j] - ch[j]); }
        }
        for (int64_t e = 0; e < V * H; e++) We[e] = W[e] + Wf[e];
        for (int64_t i = 0; i < V; i++) ae[i] = a[i] + af[i];
        for (int64_t j = 0; j < H; j++) be[j] = b[j] + bf[j];
        for (int64_t e = 0; e < L * H; e++) Ue[e] = U[e] + Uf[e];
        for (int64_t l = 0; l < L; l++) ce[l] = c[l] + cf[l];
        qo_model eff = {V, H, L, gaussian, We, ae, be, Ue, ce, m.threads};
        for (int64_t sN = 0; sN < k; sN++)
            for (int64_t i = 0; i < B; i++) {
                double *cv = chv + i * V, *ch = chh + i * H, *cy = L ? chy + i * L : NULL;
                sample_hidden(&eff, cv, cy, &st, ch, tmpH);
                visible_mean(&eff, ch, cv);
                for (int64_t q = 0; q < V; q++) cv[q] = gaussian ? cv[q] + 1.0 * s_randn(&st) : logistic(cv[q]);
                for (int64_t q = 0; q < V; q++) cv[q] = (s_rand(&st) < cv[q]) ? 1.0 : 0.0;
                if (L) sample_label(&eff, ch, &st, cy);
            }
    }
    free(Wf); free(af); free(bf); free(Uf); free(cf); free(We); free(ae); free(be); free(Ue); free(ce); free(hd); free(tmpH);
    return st.err;
}

/*
 * contrastive_divergence!  (src/training/train_cd.jl:2-21, classifier :23-43).
 * batch_size == 1 is the reference (online, in-place rank-2 update src/rbm.jl:101-118,
 * 138-150); batch_size > 1 is the declared mini-batch extension (see qarbom_oracle.py).
 */
int qo_cd_epoch(int64_t V, int64_t H, int64_t L, int gaussian, double *W, double *a, double *b, double *U, double *c,
                const double *X, const double *Y, int64_t N, int64_t B, int64_t k, double lr, double llr,
                const double *u, int64_t un, const double *z, int64_t zn, uint64_t seed,
                int64_t first_batch, int64_t n_batches, int threads) {
    qo_model m = {V, H, L, gaussian, W, a, b, U, c, threads < 1 ? 1 : threads};
    qo_stream st; stream_init(&st, u, un, z, zn, seed);
    int64_t nb_total = N / B;
    if (n_batches < 0 || first_batch + n_batches > nb_total) n_batches = nb_total - first_batch;
    double *hd = (double *)malloc(sizeof(double) * H), *hm = (double *)malloc(sizeof(double) * H);
    double *hs = (double *)malloc(sizeof(double) * H), *vm = (double *)malloc(sizeof(double) * V);
    double *ym = (double *)malloc(sizeof(double) * (L ? L : 1));
    for (int64_t mb = first_batch; mb < first_batch + n_batches; mb++) {
        double *dW = NULL, *dU = NULL, *da = NULL, *db = NULL, *dc = NULL;
        if (B > 1) {
            dW = (double *)calloc(V * H, sizeof(double)); da = (double *)calloc(V, sizeof(double));
            db = (double *)calloc(H, sizeof(double));
            if (L) { dU = (double *)calloc(L * H, sizeof(double)); dc = (double *)calloc(L, sizeof(double)); }
        }
        for (int64_t idx = 0; idx < B; idx++) {
            const double *vd = X + (mb * B + idx) * V;
            const double *yd = L ? Y + (mb * B + idx) * L : NULL;
            cond_prob_h(&m, vd, yd, hd);                         /* train_cd.jl:7 / :29 */
            memcpy(vm, vd, sizeof(double) * V);
            if (L) memcpy(ym, yd, sizeof(double) * L);
            for (int64_t s = 0; s < k; s++) {                    /* gibbs.jl:57-72 */
                sample_hidden(&m, vm, L ? ym : NULL, &st, hs, hm);
                sample_visible(&m, hs, &st, vm);
                if (L) sample_label(&m, hs, &st, ym);
            }
            cond_prob_h(&m, vm, NULL, hm);                       /* train_cd.jl:12 / :34 (2-arg) */
            double *t = outer_diff(&m, V, vd, hd, vm, hm);
            double *tu = L ? outer_diff(&m, L, yd, hd, ym, hm) : NULL;
            if (B == 1) {                                        /* src/rbm.jl:112-116,146-148 */
                axpy_inplace(&m, W, t, lr, V * H); free(t);
                for (int64_t i = 0; i < V; i++) a[i] += lr * (vd[i] - vm[i]);
                for (int64_t j = 0; j < H; j++) b[j] += lr * (hd[j] - hm[j]);
                if (L) { axpy_inplace(&m, U, tu, llr, L * H); free(tu);
                         for (int64_t l = 0; l < L; l++) c[l] += llr * (yd[l] - ym[l]); }
            } else {
                dW = add_new(&m, dW, t, V * H);
                if (L) dU = add_new(&m, dU, tu, L * H);
                for (int64_t i = 0; i < V; i++) da[i] += vd[i] - vm[i];
                for (int64_t j = 0; j < H; j++) db[j] += hd[j] - hm[j];
                for (int64_t l = 0; l < L; l++) dc[l] += yd[l] - ym[l];
            }
        }
        if (B > 1) {
            axpy_inplace(&m, W, dW, lr / (double)B, V * H);
            if (L) axpy_inplace(&m, U, dU, llr / (double)B, L * H);
            for (int64_t i = 0; i < V; i++) a[i] += da[i] * (lr / (double)B);
            for (int64_t j = 0; j < H; j++) b[j] += db[j] * (lr / (double)B);
            for (int64_t l = 0; l < L; l++) c[l] += dc[l] * (llr / (double)B);
            free(dW); free(da); free(db); free(dU); free(dc);
        }
    }
    free(hd); free(hm); free(hs); free(vm); free(ym);
    return st.err;
}

/* evaluate + _evaluate(MeanSquaredError) + reconstruct: src/evaluation.jl:16-20,52-70; src/rbm.jl:220-224 */
double qo_eval_mse(int64_t V, int64_t H, int gaussian, double *W, double *a, double *b, const double *X, int64_t N,
                   const double *z, int64_t zn, uint64_t seed, int threads) {
    qo_model m = {V, H, 0, gaussian, W, a, b, NULL, NULL, threads < 1 ? 1 : threads};
    qo_stream st; stream_init(&st, NULL, 0, z, zn, seed);
    if (z) st.free_running = 0;
    double *h = (double *)malloc(sizeof(double) * H), *v = (double *)malloc(sizeof(double) * V);
    double acc = 0.0;
    for (int64_t n = 0; n < N; n++) {
        const double *x = X + n * V;
        cond_prob_h(&m, x, NULL, h);
        visible_mean(&m, h, v);
        double s = 0.0;
        for (int64_t i = 0; i < V; i++) {
            double p = gaussian ? v[i] + 1.0 * s_randn(&st) : logistic(v[i]);
            s += (x[i] - p) * (x[i] - p);
        }
        acc += s / (double)N;
    }
    free(h); free(v);
    return acc;
}

/* mean free energy; sign convention from src/qubo.jl:43-55 (see qarbom_oracle.py::free_energy) */
double qo_mean_free_energy(int64_t V, int64_t H, int64_t L, int gaussian, double *W, double *a, double *b, double *U, double *c,
                           const double *X, const double *Y, int64_t N) {
    double tot = 0.0;
    for (int64_t n = 0; n < N; n++) {
        const double *x = X + n * V; const double *y = (L && Y) ? Y + n * L : NULL;
        double f = 0.0;
        for (int64_t i = 0; i < V; i++) f += gaussian ? 0.5 * (x[i] - a[i]) * (x[i] - a[i]) : -a[i] * x[i];
        if (y) for (int64_t l = 0; l < L; l++) f -= c[l] * y[l];
        for (int64_t j = 0; j < H; j++) {
            double pre = b[j];
            for (int64_t i = 0; i < V; i++) pre += W[i + V * j] * x[i];
            if (y) for (int64_t l = 0; l < L; l++) pre += U[l + L * j] * y[l];
            f -= (pre > 0 ? pre + log1p(exp(-pre)) : log1p(exp(pre)));
        }
        tot += f;
    }
    return tot / (double)N;
}

int qo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
