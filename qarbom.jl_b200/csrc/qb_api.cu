// C ABI (include/qarbom_b200.h): handle, HBM layout, and the host-side orchestration of the
// CD-k / PCD mini-batch step on top of the kernels in simt_kernels.cuh (QB_PREC_FP32) and
// sm100_gemm.cuh (QB_PREC_BF16).  No CPU compute path exists in this library.
#include "../../include/qarbom_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <climits>
#include <cstring>
#include <memory>
#include <vector>

#include "common.cuh"
#include "online_kernels.cuh"
#include "p2p_kernels.cuh"
#include "nvls_setup.h"
#include "simt_kernels.cuh"
#include "sm100_gemm.cuh"
#include "sm100_chain.cuh"
#include "sm100_epoch.cuh"

using namespace qb;

static thread_local std::string g_last_error;

// ----------------------------------------------------------------------------- NCCL (dlopen'ed, optional)
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;  // optional (NCCL >= 2.18)
    int (*MemAlloc)(void **, size_t) = nullptr;                              // optional (NCCL >= 2.19): registered user buffers
    int (*MemFree)(void *) = nullptr;
    int (*CommRegister)(ncclComm_t, void *, size_t, void **) = nullptr;
    int (*CommDeregister)(ncclComm_t, void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
Nccl &nccl() {
    static Nccl n;
    if (!n.lib) {
        // QB_NCCL_LIB: an explicit library (the host layer points it at the nvidia-nccl wheel torch uses, whose version matches
        // the device-API headers nvls_setup.cu was built against); else whatever "libnccl.so.2" resolves to -- the copy already
        // loaded into the process (torch.distributed) wins by SONAME
        const char *names[] = {getenv("QB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        QB_CHECK(n.lib != nullptr, QB_E_NCCL, "libnccl.so.2 not found (dlopen)");
        n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.lib, "ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.lib, "ncclCommInitRank");
        n.AllReduce = (decltype(n.AllReduce))dlsym(n.lib, "ncclAllReduce");
        n.ReduceScatter = (decltype(n.ReduceScatter))dlsym(n.lib, "ncclReduceScatter");
        n.AllGather = (decltype(n.AllGather))dlsym(n.lib, "ncclAllGather");
        n.CommSplit = (decltype(n.CommSplit))dlsym(n.lib, "ncclCommSplit");
        n.MemAlloc = (decltype(n.MemAlloc))dlsym(n.lib, "ncclMemAlloc");
        n.MemFree = (decltype(n.MemFree))dlsym(n.lib, "ncclMemFree");
        n.CommRegister = (decltype(n.CommRegister))dlsym(n.lib, "ncclCommRegister");
        n.CommDeregister = (decltype(n.CommDeregister))dlsym(n.lib, "ncclCommDeregister");
        n.GroupStart = (decltype(n.GroupStart))dlsym(n.lib, "ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.lib, "ncclGroupEnd");
        n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.lib, "ncclCommDestroy");
        n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.lib, "ncclGetErrorString");
        QB_CHECK(n.GetUniqueId && n.CommInitRank && n.AllReduce && n.CommDestroy && n.ReduceScatter && n.AllGather && n.GroupStart &&
                     n.GroupEnd, QB_E_NCCL, "NCCL symbols missing");
    }
    return n;
}
// ncclConfig_t as of NCCL 2.18 (size / magic / version let newer libraries accept the shorter struct)
struct NcclConfig218 {
    size_t size; unsigned int magic; unsigned int version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs; const char *netName; int splitShare;
};
#define QB_NCCL(expr)                                                                          \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) throw qb::Error(QB_E_NCCL, std::string(#expr) + " failed: " +             \
                                     (nccl().GetErrorString ? nccl().GetErrorString(_r) : "?")); \
    } while (0)
}  // namespace

struct HalfCapture { qb::sm100::Operand A, B; qb::sm100::EpiParams ep; double flops; bool set; };

// ----------------------------------------------------------------------------- device matrices
constexpr int QB_OVERLAP_CTAS = 16;

struct Mat {
    float *f = nullptr;  // fp32 [rows][ld]        (QB_PREC_FP32)
    bf16 *b = nullptr;   // bf16 [rows][ld]        (QB_PREC_BF16)
    bf16 *bT = nullptr;  // bf16 [units][ldT]      (QB_PREC_BF16, gradient operand)
    int64_t ld = 0, ldT = 0, rows = 0, units = 0;
    bool exact = false;  // every element is known to be exactly representable in bf16 (sampled 0/1 states, UInt8 data)
    Mat row_slice(int64_t r0, int64_t n) const {  // rows [r0, r0+n); the transposed view is not sliceable
        Mat m = *this;
        if (m.f) m.f += r0 * ld;
        if (m.b) m.b += r0 * ld;
        m.bT = nullptr;
        m.rows = n;
        return m;
    }
};

struct Rng {  // how a sampling half-step gets its random numbers
    const float *inj = nullptr; int64_t ld_inj = 0;  // injected [rows][units]
    RngKey key{}; int use_rng = 0;
};

struct qb_handle {
    qb_config cfg{};
    int64_t V = 0, H = 0, L = 0, VL = 0, ldv = 0, ldh = 0, Bmax = 0, ldb = 0;
    bool bf = false, gaussian = false;
    // QB_PREC_FP32X3: the fp32 path with its contractions on the tensor cores (operands split into three bf16 components)
    bool x3 = false;
    bf16 *W3[3] = {nullptr, nullptr, nullptr}, *W3T[3] = {nullptr, nullptr, nullptr}; bool w3_dirty = true;   // components of W / its transpose
    bf16 *x3_buf[2][3] = {}; size_t x3_cap[2] = {0, 0};   // components of the operands split on the fly
    bool gauss_threshold = false;  // FastPCD chain advance on Gaussian visibles: v = 1[u < mean + z] (src/gibbs.jl:22-25)
    bool fast_pcd_pair_chains = false;  // classifier FastPCD: pair sample i with chain i instead of the reference's chain 1
    int num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[8] = {};
    // flat parameter / gradient / velocity vectors: [W (H x ldv) | acat (ldv) | b (ldh)]
    float *P = nullptr, *G = nullptr, *Vel = nullptr;
    float *P_best = nullptr;
    // online (B = 1) CD kernel: fp32 transposed weights, exchange vectors, grid barrier, phase timers
    float *W2 = nullptr, *online_vec = nullptr; unsigned int *online_bar = nullptr; unsigned long long *online_ns = nullptr;
    bool online_kernel = true; int online_wpb = 0;
    float *F = nullptr, *F2 = nullptr, *P_eff = nullptr;  // FastPCD: fast weights [W | a | b], their transpose, effective parameters  // device-side best-model snapshot (copy_rbm / copy_rbm! without a host round trip)
    int64_t nW = 0, nP = 0;
    bf16 *Wb = nullptr, *WbT = nullptr;
    double momentum = 0.0, weight_decay = 0.0;
    // "stale_chains" mode (declared extension, bf16 PCD only): the allreduce + update of step t run on
    // comm_stream while the chains of step t+1 advance with the previous weight version.  Second
    // shadow / bias / gradient buffers make the two versions coexist.
    bool stale_chains = false, pending_update = false;
    bf16 *Wb_alt = nullptr, *WbT_alt = nullptr;
    float *G_alt = nullptr, *bias_buf[2] = {nullptr, nullptr};
    float *bias_rd = nullptr;  // [acat | b] as read by the GEMM epilogues (P's own bias region unless stale mode)
    int bias_cur = 0;
    cudaStream_t comm_stream = nullptr;
    // overlap needs SMs: the exchange of the stale mode runs on a communicator limited to QB_OVERLAP_CTAS CTAs while the
    // GEMMs leave that many SMs free (their persistent CTAs own a whole SM each, NCCL could not start beside them)
    ncclComm_t comm_overlap = nullptr;
    int gemm_sms = 148;
    // exchange buffers re-homed into ncclMemAlloc memory and registered with the communicator (zero-copy / NVLS paths)
    std::vector<void *> nccl_mem, nccl_reg; bool reg_tried = false;
    // "state_exchange" (multi-GPU PCD, bf16, binary model without labels; exact semantics): the positive product is
    // reduce-scattered BEHIND the chain advance, and instead of reducing the negative product the ranks all-gather their
    // chain states as BITS (1/16 of bf16) and each computes its own shard of the full negative product.
    bool state_exchange = false;
    uint8_t *bits_loc = nullptr, *bits_all = nullptr; bf16 *cvT_all = nullptr, *chT_all = nullptr; int64_t sx_rows = 0;
    cudaEvent_t ev_grad = nullptr, ev_upd = nullptr;
    // resident dataset
    Mat data; int64_t N = 0;
    bf16 *dataT = nullptr; int64_t dataT_B = 0, dataT_nb = 0, dataT_ld = 0;
    float *datasum = nullptr;          // [nb][ldv] column sums of every resident mini-batch (data side of delta_a)
    const float *cur_datasum = nullptr;  // sums of the mini-batch being processed (nullptr: compute them)
    // qb_pcd_step_host: the batch reconstruction error is taken inside the step, from the positive-phase p(h | x) that the
    // step computes anyway (one GEMM instead of two; weights before the update).  Binary models without label units.
    bool step_recon_wanted = false, step_recon_done = false;
    bool bias_grads_clean = false;       // the bias region of G is zero (left so by the previous update)
    bool datasum_in_update = false;      // the cached data sums are added by the update kernel, not by the init kernel
    float datasum_scale = 1.f;
    // persistent chains and work buffers (Bmax rows)
    Mat cv, ch; int64_t n_chains = 0; bool chains_binary = false;
    Mat phd, hs, vm, phm, hb;  // pos-phase probs, sampled hidden, CD model visible, CD final probs, host-batch staging
    // host-batch prefetch (qb_set_next_host_batch): second staging matrix filled on copy_stream while a step runs
    Mat hb2; void *staging2 = nullptr; size_t staging2_bytes = 0;
    const void *next_X = nullptr, *next_Y = nullptr; int next_xdt = 0, next_ydt = 0;   // hint for the next step_host call
    const void *pf_X = nullptr, *pf_Y = nullptr; int pf_xdt = 0, pf_ydt = 0; int64_t pf_B = 0; bool pf_valid = false; cudaEvent_t ev_pf = nullptr;
    float *pre_h = nullptr, *pre_v = nullptr, *lab_pre = nullptr, *softplus_rows = nullptr;
    float *out_f32 = nullptr;  // fp32 [Bmax][max(ldv, ldh)] API output staging
    float *znoise = nullptr;   // fp32 [Bmax][V] injected reconstruct noise (Gaussian visibles)
    float *rand_rows = nullptr;  // fp32 [Bmax][max(VL, H)] caller-supplied random numbers of qb_sample_hidden / qb_sample_visible
    double *dscal = nullptr;   // device scalar accumulators [4]
    void *staging = nullptr; size_t staging_bytes = 0;
    // injected streams for the next step
    float *injUh = nullptr, *injUv = nullptr, *injZv = nullptr; int64_t inj_k = 0, inj_B = 0; bool inj_active = false;
    size_t injUh_cap = 0, injUv_cap = 0, injZv_cap = 0;
    uint64_t seed = 0, draw = 0;
    ncclComm_t comm = nullptr;
    // NVLink peer-memory exchange fused with the update (p2p_kernels.cuh): peers' G / shadows / flags / masters
    bool p2p = false;
    float *peerG[QB_MAX_PEERS] = {}, *peerP[QB_MAX_PEERS] = {};
    bf16 *peerWb[QB_MAX_PEERS] = {}, *peerWbT[QB_MAX_PEERS] = {};
    unsigned int *peerFlags[QB_MAX_PEERS] = {}, *p2p_flags = nullptr, *p2p_grid_bar = nullptr;
    unsigned int p2p_epoch = 1; unsigned long long *p2p_ns = nullptr; int p2p_blocks_per_sm = 3;
    unsigned int *p2p_timeout_host = nullptr, *p2p_timeout_dev = nullptr;
    int64_t shard_rows = 0;
    // NVLS variant of that kernel (option "nvls_exchange"): G, the row-major shadow and the barrier flags live in NCCL
    // symmetric windows; the kernel reduces through the switch (multimem.ld_reduce) and pushes through it (multimem.st)
    bool nvls = false; void *nvls_state = nullptr; const float *G_mc = nullptr; bf16 *Wb_mc = nullptr;
    // split exchange (binary PCD on the NVLS path): uint16 negative products in a fourth window; option "split_exchange"
    // 0 never, 1 automatic (half-steps of at most one wave minus the SMs the background reduction needs), 2 whenever possible
    unsigned short *Gn16 = nullptr; const unsigned short *Gn16_mc = nullptr; int split_exchange_mode = 0;
    // NCCL variant of the sharded update (option "sharded_update", bf16, H % world == 0): reduce-scatter of the fp32
    // gradient, update of the owned hidden rows only, all-gather of the bf16 shadow, local transpose.  Moves 3/4 of the
    // allreduce bytes and does 1/world of the update work; the fp32 master is then sharded like in the p2p path.
    bool sharded_update = false, master_sharded = false;
    bool exchange_done = false;  // the step already reduced its gradient shard and the bias gradients (state_exchange)
    int64_t launches = 0, gemm_launches = 0;
    int64_t gemm_hist[QB_GEMM_HIST_SLOTS] = {};
    double gemm_flops = 0.0, gemm_issued_flops = 0.0;   // algorithmic / extra component-product flops (QB_PREC_FP32X3)
    int force_bn = 0, use_2cta = 1;
    // fused chain launches (sm100_chain.cuh): half-steps recorded between chain_begin / chain_end run as ONE persistent kernel
    int chain_pairs = 0;                            // option "chain_pairs" (experiments): CTA pairs of the chain kernel (0 = all)
    int fused_chain_mode = 1; int chain_bn = 0;     // options "fused_chain" (0 never, 1 automatic, 2 always), "chain_bn" (0 = automatic)
    int chain_quad_mode = 0;                        // option "chain_quad": 4-CTA-cluster chain kernel with multicast weights (0 never -- the default: measured slower --, 1 one-wave shapes, 2 whenever possible)
    int rec_depth = 0; sm100::ChainBuilder *rec = nullptr;
    unsigned int *chain_cnt = nullptr; int64_t chain_cnt_region = 0, chain_launch_no = 0;
    unsigned int *chain_err_host = nullptr, *chain_err_dev = nullptr;
    int64_t chain_launches = 0;
    // persistent multi-mini-batch kernel (sm100_epoch.cuh): option "epoch_kernel" 0 never, 1 automatic (small models), 2 always
    int epoch_kernel_mode = 1; sm100::EpochBuilder *epoch = nullptr;
    unsigned int *epoch_cnt = nullptr; size_t epoch_cnt_cap = 0;
    struct HalfCapture *capture = nullptr;   // describe-only mode of op_hidden / op_visible (program builders)
    int64_t epoch_launches = 0;
    bool async_steps = false;   // step calls return without a stream sync (qb_sync to wait)
    bool time_gemms = false;    // bracket every GEMM launch with CUDA events on the launching stream
    std::vector<cudaEvent_t> gemm_ev; size_t gemm_ev_used = 0; int64_t gemm_timed = 0; double gemm_timed_flops = 0.0;
    cudaEvent_t timer_ev[2] = {};
    std::vector<cudaEvent_t> epoch_ev;   // 5 per mini-batch step of a timed epoch, resolved after the epoch
    std::vector<cudaEvent_t> upd_ev; size_t upd_ev_used = 0; int64_t upd_timed = 0;
};

namespace {

template <typename T>
T *dalloc(size_t n) {
    T *p = nullptr;
    QB_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    QB_CUDA(cudaMemset(p, 0, (n ? n : 1) * sizeof(T)));
    return p;
}

Mat alloc_mat(const qb_handle *h, int64_t rows, int64_t units, bool want_T) {
    Mat m;
    m.rows = rows; m.units = units;
    m.ld = round_up(units, 8);
    m.ldT = round_up(rows, 8);
    if (h->bf) {
        m.b = dalloc<bf16>(rows * m.ld);
        if (want_T) m.bT = dalloc<bf16>(units * m.ldT);
    } else {
        m.f = dalloc<float>(rows * m.ld);
    }
    return m;
}
void free_mat(Mat &m) {
    cudaFree(m.f); cudaFree(m.b); cudaFree(m.bT);
    m = Mat();
}

inline dim3 grid1d(int64_t n, int block = 256) { return dim3((unsigned)ceil_div(n, block)); }

float *Wf(qb_handle *h) { return h->P; }
float *acat(qb_handle *h) { return h->P + h->nW; }
float *bh(qb_handle *h) { return h->P + h->nW + h->ldv; }
float *bias_a(qb_handle *h) { return h->bias_rd ? h->bias_rd : acat(h); }
float *bias_b(qb_handle *h) { return h->bias_rd ? h->bias_rd + h->ldv : bh(h); }
float *G_W(qb_handle *h) { return h->G; }
float *G_a(qb_handle *h) { return h->G + h->nW; }
float *G_b(qb_handle *h) { return h->G + h->nW + h->ldv; }

void ensure_staging(qb_handle *h, size_t bytes) {
    if (h->staging_bytes >= bytes) return;
    cudaFree(h->staging);
    h->staging = nullptr; h->staging_bytes = 0;
    QB_CUDA(cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
}

size_t dt_size(int dt) {
    switch (dt) {
        case QB_DT_F64: return 8;
        case QB_DT_I64: return 8;
        case QB_DT_F32: return 4;
        case QB_DT_U8: return 1;
    }
    throw Error(QB_E_ARG, "unknown element type code");
}

// host rows (n x cols of dtype) -> device matrix columns [col0, col0+cols) of dst rows [r0, r0+n)
void upload_rows_on(qb_handle *h, cudaStream_t st, void *&staging, size_t &staging_bytes, const void *src, int dt, int64_t n,
                    int64_t cols, Mat &dst, int64_t r0, int64_t col0);

void upload_rows(qb_handle *h, const void *src, int dt, int64_t n, int64_t cols, Mat &dst, int64_t r0, int64_t col0) {
    upload_rows_on(h, h->stream, h->staging, h->staging_bytes, src, dt, n, cols, dst, r0, col0);
}

void upload_rows_on(qb_handle *h, cudaStream_t st, void *&staging, size_t &staging_bytes, const void *src, int dt, int64_t n,
                    int64_t cols, Mat &dst, int64_t r0, int64_t col0) {
    if (n == 0 || cols == 0) return;
    // UInt8 values are exact in bf16; anything else may not be (a later column block can only take exactness away)
    dst.exact = (dt == QB_DT_U8) && (col0 == 0 || dst.exact);
    const size_t es = dt_size(dt);
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)((64u << 20) / (cols * es)));
    for (int64_t r = 0; r < n; r += chunk_rows) {
        const int64_t nr = std::min(chunk_rows, n - r);
        if (staging_bytes < (size_t)nr * cols * es) {
            cudaFree(staging); staging = nullptr; staging_bytes = 0;
            QB_CUDA(cudaMalloc(&staging, (size_t)nr * cols * es));
            staging_bytes = (size_t)nr * cols * es;
        }
        QB_CUDA(cudaMemcpyAsync(staging, (const char *)src + (size_t)r * cols * es, (size_t)nr * cols * es,
                                cudaMemcpyHostToDevice, st));
        float *df = dst.f ? dst.f + (r0 + r) * dst.ld : nullptr;
        bf16 *db = dst.b ? dst.b + (r0 + r) * dst.ld : nullptr;
        const int64_t tot = nr * cols;
        if ((cols % 8) == 0 && (col0 % 8) == 0 && (dst.ld % 8) == 0) {   // eight elements per thread, 16-byte stores
            const int64_t groups = tot / 8;
            switch (dt) {
                case QB_DT_F64: convert_rows8_kernel<double, 128><<<grid1d(groups, 128), 128, 0, st>>>((const double *)staging, cols, nr, df, db, dst.ld, col0); break;
                case QB_DT_I64: convert_rows8_kernel<long long, 128><<<grid1d(groups, 128), 128, 0, st>>>((const long long *)staging, cols, nr, df, db, dst.ld, col0); break;
                case QB_DT_F32: convert_rows8_kernel<float, 128><<<grid1d(groups, 128), 128, 0, st>>>((const float *)staging, cols, nr, df, db, dst.ld, col0); break;
                default: convert_rows8_kernel<unsigned char, 256><<<grid1d(groups, 256), 256, 0, st>>>((const unsigned char *)staging, cols, nr, df, db, dst.ld, col0); break;
            }
        } else switch (dt) {
            case QB_DT_F64: convert_rows_kernel<double><<<grid1d(tot), 256, 0, st>>>((const double *)staging, cols, nr, df, db, dst.ld, col0); break;
            case QB_DT_I64: convert_rows_kernel<long long><<<grid1d(tot), 256, 0, st>>>((const long long *)staging, cols, nr, df, db, dst.ld, col0); break;
            case QB_DT_F32: convert_rows_kernel<float><<<grid1d(tot), 256, 0, st>>>((const float *)staging, cols, nr, df, db, dst.ld, col0); break;
            default: convert_rows_kernel<unsigned char><<<grid1d(tot), 256, 0, st>>>((const unsigned char *)staging, cols, nr, df, db, dst.ld, col0); break;
        }
        QB_CUDA(cudaGetLastError());
        h->launches++;
        // no sync: the next copy into the staging buffer is stream-ordered after this kernel, and
        // cudaMemcpyAsync from pageable memory returns only once the source has been staged
    }
}

// device matrix columns -> host doubles (n x cols)
void download_rows(qb_handle *h, const Mat &src, int64_t r0, int64_t col0, int64_t n, int64_t cols, double *dst) {
    if (!dst || n == 0 || cols == 0) return;
    ensure_staging(h, (size_t)n * cols * sizeof(double));
    if (src.f) export_rows_kernel<float><<<grid1d(n * cols), 256, 0, h->stream>>>(src.f + r0 * src.ld, src.ld, col0, n, cols, (double *)h->staging);
    else export_rows_kernel<bf16><<<grid1d(n * cols), 256, 0, h->stream>>>(src.b + r0 * src.ld, src.ld, col0, n, cols, (double *)h->staging);
    QB_CUDA(cudaGetLastError());
    h->launches++;
    QB_CUDA(cudaMemcpyAsync(dst, h->staging, (size_t)n * cols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QB_CUDA(cudaStreamSynchronize(h->stream));
}

// dst[c][r] = src[r][c] on stream st: 128-bit loads / stores when the leading dimensions and bases allow (they do for every
// buffer this library allocates: ld % 8 == 0, 256-byte aligned bases; row slices keep 16-byte alignment)
void transpose_on(qb_handle *h, cudaStream_t st, const bf16 *src, int64_t lds, int64_t rows, int64_t cols, bf16 *dst, int64_t ldd) {
    if ((lds % 8) == 0 && (ldd % 8) == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0) {
        dim3 g((unsigned)ceil_div(cols, 128), (unsigned)ceil_div(rows, 64));
        if (st == h->copy_stream && h->next_xdt != QB_DT_U8)   // wide element types: keep the copy stream moving beside the GEMM CTAs
            transpose_bf16_wide_t_kernel<128><<<g, 128, 0, st>>>(src, lds, (int)rows, (int)cols, dst, ldd);
        else transpose_bf16_wide_kernel<<<g, 256, 0, st>>>(src, lds, (int)rows, (int)cols, dst, ldd);
    } else {
        dim3 g((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
        transpose_bf16_kernel<<<g, 256, 0, st>>>(src, lds, (int)rows, (int)cols, dst, ldd);
    }
    QB_CUDA(cudaGetLastError());
    h->launches++;
}
void transpose_into(qb_handle *h, const bf16 *src, int64_t lds, int64_t rows, int64_t cols, bf16 *dst, int64_t ldd) {
    transpose_on(h, h->stream, src, lds, rows, cols, dst, ldd);
}

RngKey make_key(qb_handle *h) {
    RngKey k;
    k.seed_lo = (uint32_t)h->seed; k.seed_hi = (uint32_t)(h->seed >> 32);
    k.draw_lo = (uint32_t)h->draw; k.draw_hi = (uint32_t)(h->draw >> 32) & 0x7FFFFFFFu;
    h->draw++;
    return k;
}
Rng philox_rng(qb_handle *h) { Rng r; r.key = make_key(h); r.use_rng = 1; return r; }
Rng injected_rng(const float *p, int64_t ld) { Rng r; r.inj = p; r.ld_inj = ld; return r; }

int64_t row0_of(const qb_handle *h, int64_t B) { return (int64_t)h->cfg.rank * B; }

// GEMM bookkeeping: algorithmic flops and (optionally) CUDA-event timing of every GEMM launch
struct GemmScope {
    qb_handle *h; bool timed = false; size_t i = 0;
    GemmScope(qb_handle *h_, double flops) : h(h_) {
        h->launches++; h->gemm_launches++; h->gemm_flops += flops;
        if (h->time_gemms && h->gemm_ev_used + 2 <= 65536) {
            while (h->gemm_ev.size() < h->gemm_ev_used + 2) {
                cudaEvent_t e; QB_CUDA(cudaEventCreate(&e)); h->gemm_ev.push_back(e);
            }
            i = h->gemm_ev_used; h->gemm_ev_used += 2; timed = true;
            h->gemm_timed++; h->gemm_timed_flops += flops;
            QB_CUDA(cudaEventRecord(h->gemm_ev[i], h->stream));
        }
        sm100::last_shape().kind = -1;
    }
    ~GemmScope() {
        if (timed) cudaEventRecord(h->gemm_ev[i + 1], h->stream);
        // launch histogram: slot = (pair, log2(bn / 32), kind); the SIMT GEMM of the fp32 path counts in the last slot
        const sm100::LaunchShape &ls = sm100::last_shape();
        int slot = QB_GEMM_HIST_SLOTS - 1;
        if (ls.kind >= 0 && ls.kind < QB_GEMM_KINDS) {
            const int b = ls.bn == 256 ? 3 : ls.bn == 128 ? 2 : ls.bn == 64 ? 1 : 0;
            slot = (ls.pair * 4 + b) * QB_GEMM_KINDS + ls.kind;
        }
        h->gemm_hist[slot]++;
    }
};

// QB_PREC_FP32X3: C = alpha A B + beta C + bias on tcgen05 with fp32-grade accuracy.  Every operand is split into bf16
// components x = hi + mid + lo (an operand known to be bf16-exact has one), the products (i, j) with i + j <= 2 -- the ones
// above 2^-24 relative -- accumulate into one fp32 TMEM tile, smallest terms first.  The weights' components (both
// orientations) are cached and refreshed after every update; states and probabilities are split on the fly.
void x3_split(qb_handle *h, const float *src, int64_t s_row, int64_t s_k, int64_t rows, int64_t K, bf16 *const c[3], int ncomp, int64_t ldk) {
    dim3 g((unsigned)ceil_div(ldk, 32), (unsigned)ceil_div(rows, 32));
    split3_kernel<<<g, 256, 0, h->stream>>>(src, s_row, s_k, (int)rows, (int)K, c[0], ncomp > 1 ? c[1] : nullptr, ncomp > 1 ? c[2] : nullptr, ldk);
    QB_CUDA(cudaGetLastError());
    h->launches++;
}

void x3_ensure_weights(qb_handle *h) {
    if (!h->w3_dirty) return;
    for (int i = 0; i < 3; i++) {
        if (!h->W3[i]) h->W3[i] = dalloc<bf16>(h->H * h->ldv);
        if (!h->W3T[i]) h->W3T[i] = dalloc<bf16>(h->VL * h->ldh);
    }
    x3_split(h, h->P, h->ldv, 1, h->H, h->VL, h->W3, 3, h->ldv);
    x3_split(h, h->P, 1, h->ldv, h->VL, h->H, h->W3T, 3, h->ldh);
    h->w3_dirty = false;
}

// components of one operand, as [rows][K] K-major matrices: the weight cache when `src` views W, else a fresh split
int x3_operand(qb_handle *h, int slot, const float *src, int64_t s_row, int64_t s_k, int64_t rows, int64_t K, bool exact,
               sm100::Operand out[3]) {
    const float *W = h->P;
    if (src >= W && src < W + h->nW) {
        x3_ensure_weights(h);
        const int64_t off = src - W;
        if (s_k == 1 && s_row == h->ldv) {            // rows of W: (hidden unit, visible column)
            QB_CHECK(off % 8 == 0, QB_E_UNSUPPORTED, "x3: misaligned view of W");
            for (int i = 0; i < 3; i++) out[i] = sm100::Operand{h->W3[i] + off, rows, K, h->ldv};
            return 3;
        }
        if (s_row == 1 && s_k == h->ldv) {            // rows of W': (visible column, hidden unit); off = first column
            QB_CHECK(off < h->ldv, QB_E_UNSUPPORTED, "x3: unsupported view of W");
            for (int i = 0; i < 3; i++) out[i] = sm100::Operand{h->W3T[i] + off * h->ldh, rows, K, h->ldh};
            return 3;
        }
    }
    const int ncomp = exact ? 1 : 3;
    const int64_t ldk = round_up(K, 8);
    const size_t need = (size_t)rows * ldk;
    if (h->x3_cap[slot] < need) {
        for (int i = 0; i < 3; i++) { cudaFree(h->x3_buf[slot][i]); h->x3_buf[slot][i] = dalloc<bf16>(need); }
        h->x3_cap[slot] = need;
    }
    x3_split(h, src, s_row, s_k, rows, K, h->x3_buf[slot], ncomp, ldk);
    for (int i = 0; i < ncomp; i++) out[i] = sm100::Operand{h->x3_buf[slot][i], rows, K, ldk};
    return ncomp;
}

void x3_gemm(qb_handle *h, const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, float *C, int64_t ldc,
             int64_t M, int64_t N, int64_t K, float alpha, float beta, const float *bias, bool a_exact, bool b_exact) {
    sm100::Operand ac[3], bc[3];
    const int na = x3_operand(h, 0, A, sam, sak, M, K, a_exact, ac);
    const int nb = x3_operand(h, 1, B, sbn, sbk, N, K, b_exact, bc);
    sm100::Product prods[sm100::MAX_PHASES];
    int n = 0;
    for (int sum = 2; sum >= 0; sum--)          // smallest components first
        for (int i = std::min(sum, na - 1); i >= 0; i--) {
            const int j = sum - i;
            if (j >= nb) continue;
            prods[n++] = sm100::Product{ac[i], bc[j], false};
        }
    sm100::EpiParams ep{};
    ep.mode = sm100::EPI_GRAD; ep.M = (int)M; ep.N = (int)N; ep.out_f32 = C; ep.ld_out = ldc;
    ep.g_affine = 1; ep.g_alpha = alpha; ep.g_beta = beta; ep.g_bias = bias;
    GemmScope gs(h, 2.0 * (double)M * (double)N * (double)K);
    h->gemm_issued_flops += 2.0 * (double)M * (double)N * (double)K * (double)(n - 1);   // component products beyond the algorithmic one
    sm100::launch_products(h->stream, prods, n, ep, h->gemm_sms, h->force_bn, h->use_2cta);
}

void simt_gemm(qb_handle *h, const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbk, int64_t sbn, float *C,
               int64_t ldc, int64_t M, int64_t N, int64_t K, float alpha, float beta, const float *bias, bool a_exact = false,
               bool b_exact = false) {
    if (h->x3) { x3_gemm(h, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, alpha, beta, bias, a_exact, b_exact); return; }
    GemmScope gs(h, 2.0 * (double)M * (double)N * (double)K);
    dim3 g((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
    simt_gemm_kernel<<<g, 256, 0, h->stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, (int)M, (int)N, (int)K, alpha, beta, bias);
    QB_CUDA(cudaGetLastError());
}

// ----------------------------------------------------------------------------- fused chains
constexpr int QB_CHAIN_REGIONS = 32;

void chain_flush(qb_handle *h);

void chain_begin(qb_handle *h) {
    if (h->rec_depth++ > 0) return;
    if (!h->rec) h->rec = new sm100::ChainBuilder();
    h->rec->prog.n_steps = 0;
}
void chain_end(qb_handle *h) {
    if (--h->rec_depth > 0) return;
    chain_flush(h);
}
struct ChainRegion {   // records the half-steps issued inside a scope, launches them when the scope ends
    qb_handle *h;
    explicit ChainRegion(qb_handle *h_) : h(h_) { chain_begin(h); }
    ~ChainRegion() noexcept(false) {
        if (std::uncaught_exceptions()) { h->rec_depth = 0; if (h->rec) h->rec->prog.n_steps = 0; return; }
        chain_end(h);
    }
};

// Width (batch rows) of the chain kernel's 256-unit pair tiles.  Measured (tools/chain_probe.py, profiles/r2b_chain_probe.txt):
// up to ONE wave of 256 x 256 tiles per half-step every pair would hold a single tile whose epilogue is exposed, and 256 x 128
// tiles win (cfg5 at 1024 rows: 625 vs 679 us per step; 512 rows: 449 vs 637); from there on the wide tile wins (2048 rows:
// 991 vs 1159; 8192 rows: 3962 vs 4911).
int chain_tile_width(qb_handle *h, int64_t M, int64_t N) {
    if (h->chain_bn) return h->chain_bn;
    if (h->force_bn >= 64) return h->force_bn;
    if (N <= 64) return 64;
    return (int64_t)ceil_div(N, 256) * ceil_div(M, 256) > h->gemm_sms / 2 ? 256 : 128;
}
// The pair-tile chain wins from about three quarters of a wave of its tiles per half-step on (cfg5: 512 rows, 64 tiles of
// 256 x 128: 449 vs 474 us per step; 1024 rows: 625 vs 699; 2048 rows: 991 vs 1122; 8192 rows: a tie) and loses on small
// models, whose half-steps want many small single-CTA tiles.
// Option "fused_chain": 1 = automatic, 2 = always, 0 = never.
// `units`: the larger layer of the model, so that the hidden and the visible half-steps of one chain get the same answer.
bool chain_pays(qb_handle *h, int64_t units, int64_t N) {
    if (h->fused_chain_mode == 2) return true;
    return (int64_t)ceil_div(N, chain_tile_width(h, units, N)) * ceil_div(units, 256) >= (h->gemm_sms * 3) / 8;
}
// The quad variant (gibbs_chain_kernel_quad): two CTA pairs in a 4-CTA cluster share a 256-unit x 256-row tile (128-row blocks
// per pair, weights multicast between the pairs), so every pair alternates between two independent tiles at one wave.  Bit-
// identical results; with the warp-uniform issue loops it is level with the plain 256 x 128 pair chain (624 vs 625 us per cfg5
// step at 1024 rows, profiles/r2b_chain_probe.txt), which gets the same alternation without clusters of four -- the multicast
// buys nothing measurable.  (Its first measurement, 1028 vs 723 us, was taken when every narrow tile was bound by the
// single-lane issue loop)  Off by default; kept as an option with its parity test.
bool chain_quad_wanted(qb_handle *h, const sm100::EpiParams &ep) {
    if (h->chain_quad_mode == 0 || h->fused_chain_mode == 0 || h->chain_bn != 0 || h->force_bn != 0 || (ep.N % 256) != 0) return false;
    if (h->gemm_sms < h->num_sms) return false;   // (stale-chains mode reserves SMs for the exchange: pair kernels only)
    if (h->chain_quad_mode == 2) return true;
    const int64_t tiles = (int64_t)(ep.N / 256) * ceil_div(ep.M, 256);
    return tiles >= h->num_sms / 8 && tiles < h->gemm_sms / 2 + h->gemm_sms / 4;
}

void launch_single(qb_handle *h, const sm100::Operand &A, const sm100::Operand &B, const sm100::EpiParams &ep, double flops) {
    GemmScope gs(h, flops);
    sm100::launch(h->stream, A, B, nullptr, nullptr, ep, h->gemm_sms, h->force_bn, h->use_2cta);
}

// does step i (about to be appended, reading B and writing the row-major / transposed outputs of ep) only touch buffers of
// earlier steps that are its ancestors along the dependency chain?  (row-block counters order exactly those)
bool chain_hazard_free(const sm100::ChainBuilder &cb, const sm100::Operand &B, const sm100::EpiParams &ep, int dep) {
    bool anc[sm100::CHAIN_MAX_STEPS] = {};
    for (int a = dep; a >= 0; a = cb.prog.step[a].dep_step) anc[a] = true;
    const void *outs[4] = {ep.state_rm, ep.prob_rm, ep.state_T, ep.prob_T};
    for (int j = 0; j < cb.prog.n_steps; j++) {
        if (anc[j]) continue;
        const sm100::ChainStep &S = cb.prog.step[j];
        const void *jouts[4] = {S.ep.state_rm, S.ep.prob_rm, S.ep.state_T, S.ep.prob_T};
        const void *jin = cb.ops[S.b_map].ptr;
        for (const void *o : outs) {
            if (!o) continue;
            if (o == jin) return false;                      // write after read of a non-ancestor
            for (const void *q : jouts) if (q == o) return false;   // write after write
        }
        for (const void *q : jouts) if (q && q == (const void *)B.ptr) return false;   // read of a non-ancestor's output (cannot happen: dep is the latest writer)
    }
    return true;
}

// one fused half-step GEMM of the bf16 path: recorded into the open chain region, or launched right away
void gemm_half_step(qb_handle *h, const sm100::Operand &A, const sm100::Operand &B, const sm100::EpiParams &ep, double flops) {
    if (h->capture) { h->capture->A = A; h->capture->B = B; h->capture->ep = ep; h->capture->flops = flops; h->capture->set = true; return; }
    const bool quad = h->rec_depth > 0 && h->use_2cta != 0 && sm100::chainable(ep) && chain_quad_wanted(h, ep);
    const bool can_chain = h->rec_depth > 0 && h->fused_chain_mode != 0 && h->use_2cta != 0 && h->force_bn != 32 && sm100::chainable(ep) &&
                           (quad || chain_pays(h, std::max<int64_t>(ep.M, A.k), ep.N));
    if (h->rec_depth > 0 && !can_chain) chain_flush(h);
    if (!can_chain) { launch_single(h, A, B, ep, flops); return; }
    sm100::ChainBuilder &cb = *h->rec;
    if (cb.prog.n_steps > 0 && (cb.prog.N != ep.N || cb.full() || cb.quad != quad)) chain_flush(h);
    for (int attempt = 0; attempt < 2; attempt++) {
        if (cb.prog.n_steps == 0) cb.begin(ep.N, quad ? sm100::QUAD_BN : chain_tile_width(h, std::max<int64_t>(ep.M, A.k), ep.N), quad);
        int dep = -1;
        for (int j = cb.prog.n_steps - 1; j >= 0; j--) {
            const sm100::EpiParams &e = cb.prog.step[j].ep;
            if ((const void *)e.state_rm == (const void *)B.ptr || (const void *)e.prob_rm == (const void *)B.ptr) { dep = j; break; }
        }
        if (cb.n_maps + 2 <= sm100::CHAIN_MAX_MAPS && chain_hazard_free(cb, B, ep, dep)) {
            cb.add(A, B, ep, dep);
            return;
        }
        chain_flush(h);   // start a new program with this step first
    }
    launch_single(h, A, B, ep, flops);
}

// CTA pairs the chain kernel runs on: all of them.  Option "chain_pairs" (experiments) limits the count.  Tried: tiles are dealt
// to the pairs round-robin in program order and a tile of half-step t + 1 depends on tiles one half-step's count T earlier, so
// with T a whole number of rounds (128 tiles on 64 pairs instead of 1.73 rounds on 74) no dependency is in flight when its
// consumer comes up -- measured slower (cfg5 at 1024 rows: 664 vs 634 us per step; 2048 rows: 1046 vs 1008): the schedule is
// not stall-bound, the ten extra pairs are worth more.
int chain_pairs_for(qb_handle *h, const sm100::ChainBuilder &) {
    const int pairs = h->gemm_sms / 2;
    return h->chain_pairs > 0 ? std::min(pairs, h->chain_pairs) : pairs;
}

void ensure_chain_flags(qb_handle *h) {
    if (h->chain_cnt) return;
    h->chain_cnt_region = (int64_t)sm100::CHAIN_MAX_STEPS * ceil_div(h->Bmax, 64);
    h->chain_cnt = dalloc<unsigned int>((size_t)QB_CHAIN_REGIONS * h->chain_cnt_region + 1);   // + the abort flag
    QB_CUDA(cudaHostAlloc((void **)&h->chain_err_host, sizeof(unsigned int), cudaHostAllocMapped));
    *h->chain_err_host = 0u;
    QB_CUDA(cudaHostGetDevicePointer((void **)&h->chain_err_dev, h->chain_err_host, 0));
}

void chain_flush(qb_handle *h) {
    if (!h->rec || h->rec->prog.n_steps == 0) return;
    sm100::ChainBuilder &cb = *h->rec;
    const int n = cb.prog.n_steps;
    cb.prog.n_steps = 0;   // (re-entrancy: launch_single below must not see the program)
    if (n == 1) {
        const sm100::ChainStep &S = cb.prog.step[0];
        launch_single(h, cb.ops[S.a_map], cb.ops[S.b_map], S.ep, cb.flops);
        return;
    }
    cb.prog.n_steps = n;
    ensure_chain_flags(h);
    // arrival counters: a fresh region of a ring per launch, the whole ring is cleared once per lap
    const int64_t region = h->chain_launch_no % QB_CHAIN_REGIONS;
    if (region == 0 && h->chain_launch_no > 0)
        QB_CUDA(cudaMemsetAsync(h->chain_cnt, 0, (size_t)QB_CHAIN_REGIONS * h->chain_cnt_region * sizeof(unsigned int), h->stream));
    h->chain_launch_no++;
    cb.prog.counters = h->chain_cnt + region * h->chain_cnt_region;
    cb.prog.error_flag = h->chain_err_dev;
    cb.prog.abort_flag = h->chain_cnt + (size_t)QB_CHAIN_REGIONS * h->chain_cnt_region;
    cb.pairs_limit = chain_pairs_for(h, cb);
    {
        GemmScope gs(h, cb.flops);
        sm100::launch_chain(h->stream, cb, h->gemm_sms);
        sm100::last_shape().bn = cb.bn; sm100::last_shape().pair = 1; sm100::last_shape().kind = sm100::K_CHAIN;
    }
    h->chain_launches++;
    cb.prog.n_steps = 0;
}

void chain_check_error(qb_handle *h) {
    if (h->p2p_timeout_host && *(volatile unsigned int *)h->p2p_timeout_host) {
        *h->p2p_timeout_host = 0u;
        throw Error(QB_E_CUDA, "peer-memory exchange: a peer rank did not reach the gradient barrier within 5 s (ranks out of step or a rank died); parameters are invalid");
    }
    if (h->chain_err_host && *(volatile unsigned int *)h->chain_err_host) {
        *h->chain_err_host = 0u;
        cudaMemsetAsync(h->chain_cnt + (size_t)QB_CHAIN_REGIONS * h->chain_cnt_region, 0, sizeof(unsigned int), h->stream);
        throw Error(QB_E_CUDA, "fused chain kernel: a row-block dependency was not satisfied in time (results of the step are invalid)");
    }
}

// ----------------------------------------------------------------------------- the three ops
struct HiddenOut {
    Mat *prob = nullptr;   // probabilities (fp32 path: .f; bf16 path: .b and/or .bT per flags)
    Mat *state = nullptr;  // sampled binary states
    bool prob_rm = false, prob_T = false, state_rm = false, state_T = false;
    float *prob_f32 = nullptr; int64_t ld_pf = 0;  // bf16 path only: extra fp32 copy of the probabilities
    int bg_what = 0; float bg_sign = 0.f;          // bf16 path: fold the hidden bias gradient
    float *softplus_rows = nullptr;                // free-energy mode (no other outputs)
    float pscale = 1.f;                            // positive phase with n_chains != batch: probabilities x (n_chains / batch)
};

// conditional_prob_h (+ gibbs_sample_hidden): in = visible-side rows [rows][kdim] (kdim = V or V+L)
void op_hidden(qb_handle *h, const Mat &in, int64_t rows, int64_t kdim, const HiddenOut &o, const Rng &rng, int64_t row0) {
    if (!h->bf) {
        simt_gemm(h, in.f, in.ld, 1, Wf(h), 1, h->ldv, h->pre_h, h->ldh, rows, h->H, kdim, 1.f, 0.f, bh(h), in.exact);
        if (o.state) o.state->exact = true;
        if (o.prob) o.prob->exact = false;
        if (o.softplus_rows) return;  // fp32 path reads pre_h directly
        act_sample_kernel<<<grid1d(rows * h->H), 256, 0, h->stream>>>(
            h->pre_h, h->ldh, (int)rows, (int)h->H, 0, o.prob ? o.prob->f : nullptr, o.prob ? o.prob->ld : 0,
            o.state ? o.state->f : nullptr, o.state ? o.state->ld : 0, rng.inj, rng.ld_inj, rng.key, row0, rng.use_rng);
        QB_CUDA(cudaGetLastError());
        h->launches++;
        return;
    }
    sm100::Operand A{h->Wb, h->H, kdim, h->ldv}, B{in.b, rows, kdim, in.ld};
    sm100::EpiParams ep{};
    ep.mode = sm100::EPI_SAMPLE; ep.M = (int)h->H; ep.N = (int)rows;
    ep.bias = bias_b(h); ep.split = (int)h->H;
    ep.act = o.softplus_rows ? sm100::ACT_GAUSSIAN : sm100::ACT_SIGMOID;
    ep.sample = (o.state != nullptr);
    if (o.state) {
        if (o.state_rm) { ep.state_rm = o.state->b; ep.ld_srm = o.state->ld; }
        if (o.state_T) { ep.state_T = o.state->bT; ep.ld_sT = o.state->ldT; }
    }
    if (o.prob) {
        if (o.prob_rm) { ep.prob_rm = o.prob->b; ep.ld_prm = o.prob->ld; }
        if (o.prob_T) { ep.prob_T = o.prob->bT; ep.ld_pT = o.prob->ldT; }
    }
    ep.prob_f32 = o.prob_f32; ep.ld_pf = o.ld_pf;
    ep.inj = rng.inj; ep.ld_inj = rng.ld_inj; ep.key = rng.key; ep.use_rng = rng.use_rng; ep.row0 = row0;
    ep.bg_what = o.bg_what; ep.bg_sign = o.bg_sign; ep.bias_grad = G_b(h);
    ep.softplus_rows = o.softplus_rows;
    ep.pscale = o.pscale;
    QB_CHECK(o.pscale == 1.f || sm100::pick_kind(ep) == sm100::K_FAST_PROB, QB_E_UNSUPPORTED, "pscale needs the fast probability epilogue");
    gemm_half_step(h, A, B, ep, 2.0 * (double)h->H * (double)rows * (double)kdim);
}

struct VisibleOut {
    Mat *state = nullptr;  // sampled visible (+label) state, or mean-field values when sample == false
    bool sample = true;
    bool state_rm = true, state_T = false;
    float *prob_f32 = nullptr; int64_t ld_pf = 0;  // fp32 copy of probabilities / means [rows][ldv]
    bool f32_from_state = false;                    // bf16 path: prob_f32 receives the (noisy) value
    int bg_sign_neg = 0;                            // bf16 path: G_a -= column sums of the state
    const Mat *ref = nullptr; double *sqerr = nullptr;  // reconstruction error vs ref rows (visible units only)
};

// conditional_prob_v / gibbs_sample_visible (+ gibbs_sample_label): in = hidden rows [rows][H]
void op_visible(qb_handle *h, const Mat &in, int64_t rows, const VisibleOut &o, const Rng &rng, int64_t row0) {
    const int act = h->gaussian ? (h->gauss_threshold ? 2 : 1) : 0;
    QB_CHECK(act != 2 || (!rng.inj && rng.use_rng), QB_E_UNSUPPORTED, "thresholded Gaussian visibles draw from the Philox stream only");
    if (!h->bf) {
        simt_gemm(h, in.f, in.ld, 1, Wf(h), h->ldv, 1, h->pre_v, h->ldv, rows, h->VL, h->H, 1.f, 0.f, acat(h), in.exact);
        if (o.state) o.state->exact = (act != 1) && o.sample;   // Bernoulli (or thresholded) samples are 0/1; means and Gaussian samples are reals
        // value written for each visible unit: sampled state (sample), probability (Bernoulli
        // mean-field) or mean + noise (Gaussian; noise only if the caller supplied a stream)
        float *st = o.state ? o.state->f : nullptr;
        int64_t lds = o.state ? o.state->ld : 0;
        float *prob_dst = o.prob_f32;
        int64_t ldp = o.ld_pf;
        if (act == 0 && !o.sample) {
            if (!prob_dst) { prob_dst = st; ldp = lds; }
            st = nullptr;
        }
        act_sample_kernel<<<grid1d(rows * h->V), 256, 0, h->stream>>>(h->pre_v, h->ldv, (int)rows, (int)h->V, act, prob_dst, ldp, st,
                                                                      lds, rng.inj, rng.ld_inj, rng.key, row0, rng.use_rng);
        QB_CUDA(cudaGetLastError());
        h->launches++;
        if (h->L > 0 && o.sample) {
            label_kernel<<<grid1d(rows, 128), 128, 0, h->stream>>>(h->pre_v, h->ldv, (int)rows, (int)h->V, (int)h->L, nullptr, 0,
                                                                   o.state->f, o.state->ld, rng.inj, rng.ld_inj, rng.key, row0);
            QB_CUDA(cudaGetLastError());
            h->launches++;
        }
        if (o.sqerr) {
            const float *recon = (act == 0) ? prob_dst : o.state->f;
            int64_t ldr = (act == 0) ? ldp : o.state->ld;
            sqerr_kernel<float><<<std::min<int64_t>(1184, ceil_div(rows * h->V, 256)), 256, 0, h->stream>>>(
                o.ref->f, o.ref->ld, recon, ldr, (int)rows, (int)h->V, o.sqerr);
            QB_CUDA(cudaGetLastError());
            h->launches++;
        }
        return;
    }
    sm100::Operand A{h->WbT, h->VL, h->H, h->ldh}, B{in.b, rows, h->H, in.ld};
    sm100::EpiParams ep{};
    ep.mode = sm100::EPI_SAMPLE; ep.M = (int)h->VL; ep.N = (int)rows;
    ep.bias = bias_a(h); ep.split = (int)h->V; ep.lab_pre = h->lab_pre; ep.ld_lab = h->L > 0 ? h->L : 1;
    ep.act = act == 2 ? sm100::ACT_GAUSS_THRESH : (act ? sm100::ACT_GAUSSIAN : sm100::ACT_SIGMOID);
    ep.sample = o.sample || (act == 1 && (rng.inj || rng.use_rng));
    if (o.state) {
        if (o.state_rm) { ep.state_rm = o.state->b; ep.ld_srm = o.state->ld; }
        if (o.state_T) { ep.state_T = o.state->bT; ep.ld_sT = o.state->ldT; }
    }
    ep.prob_f32 = o.prob_f32; ep.ld_pf = o.ld_pf; ep.f32_from_state = o.f32_from_state ? 1 : 0;
    ep.inj = rng.inj; ep.ld_inj = rng.ld_inj; ep.key = rng.key; ep.use_rng = rng.use_rng; ep.row0 = row0;
    if (o.bg_sign_neg) { ep.bg_what = 2; ep.bg_sign = -1.f; ep.bias_grad = G_a(h); }
    if (o.sqerr) { ep.sqerr = o.sqerr; ep.ref_rm = o.ref->b; ep.ld_ref = o.ref->ld; }
    ep.pscale = 1.f;
    gemm_half_step(h, A, B, ep, 2.0 * (double)h->VL * (double)rows * (double)h->H);
    // label units: finished inside the GEMM epilogue (K_FAST_LABEL) or, from the raw pre-activations, by a side kernel
    if (h->L > 0 && o.sample && o.state && sm100::pick_kind(ep) != sm100::K_FAST_LABEL && !h->capture) {
        label_bf16_kernel<<<grid1d(rows, 128), 128, 0, h->stream>>>(
            h->lab_pre, h->L, (int)rows, (int)h->V, (int)h->L, o.state_rm ? o.state->b : nullptr, o.state->ld,
            o.state_T ? o.state->bT : nullptr, o.state->ldT, rng.inj, rng.ld_inj, rng.key, row0,
            o.bg_sign_neg ? G_a(h) : nullptr);
        QB_CUDA(cudaGetLastError());
        h->launches++;
    }
}

template <typename T>
void colsum(qb_handle *h, const T *A, int64_t lda, int64_t rows, int64_t cols, float sign, float *out) {
    dim3 g((unsigned)ceil_div(cols, 32), (unsigned)std::min<int64_t>(64, ceil_div(rows, 8)));
    colsum_kernel<T><<<g, 256, 0, h->stream>>>(A, lda, (int)rows, (int)cols, sign, out);
    QB_CUDA(cudaGetLastError());
    h->launches++;
}
template <>
void colsum<bf16>(qb_handle *h, const bf16 *A, int64_t lda, int64_t rows, int64_t cols, float sign, float *out) {
    const int64_t gx = ceil_div(cols, 256);
    dim3 g((unsigned)gx, (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 8), ceil_div(4 * h->num_sms, gx))));
    colsum_bf16x8_kernel<<<g, 256, 0, h->stream>>>(A, lda, (int)rows, (int)cols, sign, out);
    QB_CUDA(cudaGetLastError());
    h->launches++;
}

// G = posH' posV - negH' negV   (and the bias gradients that were not folded into epilogues)
// `rows` data rows, `nrows` negative-phase rows (chains); pos_scale = nrows / rows multiplies the positive statistics
// (on the bf16 path it was already applied to posH by the positive-phase epilogue)
void op_grad(qb_handle *h, const Mat &posH, const Mat &posV, const bf16 *posV_T, int64_t posV_ldT, const Mat &negH,
             const Mat &negV, int64_t rows, int64_t nrows, float pos_scale) {
    if (!h->bf) {
        simt_gemm(h, posH.f, 1, posH.ld, posV.f, posV.ld, 1, G_W(h), h->ldv, h->H, h->VL, rows, pos_scale, 0.f, nullptr, posH.exact, posV.exact);
        simt_gemm(h, negH.f, 1, negH.ld, negV.f, negV.ld, 1, G_W(h), h->ldv, h->H, h->VL, nrows, -1.f, 1.f, nullptr, negH.exact, negV.exact);
        colsum<float>(h, posV.f, posV.ld, rows, h->VL, pos_scale, G_a(h));
        colsum<float>(h, negV.f, negV.ld, nrows, h->VL, -1.f, G_a(h));
        colsum<float>(h, posH.f, posH.ld, rows, h->H, pos_scale, G_b(h));
        colsum<float>(h, negH.f, negH.ld, nrows, h->H, -1.f, G_b(h));
        return;
    }
    sm100::Operand A{posH.bT, h->H, rows, posH.ldT}, B{posV_T, h->VL, rows, posV_ldT};
    sm100::Operand A2{negH.bT, h->H, nrows, negH.ldT}, B2{negV.bT, h->VL, nrows, negV.ldT};
    sm100::EpiParams ep{};
    ep.mode = sm100::EPI_GRAD; ep.M = (int)h->H; ep.N = (int)h->VL;
    ep.out_f32 = G_W(h); ep.ld_out = h->ldv;
    {
        GemmScope gs(h, 2.0 * (double)h->H * (double)h->VL * (double)(rows + nrows));
        sm100::launch(h->stream, A, B, &A2, &B2, ep, h->gemm_sms, h->force_bn, h->use_2cta);
    }
    if (!h->cur_datasum) colsum<bf16>(h, posV.b, posV.ld, rows, h->VL, pos_scale, G_a(h));  // data-side visible sums; the rest is folded
}

// bias-gradient accumulators start at the data-side visible sums when those are cached
void zero_bias_grads(qb_handle *h, float pos_scale = 1.f) {
    const int n = (int)(h->ldv + h->ldh);
    // single GPU, synchronous update: the update kernel adds the cached data sums itself and leaves the
    // accumulators at zero, so a step needs no init launch at all
    h->datasum_in_update = h->bf && h->cur_datasum != nullptr && h->cfg.world == 1 && !h->stale_chains;
    h->datasum_scale = pos_scale;
    // multi-GPU / stale mode: the data sums must be inside G before the exchange
    const float *init_sum = (h->bf && !h->datasum_in_update) ? h->cur_datasum : nullptr;
    const bool clean = h->bias_grads_clean;
    h->bias_grads_clean = false;  // this step dirties the accumulators; the update kernel re-establishes the invariant
    if (clean && !init_sum) return;
    init_bias_grads_kernel<<<grid1d(n), 256, 0, h->stream>>>(G_a(h), init_sum, pos_scale, (int)h->ldv, n);
    QB_CUDA(cudaGetLastError());
    h->launches++;
}

// update_rbm!(rbm, dW, [dU,] da, db, [dc,] lr/B [, llr/B]) after the data-parallel sum of G.
// `st` is the stream it runs on; (Wb_out, WbT_out, bias_out) the shadow set it refreshes.
void p2p_update(qb_handle *h, double lr, double llr, int64_t B_local);

void op_update_on(qb_handle *h, double lr, double llr, int64_t B_local, cudaStream_t st, float *G, bf16 *Wb_out, bf16 *WbT_out,
                  float *bias_out) {
    if (h->p2p && st == h->stream && G == h->G && Wb_out == h->Wb) { p2p_update(h, lr, llr, B_local); return; }
    ncclComm_t comm = (st == h->comm_stream && h->comm_overlap) ? h->comm_overlap : h->comm;
    const bool sharded = comm && h->sharded_update && h->bf && (h->H % h->cfg.world) == 0;
    const int64_t sr = sharded ? h->H / h->cfg.world : h->H, hr0 = sharded ? sr * h->cfg.rank : 0;
    if (sharded && h->exchange_done) {
        h->master_sharded = true;
    } else if (sharded) {
        // in-place reduce-scatter: this rank ends up with the summed gradient of the hidden rows it owns; the biases are tiny
        QB_NCCL(nccl().GroupStart());
        QB_NCCL(nccl().ReduceScatter(G, G + hr0 * h->ldv, (size_t)(sr * h->ldv), /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, st));
        QB_NCCL(nccl().AllReduce(G + h->nW, G + h->nW, (size_t)(h->ldv + h->ldh), 7, 0, comm, st));
        QB_NCCL(nccl().GroupEnd());
        h->master_sharded = true;
    } else if (comm) {
        QB_NCCL(nccl().AllReduce(G, G, (size_t)h->nP, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, st));
    }
    const double Bg = (double)B_local * (double)h->cfg.world;
    const float lr_e = (float)(lr / Bg), llr_e = (float)(llr / Bg);
    const bool use_vel = (h->momentum != 0.0 || h->weight_decay != 0.0);
    if (use_vel && !h->Vel) h->Vel = dalloc<float>(h->nP);
    // one launch: W tiles plus one extra row of blocks for the bias region [acat | b]
    dim3 g((unsigned)ceil_div(h->VL, 128), (unsigned)ceil_div(sr, 64) + 1);
    const int nb = (int)(h->ldv + h->ldh);
    const bool own_step = (st == h->stream && G == h->G);
    BiasTail bt{};
    bt.P = acat(h); bt.G = G + h->nW; bt.Vel = use_vel ? h->Vel + h->nW : nullptr;
    bt.datasum = (own_step && h->datasum_in_update) ? h->cur_datasum : nullptr; bt.dscale = h->datasum_scale;
    bt.copy_out = bias_out; bt.n_vis = (int)h->ldv; bt.n_total = nb;
    bt.zero_after = (own_step && h->cfg.world == 1 && !h->stale_chains) ? 1 : 0;
    size_t ei = 0; bool timed = false;
    if (h->time_gemms && h->upd_ev_used + 2 <= 4096) {
        while (h->upd_ev.size() < h->upd_ev_used + 2) { cudaEvent_t e; QB_CUDA(cudaEventCreate(&e)); h->upd_ev.push_back(e); }
        ei = h->upd_ev_used; h->upd_ev_used += 2; h->upd_timed++; timed = true;
        QB_CUDA(cudaEventRecord(h->upd_ev[ei], st));
    }
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = g; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        float *Wp = Wf(h) + hr0 * h->ldv, *Velp = use_vel ? h->Vel + hr0 * h->ldv : nullptr;
        const float *Gp = G + hr0 * h->ldv;
        bf16 *Wb_rows = Wb_out + hr0 * h->ldv;
        bf16 *WbT_cols = sharded ? nullptr : WbT_out;   // sharded: the transposed shadow is rebuilt after the all-gather
        int Hh = (int)sr, VLl = (int)h->VL, Vv = (int)h->V;
        int64_t ldv = h->ldv, ldh = h->ldh;
        float mom = use_vel ? (float)h->momentum : 0.f, wd = use_vel ? (float)h->weight_decay : 0.f;
        if (use_vel) QB_CUDA(cudaLaunchKernelEx(&cfg, update_w_kernel<true>, Wp, Gp, Velp, Hh, VLl, Vv, ldv, lr_e, llr_e, mom, wd, Wb_rows, WbT_cols, ldh, bt));
        else QB_CUDA(cudaLaunchKernelEx(&cfg, update_w_kernel<false>, Wp, Gp, Velp, Hh, VLl, Vv, ldv, lr_e, llr_e, mom, wd, Wb_rows, WbT_cols, ldh, bt));
    }
    if (timed) QB_CUDA(cudaEventRecord(h->upd_ev[ei + 1], st));
    QB_CUDA(cudaGetLastError());
    h->launches += 1;
    h->w3_dirty = true;
    if (sharded) {
        QB_NCCL(nccl().AllGather(Wb_out + hr0 * h->ldv, Wb_out, (size_t)(sr * h->ldv), /*ncclBfloat16*/ 9, comm, st));
        dim3 gt((unsigned)ceil_div(h->VL, 128), (unsigned)ceil_div(h->H, 64));
        transpose_bf16_wide_kernel<<<gt, 256, 0, st>>>(Wb_out, h->ldv, (int)h->H, (int)h->VL, WbT_out, h->ldh);
        QB_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    if (own_step) h->bias_grads_clean = bt.zero_after != 0;
}
void op_update(qb_handle *h, double lr, double llr, int64_t B_local) {
    op_update_on(h, lr, llr, B_local, h->stream, h->G, h->Wb, h->WbT, h->bias_rd);
}

// reduce-scatter + update + all-gather of the bf16 shadows in one kernel over NVLink peer memory
// fills the launch parameters shared by the exchange kernels
void p2p_fill(qb_handle *h, P2pParams &p) {
    p.rank = h->cfg.rank; p.world = h->cfg.world;
    for (int r = 0; r < p.world; r++) {
        p.G[r] = h->peerG[r]; p.Wb[r] = h->peerWb[r]; p.WbT[r] = h->peerWbT[r]; p.flags[r] = h->peerFlags[r];
    }
    p.G_mc = h->G_mc; p.Wb_mc = h->Wb_mc; p.Gn16_mc = h->Gn16_mc;
    p.H = (int)h->H; p.VL = (int)h->VL; p.V = (int)h->V; p.ldv = h->ldv; p.ldh = h->ldh; p.nW = h->nW;
    p.h_begin = (int)std::min<int64_t>(h->H, (int64_t)p.rank * h->shard_rows);
    p.h_end = (int)std::min<int64_t>(h->H, (int64_t)(p.rank + 1) * h->shard_rows);
    if (!h->p2p_timeout_host) {
        QB_CUDA(cudaHostAlloc((void **)&h->p2p_timeout_host, sizeof(unsigned int), cudaHostAllocMapped));
        *h->p2p_timeout_host = 0u;
        QB_CUDA(cudaHostGetDevicePointer((void **)&h->p2p_timeout_dev, h->p2p_timeout_host, 0));
    }
    p.timeout_flag = h->p2p_timeout_dev;
}

// split exchange, background part: after the positive product is in G (stream order on `st`), tell the peers and reduce the
// owned rows in place through the switch.  Runs on the communication stream beside the chain advance.
void split_reduce_positive(qb_handle *h, cudaStream_t st) {
    P2pParams p{};
    p2p_fill(h, p);
    p.epoch = h->p2p_epoch; h->p2p_epoch += 1;
    nvls_signal_kernel<<<1, 32, 0, st>>>(p);
    QB_CUDA(cudaGetLastError());
    const int64_t n4 = (int64_t)(p.h_end - p.h_begin) * h->ldv / 4;
    // at most ONE block per SM: a persistent GEMM CTA (61440 registers) and one of these blocks (4096) fill the register file
    // exactly -- two of them on an SM would keep the chain kernel, launched right after, from starting there until they finish
    // (measured: the whole reduction then serialises in front of the chain advance)
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(h->num_sms, ceil_div(n4, 4 * 128)));
    nvls_reduce_shard_kernel<<<blocks, 128, 0, st>>>(p, h->G);
    QB_CUDA(cudaGetLastError());
    h->launches += 2;
}

void p2p_update(qb_handle *h, double lr, double llr, int64_t B_local, bool split);
void p2p_update(qb_handle *h, double lr, double llr, int64_t B_local) { p2p_update(h, lr, llr, B_local, false); }
void p2p_update(qb_handle *h, double lr, double llr, int64_t B_local, bool split) {
    const double Bg = (double)B_local * (double)h->cfg.world;
    const bool use_vel = (h->momentum != 0.0 || h->weight_decay != 0.0);
    if (use_vel && !h->Vel) h->Vel = dalloc<float>(h->nP);
    P2pParams p{};
    p.rank = h->cfg.rank; p.world = h->cfg.world;
    for (int r = 0; r < p.world; r++) {
        p.G[r] = h->peerG[r]; p.Wb[r] = h->peerWb[r]; p.WbT[r] = h->peerWbT[r]; p.flags[r] = h->peerFlags[r];
    }
    p.G_mc = h->G_mc; p.Wb_mc = h->Wb_mc; p.Gn16_mc = h->Gn16_mc;
    p.W = h->P; p.Vel = use_vel ? h->Vel : nullptr; p.bias_rd = h->bias_rd;
    p.H = (int)h->H; p.VL = (int)h->VL; p.V = (int)h->V; p.ldv = h->ldv; p.ldh = h->ldh; p.nW = h->nW;
    p.h_begin = (int)std::min<int64_t>(h->H, (int64_t)p.rank * h->shard_rows);
    p.h_end = (int)std::min<int64_t>(h->H, (int64_t)(p.rank + 1) * h->shard_rows);
    p.lr = (float)(lr / Bg); p.llr = (float)(llr / Bg); p.momentum = (float)h->momentum; p.wd = (float)h->weight_decay;
    p.epoch = h->p2p_epoch; h->p2p_epoch += 2;
    p.grid_bar = h->p2p_grid_bar; p.phase_ns = h->p2p_ns;
    if (!h->p2p_timeout_host) {
        QB_CUDA(cudaHostAlloc((void **)&h->p2p_timeout_host, sizeof(unsigned int), cudaHostAllocMapped));
        *h->p2p_timeout_host = 0u;
        QB_CUDA(cudaHostGetDevicePointer((void **)&h->p2p_timeout_dev, h->p2p_timeout_host, 0));
    }
    if (*(volatile unsigned int *)h->p2p_timeout_host) {   // raised by an earlier launch: the peers are out of step
        *h->p2p_timeout_host = 0u;
        throw Error(QB_E_CUDA, "peer-memory exchange: a peer rank did not reach the gradient barrier within 5 s (ranks out of step or a rank died); parameters are invalid");
    }
    p.timeout_flag = h->p2p_timeout_dev;
    const int p2p_grid = h->num_sms * h->p2p_blocks_per_sm;
    QB_CUDA(cudaMemsetAsync(h->p2p_grid_bar, 0, sizeof(unsigned int), h->stream));
    size_t ei = 0; bool timed = false;
    if (h->time_gemms && h->upd_ev_used + 2 <= 4096) {
        while (h->upd_ev.size() < h->upd_ev_used + 2) { cudaEvent_t e; QB_CUDA(cudaEventCreate(&e)); h->upd_ev.push_back(e); }
        ei = h->upd_ev_used; h->upd_ev_used += 2; h->upd_timed++; timed = true;
        QB_CUDA(cudaEventRecord(h->upd_ev[ei], h->stream));
    }
    void *args[] = {(void *)&p};
    if (h->nvls) {
        if (split) {
            if (use_vel) QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<true, true, true>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
            else QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<false, true, true>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
        } else if (use_vel) QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<true, true>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
        else QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<false, true>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
        // every rank's rows of the row-major shadow have landed (second barrier of the kernel): rebuild the transposed one
        dim3 gt((unsigned)ceil_div(h->VL, 128), (unsigned)ceil_div(h->H, 64));
        transpose_bf16_wide_kernel<<<gt, 256, 0, h->stream>>>(h->Wb, h->ldv, (int)h->H, (int)h->VL, h->WbT, h->ldh);
        QB_CUDA(cudaGetLastError());
        h->launches += 1;
        h->master_sharded = true;
    } else if (use_vel) QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<true>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
    else QB_CUDA(cudaLaunchCooperativeKernel((void *)p2p_reduce_update_kernel<false>, dim3(p2p_grid), dim3(256), args, 0, h->stream));
    if (timed) QB_CUDA(cudaEventRecord(h->upd_ev[ei + 1], h->stream));
    h->launches += 1;
    h->w3_dirty = true;
}

// Option "nvls_exchange" = 1 (COLLECTIVE: every rank, after qb_comm_init, between steps): move the gradient, the row-major
// bf16 shadow and the barrier flags into NCCL symmetric windows and switch the data-parallel exchange to the NVLS variant of
// p2p_reduce_update_kernel.  Throws QB_E_UNSUPPORTED with the reason when multicast is not available (nothing changes then).
void nvls_enable(qb_handle *h) {
    if (h->nvls) return;
    QB_CHECK(h->comm != nullptr, QB_E_STATE, "nvls_exchange needs a communicator: call qb_comm_init first");
    QB_CHECK(h->bf && h->cfg.world > 1 && h->cfg.world <= QB_MAX_PEERS, QB_E_UNSUPPORTED, "nvls_exchange needs bf16 precision and 2..8 ranks");
    QB_CHECK(!h->p2p && !h->stale_chains && !h->state_exchange, QB_E_UNSUPPORTED, "nvls_exchange excludes qb_p2p_import, stale_chains and state_exchange");
    QB_CHECK((h->H % h->cfg.world) == 0, QB_E_UNSUPPORTED, "nvls_exchange needs n_hidden to be a multiple of the number of ranks (equal shards of hidden rows)");
    QB_CHECK(h->nccl_mem.empty(), QB_E_UNSUPPORTED, "nvls_exchange and QB_NCCL_REGISTER are mutually exclusive");
    QB_CUDA(cudaStreamSynchronize(h->stream));
    QbNvlsBuffers io{};
    io.n = 4;
    io.bytes[0] = (size_t)h->nP * sizeof(float);
    io.bytes[1] = (size_t)h->H * h->ldv * sizeof(bf16);
    io.bytes[2] = 4096;
    io.bytes[3] = (size_t)h->H * h->ldv * sizeof(unsigned short);   // negative products of the split exchange
    const char *why = qb_nvls_setup(nccl().lib, h->comm, &io);
    if (why) throw Error(QB_E_UNSUPPORTED, std::string("NVSwitch multicast exchange unavailable: ") + why);
    QB_CHECK(io.lsa_rank == h->cfg.rank && io.lsa_size == h->cfg.world, QB_E_NCCL, "NCCL's load/store team does not match the ranks of this job");
    if (cudaMemcpy(io.local[1], h->Wb, io.bytes[1], cudaMemcpyDeviceToDevice) != cudaSuccess) {
        qb_nvls_teardown(io.state);
        throw Error(QB_E_CUDA, "copying the shadow into the symmetric window failed");
    }
    cudaFree(h->G); cudaFree(h->Wb);
    h->G = (float *)io.local[0]; h->Wb = (bf16 *)io.local[1];
    h->G_mc = (const float *)io.mc[0]; h->Wb_mc = (bf16 *)io.mc[1];
    h->Gn16 = (unsigned short *)io.local[3]; h->Gn16_mc = (const unsigned short *)io.mc[3];
    for (int r = 0; r < h->cfg.world; r++) {
        h->peerG[r] = (float *)io.peer[0][r]; h->peerWb[r] = (bf16 *)io.peer[1][r]; h->peerWbT[r] = nullptr;
        h->peerFlags[r] = (unsigned int *)io.peer[2][r]; h->peerP[r] = nullptr;
    }
    if (!h->p2p_grid_bar) { h->p2p_grid_bar = dalloc<unsigned int>(1); h->p2p_ns = dalloc<unsigned long long>(2); }
    h->shard_rows = h->H / h->cfg.world;
    h->bias_grads_clean = false;   // the fresh gradient buffer is zero everywhere, but the flag's other invariants (datasum) are per step
    h->nvls_state = io.state;
    h->nvls = true; h->p2p = true; h->sharded_update = true;
}

// p2p mode: the fp32 master is sharded by hidden rows; pull the other ranks' shards (collective by convention:
// every rank is between steps when any rank calls this)
void p2p_gather_master(qb_handle *h) {
    if (h->master_sharded && h->comm && (!h->p2p || h->nvls)) {  // NCCL sharded update / NVLS kernel (equal shards): in-place all-gather of the owned hidden rows
        const int64_t sr = h->H / h->cfg.world;
        QB_NCCL(nccl().AllGather(h->P + sr * h->cfg.rank * h->ldv, h->P, (size_t)(sr * h->ldv), /*ncclFloat32*/ 7, h->comm, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
        h->master_sharded = false;
        return;
    }
    if (!h->p2p || h->nvls) return;
    QB_CUDA(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->cfg.world; r++) {
        if (r == h->cfg.rank) continue;
        const int64_t b = std::min<int64_t>(h->H, (int64_t)r * h->shard_rows), e = std::min<int64_t>(h->H, (int64_t)(r + 1) * h->shard_rows);
        if (e > b)
            QB_CUDA(cudaMemcpyAsync(h->P + b * h->ldv, h->peerP[r] + b * h->ldv, (size_t)(e - b) * h->ldv * sizeof(float),
                                    cudaMemcpyDefault, h->stream));
    }
    QB_CUDA(cudaStreamSynchronize(h->stream));
}

// stale-chains mode: make the asynchronously updated weight version the current one
void adopt_pending_update(qb_handle *h) {
    if (!h->pending_update) return;
    QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_upd, 0));
    std::swap(h->Wb, h->Wb_alt);
    std::swap(h->WbT, h->WbT_alt);
    h->bias_cur ^= 1;
    h->bias_rd = h->bias_buf[h->bias_cur];
    h->pending_update = false;
}

// ... or only order the compute stream after it (reads of the fp32 masters, timers), keeping the
// chains one version behind
void wait_pending_update(qb_handle *h) {
    if (h->pending_update) QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_upd, 0));
}

// second stream + events + (when a communicator exists) a communicator limited to QB_OVERLAP_CTAS CTAs, for exchanges that
// run under GEMMs.  Collective when it has to split the communicator.
void ensure_overlap_resources(qb_handle *h) {
    if (!h->comm_stream) {
        QB_CUDA(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        QB_CUDA(cudaEventCreateWithFlags(&h->ev_grad, cudaEventDisableTiming));
        QB_CUDA(cudaEventCreateWithFlags(&h->ev_upd, cudaEventDisableTiming));
    }
    if (h->comm && !h->comm_overlap && nccl().CommSplit) {
        NcclConfig218 c{};
        c.size = sizeof(c); c.magic = 0xcafebeef; c.version = 21800;
        c.blocking = INT_MIN; c.cgaClusterSize = INT_MIN; c.minCTAs = INT_MIN; c.maxCTAs = QB_OVERLAP_CTAS; c.netName = nullptr;
        c.splitShare = INT_MIN;
        QB_NCCL(nccl().CommSplit(h->comm, 0, h->cfg.rank, &h->comm_overlap, &c));
    }
}

void set_stale_mode(qb_handle *h, bool on) {
    QB_CHECK(!on || (h->bf && h->L == 0), QB_E_UNSUPPORTED, "stale_chains needs bf16 precision and a model without label units");
    QB_CHECK(!on || !h->p2p, QB_E_UNSUPPORTED, "stale_chains uses the NCCL path: do not import peer memory (qb_p2p_import) on this handle");
    adopt_pending_update(h);
    QB_CUDA(cudaStreamSynchronize(h->stream));
    if (on && !h->Wb_alt) {
        h->Wb_alt = dalloc<bf16>(h->H * h->ldv);
        h->WbT_alt = dalloc<bf16>(h->VL * h->ldh);
        h->G_alt = dalloc<float>(h->nP);
        h->bias_buf[0] = dalloc<float>(h->ldv + h->ldh);
        h->bias_buf[1] = dalloc<float>(h->ldv + h->ldh);
    }
    if (on) ensure_overlap_resources(h);  // collective: every rank switches the mode at the same point
    h->gemm_sms = ((on || h->state_exchange) && h->comm_overlap) ? std::max(2, (h->num_sms - QB_OVERLAP_CTAS) & ~1) : h->num_sms;
    if (on) {
        h->bias_cur = 0;
        QB_CUDA(cudaMemcpyAsync(h->bias_buf[0], acat(h), (h->ldv + h->ldh) * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
        h->bias_rd = h->bias_buf[0];
    } else {
        h->bias_rd = nullptr;
    }
    h->stale_chains = on;
}

void refresh_shadows(qb_handle *h) {
    h->w3_dirty = true;
    if (!h->bf) return;
    dim3 g((unsigned)ceil_div(h->VL, 64), (unsigned)ceil_div(h->H, 64));
    shadow_kernel<<<g, 256, 0, h->stream>>>(Wf(h), (int)h->H, (int)h->VL, h->ldv, h->Wb, h->WbT, h->ldh);
    QB_CUDA(cudaGetLastError());
    h->launches++;
}

// transposed copies of every resident mini-batch for batch size B (gradient GEMM operand), built lazily
void ensure_dataT(qb_handle *h, int64_t B) {
    if (!h->bf || h->dataT_B == B) return;
    cudaFree(h->dataT);
    h->dataT = nullptr;
    const int64_t nb = h->N / B;
    h->dataT_ld = round_up(B, 8);
    h->dataT = dalloc<bf16>((size_t)std::max<int64_t>(nb, 1) * h->VL * h->dataT_ld);
    cudaFree(h->datasum);
    h->datasum = dalloc<float>((size_t)std::max<int64_t>(nb, 1) * h->ldv);
    for (int64_t m = 0; m < nb; m++) {
        transpose_into(h, h->data.b + m * B * h->data.ld, h->data.ld, B, h->VL, h->dataT + m * h->VL * h->dataT_ld, h->dataT_ld);
        colsum<bf16>(h, h->data.b + m * B * h->data.ld, h->data.ld, B, h->VL, 1.f, h->datasum + m * h->ldv);
    }
    h->dataT_B = B; h->dataT_nb = nb;
}

struct StepRng {  // per Gibbs step s: hidden uniforms, visible uniforms / normals (labels ride on the visible draw)
    qb_handle *h; int64_t B; bool inj;
    Rng hidden(int64_t s) { return inj ? injected_rng(h->injUh + s * B * h->H, h->H) : philox_rng(h); }
    Rng visible(int64_t s) {
        if (!inj) return philox_rng(h);
        if (h->gaussian && h->L == 0) return injected_rng(h->injZv + s * B * h->V, h->V);
        // Gaussian classifier: normals for visibles and uniforms for labels live in one [B][VL] array
        return injected_rng(h->injUv + s * B * h->VL, h->VL);
    }
};

void check_step_args(qb_handle *h, int64_t B, int64_t k, bool allow_k0 = false) {
    QB_CHECK(B >= 1 && B <= h->Bmax, QB_E_ARG, "batch rows must be in [1, max_batch]");
    QB_CHECK(k >= (allow_k0 ? 0 : 1), QB_E_ARG, "gibbs steps must be >= 1 (>= 0 for qb_pcd_step: externally supplied negative samples)");
    QB_CHECK(h->cfg.world == 1 || (B % 4) == 0, QB_E_ARG, "multi-GPU requires local batch rows % 4 == 0");
    if (h->inj_active) QB_CHECK(h->inj_k >= k, QB_E_ARG, "injected streams hold fewer Gibbs steps than the step needs");
}

// k Gibbs steps on the persistent chains (_update_fantasy_data!, src/fantasy_data.jl:12-29)
void advance_chains(qb_handle *h, int64_t B, int64_t k, StepRng &sr, bool fold_bias) {
    const int64_t row0 = row0_of(h, B);
    ChainRegion region(h);   // bf16: the 2k half-steps run as one persistent kernel (sm100_chain.cuh)
    for (int64_t s = 0; s < k; s++) {
        const bool last = (s == k - 1);
        HiddenOut ho; ho.state = &h->ch; ho.state_rm = true; ho.state_T = last;
        if (last && fold_bias) { ho.bg_what = 2; ho.bg_sign = -1.f; }
        op_hidden(h, h->cv, B, h->VL, ho, sr.hidden(s), row0);
        VisibleOut vo; vo.state = &h->cv; vo.sample = true; vo.state_rm = true; vo.state_T = last;
        vo.bg_sign_neg = (last && fold_bias) ? 1 : 0;
        op_visible(h, h->ch, B, vo, sr.visible(s), row0);
    }
    h->chains_binary = true;
}

// bf16 path, classifier PCD before the first chain update: chains still hold the raw U[0,1) floats of
// _init_fantasy_data, whose column sums were never folded by an epilogue.
void fold_chain_bias_explicit(qb_handle *h, int64_t B) {
    colsum<bf16>(h, h->cv.b, h->cv.ld, B, h->VL, -1.f, G_a(h));
    colsum<bf16>(h, h->ch.b, h->ch.ld, B, h->H, -1.f, G_b(h));
}

// stale-chains variant of the PCD step (see qb_handle::stale_chains)
void pcd_step_stale(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr) {
    const int64_t Nc = h->n_chains;
    const float ps = (float)((double)Nc / (double)B);
    QB_CHECK(!h->inj_active || h->inj_B == Nc, QB_E_ARG, "injected streams must have one row per chain");
    StepRng sr{h, Nc, h->inj_active};
    const int64_t row0 = row0_of(h, B);
    zero_bias_grads(h, ps);
    advance_chains(h, Nc, k, sr, true);  // previous weight version; overlaps the in-flight allreduce + update
    adopt_pending_update(h);             // from here on: the latest weights
    HiddenOut ho; ho.prob = &h->phd; ho.prob_T = true; ho.bg_what = 1; ho.bg_sign = 1.f; ho.pscale = ps;
    op_hidden(h, xb, B, h->VL, ho, Rng(), row0);
    op_grad(h, h->phd, xb, xbT, xb_ldT, h->ch, h->cv, B, Nc, ps);
    QB_CUDA(cudaEventRecord(h->ev_grad, h->stream));
    QB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_grad, 0));
    op_update_on(h, lr, llr, Nc, h->comm_stream, h->G, h->Wb_alt, h->WbT_alt, h->bias_buf[h->bias_cur ^ 1]);
    QB_CUDA(cudaEventRecord(h->ev_upd, h->comm_stream));
    h->pending_update = true;
    std::swap(h->G, h->G_alt);  // the next step accumulates into the other gradient buffer
    h->inj_active = false;
}

// Move the gradient and the row-major bf16 shadow into ncclMemAlloc memory and register them with the communicator, so
// that the in-place reduce-scatter / all-gather of the sharded update can take NCCL's registered-buffer (zero-copy, NVLS)
// paths.  Best effort: any failure leaves the plain cudaMalloc buffers in place.
void register_exchange_buffers(qb_handle *h) {
    Nccl &n = nccl();
    if (!n.MemAlloc || !n.MemFree || !n.CommRegister || !n.CommDeregister) return;
    auto rehome = [&](void **slot, size_t bytes) {
        void *fresh = nullptr;
        if (n.MemAlloc(&fresh, bytes) != 0 || !fresh) return;
        if (cudaMemcpy(fresh, *slot, bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) { n.MemFree(fresh); cudaGetLastError(); return; }
        void *reg = nullptr;
        if (n.CommRegister(h->comm, fresh, bytes, &reg) != 0) { n.MemFree(fresh); return; }
        cudaFree(*slot);
        *slot = fresh;
        h->nccl_mem.push_back(fresh);
        if (reg) h->nccl_reg.push_back(reg);
    };
    QB_CUDA(cudaStreamSynchronize(h->stream));
    rehome((void **)&h->G, (size_t)h->nP * sizeof(float));
    rehome((void **)&h->Wb, (size_t)h->H * h->ldv * sizeof(bf16));
}

// first data-parallel step with the sharded update (after a possible qb_p2p_import, whose CUDA IPC export needs plain
// cudaMalloc memory): move the exchange buffers into registered NCCL memory
void maybe_register_exchange_buffers(qb_handle *h) {
    if (h->reg_tried || !h->comm || !h->sharded_update || h->p2p) return;
    h->reg_tried = true;
    // measured on 2 and 8 B200: no difference to unregistered buffers at these message sizes, so it stays opt-in
    if (getenv("QB_NCCL_REGISTER")) register_exchange_buffers(h);
}

// "state_exchange" variant of the multi-GPU PCD step (see qb_handle::state_exchange).  Same arithmetic as pcd_step_impl: all
// GEMMs of the step use W_t, so the positive phase may run before the chain advance; G = sum_ranks(X' Hd) - Cv_all' Ch_all.
void pcd_step_state_exchange(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr) {
    const int world = h->cfg.world, rank = h->cfg.rank;
    const int64_t sr = h->H / world, hr0 = sr * rank, Kall = (int64_t)world * B, units = h->VL + h->H, bpr = B / 8;
    if (h->sx_rows != B) {
        cudaFree(h->bits_loc); cudaFree(h->bits_all); cudaFree(h->cvT_all); cudaFree(h->chT_all);
        h->bits_loc = dalloc<uint8_t>(units * bpr); h->bits_all = dalloc<uint8_t>((size_t)world * units * bpr);
        h->cvT_all = dalloc<bf16>(h->VL * Kall); h->chT_all = dalloc<bf16>(sr * Kall);
        h->sx_rows = B;
    }
    QB_CHECK(!h->inj_active || h->inj_B == B, QB_E_ARG, "injected streams must have one row per chain");
    StepRng sr_rng{h, B, h->inj_active};
    const int64_t row0 = row0_of(h, B);
    zero_bias_grads(h, 1.f);
    // positive phase and positive product first; its reduce-scatter then runs behind the chain advance
    HiddenOut ho; ho.prob = &h->phd; ho.prob_T = true; ho.bg_what = 1; ho.bg_sign = 1.f;
    op_hidden(h, xb, B, h->VL, ho, Rng(), row0);
    {
        sm100::Operand A{h->phd.bT, h->H, B, h->phd.ldT}, Bo{xbT, h->VL, B, xb_ldT};
        sm100::EpiParams ep{};
        ep.mode = sm100::EPI_GRAD; ep.M = (int)h->H; ep.N = (int)h->VL; ep.out_f32 = G_W(h); ep.ld_out = h->ldv;
        GemmScope gs(h, 2.0 * (double)h->H * (double)h->VL * (double)B);
        sm100::launch(h->stream, A, Bo, nullptr, nullptr, ep, h->gemm_sms, h->force_bn, h->use_2cta);
    }
    if (!h->cur_datasum) colsum<bf16>(h, xb.b, xb.ld, B, h->VL, 1.f, G_a(h));
    QB_CUDA(cudaEventRecord(h->ev_grad, h->stream));
    QB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_grad, 0));
    QB_NCCL(nccl().ReduceScatter(h->G, h->G + hr0 * h->ldv, (size_t)(sr * h->ldv), /*ncclFloat32*/ 7, /*ncclSum*/ 0,
                                 h->comm_overlap ? h->comm_overlap : h->comm, h->comm_stream));
    QB_CUDA(cudaEventRecord(h->ev_upd, h->comm_stream));
    advance_chains(h, B, k, sr_rng, true);
    // chain states of every rank, as bits: [visible units | hidden units] x B / 8 bytes per rank
    pack_state_bits_kernel<<<grid1d(h->VL * bpr), 256, 0, h->stream>>>(h->cv.bT, h->cv.ldT, (int)h->VL, (int)B, h->bits_loc);
    pack_state_bits_kernel<<<grid1d(h->H * bpr), 256, 0, h->stream>>>(h->ch.bT, h->ch.ldT, (int)h->H, (int)B, h->bits_loc + h->VL * bpr);
    QB_CUDA(cudaGetLastError());
    QB_NCCL(nccl().GroupStart());
    QB_NCCL(nccl().AllGather(h->bits_loc, h->bits_all, (size_t)(units * bpr), /*ncclUint8*/ 1, h->comm, h->stream));
    QB_NCCL(nccl().AllReduce(h->G + h->nW, h->G + h->nW, (size_t)(h->ldv + h->ldh), 7, 0, h->comm, h->stream));
    QB_NCCL(nccl().GroupEnd());
    unpack_state_bits_kernel<<<grid1d(h->VL * world * bpr), 256, 0, h->stream>>>(h->bits_all, world, (int)units, 0, (int)h->VL, (int)B,
                                                                                 h->cvT_all, Kall);
    unpack_state_bits_kernel<<<grid1d(sr * world * bpr), 256, 0, h->stream>>>(h->bits_all, world, (int)units, (int)(h->VL + hr0), (int)sr,
                                                                              (int)B, h->chT_all, Kall);
    QB_CUDA(cudaGetLastError());
    h->launches += 4;
    QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_upd, 0));  // the reduced positive part of the owned hidden rows has arrived
    {   // owned shard of the negative product over the chains of ALL ranks, subtracted in place
        sm100::Operand A{h->chT_all, sr, Kall, Kall}, Bo{h->cvT_all, h->VL, Kall, Kall};
        sm100::EpiParams ep{};
        ep.mode = sm100::EPI_GRAD; ep.M = (int)sr; ep.N = (int)h->VL; ep.out_f32 = G_W(h) + hr0 * h->ldv; ep.ld_out = h->ldv; ep.grad_sub = 1;
        GemmScope gs(h, 2.0 * (double)sr * (double)h->VL * (double)Kall);
        sm100::launch(h->stream, A, Bo, nullptr, nullptr, ep, h->gemm_sms, h->force_bn, h->use_2cta);
    }
    h->exchange_done = true;  // op_update_on: gradient shard and bias gradients are already reduced
    op_update(h, lr, llr, B);
    h->exchange_done = false;
    h->inj_active = false;
}

bool split_exchange_wanted(qb_handle *h, int64_t B, int64_t k, const bf16 *xbT, cudaEvent_t *ev);
void pcd_step_split(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr);

void pcd_step_impl(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr,
                   cudaEvent_t *ev) {
    QB_CHECK(h->n_chains >= 1, QB_E_STATE, "qb_init_chains has not been called");
    maybe_register_exchange_buffers(h);
    if (h->stale_chains) { pcd_step_stale(h, xb, xbT, xb_ldT, B, k, lr, llr); return; }
    if (split_exchange_wanted(h, B, k, xbT, ev)) { pcd_step_split(h, xb, xbT, xb_ldT, B, k, lr, llr); return; }
    if (h->state_exchange && h->comm && h->sharded_update && !h->p2p && h->L == 0 && !h->gaussian && k > 0 && h->n_chains == B &&
        (B % 8) == 0 && xbT != nullptr && ev == nullptr) {
        pcd_step_state_exchange(h, xb, xbT, xb_ldT, B, k, lr, llr);
        return;
    }
    // Nc chains; Nc == B is the reference (train_pcd.jl:162).  Nc != B (declared extension): all chains advance,
    // negative statistics are averaged over all of them: G = (Nc/B) sum_data - sum_chains, applied with lr / Nc.
    const int64_t Nc = h->n_chains;
    const float ps = (float)((double)Nc / (double)B);
    QB_CHECK(!h->inj_active || h->inj_B == Nc, QB_E_ARG, "injected streams must have one row per chain");
    QB_CHECK(h->cfg.world == 1 || (Nc % 4) == 0, QB_E_ARG, "multi-GPU requires local chains % 4 == 0");
    StepRng sr{h, Nc, h->inj_active};
    const int64_t row0 = row0_of(h, B);
    zero_bias_grads(h, ps);
    const bool clf = h->L > 0;
    if (ev) QB_CUDA(cudaEventRecord(ev[0], h->stream));
    // bf16: the chain advance and the positive phase (independent of each other) are recorded into one fused launch; its tiles
    // come last in the schedule and fill the tail of the chain.  With phase timing (ev) the phases stay separate launches.
    const bool fuse_pos = h->bf && !clf && k > 0 && ev == nullptr;
    if (fuse_pos) chain_begin(h);
    if (!clf && k > 0) advance_chains(h, Nc, k, sr, h->bf);  // chains FIRST: train_pcd.jl:14-16
    else if (h->bf) {
        // classifier: gradient uses the chains as they are (train_pcd.jl:57-86); their transposes and
        // column sums come from the previous step's epilogues, except before the very first update.
        // k == 0: the negative samples were supplied by the caller (persistent_qubo!, train_quantum_pcd.jl:10-37)
        if (!h->chains_binary || k == 0) {
            transpose_into(h, h->cv.b, h->cv.ld, Nc, h->VL, h->cv.bT, h->cv.ldT);
            transpose_into(h, h->ch.b, h->ch.ld, Nc, h->H, h->ch.bT, h->ch.ldT);
        }
        fold_chain_bias_explicit(h, Nc);
    }
    if (ev) QB_CUDA(cudaEventRecord(ev[1], h->stream));
    HiddenOut ho; ho.prob = &h->phd; ho.prob_T = true; ho.bg_what = 1; ho.bg_sign = 1.f; ho.pscale = ps;  // conditional_prob_h(v_data[, y_data])
    const bool fused_recon = h->step_recon_wanted && !clf && !h->gaussian && Nc == B;
    if (fused_recon) ho.prob_rm = true;  // the reconstruction below reads p(h | x) row-major
    try { op_hidden(h, xb, B, h->VL, ho, Rng(), row0); } catch (...) { if (fuse_pos) { h->rec_depth = 0; h->rec->prog.n_steps = 0; } throw; }
    if (fuse_pos) chain_end(h);
    if (fused_recon) {  // reconstruct(x) = sigma(a + W p(h | x)) against x, squared error summed into dscal[0] (src/rbm.jl:220-224)
        QB_CUDA(cudaMemsetAsync(h->dscal, 0, sizeof(double), h->stream));
        VisibleOut vo; vo.state = &h->vm; vo.sample = false; vo.state_rm = !h->bf; vo.ref = &xb; vo.sqerr = h->dscal;
        op_visible(h, h->phd, B, vo, Rng(), row0);
        h->step_recon_done = true;
    }
    if (ev) QB_CUDA(cudaEventRecord(ev[2], h->stream));
    op_grad(h, h->phd, xb, xbT, xb_ldT, h->ch, h->cv, B, Nc, ps);
    op_update(h, lr, llr, Nc);
    if (ev) QB_CUDA(cudaEventRecord(ev[3], h->stream));
    if (clf && k > 0) advance_chains(h, Nc, k, sr, false);  // classifier: chains AFTER the update, train_pcd.jl:89-92
    if (ev) QB_CUDA(cudaEventRecord(ev[4], h->stream));
    h->inj_active = false;
}

// Split gradient exchange (see p2p_kernels.cuh): same arithmetic as pcd_step_impl on a binary model without label units --
// every GEMM of the step uses W_t, so the positive phase and its product may run before the chain advance.
bool split_exchange_wanted(qb_handle *h, int64_t B, int64_t k, const bf16 *xbT, cudaEvent_t *ev) {
    if (!h->nvls || h->split_exchange_mode == 0 || !h->bf || h->L > 0 || h->gaussian || k < 1 || h->n_chains != B || xbT == nullptr || ev != nullptr ||
        h->step_recon_wanted || h->inj_active || h->stale_chains || (int64_t)h->cfg.world * B >= 65536)
        return false;
    if (h->split_exchange_mode == 2) return true;
    // automatic: the background reduction needs block slots beside the chain kernel -- plenty when a half-step is at most one
    // wave of pair tiles minus a few pairs (1024 rows of cfg5: 64 tiles on 74 pairs); a fuller machine would only delay it
    const int64_t tiles = ceil_div(B, 256) * ceil_div(std::max(h->H, h->VL), 256);
    return tiles <= h->num_sms / 2 - 8;
}
void pcd_step_split(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr) {
    if (!h->comm_stream) {   // second stream + events for the background reduction
        QB_CUDA(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        QB_CUDA(cudaEventCreateWithFlags(&h->ev_grad, cudaEventDisableTiming));
        QB_CUDA(cudaEventCreateWithFlags(&h->ev_upd, cudaEventDisableTiming));
    }
    StepRng sr{h, B, false};
    const int64_t row0 = row0_of(h, B);
    zero_bias_grads(h, 1.f);
    // positive phase and positive product first ...
    HiddenOut ho; ho.prob = &h->phd; ho.prob_T = true; ho.bg_what = 1; ho.bg_sign = 1.f;
    op_hidden(h, xb, B, h->VL, ho, Rng(), row0);
    {
        sm100::Operand A{h->phd.bT, h->H, B, h->phd.ldT}, Bo{xbT, h->VL, B, xb_ldT};
        sm100::EpiParams ep{};
        ep.mode = sm100::EPI_GRAD; ep.M = (int)h->H; ep.N = (int)h->VL; ep.out_f32 = G_W(h); ep.ld_out = h->ldv;
        GemmScope gs(h, 2.0 * (double)h->H * (double)h->VL * (double)B);
        sm100::launch(h->stream, A, Bo, nullptr, nullptr, ep, h->gemm_sms, h->force_bn, h->use_2cta);
    }
    if (!h->cur_datasum) colsum<bf16>(h, xb.b, xb.ld, B, h->VL, 1.f, G_a(h));
    // ... its reduction runs behind the chain advance
    QB_CUDA(cudaEventRecord(h->ev_grad, h->stream));
    QB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_grad, 0));
    split_reduce_positive(h, h->comm_stream);
    QB_CUDA(cudaEventRecord(h->ev_upd, h->comm_stream));
    advance_chains(h, B, k, sr, true);
    {   // negative product of this rank's chains: exact integer counts, stored as uint16
        sm100::Operand A{h->ch.bT, h->H, B, h->ch.ldT}, Bo{h->cv.bT, h->VL, B, h->cv.ldT};
        sm100::EpiParams ep{};
        ep.mode = sm100::EPI_GRAD; ep.M = (int)h->H; ep.N = (int)h->VL; ep.out_u16 = h->Gn16; ep.ld_out = h->ldv;
        GemmScope gs(h, 2.0 * (double)h->H * (double)h->VL * (double)B);
        sm100::launch(h->stream, A, Bo, nullptr, nullptr, ep, h->gemm_sms, h->force_bn, h->use_2cta);
    }
    QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_upd, 0));   // the reduced positive product of the owned rows is in place
    p2p_update(h, lr, llr, B, /*split*/ true);
    h->inj_active = false;
}

void cd_step_impl(qb_handle *h, const Mat &xb, const bf16 *xbT, int64_t xb_ldT, int64_t B, int64_t k, double lr, double llr,
                  cudaEvent_t *ev) {
    QB_CHECK(!h->inj_active || h->inj_B == B, QB_E_ARG, "injected streams must have one row per mini-batch row");
    maybe_register_exchange_buffers(h);
    StepRng sr{h, B, h->inj_active};
    const int64_t row0 = row0_of(h, B);
    zero_bias_grads(h);
    if (ev) QB_CUDA(cudaEventRecord(ev[0], h->stream));
    // positive phase fused with the first hidden sample of the chain (same p(h | v_data[, y_data]))
    HiddenOut h0; h0.prob = &h->phd; h0.prob_T = true; h0.bg_what = 1; h0.bg_sign = 1.f; h0.state = &h->hs; h0.state_rm = true;
    {
    ChainRegion region(h);   // bf16: the 2k+1 forward contractions run as one persistent kernel (with phase timing: per launch)
    op_hidden(h, xb, B, h->VL, h0, sr.hidden(0), row0);
    if (ev) { chain_flush(h); QB_CUDA(cudaEventRecord(ev[1], h->stream)); }
    for (int64_t s = 0; s < k; s++) {
        const bool last = (s == k - 1);
        if (s > 0) {
            HiddenOut hs; hs.state = &h->hs; hs.state_rm = true;
            op_hidden(h, h->vm, B, h->VL, hs, sr.hidden(s), row0);
        }
        VisibleOut vo; vo.state = &h->vm; vo.sample = true; vo.state_rm = true; vo.state_T = last; vo.bg_sign_neg = (last && h->bf) ? 1 : 0;
        op_visible(h, h->hs, B, vo, sr.visible(s), row0);
    }
    // h_model = conditional_prob_h(rbm, v_model): probabilities, 2-arg form (ignores y_model, train_cd.jl:34)
    HiddenOut hm; hm.prob = &h->phm; hm.prob_T = true; hm.bg_what = 1; hm.bg_sign = -1.f;
    op_hidden(h, h->vm, B, h->V, hm, Rng(), row0);
    }
    if (ev) QB_CUDA(cudaEventRecord(ev[2], h->stream));
    op_grad(h, h->phd, xb, xbT, xb_ldT, h->phm, h->vm, B, B, 1.f);
    op_update(h, lr, llr, B);
    if (ev) QB_CUDA(cudaEventRecord(ev[3], h->stream));
    h->inj_active = false;
}

// ----------------------------------------------------------------------------- persistent multi-mini-batch kernel
void ensure_dataT(qb_handle *h, int64_t B);
void ensure_chain_flags(qb_handle *h);

// describe one half-step through the regular op_* code (nothing is launched)
HalfCapture describe_hidden(qb_handle *h, const Mat &in, int64_t rows, int64_t kdim, const HiddenOut &o, bool sample) {
    HalfCapture c{}; h->capture = &c;
    Rng r; if (sample) { r.use_rng = 1; r.key.seed_lo = (uint32_t)h->seed; r.key.seed_hi = (uint32_t)(h->seed >> 32); }
    try { op_hidden(h, in, rows, kdim, o, r, 0); } catch (...) { h->capture = nullptr; throw; }
    h->capture = nullptr;
    return c;
}
HalfCapture describe_visible(qb_handle *h, const Mat &in, int64_t rows, const VisibleOut &o) {
    HalfCapture c{}; h->capture = &c;
    Rng r; r.use_rng = 1; r.key.seed_lo = (uint32_t)h->seed; r.key.seed_hi = (uint32_t)(h->seed >> 32);
    try { op_visible(h, in, rows, o, r, 0); } catch (...) { h->capture = nullptr; throw; }
    h->capture = nullptr;
    return c;
}

// n_steps consecutive mini-batch steps on resident batches (first + i) mod nb in ONE launch.  Returns false (nothing done)
// when the configuration is outside what the kernel covers; the caller then runs the steps one launch sequence at a time.
// (measured, tools/epoch_probe.py: 32-row tiles beat 64-row tiles at 256 rows -- 25.7 vs 31.6 us per cfg2 step)
int epoch_tile_width(const qb_handle *h, int64_t B) { return h->force_bn ? h->force_bn : (B <= 256 ? 32 : 64); }

// the cheap part of the eligibility test (the per-half-step epilogue check happens while the program is built)
bool epoch_kernel_eligible(const qb_handle *h, bool pcd, int64_t B, int64_t k) {
    if (h->epoch_kernel_mode == 0 || !h->bf || h->cfg.world != 1 || h->stale_chains || h->inj_active || h->time_gemms) return false;
    if (h->momentum != 0.0 || h->weight_decay != 0.0 || h->use_2cta == 2 || (h->force_bn != 0 && h->force_bn != 32 && h->force_bn != 64)) return false;
    if (h->N < B || k < 1 || 2 * k + 2 > sm100::EPOCH_MAX_STEPS) return false;
    if (pcd && (h->L > 0 || h->n_chains != B)) return false;
    if (h->gaussian && h->L > 0) return false;   // Gaussian visibles + label units: generic epilogue only
    // automatic: models whose half-steps are a fraction of one wave of 128 x BN tiles -- there the launch latencies of the
    // step-by-step path dominate.  Measured (tools/epoch_probe.py, profiles/r2_epoch_probe.txt): cfg2 -19 %, cfg1 tie; from
    // about one wave on (cfg4 +18 %, cfg3 +8 %) one launch per GEMM is the better schedule.
    const int64_t tiles = ceil_div(std::max(h->H, h->VL), 128) * ceil_div(B, epoch_tile_width(h, B));
    return h->epoch_kernel_mode == 2 || (B <= 256 && 2 * tiles <= h->num_sms);   // (cfg4: 512 rows, 64 tiles: 48.7 vs 41.5 us)
}

bool epoch_kernel_run(qb_handle *h, bool pcd, int64_t first, int64_t n_steps, int64_t B, int64_t k, double lr, double llr) {
    if (n_steps < 1 || !epoch_kernel_eligible(h, pcd, B, k)) return false;
    const int bn = epoch_tile_width(h, B);
    ensure_dataT(h, B);
    const int64_t nb = h->dataT_nb;
    if (nb < 1) return false;
    if (!h->epoch) h->epoch = new sm100::EpochBuilder();
    sm100::EpochBuilder &eb = *h->epoch;
    eb.begin(bn);
    Mat data_all = h->data;   // the whole dataset: the mini-batch is a row offset (b_row_mb = B rows per batch index)
    const sm100::Operand dataT_all{h->dataT, nb * h->VL, B, h->dataT_ld};
    auto half = [&](const HalfCapture &c, int b_row_mb, int dep, int draw_off) -> int {
        if (!c.set || !sm100::chainable(c.ep)) return -2;
        sm100::EpiParams ep = c.ep; ep.N = (int)B;
        sm100::Operand Bop = c.B;
        if (b_row_mb) Bop.rows = h->N;   // view of the whole dataset
        return eb.add_half_step(c.A, Bop, b_row_mb, ep, dep, draw_off);
    };
    int last = -1, d0 = -1, d1 = -1;
    const sm100::Operand *negA = nullptr, *negB = nullptr;
    sm100::Operand phdT{h->phd.bT, h->H, B, h->phd.ldT}, phmT{h->phm.bT, h->H, B, h->phm.ldT}, vmT{h->vm.bT, h->VL, B, h->vm.ldT},
        chT{h->ch.bT, h->H, B, h->ch.ldT}, cvT{h->cv.bT, h->VL, B, h->cv.ldT};
    if (!pcd) {
        HiddenOut h0; h0.prob = &h->phd; h0.prob_T = true; h0.bg_what = 1; h0.bg_sign = 1.f; h0.state = &h->hs; h0.state_rm = true;
        last = half(describe_hidden(h, data_all, B, h->VL, h0, true), (int)B, -1, 0);
        if (last < 0) return false;
        for (int64_t s = 0; s < k; s++) {
            const bool fin = (s == k - 1);
            if (s > 0) {
                HiddenOut hs; hs.state = &h->hs; hs.state_rm = true;
                last = half(describe_hidden(h, h->vm, B, h->VL, hs, true), 0, last, (int)(2 * s));
                if (last < 0) return false;
            }
            VisibleOut vo; vo.state = &h->vm; vo.sample = true; vo.state_rm = true; vo.state_T = fin; vo.bg_sign_neg = fin ? 1 : 0;
            last = half(describe_visible(h, h->hs, B, vo), 0, last, (int)(2 * s + 1));
            if (last < 0) return false;
        }
        HiddenOut hm; hm.prob = &h->phm; hm.prob_T = true; hm.bg_what = 1; hm.bg_sign = -1.f;
        last = half(describe_hidden(h, h->vm, B, h->V, hm, false), 0, last, 0);
        if (last < 0) return false;
        d0 = last; negA = &phmT; negB = &vmT;
    } else {
        for (int64_t s = 0; s < k; s++) {
            const bool fin = (s == k - 1);
            HiddenOut ho; ho.state = &h->ch; ho.state_rm = true; ho.state_T = fin;
            if (fin) { ho.bg_what = 2; ho.bg_sign = -1.f; }
            last = half(describe_hidden(h, h->cv, B, h->VL, ho, true), 0, last, (int)(2 * s));
            if (last < 0) return false;
            VisibleOut vo; vo.state = &h->cv; vo.sample = true; vo.state_rm = true; vo.state_T = fin; vo.bg_sign_neg = fin ? 1 : 0;
            last = half(describe_visible(h, h->ch, B, vo), 0, last, (int)(2 * s + 1));
            if (last < 0) return false;
        }
        d1 = last;
        HiddenOut ho; ho.prob = &h->phd; ho.prob_T = true; ho.bg_what = 1; ho.bg_sign = 1.f;
        d0 = half(describe_hidden(h, data_all, B, h->VL, ho, false), (int)B, -1, 0);
        if (d0 < 0) return false;
        negA = &chT; negB = &cvT;
    }
    eb.add_grad_update(phdT, dataT_all, (int)h->VL, *negA, *negB, (int)h->H, (int)h->VL, d0, d1);
    sm100::EpochProgram &P = eb.prog;
    P.draws_per_iter = (int)(2 * k);
    P.n_batches = (int)nb;
    P.W = h->P; P.Wb = h->Wb; P.WbT = h->WbT; P.ldv = h->ldv; P.ldh = h->ldh; P.V = (int)h->V;
    P.lr = (float)(lr / (double)B); P.llr = (float)(llr / (double)B);
    P.bias_P = acat(h); P.bias_G = G_a(h); P.datasum = h->datasum; P.datasum_ld = h->ldv; P.dscale = 1.f;
    P.n_vis = (int)h->ldv; P.n_bias = (int)(h->ldv + h->ldh);
    ensure_chain_flags(h);
    P.error_flag = h->chain_err_dev;
    P.abort_flag = h->chain_cnt + (size_t)QB_CHAIN_REGIONS * h->chain_cnt_region;
    // the bias-gradient accumulators must start at zero (the kernel leaves them at zero)
    if (!h->bias_grads_clean) {
        const int n = (int)(h->ldv + h->ldh);
        init_bias_grads_kernel<<<grid1d(n), 256, 0, h->stream>>>(G_a(h), nullptr, 1.f, (int)h->ldv, n);
        QB_CUDA(cudaGetLastError());
        h->launches++;
    }
    constexpr int64_t MAX_ITERS = 2048;
    for (int64_t done = 0; done < n_steps; done += MAX_ITERS) {
        const int64_t it = std::min(MAX_ITERS, n_steps - done);
        const size_t need = (size_t)it * P.n_steps * (P.max_tiles_n + 1);
        if (h->epoch_cnt_cap < need) {
            cudaFree(h->epoch_cnt);
            h->epoch_cnt = dalloc<unsigned int>(need);
            h->epoch_cnt_cap = need;
        } else {
            QB_CUDA(cudaMemsetAsync(h->epoch_cnt, 0, need * sizeof(unsigned int), h->stream));
        }
        P.n_iters = (int)it;
        P.batch0 = (int)((first + done) % nb);
        P.draw0 = h->draw;
        P.rb_cnt = h->epoch_cnt;
        P.tot_cnt = h->epoch_cnt + (size_t)it * P.n_steps * P.max_tiles_n;
#ifdef QB_EPOCH_TRACE
        // development builds: per-tile %globaltimer stamps of the first launch that finds QB_EPOCH_TRACE=<file> set
        const char *trace_path = getenv("QB_EPOCH_TRACE");
        const size_t trace_n = (size_t)it * P.tiles_per_iter * 32;
        P.trace = nullptr;
        if (trace_path && it <= 256) { P.trace = dalloc<unsigned long long>(trace_n); QB_CUDA(cudaMemsetAsync(P.trace, 0, trace_n * 8, h->stream)); }
#endif
        {
            GemmScope gs(h, eb.flops_per_iter * (double)it);
            sm100::launch_epoch(h->stream, eb, h->gemm_sms);
            sm100::last_shape().bn = bn; sm100::last_shape().pair = 0; sm100::last_shape().kind = sm100::K_GRAD_UPDATE;
        }
#ifdef QB_EPOCH_TRACE
        if (P.trace) {
            std::vector<unsigned long long> host(trace_n);
            QB_CUDA(cudaStreamSynchronize(h->stream));
            QB_CUDA(cudaMemcpy(host.data(), P.trace, trace_n * 8, cudaMemcpyDeviceToHost));
            if (FILE *f = fopen(trace_path, "wb")) {
                long long hdr[4 + sm100::EPOCH_MAX_STEPS] = {it, P.tiles_per_iter, P.n_steps, std::min(h->gemm_sms, 1 << 30)};
                for (int s2 = 0; s2 < P.n_steps; s2++) hdr[4 + s2] = P.step[s2].tile_base;
                fwrite(hdr, sizeof(hdr), 1, f);
                fwrite(host.data(), 8, trace_n, f);
                fclose(f);
            }
            cudaFree(P.trace); P.trace = nullptr;
        }
#endif
        h->draw += (uint64_t)(2 * k) * (uint64_t)it;
        h->epoch_launches++;
    }
    h->bias_grads_clean = true;
    h->chains_binary = h->chains_binary || pcd;
    h->cur_datasum = nullptr;
    return true;
}

// resident mini-batch `m` of B rows as (row-major view, transposed pointer)
void batch_view(qb_handle *h, int64_t m, int64_t B, Mat &xb, const bf16 *&xbT, int64_t &ldT) {
    QB_CHECK(h->N > 0, QB_E_STATE, "qb_set_data has not been called");
    QB_CHECK(m >= 0 && (m + 1) * B <= h->N, QB_E_ARG, "mini-batch index out of range");
    xb = h->data.row_slice(m * B, B);
    xbT = nullptr; ldT = 0; h->cur_datasum = nullptr;
    if (h->bf) {
        ensure_dataT(h, B);
        xbT = h->dataT + m * h->VL * h->dataT_ld;
        ldT = h->dataT_ld;
        h->cur_datasum = h->datasum + m * h->ldv;
    }
}

void stage_host_batch(qb_handle *h, const void *X, int xdt, const void *Y, int ydt, int64_t B) {
    QB_CHECK(X != nullptr, QB_E_ARG, "X is NULL");
    QB_CHECK(h->L == 0 || Y != nullptr, QB_E_ARG, "labels required for classifier models");
    const bool hit = h->pf_valid && h->pf_X == X && h->pf_B == B && h->pf_xdt == xdt && (h->L == 0 || (h->pf_Y == Y && h->pf_ydt == ydt));
    if (h->pf_valid && !hit) QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pf, 0));   // hb2 / staging2 are still being filled: order the reuse after it
    if (hit) {
        // this batch was uploaded on copy_stream while the previous step was computing
        QB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pf, 0));
        std::swap(h->hb, h->hb2);
    } else {
        upload_rows(h, X, xdt, B, h->V, h->hb, 0, 0);
        if (h->L > 0) upload_rows(h, Y, ydt, B, h->L, h->hb, 0, h->V);
        h->hb.exact = xdt == QB_DT_U8 && (h->L == 0 || ydt == QB_DT_U8);
        if (h->bf) transpose_into(h, h->hb.b, h->hb.ld, B, h->VL, h->hb.bT, h->hb.ldT);
    }
    h->pf_valid = false;
    if (h->next_X) {  // qb_set_next_host_batch: start the next upload now, it overlaps this step's kernels
        if (!h->hb2.b && !h->hb2.f) {
            h->hb2 = alloc_mat(h, h->Bmax, h->VL, true);
            QB_CUDA(cudaEventCreateWithFlags(&h->ev_pf, cudaEventDisableTiming));
        }
        upload_rows_on(h, h->copy_stream, h->staging2, h->staging2_bytes, h->next_X, h->next_xdt, B, h->V, h->hb2, 0, 0);
        if (h->L > 0) upload_rows_on(h, h->copy_stream, h->staging2, h->staging2_bytes, h->next_Y, h->next_ydt, B, h->L, h->hb2, 0, h->V);
        if (h->bf) transpose_on(h, h->copy_stream, h->hb2.b, h->hb2.ld, B, h->VL, h->hb2.bT, h->hb2.ldT);
        QB_CUDA(cudaEventRecord(h->ev_pf, h->copy_stream));
        h->pf_X = h->next_X; h->pf_Y = h->next_Y; h->pf_xdt = h->next_xdt; h->pf_ydt = h->next_ydt; h->pf_B = B; h->pf_valid = true;
        h->next_X = nullptr; h->next_Y = nullptr;
    }
}

// sum over rows of sum_units (x - reconstruct(x))^2 for rows of `x` (chunked by Bmax), accumulated into dscal[0]
void recon_sqerr(qb_handle *h, const Mat &x, int64_t n, const double *Z_host, int64_t global_row0) {
    for (int64_t r = 0; r < n; r += h->Bmax) {
        const int64_t nr = std::min(h->Bmax, n - r);
        Mat xs = x.row_slice(r, nr);
        HiddenOut ho; ho.prob = &h->phd; ho.prob_rm = true;
        op_hidden(h, xs, nr, h->V, ho, Rng(), 0);
        Rng rng;
        if (h->gaussian) {  // reconstruct on a GRBM draws N(0,1) noise per unit (src/rbm.jl:187)
            if (Z_host) {
                Mat zf; zf.f = h->znoise; zf.ld = h->V; zf.rows = nr; zf.units = h->V;
                upload_rows(h, Z_host + (size_t)r * h->V, QB_DT_F64, nr, h->V, zf, 0, 0);
                rng = injected_rng(h->znoise, h->V);
            } else {
                rng = philox_rng(h);
            }
        }
        VisibleOut vo; vo.state = &h->vm; vo.sample = false; vo.state_rm = !h->bf; vo.ref = &xs; vo.sqerr = h->dscal;
        op_visible(h, h->phd, nr, vo, rng, global_row0 + r);
    }
}

double read_scalar(qb_handle *h, int i) {
    double v = 0.0;
    QB_CUDA(cudaMemcpyAsync(&v, h->dscal + i, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QB_CUDA(cudaStreamSynchronize(h->stream));
    return v;
}

template <typename F>
int32_t guarded(F &&fn) {
    try {
        fn();
        return QB_OK;
    } catch (const qb::Error &e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return QB_E_CUDA;
    }
}
#define QB_H_KEEP(h) QB_CHECK((h) != nullptr, QB_E_ARG, "handle is NULL"); QB_CUDA(cudaSetDevice((h)->cfg.device))
#define QB_H(h) QB_H_KEEP(h); adopt_pending_update(h)

}  // namespace

// ============================================================================= C ABI
extern "C" {

int32_t qb_version(void) { return QB_VERSION; }
const char *qb_last_error(void) { return g_last_error.c_str(); }

int32_t qb_device_count(int32_t *count) {
    return guarded([&] {
        QB_CHECK(count != nullptr, QB_E_ARG, "count is NULL");
        int n = 0;
        QB_CUDA(cudaGetDeviceCount(&n));
        *count = n;
    });
}

int32_t qb_create(qb_handle **out, const qb_config *cfg) {
    return guarded([&] {
        QB_CHECK(out && cfg, QB_E_ARG, "NULL argument");
        QB_CHECK(cfg->n_visible > 0 && cfg->n_hidden > 0 && cfg->n_classifiers >= 0, QB_E_ARG, "V, H must be > 0 and L >= 0");
        QB_CHECK(cfg->max_batch > 0, QB_E_ARG, "max_batch must be > 0");
        QB_CHECK(cfg->precision == QB_PREC_FP32 || cfg->precision == QB_PREC_BF16 || cfg->precision == QB_PREC_FP32X3, QB_E_ARG, "unknown precision");
        QB_CHECK(cfg->visible_kind == QB_VISIBLE_BERNOULLI || cfg->visible_kind == QB_VISIBLE_GAUSSIAN, QB_E_ARG, "unknown visible kind");
        QB_CHECK(cfg->world >= 1 && cfg->rank >= 0 && cfg->rank < cfg->world, QB_E_ARG, "bad rank/world");
        int ndev = 0;
        QB_CUDA(cudaGetDeviceCount(&ndev));
        QB_CHECK(ndev > 0, QB_E_CUDA, "no CUDA device visible: qarbom_b200 has no CPU path");
        QB_CHECK(cfg->device >= 0 && cfg->device < ndev, QB_E_ARG, "device ordinal out of range");
        QB_CUDA(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        QB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        QB_CHECK(prop.major == 10, QB_E_CUDA,
                 std::string("qarbom_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
        std::unique_ptr<qb_handle> h(new qb_handle());
        h->cfg = *cfg;
        h->V = cfg->n_visible; h->H = cfg->n_hidden; h->L = cfg->n_classifiers; h->VL = h->V + h->L;
        h->ldv = round_up(h->VL, 8); h->ldh = round_up(h->H, 8);
        h->Bmax = cfg->max_batch; h->ldb = round_up(h->Bmax, 8);
        h->bf = cfg->precision == QB_PREC_BF16;
        h->x3 = cfg->precision == QB_PREC_FP32X3;
        h->gaussian = cfg->visible_kind == QB_VISIBLE_GAUSSIAN;
        h->num_sms = prop.multiProcessorCount;
        h->gemm_sms = h->num_sms;
        QB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        QB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (auto &e : h->ev) QB_CUDA(cudaEventCreate(&e));
        for (auto &e : h->timer_ev) QB_CUDA(cudaEventCreate(&e));
        h->nW = h->H * h->ldv; h->nP = h->nW + h->ldv + h->ldh;
        h->P = dalloc<float>(h->nP); h->G = dalloc<float>(h->nP);
        if (h->x3) sm100::init_device();
        if (h->bf) {
            sm100::init_device();
            h->Wb = dalloc<bf16>(h->H * h->ldv);
            h->WbT = dalloc<bf16>(h->VL * h->ldh);
        }
        h->cv = alloc_mat(h.get(), h->Bmax, h->VL, true);
        h->ch = alloc_mat(h.get(), h->Bmax, h->H, true);
        h->phd = alloc_mat(h.get(), h->Bmax, h->H, true);
        h->hs = alloc_mat(h.get(), h->Bmax, h->H, false);
        h->vm = alloc_mat(h.get(), h->Bmax, h->VL, true);
        h->phm = alloc_mat(h.get(), h->Bmax, h->H, true);
        h->hb = alloc_mat(h.get(), h->Bmax, h->VL, true);
        h->pre_h = dalloc<float>(h->Bmax * h->ldh);
        h->pre_v = dalloc<float>(h->Bmax * h->ldv);
        h->lab_pre = dalloc<float>(h->Bmax * std::max<int64_t>(h->L, 1));
        h->softplus_rows = dalloc<float>(h->Bmax);
        h->out_f32 = dalloc<float>(h->Bmax * std::max(h->ldv, h->ldh));
        h->dscal = dalloc<double>(4);
        if (h->gaussian) h->znoise = dalloc<float>(h->Bmax * h->V);
        QB_CUDA(cudaDeviceSynchronize());
        *out = h.release();
    });
}

int32_t qb_destroy(qb_handle *h) {
    return guarded([&] {
        if (!h) return;
        cudaSetDevice(h->cfg.device);
        cudaDeviceSynchronize();
        auto dfree = [&](void *p) {  // buffers re-homed by register_exchange_buffers belong to NCCL's allocator
            if (p && std::find(h->nccl_mem.begin(), h->nccl_mem.end(), p) != h->nccl_mem.end()) nccl().MemFree(p);
            else cudaFree(p);
        };
        if (h->nvls_state) { qb_nvls_teardown(h->nvls_state); h->nvls_state = nullptr; h->G = nullptr; h->Wb = nullptr; }   // (frees G / Wb / flags windows)
        if (h->comm) for (void *r : h->nccl_reg) nccl().CommDeregister(h->comm, r);
        if (h->comm_overlap) nccl().CommDestroy(h->comm_overlap);
        if (h->comm) nccl().CommDestroy(h->comm);
        if (h->p2p && !h->nvls)
            for (int r = 0; r < h->cfg.world; r++)
                if (r != h->cfg.rank) {
                    cudaIpcCloseMemHandle(h->peerG[r]); cudaIpcCloseMemHandle(h->peerWb[r]); cudaIpcCloseMemHandle(h->peerWbT[r]);
                    cudaIpcCloseMemHandle(h->peerFlags[r]); cudaIpcCloseMemHandle(h->peerP[r]);
                }
        for (int i = 0; i < 3; i++) { cudaFree(h->W3[i]); cudaFree(h->W3T[i]); cudaFree(h->x3_buf[0][i]); cudaFree(h->x3_buf[1][i]); }
        cudaFree(h->epoch_cnt); delete h->epoch;
        cudaFree(h->chain_cnt); if (h->chain_err_host) cudaFreeHost(h->chain_err_host); delete h->rec;
        cudaFree(h->p2p_flags); cudaFree(h->p2p_grid_bar); cudaFree(h->p2p_ns); if (h->p2p_timeout_host) cudaFreeHost(h->p2p_timeout_host);
        cudaFree(h->P); dfree(h->G); cudaFree(h->Vel); cudaFree(h->P_best); cudaFree(h->F); cudaFree(h->F2); cudaFree(h->P_eff); cudaFree(h->W2); cudaFree(h->online_vec); cudaFree(h->online_bar); cudaFree(h->online_ns); dfree(h->Wb); cudaFree(h->WbT);
        cudaFree(h->bits_loc); cudaFree(h->bits_all); cudaFree(h->cvT_all); cudaFree(h->chT_all);
        dfree(h->Wb_alt); cudaFree(h->WbT_alt); dfree(h->G_alt); cudaFree(h->bias_buf[0]); cudaFree(h->bias_buf[1]);
        if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
        if (h->ev_grad) cudaEventDestroy(h->ev_grad);
        if (h->ev_upd) cudaEventDestroy(h->ev_upd);
        free_mat(h->data); cudaFree(h->dataT); cudaFree(h->datasum);
        free_mat(h->cv); free_mat(h->ch); free_mat(h->phd); free_mat(h->hs); free_mat(h->vm); free_mat(h->phm); free_mat(h->hb); free_mat(h->hb2);
        cudaFree(h->staging2); if (h->ev_pf) cudaEventDestroy(h->ev_pf);
        cudaFree(h->pre_h); cudaFree(h->pre_v); cudaFree(h->lab_pre); cudaFree(h->softplus_rows); cudaFree(h->out_f32);
        cudaFree(h->dscal); cudaFree(h->znoise); cudaFree(h->rand_rows); cudaFree(h->staging); cudaFree(h->injUh); cudaFree(h->injUv); cudaFree(h->injZv);
        for (auto &e : h->ev) cudaEventDestroy(e);
        for (auto &e : h->timer_ev) cudaEventDestroy(e);
        for (auto &e : h->gemm_ev) cudaEventDestroy(e);
        for (auto &e : h->epoch_ev) cudaEventDestroy(e);
        for (auto &e : h->upd_ev) cudaEventDestroy(e);
        cudaStreamDestroy(h->stream); cudaStreamDestroy(h->copy_stream);
        delete h;
    });
}

int32_t qb_set_params(qb_handle *h, const double *W, const double *a, const double *b, const double *U, const double *c) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(W && a && b, QB_E_ARG, "W, a, b must not be NULL");
        QB_CHECK(h->L == 0 || (U && c), QB_E_ARG, "U, c required for classifier models");
        // Julia column-major V x H == row-major [H][V]; U (L x H) == [H][L]
        std::vector<float> p((size_t)h->nP, 0.f);
        for (int64_t j = 0; j < h->H; j++) {
            for (int64_t i = 0; i < h->V; i++) p[j * h->ldv + i] = (float)W[i + h->V * j];
            for (int64_t l = 0; l < h->L; l++) p[j * h->ldv + h->V + l] = (float)U[l + h->L * j];
        }
        for (int64_t i = 0; i < h->V; i++) p[h->nW + i] = (float)a[i];
        for (int64_t l = 0; l < h->L; l++) p[h->nW + h->V + l] = (float)c[l];
        for (int64_t j = 0; j < h->H; j++) p[h->nW + h->ldv + j] = (float)b[j];
        QB_CUDA(cudaMemcpyAsync(h->P, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        h->master_sharded = false;
        if (h->Vel) QB_CUDA(cudaMemsetAsync(h->Vel, 0, h->nP * sizeof(float), h->stream));
        refresh_shadows(h);
        if (h->bias_rd) QB_CUDA(cudaMemcpyAsync(h->bias_rd, acat(h), (h->ldv + h->ldh) * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int32_t qb_get_params(qb_handle *h, double *W, double *a, double *b, double *U, double *c) {
    return guarded([&] {
        QB_H_KEEP(h); wait_pending_update(h);
        p2p_gather_master(h);
        std::vector<float> p((size_t)h->nP);
        QB_CUDA(cudaMemcpyAsync(p.data(), h->P, p.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t j = 0; j < h->H; j++) {
            if (W) for (int64_t i = 0; i < h->V; i++) W[i + h->V * j] = (double)p[j * h->ldv + i];
            if (U) for (int64_t l = 0; l < h->L; l++) U[l + h->L * j] = (double)p[j * h->ldv + h->V + l];
        }
        if (a) for (int64_t i = 0; i < h->V; i++) a[i] = (double)p[h->nW + i];
        if (c) for (int64_t l = 0; l < h->L; l++) c[l] = (double)p[h->nW + h->V + l];
        if (b) for (int64_t j = 0; j < h->H; j++) b[j] = (double)p[h->nW + h->ldv + j];
    });
}

// best_rbm = copy_rbm(rbm) / copy_rbm!(rbm, best_rbm) (src/training/train_cd.jl:60,99-102) kept in HBM
int32_t qb_snapshot_params(qb_handle *h) {
    return guarded([&] {
        QB_H(h);
        if (!h->P_best) h->P_best = dalloc<float>(h->nP);
        p2p_gather_master(h);
        QB_CUDA(cudaMemcpyAsync(h->P_best, h->P, h->nP * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
    });
}
// copy_rbm!(best_rbm, rbm) (src/training/train_cd.jl:109-111)
int32_t qb_restore_params(qb_handle *h) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(h->P_best != nullptr, QB_E_STATE, "qb_snapshot_params has not been called");
        QB_CUDA(cudaMemcpyAsync(h->P, h->P_best, h->nP * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        h->master_sharded = false;  // the snapshot holds every row
        refresh_shadows(h);
        if (h->bias_rd) QB_CUDA(cudaMemcpyAsync(h->bias_rd, acat(h), (h->ldv + h->ldh) * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        QB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int32_t qb_set_optimizer(qb_handle *h, double momentum, double weight_decay) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(momentum >= 0.0 && momentum < 1.0 && weight_decay >= 0.0, QB_E_ARG, "momentum in [0,1), weight_decay >= 0");
        h->momentum = momentum; h->weight_decay = weight_decay;
    });
}

int32_t qb_set_data(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t n_rows) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(X != nullptr && n_rows > 0, QB_E_ARG, "X is NULL or n_rows <= 0");
        QB_CHECK(h->L == 0 || Y != nullptr, QB_E_ARG, "labels required for classifier models");
        free_mat(h->data);
        cudaFree(h->dataT); h->dataT = nullptr; h->dataT_B = 0;
        cudaFree(h->datasum); h->datasum = nullptr; h->cur_datasum = nullptr;
        h->data = alloc_mat(h, n_rows, h->VL, false);
        h->N = n_rows;
        upload_rows(h, X, x_dtype, n_rows, h->V, h->data, 0, 0);
        if (h->L > 0) upload_rows(h, Y, y_dtype, n_rows, h->L, h->data, 0, h->V);
        h->data.exact = x_dtype == QB_DT_U8 && (h->L == 0 || y_dtype == QB_DT_U8);   // 0..255 are exact in bf16
        QB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int32_t qb_init_chains(qb_handle *h, int64_t n_chains, const double *v, const double *hid, const double *y) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(n_chains >= 1 && n_chains <= h->Bmax, QB_E_ARG, "n_chains must be in [1, max_batch]");
        h->n_chains = n_chains; h->chains_binary = false;
        h->cv.exact = false; h->ch.exact = false;
        if (v) {
            QB_CHECK(hid != nullptr && (h->L == 0 || y != nullptr), QB_E_ARG, "v, h (and y) must all be given");
            upload_rows(h, v, QB_DT_F64, n_chains, h->V, h->cv, 0, 0);
            if (h->L > 0) upload_rows(h, y, QB_DT_F64, n_chains, h->L, h->cv, 0, h->V);
            upload_rows(h, hid, QB_DT_F64, n_chains, h->H, h->ch, 0, 0);
        } else {
            // rand(V), rand(H)[, rand(L)] per chain (src/fantasy_data.jl:66-86) from the Philox stream
            const int64_t row0 = row0_of(h, n_chains);
            RngKey kv = make_key(h), kh = make_key(h);
            uniform_fill_kernel<<<grid1d(n_chains * h->VL), 256, 0, h->stream>>>(h->cv.f, h->cv.b, h->cv.ld, n_chains, h->VL, kv, row0);
            uniform_fill_kernel<<<grid1d(n_chains * h->H), 256, 0, h->stream>>>(h->ch.f, h->ch.b, h->ch.ld, n_chains, h->H, kh, row0);
            QB_CUDA(cudaGetLastError());
            h->launches += 2;
        }
        QB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int32_t qb_get_chains(qb_handle *h, double *v, double *hid, double *y) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(h->n_chains > 0, QB_E_STATE, "no chains");
        download_rows(h, h->cv, 0, 0, h->n_chains, h->V, v);
        download_rows(h, h->ch, 0, 0, h->n_chains, h->H, hid);
        if (h->L > 0) download_rows(h, h->cv, 0, h->V, h->n_chains, h->L, y);
    });
}

int32_t qb_get_chain_rows(qb_handle *h, int64_t row0, int64_t n_rows, double *v, double *hid, double *y) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(h->n_chains > 0, QB_E_STATE, "no chains");
        QB_CHECK(row0 >= 0 && n_rows >= 1 && row0 + n_rows <= h->n_chains, QB_E_ARG, "chain row range out of bounds");
        download_rows(h, h->cv, row0, 0, n_rows, h->V, v);
        download_rows(h, h->ch, row0, 0, n_rows, h->H, hid);
        if (h->L > 0) download_rows(h, h->cv, row0, h->V, n_rows, h->L, y);
    });
}

int32_t qb_seed(qb_handle *h, uint64_t seed) {
    return guarded([&] { QB_H(h); h->seed = seed; h->draw = 0; });
}

int32_t qb_inject_streams(qb_handle *h, const double *U_h, const double *U_v, const double *U_y, const double *Z_v, int64_t k,
                          int64_t B) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(U_h != nullptr && k >= 1 && B >= 1, QB_E_ARG, "U_h required, k >= 1, B >= 1");
        QB_CHECK(h->gaussian ? (Z_v != nullptr) : (U_v != nullptr), QB_E_ARG, "U_v (Bernoulli) or Z_v (Gaussian) required");
        QB_CHECK(h->L == 0 || U_y != nullptr, QB_E_ARG, "U_y required for classifier models");
        auto put = [&](float *&buf, size_t &cap, const std::vector<float> &src) {
            if (cap < src.size()) { cudaFree(buf); buf = nullptr; QB_CUDA(cudaMalloc(&buf, src.size() * sizeof(float))); cap = src.size(); }
            QB_CUDA(cudaMemcpyAsync(buf, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            QB_CUDA(cudaStreamSynchronize(h->stream));
        };
        std::vector<float> uh((size_t)k * B * h->H);
        for (size_t i = 0; i < uh.size(); i++) uh[i] = (float)U_h[i];
        put(h->injUh, h->injUh_cap, uh);
        if (h->gaussian && h->L == 0) {
            std::vector<float> z((size_t)k * B * h->V);
            for (size_t i = 0; i < z.size(); i++) z[i] = (float)Z_v[i];
            put(h->injZv, h->injZv_cap, z);
        } else {
            // one [k][B][VL] array: visible columns hold uniforms (Bernoulli) or normals (Gaussian), label columns uniforms
            std::vector<float> uv((size_t)k * B * h->VL);
            const double *vis = h->gaussian ? Z_v : U_v;
            for (int64_t s = 0; s < k; s++)
                for (int64_t r = 0; r < B; r++) {
                    float *dst = uv.data() + ((size_t)s * B + r) * h->VL;
                    for (int64_t i = 0; i < h->V; i++) dst[i] = (float)vis[((size_t)s * B + r) * h->V + i];
                    for (int64_t l = 0; l < h->L; l++) dst[h->V + l] = (float)U_y[((size_t)s * B + r) * h->L + l];
                }
            put(h->injUv, h->injUv_cap, uv);
        }
        h->inj_k = k; h->inj_B = B; h->inj_active = true;
    });
}

int32_t qb_pcd_step(qb_handle *h, int64_t batch_index, int64_t B, int64_t k, double lr, double llr) {
    return guarded([&] {
        QB_H_KEEP(h);
        check_step_args(h, B, k, /*allow_k0=*/!h->stale_chains);
        Mat xb; const bf16 *xbT; int64_t ldT;
        batch_view(h, batch_index, B, xb, xbT, ldT);
        pcd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
        if (!h->async_steps) { QB_CUDA(cudaStreamSynchronize(h->stream)); chain_check_error(h); }
    });
}

int32_t qb_cd_step(qb_handle *h, int64_t batch_index, int64_t B, int64_t k, double lr, double llr) {
    return guarded([&] {
        QB_H(h);
        check_step_args(h, B, k);
        Mat xb; const bf16 *xbT; int64_t ldT;
        batch_view(h, batch_index, B, xb, xbT, ldT);
        cd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
        if (!h->async_steps) { QB_CUDA(cudaStreamSynchronize(h->stream)); chain_check_error(h); }
    });
}

// contrastive_divergence! with the reference's own batch size 1 (src/training/train_cd.jl:2-21): one cooperative
// persistent kernel for the whole epoch (online_kernels.cuh) instead of 2k+3 launches per sample
static bool online_cd_epoch(qb_handle *h, int64_t k, double lr, double llr, double times[3]) {
    if (!h->online_kernel || h->cfg.world > 1 || h->momentum != 0.0 || h->weight_decay != 0.0) return false;
    const int64_t N = h->N;
    if (h->inj_active) QB_CHECK(h->inj_B == N && h->inj_k >= k, QB_E_ARG, "injected streams must cover every sample of the epoch");
    if (!h->W2) {
        h->W2 = dalloc<float>(h->VL * h->ldh);
        h->online_vec = dalloc<float>(2 * (3 * h->H + h->VL));
        h->online_bar = dalloc<unsigned int>(1);
        h->online_ns = dalloc<unsigned long long>(3);
    }
    {
        dim3 g((unsigned)ceil_div(h->VL, 32), (unsigned)ceil_div(h->H, 32));
        transpose_f32_kernel<<<g, 256, 0, h->stream>>>(Wf(h), h->ldv, (int)h->H, (int)h->VL, h->W2, h->ldh);
        QB_CUDA(cudaGetLastError());
    }
    QB_CUDA(cudaMemsetAsync(h->online_bar, 0, sizeof(unsigned int), h->stream));
    OnlineCdParams p{};
    p.W1 = Wf(h); p.ldv = h->ldv; p.W2 = h->W2; p.ldh = h->ldh; p.a = acat(h); p.b = bh(h);
    p.Xf = h->data.f; p.Xb = h->data.b; p.ldx = h->data.ld;
    p.V = (int)h->VL; p.Vx = (int)h->V; p.H = (int)h->H; p.n_samples = (int)N; p.k = (int)k; p.gaussian = h->gaussian ? 1 : 0;
    p.lr = (float)lr; p.llr = (float)llr;
    p.vec = h->online_vec; p.barrier = h->online_bar; p.phase_ns = h->online_ns;
    if (h->inj_active) { p.injUh = h->injUh; p.injUv = (h->gaussian && h->L == 0) ? h->injZv : h->injUv; }   // classifier streams: one [k][N][V + L] array
    p.key0.seed_lo = (uint32_t)h->seed; p.key0.seed_hi = (uint32_t)(h->seed >> 32);
    p.key0.draw_lo = (uint32_t)h->draw; p.key0.draw_hi = (uint32_t)(h->draw >> 32) & 0x7FFFFFFFu;
    // fewer, fatter CTAs make the grid barrier cheaper: one row per warp, up to 32 warps per CTA
    const int64_t rows = std::max(h->H, h->VL);
    const int wpb = h->online_wpb > 0 ? h->online_wpb : (rows >= 148 * 16 ? 32 : (rows >= 148 * 4 ? 16 : 8));
    int grid = (int)std::min<int64_t>(h->num_sms, std::max<int64_t>(1, ceil_div(rows, wpb)));
    const size_t smem_vec = (size_t)(2 * h->VL + 2 * h->H) * sizeof(float);
    QB_CHECK(smem_vec <= 200 * 1024, QB_E_UNSUPPORTED, "model too large for the online kernel's shared-memory vectors");
    // rows owned by a warp stay in shared memory for the whole epoch when they fit
    const int64_t warps = (int64_t)grid * wpb;
    const size_t smem_rows = (size_t)wpb * (size_t)(ceil_div(h->H, warps) * h->VL + ceil_div(h->VL, warps) * h->H) * sizeof(float);
    const bool cache = smem_vec + smem_rows <= 200 * 1024;
    const size_t smem = smem_vec + (cache ? smem_rows : 0);
    void *args[] = {(void *)&p};
    if (cache) {
        if (smem > 48 * 1024) QB_CUDA(cudaFuncSetAttribute(online_cd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QB_CUDA(cudaLaunchCooperativeKernel((void *)online_cd_kernel<true>, dim3(grid), dim3(32 * wpb), args, smem, h->stream));
    } else {
        if (smem > 48 * 1024) QB_CUDA(cudaFuncSetAttribute(online_cd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QB_CUDA(cudaLaunchCooperativeKernel((void *)online_cd_kernel<false>, dim3(grid), dim3(32 * wpb), args, smem, h->stream));
    }
    h->launches += 2;
    h->draw += 2ull * (uint64_t)k * (uint64_t)N;
    refresh_shadows(h);
    unsigned long long ns[3] = {0, 0, 0};
    QB_CUDA(cudaMemcpyAsync(ns, h->online_ns, sizeof(ns), cudaMemcpyDeviceToHost, h->stream));
    QB_CUDA(cudaStreamSynchronize(h->stream));
    if (times) for (int i = 0; i < 3; i++) times[i] = (double)ns[i] * 1e-9;
    h->inj_active = false;
    return true;
}

// fast_persistent_contrastive_divergence! (src/training/train_fast_pcd.jl:1-57): one epoch
static void fast_pcd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double flr, double llr, double fllr, double times[3]) {
    check_step_args(h, B, k);
    QB_CHECK(!h->inj_active, QB_E_STATE, "FastPCD epochs draw from the Philox stream (injected streams are per step)");
    QB_CHECK(h->cfg.world == 1, QB_E_UNSUPPORTED, "FastPCD updates the weights after every sample and does not shard");
    QB_CHECK(h->momentum == 0.0 && h->weight_decay == 0.0, QB_E_UNSUPPORTED, "FastPCD has no momentum / weight decay");
    QB_CHECK(h->n_chains == B, QB_E_STATE, "FastPCD pairs sample i with chain i: qb_init_chains(B) first");
    QB_CHECK(h->N > 0, QB_E_STATE, "qb_set_data has not been called");
    // N < B: _set_mini_batches drops the whole (short) batch, the epoch trains nothing (src/utils.jl:13-26)
    if (!h->W2) {
        h->W2 = dalloc<float>(h->VL * h->ldh);
        h->online_vec = dalloc<float>(2 * (3 * h->H + h->VL));
        h->online_bar = dalloc<unsigned int>(1);
        h->online_ns = dalloc<unsigned long long>(3);
    }
    if (!h->F) { h->F = dalloc<float>(h->nP); h->F2 = dalloc<float>(h->VL * h->ldh); h->P_eff = dalloc<float>(h->nP); }
    QB_CUDA(cudaMemsetAsync(h->F, 0, h->nP * sizeof(float), h->stream));          // W_fast = U_fast = a_fast = b_fast = c_fast = 0 every epoch (:12-14,:73-77)
    QB_CUDA(cudaMemsetAsync(h->F2, 0, h->VL * h->ldh * sizeof(float), h->stream));
    {
        dim3 g((unsigned)ceil_div(h->VL, 32), (unsigned)ceil_div(h->H, 32));
        transpose_f32_kernel<<<g, 256, 0, h->stream>>>(Wf(h), h->ldv, (int)h->H, (int)h->VL, h->W2, h->ldh);
        QB_CUDA(cudaGetLastError());
    }
    const int64_t rows = std::max(h->H, h->VL);
    const int wpb = h->online_wpb > 0 ? h->online_wpb : (rows >= 148 * 16 ? 32 : (rows >= 148 * 4 ? 16 : 8));
    const int grid = (int)std::min<int64_t>(h->num_sms, std::max<int64_t>(1, ceil_div(rows, wpb)));
    const int64_t warps = (int64_t)grid * wpb;
    const size_t smem_vec = (size_t)(2 * h->VL + 2 * h->H) * sizeof(float);
    const size_t smem_rows = 2 * (size_t)wpb * (size_t)(ceil_div(h->H, warps) * h->VL + ceil_div(h->VL, warps) * h->H) * sizeof(float);
    QB_CHECK(smem_vec <= 200 * 1024, QB_E_UNSUPPORTED, "model too large for the online kernel's shared-memory vectors");
    const bool cache = smem_vec + smem_rows <= 200 * 1024;
    const size_t smem = smem_vec + (cache ? smem_rows : 0);
    if (smem > 48 * 1024) {
        if (cache) QB_CUDA(cudaFuncSetAttribute(online_fastpcd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else QB_CUDA(cudaFuncSetAttribute(online_fastpcd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int64_t nb = h->N / B;
    float t_online = 0.f, t_chain = 0.f;
    float *P_true = h->P;
    for (int64_t mb = 0; mb < nb; mb++) {
        Mat xb = h->data.row_slice(mb * B, B);
        FastPcdParams p{};
        p.W1 = P_true; p.F1 = h->F; p.ldv = h->ldv; p.W2 = h->W2; p.F2 = h->F2; p.ldh = h->ldh;
        p.a = P_true + h->nW; p.b = P_true + h->nW + h->ldv; p.fa = h->F + h->nW; p.fb = h->F + h->nW + h->ldv;
        p.Xf = xb.f; p.Xb = xb.b; p.ldx = xb.ld;
        p.CVf = h->cv.f; p.CVb = h->cv.b; p.ldcv = h->cv.ld; p.CHf = h->ch.f; p.CHb = h->ch.b; p.ldch = h->ch.ld;
        p.V = (int)h->VL; p.Vx = (int)h->V; p.H = (int)h->H; p.B = (int)B;
        // classifier: the reference never advances batch_index, every sample pairs with chain 1 (train_fast_pcd.jl:77-119)
        p.chain_stride = (h->L == 0 || h->fast_pcd_pair_chains) ? 1 : 0;
        p.lr = (float)(lr / (double)B); p.flr = (float)(flr / (double)B);
        p.llr = (float)(llr / (double)B); p.fllr = (float)(fllr / (double)B);
        p.vec = h->online_vec; p.barrier = h->online_bar;
        QB_CUDA(cudaMemsetAsync(h->online_bar, 0, sizeof(unsigned int), h->stream));
        if (times) QB_CUDA(cudaEventRecord(h->ev[0], h->stream));
        void *args[] = {(void *)&p};
        if (cache) QB_CUDA(cudaLaunchCooperativeKernel((void *)online_fastpcd_kernel<true>, dim3(grid), dim3(32 * wpb), args, smem, h->stream));
        else QB_CUDA(cudaLaunchCooperativeKernel((void *)online_fastpcd_kernel<false>, dim3(grid), dim3(32 * wpb), args, smem, h->stream));
        if (times) QB_CUDA(cudaEventRecord(h->ev[1], h->stream));
        // chains advance with the effective parameters W + W_fast, a + a_fast, b + b_fast (src/fantasy_data.jl:31-47)
        add_vectors_kernel<<<grid1d(h->nP), 256, 0, h->stream>>>(P_true, h->F, h->P_eff, h->nP);
        QB_CUDA(cudaGetLastError());
        h->launches += 2;
        h->P = h->P_eff;
        float *saved_bias_rd = h->bias_rd; h->bias_rd = nullptr;
        h->gauss_threshold = h->gaussian;  // the fast gibbs_sample_visible thresholds even Gaussian "probabilities" (src/gibbs.jl:22-25)
        try {
            refresh_shadows(h);
            StepRng sr{h, B, false};
            advance_chains(h, B, k, sr, false);
        } catch (...) { h->P = P_true; h->bias_rd = saved_bias_rd; h->gauss_threshold = false; throw; }
        h->P = P_true; h->bias_rd = saved_bias_rd; h->gauss_threshold = false;
        if (times) {
            QB_CUDA(cudaEventRecord(h->ev[2], h->stream));
            QB_CUDA(cudaStreamSynchronize(h->stream));
            float a = 0, b = 0;
            QB_CUDA(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
            QB_CUDA(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
            t_online += a; t_chain += b;
        }
    }
    refresh_shadows(h);
    QB_CUDA(cudaStreamSynchronize(h->stream));
    if (times) { times[0] = 0.0; times[1] = t_chain * 1e-3; times[2] = t_online * 1e-3; }  // sample+update are fused in the online kernel
}

static void epoch_impl(qb_handle *h, bool pcd, int64_t B, int64_t k, double lr, double llr, double times[3]) {
    check_step_args(h, B, k);
    if (!pcd && B == 1 && h->N >= 1 && online_cd_epoch(h, k, lr, llr, times)) return;
    QB_CHECK(!h->inj_active, QB_E_STATE, "injected streams are per step (or per online CD epoch); use qb_pcd_step / qb_cd_step");
    QB_CHECK(h->N > 0, QB_E_STATE, "qb_set_data has not been called");
    const int64_t nb = h->N / B;  // _set_mini_batches: trailing remainder dropped -- all of it when N < B (src/utils.jl:13-26)
    double t[3] = {0, 0, 0};
    // Small models: the epoch is one launch of the persistent multi-mini-batch kernel.  When phase times are wanted the
    // first mini-batches run one launch sequence at a time with phase events (same arithmetic, same results) and their
    // split is scaled to the event-timed duration of the whole epoch.
    if (nb > 0 && h->epoch_kernel_mode != 0) {
        const int64_t probe = (times != nullptr && !(pcd && h->stale_chains)) ? std::min<int64_t>(4, nb) : 0;
        if (epoch_kernel_eligible(h, pcd, B, k)) {
            QB_CUDA(cudaEventRecord(h->ev[6], h->stream));
            cudaEvent_t pe[5];
            double f[3] = {0, 0, 0};
            for (int64_t m = 0; m < probe; m++) {
                Mat xb; const bf16 *xbT; int64_t ldT;
                batch_view(h, m, B, xb, xbT, ldT);
                while ((int64_t)h->epoch_ev.size() < 5 * (m + 1)) { cudaEvent_t e; QB_CUDA(cudaEventCreate(&e)); h->epoch_ev.push_back(e); }
                for (int i = 0; i < 5; i++) pe[i] = h->epoch_ev[5 * m + i];
                if (pcd) pcd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, pe);
                else cd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, pe);
            }
            if (epoch_kernel_run(h, pcd, probe, nb - probe, B, k, lr, llr) || nb == probe) {
                QB_CUDA(cudaEventRecord(h->ev[7], h->stream));
                QB_CUDA(cudaStreamSynchronize(h->stream));
                chain_check_error(h);
                if (times) {
                    for (int64_t m = 0; m < probe; m++) {
                        cudaEvent_t *e = h->epoch_ev.data() + 5 * m;
                        float a = 0, b = 0, c = 0, d = 0;
                        QB_CUDA(cudaEventElapsedTime(&a, e[0], e[1])); QB_CUDA(cudaEventElapsedTime(&b, e[1], e[2]));
                        QB_CUDA(cudaEventElapsedTime(&c, e[2], e[3]));
                        if (pcd) { QB_CUDA(cudaEventElapsedTime(&d, e[3], e[4])); f[1] += a + d; f[0] += b; f[2] += c; }
                        else { f[0] += a; f[1] += b; f[2] += c; }
                    }
                    float total = 0.f;
                    QB_CUDA(cudaEventElapsedTime(&total, h->ev[6], h->ev[7]));
                    const double sum = f[0] + f[1] + f[2];
                    for (int i = 0; i < 3; i++) times[i] = sum > 0 ? total * 1e-3 * f[i] / sum : 0.0;
                }
                return;
            }
            // not covered after all: the probe steps are done, continue with the remaining mini-batches below
            for (int64_t m = probe; m < nb; m++) {
                Mat xb; const bf16 *xbT; int64_t ldT;
                batch_view(h, m, B, xb, xbT, ldT);
                if (pcd) pcd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
                else cd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
            }
            QB_CUDA(cudaEventRecord(h->ev[7], h->stream));
            QB_CUDA(cudaStreamSynchronize(h->stream));
            chain_check_error(h);
            if (times) {
                float total = 0.f;
                QB_CUDA(cudaEventElapsedTime(&total, h->ev[6], h->ev[7]));
                times[0] = 0.0; times[1] = total * 1e-3; times[2] = 0.0;
            }
            return;
        }
    }
    // Phase times come from CUDA events recorded around the phases of every step and resolved AFTER the epoch (in chunks of
    // QB_EPOCH_EV_STEPS steps): no host synchronisation between mini-batches, the stream stays full.
    const bool timed = (times != nullptr) && !(pcd && h->stale_chains);  // the overlapped step has no phase boundaries
    constexpr int64_t QB_EPOCH_EV_STEPS = 1024;
    auto resolve = [&](int64_t n_steps) {
        QB_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t i = 0; i < n_steps; i++) {
            cudaEvent_t *e = h->epoch_ev.data() + 5 * i;
            float a = 0, b = 0, c = 0, d = 0;
            QB_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
            QB_CUDA(cudaEventElapsedTime(&b, e[1], e[2]));
            QB_CUDA(cudaEventElapsedTime(&c, e[2], e[3]));
            if (pcd) {
                QB_CUDA(cudaEventElapsedTime(&d, e[3], e[4]));
                t[1] += (a + d) * 1e-3; t[0] += b * 1e-3; t[2] += c * 1e-3;  // gibbs, sample, update
            } else {
                t[0] += a * 1e-3; t[1] += b * 1e-3; t[2] += c * 1e-3;
            }
        }
    };
    // long epochs of short steps: the phase split is sampled on every 16th mini-batch (event records between the launches of
    // a 30 us step cost a tenth of it) and scaled to the event-timed duration of the whole epoch
    const int64_t stride = nb > 64 ? 16 : 1;
    if (timed) QB_CUDA(cudaEventRecord(h->ev[6], h->stream));
    int64_t pending = 0;
    for (int64_t m = 0; m < nb; m++) {
        Mat xb; const bf16 *xbT; int64_t ldT;
        batch_view(h, m, B, xb, xbT, ldT);
        cudaEvent_t *ev = nullptr;
        if (timed && m % stride == 0) {
            while ((int64_t)h->epoch_ev.size() < 5 * (pending + 1)) { cudaEvent_t e; QB_CUDA(cudaEventCreate(&e)); h->epoch_ev.push_back(e); }
            ev = h->epoch_ev.data() + 5 * pending;
        }
        if (pcd) pcd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, ev);
        else cd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, ev);
        if (ev && ++pending == QB_EPOCH_EV_STEPS) { resolve(pending); pending = 0; }
    }
    if (timed) QB_CUDA(cudaEventRecord(h->ev[7], h->stream));
    if (timed && pending) resolve(pending);
    QB_CUDA(cudaStreamSynchronize(h->stream));
    chain_check_error(h);
    if (timed && nb > 0) {
        float total = 0.f;
        QB_CUDA(cudaEventElapsedTime(&total, h->ev[6], h->ev[7]));
        const double sum = t[0] + t[1] + t[2];
        if (stride > 1 && sum > 0) { const double f = total * 1e-3 / sum; t[0] *= f; t[1] *= f; t[2] *= f; }
    }
    if (times) { times[0] = t[0]; times[1] = t[1]; times[2] = t[2]; }
}

int32_t qb_train_steps(qb_handle *h, int32_t pcd, int64_t first_batch, int64_t n_steps, int64_t B, int64_t k, double lr, double llr) {
    return guarded([&] {
        if (pcd) { QB_H_KEEP(h); } else { QB_H(h); }
        check_step_args(h, B, k);
        QB_CHECK(!h->inj_active, QB_E_STATE, "injected streams are per step: use qb_pcd_step / qb_cd_step");
        QB_CHECK(h->N >= B, QB_E_STATE, "the resident dataset holds no full mini-batch");
        QB_CHECK(first_batch >= 0 && n_steps >= 0, QB_E_ARG, "first_batch and n_steps must be >= 0");
        QB_CHECK(!pcd || h->n_chains >= 1, QB_E_STATE, "qb_init_chains has not been called");
        const int64_t nb = h->N / B;
        if (n_steps > 0 && !epoch_kernel_run(h, pcd != 0, first_batch % nb, n_steps, B, k, lr, llr)) {
            for (int64_t i = 0; i < n_steps; i++) {
                Mat xb; const bf16 *xbT; int64_t ldT;
                batch_view(h, (first_batch + i) % nb, B, xb, xbT, ldT);
                if (pcd) pcd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
                else cd_step_impl(h, xb, xbT, ldT, B, k, lr, llr, nullptr);
            }
        }
        if (!h->async_steps) { QB_CUDA(cudaStreamSynchronize(h->stream)); chain_check_error(h); }
    });
}

int32_t qb_pcd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double llr, double times[3]) {
    return guarded([&] { QB_H_KEEP(h); epoch_impl(h, true, B, k, lr, llr, times); });
}
int32_t qb_fast_pcd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double fast_lr, double llr, double fast_llr, double times[3]) {
    return guarded([&] { QB_H(h); fast_pcd_epoch(h, B, k, lr, fast_lr, llr, fast_llr, times); });
}
int32_t qb_cd_epoch(qb_handle *h, int64_t B, int64_t k, double lr, double llr, double times[3]) {
    return guarded([&] { QB_H(h); epoch_impl(h, false, B, k, lr, llr, times); });
}

static void step_host_impl(qb_handle *h, bool pcd, const void *X, int xdt, const void *Y, int ydt, int64_t B, int64_t k,
                           double lr, double llr, double *recon) {
    check_step_args(h, B, k);
    h->cur_datasum = nullptr;
    stage_host_batch(h, X, xdt, Y, ydt, B);
    h->step_recon_wanted = recon != nullptr; h->step_recon_done = false;
    try {
        if (pcd) pcd_step_impl(h, h->hb, h->hb.bT, h->hb.ldT, B, k, lr, llr, nullptr);
        else cd_step_impl(h, h->hb, h->hb.bT, h->hb.ldT, B, k, lr, llr, nullptr);
    } catch (...) { h->step_recon_wanted = false; throw; }
    h->step_recon_wanted = false;
    if (recon) {
        if (!h->step_recon_done) {  // other models / modes: reconstruction of the batch after the step (two GEMMs)
            QB_CUDA(cudaMemsetAsync(h->dscal, 0, sizeof(double), h->stream));
            recon_sqerr(h, h->hb, B, nullptr, row0_of(h, B));
        }
        *recon = read_scalar(h, 0);
    } else {
        QB_CUDA(cudaStreamSynchronize(h->stream));
    }
}

int32_t qb_pcd_step_host(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t B, int64_t k,
                         double lr, double llr, double *recon_sq_err) {
    return guarded([&] { QB_H(h); step_host_impl(h, true, X, x_dtype, Y, y_dtype, B, k, lr, llr, recon_sq_err); });
}
int32_t qb_cd_step_host(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t B, int64_t k,
                        double lr, double llr, double *recon_sq_err) {
    return guarded([&] { QB_H(h); step_host_impl(h, false, X, x_dtype, Y, y_dtype, B, k, lr, llr, recon_sq_err); });
}

int32_t qb_set_next_host_batch(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(X != nullptr, QB_E_ARG, "X is NULL");
        QB_CHECK(h->L == 0 || Y != nullptr, QB_E_ARG, "labels required for classifier models");
        dt_size(x_dtype);
        h->next_X = X; h->next_Y = Y; h->next_xdt = x_dtype; h->next_ydt = y_dtype;
    });
}

int32_t qb_eval_mse(qb_handle *h, const void *X, int32_t x_dtype, int64_t n_rows, const double *Z, double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(out != nullptr, QB_E_ARG, "out is NULL");
        QB_CUDA(cudaMemsetAsync(h->dscal, 0, sizeof(double), h->stream));
        if (X == nullptr) {
            QB_CHECK(h->N > 0, QB_E_STATE, "no resident dataset");
            n_rows = h->N;
            recon_sqerr(h, h->data, h->N, Z, (int64_t)h->cfg.rank * h->N);
        } else {
            QB_CHECK(n_rows > 0, QB_E_ARG, "n_rows <= 0");
            for (int64_t r = 0; r < n_rows; r += h->Bmax) {
                const int64_t nr = std::min(h->Bmax, n_rows - r);
                upload_rows(h, (const char *)X + (size_t)r * h->V * dt_size(x_dtype), x_dtype, nr, h->V, h->hb, 0, 0);
                recon_sqerr(h, h->hb, nr, Z ? Z + (size_t)r * h->V : nullptr, r);
            }
        }
        *out = read_scalar(h, 0) / (double)n_rows;  // _evaluate(MeanSquaredError): sum over units, / dataset_size
    });
}

int32_t qb_eval_free_energy(qb_handle *h, const void *X, int32_t x_dtype, const void *Y, int32_t y_dtype, int64_t n_rows,
                            double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(out != nullptr, QB_E_ARG, "out is NULL");
        QB_CUDA(cudaMemsetAsync(h->dscal, 0, 2 * sizeof(double), h->stream));
        const bool resident = (X == nullptr);
        if (resident) { QB_CHECK(h->N > 0, QB_E_STATE, "no resident dataset"); n_rows = h->N; }
        QB_CHECK(n_rows > 0, QB_E_ARG, "n_rows <= 0");
        const bool with_y = h->L > 0 && (resident || Y != nullptr);
        const int64_t kdim = with_y ? h->VL : h->V;
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            Mat xs;
            if (resident) xs = h->data.row_slice(r, nr);
            else {
                upload_rows(h, (const char *)X + (size_t)r * h->V * dt_size(x_dtype), x_dtype, nr, h->V, h->hb, 0, 0);
                if (with_y) upload_rows(h, (const char *)Y + (size_t)r * h->L * dt_size(y_dtype), y_dtype, nr, h->L, h->hb, 0, h->V);
                xs = h->hb.row_slice(0, nr);
            }
            HiddenOut ho;
            if (h->bf) {
                QB_CUDA(cudaMemsetAsync(h->softplus_rows, 0, nr * sizeof(float), h->stream));
                ho.softplus_rows = h->softplus_rows;
            } else {
                ho.softplus_rows = h->pre_h;  // marker: fp32 path keeps the pre-activations in pre_h
            }
            op_hidden(h, xs, nr, kdim, ho, Rng(), 0);
            const int threads = 256, warps_per_block = threads / 32;
            dim3 g((unsigned)ceil_div(nr, warps_per_block));
            if (h->bf) free_energy_kernel<bf16><<<g, threads, 0, h->stream>>>(xs.b, xs.ld, nullptr, 0, h->softplus_rows, acat(h), (int)nr, (int)h->V, (int)kdim, (int)h->H, h->gaussian ? 1 : 0, h->dscal);
            else free_energy_kernel<float><<<g, threads, 0, h->stream>>>(xs.f, xs.ld, h->pre_h, h->ldh, nullptr, acat(h), (int)nr, (int)h->V, (int)kdim, (int)h->H, h->gaussian ? 1 : 0, h->dscal);
            QB_CUDA(cudaGetLastError());
            h->launches++;
        }
        *out = read_scalar(h, 0) / (double)n_rows;
    });
}

int32_t qb_reconstruct(qb_handle *h, const double *v, int64_t n_rows, const double *Z, double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(v && out && n_rows > 0, QB_E_ARG, "NULL argument or n_rows <= 0");
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            upload_rows(h, v + (size_t)r * h->V, QB_DT_F64, nr, h->V, h->hb, 0, 0);
            HiddenOut ho; ho.prob = &h->phd; ho.prob_rm = true;
            op_hidden(h, h->hb, nr, h->V, ho, Rng(), 0);
            Rng rng;
            if (h->gaussian) {
                if (Z) {
                    Mat zf; zf.f = h->znoise; zf.ld = h->V; zf.rows = nr; zf.units = h->V;
                    upload_rows(h, Z + (size_t)r * h->V, QB_DT_F64, nr, h->V, zf, 0, 0);
                    rng = injected_rng(h->znoise, h->V);
                } else {
                    rng = philox_rng(h);
                }
            }
            Mat of; of.f = h->out_f32; of.ld = h->ldv; of.rows = nr; of.units = h->VL;
            VisibleOut vo; vo.sample = false;
            if (h->bf) { vo.state = nullptr; vo.prob_f32 = h->out_f32; vo.ld_pf = h->ldv; vo.f32_from_state = h->gaussian; }
            else vo.state = &of;
            op_visible(h, h->phd, nr, vo, rng, r);
            download_rows(h, of, 0, 0, nr, h->V, out + (size_t)r * h->V);
        }
    });
}

// gibbs_sample_hidden (src/gibbs.jl:3-6, 29-32) for n_rows rows
int32_t qb_sample_hidden(qb_handle *h, const double *v, const double *y, int64_t n_rows, const double *U, double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(v && out && n_rows > 0, QB_E_ARG, "NULL argument or n_rows <= 0");
        QB_CHECK(y == nullptr || h->L > 0, QB_E_ARG, "y given for a model without label units");
        if (U && !h->rand_rows) h->rand_rows = dalloc<float>(h->Bmax * std::max(h->VL, h->H));
        const Rng philox = U ? Rng() : philox_rng(h);  // one draw id for the whole call, rows keyed by their index
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            upload_rows(h, v + (size_t)r * h->V, QB_DT_F64, nr, h->V, h->hb, 0, 0);
            if (y) upload_rows(h, y + (size_t)r * h->L, QB_DT_F64, nr, h->L, h->hb, 0, h->V);
            Rng rng = philox;
            if (U) {
                Mat uf; uf.f = h->rand_rows; uf.ld = h->H; uf.rows = nr; uf.units = h->H;
                upload_rows(h, U + (size_t)r * h->H, QB_DT_F64, nr, h->H, uf, 0, 0);
                rng = injected_rng(h->rand_rows, h->H);
            }
            HiddenOut ho; ho.state = &h->hs; ho.state_rm = true;
            op_hidden(h, h->hb, nr, y ? h->VL : h->V, ho, rng, r);
            download_rows(h, h->hs, 0, 0, nr, h->H, out + (size_t)r * h->H);
        }
    });
}

// gibbs_sample_visible (src/gibbs.jl:13-20) and gibbs_sample_label (src/gibbs.jl:46-49) for n_rows rows
int32_t qb_sample_visible(qb_handle *h, const double *hid, int64_t n_rows, const double *R, const double *Ry, double *out_v,
                          double *out_y) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(hid && out_v && n_rows > 0, QB_E_ARG, "NULL argument or n_rows <= 0");
        QB_CHECK(h->L == 0 || R == nullptr || Ry != nullptr, QB_E_ARG, "label uniforms Ry are required with R on a classifier");
        QB_CHECK(out_y == nullptr || h->L > 0, QB_E_ARG, "out_y given for a model without label units");
        if (R && !h->rand_rows) h->rand_rows = dalloc<float>(h->Bmax * std::max(h->VL, h->H));
        const Rng philox = R ? Rng() : philox_rng(h);
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            upload_rows(h, hid + (size_t)r * h->H, QB_DT_F64, nr, h->H, h->hs, 0, 0);
            Rng rng = philox;
            if (R) {  // one [rows][V + L] array: visible uniforms (normals for Gaussian visibles), then label uniforms
                Mat uf; uf.f = h->rand_rows; uf.ld = h->VL; uf.rows = nr; uf.units = h->VL;
                upload_rows(h, R + (size_t)r * h->V, QB_DT_F64, nr, h->V, uf, 0, 0);
                if (h->L > 0) upload_rows(h, Ry + (size_t)r * h->L, QB_DT_F64, nr, h->L, uf, 0, h->V);
                rng = injected_rng(h->rand_rows, h->VL);
            }
            VisibleOut vo; vo.state = &h->vm; vo.sample = true; vo.state_rm = true;
            op_visible(h, h->hs, nr, vo, rng, r);
            download_rows(h, h->vm, 0, 0, nr, h->V, out_v + (size_t)r * h->V);
            if (out_y) download_rows(h, h->vm, 0, h->V, nr, h->L, out_y + (size_t)r * h->L);
        }
    });
}

int32_t qb_conditional_prob_h(qb_handle *h, const double *v, const double *y, int64_t n_rows, double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(v && out && n_rows > 0, QB_E_ARG, "NULL argument or n_rows <= 0");
        QB_CHECK(y == nullptr || h->L > 0, QB_E_ARG, "y given for a model without label units");
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            upload_rows(h, v + (size_t)r * h->V, QB_DT_F64, nr, h->V, h->hb, 0, 0);
            if (y) upload_rows(h, y + (size_t)r * h->L, QB_DT_F64, nr, h->L, h->hb, 0, h->V);
            Mat of; of.f = h->out_f32; of.ld = h->ldh; of.rows = nr; of.units = h->H;
            HiddenOut ho;
            if (h->bf) { ho.prob_f32 = h->out_f32; ho.ld_pf = h->ldh; }
            else ho.prob = &of;
            op_hidden(h, h->hb, nr, y ? h->VL : h->V, ho, Rng(), 0);
            download_rows(h, of, 0, 0, nr, h->H, out + (size_t)r * h->H);
        }
    });
}

// conditional_prob_y_given_v (src/rbm.jl:191-194): softmax(c + U * sigma(b + W'v)) for n_rows rows
int32_t qb_conditional_prob_y_given_v(qb_handle *h, const void *X, int32_t x_dtype, int64_t n_rows, double *out) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(h->L > 0, QB_E_ARG, "model has no label units");
        QB_CHECK(out != nullptr, QB_E_ARG, "out is NULL");
        const bool resident = (X == nullptr);
        if (resident) { QB_CHECK(h->N > 0, QB_E_STATE, "no resident dataset"); n_rows = h->N; }
        QB_CHECK(n_rows > 0, QB_E_ARG, "n_rows <= 0");
        for (int64_t r = 0; r < n_rows; r += h->Bmax) {
            const int64_t nr = std::min(h->Bmax, n_rows - r);
            Mat xs;
            if (resident) xs = h->data.row_slice(r, nr);
            else {
                upload_rows(h, (const char *)X + (size_t)r * h->V * dt_size(x_dtype), x_dtype, nr, h->V, h->hb, 0, 0);
                xs = h->hb.row_slice(0, nr);
            }
            HiddenOut ho; ho.prob = &h->phd; ho.prob_rm = true;
            op_hidden(h, xs, nr, h->V, ho, Rng(), 0);  // 2-arg conditional_prob_h: labels are not inputs here
            // label pre-activations c + U p_h: a GEMM over the L label rows of the visible-side weights only
            if (h->bf) {
                sm100::Operand A{h->WbT + h->V * h->ldh, h->L, h->H, h->ldh}, B{h->phd.b, nr, h->H, h->phd.ld};
                sm100::EpiParams ep{};
                ep.mode = sm100::EPI_SAMPLE; ep.M = (int)h->L; ep.N = (int)nr; ep.bias = bias_a(h) + h->V;
                ep.act = sm100::ACT_SIGMOID; ep.sample = 0; ep.split = 0; ep.lab_pre = h->lab_pre; ep.ld_lab = h->L; ep.pscale = 1.f;
                GemmScope gs(h, 2.0 * (double)h->L * (double)nr * (double)h->H);
                sm100::launch(h->stream, A, B, nullptr, nullptr, ep, h->num_sms, h->force_bn, 0);
            } else {
                simt_gemm(h, h->phd.f, h->phd.ld, 1, Wf(h) + h->V, h->ldv, 1, h->lab_pre, h->L, nr, h->L, h->H, 1.f, 0.f, acat(h) + h->V);
            }
            ensure_staging(h, (size_t)nr * h->L * sizeof(double));
            label_softmax_kernel<<<grid1d(nr, 128), 128, 0, h->stream>>>(h->lab_pre, h->L, (int)nr, (int)h->L, (double *)h->staging);
            QB_CUDA(cudaGetLastError());
            h->launches++;
            QB_CUDA(cudaMemcpyAsync(out + (size_t)r * h->L, h->staging, (size_t)nr * h->L * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            QB_CUDA(cudaStreamSynchronize(h->stream));
        }
    });
}

// ----------------------------------------------------------------------------- train!'s epoch loop
// _diverged (src/evaluation.jl:132-168): Accuracy must not decrease / fall below its maximum, every other metric must not
// increase / exceed its minimum.  NaN compares false, as in Julia.
static bool metric_diverged(const std::vector<double> &hist, int metric) {
    if (hist.size() <= 1) return false;
    const double last = hist.back(), prev = hist[hist.size() - 2];
    if (metric == QB_METRIC_ACCURACY) {
        if (last < prev) return true;
        double mx = hist[0];
        for (double v : hist) if (v > mx) mx = v;   // maximum(): NaN-free by construction (accuracy is a count / N)
        return last < mx;
    }
    if (last > prev) return true;
    double mn = hist[0];
    bool any_nan = false;
    for (double v : hist) { if (v != v) any_nan = true; else if (mn != mn || v < mn) mn = v; }
    if (any_nan) return false;   // minimum() of a vector holding NaN is NaN in Julia, and x > NaN is false
    return last > mn;
}

int32_t qb_train(qb_handle *h, const qb_train_config *c, double *history, double times[3], int32_t *epochs_run) {
    // plain C++ over the entry points above (each of them is guarded): status codes are passed through, nothing throws here
    if (!h || !c || !history) { g_last_error = "qb_train: NULL argument"; return QB_E_ARG; }
    auto fail = [&](int32_t code, const char *msg) { g_last_error = msg; return code; };
    if (c->n_epochs < 0 || c->batch_size < 1 || c->gibbs_steps < 1 || !c->learning_rate) return fail(QB_E_ARG, "qb_train: n_epochs >= 0, batch_size >= 1, gibbs_steps >= 1 and learning_rate are required");
    if (c->method < QB_METHOD_CD || c->method > QB_METHOD_FAST_PCD) return fail(QB_E_ARG, "qb_train: unknown method");
    if (c->n_metrics < 1 || c->n_metrics > 8 || !c->metrics) return fail(QB_E_ARG, "qb_train: 1..8 metrics required");
    const bool clf = h->L > 0;
    if (clf && !c->label_learning_rate) return fail(QB_E_ARG, "qb_train: label_learning_rate is required for classifier models");
    int stop_col = -1;
    for (int i = 0; i < c->n_metrics; i++) {
        const int m = c->metrics[i];
        if (m < QB_METRIC_MSE || m > QB_METRIC_FREE_ENERGY) return fail(QB_E_ARG, "qb_train: unknown metric");
        if ((m == QB_METRIC_ACCURACY || m == QB_METRIC_CROSS_ENTROPY) && (!clf || !c->Y_eval))
            return fail(QB_E_ARG, "qb_train: Accuracy / CrossEntropy need a classifier model and Y_eval");
        if (m == c->stopping_metric && stop_col < 0) stop_col = i;
    }
    if (stop_col < 0) return fail(QB_E_ARG, "qb_train: stopping_metric is not among the metrics (the reference raises KeyError)");
    const bool test = c->X_eval != nullptr;
    if (test && c->n_eval < 1) return fail(QB_E_ARG, "qb_train: n_eval must be >= 1 with X_eval");
    const int64_t n_eval = test ? c->n_eval : h->N;
    if (n_eval < 1) return fail(QB_E_STATE, "qb_train: no resident dataset (qb_set_data)");
    const int L = (int)h->L, nm = c->n_metrics;
    std::vector<double> prob;   // p(y | v) of the evaluated rows, shared by Accuracy and CrossEntropy

    auto label_at = [&](int64_t row, int l) -> double {
        const char *p = (const char *)c->Y_eval + ((size_t)row * L + l) * dt_size(c->y_dtype);
        switch (c->y_dtype) {
            case QB_DT_F64: return *(const double *)p;
            case QB_DT_I64: return (double)*(const int64_t *)p;
            case QB_DT_F32: return (double)*(const float *)p;
            default: return (double)*(const uint8_t *)p;
        }
    };
    auto evaluate = [&](double *row) -> int32_t {
        bool have_prob = false;
        for (int i = 0; i < nm; i++) {
            const int m = c->metrics[i];
            int32_t rc = QB_OK;
            if (m == QB_METRIC_MSE) rc = qb_eval_mse(h, test ? c->X_eval : nullptr, c->x_dtype, test ? n_eval : 0, nullptr, &row[i]);
            else if (m == QB_METRIC_FREE_ENERGY) rc = qb_eval_free_energy(h, test ? c->X_eval : nullptr, c->x_dtype, nullptr, 0, test ? n_eval : 0, &row[i]);
            else {
                if (!have_prob) {
                    prob.resize((size_t)n_eval * L);
                    rc = qb_conditional_prob_y_given_v(h, test ? c->X_eval : nullptr, c->x_dtype, test ? n_eval : 0, prob.data());
                    have_prob = rc == QB_OK;
                }
                if (rc == QB_OK) {
                    // evaluate(rbm::RBMClassifiers, ...) (src/evaluation.jl:28-50): rounded_y_pred = 1 where p equals the row maximum
                    double acc = 0.0, ce = 0.0;
                    for (int64_t r = 0; r < n_eval; r++) {
                        const double *p = prob.data() + (size_t)r * L;
                        double mx = p[0];
                        for (int l = 1; l < L; l++) if (p[l] > mx) mx = p[l];
                        bool all_eq = true;
                        double s = 0.0;
                        for (int l = 0; l < L; l++) {
                            const double pred = (p[l] == mx) ? 1.0 : 0.0, y = label_at(r, l);
                            if (std::llround(y) != (long long)pred) all_eq = false;
                            s += y * std::log(pred) + (1.0 - y) * std::log(1.0 - pred);   // literally: 0 * -Inf = NaN
                        }
                        acc += (all_eq ? 1.0 : 0.0) / (double)n_eval;
                        ce += s / (double)n_eval;
                    }
                    row[i] = (m == QB_METRIC_ACCURACY) ? acc : ce;
                }
            }
            if (rc != QB_OK) return rc;
        }
        return QB_OK;
    };

    int32_t rc = qb_snapshot_params(h);                 // best_rbm = copy_rbm(rbm)
    if (rc != QB_OK) return rc;
    if ((rc = evaluate(history)) != QB_OK) return rc;   // initial_evaluation
    if (c->method != QB_METHOD_CD && (rc = qb_init_chains(h, c->batch_size, nullptr, nullptr, nullptr)) != QB_OK) return rc;
    std::vector<double> stop_hist;
    int patience = c->patience, ran = c->n_epochs;
    double tot[3] = {0, 0, 0};
    for (int e = 1; e <= c->n_epochs; e++) {
        double t[3] = {0, 0, 0};
        const double lr = c->learning_rate[e - 1], llr = clf ? c->label_learning_rate[e - 1] : 0.0;
        if (c->method == QB_METHOD_FAST_PCD) rc = qb_fast_pcd_epoch(h, c->batch_size, c->gibbs_steps, lr, c->fast_learning_rate, llr, c->fast_label_learning_rate, t);
        else if (c->method == QB_METHOD_PCD) rc = qb_pcd_epoch(h, c->batch_size, c->gibbs_steps, lr, llr, t);
        else rc = qb_cd_epoch(h, c->batch_size, c->gibbs_steps, lr, llr, t);
        if (rc != QB_OK) return rc;
        for (int i = 0; i < 3; i++) tot[i] += t[i];
        double *row = history + (size_t)e * nm;
        if ((rc = evaluate(row)) != QB_OK) return rc;
        stop_hist.push_back(row[stop_col]);
        int flags = 0;
        bool stop = false;
        if (metric_diverged(stop_hist, c->stopping_metric)) {
            flags |= QB_EPOCH_DIVERGED;
            if (c->early_stopping) {
                if (patience == 0) { flags |= QB_EPOCH_EARLY_STOP; stop = true; }
                else patience -= 1;
            }
        } else {
            patience = c->patience;
            if (c->store_best_rbm) {
                if ((rc = qb_snapshot_params(h)) != QB_OK) return rc;   // copy_rbm!(rbm, best_rbm)
                flags |= QB_EPOCH_NEW_BEST;
            }
        }
        if (stop) ran = e;
        // the reference breaks out BEFORE logging the epoch it stops at (train_cd.jl:91-96): the callback still sees it, flagged
        if (c->on_epoch) c->on_epoch(c->user, e, row, t, flags);
        if (stop) break;
    }
    if (c->store_best_rbm && (rc = qb_restore_params(h)) != QB_OK) return rc;   // copy_rbm!(best_rbm, rbm)
    if (times) for (int i = 0; i < 3; i++) times[i] = tot[i];
    if (epochs_run) *epochs_run = ran;
    return QB_OK;
}

int32_t qb_comm_unique_id(void *id128) {
    return guarded([&] {
        QB_CHECK(id128 != nullptr, QB_E_ARG, "id buffer is NULL");
        ncclUniqueId id;
        QB_NCCL(nccl().GetUniqueId(&id));
        memcpy(id128, &id, sizeof(id));
    });
}

int32_t qb_comm_init(qb_handle *h, const void *id128) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(id128 != nullptr, QB_E_ARG, "id buffer is NULL");
        QB_CHECK(h->cfg.world > 1, QB_E_ARG, "world == 1 needs no communicator");
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        QB_NCCL(nccl().CommInitRank(&h->comm, h->cfg.world, id, h->cfg.rank));
        // default exchange in bf16 mode: reduce-scatter + shard update + bf16 all-gather (measured 6 % faster than allreduce +
        // full update at 8 GPUs, equal at 2); option "sharded_update" = 0 goes back to the allreduce
        h->sharded_update = h->bf && (h->H % h->cfg.world) == 0;

    });
}

// ---- NVLink peer-memory path: 5 IPC handles per rank (G, Wb, WbT, flags, P), 64 bytes each
int32_t qb_p2p_export(qb_handle *h, void *handles320) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(handles320 != nullptr, QB_E_ARG, "handle buffer is NULL");
        QB_CHECK(h->bf && h->cfg.world > 1 && h->cfg.world <= QB_MAX_PEERS, QB_E_UNSUPPORTED, "peer-memory exchange needs bf16 precision and 2..8 ranks");
        QB_CHECK(!h->stale_chains, QB_E_UNSUPPORTED, "peer-memory exchange and stale_chains are mutually exclusive");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        if (!h->p2p_flags) { h->p2p_flags = dalloc<unsigned int>(QB_MAX_PEERS); h->p2p_grid_bar = dalloc<unsigned int>(1); h->p2p_ns = dalloc<unsigned long long>(2); }
        cudaIpcMemHandle_t hd[5];
        QB_CUDA(cudaIpcGetMemHandle(&hd[0], h->G));
        QB_CUDA(cudaIpcGetMemHandle(&hd[1], h->Wb));
        QB_CUDA(cudaIpcGetMemHandle(&hd[2], h->WbT));
        QB_CUDA(cudaIpcGetMemHandle(&hd[3], h->p2p_flags));
        QB_CUDA(cudaIpcGetMemHandle(&hd[4], h->P));
        memcpy(handles320, hd, sizeof(hd));
    });
}
int32_t qb_p2p_import(qb_handle *h, const void *all_handles) {
    return guarded([&] {
        QB_H(h);
        QB_CHECK(all_handles != nullptr && h->p2p_flags != nullptr, QB_E_STATE, "call qb_p2p_export on every rank first");
        const cudaIpcMemHandle_t *hd = (const cudaIpcMemHandle_t *)all_handles;
        for (int r = 0; r < h->cfg.world; r++) {
            if (r == h->cfg.rank) {
                h->peerG[r] = h->G; h->peerWb[r] = h->Wb; h->peerWbT[r] = h->WbT; h->peerFlags[r] = h->p2p_flags; h->peerP[r] = h->P;
                continue;
            }
            void *ptr[5];
            for (int i = 0; i < 5; i++) QB_CUDA(cudaIpcOpenMemHandle(&ptr[i], hd[r * 5 + i], cudaIpcMemLazyEnablePeerAccess));
            h->peerG[r] = (float *)ptr[0]; h->peerWb[r] = (bf16 *)ptr[1]; h->peerWbT[r] = (bf16 *)ptr[2];
            h->peerFlags[r] = (unsigned int *)ptr[3]; h->peerP[r] = (float *)ptr[4];
        }
        h->shard_rows = round_up(ceil_div(h->H, h->cfg.world), 64);
        h->p2p = true;
    });
}

int32_t qb_get_counters(qb_handle *h, int64_t *kernel_launches, int64_t *gemm_launches, double *gemm_flops, int64_t *timed_gemms,
                        double *timed_gemm_ms, double *timed_gemm_flops) {
    return guarded([&] {
        QB_H_KEEP(h);
        if (kernel_launches) *kernel_launches = h->launches;
        if (gemm_launches) *gemm_launches = h->gemm_launches;
        if (gemm_flops) *gemm_flops = h->gemm_flops;
        double ms = 0.0;
        if (h->gemm_ev_used) {
            QB_CUDA(cudaStreamSynchronize(h->stream));
            for (size_t i = 0; i + 1 < h->gemm_ev_used; i += 2) {
                float t = 0.f;
                QB_CUDA(cudaEventElapsedTime(&t, h->gemm_ev[i], h->gemm_ev[i + 1]));
                ms += t;
            }
        }
        if (timed_gemms) *timed_gemms = h->gemm_timed;
        if (timed_gemm_ms) *timed_gemm_ms = ms;
        if (timed_gemm_flops) *timed_gemm_flops = h->gemm_timed_flops;
    });
}
int32_t qb_get_gemm_histogram(qb_handle *h, int64_t *counts, int32_t n_slots) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(counts != nullptr && n_slots >= 1, QB_E_ARG, "counts is NULL or n_slots < 1");
        for (int i = 0; i < n_slots; i++) counts[i] = i < QB_GEMM_HIST_SLOTS ? h->gemm_hist[i] : 0;
    });
}
int32_t qb_get_issued_flops(qb_handle *h, double *issued_flops) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(issued_flops != nullptr, QB_E_ARG, "issued_flops is NULL");
        *issued_flops = h->gemm_flops + h->gemm_issued_flops;
    });
}
int32_t qb_reset_counters(qb_handle *h) {
    return guarded([&] {
        QB_H_KEEP(h);
        h->launches = 0; h->gemm_launches = 0; h->gemm_flops = 0.0; h->gemm_issued_flops = 0.0;
        for (auto &c : h->gemm_hist) c = 0;
        h->gemm_ev_used = 0; h->gemm_timed = 0; h->gemm_timed_flops = 0.0;
        h->upd_ev_used = 0; h->upd_timed = 0;
    });
}
int32_t qb_set_option(qb_handle *h, const char *key, double value) {
    return guarded([&] {
        QB_H_KEEP(h);
        QB_CHECK(key != nullptr, QB_E_ARG, "key is NULL");
        if (!strcmp(key, "force_bn")) {
            int bn = (int)value;
            QB_CHECK(bn == 0 || bn == 32 || bn == 64 || bn == 128 || bn == 256, QB_E_ARG, "force_bn must be 0/32/64/128/256");
            h->force_bn = bn;
        } else if (!strcmp(key, "split_exchange")) {
            QB_CHECK(value == 0.0 || value == 1.0 || value == 2.0, QB_E_ARG, "split_exchange must be 0, 1 or 2");
            h->split_exchange_mode = (int)value;
        } else if (!strcmp(key, "chain_quad")) {
            QB_CHECK(value == 0.0 || value == 1.0 || value == 2.0, QB_E_ARG, "chain_quad must be 0, 1 or 2");
            h->chain_quad_mode = (int)value;
        } else if (!strcmp(key, "nvls_exchange")) {
            QB_CHECK(value == 1.0, QB_E_ARG, "nvls_exchange can only be switched on (1)");
            nvls_enable(h);
        } else if (!strcmp(key, "p2p_blocks_per_sm")) {
            QB_CHECK(value >= 1 && value <= 4, QB_E_ARG, "p2p_blocks_per_sm must be 1..4");
            h->p2p_blocks_per_sm = (int)value;
        } else if (!strcmp(key, "online_wpb")) {
            QB_CHECK(value == 0 || value == 4 || value == 8 || value == 16 || value == 32, QB_E_ARG, "online_wpb must be 0/4/8/16/32");
            h->online_wpb = (int)value;
        } else if (!strcmp(key, "online_kernel")) {
            h->online_kernel = value != 0.0;
        } else if (!strcmp(key, "sharded_update")) {
            QB_CHECK(value == 0.0 || (h->bf && h->cfg.world > 1 && (h->H % h->cfg.world) == 0), QB_E_UNSUPPORTED,
                     "sharded_update needs bf16 precision, world > 1 and n_hidden % world == 0");
            QB_CHECK(!h->master_sharded || value != 0.0, QB_E_STATE, "gather the master (qb_get_params) before switching sharded_update off");
            h->sharded_update = value != 0.0;
        } else if (!strcmp(key, "state_exchange")) {
            QB_CHECK(value == 0.0 || (h->bf && h->comm && h->sharded_update), QB_E_UNSUPPORTED,
                     "state_exchange needs bf16 precision, a communicator and the sharded update");
            QB_CUDA(cudaStreamSynchronize(h->stream));
            h->state_exchange = value != 0.0;
            if (h->state_exchange) ensure_overlap_resources(h);  // collective (communicator split)
            h->gemm_sms = ((h->stale_chains || h->state_exchange) && h->comm_overlap) ? std::max(2, (h->num_sms - QB_OVERLAP_CTAS) & ~1) : h->num_sms;
        } else if (!strcmp(key, "fused_chain")) {
            QB_CHECK(value == 0.0 || value == 1.0 || value == 2.0, QB_E_ARG, "fused_chain must be 0, 1 or 2");
            h->fused_chain_mode = (int)value;
        } else if (!strcmp(key, "epoch_kernel")) {
            QB_CHECK(value == 0.0 || value == 1.0 || value == 2.0, QB_E_ARG, "epoch_kernel must be 0, 1 or 2");
            h->epoch_kernel_mode = (int)value;
        } else if (!strcmp(key, "chain_pairs")) {
            QB_CHECK(value >= -1.0 && value <= 1024.0, QB_E_ARG, "chain_pairs must be 0 (all) or a count");
            h->chain_pairs = (int)value;
        } else if (!strcmp(key, "chain_bn")) {
            const int bn = (int)value;
            QB_CHECK(bn == 0 || bn == 64 || bn == 128 || bn == 256, QB_E_ARG, "chain_bn must be 0/64/128/256");
            h->chain_bn = bn;
        } else if (!strcmp(key, "fast_pcd_pair_chains")) {
            h->fast_pcd_pair_chains = value != 0.0;
        } else if (!strcmp(key, "stale_chains")) {
            set_stale_mode(h, value != 0.0);
        } else if (!strcmp(key, "use_2cta")) {
            QB_CHECK(value == 0.0 || value == 1.0 || value == 2.0, QB_E_ARG, "use_2cta must be 0, 1 or 2");
            h->use_2cta = (int)value;
        } else if (!strcmp(key, "async_steps")) {
            h->async_steps = value != 0.0;
        } else if (!strcmp(key, "time_gemms")) {
            h->time_gemms = value != 0.0;
        } else {
            throw Error(QB_E_ARG, std::string("unknown option ") + key);
        }
    });
}
// p2p mode: accumulated ns inside the fused exchange kernel: [waiting for the peers' gradients, reduce + update + push]
int32_t qb_get_p2p_timing(qb_handle *h, double *wait_ms, double *work_ms) {
    return guarded([&] {
        QB_H_KEEP(h);
        unsigned long long ns[2] = {0, 0};
        if (h->p2p_ns) {
            QB_CUDA(cudaStreamSynchronize(h->stream));
            QB_CUDA(cudaMemcpy(ns, h->p2p_ns, sizeof(ns), cudaMemcpyDeviceToHost));
            QB_CUDA(cudaMemset(h->p2p_ns, 0, sizeof(ns)));
        }
        if (wait_ms) *wait_ms = ns[0] * 1e-6;
        if (work_ms) *work_ms = ns[1] * 1e-6;
    });
}
int32_t qb_get_update_timing(qb_handle *h, int64_t *timed_updates, double *timed_update_ms) {
    return guarded([&] {
        QB_H_KEEP(h);
        double ms = 0.0;
        if (h->upd_ev_used) {
            QB_CUDA(cudaStreamSynchronize(h->stream));
            for (size_t i = 0; i + 1 < h->upd_ev_used; i += 2) {
                float t = 0.f;
                QB_CUDA(cudaEventElapsedTime(&t, h->upd_ev[i], h->upd_ev[i + 1]));
                ms += t;
            }
        }
        if (timed_updates) *timed_updates = h->upd_timed;
        if (timed_update_ms) *timed_update_ms = ms;
    });
}
int32_t qb_sync(qb_handle *h) {
    return guarded([&] { QB_H_KEEP(h); wait_pending_update(h); QB_CUDA(cudaStreamSynchronize(h->stream)); chain_check_error(h); });
}
int32_t qb_timer_start(qb_handle *h) {
    return guarded([&] { QB_H_KEEP(h); QB_CUDA(cudaEventRecord(h->timer_ev[0], h->stream)); });
}
int32_t qb_timer_stop(qb_handle *h, double *elapsed_ms) {
    return guarded([&] {
        QB_H_KEEP(h); wait_pending_update(h);
        QB_CHECK(elapsed_ms != nullptr, QB_E_ARG, "elapsed_ms is NULL");
        QB_CUDA(cudaEventRecord(h->timer_ev[1], h->stream));  // after QB_H: includes the last asynchronous update
        QB_CUDA(cudaEventSynchronize(h->timer_ev[1]));
        float t = 0.f;
        QB_CUDA(cudaEventElapsedTime(&t, h->timer_ev[0], h->timer_ev[1]));
        *elapsed_ms = t;
        chain_check_error(h);
    });
}

}  // extern "C"
