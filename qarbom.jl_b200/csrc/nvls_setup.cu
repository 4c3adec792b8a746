// NVSwitch multicast ("NVLS") addresses for the fused gradient exchange (p2p_kernels.cuh, NVLS variant), obtained through
// NCCL's symmetric-memory windows -- the supported way for a user kernel to get multimem / peer pointers to buffers of every
// rank of a communicator (NCCL >= 2.28: ncclMemAlloc + ncclCommWindowRegister(NCCL_WIN_COLL_SYMMETRIC) + ncclDevCommCreate
// with lsaMultimem).  Only HOST-side setup lives here; the kernels issue multimem.ld_reduce / multimem.st themselves on the
// raw addresses, so no NCCL device code is linked.
//
// This is a separate translation unit because it is the only one compiled against NCCL's own headers (struct layouts of
// ncclDevComm / ncclWindow_vidmem are version-specific: the runtime library must be the same major.minor as the headers,
// checked below).  Built with -DQB_HAVE_NCCL_DEVICE_API when the headers are found at build time (build.py); otherwise, or
// when the runtime library is older or the GPUs have no multicast support, qb_nvls_setup reports why and the caller keeps
// exchanging gradients with NCCL collectives.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "nvls_setup.h"

#ifdef QB_HAVE_NCCL_DEVICE_API
#include <nccl.h>
#include <nccl_device/impl/comm__types.h>

namespace {
struct Api {
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*MemAlloc)(void **, size_t) = nullptr;
    ncclResult_t (*MemFree)(void *) = nullptr;
    ncclResult_t (*WindowRegister)(ncclComm_t, void *, size_t, ncclWindow_t *, int) = nullptr;
    ncclResult_t (*WindowDeregister)(ncclComm_t, ncclWindow_t) = nullptr;
    ncclResult_t (*DevCommCreate)(ncclComm_t, ncclDevCommRequirements_t const *, ncclDevComm_t *) = nullptr;
    ncclResult_t (*DevCommDestroy)(ncclComm_t, ncclDevComm_t const *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *(*GetLastError)(ncclComm_t) = nullptr;
};
struct State {
    Api api;
    ncclComm_t comm = nullptr;
    ncclDevComm devComm;
    bool have_devcomm = false;
    ncclWindow_t win[QB_NVLS_MAX_BUFFERS] = {};
    void *mem[QB_NVLS_MAX_BUFFERS] = {};
    int n = 0;
};
thread_local char g_msg[512];
const char *fail(const char *what, const Api &api, ncclResult_t r, ncclComm_t comm) {
    snprintf(g_msg, sizeof(g_msg), "%s: %s%s%s", what, api.GetErrorString ? api.GetErrorString(r) : "NCCL error",
             (api.GetLastError && comm) ? " -- " : "", (api.GetLastError && comm) ? api.GetLastError(comm) : "");
    return g_msg;
}
inline char *add4g(char *base, uint32_t delta4g) { return base + ((uint64_t)delta4g << 32); }
}  // namespace

extern "C" const char *qb_nvls_setup(void *nccl_lib, void *comm_, QbNvlsBuffers *io) {
    if (!nccl_lib || !comm_ || !io || io->n < 1 || io->n > QB_NVLS_MAX_BUFFERS) return "qb_nvls_setup: bad arguments";
    State *st = new State();
    Api &a = st->api;
    a.GetVersion = (decltype(a.GetVersion))dlsym(nccl_lib, "ncclGetVersion");
    a.MemAlloc = (decltype(a.MemAlloc))dlsym(nccl_lib, "ncclMemAlloc");
    a.MemFree = (decltype(a.MemFree))dlsym(nccl_lib, "ncclMemFree");
    a.WindowRegister = (decltype(a.WindowRegister))dlsym(nccl_lib, "ncclCommWindowRegister");
    a.WindowDeregister = (decltype(a.WindowDeregister))dlsym(nccl_lib, "ncclCommWindowDeregister");
    a.DevCommCreate = (decltype(a.DevCommCreate))dlsym(nccl_lib, "ncclDevCommCreate");
    a.DevCommDestroy = (decltype(a.DevCommDestroy))dlsym(nccl_lib, "ncclDevCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(nccl_lib, "ncclGetErrorString");
    a.GetLastError = (decltype(a.GetLastError))dlsym(nccl_lib, "ncclGetLastError");
    int version = 0;
    if (!a.GetVersion || a.GetVersion(&version) != ncclSuccess) { delete st; return "ncclGetVersion unavailable"; }
    if (!a.MemAlloc || !a.MemFree || !a.WindowRegister || !a.WindowDeregister || !a.DevCommCreate || !a.DevCommDestroy) {
        snprintf(g_msg, sizeof(g_msg), "the loaded NCCL (%d) has no device API (symmetric windows / ncclDevCommCreate need >= 2.28)", version);
        delete st;
        return g_msg;
    }
    if (version / 100 != NCCL_VERSION_CODE / 100) {   // same major.minor: ncclDevComm / ncclWindow_vidmem layouts are per release
        snprintf(g_msg, sizeof(g_msg), "NCCL runtime %d does not match the headers this library was built against (%d)", version,
                 (int)NCCL_VERSION_CODE);
        delete st;
        return g_msg;
    }
    st->comm = (ncclComm_t)comm_;
    st->n = io->n;
    ncclResult_t r;
    for (int i = 0; i < io->n; i++) {
        const size_t bytes = (io->bytes[i] + 4095) & ~(size_t)4095;
        if ((r = a.MemAlloc(&st->mem[i], bytes)) != ncclSuccess) { const char *m = fail("ncclMemAlloc", a, r, st->comm); qb_nvls_teardown(st); return m; }
        if (cudaMemset(st->mem[i], 0, bytes) != cudaSuccess) { qb_nvls_teardown(st); return "cudaMemset of a symmetric buffer failed"; }
        if ((r = a.WindowRegister(st->comm, st->mem[i], bytes, &st->win[i], NCCL_WIN_COLL_SYMMETRIC)) != ncclSuccess || st->win[i] == nullptr) {
            const char *m = (r != ncclSuccess) ? fail("ncclCommWindowRegister", a, r, st->comm) : "ncclCommWindowRegister returned no symmetric window";
            qb_nvls_teardown(st);
            return m;
        }
    }
    ncclDevCommRequirements_t reqs;
    memset(&reqs, 0, sizeof(reqs));
    reqs.lsaMultimem = true;
    memset(&st->devComm, 0, sizeof(st->devComm));
    if ((r = a.DevCommCreate(st->comm, &reqs, &st->devComm)) != ncclSuccess) {
        const char *m = fail("ncclDevCommCreate(lsaMultimem)", a, r, st->comm);
        qb_nvls_teardown(st);
        return m;
    }
    st->have_devcomm = true;
    if (st->devComm.lsaMultimem.mcBasePtr == nullptr) { qb_nvls_teardown(st); return "no multicast mapping (NVLS disabled or unsupported on these GPUs)"; }
    if (st->devComm.lsaSize != st->devComm.nRanks || st->devComm.nRanks > QB_NVLS_MAX_PEERS) {
        snprintf(g_msg, sizeof(g_msg), "only %d of %d ranks are load/store accessible from this GPU", st->devComm.lsaSize, st->devComm.nRanks);
        qb_nvls_teardown(st);
        return g_msg;
    }
    io->lsa_rank = st->devComm.lsaRank;
    io->lsa_size = st->devComm.lsaSize;
    for (int i = 0; i < io->n; i++) {
        ncclWindow_vidmem w;
        if (cudaMemcpy(&w, st->win[i], sizeof(w), cudaMemcpyDeviceToHost) != cudaSuccess) { qb_nvls_teardown(st); return "reading a window descriptor failed"; }
        io->local[i] = st->mem[i];
        io->mc[i] = (char *)st->devComm.lsaMultimem.mcBasePtr + (uint64_t)w.mcOffset4K * 4096;
        // rank p's buffer as seen from this GPU; slot lsaRank is a second mapping of my own buffer (an alias of st->mem[i],
        // same physical memory), so the local pointer handed out is the allocation itself
        for (int p = 0; p < io->lsa_size; p++) io->peer[i][p] = add4g(w.lsaFlatBase, (uint32_t)p * w.stride4G);
        if (w.lsaRank != st->devComm.lsaRank || w.stride4G == 0) { qb_nvls_teardown(st); return "unexpected symmetric window descriptor"; }
        io->peer[i][w.lsaRank] = st->mem[i];
    }
    io->state = st;
    return nullptr;
}

extern "C" void qb_nvls_teardown(void *state) {
    State *st = (State *)state;
    if (!st) return;
    if (st->have_devcomm) st->api.DevCommDestroy(st->comm, &st->devComm);
    for (int i = 0; i < QB_NVLS_MAX_BUFFERS; i++) {
        if (st->win[i]) st->api.WindowDeregister(st->comm, st->win[i]);
        if (st->mem[i]) st->api.MemFree(st->mem[i]);
    }
    delete st;
}

#else  // built without NCCL's device-API headers

extern "C" const char *qb_nvls_setup(void *, void *, QbNvlsBuffers *) {
    return "this library was built without NCCL's device-API headers (nccl_device/*.h, NCCL >= 2.28)";
}
extern "C" void qb_nvls_teardown(void *) {}

#endif
