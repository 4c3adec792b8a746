// Host-side setup of NVSwitch-multicast ("NVLS") and peer addresses for symmetric buffers (see nvls_setup.cu).
#pragma once
#include <cstddef>

constexpr int QB_NVLS_MAX_BUFFERS = 4;
constexpr int QB_NVLS_MAX_PEERS = 8;

struct QbNvlsBuffers {
    int n;                                              // in: number of buffers (every rank: same count, same sizes, same order)
    size_t bytes[QB_NVLS_MAX_BUFFERS];                  // in
    void *local[QB_NVLS_MAX_BUFFERS];                   // out: this rank's buffer (zero-filled ncclMemAlloc memory)
    void *mc[QB_NVLS_MAX_BUFFERS];                      // out: multicast address -- multimem.ld_reduce sums / multimem.st writes ALL ranks' buffers
    void *peer[QB_NVLS_MAX_BUFFERS][QB_NVLS_MAX_PEERS]; // out: unicast address of rank p's buffer
    int lsa_rank, lsa_size;                             // out
    void *state;                                        // out: pass to qb_nvls_teardown
};

// COLLECTIVE over the communicator.  Returns nullptr on success, else a message saying why multicast is unavailable
// (nothing is left allocated then).
extern "C" const char *qb_nvls_setup(void *nccl_lib, void *comm, QbNvlsBuffers *io);
extern "C" void qb_nvls_teardown(void *state);
