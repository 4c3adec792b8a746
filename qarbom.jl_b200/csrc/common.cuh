// Shared device/host helpers: error plumbing, Philox4x32-10, uniform/normal conversion.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace qb {

typedef __nv_bfloat16 bf16;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define QB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw qb::Error(-2, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                    ":" + std::to_string(__LINE__) + ")");                              \
    } while (0)

#define QB_CHECK(cond, code, msg)                   \
    do {                                            \
        if (!(cond)) throw qb::Error((code), (msg)); \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ----------------------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al., SC'11).  The stream contract of this library:
//   uniform (draw, row, unit)  = u23( philox(ctr = (unit, row >> 2, draw_lo, draw_hi), key = seed)[row & 3] )
//   normal  (draw, row, unit)  = box_muller( philox(ctr = (unit, row >> 2, draw_lo, draw_hi | 1<<31), key) )[row & 3]
//                                (words x,y -> rows 4g, 4g+1 as r cos, r sin; words z,w -> rows 4g+2, 4g+3)
// with `row` the GLOBAL row (chain / mini-batch row) so results do not depend on the
// number of GPUs.  tests/philox_ref.py restates this in numpy.
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                          uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 24-bit uniform in [0,1): exactly representable in fp32 and fp64 (Box-Muller inputs).
__host__ __device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// 23-bit uniform in [0,1) = (x >> 9) * 2^-23, built from the mantissa bits: no int->float conversion
// (I2F shares the quarter-rate XU pipe with the sigmoid's ex2/rcp in the GEMM epilogue).
__device__ __forceinline__ float u23(uint32_t x) { return __uint_as_float(0x3F800000u | (x >> 9)) - 1.0f; }

struct RngKey {
    uint32_t seed_lo, seed_hi;
    uint32_t draw_lo, draw_hi;  // one id per sampling half-step, identical on every rank
};

#ifdef __CUDACC__
__device__ __forceinline__ Philox4 rng_uniform4(const RngKey &k, uint32_t unit, uint32_t row_group) {
    return philox4x32_10(unit, row_group, k.draw_lo, k.draw_hi, k.seed_lo, k.seed_hi);
}
__device__ __forceinline__ float rng_uniform(const RngKey &k, uint32_t unit, uint64_t row) {
    Philox4 p = rng_uniform4(k, unit, (uint32_t)(row >> 2));
    uint32_t r = (uint32_t)(row & 3);
    uint32_t v = r == 0 ? p.x : (r == 1 ? p.y : (r == 2 ? p.z : p.w));
    return u23(v);
}
// Box-Muller on two 23-bit uniforms built from mantissa bits (exact in fp32 and fp64): u1 = 1 - (a >> 9) 2^-23 in (0, 1],
// u2 = (b >> 9) 2^-23 in [0, 1);  z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1)  (|z| <= 5.65)
__device__ __forceinline__ void box_muller2(uint32_t a, uint32_t b, float &z0, float &z1) {
    const float u1 = 1.0f - u23(a), u2 = u23(b);
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z0 = r * cs; z1 = r * sn;
}
// four normals for rows (4*group .. 4*group+3)
__device__ __forceinline__ void rng_normal4(const RngKey &k, uint32_t unit, uint32_t row_group, float (&z)[4]) {
    Philox4 p = philox4x32_10(unit, row_group, k.draw_lo, k.draw_hi | 0x80000000u, k.seed_lo, k.seed_hi);
    box_muller2(p.x, p.y, z[0], z[1]);
    box_muller2(p.z, p.w, z[2], z[3]);
}
__device__ __forceinline__ float rng_normal(const RngKey &k, uint32_t unit, uint64_t row) {
    float z[4];
    rng_normal4(k, unit, (uint32_t)(row >> 2), z);
    const uint32_t r = (uint32_t)(row & 3);
    return r == 0 ? z[0] : (r == 1 ? z[1] : (r == 2 ? z[2] : z[3]));
}

// logistic: exact-ish fp32 (expf + IEEE divide) for the validation path ...
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }
// ... and 2 MUFU (ex2 + rcp) for tensor-core epilogues (|err| ~ 2e-7 absolute)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace qb
