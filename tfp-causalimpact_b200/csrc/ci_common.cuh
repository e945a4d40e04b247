// ci_common.cuh -- shared device helpers for the sm_100a state-space kernels.
//
// Tile layout in HBM / shared memory (built once per problem by ci_set_data):
//   time is cut into blocks of TB = 32*KS steps; one tile holds TB rows of
//   [x_0 .. x_{p-1}, y]  (y = NaN marks a masked / padded step, x = 0 there),
//   row stride `ld` = (p+1) rounded up to ODD, plus ONE pad word after every
//   32 rows.  Element (tl, j) of a tile sits at  tl*ld + (tl>>5) + j.
//   With lane L owning the KS consecutive steps tl = KS*L + k this makes the
//   32 lanes of a warp hit 32 distinct banks for any fixed (k, j), and the
//   transposed access (fixed tl, j = lane) is contiguous.  A tile is a single
//   contiguous 16-byte-aligned range, so it is fetched with one
//   cp.async.bulk (TMA 1-D) into a shared-memory stage.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ci {

constexpr int KS = 8;            // consecutive time steps owned by one lane
constexpr int TB = 32 * KS;      // time steps per tile
constexpr int DSLOTS = 4;        // theta components per lane (dim <= 32*DSLOTS)
constexpr int MAX_DIM = 32 * DSLOTS;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline int tile_ld(int p) { return (p + 1) | 1; }
__host__ __device__ inline int tile_elems(int p) {
  int e = TB * tile_ld(p) + TB / 32;
  return (e + 3) & ~3;           // keep every tile 16-byte aligned (f32 and f64)
}
__host__ __device__ inline int tile_off(int tl, int ld) { return tl * ld + (tl >> 5); }

// ---------------------------------------------------------------------------
// numerics traits
// ---------------------------------------------------------------------------
template <typename R> struct Num;
template <> struct Num<float> {
  // MUFU.RCP (1 ulp) + one Newton step: ~0.5 ulp, no slow-path branch.  Inputs
  // are innovation variances F = P + sigma^2 > 0, never denormal in practice.
  static __device__ __forceinline__ float rcp(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r); }
  static __device__ __forceinline__ float rcp_fast(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  // 2^-floor(log2 x) for normal x > 0: an EXACT power-of-two rescaling factor (3 integer ops
  // instead of a MUFU.RCP on the critical path of every Moebius scan level)
  static __device__ __forceinline__ float pow2_rescale(float x) {
    return __int_as_float((254 - ((__float_as_int(x) >> 23) & 255)) << 23); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float log_fast(float x) { return __logf(x); }
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double pow2_rescale(double x) {
    return __longlong_as_double((long long)(2046 - ((__double_as_longlong(x) >> 52) & 2047)) << 52); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double log_fast(double x) { return ::log(x); }
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double nan() { return CUDART_NAN; }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Sum EACH of 16 per-lane values over the warp with 16 shuffles (a butterfly that halves the
// number of live values per level) instead of 16 x 5.  On return v[0] of lane L holds the
// warp total of value index L >> 1 (fixed order: deterministic).
template <typename R> __device__ __forceinline__ void warp_multi_sum16(R (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const R send = up ? v[i] : v[i + half];
      const R keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, off);
    }
  }
  v[0] += __shfl_xor_sync(FULL, v[0], 1);
}

// ---------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA 1-D) wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: every random number is a
// pure function of (seed, stream ids).  oracle/philox_np.py restates it.
// ---------------------------------------------------------------------------
struct Philox {
  static __host__ __device__ __forceinline__ void round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                        uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  static __host__ __device__ __forceinline__ uint4 gen(uint64_t seed, uint32_t c0, uint32_t c1,
                                                       uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c0, c1, c2, c3, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// uniform in (0,1), exactly representable in float32: ((x>>8)+0.5) * 2^-24
template <typename R> __device__ __forceinline__ R u01(uint32_t x) {
  return ((R)(x >> 8) + (R)0.5) * (R)5.9604644775390625e-08;
}
// Box-Muller: two normals from two uniforms.
template <typename R> __device__ __forceinline__ void box_muller(uint32_t x0, uint32_t x1, R& z0,
                                                                 R& z1) {
  const R u = u01<R>(x0), w = u01<R>(x1);
#ifdef CI_FAST_RNG
  if (sizeof(R) == 4) {
    // MUFU-based log / sin / cos: |error| ~ 1e-6 on a standard normal, far below the Monte-Carlo
    // error of anything built from the draws; ~15 instructions instead of ~70.
    const float rad_f = __fsqrt_rn(-2.0f * __logf((float)u));
    float fs, fc;
    __sincosf(6.283185307179586f * (float)w, &fs, &fc);
    z0 = (R)(rad_f * fc); z1 = (R)(rad_f * fs);
    return;
  }
#endif
  const R rad = Num<R>::sqrt((R)-2 * Num<R>::log(u));
  R s, c;
  if (sizeof(R) == 4) { float fs, fc; sincospif(2.0f * (float)w, &fs, &fc); s = fs; c = fc; }
  else { double ds, dc; sincospi(2.0 * (double)w, &ds, &dc); s = (R)ds; c = (R)dc; }
  z0 = rad * c; z1 = rad * s;
}
// RNG stream ids (counter word c1)
enum : uint32_t { RNG_MOMENTUM = 1, RNG_ACCEPT = 2, RNG_LEAPFROG = 3, RNG_SMOOTH = 4,
                  RNG_PREDICT = 5 };

}  // namespace ci
