// synth.cuh -- counter-based synthetic fitted tensor, identical bit for bit on the
// device (here) and on the host (metalquicha_b200/synth.py, NumPy uint64/float64).
//
// Element (mu >= nu) of auxiliary slab q:
//     r     = mix64(key(seed, q) ^ (mu << 32 | nu))
//     s     = sum of the four 16-bit fields of r                (Irwin-Hall, ~N)
//     x     = (s - 131070) * XNORM
//     d     = (mu - nu) * (8/n);   decay = 1 / (1 + d*d)
//     value = (x * decay) * scale
// Only IEEE add/mul/div are used, each individually rounded (no FMA contraction),
// which is what makes the two implementations agree exactly.
#pragma once
#include <cstdint>

namespace mqcb200 {

#define MQCB200_SYNTH_XNORM 2.6429156645611437e-05 /* 1/sqrt(4*(65536^2-1)/12) */

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__host__ __device__ inline double synth_value(uint64_t seed, uint32_t q, uint32_t mu, uint32_t nu,
                                              double inv_width, double scale) {
  const uint64_t key = mix64(seed ^ mix64(0x51ED270B0B4F3A1Dull + (uint64_t)q));
  const uint64_t r = mix64(key ^ (((uint64_t)mu << 32) | (uint64_t)nu));
  const int s = (int)(r & 0xFFFF) + (int)((r >> 16) & 0xFFFF) + (int)((r >> 32) & 0xFFFF) + (int)(r >> 48);
#ifdef __CUDA_ARCH__
  const double x = __dmul_rn((double)(s - 131070), MQCB200_SYNTH_XNORM);
  const double d = __dmul_rn((double)(mu - nu), inv_width);
  const double decay = __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(d, d)));
  return __dmul_rn(__dmul_rn(x, decay), scale);
#else
  const double x = (double)(s - 131070) * MQCB200_SYNTH_XNORM;
  const double d = (double)(mu - nu) * inv_width;
  volatile double dd = d * d;  // keep the host compiler from contracting
  const double decay = 1.0 / (1.0 + dd);
  volatile double xd = x * decay;
  return xd * scale;
#endif
}

}  // namespace mqcb200
