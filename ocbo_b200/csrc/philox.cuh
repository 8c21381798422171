// Philox4x32-10 counter-based generator + Box-Muller, bit-for-bit the layout
// contract restated in oracle/philox.py (checked against the Random123
// known-answer vectors there): results depend on (seed, trial, element index)
// only, never on the launch geometry or the world size.
#pragma once
#include <stdint.h>

namespace ocbo {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                              uint32_t& c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
}

// pair q of trial `trial` -> two standard normals (elements 2q and 2q+1)
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint32_t trial, uint64_t q,
                                                   double& z0, double& z1) {
  uint32_t c0 = (uint32_t)(q & 0xffffffffu), c1 = (uint32_t)(q >> 32), c2 = trial, c3 = 0x0CB0u;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  const uint64_t a = (((uint64_t)c0 << 32) | c1) >> 11;
  const uint64_t b = (((uint64_t)c2 << 32) | c3) >> 11;
  const double u1 = ((double)a + 0.5) * 1.1102230246251565e-16;  // 2^-53
  const double u2 = ((double)b + 0.5) * 1.1102230246251565e-16;
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

}  // namespace ocbo
