// euler_b200/csrc/rng.cuh — the reference's random stream (misc/rng.c:5-20 xorshift64* keeping
// the high 32 bits; randf() main.c:203-207) with JUMP-AHEAD, so that every marker a source cell
// appends gets exactly the draws the reference's sequential loop (main.c:284-291) would give it.
// Host and device: the same functions are checked on the CPU in tests/test_kernel_arith_host.py.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace euler {

// xorshift64 state transition is linear over GF(2) (misc/rng.c:7-9); jump[j] holds the 64
// columns of T^(2^j), so any number of draws can be skipped in O(64 log k).
__host__ __device__ __forceinline__ unsigned long long rng_jump(const unsigned long long* __restrict__ jump,
                                                       unsigned long long s, unsigned long long k) {
  for (int j = 0; k != 0; ++j, k >>= 1) {
    if (k & 1ull) {
      const unsigned long long* col = jump + (size_t)j * 64;
      unsigned long long out = 0;
      for (int b = 0; b < 64; ++b)
        if ((s >> b) & 1ull) out ^= col[b];
      s = out;
    }
  }
  return s;
}
__host__ __device__ __forceinline__ unsigned long long rng_step(unsigned long long s) {
  s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
  return s;
}
__host__ __device__ __forceinline__ float rng_float(unsigned long long s) {
  const unsigned int bits = (unsigned int)((s * 0x2545F4914F6CDD1Dull) >> 32);
  return (float)((double)bits / (double)0xFFFFFFFFu);        // main.c:206
}

// jump table: 64 matrices T^(2^j), 64 columns each (host, once per handle)
inline void rng_build_jump_table(unsigned long long* t) {
  // column b of T: image of the basis vector e_b under one xorshift64 step
  for (int b = 0; b < 64; ++b) {
    unsigned long long s = 1ull << b;
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    t[b] = s;
  }
  for (int j = 1; j < 64; ++j) {
    const unsigned long long* prev = t + (size_t)(j - 1) * 64;
    unsigned long long* cur = t + (size_t)j * 64;
    for (int b = 0; b < 64; ++b) {
      unsigned long long s = prev[b], out = 0;
      for (int k = 0; k < 64; ++k)
        if ((s >> k) & 1ull) out ^= prev[k];
      cur[b] = out;
    }
  }
}

}  // namespace euler
