// Philox4x32-10 counter-based generator shared by the environment and heuristic kernels: every stream is a pure
// function of (seed, instance id, epoch, kind), so results do not depend on the launch shape or the GPU count.
// kinds in use: 0 profile table, 1 node capacities, 2 rule draw, 3 categorical action sample, 4 random heuristic.
#pragma once
#include <stdint.h>

namespace tpc {

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: streams are a pure function of (seed, ids) ----
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ inline U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    U4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// sequential draw helper: draw i of stream (seed, id, epoch, kind) = word i%4 of block i/4
struct Stream {
  uint32_t k0, k1, c0, c1, c2, kind;
  uint32_t blk = 0xffffffffu;
  U4 cur;
  __device__ uint32_t draw(uint32_t i) {
    const uint32_t b = i >> 2;
    if (b != blk) {
      cur = philox4x32_10(U4{c0, c1, c2, (kind << 28) | b}, k0, k1);
      blk = b;
    }
    const uint32_t w = i & 3;
    return w == 0 ? cur.x : (w == 1 ? cur.y : (w == 2 ? cur.z : cur.w));
  }
  // uniform integer in [lo, hi)
  __device__ int randint(uint32_t i, int lo, int hi) {
    return lo + (int)__umulhi(draw(i), (uint32_t)(hi - lo));
  }
};

__device__ __forceinline__ Stream make_stream(uint64_t seed, int64_t id, int64_t epoch, uint32_t kind) {
  Stream s;
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  s.c0 = (uint32_t)id; s.c1 = (uint32_t)((uint64_t)id >> 32);
  s.c2 = (uint32_t)epoch;
  s.kind = kind;
  return s;
}

// round_half_up(n, 2) = floor(n * 100 + 0.5) / 100 in fp32 (environment/custom/resource_v3/misc/utils.py:19-22)
__device__ __forceinline__ float rhu2(float n) {
  return __fdiv_rn(floorf(__fadd_rn(__fmul_rn(n, 100.0f), 0.5f)), 100.0f);
}

}  // namespace tpc
