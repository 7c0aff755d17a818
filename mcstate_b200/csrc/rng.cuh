// xoshiro256 streams, one per particle, resident in registers for a whole launch.
//
// Replaces dust's per-particle generator underneath model$run / model$filter
// (R/particle_filter_state.R:65,151).  What the reference tree pins: the
// algorithm name "xoshiro256plus" (tests/testthat/test-pmcmc-parallel.R:41),
// 32 bytes per stream (R/pmcmc_parallel.R:130), stream partition per parameter
// set (tests/testthat/test-particle-filter.R:1316-1318), long-jump per chain
// (R/pmcmc_parallel.R:133,147).  The seeding idiom and jump polynomials are
// dust's [SURVEY.md A.1].
#pragma once
#include <cstdint>

#define MCB_HD __host__ __device__ __forceinline__

namespace mcb {

struct Xoshiro {
  uint64_t s0, s1, s2, s3;
};

MCB_HD uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// SCR = 0: "+" scrambler, SCR = 1: "**" scrambler; the state walk is shared.
template <int SCR>
MCB_HD uint64_t xo_next(Xoshiro &r) {
  const uint64_t out = SCR == 0 ? r.s0 + r.s3 : rotl64(r.s1 * 5, 7) * 9;
  const uint64_t t = r.s1 << 17;
  r.s2 ^= r.s0;
  r.s3 ^= r.s1;
  r.s1 ^= r.s2;
  r.s0 ^= r.s3;
  r.s2 ^= t;
  r.s3 = rotl64(r.s3, 45);
  return out;
}

// top 53 bits -> [0, 1)
template <int SCR>
MCB_HD double xo_real(Xoshiro &r) {
  return (double)(xo_next<SCR>(r) >> 11) * 0x1.0p-53;
}

inline uint64_t splitmix64_step(uint64_t &state) {
  uint64_t z = (state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// integer seed -> first stream: each round's output seeds the next round
inline void xo_seed(uint64_t seed, uint64_t s[4]) {
  for (int i = 0; i < 4; ++i) {
    uint64_t st = seed;
    seed = splitmix64_step(st);
    s[i] = seed;
  }
}

inline void xo_apply_poly(uint64_t s[4], const uint64_t poly[4]) {
  Xoshiro r{s[0], s[1], s[2], s[3]};
  uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int w = 0; w < 4; ++w)
    for (int b = 0; b < 64; ++b) {
      if ((poly[w] >> b) & 1ULL) { a0 ^= r.s0; a1 ^= r.s1; a2 ^= r.s2; a3 ^= r.s3; }
      xo_next<0>(r);
    }
  s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
}

static const uint64_t kJumpPoly[4] = {0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL,
                                      0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL};
static const uint64_t kLongJumpPoly[4] = {0x76e15d3efefdcbbfULL, 0xc5004e441c522fb3ULL,
                                          0x77710069854ee241ULL, 0x39109bb02acbe635ULL};

inline void xo_jump(uint64_t s[4]) { xo_apply_poly(s, kJumpPoly); }
inline void xo_long_jump(uint64_t s[4]) { xo_apply_poly(s, kLongJumpPoly); }

// ---------------------------------------------------------------------------
// Parallel stream derivation.  dust derives stream i by jumping i times from
// stream 0 -- serial in N.  A jump is a linear map J over GF(2)^256, so stream
// i = J^i s0 and J^i factors over the bits of i.  The host squares the 256x256
// bit matrix of J 31 times (columns stored as 4 words); on the device every
// particle applies the matrices of its set bits: O(log N) instead of O(N).
// ---------------------------------------------------------------------------
constexpr int kJumpLevels = 32;
// [level][column 0..255][word 0..3]
struct JumpTable {
  uint64_t col[kJumpLevels][256][4];
};

inline void jump_table_build(JumpTable &t) {
  for (int c = 0; c < 256; ++c) {
    uint64_t e[4] = {0, 0, 0, 0};
    e[c >> 6] = 1ULL << (c & 63);
    xo_jump(e);
    for (int w = 0; w < 4; ++w) t.col[0][c][w] = e[w];
  }
  for (int l = 1; l < kJumpLevels; ++l)
    for (int c = 0; c < 256; ++c) {
      uint64_t acc[4] = {0, 0, 0, 0};
      const uint64_t *v = t.col[l - 1][c];
      for (int b = 0; b < 256; ++b)
        if ((v[b >> 6] >> (b & 63)) & 1ULL)
          for (int w = 0; w < 4; ++w) acc[w] ^= t.col[l - 1][b][w];
      for (int w = 0; w < 4; ++w) t.col[l][c][w] = acc[w];
    }
}

}  // namespace mcb
