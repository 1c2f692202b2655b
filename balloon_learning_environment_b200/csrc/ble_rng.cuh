// Philox4x32-10 counter-based generator shared by the reset, agent and learner kernels.
// Included from inside `namespace ble`.
#pragma once

struct Philox {          // Philox4x32-10, one stream per (seed, balloon)
  uint32_t key[2], ctr[4], out[4];
  int have;
  __device__ void init(uint64_t seed, uint64_t stream) {
    key[0] = uint32_t(seed); key[1] = uint32_t(seed >> 32);
    ctr[0] = 0; ctr[1] = 0; ctr[2] = uint32_t(stream); ctr[3] = uint32_t(stream >> 32);
    have = 0;
  }
  __device__ void round(uint32_t* c, const uint32_t* k) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ uint32_t next() {
    if (have == 0) {
      uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
      uint32_t k[2] = {key[0], key[1]};
      for (int i = 0; i < 10; ++i) { round(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
      out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
      if (++ctr[0] == 0) ++ctr[1];
      have = 4;
    }
    return out[--have];
  }
  __device__ double uniform() {       // [0, 1) with 53 bits
    const uint64_t a = next(), b = next();
    return double(((a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ float uniform_f32() { return float(next() >> 8) * (1.0f / 16777216.0f); }   // [0,1), 24 bits
  __device__ double normal() {        // Box-Muller
    const double u1 = 1.0 - uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cos(2.0 * kPi * u2);
  }
  __device__ double gamma(double a) { // Marsaglia-Tsang, a >= 1
    const double dd = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
    for (;;) {
      const double x = normal();
      double v = 1.0 + c * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = 1.0 - uniform();
      if (log(u) < 0.5 * x * x + dd - dd * v + dd * log(v)) return dd * v;
    }
  }
};
