// mt64_jump.cu — jump-ahead for std::mt19937_64 (host code; no kernels in this file).
//
// The reference draws the bootstrap replicates from ONE generator (src/BootstrapSample.cpp:48-60: seeded once,
// src/BootstrapSample.cpp:60-73 called per replicate), so replicate r starts r * bootstrap_count outputs into the stream.
// With the replicates spread over GPUs a rank has to stand at the start of ITS replicates without producing the draws of
// everybody else's.  MT19937-64 is linear over GF(2): with V_t = (upper 33 bits of x[t], x[t+1], ..., x[t+311]) the
// 19937-bit state at time t, V_{t+1} = A V_t, and for g(z) = z^J mod phi(z) (phi = the minimal polynomial of A, degree 19937)
//     V_{t+J} = g(A) V_t = XOR over the set coefficients g_i of V_{t+i},
// i.e. word m of the state J steps ahead is the XOR of the raw sequence words x[t+i+m] over the set bits i of g — 19937
// ordinary steps of the recurrence and ~10^4 x 312 word XORs, whatever J is.  phi is found once per process with
// Berlekamp-Massey on one output bit of the recurrence (2 x 19937 bits), z^J mod phi by square-and-multiply on bit vectors.
// Checked at first use against the plain recurrence; tests/test_mt64_jump.py holds it to an independent Python MT19937-64.
#include "common.cuh"

#include <array>
#include <map>
#include <memory>
#include <mutex>

#include "../../include/msweep_b200.h"

namespace mswb {
namespace {

constexpr int NW = 312, MM = 156, DEG = 19937;
constexpr int PW = 313;                      // words of a polynomial of degree <= 19937 (+ slack for shifted copies)
constexpr uint64_t UPPER = 0xFFFFFFFF80000000ull, LOWER = 0x7FFFFFFFull, MATRIX_A = 0xB5026F5AA96619E9ull;

inline uint64_t twist(uint64_t far, uint64_t cur, uint64_t nxt) {
  const uint64_t y = (cur & UPPER) | (nxt & LOWER);
  return far ^ (y >> 1) ^ ((y & 1ull) ? MATRIX_A : 0ull);
}

// raw[0..311] given; fills raw[312..n) with the recurrence
void extend(std::vector<uint64_t> &raw, size_t n) {
  const size_t have = raw.size();
  raw.resize(n);
  for (size_t k = have; k < n; ++k) raw[k] = twist(raw[k - NW + MM], raw[k - NW], raw[k - NW + 1]);
}

using Poly = std::array<uint64_t, PW>;

inline bool get_bit(const uint64_t *p, int i) { return (p[i >> 6] >> (i & 63)) & 1ull; }

struct Field {
  Poly phi{};                                  // minimal polynomial, bit i = coefficient of z^i
  std::vector<Poly> phi_shift;                 // phi << s, s = 0..63 (word-aligned XORs in the reduction)
  bool ok = false;

  // Berlekamp-Massey over GF(2) on s_t = top bit of x[t]; bit vectors throughout
  void find_phi() {
    std::vector<uint64_t> raw(NW);
    raw[0] = 5489ull;
    for (int i = 1; i < NW; ++i) raw[i] = 6364136223846793005ull * (raw[i - 1] ^ (raw[i - 1] >> 62)) + (uint64_t)i;
    const int n_bits = 2 * DEG + 64;
    extend(raw, (size_t)n_bits);
    std::vector<uint64_t> C(PW + 1, 0), Bs(PW + 1, 0), R(PW + 1, 0), T(PW + 1);
    C[0] = 1; Bs[0] = 2;                        // C = 1, B = 1 shifted by m = 1
    int L = 0;
    for (int N = 0; N < n_bits; ++N) {
      const int words = std::min(PW + 1, (std::max(L, N - L) + 2) / 64 + 2);
      uint64_t carry = raw[N] >> 63;            // R = (R << 1) | s_N: bit i of R is s_{N-i}
      for (int w = 0; w < words; ++w) { const uint64_t nc = R[w] >> 63; R[w] = (R[w] << 1) | carry; carry = nc; }
      uint64_t acc = 0;
      for (int w = 0; w <= L / 64; ++w) acc ^= C[w] & R[w];
      if (__builtin_parityll(acc)) {
        if (2 * L <= N) {
          T = C;
          for (int w = 0; w < words; ++w) C[w] ^= Bs[w];
          L = N + 1 - L;
          Bs = T;
        } else {
          for (int w = 0; w < words; ++w) C[w] ^= Bs[w];
        }
      }
      carry = 0;
      for (int w = 0; w < words; ++w) { const uint64_t nc = Bs[w] >> 63; Bs[w] = (Bs[w] << 1) | carry; carry = nc; }
      if (L > DEG) return;                      // not the sequence we think it is: ok stays false
    }
    if (L != DEG) return;
    // C(z) = 1 + c_1 z + ... + c_L z^L is the connection polynomial; phi is its reciprocal
    for (int i = 0; i <= DEG; ++i) if (get_bit(C.data(), i)) phi[(DEG - i) >> 6] |= 1ull << ((DEG - i) & 63);
    phi_shift.assign(64, Poly{});
    for (int s = 0; s < 64; ++s) {
      Poly &q = phi_shift[s];
      for (int w = 0; w < PW; ++w) {
        q[w] = phi[w] << s;
        if (s && w) q[w] |= phi[w - 1] >> (64 - s);
      }
    }
    ok = true;
  }

  // p (2 * PW words, degree < 2 * DEG) -> p mod phi in the low PW words
  void reduce(uint64_t *p) const {
    for (int i = 2 * DEG - 1; i >= DEG; --i) {
      if (!get_bit(p, i)) continue;
      const int sh = i - DEG;
      const Poly &q = phi_shift[sh & 63];
      uint64_t *dst = p + (sh >> 6);
      for (int w = 0; w < PW; ++w) dst[w] ^= q[w];
    }
  }

  // z^J mod phi
  Poly power_of_z(uint64_t J) const {
    static const auto spread = [] {             // byte -> its bits on the even positions of 16 bits
      std::array<uint16_t, 256> t{};
      for (int b = 0; b < 256; ++b) { uint16_t v = 0; for (int k = 0; k < 8; ++k) if (b >> k & 1) v |= (uint16_t)(1u << (2 * k)); t[b] = v; }
      return t;
    }();
    std::vector<uint64_t> r(2 * PW + 2, 0), sq(2 * PW + 2);
    r[0] = 1;
    for (int bit = 63; bit >= 0; --bit) {
      // square: coefficient i moves to 2 i
      std::fill(sq.begin(), sq.end(), 0);
      for (int w = 0; w < PW; ++w) {
        const uint64_t v = r[w];
        uint64_t lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) {
          lo |= (uint64_t)spread[(v >> (8 * b)) & 0xff] << (16 * b);
          hi |= (uint64_t)spread[(v >> (32 + 8 * b)) & 0xff] << (16 * b);
        }
        sq[2 * w] = lo; sq[2 * w + 1] = hi;
      }
      reduce(sq.data());
      std::copy(sq.begin(), sq.begin() + PW, r.begin());
      if ((J >> bit) & 1ull) {                  // times z
        uint64_t carry = 0;
        for (int w = 0; w < PW; ++w) { const uint64_t nc = r[w] >> 63; r[w] = (r[w] << 1) | carry; carry = nc; }
        if (get_bit(r.data(), DEG)) for (int w = 0; w < PW; ++w) r[w] ^= phi[w];
      }
    }
    Poly g{};
    std::copy(r.begin(), r.begin() + PW, g.begin());
    return g;
  }
};

void apply(const Poly &g, const uint64_t *state, uint64_t *out) {
  std::vector<uint64_t> raw(state, state + NW);
  extend(raw, (size_t)DEG + NW);
  uint64_t acc[NW] = {0};
  for (int w = 0; w < PW; ++w) {
    uint64_t bits = g[w];
    while (bits) {
      const int i = 64 * w + __builtin_ctzll(bits);
      bits &= bits - 1;
      const uint64_t *src = raw.data() + i;
      for (int m = 0; m < NW; ++m) acc[m] ^= src[m];
    }
  }
  std::copy(acc, acc + NW, out);
}

const Field &field() {
  static Field f;
  static std::once_flag once;
  std::call_once(once, [] {
    f.find_phi();
    if (!f.ok) return;
    // self-check against the plain recurrence before anything relies on it
    std::vector<uint64_t> raw(NW);
    raw[0] = 20231017ull;
    for (int i = 1; i < NW; ++i) raw[i] = 6364136223846793005ull * (raw[i - 1] ^ (raw[i - 1] >> 62)) + (uint64_t)i;
    const uint64_t J = 54321;
    uint64_t jumped[NW];
    apply(f.power_of_z(J), raw.data(), jumped);
    extend(raw, (size_t)J + NW);
    bool same = ((jumped[0] ^ raw[J]) & UPPER) == 0;
    for (int m = 1; m < NW && same; ++m) same = jumped[m] == raw[J + m];
    f.ok = same;
  });
  return f;
}

std::mutex g_cache_mutex;
std::map<uint64_t, std::shared_ptr<const Poly>> g_cache;      // J -> z^J mod phi (a bootstrap run needs two of them)

} // namespace

bool mt64_jump_available() { return field().ok; }

// state[312]: the generator's words at a refill boundary (std::mt19937_64 right after seeding, or with all 312 words
// consumed); out[312]: the words from which the stream continues after n_outputs more outputs, again at a refill boundary.
// The low 31 bits of out[0] are not part of the generator's state (nothing ever reads them).
void mt64_jump(const uint64_t *state, uint64_t n_outputs, uint64_t *out) {
  if (n_outputs == 0) { std::copy(state, state + NW, out); return; }
  const Field &f = field();
  if (!f.ok) throw Error("mt19937_64 jump-ahead failed its self-check");
  std::shared_ptr<const Poly> g;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_cache.find(n_outputs);
    if (it != g_cache.end()) g = it->second;
  }
  if (!g) {
    g = std::make_shared<const Poly>(f.power_of_z(n_outputs));
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    if (g_cache.size() >= 64) g_cache.clear();
    g_cache[n_outputs] = g;
  }
  apply(*g, state, out);
}

// Coefficients of z^J mod phi as 313 words (bit i = coefficient of z^i): the device applies them itself when a replicate's
// stream is generated in parallel segments (bootstrap.cu: mt64_chain_kernel).
void mt64_jump_poly(uint64_t n_outputs, uint64_t *bits_out) {
  const Field &f = field();
  if (!f.ok) throw Error("mt19937_64 jump-ahead failed its self-check");
  std::shared_ptr<const Poly> g;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_cache.find(n_outputs);
    if (it != g_cache.end()) g = it->second;
  }
  if (!g) {
    g = std::make_shared<const Poly>(f.power_of_z(n_outputs));
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    if (g_cache.size() >= 64) g_cache.clear();
    g_cache[n_outputs] = g;
  }
  std::copy(g->begin(), g->end(), bits_out);
}

} // namespace mswb

extern "C" int mswb_mt64_jump(const uint64_t *state, uint64_t n_outputs, uint64_t *state_out) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(state && state_out, "NULL argument");
    mswb::mt64_jump(state, n_outputs, state_out);
  });
}
