// host_bigint.hpp -- small unsigned big-integer for ONE-TIME host-side constant setup.
//
// The reference derives its per-field constants with GMP at pairing_init time
// (arith/montfp.c:579-599: R, R^3, -p^-1; ecc/f_param.c:408-444; ecc/d_param.c:1035-1049).  This
// class does the same job without GMP: parameter text -> limbs, R^2 mod p, -p^-1 mod 2^32, the
// final-exponent cofactors and the Frobenius constants.  It is never used to compute a pairing.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

namespace pbcb200 {

class BigUInt {
 public:
  std::vector<uint32_t> w;  // little-endian, normalised (no leading zero words)

  BigUInt() {}
  BigUInt(uint64_t v) { while (v) { w.push_back((uint32_t)v); v >>= 32; } }

  static bool from_dec(const std::string& s, BigUInt* out) {
    BigUInt r;
    if (s.empty()) return false;
    for (char c : s) {
      if (c < '0' || c > '9') return false;
      r.mul_small(10);
      r.add_small((uint32_t)(c - '0'));
    }
    *out = r;
    return true;
  }

  bool is_zero() const { return w.empty(); }
  size_t bits() const {
    if (w.empty()) return 0;
    uint32_t t = w.back();
    size_t b = 0;
    while (t) { b++; t >>= 1; }
    return (w.size() - 1) * 32 + b;
  }
  bool bit(size_t i) const { return (i >> 5) < w.size() && ((w[i >> 5] >> (i & 31)) & 1u); }
  uint32_t word(size_t i) const { return i < w.size() ? w[i] : 0; }
  void to_words(uint32_t* out, size_t n) const { for (size_t i = 0; i < n; i++) out[i] = word(i); }

  static int cmp(const BigUInt& a, const BigUInt& b) {
    if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
    for (size_t i = a.w.size(); i-- > 0;)
      if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
  }
  bool operator==(const BigUInt& o) const { return cmp(*this, o) == 0; }
  bool operator<(const BigUInt& o) const { return cmp(*this, o) < 0; }

  BigUInt operator+(const BigUInt& o) const {
    BigUInt r;
    uint64_t c = 0;
    size_t n = std::max(w.size(), o.w.size());
    for (size_t i = 0; i < n || c; i++) {
      c += (uint64_t)word(i) + o.word(i);
      r.w.push_back((uint32_t)c);
      c >>= 32;
    }
    r.trim();
    return r;
  }
  // requires *this >= o
  BigUInt operator-(const BigUInt& o) const {
    BigUInt r;
    int64_t c = 0;
    for (size_t i = 0; i < w.size(); i++) {
      int64_t t = (int64_t)w[i] - o.word(i) + c;
      c = t < 0 ? -1 : 0;
      r.w.push_back((uint32_t)(t & 0xffffffffll));
    }
    r.trim();
    return r;
  }
  BigUInt operator*(const BigUInt& o) const {
    BigUInt r;
    if (w.empty() || o.w.empty()) return r;
    r.w.assign(w.size() + o.w.size(), 0);
    for (size_t i = 0; i < w.size(); i++) {
      uint64_t c = 0;
      for (size_t j = 0; j < o.w.size() || c; j++) {
        c += (uint64_t)r.w[i + j] + (uint64_t)w[i] * o.word(j);
        r.w[i + j] = (uint32_t)c;
        c >>= 32;
      }
    }
    r.trim();
    return r;
  }
  BigUInt shl(size_t k) const {
    BigUInt r;
    if (w.empty()) return r;
    r.w.assign(w.size() + k / 32 + 1, 0);
    for (size_t i = 0; i < w.size(); i++) {
      uint64_t v = (uint64_t)w[i] << (k & 31);
      r.w[i + k / 32] |= (uint32_t)v;
      r.w[i + k / 32 + 1] |= (uint32_t)(v >> 32);
    }
    r.trim();
    return r;
  }
  // binary long division: q = a / b, r = a % b
  static void divmod(const BigUInt& a, const BigUInt& b, BigUInt* q, BigUInt* r) {
    BigUInt quo, rem;
    quo.w.assign(a.w.size(), 0);
    for (size_t i = a.bits(); i-- > 0;) {
      rem = rem.shl(1);
      if (a.bit(i)) rem.add_small(1);
      if (!(rem < b)) {
        rem = rem - b;
        quo.w[i >> 5] |= 1u << (i & 31);
      }
    }
    quo.trim();
    if (q) *q = quo;
    if (r) *r = rem;
  }
  BigUInt operator%(const BigUInt& m) const { BigUInt r; divmod(*this, m, nullptr, &r); return r; }
  BigUInt operator/(const BigUInt& m) const { BigUInt q; divmod(*this, m, &q, nullptr); return q; }

  static BigUInt mulmod(const BigUInt& a, const BigUInt& b, const BigUInt& m) { return (a * b) % m; }
  static BigUInt submod(const BigUInt& a, const BigUInt& b, const BigUInt& m) {
    return a < b ? (a + m) - b : a - b;
  }
  static BigUInt addmod(const BigUInt& a, const BigUInt& b, const BigUInt& m) {
    BigUInt s = a + b;
    return s < m ? s : s - m;
  }
  static BigUInt powmod(const BigUInt& a, const BigUInt& e, const BigUInt& m) {
    BigUInt r(1), base = a % m;
    for (size_t i = e.bits(); i-- > 0;) {
      r = mulmod(r, r, m);
      if (e.bit(i)) r = mulmod(r, base, m);
    }
    return r;
  }
  // prime modulus
  static BigUInt invmod(const BigUInt& a, const BigUInt& p) { return powmod(a, p - BigUInt(2), p); }

 private:
  void trim() { while (!w.empty() && w.back() == 0) w.pop_back(); }
  void mul_small(uint32_t m) {
    uint64_t c = 0;
    for (auto& x : w) { c += (uint64_t)x * m; x = (uint32_t)c; c >>= 32; }
    if (c) w.push_back((uint32_t)c);
  }
  void add_small(uint32_t a) {
    uint64_t c = a;
    for (size_t i = 0; i < w.size() && c; i++) { c += w[i]; w[i] = (uint32_t)c; c >>= 32; }
    if (c) w.push_back((uint32_t)c);
  }
};

// -p^-1 mod 2^32 by Newton iteration (arith/montfp.c:590-596 does it with mpz_invert)
inline uint32_t neg_inv32(uint32_t p0) {
  uint32_t x = 1;
  for (int i = 0; i < 6; i++) x *= 2u - p0 * x;
  return (uint32_t)(0u - x);
}

}  // namespace pbcb200
