// host_fields.hpp -- one-time host-side arithmetic in F_q^2 and F_q[x]/(cubic) on BigUInt residues,
// used only to derive the per-pairing constants the reference computes in f_init_pairing
// (ecc/f_param.c:335-447) and d_init_pairing (ecc/d_param.c:993-1095).  Never on the data path.
#pragma once
#include "host_bigint.hpp"

namespace pbcb200 {

struct HostF2 {                      // a + b s, s^2 = beta
  BigUInt a, b;
};
struct HostF2Field {
  BigUInt q, beta;
  HostF2 mul(const HostF2& x, const HostF2& y) const {
    BigUInt t0 = BigUInt::mulmod(x.a, y.a, q), t1 = BigUInt::mulmod(x.b, y.b, q);
    HostF2 r;
    r.a = BigUInt::addmod(t0, BigUInt::mulmod(t1, beta, q), q);
    r.b = BigUInt::addmod(BigUInt::mulmod(x.a, y.b, q), BigUInt::mulmod(x.b, y.a, q), q);
    return r;
  }
  HostF2 neg(const HostF2& x) const {
    HostF2 r;
    r.a = x.a.is_zero() ? x.a : q - x.a;
    r.b = x.b.is_zero() ? x.b : q - x.b;
    return r;
  }
  HostF2 inv(const HostF2& x) const {
    // (a - b s)/(a^2 - beta b^2)   (arith/fieldquadratic.c:290-309)
    BigUInt d = BigUInt::submod(BigUInt::mulmod(x.a, x.a, q),
                                BigUInt::mulmod(beta, BigUInt::mulmod(x.b, x.b, q), q), q);
    d = BigUInt::invmod(d, q);
    HostF2 r;
    r.a = BigUInt::mulmod(x.a, d, q);
    r.b = BigUInt::mulmod(x.b, d, q);
    if (!r.b.is_zero()) r.b = q - r.b;
    return r;
  }
  HostF2 pow(const HostF2& x, const BigUInt& e) const {
    HostF2 r;
    r.a = BigUInt(1);
    for (size_t i = e.bits(); i-- > 0;) {
      r = mul(r, r);
      if (e.bit(i)) r = mul(r, x);
    }
    return r;
  }
  HostF2 scale(const HostF2& x, const BigUInt& k) const {
    HostF2 r;
    r.a = BigUInt::mulmod(x.a, k, q);
    r.b = BigUInt::mulmod(x.b, k, q);
    return r;
  }
};

struct HostF3 { BigUInt c[3]; };     // c0 + c1 x + c2 x^2 modulo x^3 + m2 x^2 + m1 x + m0
struct HostF3Field {
  BigUInt q, m[3];
  HostF3 mul(const HostF3& x, const HostF3& y) const {
    BigUInt d[5];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        d[i + j] = BigUInt::addmod(d[i + j], BigUInt::mulmod(x.c[i], y.c[j], q), q);
    for (int k = 4; k >= 3; k--) {       // x^k = -x^(k-3) (m0 + m1 x + m2 x^2)
      BigUInt t = d[k];
      d[k] = BigUInt();
      for (int i = 0; i < 3; i++)
        d[k - 3 + i] = BigUInt::submod(d[k - 3 + i], BigUInt::mulmod(t, m[i], q), q);
    }
    HostF3 r;
    for (int i = 0; i < 3; i++) r.c[i] = d[i];
    return r;
  }
  HostF3 pow(const HostF3& x, const BigUInt& e) const {
    HostF3 r;
    r.c[0] = BigUInt(1);
    for (size_t i = e.bits(); i-- > 0;) {
      r = mul(r, r);
      if (e.bit(i)) r = mul(r, x);
    }
    return r;
  }
};

// F_q[x]/(x^n + m[n-1] x^(n-1) + ... + m[0]) for any n (type g: n = 5)
struct HostPolyField {
  BigUInt q;
  std::vector<BigUInt> m;
  typedef std::vector<BigUInt> El;
  size_t n() const { return m.size(); }
  El mul(const El& x, const El& y) const {
    size_t N = n();
    std::vector<BigUInt> d(2 * N - 1);
    for (size_t i = 0; i < N; i++)
      for (size_t j = 0; j < N; j++)
        d[i + j] = BigUInt::addmod(d[i + j], BigUInt::mulmod(x[i], y[j], q), q);
    for (size_t k = 2 * N - 2; k >= N; k--) {
      BigUInt t = d[k];
      d[k] = BigUInt();
      for (size_t i = 0; i < N; i++)
        d[k - N + i] = BigUInt::submod(d[k - N + i], BigUInt::mulmod(t, m[i], q), q);
    }
    d.resize(N);
    return d;
  }
  El pow(const El& x, const BigUInt& e) const {
    El r(n());
    r[0] = BigUInt(1);
    for (size_t i = e.bits(); i-- > 0;) {
      r = mul(r, r);
      if (e.bit(i)) r = mul(r, x);
    }
    return r;
  }
};

}  // namespace pbcb200
