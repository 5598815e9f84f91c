// Decimal text -> float32, one field at a time, written so that the SAME code compiles for the
// device (reader.cu) and for the host (tests/helpers/textparse_host.cpp fuzzes it against strtod).
//
// Replaces, for SURVEY.md section 8 row f-4, what the reference's scripts do on the host:
//   np.loadtxt(path, usecols=(0,1,2), dtype=np.float32)         (/root/reference/tests/correlations.py:29)
//   pd.read_csv(path, usecols=(0,1,2), delim_whitespace=True)    (/root/reference/tests/positions.py:25)
// NumPy's text reader converts a field to float32 as  (float)strtod(field) : decimal -> nearest
// double (round half to even) -> nearest float.  To be bit-identical the device has to reproduce
// the correctly rounded DOUBLE first, so the conversion below is exact integer arithmetic, not a
// chain of floating-point multiplications:
//   field = (-1)^s * w * 10^q,  w < 2^64 (up to 19 significant digits), |q| <= 27
//   q >= 0:  N = w * 5^q fits 128 bits  ->  round N to 53 bits, scale by 2^q
//   q <  0:  (w << lz << 64) / 5^-q as a 128-bit quotient (>= 64 significant bits) + remainder as
//            the sticky bit  ->  round to 53 bits, scale by 2^(-64 - lz + q)
// Everything outside that envelope (more than 19 significant digits with a non-zero tail,
// |q| > 27 after normalisation, "nan"/"inf" spellings) is reported as FIELD_SLOW and re-parsed by
// the host for that one field; nothing is ever approximated.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define JPS_HD __host__ __device__ __forceinline__
#else
#define JPS_HD inline
#endif

namespace jps {
namespace text {

enum FieldStatus {
  FIELD_OK = 0,
  FIELD_SLOW = 1,      // a number (or nan/inf) this parser does not convert exactly: host re-parses it
  FIELD_BAD = 2,       // not a number
  FIELD_MISSING = 3,   // the line has fewer columns than requested
};

typedef unsigned __int128 u128;

// characters that separate fields inside a line (np.loadtxt with delimiter=None: any whitespace)
JPS_HD bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
JPS_HD bool is_digit(char c) { return c >= '0' && c <= '9'; }

JPS_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

JPS_HD int msb128(u128 n) {                         // index of the highest set bit, n != 0
  const uint64_t hi = (uint64_t)(n >> 64), lo = (uint64_t)n;
  return hi ? 127 - clz64(hi) : 63 - clz64(lo);
}

JPS_HD double scale_pow2(double m, int e) {         // m * 2^e, exact (normal range only)
#ifdef __CUDA_ARCH__
  return ldexp(m, e);
#else
  return __builtin_ldexp(m, e);
#endif
}

// nearest double (ties to even) of  n * 2^e2 [+ a non-zero tail below n if `sticky`],  n != 0.
// With `sticky` the caller guarantees n has at least 55 significant bits.
JPS_HD double round_to_double(u128 n, int e2, bool sticky) {
  const int b = msb128(n);
  if (b <= 52) return scale_pow2((double)(uint64_t)n, e2);       // exact already
  const int shift = b - 52;
  uint64_t mant = (uint64_t)(n >> shift);                         // 53 bits
  const u128 rem = n & ((((u128)1) << shift) - 1);
  const u128 half = ((u128)1) << (shift - 1);
  if (rem > half || (rem == half && (sticky || (mant & 1)))) ++mant;   // may reach 2^53: still exact
  return scale_pow2((double)mant, e2 + shift);
}

JPS_HD uint64_t pow5(int k) {                       // k <= 27: 5^27 < 2^63
  uint64_t p = 1;
  for (int i = 0; i < k; ++i) p *= 5;
  return p;
}

// w != 0, |q| <= 27
JPS_HD double decimal_to_double(uint64_t w, int q) {
  // Clinger's fast path: w and 10^|q| are both exact doubles (w < 2^53, |q| <= 22), so ONE IEEE
  // multiplication / division is the correctly rounded result.  Covers "%.6f"-style catalogues;
  // the 128-bit integer path below takes the 16..19-digit mantissas of "%.18e".
  if (w < (1ull << 53) && q >= -22 && q <= 22) {
    double p10 = 1.0;
    const int a = q < 0 ? -q : q;
    for (int i = 0; i < a; ++i) p10 *= 10.0;         // exact: 10^22 < 2^53 * 2^22 and 5^22 < 2^53
    return q < 0 ? (double)w / p10 : (double)w * p10;
  }
  if (q >= 0) return round_to_double((u128)w * pow5(q), q, false);
  const int lz = clz64(w);
  const u128 num = ((u128)(w << lz)) << 64;
  const uint64_t d = pow5(-q);
  const u128 quo = num / d;
  const bool sticky = (num % d) != 0;
  return round_to_double(quo, -64 - lz + q, sticky);
}

// One field starting at p (first non-blank character; p < end).  A field ends at a blank, at `end`
// (end of line) or at the comment character.  On return `next` points just past the field.
JPS_HD int parse_field(const char* p, const char* end, char comment, float& out, const char*& next) {
  const char* s = p;
  bool neg = false;
  if (s < end && (*s == '+' || *s == '-')) { neg = (*s == '-'); ++s; }
  uint64_t w = 0;
  int nd = 0, dec_exp = 0;
  bool any = false, dot = false, tail = false;
  for (; s < end; ++s) {
    const char c = *s;
    if (is_digit(c)) {
      any = true;
      const int d = c - '0';
      if (w == 0 && d == 0) {                       // leading zero: no significant digit yet
        if (dot) --dec_exp;
      } else if (nd < 19) {
        w = w * 10 + (uint64_t)d;
        ++nd;
        if (dot) --dec_exp;
      } else {                                      // digit beyond the 19 kept ones
        if (d) tail = true;
        if (!dot) ++dec_exp;
      }
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      break;
    }
  }
  int status = FIELD_OK;
  if (!any) {
    // "nan", "inf", "infinity" (any case) are numbers for strtod; everything else is an error.
    // Either way skip to the end of the token; the host decides.
    const char c = (s < end) ? (char)(*s | 0x20) : '\0';
    status = (!dot && (c == 'n' || c == 'i')) ? FIELD_SLOW : FIELD_BAD;
    while (s < end && !is_blank(*s) && *s != comment) ++s;
    next = s;
    out = 0.0f;
    return status;
  }
  int e10 = 0;
  if (s < end && (*s == 'e' || *s == 'E')) {
    const char* t = s + 1;
    bool eneg = false;
    if (t < end && (*t == '+' || *t == '-')) { eneg = (*t == '-'); ++t; }
    if (t < end && is_digit(*t)) {
      for (; t < end && is_digit(*t); ++t)
        if (e10 < 100000) e10 = e10 * 10 + (*t - '0');
      if (eneg) e10 = -e10;
      s = t;
    }                                                // "1e" / "1e+": the exponent marker is junk -> BAD below
  }
  if (s < end && !is_blank(*s) && *s != comment) {   // trailing junk glued to the number
    while (s < end && !is_blank(*s) && *s != comment) ++s;
    next = s;
    out = 0.0f;
    return FIELD_BAD;
  }
  next = s;
  if (w == 0) {
    out = neg ? -0.0f : 0.0f;
    return FIELD_OK;
  }
  int q = dec_exp + e10;
  // bring q into the exact envelope when the digits allow it ("1e30", "1000000...e-40")
  while (q > 27 && w <= 0xFFFFFFFFFFFFFFFFull / 10) { w *= 10; --q; }
  while (q < -27 && w % 10 == 0) { w /= 10; ++q; }
  if (tail || q > 27 || q < -27) {
    out = 0.0f;
    return FIELD_SLOW;
  }
  const double v = decimal_to_double(w, q);
  out = (float)(neg ? -v : v);                       // second rounding, as NumPy's (float)strtod()
  return FIELD_OK;
}

// Parse the columns cols[0..ncols) (ascending or not, each >= 0) of one line [p, end).
// Returns 0 for a line with no fields (blank / comment only), 1 for a parsed row, and sets
// `status` to the worst field status met among the REQUESTED columns.
template <int MAXC>
JPS_HD int parse_line(const char* p, const char* end, char comment, const int* cols, int ncols, float* vals,
                      int& status) {
  status = FIELD_OK;
  int maxcol = 0;
  for (int c = 0; c < ncols; ++c) maxcol = cols[c] > maxcol ? cols[c] : maxcol;
  bool got[MAXC];
  for (int c = 0; c < ncols; ++c) { got[c] = false; vals[c] = 0.0f; }
  int col = 0;
  const char* s = p;
  for (;;) {
    while (s < end && is_blank(*s)) ++s;
    if (s >= end || *s == comment) break;
    if (col > maxcol) break;                         // later columns are never looked at
    bool wanted = false;
    for (int c = 0; c < ncols; ++c) wanted |= (cols[c] == col);
    if (wanted) {
      float v;
      const char* nx;
      const int st = parse_field(s, end, comment, v, nx);
      for (int c = 0; c < ncols; ++c)
        if (cols[c] == col) { vals[c] = v; got[c] = true; }
      status = st > status ? st : status;
      s = nx;
    } else {
      while (s < end && !is_blank(*s) && *s != comment) ++s;
    }
    ++col;
  }
  if (col == 0) return 0;                            // nothing on this line
  for (int c = 0; c < ncols; ++c)
    if (!got[c]) status = FIELD_MISSING;
  return 1;
}

}  // namespace text
}  // namespace jps
