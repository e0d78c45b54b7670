/* ss_pow.h -- `x ** y` of the reference, bit for bit: CPython's float_pow hands positive finite operands to libm's pow(),
 * which on x86_64 glibc 2.39 with FMA + AVX2 resolves to the __pow_fma build of sysdeps/ieee754/dbl-64/e_pow.c (log_inline +
 * exp_inline of Szabolcs Nagy's table-driven pow, < 0.52 ULP, NOT correctly rounded -- so "some pow" is not enough for
 * bit-exact welfare comparisons).  This is that build's main path restated operation by operation from its disassembly
 * (every fused multiply-add of the compiled code is an fma() here, every other operation a single IEEE operation; compile
 * with -ffp-contract=off / nvcc -fmad=false), with the tables read out of the same libm (ss_pow_tables.h, extract.py).
 * Used on both sides: oracle/ss_oracle.c checks it against libm's pow() itself (and tests/ do on millions of operands); the
 * CUDA engine runs the same code.  Domain of the main path: x > 0 finite and normal, 2^-65 <= |y| < 2^63,
 * 2^-54 <= |y log x| < 512 -- what a Cobb-Douglas welfare w ** (m / mt) with w >= 0.01 meets; anything else returns NaN
 * through *bad (callers treat that as a fault, never as a value).  x == 1 and y == 0 are answered by float_pow before libm. */
#ifndef SS_POW_H
#define SS_POW_H
#include <stdint.h>
#include "ss_pow_tables.h"

#ifdef __CUDACC__
#define SS_POW_FN __host__ __device__ static inline
#define SS_POW_FMA(a, b, c) fma((a), (b), (c))
#else
#include <math.h>
#include <string.h>
#define SS_POW_FN static inline
#define SS_POW_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

SS_POW_FN double ss_pow_asdouble(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
SS_POW_FN unsigned long long ss_pow_asuint(double d) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long u; memcpy(&u, &d, 8); return u;
#endif
}

/* tables: on the device they are read through the pointers the caller passes (global / constant memory copies) */
typedef struct {
    const unsigned long long* log_hdr;   /* ln2hi, ln2lo, A0..A6 */
    const unsigned long long* log_tab;   /* 128 x (invc, logc, logctail) */
    const unsigned long long* exp_hdr;   /* invln2N, shift, negln2hiN, negln2loN, C2..C5 */
    const unsigned long long* exp_tab;   /* 128 x (tail, value) */
} ss_pow_tables;

SS_POW_FN double ss_pow_pos(const ss_pow_tables* T, double x, double y, int* bad) {
    const unsigned long long ix = ss_pow_asuint(x), iy = ss_pow_asuint(y);
    const unsigned topx = (unsigned)(ix >> 52), topy = (unsigned)(iy >> 52) & 0x7ffu;
    if (topx - 1u > 0x7fdu || topy - 0x3beu > 0x7fu) { *bad = 1; return ss_pow_asdouble(0x7ff8000000000000ull); }
#define LH(i) ss_pow_asdouble(T->log_hdr[i])
#define EH(i) ss_pow_asdouble(T->exp_hdr[i])
    /* log_inline: x = 2^k z, z in [OFF, 2 OFF), log(x) = k ln2 + log(c) + log1p(z/c - 1) */
    const unsigned long long tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((long long)tmp >> 52);
    const double z = ss_pow_asdouble(ix - (tmp & 0xfff0000000000000ull));
    const double kd = (double)k;
    const double invc = ss_pow_asdouble(T->log_tab[3 * i]), logc = ss_pow_asdouble(T->log_tab[3 * i + 1]),
                 logctail = ss_pow_asdouble(T->log_tab[3 * i + 2]);
    const double t1 = SS_POW_FMA(kd, LH(0), logc);
    const double lo1 = SS_POW_FMA(kd, LH(1), logctail);
    const double r = SS_POW_FMA(z, invc, -1.0);
    const double ar = r * LH(2);
    const double q1 = SS_POW_FMA(r, LH(4), LH(3));
    const double q2 = SS_POW_FMA(r, LH(6), LH(5));
    const double t2 = r + t1;
    const double lo2 = (t1 - t2) + r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double lo3 = SS_POW_FMA(ar, r, -ar2);
    const double hi = t2 + ar2;
    const double q3 = SS_POW_FMA(r, LH(8), LH(7));
    const double lo4 = (t2 - hi) + ar2;
    const double q4 = SS_POW_FMA(q3, ar2, q2);
    const double q5 = SS_POW_FMA(ar2, q4, q1);
    double s = lo1 + lo2;
    s = s + lo3;
    s = s + lo4;
    const double lo = SS_POW_FMA(ar3, q5, s);
    const double lhi = hi + lo;
    const double llo = (hi - lhi) + lo;
    /* y log(x) as ehi + elo */
    const double ehi = y * lhi;
    const double e1 = SS_POW_FMA(lhi, y, -ehi);
    const double elo = SS_POW_FMA(y, llo, e1);
    const unsigned abstop = (unsigned)(ss_pow_asuint(ehi) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u > 0x3eu) { *bad = 1; return ss_pow_asdouble(0x7ff8000000000000ull); }
    /* exp_inline: exp(ehi + elo) = 2^(k/N) exp(r), r in [-ln2/2N, ln2/2N] */
    double kd2 = SS_POW_FMA(ehi, EH(0), EH(1));
    const unsigned long long ki = ss_pow_asuint(kd2);
    kd2 = kd2 - EH(1);
    double r2 = SS_POW_FMA(kd2, EH(2), ehi);
    r2 = SS_POW_FMA(kd2, EH(3), r2);
    const unsigned idx = 2u * (unsigned)(ki & 127u);
    const unsigned long long sbits = T->exp_tab[idx + 1] + (ki << 45);
    r2 = elo + r2;
    const double p1 = SS_POW_FMA(r2, EH(5), EH(4));
    const double tr = r2 + ss_pow_asdouble(T->exp_tab[idx]);
    const double rr = r2 * r2;
    const double p2 = SS_POW_FMA(r2, EH(7), EH(6));
    const double t = SS_POW_FMA(p1, rr, tr);
    const double r4 = rr * rr;
    const double tm = SS_POW_FMA(p2, r4, t);
    const double scale = ss_pow_asdouble(sbits);
    return SS_POW_FMA(tm, scale, scale);
#undef LH
#undef EH
}
#endif
