// fp64 sincos / exp / log1p for the rollout kernels: <= 1-2 ulp, every coefficient a constant-bank operand.
//
// libdevice's versions cost ~75 / 48 / 75 issued instructions per call here, a third of a rollout step,
// mostly because each 64-bit polynomial coefficient is materialised with two uniform moves.  These versions
// keep the coefficients in __constant__ arrays (DFMA takes a c[bank][offset] operand directly) and skip the
// cases a trajectory never produces; arguments outside the fast range fall back to libdevice.
//   sincos: Cody-Waite reduction by pi/2 (two FMAs, |x| < 1e5) + the fdlibm kernel polynomials on [-pi/4, pi/4]
//   exp   : x <= 0, n = rint(x log2 e), degree-13 Taylor on |r| <= ln2/2, scaling by exponent arithmetic
//   log1p : e in [1e-4, 1]: 2 atanh(s) with s = e/(2+e), or ln2 + 2 atanh((e-1)/(e+3)) above sqrt(2)-1
// All FMAs are explicit (the rollout files are compiled with -fmad=false).
#pragma once

namespace qx {

static __constant__ double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,
                                       -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                                       -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kCosC[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03,
                                       2.48015872894767294178e-05, -2.75573143513906633035e-07,
                                       2.08757232129817482790e-09, -1.13596475577881948265e-11};
// 1/k!, k = 0..13
static __constant__ double kInvFact[14] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
                                           1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
                                           1.0 / 479001600, 1.0 / 6227020800.0};
// 2/(2k+1), k = 1..11
static __constant__ double kAtanhC[11] = {2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13,
                                          2.0 / 15, 2.0 / 17, 2.0 / 19, 2.0 / 21, 2.0 / 23};
static __constant__ double kMathC[8] = {
    6.36619772367581382433e-01,   // 0: 2/pi
    1.5707963267948966,           // 1: pi/2 hi
    6.123233995736766e-17,        // 2: pi/2 lo
    1.4426950408889634,           // 3: log2(e)
    6.93147180369123816490e-01,   // 4: ln2 hi
    1.90821492927058770002e-10,   // 5: ln2 lo
    0.6931471805599453,           // 6: ln2
    0.41421356237309503};         // 7: sqrt(2) - 1

// sin/cos of NA independent arguments.  The fast path is ONE basic block holding the NA polynomial chains, so
// the compiler can interleave them (a warp issues in order: independent chains only overlap inside a block).
template <int NA>
__device__ __forceinline__ void fast_sincos_n(const double* x, double* sp, double* cp) {
  bool ok = true;
#pragma unroll
  for (int a = 0; a < NA; ++a) ok = ok && (fabs(x[a]) < 1.0e5);      // false for NaN / inf too
  if (!ok) {
#pragma unroll
    for (int a = 0; a < NA; ++a) ::sincos(x[a], &sp[a], &cp[a]);
    return;
  }
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const int q = __double2int_rn(x[a] * kMathC[0]);
    const double n = (double)q;
    double r = fma(-n, kMathC[1], x[a]);
    r = fma(-n, kMathC[2], r);
    const double z = r * r;
    // sin(r) = r + r z (S1 + z (S2 + ... ))
    double ps = fma(z, kSinC[5], kSinC[4]);
    ps = fma(z, ps, kSinC[3]);
    ps = fma(z, ps, kSinC[2]);
    ps = fma(z, ps, kSinC[1]);
    ps = fma(z, ps, kSinC[0]);
    const double s = fma(r * z, ps, r);
    // cos(r) = w + (((1 - w) - z/2) + z (z (C1 + z (C2 + ...))))
    double pc = fma(z, kCosC[5], kCosC[4]);
    pc = fma(z, pc, kCosC[3]);
    pc = fma(z, pc, kCosC[2]);
    pc = fma(z, pc, kCosC[1]);
    pc = fma(z, pc, kCosC[0]);
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double c = w + fma(z, z * pc, (1.0 - w) - hz);
    const double u = (q & 1) ? c : s, v = (q & 1) ? s : c;
    sp[a] = (q & 2) ? -u : u;
    cp[a] = ((q + 1) & 2) ? -v : v;
  }
}

__device__ __forceinline__ void fast_sincos(double x, double* sp, double* cp) { fast_sincos_n<1>(&x, sp, cp); }

// 1/d for a normal d of either sign: approximation (2^-23) + three Newton steps (last one for the rounding).
// d = 0 gives NaN instead of inf - either way the trajectory is at the Euler-angle singularity and is rejected.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}

// exp(x) for x <= 0
__device__ __forceinline__ double fast_exp_neg(double x) {
  x = fmax(x, -700.0);
  const int q = __double2int_rn(x * kMathC[3]);
  const double n = (double)q;
  double r = fma(-n, kMathC[4], x);
  r = fma(-n, kMathC[5], r);
  double p = kInvFact[13];
#pragma unroll
  for (int k = 12; k >= 0; --k) p = fma(p, r, kInvFact[k]);
  return p * __hiloint2double((q + 1023) << 20, 0);     // q in [-1010, 0]: a normal power of two
}

// reciprocal of d in [1, 8): approximation (2^-23) + two Newton steps
__device__ __forceinline__ double fast_rcp_pos(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return r;
}

// log1p(e) for e in (0, 1]
__device__ __forceinline__ double fast_log1p_unit(double e) {
  const bool big = e > kMathC[7];
  const double num = big ? e - 1.0 : e;
  const double den = big ? e + 3.0 : e + 2.0;
  const double rc = fast_rcp_pos(den);
  double s = num * rc;
  s = fma(fma(-den, s, num), rc, s);                    // one correction: s = num/den to < 1 ulp
  const double z = s * s;
  double p = kAtanhC[10];
#pragma unroll
  for (int k = 9; k >= 0; --k) p = fma(p, z, kAtanhC[k]);
  const double l = fma(s * z, p, 2.0 * s);
  return big ? l + kMathC[6] : l;
}

}  // namespace qx
