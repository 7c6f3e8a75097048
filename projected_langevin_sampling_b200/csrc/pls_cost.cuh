// Per-point likelihood cost and its derivative w.r.t. the untransformed prediction F, evaluated in registers.
// Follows the reference's formulas operation by operation (paths relative to the reference root,
// pls/ = src/projected_langevin_sampling/):
//   links     pls/link_functions.py:30-80
//   gaussian  pls/costs/gaussian.py:54-88       bernoulli  pls/costs/bernoulli.py:48-77
//   poisson   pls/costs/poisson.py:47-82        student-t  pls/costs/student_t.py:55-88
//   multimodal pls/costs/multimodal.py:37-77 (derivative: closed form of the autograd the reference runs, :79-91)
// For (cost, link) pairs without a hand-written derivative in the reference, the derivative is the chain rule
// (d cost / d mu) * link'(F), which is what its autograd fallback pls/costs/base.py:68-84 evaluates; torch's clip has
// zero gradient outside [jitter, 1 - jitter], reproduced in link_derivative.
#pragma once
#include "pls_common.cuh"

namespace pls {

__device__ __forceinline__ double clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// Arithmetic policies of the derivative functors.
//   LibMath : CUDA's IEEE double division and exp (<= 1 ulp).  Both contain a rarely taken branch / slow-path call, which
//             turns every evaluation into several basic blocks: independent evaluations cannot interleave.
//   FlatMath: branch-free equivalents for the register epilogue of the hot kernel, where four evaluations per call must
//             overlap their dependent FP64 chains.  div: MUFU.RCP64H seed, two Newton steps, one residual correction
//             (<= 1 ulp; +-0 and +-inf divisors give the IEEE results by selection; a denormal divisor is treated as 0);
//             exp: the hot loop's table-driven routine (<= 2 ulp), +inf above 709.78, clamped below -700.
struct LibMath {
  __device__ __forceinline__ double div(double a, double b) const { return a / b; }
  __device__ __forceinline__ double expo(double x) const { return exp(x); }
  __device__ __forceinline__ double logn(double x) const { return log(x); }
};
struct FlatMath {
  const double* tbl;  // 2^(j/64), j = 0..63, in shared memory
  __device__ __forceinline__ double div(double a, double b) const {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    double q = a * r;
    q = fma(fma(-b, q, a), r, q);
    const double ab = fabs(b);
    // a / +-0 = a * +-inf,  a / +-inf = a * +-0   (NaN where IEEE gives NaN)
    q = (ab < 2.2250738585072014e-308) ? a * copysign(__longlong_as_double(0x7ff0000000000000LL), b) : q;
    q = (ab > 1.7976931348623157e308) ? a * copysign(0.0, b) : q;
    return q;
  }
  __device__ __forceinline__ double expo(double x) const {
    const double v = gram_exp_fast(fmin(x, 709.782712893384), tbl);
    return (x > 709.782712893384) ? __longlong_as_double(0x7ff0000000000000LL) : ((x != x) ? x : v);
  }
  // log x = e ln2 + 2 atanh(s), x = 2^e m with m in [sqrt(1/2), sqrt(2)), s = (m - 1) / (m + 1), |s| <= 0.1716:
  // 2 atanh(s) = 2 s (1 + s^2/3 + ... + s^20/21) (truncation < 1e-18); m - 1 is exact, so there is no cancellation near 1.
  // <= 2 ulp; 0 -> -inf, negative -> NaN, +inf -> +inf, NaN -> NaN by selection; denormals are rescaled by 2^54.
  __device__ __forceinline__ double logn(double x) const {
    const bool tiny = x < 2.2250738585072014e-308;
    const double xs = tiny ? x * 18014398509481984.0 : x;
    int hi = __double2hiint(xs);
    int e = (hi >> 20) - 1023 - (tiny ? 54 : 0);
    hi = (hi & 0x000fffff) | 0x3ff00000;  // m in [1, 2)
    const bool upper = hi > 0x3ff6a09e;   // m > sqrt(2) (to the high word): halve it
    hi -= upper ? 0x00100000 : 0;
    e += upper ? 1 : 0;
    const double m = __hiloint2double(hi, __double2loint(xs));
    const double s = div(m - 1.0, m + 1.0);
    const double z = s * s;
    double q = fma(z, 1.0 / 21.0, 1.0 / 19.0);
    q = fma(q, z, 1.0 / 17.0);
    q = fma(q, z, 1.0 / 15.0);
    q = fma(q, z, 1.0 / 13.0);
    q = fma(q, z, 1.0 / 11.0);
    q = fma(q, z, 1.0 / 9.0);
    q = fma(q, z, 1.0 / 7.0);
    q = fma(q, z, 1.0 / 5.0);
    q = fma(q, z, 1.0 / 3.0);
    const double r = 2.0 * fma(s * z, q, s);
    const double ed = (double)e;
    double v = fma(ed, 6.93147180369123816490e-01, fma(ed, 1.90821492927058770002e-10, r));
    v = (x == 0.0) ? -__longlong_as_double(0x7ff0000000000000LL) : v;
    v = (x < 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : v;
    v = (x == __longlong_as_double(0x7ff0000000000000LL)) ? x : v;
    return (x != x) ? x : v;
  }
};

template <class M = LibMath>
__device__ __forceinline__ double link_transform(const pls_cost& c, double f, const M& m = M()) {
  switch (c.link_id) {
    case PLS_LINK_SQUARE:
      return f * f;
    case PLS_LINK_SIGMOID:
      return clip(m.div(1.0, 1.0 + m.expo(-f)), c.link_jitter, 1.0 - c.link_jitter);
    case PLS_LINK_PROBIT:
      return clip((1.0 + erf(m.div(f, c.probit_divisor))) / 2.0, c.link_jitter, 1.0 - c.link_jitter);
    default:
      return f;
  }
}

template <class M = LibMath>
__device__ __forceinline__ double link_derivative(const pls_cost& c, double f, const M& m = M()) {
  switch (c.link_id) {
    case PLS_LINK_SQUARE:
      return 2.0 * f;
    case PLS_LINK_SIGMOID: {
      const double s = m.div(1.0, 1.0 + m.expo(-f));
      return (s >= c.link_jitter && s <= 1.0 - c.link_jitter) ? s * (1.0 - s) : 0.0;
    }
    case PLS_LINK_PROBIT: {
      const double u = m.div(f, c.probit_divisor);
      const double p = (1.0 + erf(u)) / 2.0;
      // d/df [ (1 + erf(f / r)) / 2 ] = exp(-(f/r)^2) / (r sqrt(pi))
      return (p >= c.link_jitter && p <= 1.0 - c.link_jitter) ? m.div(m.expo(-u * u), c.probit_divisor * 1.7724538509055160273) : 0.0;
    }
    default:
      return 1.0;
  }
}

// c(y, F) for one training point
template <class M = LibMath>
__device__ __forceinline__ double cost_value(const pls_cost& c, double y, double f, const M& m = M()) {
  const double mu = link_transform(c, f, m);
  switch (c.cost_id) {
    case PLS_COST_GAUSSIAN: {
      const double e = mu - y;
      return m.div(1.0, 2.0 * c.observation_noise) * (e * e);
    }
    case PLS_COST_BERNOULLI:
      return -m.logn(mu) * y - m.logn(1.0 - mu) * (1.0 - y);
    case PLS_COST_POISSON:
      return -2.0 * (y * m.logn(fabs(f))) + mu;
    case PLS_COST_STUDENT_T: {
      const double e = mu - y;
      return 0.5 * (c.degrees_of_freedom + 1.0) * m.logn(1.0 + m.div(e * e, c.degrees_of_freedom * (c.scale * c.scale)));
    }
    case PLS_COST_MULTIMODAL: {
      const double s2 = c.observation_noise * c.observation_noise;
      const double e1 = y - mu + c.shift;
      const double e2 = y - mu;
      const double a1 = c.log_weight_1 + (-0.5 * m.div(e1 * e1, s2) - c.log_normaliser);
      const double a2 = c.log_weight_2 + (-0.5 * m.div(e2 * e2, s2) - c.log_normaliser);
      const double mx = fmax(a1, a2);
      return -(mx + m.logn(m.expo(a1 - mx) + m.expo(a2 - mx)));
    }
  }
  return 0.0;
}

// d c(y, F) / d F for one training point
template <class M = LibMath>
__device__ __forceinline__ double cost_derivative(const pls_cost& c, double y, double f, const M& m = M()) {
  if (c.closed_form) {
    if (c.cost_id == PLS_COST_GAUSSIAN && c.link_id == PLS_LINK_IDENTITY) return m.div(1.0, c.observation_noise) * (f - y);
    if (c.cost_id == PLS_COST_BERNOULLI && c.link_id == PLS_LINK_SIGMOID) {
      const double p = link_transform(c, f, m);  // the clipped probability, as bernoulli.py:72-77 uses it
      return -(y * (1.0 - p)) + (1.0 - y) * p;
    }
    if (c.cost_id == PLS_COST_POISSON && c.link_id == PLS_LINK_SQUARE) return -2.0 * m.div(y, f) + 2.0 * f;
    if (c.cost_id == PLS_COST_STUDENT_T && c.link_id == PLS_LINK_IDENTITY) {
      const double e = f - y;
      return (c.degrees_of_freedom + 1.0) * m.div(e, c.degrees_of_freedom * (c.scale * c.scale) + e * e);
    }
  }
  const double mu = link_transform(c, f, m);
  const double dmu = link_derivative(c, f, m);
  switch (c.cost_id) {
    case PLS_COST_GAUSSIAN:
      return m.div(mu - y, c.observation_noise) * dmu;
    case PLS_COST_BERNOULLI:
      return (m.div(-y, mu) + m.div(1.0 - y, 1.0 - mu)) * dmu;
    case PLS_COST_POISSON:
      return m.div(-2.0 * y, f) + dmu;
    case PLS_COST_STUDENT_T: {
      const double e = mu - y;
      return m.div((c.degrees_of_freedom + 1.0) * e, c.degrees_of_freedom * c.scale * c.scale + e * e) * dmu;
    }
    case PLS_COST_MULTIMODAL: {
      const double s2 = c.observation_noise * c.observation_noise;
      const double e1 = y - mu + c.shift;
      const double e2 = y - mu;
      const double a1 = c.log_weight_1 - m.div(0.5 * e1 * e1, s2);
      const double a2 = c.log_weight_2 - m.div(0.5 * e2 * e2, s2);
      const double mx = fmax(a1, a2);
      const double w1 = m.expo(a1 - mx), w2 = m.expo(a2 - mx);
      return m.div(-m.div(w1 * e1 + w2 * e2, w1 + w2), s2) * dmu;
    }
  }
  return 0.0;
}

}  // namespace pls
