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

// {rc_i, T_i} of FlatMath::logn: rc_i = the double nearest 1 / (1 + i/64), T_i = -log(rc_i) correctly rounded (computed with 60-digit
// decimal arithmetic), i = -19 .. 27.
static __device__ const double2 kLogTable[47] = {
    {0x1.6c16c16c16c17p+0, -0x1.68ac83e9c6a15p-2},  // c = 1 + -19/64
    {0x1.642c8590b2164p+0, -0x1.522ae0738a3d7p-2},  // c = 1 + -18/64
    {0x1.5c9882b931057p+0, -0x1.3c25277333183p-2},  // c = 1 + -17/64
    {0x1.5555555555555p+0, -0x1.269621134db91p-2},  // c = 1 + -16/64
    {0x1.4e5e0a72f0539p+0, -0x1.1178e8227e47ap-2},  // c = 1 + -15/64
    {0x1.47ae147ae147bp+0, -0x1.f991c6cb3b37ap-3},  // c = 1 + -14/64
    {0x1.4141414141414p+0, -0x1.d1037f2655e7bp-3},  // c = 1 + -13/64
    {0x1.3b13b13b13b14p+0, -0x1.a93ed3c8ad9e5p-3},  // c = 1 + -12/64
    {0x1.3521cfb2b78c1p+0, -0x1.823c16551a3c0p-3},  // c = 1 + -11/64
    {0x1.2f684bda12f68p+0, -0x1.5bf406b543db0p-3},  // c = 1 + -10/64
    {0x1.29e4129e4129ep+0, -0x1.365fcb0159014p-3},  // c = 1 + -9/64
    {0x1.2492492492492p+0, -0x1.1178e8227e47ap-3},  // c = 1 + -8/64
    {0x1.1f7047dc11f70p+0, -0x1.da7276384469ep-4},  // c = 1 + -7/64
    {0x1.1a7b9611a7b96p+0, -0x1.9335e5d594988p-4},  // c = 1 + -6/64
    {0x1.15b1e5f75270dp+0, -0x1.4d3115d207eacp-4},  // c = 1 + -5/64
    {0x1.1111111111111p+0, -0x1.08598b59e3a06p-4},  // c = 1 + -4/64
    {0x1.0c9714fbcda3bp+0, -0x1.894aa149fb34bp-5},  // c = 1 + -3/64
    {0x1.0842108421084p+0, -0x1.0415d89e74440p-5},  // c = 1 + -2/64
    {0x1.0410410410410p+0, -0x1.0205658935837p-6},  // c = 1 + -1/64
    {0x1.0000000000000p+0, 0x0.0p+0},  // c = 1 + 0/64
    {0x1.f81f81f81f820p-1, 0x1.fc0a8b0fc03c4p-7},  // c = 1 + 1/64
    {0x1.f07c1f07c1f08p-1, 0x1.f829b0e7832f8p-6},  // c = 1 + 2/64
    {0x1.e9131abf0b767p-1, 0x1.77458f632dcffp-5},  // c = 1 + 3/64
    {0x1.e1e1e1e1e1e1ep-1, 0x1.f0a30c01162a8p-5},  // c = 1 + 4/64
    {0x1.dae6076b981dbp-1, 0x1.341d7961bd1d0p-4},  // c = 1 + 5/64
    {0x1.d41d41d41d41dp-1, 0x1.6f0d28ae56b4ep-4},  // c = 1 + 6/64
    {0x1.cd85689039b0bp-1, 0x1.a926d3a4ad562p-4},  // c = 1 + 7/64
    {0x1.c71c71c71c71cp-1, 0x1.e27076e2af2eap-4},  // c = 1 + 8/64
    {0x1.c0e070381c0e0p-1, 0x1.0d77e7cd08e5bp-3},  // c = 1 + 9/64
    {0x1.bacf914c1bad0p-1, 0x1.29552f81ff521p-3},  // c = 1 + 10/64
    {0x1.b4e81b4e81b4fp-1, 0x1.44d2b6ccb7d1cp-3},  // c = 1 + 11/64
    {0x1.af286bca1af28p-1, 0x1.5ff3070a793d6p-3},  // c = 1 + 12/64
    {0x1.a98ef606a63bep-1, 0x1.7ab890210d907p-3},  // c = 1 + 13/64
    {0x1.a41a41a41a41ap-1, 0x1.9525a9cf456b6p-3},  // c = 1 + 14/64
    {0x1.9ec8e951033d9p-1, 0x1.af3c94e80bff3p-3},  // c = 1 + 15/64
    {0x1.999999999999ap-1, 0x1.c8ff7c79a9a20p-3},  // c = 1 + 16/64
    {0x1.948b0fcd6e9e0p-1, 0x1.e27076e2af2e8p-3},  // c = 1 + 17/64
    {0x1.8f9c18f9c18fap-1, 0x1.fb9186d5e3e29p-3},  // c = 1 + 18/64
    {0x1.8acb90f6bf3aap-1, 0x1.0a324e27390e2p-2},  // c = 1 + 19/64
    {0x1.8618618618618p-1, 0x1.1675cababa60fp-2},  // c = 1 + 20/64
    {0x1.8181818181818p-1, 0x1.22941fbcf7966p-2},  // c = 1 + 21/64
    {0x1.7d05f417d05f4p-1, 0x1.2e8e2bae11d31p-2},  // c = 1 + 22/64
    {0x1.78a4c8178a4c8p-1, 0x1.3a64c556945eap-2},  // c = 1 + 23/64
    {0x1.745d1745d1746p-1, 0x1.4618bc21c5ec2p-2},  // c = 1 + 24/64
    {0x1.702e05c0b8170p-1, 0x1.51aad872df82ep-2},  // c = 1 + 25/64
    {0x1.6c16c16c16c17p-1, 0x1.5d1bdbf5809cap-2},  // c = 1 + 26/64
    {0x1.6816816816817p-1, 0x1.686c81e9b14adp-2},  // c = 1 + 27/64
};

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
  // log x = e ln2 + T[i] + log1p(r):  x = 2^e m with m in [sqrt(1/2), sqrt(2)), i = round(64 (m - 1)) in [-19, 27], c_i = 1 + i/64,
  // r = m rc_i - 1 with rc_i the double nearest 1 / c_i (ONE fma: exact product, one rounding) and T[i] = -log(rc_i) tabulated for
  // that rounded rc_i, so log m = T[i] + log1p(r) holds exactly; |r| <= 0.0111 and log1p is its degree-8 Taylor polynomial
  // (truncation < 2e-17 relative).  i = 0 has rc = 1, T = 0: next to x = 1 the result is log1p(m - 1), no cancellation.  15 FP64 ops
  // and one 16-byte table load (L1-resident, 752 bytes) against ~30 for the division-based atanh series this replaces; <= 2 ulp
  // (1.7 measured over [1e-304, 1e304] and around 1); 0 -> -inf, negative -> NaN, +inf -> +inf, NaN -> NaN by selection; denormals
  // are rescaled by 2^54.
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
    // round(64 m - 64) lands in the low word of 1.5 * 2^52 + (64 m - 64)
    const int i = __double2loint(fma(m, 64.0, 6755399441055744.0 - 64.0));
    const double2 tb = __ldg(&kLogTable[i + 19]);
    const double r = fma(m, tb.x, -1.0);
    double q = fma(-1.0 / 8.0, r, 1.0 / 7.0);
    q = fma(q, r, -1.0 / 6.0);
    q = fma(q, r, 1.0 / 5.0);
    q = fma(q, r, -1.0 / 4.0);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double pr = fma(q, r * r, r);  // log1p(r)
    const double ed = (double)e;
    double v = fma(ed, 6.93147180369123816490e-01, tb.y + fma(ed, 1.90821492927058770002e-10, pr));
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
