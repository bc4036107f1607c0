// Host-side lookup-table builders for libfov360.so (init-time only; never per frame).
//
// Why on the host: the reference's transform formulas truncate float/double transcendental
// results to integers ((int)(lambda*(exp(pow(t,4))-1)), ceil(0.5*n*pow(log(..),0.25)), ...).
// A 1-ulp difference between two libm implementations flips a table entry, which moves a
// whole sample box or source pixel - far outside the 1-LSB pixel tolerance.  The tables are
// tiny (O(W+H) entries), gaze-independent and built once per frame geometry, exactly like
// the reference's own init-time create_grid_kernel (sat_decoder.cc:139-170), so they are
// evaluated here with the host libm (the same one the parity oracle uses) and uploaded.
// Everything per-frame and per-pixel runs in CUDA.
//
// Arithmetic follows the OpenCL C typing of the reference kernels: float expressions use
// expf/powf/logf, double expressions exp/pow/ceil, float->int conversions truncate, and the
// translation unit is compiled with -ffp-contract=off.
#include <cmath>
#include <cstdlib>

#include "fov360_internal.h"

namespace fov {
namespace {

inline int sign_of(int v) { return (v > 0) - (v < 0); }

// lambda = dim / (e - 1) in float (sat_decoder_sample_rect_kernel.cl:266-267).
inline float lambda_for(int dim) { return (float)dim / (expf(1.0f) - 1); }

// Forward log-rectilinear map, magnitude only, float typing
// (sat_decoder_sample_rect_kernel.cl:269-273).
inline int forward_f32(int mag, int n_red, float lambda) {
  const float t = (float)(2.0f * mag / n_red);
  const int warped = (int)(lambda * (expf(powf(t, 4.0f)) - 1));
  return mag > warped ? mag : warped;
}

// Forward map with double typing (sat_decoder_interpolate_kernel.cl:56-65).
inline int forward_f64(int mag, int n_red, float lambda) {
  const int warped = (int)(lambda * (exp(pow(2.0 * mag / n_red, 4.0)) - 1));
  return mag > warped ? mag : warped;
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

// SATDecoder grid edges: edge[t] = floor((delta(u) + delta(u+1)) / 2.0f), u = t-1-n/2
// (sat_decoder_sample_rect_kernel.cl:260-294).
void build_sat_grid_edges(int ow, int oh, int W, int H, std::vector<int16_t> &xe,
                          std::vector<int16_t> &ye) {
  auto axis = [](int n_red, int n_full, std::vector<int16_t> &e) {
    const float lambda = lambda_for(n_full);
    e.resize(n_red + 1);
    for (int t = 0; t <= n_red; ++t) {
      const int u = t - 1 - n_red / 2;
      const int a = forward_f32(std::abs(u), n_red, lambda) * sign_of(u);
      const int b = forward_f32(std::abs(u + 1), n_red, lambda) * sign_of(u + 1);
      e[t] = (int16_t)floorf((a + b) / 2.0f);
    }
  };
  axis(ow, W, xe);
  axis(oh, H, ye);
}

// Inverse-warp table for one axis, entries for d = pos - centre in [-n_full, n_full]
// (sat_decoder_interpolate_kernel.cl:38-103, 135-142; everything that does not depend on the
// gaze once d is fixed).
void build_interp_axis(int n_full, int n_red, std::vector<InterpEntry> &lut) {
  const float lambda = n_full / (expf(1.0f) - 1);  // :11-12 (int / float)
  lut.resize(2 * (size_t)n_full + 1);
  for (int d = -n_full; d <= n_full; ++d) {
    InterpEntry e;
    int u = (int)(ceil(0.5 * n_red * powf(logf(std::abs(d) / lambda + 1), 0.25f)) * sign_of(d));
    if (std::abs(u) > std::abs(d) || u == 0) u = d;
    const int d_calc = forward_f64(std::abs(u), n_red, lambda) * sign_of(u);
    const int du = (d < 0) - (d > 0);  // (pos < centre) - (pos > centre), :75-76
    const int d_min = forward_f32(std::abs(u + du), n_red, lambda) * sign_of(u);
    const int rel_lo = d_min < d_calc ? d_min : d_calc;
    const int rel_hi = d_min < d_calc ? d_calc : d_min;
    e.idx_exact = (int16_t)clampi(u + n_red / 2, 0, n_red - 1);
    e.min_u = (int16_t)(u < u + du ? u : u + du);
    e.max_u = (int16_t)(u < u + du ? u + du : u);
    e.exact = (int16_t)(d_calc == d);
    e.rel_lo = (int16_t)rel_lo;
    e.rel_hi = (int16_t)rel_hi;
    if (rel_hi == rel_lo) {
      e.ratio = 0;
    } else {
      float r = (float)(d - rel_lo) / (rel_hi - rel_lo);
      e.ratio = fminf(fmaxf(r, (float)0), (float)1);
    }
    lut[(size_t)(d + n_full)] = e;
  }
}

// ImageSampler raw deltas (image_sampler_sample_rect_kernel.cl:68-87).
void build_img_grid_axes(int ow, int oh, int W, int H, std::vector<int16_t> &xd,
                         std::vector<int16_t> &yd) {
  auto axis = [](int n_red, int n_full, std::vector<int16_t> &v) {
    const float lambda = lambda_for(n_full);
    v.resize(n_red);
    for (int i = 0; i < n_red; ++i) {
      const int u = i - n_red / 2;
      v[i] = (int16_t)(forward_f32(std::abs(u), n_red, lambda) * sign_of(u));
    }
  };
  axis(ow, W, xd);
  axis(oh, H, yd);
}

// Log-polar factors (image_sampler_sample_logpolar_kernel.cl:2-3, 31-34):
// radius[i] = exp(10 * pow(i/ow, 1)), angle[j] = (float)(j/oh * 2 * 3.14159265359).
void build_logpolar_axes(int ow, int oh, std::vector<float> &radius, std::vector<float> &cs,
                         std::vector<float> &sn) {
  radius.resize(ow);
  cs.resize(oh);
  sn.resize(oh);
  for (int i = 0; i < ow; ++i) radius[i] = expf(10.0f * powf((float)i / ow, (float)1.0));
  for (int j = 0; j < oh; ++j) {
    const float a = (float)((float)j / oh * 2.0f * 3.14159265359);
    cs[j] = cosf(a);
    sn[j] = sinf(a);
  }
}

// Direction table of the log-polar inverse warp (image_sampler_interpolate_kernel.cl:46-51):
// cos / sin of (float)j / oh * 2.0f * M_PI - the float product is promoted by the double M_PI.
void build_logpolar_directions(int oh, std::vector<double2> &dir) {
  dir.resize(oh);
  for (int j = 0; j < oh; ++j) {
    const double a = (double)((float)j / oh * 2.0f) * 3.14159265358979323846;
    dir[j].x = cos(a);
    dir[j].y = sin(a);
  }
}

// ln table of the log-polar inverse warp: i_f = ow * (log(sqrt(d2)) / 10) (:28-33) = K ln(d2) with
// K = ow / 20.  One entry per (binary exponent e, top 7 mantissa bits k) of d2: c = the integer of
// the bin when it holds a single one (every d2 < 256: r = d2 / c - 1 is then exactly 0), else its
// centre, so that |r| <= 2^-8 and ln(d2) = ln c + r - r^2/2 + r^3/3 - r^4/4 to 5e-14.
void build_logpolar_lntab(int ow, std::vector<double2> &tab) {
  const double K = ow / 20.0;
  tab.resize((size_t)kLnExponents * 128);
  for (int e = 0; e < kLnExponents; ++e)
    for (int k = 0; k < 128; ++k) {
      const double lo = ldexp(1.0 + k / 128.0, e), hi = ldexp(1.0 + (k + 1) / 128.0, e);
      double c = 0.5 * (lo + hi);
      if (ceil(hi) - ceil(lo) <= 1.0) c = fmax(ceil(lo), 1.0);  // at most one integer in [lo, hi)
      tab[(size_t)e * 128 + k].x = 1.0 / c;
      tab[(size_t)e * 128 + k].y = K * log(c);
    }
}

// Half-width of the band around n + 0.5 inside which the single-precision j_f of the inverse
// log-polar warp may round to the other index than the reference's.  Both values are within 3e-7 rad
// of the true angle (reference: float division + atanf, 1.5e-7; device: reciprocal, polynomial and
// its evaluation, 2.6e-7) times oh / 2 pi; the reference then rounds the scaled angle to float
// (<= a quarter of the spacing at "j_f + 2 oh" <= 2.75 oh, being at least one binade smaller) and
// the "+ 2 oh" sum (half a spacing), the device rounds once (half a spacing): 1.25 spacings, 1.5 here.
float logpolar_round_zone(int oh) {
  const float top = 2.75f * oh;
  const float quantum = nextafterf(top, INFINITY) - top;
  return (float)(6e-7 * oh / (2.0 * 3.14159265358979323846)) + 1.5f * quantum;
}

// Gnomonic viewport constants (projections_program.cl:25-28): (center - 0.5) is promoted to
// double by the double literals 0.5 / PI and narrowed to float on assignment; sin/cos of the float.
GnomonicView make_gnomonic_view(float cx, float cy) {
  const double PI = 3.141592653589793;
  const float phi1 = (float)((cy - 0.5) * PI);
  GnomonicView v;
  v.lambda0 = (float)((cx - 0.5) * 2.0 * PI);
  v.sin_phi1 = sinf(phi1);
  v.cos_phi1 = cosf(phi1);
  return v;
}

}  // namespace fov
