/*
 * fov_oracle.c - CPU restatement of the reference's foveation kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library, and only as the checker or the reported
 * CPU baseline.  The product (fov360 CUDA library) never links or calls it.
 *
 * Parity pinning: the reference repository has NO tests, golden vectors or
 * known-answer values for this path.  This restatement is
 * therefore pinned against the reference ITSELF: oracle/_ref/libfovref.so is the
 * reference's own .cl kernel sources compiled by g++ through ref_shim/clshim.h,
 * tests/test_oracle_vs_ref.py checks every function below bit-for-bit against it
 * (when /root/reference is present), and tests/golden/ holds hashes generated
 * from that library by tests/golden/make_golden.py.
 *
 * Arithmetic conventions (they decide truncation results, so they are part of
 * the contract): float expressions use the float libm entry points (expf, powf,
 * logf, ...), double expressions the double ones, exactly where OpenCL C / the
 * g++ shim would pick them; no FMA contraction (-ffp-contract=off); float->int
 * conversions truncate toward zero; uint sums wrap modulo 2^32.
 *
 * Every function cites the reference file:line it follows.  Paths are relative
 * to /root/reference/src/.
 */
#include <math.h>
#include <omp.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int g_threads = 0; /* 0 = OpenMP default */

void orc_set_threads(int n) { g_threads = n; }
int orc_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }
#define NT (g_threads > 0 ? g_threads : omp_get_max_threads())

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); }
static inline uint32_t uclamp(uint32_t v, uint32_t lo, uint32_t hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}
static inline int sgn(int v) { return (v > 0) - (v < 0); }
static inline float fclampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
/* mix(), OpenCL 1.2 6.12.4: x + (y - x) * a, each operation rounded to float. */
static inline float mixf(float a, float b, float t) { return a + (b - a) * t; }

/* lambda = dim / (e - 1), all in float: sat_decoder_sample_rect_kernel.cl:266-267 */
static inline float lambda_of(int dim) { return (float)dim / (expf(1.0f) - 1); }

/* Log-rectilinear forward map, float flavour:
 * max(|u|, (int)(lambda * (exp(pow(2|u|/n, 4)) - 1))) * sign(u)
 * sat_decoder_sample_rect_kernel.cl:269-273 (same text at :274-290,
 * image_sampler_sample_rect_kernel.cl:74-83, sat_decoder_interpolate_kernel.cl:77-89). */
static inline int delta_f32(int a /* |u| */, int n, float lambda) {
  float t = (float)(2.0f * a / n);
  return imax(a, (int)(lambda * (expf(powf(t, 4.0f)) - 1)));
}

/* Same map evaluated in double: sat_decoder_interpolate_kernel.cl:56-65. */
static inline int delta_f64(int a /* |u| */, int n, float lambda) {
  return imax(a, (int)(lambda * (exp(pow(2.0 * a / n, 4.0)) - 1)));
}

/* ------------------------------------------------------------------------- */
/* SATEncoder::EncodeFrameGPU  (sat_encoder.cc:67-135)                        */
/*   copy_image_kernel   sat_encoder_encode_kernels.cl:1-20                   */
/*   scan_rows_kernel    sat_encoder_encode_kernels.cl:44-58                  */
/*   scan_columns_kernel sat_encoder_encode_kernels.cl:60-74                  */
/* SAT layout: u32[H][W][3] dense (target_linesize = 3*W elements, :77).      */
/* ------------------------------------------------------------------------- */
void orc_sat_encode(uint32_t *sat, const uint8_t *src, int W, int H, int src_linesize) {
  const int bpp = src_linesize / W; /* :9 */
  const size_t row = (size_t)3 * W;
  /* widen + row scan (per row independent) */
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int y = 0; y < H; ++y) {
    const uint8_t *s = src + (size_t)y * src_linesize;
    uint32_t *d = sat + (size_t)y * row;
    uint32_t a0 = 0, a1 = 0, a2 = 0;
    for (int x = 0; x < W; ++x) {
      a0 += s[x * bpp + 0];
      a1 += s[x * bpp + 1];
      a2 += s[x * bpp + 2];
      d[3 * x + 0] = a0;
      d[3 * x + 1] = a1;
      d[3 * x + 2] = a2;
    }
  }
  /* column scan, in place; column chunks are independent */
  const int nt = NT;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int c = 0; c < nt; ++c) {
    size_t lo = row * c / nt, hi = row * (c + 1) / nt;
    for (int y = 1; y < H; ++y) {
      uint32_t *d = sat + (size_t)y * row;
      const uint32_t *p = d - row;
      for (size_t k = lo; k < hi; ++k) d[k] += p[k];
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SATDecoder::InitializeGrid (sat_decoder.cc:139-170) -> create_grid_kernel  */
/* sat_decoder_sample_rect_kernel.cl:243-295.                                 */
/* grid: int16[(oh+1)][(ow+1)][2]; x depends on the column only, y on the row.*/
/* ------------------------------------------------------------------------- */
static int16_t sat_grid_edge(int t, int n_out, float lambda) {
  int u = (t - 1) - n_out / 2;                            /* :260-264 */
  int d = delta_f32(abs(u), n_out, lambda) * sgn(u);      /* :269-273 */
  int dp = delta_f32(abs(u + 1), n_out, lambda) * sgn(u + 1); /* :274-279 */
  return (int16_t)floorf((d + dp) / 2.0f);                /* :293 */
}

/* Separable form: xedge[ow+1], yedge[oh+1]. */
void orc_sat_grid_edges(int16_t *xedge, int16_t *yedge, int ow, int oh, int W, int H) {
  const float lx = lambda_of(W), ly = lambda_of(H);
  for (int t = 0; t <= ow; ++t) xedge[t] = sat_grid_edge(t, ow, lx);
  for (int t = 0; t <= oh; ++t) yedge[t] = sat_grid_edge(t, oh, ly);
}

void orc_sat_create_grid(int16_t *grid, int ow, int oh, int W, int H) {
  int16_t *xe = (int16_t *)malloc(sizeof(int16_t) * (ow + 1));
  int16_t *ye = (int16_t *)malloc(sizeof(int16_t) * (oh + 1));
  orc_sat_grid_edges(xe, ye, ow, oh, W, H);
  for (int ty = 0; ty <= oh; ++ty)
    for (int tx = 0; tx <= ow; ++tx) {
      size_t p = ((size_t)ty * (ow + 1) + tx) * 2; /* :292 */
      grid[p] = xe[tx];
      grid[p + 1] = ye[ty];
    }
  free(xe);
  free(ye);
}

/* ------------------------------------------------------------------------- */
/* SATDecoder::SampleFrameRectGPU (sat_decoder.cc:301-348) ->                 */
/* sample_rect_kernel, sat_decoder_sample_rect_kernel.cl:138-241.             */
/* out: uchar4[oh][out_linesize/4]; only .xyz written, and only when the box  */
/* touches the frame (:197-200); everything else keeps its previous contents. */
/* ------------------------------------------------------------------------- */
void orc_sat_sample_rect(uint8_t *out, int ow, int oh, int out_linesize, const uint32_t *sat,
                         int W, int H, const int16_t *grid, float cx, float cy) {
  const int gw = ow + 1;
  const int o_linesize = out_linesize / 4; /* :153 */
  const int cxp = (int)(cx * W);           /* :176, float mul then truncation */
  const int cyp = (int)(cy * H);
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int j = 0; j < oh; ++j) {
    for (int i = 0; i < ow; ++i) {
      int dx = grid[((size_t)(j + 1) * gw + (i + 1)) * 2];      /* :168-169 */
      int dxm = grid[((size_t)(j + 1) * gw + i) * 2];           /* :170-171 */
      int dy = grid[((size_t)(j + 1) * gw + (i + 1)) * 2 + 1];  /* :172-173 */
      int dym = grid[((size_t)j * gw + (i + 1)) * 2 + 1];       /* :174-175 */
      int px = cxp + dx, py = cyp + dy;                         /* :176-177 */
      int mx = cxp + dxm, my = cyp + dym;                       /* :178-179 */
      if (px >= W && mx >= W) {                                 /* :181-187 */
        px -= W;
        mx -= W;
      } else if (px < 0 && mx < 0) {
        px += W;
        mx += W;
      }
      if (!(((px >= 0 && px < W) || (mx >= 0 && mx < W)) &&
            ((py >= 0 && py < H) || (my >= 0 && my < H)))) /* :197-200 */
        continue;
      px = iclamp(px, 1, W - 1); /* :201-204 */
      py = iclamp(py, 1, H - 1);
      mx = iclamp(mx, 0, px - 1);
      my = iclamp(my, 0, py - 1);
      uint8_t *o = out + ((size_t)j * o_linesize + i) * 4; /* :205 */
      /* px > 0 && py > 0 always holds after the clamp, so only :206-217 is live. */
      const uint32_t *tl = sat + ((size_t)my * W + mx) * 3;
      const uint32_t *tr = sat + ((size_t)my * W + px) * 3;
      const uint32_t *bl = sat + ((size_t)py * W + mx) * 3;
      const uint32_t *br = sat + ((size_t)py * W + px) * 3;
      uint32_t area = (uint32_t)((px - mx) * (py - my)); /* :211 */
      for (int c = 0; c < 3; ++c)
        o[c] = (uint8_t)((br[c] - tr[c] + tl[c] - bl[c]) / area); /* :212-217 */
    }
  }
}

/* ------------------------------------------------------------------------- */
/* SATDecoder::InterpolateFrameRectGPU (sat_decoder.cc:887-927) ->            */
/* interpolate_rect_kernel, sat_decoder_interpolate_kernel.cl:1-152.          */
/* Both buffers are dense uchar3 arrays with a 4-byte pixel stride.  The 4th  */
/* byte of a written pixel is the source pixel's 4th byte on the exact-hit    */
/* path (struct copy, :72) and 0 on the interpolated path (convert_uchar3).   */
/* ------------------------------------------------------------------------- */

/* Per-axis part of the kernel.  All of it depends only on (pos, centre). */
typedef struct {
  int exact;     /* delta_calculated == delta (:67) */
  int idx_exact; /* clamp(u + n_red/2, 0, n_red-1) (:69-70) */
  int idx_lo;    /* clamp(min_u + n_red/2, ...) (:118-133) */
  int idx_hi;
  float ratio; /* :135-142 */
} axis_desc;

static axis_desc interp_axis(int pos, int centre, int n_full, int n_red, float lambda,
                             int wraps) {
  axis_desc r;
  int offset = 0;
  if (wraps) { /* :26-33, x axis only */
    if (pos - centre > n_full / 2) {
      pos -= n_full;
      offset = 1;
    } else if (pos - centre < -n_full / 2) {
      pos += n_full;
      offset = 1;
    }
  }
  int d = pos - centre; /* :38, :42 */
  /* :43-48: double * float, ceil in double */
  int u = (int)(ceil(0.5 * n_red * powf(logf(abs(d) / lambda + 1), 0.25f)) * sgn(d));
  if (abs(u) > abs(d) || u == 0) u = d;                 /* :50-55 */
  int d_calc = delta_f64(abs(u), n_red, lambda) * sgn(u); /* :56-65 */
  r.exact = (d_calc == d);
  r.idx_exact = iclamp(u + n_red / 2, 0, n_red - 1);
  int du = (pos < centre) - (pos > centre);                  /* :75-76 */
  int d_min = delta_f32(abs(u + du), n_red, lambda) * sgn(u); /* :77-89 */
  int lo = imin(centre + d_min, centre + d_calc);            /* :91-98 */
  int hi = imax(centre + d_min, centre + d_calc);
  int min_u = imin(u, u + du), max_u = imax(u, u + du); /* :100-103 */
  if (wraps) {
    if (lo < 0 && !offset) min_u = max_u;        /* :105-107 */
    if (hi >= n_full && !offset) max_u = min_u;  /* :108-110 */
  } else {
    if (lo < 0) min_u = max_u;       /* :111-113 */
    if (hi >= n_full) max_u = min_u; /* :114-116 */
  }
  r.idx_lo = iclamp(min_u + n_red / 2, 0, n_red - 1);
  r.idx_hi = iclamp(max_u + n_red / 2, 0, n_red - 1);
  r.ratio = (hi == lo) ? 0 : fclampf((float)(pos - lo) / (hi - lo), (float)0, (float)1);
  return r;
}

void orc_sat_interpolate_rect(uint8_t *out, int W, int H, const uint8_t *reduced, int ow, int oh,
                              float cx, float cy) {
  const float lx = W / (expf(1.0f) - 1); /* :11-12 */
  const float ly = H / (expf(1.0f) - 1);
  const int cxp = (int)(cx * W); /* :24-25 */
  const int cyp = (int)(cy * H);
  axis_desc *xs = (axis_desc *)malloc(sizeof(axis_desc) * W);
  for (int x = 0; x < W; ++x) xs[x] = interp_axis(x, cxp, W, ow, lx, 1);
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int y = 0; y < H; ++y) {
    axis_desc ya = interp_axis(y, cyp, H, oh, ly, 0);
    for (int x = 0; x < W; ++x) {
      const axis_desc *xa = &xs[x];
      uint8_t *o = out + ((size_t)y * W + x) * 4; /* :16 */
      if (xa->exact && ya.exact) {                /* :67-72 */
        memcpy(o, reduced + ((size_t)ya.idx_exact * ow + xa->idx_exact) * 4, 4);
        continue;
      }
      const uint8_t *tl = reduced + ((size_t)ya.idx_lo * ow + xa->idx_lo) * 4;
      const uint8_t *tr = reduced + ((size_t)ya.idx_lo * ow + xa->idx_hi) * 4;
      const uint8_t *bl = reduced + ((size_t)ya.idx_hi * ow + xa->idx_lo) * 4;
      const uint8_t *br = reduced + ((size_t)ya.idx_hi * ow + xa->idx_hi) * 4;
      for (int c = 0; c < 3; ++c) { /* :143-150 */
        float l = mixf((float)tl[c], (float)bl[c], ya.ratio);
        float r = mixf((float)tr[c], (float)br[c], ya.ratio);
        o[c] = (uint8_t)(int)mixf(l, r, xa->ratio);
      }
      o[3] = 0;
    }
  }
  free(xs);
}

/* ------------------------------------------------------------------------- */
/* SATDecoder::DecodeFrameGPU (sat_decoder.cc:176-210) -> decode_kernel,      */
/* sat_decoder_decode_kernel.cl:1-58.  1x1 box = exact inverse of the SAT.    */
/* ------------------------------------------------------------------------- */
void orc_sat_decode(uint8_t *out, int out_linesize, const uint32_t *sat, int W, int H) {
  const int bpp = out_linesize / W; /* :12 */
  const size_t row = (size_t)3 * W;
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      uint8_t *o = out + (size_t)y * out_linesize + (size_t)x * bpp; /* :17 */
      const uint32_t *br = sat + y * row + (size_t)3 * x;
      for (int c = 0; c < 3; ++c) {
        uint32_t v;
        if (x > 0 && y > 0) /* :21-33 */
          v = br[c] - br[c - (ptrdiff_t)row] + br[c - (ptrdiff_t)row - 3] - br[c - 3];
        else if (x > 0) /* :34-42 */
          v = br[c] - br[c - 3];
        else if (y > 0) /* :43-51 */
          v = br[c] - br[c - (ptrdiff_t)row];
        else /* :52-57 */
          v = sat[c];
        o[c] = (uint8_t)uclamp(v, 0u, 255u);
      }
    }
}

/* ------------------------------------------------------------------------- */
/* ImageSampler::InitializeGrid (image_sampler.cc:170-202) -> create_grid_    */
/* kernel, image_sampler_sample_rect_kernel.cl:48-88.  int16[oh][ow][2], raw  */
/* deltas (no midpoints).                                                     */
/* ------------------------------------------------------------------------- */
void orc_img_grid_axes(int16_t *xd, int16_t *yd, int ow, int oh, int W, int H) {
  const float lx = lambda_of(W), ly = lambda_of(H);
  for (int i = 0; i < ow; ++i) {
    int u = i - ow / 2;
    xd[i] = (int16_t)(delta_f32(abs(u), ow, lx) * sgn(u)); /* :74-78 */
  }
  for (int j = 0; j < oh; ++j) {
    int v = j - oh / 2;
    yd[j] = (int16_t)(delta_f32(abs(v), oh, ly) * sgn(v)); /* :79-83 */
  }
}

void orc_img_create_grid(int16_t *grid, int ow, int oh, int W, int H) {
  int16_t *xd = (int16_t *)malloc(sizeof(int16_t) * ow);
  int16_t *yd = (int16_t *)malloc(sizeof(int16_t) * oh);
  orc_img_grid_axes(xd, yd, ow, oh, W, H);
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      grid[((size_t)j * ow + i) * 2] = xd[i]; /* :85-87 */
      grid[((size_t)j * ow + i) * 2 + 1] = yd[j];
    }
  free(xd);
  free(yd);
}

/* ImageSampler::SampleFrameRectGPU (image_sampler.cc:249-299) -> sample_rect_kernel,
 * image_sampler_sample_rect_kernel.cl:1-46. */
void orc_img_sample_rect(uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src, int W,
                         int H, int src_linesize, const int16_t *grid, float cx, float cy) {
  const int sbpp = src_linesize / W, obpp = out_linesize / ow; /* :9-10 */
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      int dx = grid[((size_t)j * ow + i) * 2], dy = grid[((size_t)j * ow + i) * 2 + 1];
      int x = (int)(cx * W + dx); /* :26-27: float + int, then truncation */
      int y = (int)(cy * H + dy);
      if (x >= W) /* :29-33 */
        x -= W;
      else if (x < 0)
        x += W;
      if (x >= 0 && x < W && y >= 0 && y < H) { /* :35-43 */
        uint8_t *o = out + (size_t)j * out_linesize + (size_t)i * obpp;
        const uint8_t *s = src + (size_t)y * src_linesize + (size_t)x * sbpp;
        o[0] = s[0];
        o[1] = s[1];
        o[2] = s[2];
      }
    }
}

/* ------------------------------------------------------------------------- */
/* ImageSampler::InitializeLogpolarGrid (image_sampler.cc:204-247) ->         */
/* create_logpolar_grid_kernel, image_sampler_sample_logpolar_kernel.cl:5-39. */
/* grid[j][i] = ((int)(r[i]*c[j]), (int)(r[i]*s[j])) with float r, c, s.      */
/* ------------------------------------------------------------------------- */
#define ORC_PI 3.14159265359 /* :2 */

void orc_img_logpolar_axes(float *radius, float *cs, float *sn, int ow, int oh) {
  for (int i = 0; i < ow; ++i) radius[i] = expf(10.0f * powf((float)i / ow, (float)1.0)); /* :31 */
  for (int j = 0; j < oh; ++j) {
    float a = (float)((float)j / oh * 2.0f * ORC_PI); /* :32 */
    cs[j] = cosf(a);
    sn[j] = sinf(a); /* :34 */
  }
}

void orc_img_create_logpolar_grid(int16_t *grid, int ow, int oh, int W, int H) {
  (void)W;
  (void)H;
  float *r = (float *)malloc(sizeof(float) * ow);
  float *c = (float *)malloc(sizeof(float) * oh);
  float *s = (float *)malloc(sizeof(float) * oh);
  orc_img_logpolar_axes(r, c, s, ow, oh);
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      int dx = (int)(r[i] * c[j]);
      int dy = (int)(r[i] * s[j]);
      grid[((size_t)j * ow + i) * 2] = (int16_t)dx; /* :36-38 */
      grid[((size_t)j * ow + i) * 2 + 1] = (int16_t)dy;
    }
  free(r);
  free(c);
  free(s);
}

/* ImageSampler::SampleFrameLogPolarGPU (image_sampler.cc:577-621) ->
 * sample_logpolar_kernel, image_sampler_sample_logpolar_kernel.cl:41-86. */
void orc_img_sample_logpolar(uint8_t *out, int ow, int oh, int out_linesize, const uint8_t *src,
                             int W, int H, int src_linesize, const int16_t *grid, float cx,
                             float cy) {
  const int sbpp = src_linesize / W, obpp = out_linesize / ow; /* :49-50 */
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      int x = (int)(cx * W + grid[((size_t)j * ow + i) * 2]); /* :67-70 */
      int y = (int)(cy * H + grid[((size_t)j * ow + i) * 2 + 1]);
      x = (x + 10 * W) % W; /* :73 */
      y = iclamp(y, 0, H - 1);
      if (x >= 0 && x < W && y >= 0 && y < H) { /* :76-77 */
        uint8_t *o = out + (size_t)j * out_linesize + (size_t)i * obpp;
        const uint8_t *s = src + (size_t)y * src_linesize + (size_t)x * sbpp;
        o[0] = s[0];
        o[1] = s[1];
        o[2] = s[2];
      }
    }
}

/* ImageSampler::InterpolateFrameLogPolarGPU (image_sampler.cc:780-818) ->
 * interpolate_logpolar_kernel, image_sampler_interpolate_kernel.cl:1-81.
 * Mixed float/double arithmetic follows the C++ overloads the g++ shim picks:
 * pow(int, float) and fmod(float, int) promote to double. */
void orc_img_interpolate_logpolar(uint8_t *out, int W, int H, const uint8_t *reduced, int ow,
                                  int oh, float cx, float cy) {
  const float alpha = 1.0f; /* :9 */
  const int cxp = (int)(cx * W), cyp = (int)(cy * H); /* :19-20 */
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int yy = 0; yy < H; ++yy)
    for (int xx = 0; xx < W; ++xx) {
      int x = xx, y = yy;
      uint8_t *o = out + ((size_t)yy * W + xx) * 4; /* :13 */
      if (x - cxp > W / 2) /* :21-25 */
        x -= W;
      else if (x - cxp < -W / 2)
        x += W;
      int dx = x - cxp, dy = y - cyp;
      float i_f = (dx == 0 && dy == 0)
                      ? 0.0f
                      : (float)(ow * pow(log(sqrt(pow((double)dx, (double)2.0f) +
                                                  pow((double)dy, (double)2.0f))) /
                                             10.0f,
                                         (double)(1.0f / alpha))); /* :28-33 */
      int i = iclamp((int)roundf(i_f), 0, ow - 1);                 /* :34 */
      float j_f = 0.0f;
      if (dx != 0) { /* :36-40 */
        j_f = (float)((atanf((float)dy / dx) + M_PI * (dx < 0)) * ((float)oh / (2.0 * M_PI)));
        j_f = (float)fmod((double)(j_f + 2 * oh), (double)oh);
      } else { /* :41-43 */
        j_f = (float)((M_PI_2 + M_PI * (dy < 0)) * (oh / (2.0 * M_PI)));
      }
      int j = iclamp((int)roundf(j_f), 0, oh - 1); /* :44 */
      float rad = expf(10.0f * powf((float)i / ow, alpha));
      int calc_x = (int)(cx * W + rad * cos((float)j / oh * 2.0f * M_PI)); /* :46-48 */
      int calc_y = (int)(cy * H + rad * sin((float)j / oh * 2.0f * M_PI)); /* :49-51 */
      if (calc_x == x && calc_y == y) {                                    /* :53-55 */
        memcpy(o, reduced + ((size_t)j * ow + i) * 4, 4);
        continue;
      }
      int min_i = iclamp((int)floorf(i_f), 0, ow - 1); /* :59-62 */
      int min_j = (int)floorf(j_f + oh) % oh;
      int max_i = iclamp((int)ceilf(i_f), 0, ow - 1);
      int max_j = (int)ceilf(j_f + oh) % oh;
      const uint8_t *tl = reduced + ((size_t)min_j * ow + min_i) * 4;
      const uint8_t *tr = reduced + ((size_t)min_j * ow + max_i) * 4;
      const uint8_t *bl = reduced + ((size_t)max_j * ow + min_i) * 4;
      const uint8_t *br = reduced + ((size_t)max_j * ow + max_i) * 4;
      float ir = i_f - floorf(i_f), jr = j_f - floorf(j_f); /* :69-70 */
      for (int c = 0; c < 3; ++c) {                         /* :71-79 */
        float l = mixf((float)tl[c], (float)bl[c], jr);
        float r = mixf((float)tr[c], (float)br[c], jr);
        o[c] = (uint8_t)(int)mixf(l, r, ir);
      }
      o[3] = 0;
    }
}

/* ImageSampler::ApplyLogPolarGaussianBlur (image_sampler.cc:820-857) ->
 * logpolar_gaussian_blur_kernel, image_sampler_sample_logpolar_kernel.cl:88-142.
 * Dense uchar3 (4-byte stride) on both sides; linesize is not used for addressing. */
void orc_img_logpolar_blur(uint8_t *out, int ow, int oh, int linesize, const uint8_t *src) {
  (void)linesize;
  const float P1 = 0.3377, P2 = 0.1217, P3 = 0.0439; /* :111 */
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int j = 0; j < oh; ++j)
    for (int i = 0; i < ow; ++i) {
      uint8_t *o = out + ((size_t)j * ow + i) * 4;
      if (i < ow / 2) { /* :138-139 */
        memcpy(o, src + ((size_t)j * ow + i) * 4, 4);
        continue;
      }
      int jm = imax(j - 1, 0), jp = imin(j + 1, oh - 1); /* :112-121 */
      int im = imax(i - 1, 0), ip = imin(i + 1, ow - 1);
#define PX(J, I) (src + ((size_t)(J)*ow + (I)) * 4)
      for (int c = 0; c < 3; ++c) { /* :123-137 */
        float corners = (float)PX(jm, im)[c] + (float)PX(jm, ip)[c] + (float)PX(jp, im)[c] +
                        (float)PX(jp, ip)[c];
        float edges = (float)PX(jm, i)[c] + (float)PX(j, im)[c] + (float)PX(j, ip)[c] +
                      (float)PX(jp, i)[c];
        float v = P3 * corners + P2 * edges + P1 * (float)PX(j, i)[c];
        o[c] = (uint8_t)(int)v;
      }
#undef PX
      o[3] = 0;
    }
}

/* FNV-1a 64-bit over a raw buffer: the hash the golden fixtures are keyed on. */
/* ---- Projections (section 8(f) rank 3) -------------------------------------------------- */

/* gnomonic_kernel, projections_program.cl:7-47: inverse gnomonic projection of a tw x th
 * viewport (fixed tangent-plane extent 6 x 3) centred on `center` out of a W x H equirectangular
 * frame; both buffers are dense uchar3 arrays (4-byte pixels), the whole 4-byte pixel is copied.
 * Typing follows OpenCL C: the #define'd PI / PI_2 are double literals, so phi1, lambda0, the two
 * fmod wraps and the final divisions are evaluated in double and narrowed to float on assignment;
 * everything else is float (sqrtf, atanf, sinf, cosf, asinf, atan2f). */
void orc_gnomonic(uint8_t *out, int tw, int th, const uint8_t *src, int W, int H, float cx,
                  float cy) {
  const double PI = 3.141592653589793, PI_2 = 1.5707963267948966;
  const uint32_t *s32 = (const uint32_t *)src;
  uint32_t *o32 = (uint32_t *)out;
#pragma omp parallel for schedule(static) num_threads(NT)
  for (int j = 0; j < th; ++j) {
    for (int i = 0; i < tw; ++i) {
      const float u = (float)i / tw, v = (float)j / th;  /* :21 */
      const float x = 6.0f * (u - 0.5f), y = 3.0f * (v - 0.5f); /* :19, :22-23 */
      const float phi1 = (float)((cy - 0.5) * PI);          /* :26 */
      const float lambda0 = (float)((cx - 0.5) * 2.0 * PI); /* :28 */
      const float rho = sqrtf(x * x + y * y);
      const float c = atanf(rho);
      float phi = asinf(cosf(c) * sinf(phi1) + (y * sinf(c) * cosf(phi1)) / rho); /* :31 */
      float lambda = lambda0 + atan2f(x * sinf(c), (rho * cosf(phi1) * cosf(c) -
                                                    y * sinf(phi1) * sinf(c))); /* :32-34 */
      phi = (float)fmod(phi + PI_2 + 10 * PI, 2 * PI);     /* :35 */
      lambda = (float)fmod(lambda + PI + 10 * PI, 2 * PI); /* :36 */
      float su = (float)(lambda / (2.0 * PI)), sv = (float)(phi / (PI)); /* :37 */
      su = fminf(fmaxf(su, 0.0f), 0.999f); /* :38: clamp = fmin(fmax()), a NaN becomes 0 */
      sv = fminf(fmaxf(sv, 0.0f), 0.999f);
      const int sc = (int)(sv * H) * W + (int)(su * W); /* :40-41 */
      o32[(size_t)j * tw + i] = s32[sc];                /* :43-44 */
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * VideoEncoder::EncodeFrame colour conversion (video_encoder.cc:380-398):
 * sws_getContext(w, h, RGB0, w, h, YUV420P, SWS_BILINEAR) + sws_scale.  The arithmetic is
 * libswscale's (a third-party dependency; its FFmpeg 4.2 sources are vendored under
 * /root/reference/include/FFmpeg42/libswscale, cited below as sws/<file>:<line>); restated here in
 * its C (SWS_BITEXACT) form.  Pinned by tests/golden/swscale_*.npz, generated by the REAL
 * libswscale binary present in the build container (tests/golden/make_golden_swscale.py).
 *
 *  - coefficients: ITU-R BT.601 limited range, 15-bit fixed point, (int)(k * 2^15 + 0.5)
 *    (sws/utils.c:807-816);
 *  - luma: a 14-bit value per pixel with rounding term (32 << 14) + 256 and shift 9
 *    (sws/input.c:252-275 with the bgr32 wrapper's shifts, :378), widened to 15 bits by the
 *    identity horizontal filter (x 2^14 >> 13, saturating at 2^15 - 1, sws/swscale.c:96-122) and
 *    narrowed with the constant "dither" 64: (v + 64) >> 7 (sws/output.c:395-403,
 *    sws/swscale.c:351-352);
 *  - chroma: horizontal pairs are summed BEFORE the matrix (the "_half" input readers,
 *    sws/input.c:305-345; chosen because the output is horizontally subsampled and
 *    SWS_FULL_CHR_H_INP is not set, sws/utils.c:1391-1405): rounding term (256 << 15) + 512,
 *    shift 10; widened to 15 bits like luma; vertically the bilinear 2:1 filter has the four
 *    12-bit taps 512, 1536, 1536, 512 on source rows 2j-1 .. 2j+2, rows outside the frame
 *    replicate the edge row (sws/utils.c:403-418, :494 and its border fix-up), accumulated on
 *    64 << 12 and shifted by 19 (sws/output.c:380-393).
 * Preconditions (outside them libswscale builds other filters): W and H even, H >= 8. */
enum {
  kRY = 8414, kGY = 16519, kBY = 3208,     /* (int)(0.299|0.587|0.114 * 219/255 * 2^15 + 0.5) */
  kRU = -4865, kGU = -9528, kBU = 14392,   /* -(int)(0.169..), -(int)(0.331..), (int)(0.500 * 224/255 ..) */
  kRV = 14392, kGV = -12061, kBV = -2332
};

static inline uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
static inline int widen15(int v14) { return imin(v14 * 2, 32767); }

/* 15-bit chroma of the horizontal pixel pair i of one source row */
static inline void chroma15_pair(const uint8_t *row, int i, int *u15, int *v15) {
  const uint8_t *p = row + (size_t)8 * i;
  const int r = p[0] + p[4], g = p[1] + p[5], b = p[2] + p[6];
  const int rnd = (256 << 15) + 512;
  *u15 = widen15((kRU * r + kGU * g + kBU * b + rnd) >> 10);
  *v15 = widen15((kRV * r + kGV * g + kBV * b + rnd) >> 10);
}

int orc_rgb0_to_yuv420p(uint8_t *y, int y_linesize, uint8_t *u, int u_linesize, uint8_t *v,
                        int v_linesize, const uint8_t *src, int src_linesize, int W, int H) {
  if (W < 2 || H < 8 || (W & 1) || (H & 1)) return -2;
#pragma omp parallel for num_threads(NT) schedule(static)
  for (int yy = 0; yy < H; ++yy) {
    const uint8_t *row = src + (size_t)yy * src_linesize;
    for (int x = 0; x < W; ++x) {
      const uint8_t *p = row + (size_t)4 * x;
      const int y14 = (kRY * p[0] + kGY * p[1] + kBY * p[2] + (32 << 14) + 256) >> 9;
      y[(size_t)yy * y_linesize + x] = clip_u8((widen15(y14) + 64) >> 7);
    }
  }
#pragma omp parallel for num_threads(NT) schedule(static)
  for (int j = 0; j < H / 2; ++j) {
    static const int tap[4] = {512, 1536, 1536, 512};
    for (int i = 0; i < W / 2; ++i) {
      int su = 64 << 12, sv = 64 << 12;
      for (int k = 0; k < 4; ++k) {
        const int yy = iclamp(2 * j - 1 + k, 0, H - 1);
        int cu, cv;
        chroma15_pair(src + (size_t)yy * src_linesize, i, &cu, &cv);
        su += tap[k] * cu;
        sv += tap[k] * cv;
      }
      u[(size_t)j * u_linesize + i] = clip_u8(su >> 19);
      v[(size_t)j * v_linesize + i] = clip_u8(sv >> 19);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * VideoDecoder::GetFrame colour conversion (video_decoder.cc:165-170, :222):
 * sws_getContext(w, h, YUV420P, w, h, RGB0, SWS_BILINEAR) + sws_scale.  Same size and an 8-bit
 * packed RGB target select libswscale's dedicated yuv2rgb converter (sws/swscale_unscaled.c), not
 * the scaler: chroma is NOT interpolated (each U/V sample covers its 2x2 luma block) and the
 * arithmetic is the 16-bit fixed point of the x86 converter the reference's host runs
 * (sws/x86/yuv2rgb_template.c:95-98 with the coefficients of sws/yuv2rgb.c:830-837):
 *     Y' = (((Y << 3) - 128) * 9539) >> 16          (9539 = round16(255/219 * 2^13), signed)
 *     U' = (U << 3) - 1024,  V' = (V << 3) - 1024
 *     R = clip8(Y' + ((V' * 13075) >> 16))
 *     G = clip8(Y' + ((U' * -3209) >> 16) + ((V' * -6660) >> 16))
 *     B = clip8(Y' + ((U' * 16525) >> 16))           (>> = arithmetic shift: pmulhw)
 * and the 4th byte of RGB0 is written as 255.  Pinned by tests/golden/swscale_*.npz: the real
 * libswscale returns exactly this with and without SWS_BITEXACT.  W and H even. */
int orc_yuv420p_to_rgb0(uint8_t *dst, int dst_linesize, const uint8_t *y, int y_linesize,
                        const uint8_t *u, int u_linesize, const uint8_t *v, int v_linesize, int W,
                        int H) {
  if (W < 2 || H < 2 || (W & 1) || (H & 1)) return -2;
#pragma omp parallel for num_threads(NT) schedule(static)
  for (int yy = 0; yy < H; ++yy) {
    for (int x = 0; x < W; ++x) {
      const int Y = y[(size_t)yy * y_linesize + x];
      const int U = u[(size_t)(yy / 2) * u_linesize + x / 2];
      const int V = v[(size_t)(yy / 2) * v_linesize + x / 2];
      const int yt = (((Y << 3) - 128) * 9539) >> 16;
      const int up = (U << 3) - 1024, vp = (V << 3) - 1024;
      uint8_t *o = dst + (size_t)yy * dst_linesize + (size_t)4 * x;
      o[0] = clip_u8(yt + ((vp * 13075) >> 16));
      o[1] = clip_u8(yt + ((up * -3209) >> 16) + ((vp * -6660) >> 16));
      o[2] = clip_u8(yt + ((up * 16525) >> 16));
      o[3] = 255;
    }
  }
  return 0;
}

uint64_t orc_fnv1a64(const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

/* Synthetic RGB0 frame generator shared by tests and bench:
 * LCG s = s*1664525 + 1013904223 per byte, byte = s>>24, padding byte = 0. */
void orc_fill_frame_lcg(uint8_t *buf, size_t nbytes, uint32_t seed) {
  uint32_t s = seed;
  for (size_t i = 0; i < nbytes; ++i) {
    s = s * 1664525u + 1013904223u;
    buf[i] = ((i & 3) == 3) ? 0 : (uint8_t)(s >> 24);
  }
}
