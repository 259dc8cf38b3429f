/*
 * msda_oracle.c — CPU restatement of RichSem's MSDeformAttn arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped op
 * (richsem_b200/) never imports, links or calls anything under oracle/.
 *
 * What it restates (paths relative to /root/reference):
 *   forward   models/richsem/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299  (per-output loop,
 *             l-major / p-minor accumulation into one register) with the bilinear sample of
 *             :33-84;
 *   backward  :87-159 (per-sample gradient formulas) summed over channels as the D=32 kernel
 *             :301-403 does; grad_value accumulated sequentially (the CUDA kernel's atomics
 *             make its order arbitrary, so only tolerance-level agreement is defined there);
 *   indices   :285-288 (pixel coordinate = loc*size - 0.5 as a rounded multiply followed by a
 *             rounded subtract; range test) and :38-78 (floor, per-corner bounds).  This is the
 *             bit-exact contract for corner indices and level offsets.
 *
 * Pinning: checked against golden vectors produced by the reference's own
 * ms_deform_attn_core_pytorch (models/richsem/ops/functions/ms_deform_attn_func.py:41-61) in
 * tests/golden/ (generator: tests/golden/make_golden.py) — see tests/test_oracle.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile).  -ffp-contract=off
 * matters: an FMA in the coordinate computation changes floor() on the pixel lattice.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GEOM_BODY(T, FLOOR)                                                                       \
  const T w_im = (T)(x * (T)W) - (T)0.5;                                                          \
  const T h_im = (T)(y * (T)H) - (T)0.5;                                                          \
  tok[0] = tok[1] = tok[2] = tok[3] = -1;                                                         \
  *lh = 0;                                                                                        \
  *lw = 0;                                                                                        \
  if (!(h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W)) return 0;                    \
  {                                                                                               \
    const T hf = FLOOR(h_im), wf = FLOOR(w_im);                                                   \
    const int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;                               \
    *lh = h_im - hf;                                                                              \
    *lw = w_im - wf;                                                                              \
    if (h0 >= 0 && w0 >= 0) tok[0] = (int32_t)(start + (int64_t)h0 * W + w0);                     \
    if (h0 >= 0 && w1 <= W - 1) tok[1] = (int32_t)(start + (int64_t)h0 * W + w1);                 \
    if (h1 <= H - 1 && w0 >= 0) tok[2] = (int32_t)(start + (int64_t)h1 * W + w0);                 \
    if (h1 <= H - 1 && w1 <= W - 1) tok[3] = (int32_t)(start + (int64_t)h1 * W + w1);             \
  }                                                                                               \
  return 1;

static int geom_f32(float x, float y, int H, int W, int64_t start, int32_t tok[4], float* lh, float* lw) {
  GEOM_BODY(float, floorf)
}
static int geom_f64(double x, double y, int H, int W, int64_t start, int32_t tok[4], double* lh, double* lw) {
  GEOM_BODY(double, floor)
}

/* Same geometry with the pixel coordinate computed as ONE fused multiply-add, fmaf(x, W, -0.5f):
 * what nvcc's default -fmad=true makes of cuh:285-286, i.e. the compiled reference extension.
 * (MSDA_FLAG_COORDS_FMA in the library.) */
static int geom_f32_fma(float x, float y, int H, int W, int64_t start, int32_t tok[4], float* lh, float* lw) {
  const float w_im = fmaf(x, (float)W, -0.5f), h_im = fmaf(y, (float)H, -0.5f);
  tok[0] = tok[1] = tok[2] = tok[3] = -1;
  *lh = 0;
  *lw = 0;
  if (!(h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W)) return 0;
  {
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h0 = (int)hf, w0 = (int)wf, h1 = h0 + 1, w1 = w0 + 1;
    *lh = h_im - hf;
    *lw = w_im - wf;
    if (h0 >= 0 && w0 >= 0) tok[0] = (int32_t)(start + (int64_t)h0 * W + w0);
    if (h0 >= 0 && w1 <= W - 1) tok[1] = (int32_t)(start + (int64_t)h0 * W + w1);
    if (h1 <= H - 1 && w0 >= 0) tok[2] = (int32_t)(start + (int64_t)h1 * W + w0);
    if (h1 <= H - 1 && w1 <= W - 1) tok[3] = (int32_t)(start + (int64_t)h1 * W + w1);
  }
  return 1;
}

void msda_oracle_corners_f32_fma(const int64_t* shapes, const int64_t* start, const float* loc, int64_t n_qm,
                                 int L, int P, int32_t* corners) {
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n_qm; ++t)
    for (int l = 0; l < L; ++l)
      for (int p = 0; p < P; ++p) {
        const int64_t s = (t * L + l) * P + p;
        float lh, lw;
        geom_f32_fma(loc[2 * s], loc[2 * s + 1], (int)shapes[2 * l], (int)shapes[2 * l + 1], start[l],
                     corners + 4 * s, &lh, &lw);
      }
}

/* corners[b,q,m,l,p,4] : token index of each bilinear corner, -1 if it contributes nothing */
void msda_oracle_corners_f32(const int64_t* shapes, const int64_t* start, const float* loc, int64_t n_qm,
                             int L, int P, int32_t* corners) {
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n_qm; ++t)
    for (int l = 0; l < L; ++l)
      for (int p = 0; p < P; ++p) {
        const int64_t s = (t * L + l) * P + p;
        float lh, lw;
        geom_f32(loc[2 * s], loc[2 * s + 1], (int)shapes[2 * l], (int)shapes[2 * l + 1], start[l],
                 corners + 4 * s, &lh, &lw);
      }
}

#define DEFINE_ORACLE(T, SUF, GEOM)                                                               \
  void msda_oracle_forward_##SUF(const T* value, const int64_t* shapes, const int64_t* start,     \
                                 const T* loc, const T* attw, int N, int S, int M, int D, int L,  \
                                 int Lq, int P, T* out) {                                         \
    const int64_t n_qm = (int64_t)N * Lq * M;                                                     \
    _Pragma("omp parallel for schedule(static)") for (int64_t t = 0; t < n_qm; ++t) {             \
      const int m = (int)(t % M);                                                                 \
      const int64_t b = t / ((int64_t)M * Lq);                                                    \
      const T* vb = value + b * S * M * D;                                                        \
      T* o = out + t * D;                                                                         \
      for (int c = 0; c < D; ++c) o[c] = 0;                                                       \
      for (int l = 0; l < L; ++l)                                                                 \
        for (int p = 0; p < P; ++p) {                                                             \
          const int64_t s = (t * L + l) * P + p;                                                  \
          int32_t tok[4];                                                                         \
          T lh, lw;                                                                               \
          if (!GEOM(loc[2 * s], loc[2 * s + 1], (int)shapes[2 * l], (int)shapes[2 * l + 1],       \
                    start[l], tok, &lh, &lw))                                                     \
            continue;                                                                             \
          const T hh = 1 - lh, hw = 1 - lw, a = attw[s];                                          \
          const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                         \
          for (int c = 0; c < D; ++c) {                                                           \
            const T v1 = tok[0] >= 0 ? vb[((int64_t)tok[0] * M + m) * D + c] : 0;                 \
            const T v2 = tok[1] >= 0 ? vb[((int64_t)tok[1] * M + m) * D + c] : 0;                 \
            const T v3 = tok[2] >= 0 ? vb[((int64_t)tok[2] * M + m) * D + c] : 0;                 \
            const T v4 = tok[3] >= 0 ? vb[((int64_t)tok[3] * M + m) * D + c] : 0;                 \
            o[c] += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * a;                                  \
          }                                                                                       \
        }                                                                                         \
    }                                                                                             \
  }                                                                                               \
                                                                                                  \
  /* grad_value must be zero-filled by the caller; images are processed in parallel, the         \
   * samples of one image sequentially, so the result is run-to-run identical. */                 \
  void msda_oracle_backward_##SUF(const T* grad_out, const T* value, const int64_t* shapes,       \
                                  const int64_t* start, const T* loc, const T* attw, int N,       \
                                  int S, int M, int D, int L, int Lq, int P, T* grad_value,       \
                                  T* grad_loc, T* grad_attw) {                                    \
    _Pragma("omp parallel for schedule(static)") for (int64_t bm = 0; bm < (int64_t)N * M; ++bm) {\
      const int64_t b = bm / M;                                                                   \
      const int m = (int)(bm % M);                                                                \
      const T* vb = value + b * S * M * D;                                                        \
      T* gvb = grad_value + b * S * M * D;                                                        \
      for (int64_t q = 0; q < Lq; ++q) {                                                          \
        const int64_t t = (b * Lq + q) * M + m;                                                   \
        const T* go = grad_out + t * D;                                                           \
        for (int l = 0; l < L; ++l)                                                               \
          for (int p = 0; p < P; ++p) {                                                           \
            const int64_t s = (t * L + l) * P + p;                                                \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                         \
            int32_t tok[4];                                                                       \
            T lh, lw;                                                                             \
            grad_attw[s] = 0;                                                                     \
            grad_loc[2 * s] = 0;                                                                  \
            grad_loc[2 * s + 1] = 0;                                                              \
            if (!GEOM(loc[2 * s], loc[2 * s + 1], H, W, start[l], tok, &lh, &lw)) continue;       \
            const T hh = 1 - lh, hw = 1 - lw, a = attw[s];                                        \
            const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                       \
            T ga = 0, gx = 0, gy = 0;                                                             \
            for (int c = 0; c < D; ++c) {                                                         \
              const T top = go[c], tgv = top * a;                                                 \
              T gh_w = 0, gw_w = 0, val = 0;                                                      \
              if (tok[0] >= 0) {                                                                  \
                const int64_t i = ((int64_t)tok[0] * M + m) * D + c;                              \
                gh_w -= hw * vb[i]; gw_w -= hh * vb[i]; val += w1 * vb[i]; gvb[i] += w1 * tgv;    \
              }                                                                                   \
              if (tok[1] >= 0) {                                                                  \
                const int64_t i = ((int64_t)tok[1] * M + m) * D + c;                              \
                gh_w -= lw * vb[i]; gw_w += hh * vb[i]; val += w2 * vb[i]; gvb[i] += w2 * tgv;    \
              }                                                                                   \
              if (tok[2] >= 0) {                                                                  \
                const int64_t i = ((int64_t)tok[2] * M + m) * D + c;                              \
                gh_w += hw * vb[i]; gw_w -= lh * vb[i]; val += w3 * vb[i]; gvb[i] += w3 * tgv;    \
              }                                                                                   \
              if (tok[3] >= 0) {                                                                  \
                const int64_t i = ((int64_t)tok[3] * M + m) * D + c;                              \
                gh_w += lw * vb[i]; gw_w += lh * vb[i]; val += w4 * vb[i]; gvb[i] += w4 * tgv;    \
              }                                                                                   \
              ga += top * val;                                                                    \
              gx += (T)W * gw_w * tgv;                                                            \
              gy += (T)H * gh_w * tgv;                                                            \
            }                                                                                     \
            grad_attw[s] = ga;                                                                    \
            grad_loc[2 * s] = gx;                                                                 \
            grad_loc[2 * s + 1] = gy;                                                             \
          }                                                                                       \
      }                                                                                           \
    }                                                                                             \
  }

DEFINE_ORACLE(float, f32, geom_f32)
DEFINE_ORACLE(double, f64, geom_f64)

int msda_oracle_version(void) { return 1; }
