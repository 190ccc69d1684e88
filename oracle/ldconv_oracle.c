/*
 * ldconv_oracle.c -- CPU restatement of the reference LDConv, forward and backward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 * The product (experiment_yolo_b200/) never imports, links or executes it and has no CPU fallback.
 *
 * What it restates (all citations relative to /root/reference/):
 *   ultralytics/nn/modules/conv.py:350-503   class LDConv
 * The reference holds no golden vectors or tests for this path, so this oracle is pinned
 * against outputs of the reference itself: oracle/gen_golden.py imports the unmodified reference module in the
 * authoring container and writes the .npz fixtures under tests/golden/; tests/test_oracle_golden.py checks every function below
 * against those fixtures (sampling indices / clamped coordinates / resampled operand bit-exact).
 *
 * Layout follows the reference: activations NCHW fp32, offsets (B, 2N, h, w) with the first N channels the ROW (H axis)
 * offsets and the last N the COLUMN (W axis) offsets (conv.py:463-467 calls them x / y).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the coordinate and bilinear arithmetic must round exactly
 * like the reference's separate element-wise torch kernels, so no FMA contraction is allowed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define LDC_EXPORT __attribute__((visibility("default")))

/* conv.py:413-432  _get_p_n: raster base grid, base = round(sqrt(N)), N // base full rows of `base`
 * columns, then one partial row of N % base.  out[0:N] = row coordinates, out[N:2N] = column coordinates. */
LDC_EXPORT void ldc_oracle_p_n(int N, int64_t* out)
{
    int base = (int)lrint(sqrt((double)N));
    int rows = N / base;
    int mod = N % base;
    int idx = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < base; ++c) {
            out[idx] = r;
            out[N + idx] = c;
            ++idx;
        }
    for (int c = 0; c < mod; ++c) {
        out[idx] = rows;
        out[N + idx] = c;
        ++idx;
    }
}

/* output spatial size of the 3x3 / pad 1 / stride s offset conv (conv.py:356) */
LDC_EXPORT int ldc_oracle_out_size(int H, int s) { return (H - 1) / s + 1; }

/* conv.py:356,368  offset = p_conv(x): 3x3, padding 1, stride s, C -> 2N, with bias.
 * x (B,C,H,W), w (2N,C,3,3), b (2N) -> off (B,2N,h,w).  Accumulates in double and rounds once (the summation order of
 * the reference's ATen conv is not specified, so this is compared with a tolerance, never bit-exact). */
LDC_EXPORT void ldc_oracle_offset_conv(const float* x, const float* w, const float* b, float* off,
                                       int B, int C, int H, int W, int N, int s)
{
    const int h = ldc_oracle_out_size(H, s), wo = ldc_oracle_out_size(W, s);
    const int O = 2 * N;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < B; ++bi)
        for (int o = 0; o < O; ++o)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < wo; ++j) {
                    double acc = b ? (double)b[o] : 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int ky = 0; ky < 3; ++ky) {
                            int r = i * s + ky - 1;
                            if (r < 0 || r >= H) continue;
                            for (int kx = 0; kx < 3; ++kx) {
                                int k = j * s + kx - 1;
                                if (k < 0 || k >= W) continue;
                                acc += (double)w[((o * C + c) * 3 + ky) * 3 + kx] *
                                       (double)x[(((size_t)bi * C + c) * H + r) * W + k];
                            }
                        }
                    off[(((size_t)bi * O + o) * h + i) * wo + j] = (float)acc;
                }
}

static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One sampling point: conv.py:435-454 (p = p_0 + p_n + offset), :375-387 (floor / clamp / corner indices, clamped p),
 * :390-393 (the four bilinear weights; NOT (1-frac, frac) at the border: the corner indices and p are clamped
 * independently, so both weights on an axis are 1 once p leaves [0, H-1) ). */
typedef struct {
    int r0, r1, k0, k1;      /* clamped corner rows / columns */
    float pcr, pck;          /* clamped coordinates */
    float ar0, ar1, ak0, ak1; /* per-axis weights: g_lt=ar0*ak0, g_rb=ar1*ak1, g_lb=ar0*ak1, g_rt=ar1*ak0 */
    float pr, pk;            /* unclamped coordinates (for the clamp-backward indicator) */
} ldc_point;

static inline ldc_point ldc_point_make(int i, int j, int s, int pn_r, int pn_k, float off_r, float off_k, int H, int W)
{
    ldc_point q;
    /* (p_0 + p_n) is an exact small integer in float; adding the offset rounds once */
    q.pr = (float)(i * s + pn_r) + off_r;
    q.pk = (float)(j * s + pn_k) + off_k;
    float fr = floorf(q.pr), fk = floorf(q.pk);
    float hm = (float)(H - 1), wm = (float)(W - 1);
    q.r0 = (int)clampf(fr, 0.f, hm);
    q.r1 = (int)clampf(fr + 1.f, 0.f, hm);
    q.k0 = (int)clampf(fk, 0.f, wm);
    q.k1 = (int)clampf(fk + 1.f, 0.f, wm);
    q.pcr = clampf(q.pr, 0.f, hm);
    q.pck = clampf(q.pk, 0.f, wm);
    q.ar0 = 1.f + ((float)q.r0 - q.pcr);
    q.ar1 = 1.f - ((float)q.r1 - q.pcr);
    q.ak0 = 1.f + ((float)q.k0 - q.pck);
    q.ak1 = 1.f - ((float)q.k1 - q.pck);
    return q;
}

/* Sampling geometry only (no data): conv.py:366-393.
 * off (B,2N,h,w) -> idx (B,h,w,N,4) int32 = {r0,r1,k0,k1}; coord (B,h,w,N,2) = {pcr,pck};
 * g (B,h,w,N,4) = {g_lt,g_rb,g_lb,g_rt}.  Any output pointer may be NULL. */
LDC_EXPORT void ldc_oracle_grid(const float* off, const int64_t* p_n, int32_t* idx, float* coord, float* g,
                                int B, int H, int W, int h, int w, int N, int s)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j)
                for (int n = 0; n < N; ++n) {
                    float off_r = off[(((size_t)b * 2 * N + n) * h + i) * w + j];
                    float off_k = off[(((size_t)b * 2 * N + N + n) * h + i) * w + j];
                    ldc_point q = ldc_point_make(i, j, s, (int)p_n[n], (int)p_n[N + n], off_r, off_k, H, W);
                    size_t e = (((size_t)b * h + i) * w + j) * N + n;
                    if (idx) {
                        idx[e * 4 + 0] = q.r0; idx[e * 4 + 1] = q.r1;
                        idx[e * 4 + 2] = q.k0; idx[e * 4 + 3] = q.k1;
                    }
                    if (coord) { coord[e * 2 + 0] = q.pcr; coord[e * 2 + 1] = q.pck; }
                    if (g) {
                        g[e * 4 + 0] = q.ar0 * q.ak0; g[e * 4 + 1] = q.ar1 * q.ak1;
                        g[e * 4 + 2] = q.ar0 * q.ak1; g[e * 4 + 3] = q.ar1 * q.ak0;
                    }
                }
}

/* conv.py:396-405 + :494-503: four-corner gather, bilinear sum in the reference's order (lt, rb, lb, rt) and the
 * 'b c h w n -> b c (h n) w' rearrange.  x (B,C,H,W), off (B,2N,h,w) -> x_offset (B,C,h*N,w). */
LDC_EXPORT void ldc_oracle_sample(const float* x, const float* off, const int64_t* p_n, float* x_offset,
                                  int B, int C, int H, int W, int h, int w, int N, int s)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j)
                for (int n = 0; n < N; ++n) {
                    float off_r = off[(((size_t)b * 2 * N + n) * h + i) * w + j];
                    float off_k = off[(((size_t)b * 2 * N + N + n) * h + i) * w + j];
                    ldc_point q = ldc_point_make(i, j, s, (int)p_n[n], (int)p_n[N + n], off_r, off_k, H, W);
                    float g_lt = q.ar0 * q.ak0, g_rb = q.ar1 * q.ak1, g_lb = q.ar0 * q.ak1, g_rt = q.ar1 * q.ak0;
                    for (int c = 0; c < C; ++c) {
                        const float* xc = x + ((size_t)b * C + c) * H * W;
                        float v = g_lt * xc[q.r0 * W + q.k0];
                        v = v + g_rb * xc[q.r1 * W + q.k1];
                        v = v + g_lb * xc[q.r0 * W + q.k1];
                        v = v + g_rt * xc[q.r1 * W + q.k0];
                        x_offset[(((size_t)b * C + c) * (h * N) + (i * N + n)) * w + j] = v;
                    }
                }
}

/* conv.py:355,408: Conv2d(inc, outc, (N,1), stride (N,1), no bias) on x_offset (B,C,h*N,w):
 * pre[b,o,i,j] = sum_{c,n} Wc[o,c,n,0] * x_offset[b,c,i*N+n,j].  Double accumulation, one rounding. */
LDC_EXPORT void ldc_oracle_colconv(const float* x_offset, const float* wc, float* pre,
                                   int B, int C, int h, int w, int N, int O)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < O; ++o)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j) {
                    double acc = 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int n = 0; n < N; ++n)
                            acc += (double)wc[(o * C + c) * N + n] *
                                   (double)x_offset[(((size_t)b * C + c) * (h * N) + (i * N + n)) * w + j];
                    pre[(((size_t)b * O + o) * h + i) * w + j] = (float)acc;
                }
}

/* conv.py:355 nn.BatchNorm2d + nn.SiLU on pre (B,O,h,w).
 * training != 0: normalise with the batch mean / biased variance over (b,i,j), then
 *   running <- (1-momentum)*running + momentum*(mean, UNBIASED var)   (torch.nn.BatchNorm2d semantics);
 *   save_mean / save_invstd (O each) receive the batch statistics.
 * training == 0: normalise with running_mean / running_var (left untouched).
 * eps and momentum are read from the module by the caller (1e-3 / 0.03 inside a DetectionModel,
 * ultralytics/utils/torch_utils.py:342-352). */
LDC_EXPORT void ldc_oracle_bn_silu(const float* pre, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float* save_mean, float* save_invstd, float* out,
                                   int B, int O, int h, int w, float eps, float momentum, int training)
{
    const size_t hw = (size_t)h * w;
    const double M = (double)B * (double)hw;
    for (int o = 0; o < O; ++o) {
        double mean, var;
        if (training) {
            double sum = 0.0;
            for (int b = 0; b < B; ++b) {
                const float* p = pre + ((size_t)b * O + o) * hw;
                for (size_t t = 0; t < hw; ++t) sum += p[t];
            }
            mean = sum / M;
            double ss = 0.0;
            for (int b = 0; b < B; ++b) {
                const float* p = pre + ((size_t)b * O + o) * hw;
                for (size_t t = 0; t < hw; ++t) { double d = p[t] - mean; ss += d * d; }
            }
            var = ss / M;
            if (running_mean) running_mean[o] = (float)((1.0 - momentum) * running_mean[o] + momentum * mean);
            if (running_var) running_var[o] = (float)((1.0 - momentum) * running_var[o] + momentum * (ss / (M - 1.0)));
        } else {
            mean = running_mean[o];
            var = running_var[o];
        }
        double invstd = 1.0 / sqrt(var + (double)eps);
        if (save_mean) save_mean[o] = (float)mean;
        if (save_invstd) save_invstd[o] = (float)invstd;
        for (int b = 0; b < B; ++b) {
            const float* p = pre + ((size_t)b * O + o) * hw;
            float* q = out + ((size_t)b * O + o) * hw;
            for (size_t t = 0; t < hw; ++t) {
                double z = (p[t] - mean) * invstd * gamma[o] + beta[o];
                q[t] = (float)(z / (1.0 + exp(-z)));
            }
        }
    }
}

/* Backward of SiLU(BN(pre)) (autograd of conv.py:355's Sequential tail).
 * grad_out (B,O,h,w) -> grad_pre (B,O,h,w), grad_gamma (O), grad_beta (O).
 * training: mean / invstd are the batch statistics and the gradient flows through them;
 * eval: they are the running statistics, treated as constants. */
LDC_EXPORT void ldc_oracle_bn_silu_bwd(const float* pre, const float* grad_out, const float* gamma, const float* beta,
                                       const float* mean, const float* invstd, float* grad_pre, float* grad_gamma,
                                       float* grad_beta, int B, int O, int h, int w, int training)
{
    const size_t hw = (size_t)h * w;
    const double M = (double)B * (double)hw;
    for (int o = 0; o < O; ++o) {
        double mu = mean[o], is = invstd[o], ga = gamma[o], be = beta[o];
        double sum_dz = 0.0, sum_dz_xh = 0.0;
        for (int b = 0; b < B; ++b) {
            const float* p = pre + ((size_t)b * O + o) * hw;
            const float* gy = grad_out + ((size_t)b * O + o) * hw;
            for (size_t t = 0; t < hw; ++t) {
                double xh = (p[t] - mu) * is;
                double z = ga * xh + be;
                double sg = 1.0 / (1.0 + exp(-z));
                double dz = gy[t] * sg * (1.0 + z * (1.0 - sg));
                sum_dz += dz;
                sum_dz_xh += dz * xh;
            }
        }
        if (grad_gamma) grad_gamma[o] = (float)sum_dz_xh;
        if (grad_beta) grad_beta[o] = (float)sum_dz;
        for (int b = 0; b < B; ++b) {
            const float* p = pre + ((size_t)b * O + o) * hw;
            const float* gy = grad_out + ((size_t)b * O + o) * hw;
            float* gp = grad_pre + ((size_t)b * O + o) * hw;
            for (size_t t = 0; t < hw; ++t) {
                double xh = (p[t] - mu) * is;
                double z = ga * xh + be;
                double sg = 1.0 / (1.0 + exp(-z));
                double dz = gy[t] * sg * (1.0 + z * (1.0 - sg));
                double d = training ? (dz - sum_dz / M - xh * sum_dz_xh / M) : dz;
                gp[t] = (float)(ga * is * d);
            }
        }
    }
}

/* Backward of the (N,1) column conv: grad_pre (B,O,h,w) ->
 *   grad_x_offset (B,C,h*N,w) = sum_o Wc[o,c,n] * grad_pre[b,o,i,j]
 *   grad_wc (O,C,N)           = sum_{b,i,j} grad_pre[b,o,i,j] * x_offset[b,c,i*N+n,j]         */
LDC_EXPORT void ldc_oracle_colconv_bwd(const float* grad_pre, const float* x_offset, const float* wc,
                                       float* grad_x_offset, float* grad_wc, int B, int C, int h, int w, int N, int O)
{
    if (grad_x_offset) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
            for (int c = 0; c < C; ++c)
                for (int i = 0; i < h; ++i)
                    for (int n = 0; n < N; ++n)
                        for (int j = 0; j < w; ++j) {
                            double acc = 0.0;
                            for (int o = 0; o < O; ++o)
                                acc += (double)wc[(o * C + c) * N + n] *
                                       (double)grad_pre[(((size_t)b * O + o) * h + i) * w + j];
                            grad_x_offset[(((size_t)b * C + c) * (h * N) + (i * N + n)) * w + j] = (float)acc;
                        }
    }
    if (grad_wc) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int o = 0; o < O; ++o)
            for (int c = 0; c < C; ++c)
                for (int n = 0; n < N; ++n) {
                    double acc = 0.0;
                    for (int b = 0; b < B; ++b)
                        for (int i = 0; i < h; ++i)
                            for (int j = 0; j < w; ++j)
                                acc += (double)grad_pre[(((size_t)b * O + o) * h + i) * w + j] *
                                       (double)x_offset[(((size_t)b * C + c) * (h * N) + (i * N + n)) * w + j];
                    grad_wc[(o * C + c) * N + n] = (float)acc;
                }
    }
}

/* Backward of the bilinear resampling (autograd of conv.py:386-405; closed form in SURVEY.md Appendix A).
 * grad_x_offset (B,C,h*N,w) ->
 *   grad_x (B,C,H,W) += scatter of g * corner weight           (the four GatherBackward / scatter_add_)
 *   grad_off (B,2N,h,w): row part  1[0<=pr<=H-1] * sum_c g*(-ak0*x00 + ak1*x11 - ak1*x01 + ak0*x10)
 *                        col part  1[0<=pk<=W-1] * sum_c g*(-ar0*x00 + ar1*x11 + ar0*x01 - ar1*x10)
 * (floor / indices carry no gradient: p.detach(), conv.py:376; the indicator is torch.clamp's backward, inclusive.)
 * grad_x must be zero-initialised (or hold the gradient to accumulate into).  Serial over samples: the scatter
 * has write conflicts and this is a checker, not a fast path.  Accumulates in double. */
LDC_EXPORT void ldc_oracle_sample_bwd(const float* grad_x_offset, const float* x, const float* off, const int64_t* p_n,
                                      double* grad_x, float* grad_off, int B, int C, int H, int W, int h, int w, int N,
                                      int s)
{
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j)
                for (int n = 0; n < N; ++n) {
                    size_t o_r = (((size_t)b * 2 * N + n) * h + i) * w + j;
                    size_t o_k = (((size_t)b * 2 * N + N + n) * h + i) * w + j;
                    ldc_point q = ldc_point_make(i, j, s, (int)p_n[n], (int)p_n[N + n], off[o_r], off[o_k], H, W);
                    double ar0 = q.ar0, ar1 = q.ar1, ak0 = q.ak0, ak1 = q.ak1;
                    double acc_r = 0.0, acc_k = 0.0;
                    for (int c = 0; c < C; ++c) {
                        const float* xc = x + ((size_t)b * C + c) * H * W;
                        double* gx = grad_x ? grad_x + ((size_t)b * C + c) * H * W : NULL;
                        double g = grad_x_offset[(((size_t)b * C + c) * (h * N) + (i * N + n)) * w + j];
                        double x00 = xc[q.r0 * W + q.k0], x11 = xc[q.r1 * W + q.k1];
                        double x01 = xc[q.r0 * W + q.k1], x10 = xc[q.r1 * W + q.k0];
                        if (gx) {
                            gx[q.r0 * W + q.k0] += g * ar0 * ak0;
                            gx[q.r1 * W + q.k1] += g * ar1 * ak1;
                            gx[q.r0 * W + q.k1] += g * ar0 * ak1;
                            gx[q.r1 * W + q.k0] += g * ar1 * ak0;
                        }
                        acc_r += g * (-ak0 * x00 + ak1 * x11 - ak1 * x01 + ak0 * x10);
                        acc_k += g * (-ar0 * x00 + ar1 * x11 + ar0 * x01 - ar1 * x10);
                    }
                    if (grad_off) {
                        int in_r = (q.pr >= 0.f) && (q.pr <= (float)(H - 1));
                        int in_k = (q.pk >= 0.f) && (q.pk <= (float)(W - 1));
                        grad_off[o_r] = in_r ? (float)acc_r : 0.f;
                        grad_off[o_k] = in_k ? (float)acc_k : 0.f;
                    }
                }
}

/* Backward of the offset conv (standard conv2d backward of conv.py:356):
 *   grad_x (B,C,H,W) += conv_transpose(grad_off, w, stride s, padding 1)   (second gradient path into x)
 *   grad_w (2N,C,3,3), grad_b (2N). */
LDC_EXPORT void ldc_oracle_offset_conv_bwd(const float* grad_off, const float* x, const float* w, double* grad_x,
                                           float* grad_w, float* grad_b, int B, int C, int H, int W, int N, int s)
{
    const int h = ldc_oracle_out_size(H, s), wo = ldc_oracle_out_size(W, s);
    const int O = 2 * N;
    double* gw = grad_w ? (double*)calloc((size_t)O * C * 9, sizeof(double)) : NULL;
    double* gb = grad_b ? (double*)calloc((size_t)O, sizeof(double)) : NULL;
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < O; ++o)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < wo; ++j) {
                    double g = grad_off[(((size_t)b * O + o) * h + i) * wo + j];
                    if (gb) gb[o] += g;
                    for (int c = 0; c < C; ++c)
                        for (int ky = 0; ky < 3; ++ky) {
                            int r = i * s + ky - 1;
                            if (r < 0 || r >= H) continue;
                            for (int kx = 0; kx < 3; ++kx) {
                                int k = j * s + kx - 1;
                                if (k < 0 || k >= W) continue;
                                size_t xi = (((size_t)b * C + c) * H + r) * W + k;
                                int wi = ((o * C + c) * 3 + ky) * 3 + kx;
                                if (grad_x) grad_x[xi] += g * (double)w[wi];
                                if (gw) gw[wi] += g * (double)x[xi];
                            }
                        }
                }
    if (gw) { for (int t = 0; t < O * C * 9; ++t) grad_w[t] = (float)gw[t]; free(gw); }
    if (gb) { for (int t = 0; t < O; ++t) grad_b[t] = (float)gb[t]; free(gb); }
}

LDC_EXPORT int ldc_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
