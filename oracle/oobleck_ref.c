/* oobleck_ref.c -- plain C restatement of the layer primitives of the Oobleck / sigmaVAE hot path.
 * TEST INFRASTRUCTURE ONLY (loaded by tests/test_oracle_c.py through ctypes; never by the product).
 *
 * Independent of torch: used to cross-check oracle/oobleck_oracle.py (which leans on torch's CPU conv
 * kernels, exactly as the reference does) with straight loops and double accumulation.
 * Reference lines followed (relative to /root/reference):
 *   ref_weight_norm      torch.nn.utils.weight_norm dim=0, as used by dac.nn.layers.WNConv1d /
 *                        WNConvTranspose1d (call sites stable_audio_tools/models/autoencoders.py:49,52,76,98)
 *   ref_snake_beta       stable_audio_tools/models/blocks.py:301-302,331-339
 *   ref_conv1d           nn.Conv1d semantics of autoencoders.py:49-53,76-77,133,141,168,184
 *   ref_conv_transpose1d nn.ConvTranspose1d semantics of autoencoders.py:98-100
 *   ref_residual_unit    autoencoders.py:39-62
 *   ref_vae_sample       stable_audio_tools/models/bottleneck.py:51-62
 *   ref_sigma_sample     model_sigmaVAE.py:187-213
 * Layout everywhere: [B, C, T] row-major, float32.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* w[i,:] = v[i,:] * g[i] / ||v[i,:]||_2 */
void ref_weight_norm(const float* v, const float* g, float* w, int dim0, int inner) {
  for (int i = 0; i < dim0; ++i) {
    double s = 0.0;
    for (int j = 0; j < inner; ++j) s += (double)v[(size_t)i * inner + j] * v[(size_t)i * inner + j];
    const float scale = g[i] / (float)sqrt(s);
    for (int j = 0; j < inner; ++j) w[(size_t)i * inner + j] = v[(size_t)i * inner + j] * scale;
  }
}

/* y = x + 1/(exp(beta)+1e-9) * sin(x*exp(alpha))^2, per channel */
void ref_snake_beta(const float* x, float* y, const float* alpha, const float* beta, int logscale, int B, int C,
                    long T) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float a = alpha[c], bt = beta[c];
      if (logscale) { a = expf(a); bt = expf(bt); }
      const float inv_b = 1.0f / (bt + 0.000000001f);
      const size_t base = ((size_t)b * C + c) * T;
      for (long t = 0; t < T; ++t) {
        const float s = sinf(x[base + t] * a);
        y[base + t] = x[base + t] + inv_b * (s * s);
      }
    }
}

/* w: [Cout, Cin, K]; y: [B, Cout, T_out], T_out = (T + 2p - d(K-1) - 1)/s + 1 */
void ref_conv1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, long T,
                int K, int stride, int dilation, int padding) {
  const long T_out = (T + 2L * padding - (long)dilation * (K - 1) - 1) / stride + 1;
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      float* yr = y + ((size_t)b * Cout + co) * T_out;
      for (long t = 0; t < T_out; ++t) {
        double acc = bias ? bias[co] : 0.0;
        for (int ci = 0; ci < Cin; ++ci) {
          const float* xr = x + ((size_t)b * Cin + ci) * T;
          const float* wr = w + ((size_t)co * Cin + ci) * K;
          for (int k = 0; k < K; ++k) {
            const long ti = t * stride + (long)k * dilation - padding;
            if (ti >= 0 && ti < T) acc += (double)xr[ti] * wr[k];
          }
        }
        yr[t] = (float)acc;
      }
    }
}

/* w: [Cin, Cout, K]; y: [B, Cout, T_out], T_out = (T-1)s - 2p + (K-1) + 1 */
void ref_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                          long T, int K, int stride, int padding) {
  const long T_out = (T - 1) * stride - 2L * padding + (K - 1) + 1;
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      float* yr = y + ((size_t)b * Cout + co) * T_out;
      for (long t = 0; t < T_out; ++t) {
        double acc = bias ? bias[co] : 0.0;
        for (int k = 0; k < K; ++k) {
          const long num = t + padding - k;
          if (num < 0 || num % stride) continue;
          const long ti = num / stride;
          if (ti >= T) continue;
          for (int ci = 0; ci < Cin; ++ci)
            acc += (double)x[((size_t)b * Cin + ci) * T + ti] * w[((size_t)ci * Cout + co) * K + k];
        }
        yr[t] = (float)acc;
      }
    }
}

/* x + conv1(snake(conv7_dil(snake(x)))) with already folded weights w7 [C,C,7], w1 [C,C,1] */
void ref_residual_unit(const float* x, float* y, const float* a0, const float* b0, const float* w7,
                       const float* bias7, const float* a1, const float* b1, const float* w1, const float* bias1,
                       int B, int C, long T, int dilation) {
  const size_t n = (size_t)B * C * T;
  float* h0 = (float*)malloc(n * sizeof(float));
  float* h1 = (float*)malloc(n * sizeof(float));
  ref_snake_beta(x, h0, a0, b0, 1, B, C, T);
  ref_conv1d(h0, w7, bias7, h1, B, C, C, T, 7, 1, dilation, 3 * dilation);
  ref_snake_beta(h1, h0, a1, b1, 1, B, C, T);
  ref_conv1d(h0, w1, bias1, h1, B, C, C, T, 1, 1, 1, 0);
  for (size_t i = 0; i < n; ++i) y[i] = h1[i] + x[i];
  free(h0);
  free(h1);
}

/* latents = noise*scale + mean (two roundings); returns kl = (mean^2+var-log var-1).sum(1).mean() */
double ref_vae_sample(const float* mean, const float* scale, const float* noise, float* out, int B, int D, long T) {
  double kl = 0.0;
  const size_t n = (size_t)B * D * T;
  for (size_t i = 0; i < n; ++i) {
    volatile float t = noise[i] * scale[i]; /* volatile: forbid FMA contraction */
    out[i] = t + mean[i];
    const float sp = scale[i] > 20.f ? scale[i] : log1pf(expf(scale[i]));
    const float stdev = sp + 1e-4f;
    const float var = stdev * stdev;
    kl += (double)(mean[i] * mean[i] + var - logf(var) - 1.f);
  }
  return kl / ((double)B * (double)T);
}

/* mean + std*noise (two roundings) */
void ref_sigma_sample(const float* mean, const float* noise, float* out, size_t n, float std) {
  for (size_t i = 0; i < n; ++i) {
    volatile float t = std * noise[i];
    out[i] = mean[i] + t;
  }
}
