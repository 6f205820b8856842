// HBM-bound kernels around the convolutions: BatchNorm finalize / backward, max-pooling, global
// average pooling, input packing, L2-normalise, momentum (EMA) update, queue enqueue, Adam.
// All activations are channels-last fp32 rows [pixels, ld]; channel counts are multiples of 4 and
// every kernel moves float4 per thread with consecutive threads on consecutive addresses.
#include "common.cuh"
#include "coclr_b200.h"

namespace coclr {

COCLR_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
COCLR_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

COCLR_DEVINL float4 affine_relu(float4 v, float4 sc, float4 sh, int relu) {
  v.x = fmaf(v.x, sc.x, sh.x);
  v.y = fmaf(v.y, sc.y, sh.y);
  v.z = fmaf(v.z, sc.z, sh.z);
  v.w = fmaf(v.w, sc.w, sh.w);
  if (relu) {
    v.x = fmaxf(v.x, 0.f);
    v.y = fmaxf(v.y, 0.f);
    v.z = fmaxf(v.z, 0.f);
    v.w = fmaxf(v.w, 0.f);
  }
  return v;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalize: per-channel sums -> (scale, shift) for the consumer prologue, saved
// (mean, rstd) for backward, running-stat update.  nn.BatchNorm3d train-mode semantics
// (backbone/s3dg.py:16,46-47): biased variance for normalisation, unbiased for running_var.
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const coclr_bn_finalize_t P) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  float mean, var;
  if (P.training) {
    const double n = (double)P.count;
    const double m = P.sum[c] / n;
    double v = P.sumsq[c] / n - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    const float unbiased = (float)(P.count > 1 ? v * n / (n - 1.0) : v);
    P.running_mean[c] = (1.f - P.momentum) * P.running_mean[c] + P.momentum * mean;
    P.running_var[c] = (1.f - P.momentum) * P.running_var[c] + P.momentum * unbiased;
  } else {
    mean = P.running_mean[c];
    var = P.running_var[c];
  }
  const float rstd_exact = 1.f / sqrtf(var + P.eps);
  const float sc = P.gamma[c] * rstd_exact;
  P.scale[c] = sc;
  P.shift[c] = P.beta[c] - mean * sc;
  if (P.save_mean) P.save_mean[c] = mean;
  if (P.save_rstd) P.save_rstd[c] = rstd_exact;
}

// ------------------------------------------------------------------------------------------------
// column-reduce helper: thread t < A owns channel group (t % C4) and rows (t / C4) + k*(A / C4)
// ------------------------------------------------------------------------------------------------
static constexpr int kColThreads = 256;

// BN backward, phase 1: s1[c] = sum dz, s2[c] = sum dz * xhat, dz = dA * [scale*y+shift > 0]
__global__ void __launch_bounds__(kColThreads) bn_bwd_reduce_kernel(const coclr_bn_bwd_t P) {
  const int C4 = P.C >> 2;
  const int A = (kColThreads / C4) * C4;
  const int R = A / C4;
  __shared__ float red[8][kColThreads];
  const int t = threadIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (t < A) {
    const int cg = t % C4, rs = t / C4;
    const int c = cg * 4;
    const float4 sc = ld4(P.scale + c), sh = ld4(P.shift + c), mu = ld4(P.mean + c), rs4 = ld4(P.rstd + c);
    const long rows_per = ((long)P.M + gridDim.x - 1) / gridDim.x;
    const long r_begin = (long)blockIdx.x * rows_per;
    const long r_end = min((long)P.M, r_begin + rows_per);
    for (long r = r_begin + rs; r < r_end; r += R) {
      const float4 y = ld4(P.y + r * P.ld + P.coff + c);
      const float4 da = ld4(P.dA + r * P.ld + P.coff + c);
      float4 dz;
      dz.x = (!P.relu || fmaf(y.x, sc.x, sh.x) > 0.f) ? da.x : 0.f;
      dz.y = (!P.relu || fmaf(y.y, sc.y, sh.y) > 0.f) ? da.y : 0.f;
      dz.z = (!P.relu || fmaf(y.z, sc.z, sh.z) > 0.f) ? da.z : 0.f;
      dz.w = (!P.relu || fmaf(y.w, sc.w, sh.w) > 0.f) ? da.w : 0.f;
      acc[0] += dz.x; acc[1] += dz.y; acc[2] += dz.z; acc[3] += dz.w;
      acc[4] += dz.x * ((y.x - mu.x) * rs4.x);
      acc[5] += dz.y * ((y.y - mu.y) * rs4.y);
      acc[6] += dz.z * ((y.z - mu.z) * rs4.z);
      acc[7] += dz.w * ((y.w - mu.w) * rs4.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j][t] = acc[j];
  __syncthreads();
  if (t < C4) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += (double)red[j][t + k * C4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(P.sums + t * 4 + j, s[j]);
      atomicAdd(P.sums + P.C + t * 4 + j, s[4 + j]);
    }
  }
}

// BN backward, phase 2 (in place on dA): dY = scale * (dz - s1/n - xhat * s2/n); block 0 also
// writes dgamma = s2, dbeta = s1.
__global__ void __launch_bounds__(kColThreads) bn_bwd_apply_kernel(const coclr_bn_bwd_t P) {
  const int C4 = P.C >> 2;
  const int A = (kColThreads / C4) * C4;
  const int R = A / C4;
  const int t = threadIdx.x;
  if (t >= A) return;
  const int cg = t % C4, rs = t / C4;
  const int c = cg * 4;
  const float4 sc = ld4(P.scale + c), sh = ld4(P.shift + c), mu = ld4(P.mean + c), rs4 = ld4(P.rstd + c);
  const double inv_n = 1.0 / (double)P.M;
  float m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double s1 = P.sums[c + j], s2 = P.sums[P.C + c + j];
    m1[j] = (float)(s1 * inv_n);
    m2[j] = (float)(s2 * inv_n);
    if (blockIdx.x == 0 && rs == 0) {
      if (P.dgamma) P.dgamma[c + j] += (float)s2;  // gradient buffers accumulate (zeroed once per step)
      if (P.dbeta) P.dbeta[c + j] += (float)s1;
    }
  }
  const long rows_per = ((long)P.M + gridDim.x - 1) / gridDim.x;
  const long r_begin = (long)blockIdx.x * rows_per;
  const long r_end = min((long)P.M, r_begin + rows_per);
  for (long r = r_begin + rs; r < r_end; r += R) {
    const float4 y = ld4(P.y + r * P.ld + P.coff + c);
    float* dp = P.dA + r * P.ld + P.coff + c;
    const float4 da = ld4(dp);
    float4 o;
    {
      const float dz = (!P.relu || fmaf(y.x, sc.x, sh.x) > 0.f) ? da.x : 0.f;
      o.x = sc.x * (dz - m1[0] - ((y.x - mu.x) * rs4.x) * m2[0]);
    }
    {
      const float dz = (!P.relu || fmaf(y.y, sc.y, sh.y) > 0.f) ? da.y : 0.f;
      o.y = sc.y * (dz - m1[1] - ((y.y - mu.y) * rs4.y) * m2[1]);
    }
    {
      const float dz = (!P.relu || fmaf(y.z, sc.z, sh.z) > 0.f) ? da.z : 0.f;
      o.z = sc.z * (dz - m1[2] - ((y.z - mu.z) * rs4.z) * m2[2]);
    }
    {
      const float dz = (!P.relu || fmaf(y.w, sc.w, sh.w) > 0.f) ? da.w : 0.f;
      o.w = sc.w * (dz - m1[3] - ((y.w - mu.w) * rs4.w) * m2[3]);
    }
    st4(dp, o);
  }
}

// bias + ReLU backward for the projection head (model/pretrain.py:52-53): in place dz = dA*[h+b>0],
// dbias = sum_rows dz.  One CTA, rows are few (the batch).
__global__ void bias_relu_bwd_kernel(const float* h, const float* bias, float* dA, float* dbias, int M, int C) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float s = 0.f;
    const float b = bias[c];
    for (int r = 0; r < M; ++r) {
      const float dz = (h[(size_t)r * C + c] + b > 0.f) ? dA[(size_t)r * C + c] : 0.f;
      dA[(size_t)r * C + c] = dz;
      s += dz;
    }
    dbias[c] += s;
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool3d (nn.MaxPool3d; backbone/s3dg.py:105,151,162,173,190): -inf padding, first maximum wins.
// The input affine+ReLU (pending BatchNorm of the producer) is applied on load; the output is final.
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const coclr_pool_t P) {
  const int C4 = P.C >> 2;
  const long total = (long)P.B * P.To * P.Ho * P.Wo * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    long r = i / C4;
    const int xo = (int)(r % P.Wo); r /= P.Wo;
    const int yo = (int)(r % P.Ho); r /= P.Ho;
    const int to = (int)(r % P.To);
    const int b = (int)(r / P.To);
    const int c = cg * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.scale) { sc = ld4(P.scale + c); sh = ld4(P.shift + c); }
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 bi = make_uchar4(0, 0, 0, 0);
    int tap = 0;
    for (int a = 0; a < P.g.kt; ++a) {
      const int ti = to * P.g.st - P.g.pt + a;
      for (int bb = 0; bb < P.g.kh; ++bb) {
        const int yi = yo * P.g.sh - P.g.ph + bb;
        for (int cc = 0; cc < P.g.kw; ++cc, ++tap) {
          const int xi = xo * P.g.sw - P.g.pw + cc;
          if ((unsigned)ti >= (unsigned)P.Ti || (unsigned)yi >= (unsigned)P.Hi || (unsigned)xi >= (unsigned)P.Wi) continue;
          float4 v = ld4(P.x + ((((long)b * P.Ti + ti) * P.Hi + yi) * P.Wi + xi) * P.ldx + P.x_coff + c);
          if (P.scale) v = affine_relu(v, sc, sh, P.relu);
          if (v.x > best.x) { best.x = v.x; bi.x = (unsigned char)tap; }
          if (v.y > best.y) { best.y = v.y; bi.y = (unsigned char)tap; }
          if (v.z > best.z) { best.z = v.z; bi.z = (unsigned char)tap; }
          if (v.w > best.w) { best.w = v.w; bi.w = (unsigned char)tap; }
        }
      }
    }
    const long o = ((((long)b * P.To + to) * P.Ho + yo) * P.Wo + xo);
    st4(P.y + o * P.ldy + P.y_coff + c, best);
    if (P.idx) *reinterpret_cast<uchar4*>(P.idx + o * P.C + c) = bi;
  }
}

// gather form of the backward: dX[in] (+)= sum over windows whose arg-max is `in` of dY[out]
__global__ void maxpool_bwd_kernel(const coclr_pool_t P) {
  const int C4 = P.C >> 2;
  const long total = (long)P.B * P.Ti * P.Hi * P.Wi * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    long r = i / C4;
    const int xi = (int)(r % P.Wi); r /= P.Wi;
    const int yi = (int)(r % P.Hi); r /= P.Hi;
    const int ti = (int)(r % P.Ti);
    const int b = (int)(r / P.Ti);
    const int c = cg * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int tap = 0;
    for (int a = 0; a < P.g.kt; ++a) {
      int nt = ti + P.g.pt - a;
      const bool vt = nt >= 0 && (nt % P.g.st) == 0 && (nt / P.g.st) < P.To;
      nt /= P.g.st;
      for (int bb = 0; bb < P.g.kh; ++bb) {
        int ny = yi + P.g.ph - bb;
        const bool vy = ny >= 0 && (ny % P.g.sh) == 0 && (ny / P.g.sh) < P.Ho;
        ny /= P.g.sh;
        for (int cc = 0; cc < P.g.kw; ++cc, ++tap) {
          int nx = xi + P.g.pw - cc;
          const bool vx = nx >= 0 && (nx % P.g.sw) == 0 && (nx / P.g.sw) < P.Wo;
          nx /= P.g.sw;
          if (!(vt && vy && vx)) continue;
          const long o = ((((long)b * P.To + nt) * P.Ho + ny) * P.Wo + nx);
          const uchar4 id = *reinterpret_cast<const uchar4*>(P.idx + o * P.C + c);
          if (id.x == tap || id.y == tap || id.z == tap || id.w == tap) {
            const float4 d = ld4(P.dy + o * P.ldy + P.y_coff + c);
            if (id.x == tap) acc.x += d.x;
            if (id.y == tap) acc.y += d.y;
            if (id.z == tap) acc.z += d.z;
            if (id.w == tap) acc.w += d.w;
          }
        }
      }
    }
    float* dp = P.dx + ((((long)b * P.Ti + ti) * P.Hi + yi) * P.Wi + xi) * P.ldx + P.x_coff + c;
    if (P.accumulate) {
      const float4 old = ld4(dp);
      acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
    }
    st4(dp, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// AdaptiveAvgPool3d((1,1,1)) over relu(scale*y+shift) (model/pretrain.py:51) and its backward
// ------------------------------------------------------------------------------------------------
__global__ void avgpool_fwd_kernel(const float* x, int ld, int coff, const float* scale, const float* shift, int relu,
                                   float* out, int B, int Pn, int C) {
  const int C4 = C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C4) return;
  const int cg = i % C4, b = i / C4;
  const int c = cg * 4;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) { sc = ld4(scale + c); sh = ld4(shift + c); }
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < Pn; ++p) {
    float4 v = ld4(x + ((long)b * Pn + p) * ld + coff + c);
    if (scale) v = affine_relu(v, sc, sh, relu);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float inv = 1.f / (float)Pn;
  st4(out + (long)b * C + c, make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv));
}
__global__ void avgpool_bwd_kernel(const float* dfeat, float* dA, int ld, int coff, int B, int Pn, int C) {
  const int C4 = C >> 2;
  const long total = (long)B * Pn * C4;
  const float inv = 1.f / (float)Pn;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    const long r = i / C4;
    const int b = (int)(r / Pn);
    const float4 d = ld4(dfeat + (long)b * C + cg * 4);
    st4(dA + r * ld + coff + cg * 4, make_float4(d.x * inv, d.y * inv, d.z * inv, d.w * inv));
  }
}

// ------------------------------------------------------------------------------------------------
// clip packing: x[b, c, t, h, w] (c < 3, arbitrary batch stride: the reference's block[:, i]
// .contiguous() copies, model/pretrain.py:149-150, are folded in) -> [b, thw, 4] with channel 3 = 0
// ------------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* x, long batch_stride, long chan_stride, int Cin, float* out, int B,
                                  long thw, const long* __restrict__ batch_index) {
  const long total = (long)B * thw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / thw, p = i - b * thw;
    const long sb = batch_index ? batch_index[b] : b;  // shuffle-BN gather folded in (pretrain.py:124)
    const float* s = x + sb * batch_stride + p;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = s[0];
    if (Cin > 1) v.y = s[chan_stride];
    if (Cin > 2) v.z = s[2 * chan_stride];
    if (Cin > 3) v.w = s[3 * chan_stride];
    st4(out + i * 4, v);
  }
}

// ------------------------------------------------------------------------------------------------
// F.normalize(z + bias, dim=1) (model/pretrain.py:154,167) and backward; one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* z, const float* bias, float* q, float* inv_norm, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = z[(long)row * D + c] + (bias ? bias[c] : 0.f);
    ss += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < D; c += 32) q[(long)row * D + c] = (z[(long)row * D + c] + (bias ? bias[c] : 0.f)) * inv;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}
// dz = (dq - q * <q, dq>) * inv_norm
__global__ void l2norm_bwd_kernel(const float* q, const float* dq, const float* inv_norm, float* dz, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot += q[(long)row * D + c] * dq[(long)row * D + c];
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float inv = inv_norm[row];
  for (int c = lane; c < D; c += 32)
    dz[(long)row * D + c] = (dq[(long)row * D + c] - q[(long)row * D + c] * dot) * inv;
}
__global__ void colsum_small_kernel(const float* x, float* out, int M, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < M; ++r) s += x[(long)r * C + c];
  out[c] += s;
}

// ------------------------------------------------------------------------------------------------
// momentum update of the key encoder: k = k*m + q*(1-m) (model/pretrain.py:76-80), bit-exact
// with the reference's two roundings (no FMA contraction), one launch for all parameters
// ------------------------------------------------------------------------------------------------
__global__ void ema_kernel(float* __restrict__ k, const float* __restrict__ q, float m, float one_minus_m, long n) {
  const long n4 = n >> 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(k)[i];
    const float4 b = reinterpret_cast<const float4*>(q)[i];
    a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(b.x, one_minus_m));
    a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(b.y, one_minus_m));
    a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(b.z, one_minus_m));
    a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(b.w, one_minus_m));
    reinterpret_cast<float4*>(k)[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (n4 << 2) + threadIdx.x;
    k[i] = __fadd_rn(__fmul_rn(k[i], m), __fmul_rn(q[i], one_minus_m));
  }
}

// queue[:, ptr:ptr+n] = keys^T (model/pretrain.py:82-96); consecutive threads write consecutive columns
__global__ void enqueue_kernel(float* queue, const float* keys, int dim, int K, int ptr, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dim * n) return;
  const int c = i / n, j = i - c * n;
  queue[(long)c * K + ptr + j] = keys[(long)j * dim + c];
}

// torch.optim.Adam with coupled L2 weight decay (main_nce.py:190-200) over one flat buffer
__global__ void adam_kernel(const coclr_adam_t P) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long)gridDim.x * blockDim.x) {
    const float p = P.param[i];
    float g = P.grad[i] * P.grad_scale;
    g = g + P.weight_decay * p;
    float m = P.exp_avg[i];
    float v = P.exp_avg_sq[i];
    m = m + (1.f - P.beta1) * (g - m);
    v = v * P.beta2 + (1.f - P.beta2) * g * g;
    const float denom = sqrtf(v) / P.bc2_sqrt + P.eps;
    P.param[i] = p - P.step_size * (m / denom);
    P.exp_avg[i] = m;
    P.exp_avg_sq[i] = v;
  }
}

static inline int grid_for(long total, int threads, int cap) {
  long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace coclr

using namespace coclr;
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? COCLR_OK : COCLR_E_LAUNCH)

extern "C" int coclr_bn_finalize(const coclr_bn_finalize_t* p, coclr_stream_t stream) {
  if (!p || !p->scale || !p->shift || !p->gamma || !p->beta || p->C <= 0) return COCLR_E_ARG;
  if (p->training && (!p->sum || !p->sumsq || p->count <= 0)) return COCLR_E_ARG;
  bn_finalize_kernel<<<(p->C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p);
  return LAUNCH_OK();
}

extern "C" int coclr_bn_bwd(const coclr_bn_bwd_t* p, int num_sms, coclr_stream_t stream) {
  if (!p || !p->y || !p->dA || !p->sums || p->C % 4 || p->C > 1024 || p->ld % 4 || p->coff % 4) return COCLR_E_ARG;
  const int C4 = p->C / 4;
  const int R = kColThreads / C4;
  long slabs = ((long)p->M + (long)R * 16 - 1) / ((long)R * 16);
  int grid = (int)(slabs < (long)num_sms * 8 ? slabs : (long)num_sms * 8);
  if (grid < 1) grid = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(p->sums, 0, sizeof(double) * 2 * p->C, s) != cudaSuccess) return COCLR_E_LAUNCH;
  bn_bwd_reduce_kernel<<<grid, kColThreads, 0, s>>>(*p);
  bn_bwd_apply_kernel<<<grid, kColThreads, 0, s>>>(*p);
  return LAUNCH_OK();
}

extern "C" int coclr_bias_relu_bwd(const float* h, const float* bias, float* dA, float* dbias, int M, int C,
                                   coclr_stream_t stream) {
  if (!h || !bias || !dA || !dbias) return COCLR_E_ARG;
  bias_relu_bwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h, bias, dA, dbias, M, C);
  return LAUNCH_OK();
}

extern "C" int coclr_maxpool_fwd(const coclr_pool_t* p, coclr_stream_t stream) {
  if (!p || !p->x || !p->y || p->C % 4 || p->ldx % 4 || p->ldy % 4) return COCLR_E_ARG;
  if (p->g.kt * p->g.kh * p->g.kw > 255) return COCLR_E_ARG;
  const long total = (long)p->B * p->To * p->Ho * p->Wo * (p->C / 4);
  maxpool_fwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(*p);
  return LAUNCH_OK();
}
extern "C" int coclr_maxpool_bwd(const coclr_pool_t* p, coclr_stream_t stream) {
  if (!p || !p->dx || !p->dy || !p->idx || p->C % 4) return COCLR_E_ARG;
  const long total = (long)p->B * p->Ti * p->Hi * p->Wi * (p->C / 4);
  maxpool_bwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(*p);
  return LAUNCH_OK();
}

extern "C" int coclr_avgpool_fwd(const float* x, int ld, int coff, const float* scale, const float* shift, int relu,
                                 float* out, int B, int Pn, int C, coclr_stream_t stream) {
  if (!x || !out || C % 4) return COCLR_E_ARG;
  avgpool_fwd_kernel<<<(B * (C / 4) + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ld, coff, scale, shift, relu, out,
                                                                                B, Pn, C);
  return LAUNCH_OK();
}
extern "C" int coclr_avgpool_bwd(const float* dfeat, float* dA, int ld, int coff, int B, int Pn, int C,
                                 coclr_stream_t stream) {
  if (!dfeat || !dA || C % 4) return COCLR_E_ARG;
  const long total = (long)B * Pn * (C / 4);
  avgpool_bwd_kernel<<<grid_for(total, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(dfeat, dA, ld, coff, B, Pn, C);
  return LAUNCH_OK();
}

extern "C" int coclr_pack_input(const float* x, long batch_stride, long chan_stride, int Cin, float* out, int B,
                                long thw, const long* batch_index, coclr_stream_t stream) {
  if (!x || !out || Cin < 1 || Cin > 4) return COCLR_E_ARG;
  pack_input_kernel<<<grid_for((long)B * thw, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, batch_stride,
                                                                                             chan_stride, Cin, out, B, thw, batch_index);
  return LAUNCH_OK();
}

extern "C" int coclr_l2norm_fwd(const float* z, const float* bias, float* q, float* inv_norm, int B, int D,
                                coclr_stream_t stream) {
  if (!z || !q) return COCLR_E_ARG;
  l2norm_fwd_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(z, bias, q, inv_norm, B, D);
  return LAUNCH_OK();
}
extern "C" int coclr_l2norm_bwd(const float* q, const float* dq, const float* inv_norm, float* dz, float* dbias, int B,
                                int D, coclr_stream_t stream) {
  if (!q || !dq || !inv_norm || !dz) return COCLR_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  l2norm_bwd_kernel<<<(B + 3) / 4, 128, 0, s>>>(q, dq, inv_norm, dz, B, D);
  if (dbias) colsum_small_kernel<<<(D + 127) / 128, 128, 0, s>>>(dz, dbias, B, D);
  return LAUNCH_OK();
}

extern "C" int coclr_ema_update(float* k, const float* q, float m, float one_minus_m, long n, int num_sms,
                                coclr_stream_t stream) {
  if (!k || !q || n < 0) return COCLR_E_ARG;
  if (((uintptr_t)k | (uintptr_t)q) & 15) return COCLR_E_ARG;
  ema_kernel<<<grid_for(n / 4 + 1, 256, num_sms * 8), 256, 0, (cudaStream_t)stream>>>(k, q, m, one_minus_m, n);
  return LAUNCH_OK();
}

extern "C" int coclr_queue_enqueue(float* queue, const float* keys, int dim, int K, int ptr, int n,
                                   coclr_stream_t stream) {
  if (!queue || !keys || ptr < 0 || ptr + n > K) return COCLR_E_ARG;
  enqueue_kernel<<<(dim * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(queue, keys, dim, K, ptr, n);
  return LAUNCH_OK();
}

extern "C" int coclr_adam_step(const coclr_adam_t* p, int num_sms, coclr_stream_t stream) {
  if (!p || !p->param || !p->grad || !p->exp_avg || !p->exp_avg_sq) return COCLR_E_ARG;
  adam_kernel<<<grid_for(p->n, 256, num_sms * 8), 256, 0, (cudaStream_t)stream>>>(*p);
  return LAUNCH_OK();
}
