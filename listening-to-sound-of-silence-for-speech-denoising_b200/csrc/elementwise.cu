// HBM-bound elementwise / reduction kernels of the hot path:
//   complex-ratio-mask recovery (fwd/bwd)      M2/transform.py:156-169
//   MSE / BCE-with-logits losses (fwd + grad)  M2/agent.py:172-190, M1/agent.py:185-202
//   Adam step                                  M2/agent.py:167-170 (torch.optim.Adam defaults)
//   BatchNorm (train/eval) + ReLU/PReLU fwd/bwd over NHWC activations
//   layout changes (NCHW <-> NHWC, reflect borders, nearest resize, concat slices)
#include "common.cuh"
#include <stdlib.h>
#include "sos_b200.h"

namespace {

constexpr int kThreads = 256;
inline int grid_for(long long n, int per_block = kThreads, int max_blocks = 148 * 16) {
  long long g = ceil_div_ll(n, per_block);
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// Smallest grid >= g for which g * kThreads is a multiple of `groups` (a thread then keeps one channel group for the whole run).
inline int grid_mult(int g, int groups) {
  int a = groups, b = kThreads;
  while (b) { const int t = a % b; a = b; b = t; }
  const int m = groups / a;
  return (g + m - 1) / m * m;
}

// ----------------------------------------------------------------------------- cRM
__device__ __forceinline__ float crm_to_m(float c, float inv_a, float b) {
  return inv_a * (logf(c / (1.f - c + 1e-8f) + 1e-10f) + b);
}

__global__ void icrm_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ crm, float* __restrict__ rec,
                                long long plane, long long total, float inv_a, float b) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long bi = e / plane, r = e - bi * plane;
    const long long i0 = bi * 2 * plane + r, i1 = i0 + plane;
    const float mr = crm_to_m(crm[i0], inv_a, b), mi = crm_to_m(crm[i1], inv_a, b);
    const float yr = Y[i0], yi = Y[i1];
    rec[i0] = mr * yr - mi * yi;
    rec[i1] = mr * yi + mi * yr;
  }
}

__device__ __forceinline__ float crm_dm_dc(float c, float inv_a) {
  const float den = 1.f - c + 1e-8f;
  const float u = c / den;
  return inv_a * ((1.f + 1e-8f) / (den * den)) / (u + 1e-10f);
}

__global__ void icrm_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ crm, const float* __restrict__ grec,
                                float* __restrict__ gcrm, long long plane, long long total, float inv_a) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long bi = e / plane, r = e - bi * plane;
    const long long i0 = bi * 2 * plane + r, i1 = i0 + plane;
    const float yr = Y[i0], yi = Y[i1], gr = grec[i0], gi = grec[i1];
    gcrm[i0] = (gr * yr + gi * yi) * crm_dm_dc(crm[i0], inv_a);
    gcrm[i1] = (gi * yr - gr * yi) * crm_dm_dc(crm[i1], inv_a);
  }
}

// ----------------------------------------------------------------------------- losses
__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long n, float* __restrict__ loss_sum,
                           float* __restrict__ grad, float gscale) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float d = pred[e] - tgt[e];
    acc = fmaf(d, d, acc);
    if (grad) grad[e] = d * gscale;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, acc);
}

__global__ void bce_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ loss_sum,
                           float* __restrict__ grad, float gscale) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e], t = y[e];
    acc += fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    if (grad) grad[e] = (1.f / (1.f + expf(-v)) - t) * gscale;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, acc);
}

// ----------------------------------------------------------------------------- Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_over_bc1, float b1, float b2, float eps, float inv_sqrt_bc2, float gscale) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float gr = g[e] * gscale;
    const float mm = b1 * m[e] + (1.f - b1) * gr;
    const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
    m[e] = mm;
    v[e] = vv;
    p[e] -= lr_over_bc1 * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
  }
}

__global__ void round_tf32_kernel(float* __restrict__ x, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] = tf32_rna(x[e]);
}

// Device-resident optimiser clock (CUDA-graph friendly: a replayed graph must not bake the step count or the learning rate in
// as kernel arguments).  state = [lr, step, lr / (1 - b1^step), 1 / sqrt(1 - b2^step)]; the tick kernel advances the step.
__global__ void adam_tick_kernel(float* __restrict__ state, float b1, float b2) {
  const double step = (double)state[1] + 1.0;
  state[1] = (float)step;
  state[2] = (float)((double)state[0] / (1.0 - pow((double)b1, step)));
  state[3] = (float)(1.0 / sqrt(1.0 - pow((double)b2, step)));
}
__global__ void adam_dev_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                long long n4, const float* __restrict__ state, float b1, float b2, float eps, float gscale) {
  const float lr_over_bc1 = state[2], inv_sqrt_bc2 = state[3];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = g[e];
    float4 p4 = p[e], m4 = m[e], v4 = v[e];
    const float gr[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
    float* pp = &p4.x; float* mm = &m4.x; float* vv = &v4.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mm[i] = b1 * mm[i] + (1.f - b1) * gr[i];
      vv[i] = b2 * vv[i] + (1.f - b2) * gr[i] * gr[i];
      pp[i] -= lr_over_bc1 * mm[i] / (sqrtf(vv[i]) * inv_sqrt_bc2 + eps);
    }
    p[e] = p4; m[e] = m4; v[e] = v4;
  }
}

// ----------------------------------------------------------------------------- views
struct View {
  int H, W;     // logical extent
  int Hp, Wp;   // buffer extent
  int ph, pw;   // offset of the logical window inside the buffer
  int ld;       // channels per pixel in the buffer
  int coff;     // first channel of the slice
};
__device__ __forceinline__ long long view_pix(const View& v, long long pix) {   // pix = (n*H + h)*W + w
  const int w = (int)(pix % v.W);
  const long long t = pix / v.W;
  const int h = (int)(t % v.H);
  const long long n = t / v.H;
  return ((n * v.Hp + h + v.ph) * v.Wp + (w + v.pw)) * (long long)v.ld + v.coff;
}

// ----------------------------------------------------------------------------- BatchNorm
constexpr int kUnroll = 4;          // independent 16-byte loads per array kept in flight by a thread
__device__ __forceinline__ float4 ld_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
// y dense [P][C]; each thread owns one float4 channel group and strides over rows.
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const float* __restrict__ y, long long P, int C,
                                                             float* __restrict__ partial /*[grid][2][C]*/) {
  extern __shared__ float sm[];                 // [rows][2][C]
  const int cg = C >> 2;
  const int rows = kThreads / cg;
  const int r = threadIdx.x / cg, c4 = threadIdx.x - r * cg;
  float4 s = {0, 0, 0, 0}, q = {0, 0, 0, 0};
  if (r < rows) {
    const long long stride = (long long)gridDim.x * rows;
    for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += stride * 4) {
      float4 v4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (p0 + u * stride < P) v4[u] = __ldcs(reinterpret_cast<const float4*>(y + (p0 + u * stride) * C + c4 * 4));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (p0 + u * stride >= P) break;
        const float4 v = v4[u];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
    }
    float* d = sm + (size_t)r * 2 * C;
    *reinterpret_cast<float4*>(d + c4 * 4) = s;
    *reinterpret_cast<float4*>(d + C + c4 * 4) = q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += sm[(size_t)rr * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// Reduction of the per-block partial sums: a block owns kFinCh = 8 channels (one 32-byte sector per partial row); kFinRows = 32
// thread rows stride over the G partial blocks (G is ~600: ~19 dependent loads per thread), then combine through shared
// memory.  NQ quantities per channel ([g][NQ][C] layout).  Thread t: channel t & 7, row t >> 3; the channel totals end up in
// the threads of row 0 (lanes 0..7 of warp 0).
constexpr int kFinCh = 8, kFinRows = 32;
__device__ __forceinline__ float sum8(float v) {       // over lanes 0..7 (only they call it)
  v += __shfl_xor_sync(0xffu, v, 4);
  v += __shfl_xor_sync(0xffu, v, 2);
  v += __shfl_xor_sync(0xffu, v, 1);
  return v;
}
template <int NQ>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int G, int C, int c, int gl, double (&out)[NQ],
                                                double (*sm)[NQ][kFinCh]) {
  double a[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) a[q] = 0;
  if (c < C)
    for (int g = gl; g < G; g += kFinRows) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) a[q] += partial[((size_t)g * NQ + q) * C + c];
    }
#pragma unroll
  for (int q = 0; q < NQ; ++q) sm[gl][q][threadIdx.x & (kFinCh - 1)] = a[q];
  __syncthreads();
  if (gl != 0) return;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double t = 0;
    for (int r = 0; r < kFinRows; ++r) t += sm[r][q][threadIdx.x & (kFinCh - 1)];
    out[q] = t;
  }
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int G, int C, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                          float momentum, float* __restrict__ running_mean,
                                                          float* __restrict__ running_var, float* __restrict__ mean_out,
                                                          float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                                          float* __restrict__ shift_out) {
  __shared__ double sm[kFinRows][2][kFinCh];
  const int c = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1)), gl = threadIdx.x / kFinCh;
  double r[2];
  reduce_partials<2>(partial, G, C, c, gl, r, sm);
  if (gl != 0 || c >= C) return;
  const double mean = r[0] / count;
  double var = r[1] / count - mean * mean;
  if (var < 0) var = 0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  const float sc = gamma[c] * invstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - (float)mean * sc;
  if (running_mean) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : v * slope;
  return v;
}

// z(view) = act(y * scale + shift);  y dense [P][C]
__global__ void __launch_bounds__(kThreads) bn_act_kernel(const float* __restrict__ y, float* __restrict__ z, long long P, int C,
                                                           const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                           const float* __restrict__ slope_ptr, View zv) {
  const int cg = C >> 2;
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  const bool rnd = (act & SOS_ACT_ROUND_TF32) != 0;
  act &= SOS_ACT_MASK;
  const long long total = P * cg;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / cg;
    const int c = (int)(e - p * cg) * 4;
    const float4 v = *reinterpret_cast<const float4*>(y + p * C + c);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    float4 o;
    o.x = act_fwd(fmaf(v.x, sc.x, sh.x), act, slope);
    o.y = act_fwd(fmaf(v.y, sc.y, sh.y), act, slope);
    o.z = act_fwd(fmaf(v.z, sc.z, sh.z), act, slope);
    o.w = act_fwd(fmaf(v.w, sc.w, sh.w), act, slope);
    if (rnd) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
    *reinterpret_cast<float4*>(z + view_pix(zv, p) + c) = o;
  }
}

// Backward pass 1: per-channel sums of dpre and dpre*xhat (+ PReLU slope grad).
//   pre = y*scale+shift, dpre = dz * act'(pre), xhat = (y-mean)*invstd
template <int NQ>   // 3 sums, or 4 with the sum of dpre^2 (half-precision gradient scaling)
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const float* __restrict__ dz, View dzv, const float* __restrict__ y,
                                                                  long long P, int C, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, int act,
                                                                  const float* __restrict__ slope_ptr,
                                                                  float* __restrict__ partial /*[grid][NQ][C]*/) {
  extern __shared__ float sm[];                 // [NQ][C] block totals (shared-memory atomics: a small footprint lets this
                                                // HBM-bound kernel share an SM with a tensor-core kernel of another stream)
  for (int i = threadIdx.x; i < NQ * C; i += kThreads) sm[i] = 0.f;
  __syncthreads();
  const int cg = C >> 2;
  const int rows = kThreads / cg;
  const int r = threadIdx.x / cg, c4 = threadIdx.x - r * cg;
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0}, s4[4] = {0, 0, 0, 0};
  if (r < rows) {
    const int c = c4 * 4;
    float sc[4], sh[4], mu[4], is[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { sc[i] = scale[c + i]; sh[i] = shift[c + i]; mu[i] = mean[c + i]; is[i] = invstd[c + i]; }
    const long long stride = (long long)gridDim.x * rows;
    const bool dense = dzv.Hp == dzv.H && dzv.Wp == dzv.W && dzv.ph == 0 && dzv.pw == 0 && dzv.ld == C && dzv.coff == 0;
    for (long long p0 = (long long)blockIdx.x * rows + r; p0 < P; p0 += stride * kUnroll) {
      float4 yv4[kUnroll], dz4[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {                // kUnroll independent row loads in flight per thread
        const long long p = p0 + u * stride;
        if (p < P) {
          yv4[u] = ld_stream(y + p * C + c);
          dz4[u] = ld_stream(dz + (dense ? p * C : view_pix(dzv, p)) + c);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (p0 + u * stride >= P) break;
        const float yv[4] = {yv4[u].x, yv4[u].y, yv4[u].z, yv4[u].w}, dv[4] = {dz4[u].x, dz4[u].y, dz4[u].z, dz4[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pre = fmaf(yv[i], sc[i], sh[i]);
          float dpre = dv[i];
          if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
          else if (act == 2) { if (pre <= 0.f) { s3[i] = fmaf(dv[i], pre, s3[i]); dpre *= slope; } }
          s1[i] += dpre;
          s2[i] = fmaf(dpre, (yv[i] - mu[i]) * is[i], s2[i]);
          if (NQ == 4) s4[i] = fmaf(dpre, dpre, s4[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(sm + c + i, s1[i]);
      atomicAdd(sm + C + c + i, s2[i]);
      atomicAdd(sm + 2 * C + c + i, s3[i]);
      if (NQ == 4) atomicAdd(sm + 3 * C + c + i, s4[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * C; i += kThreads) partial[(size_t)blockIdx.x * NQ * C + i] = sm[i];
}

// Backward finalize: dgamma, dbeta, dslope (atomically accumulated; zeroed by the caller) and the two per-channel means
// used by pass 2.  Same 32-channel blocks as bn_finalize_kernel.
template <int NQ>
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int G, int C, double count,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ dslope, float* __restrict__ m1, float* __restrict__ m2,
                                                              const float* __restrict__ scale, float* __restrict__ dy_sumsq, int accumulate,
                                                              int c_real) {
  __shared__ double sm[kFinRows][NQ][kFinCh];
  const int c = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1)), gl = threadIdx.x / kFinCh;
  double r[NQ];
  reduce_partials<NQ>(partial, G, C, c, gl, r, sm);
  if (gl != 0) return;
  float s3 = 0.f;
  if (c < C) {
    if (c < c_real) {                                // (channels >= c_real are zero padding of the map, not parameters)
      if (accumulate) {                              // straight into the parameters' .grad
        dbeta[c] += (float)r[0];
        dgamma[c] += (float)r[1];
      } else {
        dbeta[c] = (float)r[0];
        dgamma[c] = (float)r[1];
      }
    }
    m1[c] = (float)(r[0] / count);
    m2[c] = (float)(r[1] / count);
    s3 = (float)r[2];
  }
  if (NQ == 4) {
    // sum over pixels of dy^2 with dy = scale (dpre - m1 - xhat m2):  scale^2 (sum dpre^2 - P m1^2 - P m2^2)
    float q = 0.f;
    if (c < C) {
      const double a = r[0] / count, b = r[1] / count, sc = (double)scale[c];
      const double v = sc * sc * (r[NQ - 1] - count * a * a - count * b * b);
      q = v > 0 ? (float)v : 0.f;
    }
    q = sum8(q);
    if (threadIdx.x == 0) atomicAdd(dy_sumsq, q);
  }
  if (dslope) {
    s3 = sum8(s3);
    if (threadIdx.x == 0) atomicAdd(dslope, s3);
  }
}

// Backward pass 2: dy = scale * (dpre - m1 - xhat*m2)      (scale = gamma*invstd)
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const float* __restrict__ dz, View dzv, const float* __restrict__ y,
                                                                 float* __restrict__ dy, long long P, int C,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const float* __restrict__ m1, const float* __restrict__ m2, int act,
                                                                 const float* __restrict__ slope_ptr) {
  const int cg = C >> 2;
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  const bool rnd = (act & SOS_ACT_ROUND_TF32) != 0;
  act &= SOS_ACT_MASK;
  const long long total = P * cg;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / cg;
    const int c = (int)(e - p * cg) * 4;
    const float4 yv4 = *reinterpret_cast<const float4*>(y + p * C + c);
    const float4 dz4 = *reinterpret_cast<const float4*>(dz + view_pix(dzv, p) + c);
    const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w}, dv[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float sc = scale[c + i];
      const float pre = fmaf(yv[i], sc, shift[c + i]);
      float dpre = dv[i];
      if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
      else if (act == 2) dpre = pre > 0.f ? dpre : dpre * slope;
      const float xhat = (yv[i] - mean[c + i]) * invstd[c + i];
      o[i] = sc * (dpre - m1[c + i] - xhat * m2[c + i]);
      if (rnd) o[i] = tf32_rna(o[i]);
    }
    *reinterpret_cast<float4*>(dy + p * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}


// ----------------------------------------------------------------------------- dense fast paths
// When every operand is a dense [P][C] array, float4 element e lives at 4*e in all of them: no pixel arithmetic, and each thread
// keeps kUnroll independent 16-byte loads per array in flight (the grid-stride float4 loops above leave ~32 KB per SM in flight,
// which caps them near 4.5 TB/s; HBM needs ~45 KB per SM).

__global__ void __launch_bounds__(kThreads) bn_act_dense_kernel(const float* __restrict__ y, float* __restrict__ z, unsigned total, unsigned cg,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                                 const float* __restrict__ slope_ptr) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  const bool rnd = (act & SOS_ACT_ROUND_TF32) != 0;
  act &= SOS_ACT_MASK;
  const unsigned span = gridDim.x * blockDim.x;
  for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += span * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) v[i] = ld_stream(y + (size_t)e * 4);
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      const unsigned c = (e % cg) * 4;
      const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
      float4 o;
      o.x = act_fwd(fmaf(v[i].x, sc.x, sh.x), act, slope);
      o.y = act_fwd(fmaf(v[i].y, sc.y, sh.y), act, slope);
      o.z = act_fwd(fmaf(v[i].z, sc.z, sh.z), act, slope);
      o.w = act_fwd(fmaf(v[i].w, sc.w, sh.w), act, slope);
      if (rnd) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
      *reinterpret_cast<float4*>(z + (size_t)e * 4) = o;
    }
  }
}

// mode 0: BatchNorm backward pass 2 (needs mean / invstd / m1 / m2); mode 1: eval-mode affine backward
template <int MODE>
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_dense_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                                       float* __restrict__ dy, unsigned total, unsigned cg,
                                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                       const float* __restrict__ m1, const float* __restrict__ m2, int act,
                                                                       const float* __restrict__ slope_ptr) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  const bool rnd = (act & SOS_ACT_ROUND_TF32) != 0;
  act &= SOS_ACT_MASK;
  const unsigned span = gridDim.x * blockDim.x;
  for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += span * kUnroll) {
    float4 yv4[kUnroll], dz4[kUnroll];
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) {
        yv4[i] = ld_stream(y + (size_t)e * 4);
        dz4[i] = ld_stream(dz + (size_t)e * 4);
      }
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      const unsigned c = (e % cg) * 4;
      const float yv[4] = {yv4[i].x, yv4[i].y, yv4[i].z, yv4[i].w}, dv[4] = {dz4[i].x, dz4[i].y, dz4[i].z, dz4[i].w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float sc = scale[c + k];
        const float pre = fmaf(yv[k], sc, shift[c + k]);
        float dpre = dv[k];
        if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
        else if (act == 2) dpre = pre > 0.f ? dpre : dpre * slope;
        if (MODE == 0) {
          const float xhat = (yv[k] - mean[c + k]) * invstd[c + k];
          o[k] = sc * (dpre - m1[c + k] - xhat * m2[c + k]);
        } else {
          o[k] = sc * dpre;
        }
        if (rnd) o[k] = tf32_rna(o[k]);
      }
      *reinterpret_cast<float4*>(dy + (size_t)e * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ----------------------------------------------------------------------------- half-precision operand producers
// z half dense [P][C] = act(y * scale + shift): a thread owns 8 channels (two 16-byte loads, one 16-byte store).  The grid is
// sized so that gridDim * blockDim is a multiple of C/8: a thread then meets the same 8 channels in every iteration and keeps
// their coefficients in registers.
__global__ void __launch_bounds__(kThreads) bn_act_half_kernel(const float* __restrict__ y, uint4* __restrict__ z, unsigned total, unsigned cg8,
                                                                const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                                const float* __restrict__ slope_ptr) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  const unsigned span = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c = (tid % cg8) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); }
  for (unsigned e0 = tid; e0 < total; e0 += span * 4) {
    float4 v[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) {
        v[i][0] = ld_stream(y + (size_t)e * 8);
        v[i][1] = ld_stream(y + (size_t)e * 8 + 4);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      const float in[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = act_fwd(fmaf(in[k], sc[k], sh[k]), act, slope);
      z[e] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
    }
  }
}

// BatchNorm backward pass 2 with a scaled half output: dy = half(s * scale (dpre - m1 - xhat m2)), s = 2^e from the tensor's
// sum of squares (scal[2], written by the finalize kernel).  Block 0 publishes scal[0] = s, scal[1] = 1/s.  Same thread ->
// channel-group mapping as bn_act_half_kernel.
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_half_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                                      uint4* __restrict__ dy, unsigned total, unsigned cg8,
                                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                      const float* __restrict__ m1, const float* __restrict__ m2, int act,
                                                                      const float* __restrict__ slope_ptr, float* __restrict__ scal, float count) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  const int ex = half_scale_exp(scal[2], count);
  const float s = pow2i(ex);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[0] = s;
    scal[1] = pow2i(-ex);
  }
  const unsigned span = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c = (tid % cg8) * 8;
  float sc[8], sh[8], mu[8], is[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); mu[k] = __ldg(mean + c + k); is[k] = __ldg(invstd + c + k);
    a1[k] = __ldg(m1 + c + k); a2[k] = __ldg(m2 + c + k);
  }
  for (unsigned e0 = tid; e0 < total; e0 += span * 2) {
    float4 yv4[2][2], dz4[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) {
        yv4[i][0] = ld_stream(y + (size_t)e * 8);
        yv4[i][1] = ld_stream(y + (size_t)e * 8 + 4);
        dz4[i][0] = ld_stream(dz + (size_t)e * 8);
        dz4[i][1] = ld_stream(dz + (size_t)e * 8 + 4);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      const float yv[8] = {yv4[i][0].x, yv4[i][0].y, yv4[i][0].z, yv4[i][0].w, yv4[i][1].x, yv4[i][1].y, yv4[i][1].z, yv4[i][1].w};
      const float dv[8] = {dz4[i][0].x, dz4[i][0].y, dz4[i][0].z, dz4[i][0].w, dz4[i][1].x, dz4[i][1].y, dz4[i][1].z, dz4[i][1].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float pre = fmaf(yv[k], sc[k], sh[k]);
        float dpre = dv[k];
        if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
        else if (act == 2) dpre = pre > 0.f ? dpre : dpre * slope;
        const float xhat = (yv[k] - mu[k]) * is[k];
        o[k] = s * (sc[k] * (dpre - a1[k] - xhat * a2[k]));
      }
      dy[e] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
    }
  }
}

// ----------------------------------------------------------------------------- half-STORAGE BatchNorm passes
// Second generation of the three kernels above: the raw conv output y and the incoming gradient dz may themselves be stored as
// half (y: the conv epilogue rounds its fp32 accumulators once, AFTER taking the batch statistics from them; dz: the data-gradient
// epilogue of the next layer stores its accumulators unscaled, i.e. still multiplied by that layer's power-of-two operand scale,
// whose inverse arrives as the device scalar `in_inv`).  Per element a layer then moves 4 B forward (was 6) and 10 B backward
// (was 18).  BatchNorm backward is linear in dz, so the kernels work on the stored (scaled) values throughout: m1, m2 and the
// new operand scale are in those units, the parameter gradients are multiplied by in_inv when they are written, and the
// published scale of dy composes both (scal[0] = s * s_in, scal[1] = 1/s * in_inv).
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p, size_t e) { a = ld_stream(p + e * 8); b = ld_stream(p + e * 8 + 4); }
  __device__ __forceinline__ void get(float (&v)[8]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
};
template <> struct Raw8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* p, size_t e) { u = __ldcs(reinterpret_cast<const uint4*>(p) + e); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      v[2 * k] = f.x; v[2 * k + 1] = f.y;
    }
  }
};

template <typename TY>
__global__ void __launch_bounds__(kThreads) bn_act_h_kernel(const TY* __restrict__ y, uint4* __restrict__ z, unsigned total, unsigned cg8,
                                                            const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                            const float* __restrict__ slope_ptr) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  const unsigned span = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c = (tid % cg8) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); }
  constexpr int U = 8;
  for (unsigned e0 = tid; e0 < total; e0 += span * U) {
    Raw8<TY> v[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) v[i].load(y, e);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      float in[8], o[8];
      v[i].get(in);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = act_fwd(fmaf(in[k], sc[k], sh[k]), act, slope);
      z[e] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
    }
  }
}

// Backward pass 1 (sums of dpre, dpre * xhat, PReLU slope term, dpre^2).  Block-local mapping: thread t owns channel group
// t % cg8 for rows (t / cg8) + k * rows_per_block (threads beyond rows_per_block * cg8 idle: 4 of 256 for 48 / 96 channels).
// (two blocks per SM: <= 128 registers, so that ~64 KB of loads are in flight per SM; PRELU = false drops the slope sums)
template <typename TY, typename TDZ, bool PRELU>
__global__ void __launch_bounds__(kThreads, 2) bn_bwd_reduce_h_kernel(const TDZ* __restrict__ dz, const TY* __restrict__ y, long long P, int C,
                                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                                      const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                                                                      const float* __restrict__ slope_ptr, float* __restrict__ partial /*[grid][4][C]*/,
                                                                      float* __restrict__ zero_me) {
  extern __shared__ float sm[];                 // [rows per block][4][C]: every thread's sums, added up per (quantity, channel) below
  if (zero_me && blockIdx.x == 0 && threadIdx.x == 0) *zero_me = 0.f;      // (the finalize kernel accumulates sum dy^2 there: no fill launch)
  const int cg8 = C >> 3;
  const int rpb = kThreads / cg8;
  const int r = threadIdx.x / cg8, cgi = threadIdx.x - r * cg8;
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  if (r < rpb) {
    const int c = cgi * 8;
    float sc[8], sh[8], mu[8], is[8];
    float s1[8], s2[8], s3[8], s4[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); mu[k] = __ldg(mean + c + k); is[k] = __ldg(invstd + c + k);
      s1[k] = s2[k] = s3[k] = s4[k] = 0.f;
    }
    const long long stride = (long long)gridDim.x * rpb;
    constexpr int U = (sizeof(TY) + sizeof(TDZ) > 4) ? 2 : 4;       // 16-byte loads in flight per array and thread
    for (long long p0 = (long long)blockIdx.x * rpb + r; p0 < P; p0 += stride * U) {
      Raw8<TY> yr[U];
      Raw8<TDZ> dr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long pp = p0 + u * stride;
        if (pp < P) {
          yr[u].load(y, (size_t)pp * cg8 + cgi);
          dr[u].load(dz, (size_t)pp * cg8 + cgi);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p0 + u * stride >= P) break;
        float yv[8], dv[8];
        yr[u].get(yv);
        dr[u].get(dv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float pre = fmaf(yv[k], sc[k], sh[k]);
          float dpre = dv[k];
          if (PRELU) { if (pre <= 0.f) { s3[k] = fmaf(dv[k], pre, s3[k]); dpre *= slope; } }
          else if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
          s1[k] += dpre;
          s2[k] = fmaf(dpre, (yv[k] - mu[k]) * is[k], s2[k]);
          s4[k] = fmaf(dpre, dpre, s4[k]);
        }
      }
    }
    // (plain stores + a column sum below: shared-memory atomics of 21-42 threads per address serialised the end of every block)
    float* mine = sm + (size_t)r * 4 * C + c;
#pragma unroll
    for (int k = 0; k < 8; k += 4) {
      *reinterpret_cast<float4*>(mine + k) = make_float4(s1[k], s1[k + 1], s1[k + 2], s1[k + 3]);
      *reinterpret_cast<float4*>(mine + C + k) = make_float4(s2[k], s2[k + 1], s2[k + 2], s2[k + 3]);
      *reinterpret_cast<float4*>(mine + 2 * C + k) = make_float4(s3[k], s3[k + 1], s3[k + 2], s3[k + 3]);
      *reinterpret_cast<float4*>(mine + 3 * C + k) = make_float4(s4[k], s4[k + 1], s4[k + 2], s4[k + 3]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += kThreads) {
    float t = 0.f;
    for (int rr = 0; rr < rpb; ++rr) t += sm[(size_t)rr * 4 * C + i];
    partial[(size_t)blockIdx.x * 4 * C + i] = t;
  }
}

// Backward finalize for the kernels of this section: like bn_bwd_finalize_kernel<4>, with the incoming gradient's inverse scale.
__global__ void __launch_bounds__(256) bn_bwd_finalize_h_kernel(const float* __restrict__ partial, int G, int C, double count,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dslope,
                                                                float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ scale,
                                                                float* __restrict__ dy_sumsq, int accumulate, int c_real,
                                                                const float* __restrict__ in_inv_ptr) {
  __shared__ double sm[kFinRows][4][kFinCh];
  const int c = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1)), gl = threadIdx.x / kFinCh;
  double r[4];
  reduce_partials<4>(partial, G, C, c, gl, r, sm);
  if (gl != 0) return;
  const float in_inv = in_inv_ptr ? *in_inv_ptr : 1.f;
  float s3 = 0.f, q = 0.f;
  if (c < C) {
    if (c < c_real) {
      if (accumulate) {
        dbeta[c] += (float)r[0] * in_inv;
        dgamma[c] += (float)r[1] * in_inv;
      } else {
        dbeta[c] = (float)r[0] * in_inv;
        dgamma[c] = (float)r[1] * in_inv;
      }
    }
    m1[c] = (float)(r[0] / count);
    m2[c] = (float)(r[1] / count);
    s3 = (float)r[2] * in_inv;
    const double a = r[0] / count, b = r[1] / count, sc = (double)scale[c];
    const double v = sc * sc * (r[3] - count * a * a - count * b * b);
    q = v > 0 ? (float)v : 0.f;
  }
  q = sum8(q);
  if (threadIdx.x == 0) atomicAdd(dy_sumsq, q);
  if (dslope) {
    s3 = sum8(s3);
    if (threadIdx.x == 0) atomicAdd(dslope, s3);
  }
}

template <typename TY, typename TDZ>
__global__ void __launch_bounds__(kThreads, 2) bn_bwd_apply_h_kernel(const TDZ* __restrict__ dz, const TY* __restrict__ y, uint4* __restrict__ dy,
                                                                  unsigned total, unsigned cg8, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ m1,
                                                                  const float* __restrict__ m2, int act, const float* __restrict__ slope_ptr,
                                                                  float* __restrict__ scal, float count, const float* __restrict__ in_inv_ptr) {
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  act &= SOS_ACT_MASK;
  const int ex = half_scale_exp(scal[2], count);
  const float s = pow2i(ex);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float in_inv = in_inv_ptr ? *in_inv_ptr : 1.f;
    scal[0] = s / in_inv;                              // (powers of two: exact)
    scal[1] = pow2i(-ex) * in_inv;
  }
  const unsigned span = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c = (tid % cg8) * 8;
  float sc[8], sh[8], mu[8], is[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = __ldg(scale + c + k); sh[k] = __ldg(shift + c + k); mu[k] = __ldg(mean + c + k); is[k] = __ldg(invstd + c + k);
    a1[k] = __ldg(m1 + c + k); a2[k] = __ldg(m2 + c + k);
  }
  constexpr int U = (sizeof(TY) + sizeof(TDZ) > 4) ? 2 : 4;
  for (unsigned e0 = tid; e0 < total; e0 += span * U) {
    Raw8<TY> yr[U];
    Raw8<TDZ> dr[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) {
        yr[i].load(y, e);
        dr[i].load(dz, e);
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const unsigned e = e0 + i * span;
      if (e >= total) break;
      float yv[8], dv[8], o[8];
      yr[i].get(yv);
      dr[i].get(dv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float pre = fmaf(yv[k], sc[k], sh[k]);
        float dpre = dv[k];
        if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
        else if (act == 2) dpre = pre > 0.f ? dpre : dpre * slope;
        const float xhat = (yv[k] - mu[k]) * is[k];
        o[k] = s * (sc[k] * (dpre - a1[k] - xhat * a2[k]));
      }
      dy[e] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
    }
  }
}

__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) acc = fmaf(x[e], x[e], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// out half [P][cd] = s * x [P][cs] (zero for channels >= cs); one thread per 8 output channels.
__global__ void __launch_bounds__(kThreads) to_half_kernel(const float* __restrict__ x, long long P, int cs, uint4* __restrict__ out, int cd,
                                                            float* __restrict__ scal, float count) {
  float s = 1.f;
  if (scal) {
    const int ex = half_scale_exp(scal[2], count);
    s = pow2i(ex);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      scal[0] = s;
      scal[1] = pow2i(-ex);
    }
  }
  const int g8 = cd >> 3;
  const long long total = P * g8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / g8;
    const int c = (int)(e - p * g8) * 8;
    float v[8];
    if (c + 8 <= cs && (cs & 3) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(x + p * cs + c), b = *reinterpret_cast<const float4*>(x + p * cs + c + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c + k < cs ? x[p * cs + c + k] : 0.f;
    }
    out[e] = make_uint4(pack_half2(s * v[0], s * v[1]), pack_half2(s * v[2], s * v[3]), pack_half2(s * v[4], s * v[5]), pack_half2(s * v[6], s * v[7]));
  }
}

// Eval-mode backward of z = act(y*scale + shift): dy = dz * act'(pre) * scale.
__global__ void __launch_bounds__(kThreads) affine_act_bwd_kernel(const float* __restrict__ dz, View dzv, const float* __restrict__ y,
                                                                   float* __restrict__ dy, long long P, int C,
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   int act, const float* __restrict__ slope_ptr) {
  const int cg = C >> 2;
  const float slope = slope_ptr ? *slope_ptr : 0.f;
  const bool rnd = (act & SOS_ACT_ROUND_TF32) != 0;
  act &= SOS_ACT_MASK;
  const long long total = P * cg;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / cg;
    const int c = (int)(e - p * cg) * 4;
    const float4 yv4 = *reinterpret_cast<const float4*>(y + p * C + c);
    const float4 dz4 = *reinterpret_cast<const float4*>(dz + view_pix(dzv, p) + c);
    const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w}, dv[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float sc = scale[c + i];
      const float pre = fmaf(yv[i], sc, shift[c + i]);
      float dpre = dv[i];
      if (act == 1) dpre = pre > 0.f ? dpre : 0.f;
      else if (act == 2) dpre = pre > 0.f ? dpre : dpre * slope;
      o[i] = sc * dpre;
      if (rnd) o[i] = tf32_rna(o[i]);
    }
    *reinterpret_cast<float4*>(dy + p * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Tiled transpose: out (cols, rows) = in (rows, cols)^T.
__global__ void transpose_kernel(const float* __restrict__ in, long long rows, long long cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
  }
}

// Split operands of the fp32-grade tensor-core GEMMs (LSTM input projections, MLP heads: M1/networks.py:95-98, M2/networks.py:64-70).
// A value x is written as hi = tf32(x) and lo = tf32(x - hi) (both exactly representable, so the TF32 products are exact) and
// a @ b = hi@hi + lo@hi + hi@lo runs as ONE 3-tap launch of the tap GEMM: the activation operand is the stack [hi; lo] (tap t reads
// stack row {0, 1, 0}[t]) and the weight operand carries the slots [hi | hi | lo] along K.
//   out[s * slot_stride + r * ld_out + k] = part_s(src[r * sr + (k + k_shift) * sk]),  r < R, k < KP  (zero where k >= K or the
//   shifted source index leaves [0, K)); n_slots = 1: {hi} (a plain TF32 operand); 2: parts {hi, lo}; 3: {hi, hi, lo}.
// One of sr / sk is 1: the 32 x 32 tile is read along that axis and written along k, i.e. the kernel also transposes.
__global__ void split_tf32_kernel(const float* __restrict__ src, long long R, long long K, long long KP, long long sr, long long sk,
                                  long long k_shift, float* __restrict__ out, long long ld_out, long long slot_stride, int n_slots,
                                  long long src_batch_stride, long long out_batch_stride) {
  __shared__ float tile[32][33];
  const long long k0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  src += blockIdx.z * src_batch_stride;                // (batched: grid.z independent (R, K) operands, e.g. one per clip)
  out += blockIdx.z * out_batch_stride;
  if (sk == 1) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const long long r = r0 + i, k = k0 + threadIdx.x, ks = k + k_shift;
      tile[i][threadIdx.x] = (r < R && k < K && ks >= 0 && ks < K) ? src[r * sr + ks] : 0.f;
    }
  } else {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const long long k = k0 + i, r = r0 + threadIdx.x, ks = k + k_shift;
      tile[threadIdx.x][i] = (r < R && k < K && ks >= 0 && ks < K) ? src[r * sr + ks * sk] : 0.f;
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i, k = k0 + threadIdx.x;
    if (r >= R || k >= KP) continue;
    const float v = tile[i][threadIdx.x];
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    float* o = out + r * ld_out + k;
    o[0] = hi;
    if (n_slots == 2) {
      o[slot_stride] = lo;
    } else if (n_slots == 3) {
      o[slot_stride] = hi;
      o[2 * slot_stride] = lo;
    }
  }
}

// dst[i] += alpha * src[i]  (parameter gradients added straight into the flat gradient buffer)
__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, float alpha, const float* __restrict__ base) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dst[e] = fmaf(alpha, src[e], base ? base[e] : dst[e]);
}

// Sequence-major rows (T, B, C) <-> per-clip maps (B, C, T): the mask head's `h.permute(0, 2, 1).view(B, 2, 256, T)` (M2/networks.py:92-93)
// and its adjoint.  to_map = 1: out (B, C, T) = in (T, B, ld)[..., :C]; to_map = 0: out (T, B, C) = in (B, C, T).  Tiled over (t, c).
__global__ void seq_map_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int B, int C, int ld, int to_map) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (to_map) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int t = t0 + i, c = c0 + threadIdx.x;
      tile[i][threadIdx.x] = (t < T && c < C) ? in[((long long)t * B + b) * ld + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + i, t = t0 + threadIdx.x;
      if (c < C && t < T) out[((long long)b * C + c) * T + t] = tile[threadIdx.x][i];
    }
  } else {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + i, t = t0 + threadIdx.x;
      tile[i][threadIdx.x] = (t < T && c < C) ? in[((long long)b * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int t = t0 + i, c = c0 + threadIdx.x;
      if (c < C && t < T) out[((long long)t * B + b) * ld + c] = tile[threadIdx.x][i];
    }
  }
}

// y[r][c] = act(y[r][c] + bias[c]);  act 0 none, 1 relu, 3 sigmoid
__global__ void bias_act_kernel(float* __restrict__ y, long long rows, int cols, long long ld, const float* __restrict__ bias, int act) {
  const long long total = rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cols;
    const int c = (int)(e - r * cols);
    float v = y[r * ld + c] + (bias ? bias[c] : 0.f);
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 3) v = 1.f / (1.f + expf(-v));
    y[r * ld + c] = v;
  }
}

// dpre = dy * act'(y_out);  dbias[c] += sum_r dpre[r][c].  One block column-strip of 32 columns x 256 rows per iteration.
__global__ void bias_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dpre, long long rows,
                                    int cols, long long ld, int act, float* __restrict__ dbias) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols) {
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      float g = dy[r * ld + c];
      const float o = y[r * ld + c];
      if (act == 1) g = o > 0.f ? g : 0.f;
      else if (act == 3) g = g * o * (1.f - o);
      if (dpre) dpre[r * ld + c] = g;
      acc += g;
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols && dbias) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(dbias + c, s);
  }
}

// ----------------------------------------------------------------------------- layout
// NCHW (B,C,H,W) -> view of NHWC buffer; channels >= C inside the slice width Cw are zero filled.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, float* __restrict__ out, View ov, int Cw, long long P) {
  const long long total = P * Cw;
  const long long plane = (long long)ov.H * ov.W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / Cw;
    const int c = (int)(e - p * Cw);
    const long long n = p / plane, r = p - n * plane;
    out[view_pix(ov, p) + c] = c < C ? x[(n * C + c) * plane + r] : 0.f;
  }
}

// NCHW fp32 (B,C,H,W) -> dense NHWC half (B,H,W,Cw), channels >= C zero filled: a thread writes 8 channels of one pixel.
__global__ void nchw_to_nhwc_half_kernel(const float* __restrict__ x, int C, uint4* __restrict__ out, int Cw, long long plane, long long P) {
  const int g8 = Cw >> 3;
  const long long total = P * g8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / g8;
    const int c0 = (int)(e - p * g8) * 8;
    const long long n = p / plane, r = p - n * plane;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = c0 + k < C ? __ldg(x + (n * C + c0 + k) * plane + r) : 0.f;
    out[e] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, View iv, float* __restrict__ out, int C, long long P) {
  const long long plane = (long long)iv.H * iv.W;
  const long long total = P * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / (plane * C);
    const long long rem = e - n * plane * C;
    const int c = (int)(rem / plane);
    const long long r = rem - c * plane;
    out[e] = in[view_pix(iv, n * plane + r) + c];
  }
}

// dst(view, H x W) = src(view, Hs x Ws) with PyTorch 'nearest' index map; C channels copied.
// accumulate: dst += src (used for gradient fan-in)
// One block per destination row (n, h): the row-level index arithmetic (nearest source row, view offsets) is done once per
// block; threads run over the row's W * C/4 float4 elements.
__global__ void __launch_bounds__(kThreads) copy_view_kernel(const float* __restrict__ src, View sv, float* __restrict__ dst, View dv, int C,
                                                              int accumulate) {
  const int cg = C >> 2;
  const int h = blockIdx.x, n = blockIdx.y;
  int hs = h;
  if (sv.H != dv.H) hs = min((int)floorf(h * ((float)sv.H / (float)dv.H)), sv.H - 1);
  const float sw = (float)sv.W / (float)dv.W;
  const bool same_w = sv.W == dv.W;
  const float* srow = src + (((long long)n * sv.Hp + hs + sv.ph) * sv.Wp + sv.pw) * (long long)sv.ld + sv.coff;
  float* drow = dst + (((long long)n * dv.Hp + h + dv.ph) * dv.Wp + dv.pw) * (long long)dv.ld + dv.coff;
  const int total = dv.W * cg;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    const int w = e / cg;
    const int c = (e - w * cg) * 4;
    const int ws = same_w ? w : min((int)floorf(w * sw), sv.W - 1);
    float4 v = *reinterpret_cast<const float4*>(srow + (long long)ws * sv.ld + c);
    float* d = drow + (long long)w * dv.ld + c;
    if (accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(d);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(d) = v;
  }
}

// Adjoint of copy_view with resize: gsrc(view) += gdst(view) through the nearest map.  One block per destination row.  When the
// destination is not larger than the source in either axis the map is one-to-one (plain read-modify-write); otherwise several
// destination pixels hit the same source pixel and the adds are atomic.
__global__ void __launch_bounds__(kThreads) copy_view_bwd_kernel(const float* __restrict__ gdst, View dv, float* __restrict__ gsrc, View sv,
                                                                  int C) {
  const int cg = C >> 2;
  const int h = blockIdx.x, n = blockIdx.y;
  int hs = h;
  if (sv.H != dv.H) hs = min((int)floorf(h * ((float)sv.H / (float)dv.H)), sv.H - 1);
  const float sw = (float)sv.W / (float)dv.W;
  const bool same_w = sv.W == dv.W;
  const bool many_to_one = dv.H > sv.H || dv.W > sv.W;
  float* srow = gsrc + (((long long)n * sv.Hp + hs + sv.ph) * sv.Wp + sv.pw) * (long long)sv.ld + sv.coff;
  const float* drow = gdst + (((long long)n * dv.Hp + h + dv.ph) * dv.Wp + dv.pw) * (long long)dv.ld + dv.coff;
  const int total = dv.W * cg;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    const int w = e / cg;
    const int c = (e - w * cg) * 4;
    const int ws = same_w ? w : min((int)floorf(w * sw), sv.W - 1);
    const float4 v = *reinterpret_cast<const float4*>(drow + (long long)w * dv.ld + c);
    float* d = srow + (long long)ws * sv.ld + c;
    if (many_to_one) {
      atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
    } else {
      float4 o = *reinterpret_cast<float4*>(d);
      o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      *reinterpret_cast<float4*>(d) = o;
    }
  }
}

__device__ __forceinline__ int reflect_coord(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Fill the border of a padded NHWC buffer (B,Hp,Wp,C) by reflection of its interior (H x W at offset p).
// One block per buffer row (n, hp): border rows copy a whole mirrored row, interior rows only their 2*pad border columns.
__global__ void __launch_bounds__(kThreads) reflect_fill_kernel(float* __restrict__ buf, int H, int W, int pad, int C) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, cg = C >> 2;
  const int hp = blockIdx.x, n = blockIdx.y;
  const int h = hp - pad;
  const bool border_row = h < 0 || h >= H;
  const int hs = reflect_coord(h, H) + pad;
  float* drow = buf + ((long long)n * Hp + hp) * Wp * (long long)C;
  const float* srow = buf + ((long long)n * Hp + hs) * Wp * (long long)C;
  const int ncols = border_row ? Wp : 2 * pad;               // columns of this row that need filling
  const int total = ncols * cg;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    int col = e / cg;
    const int c = (e - col * cg) * 4;
    const int wp = border_row ? col : (col < pad ? col : W + col);          // interior rows: left pad, then right pad
    const int ws = reflect_coord(wp - pad, W) + pad;
    *reinterpret_cast<float4*>(drow + (long long)wp * C + c) = *reinterpret_cast<const float4*>(srow + (long long)ws * C + c);
  }
}

// Adjoint: fold border gradients back into the interior (in place).  An interior element gathers the (up to 8) border positions
// that mirror onto it, so no atomics are needed; only the elements within `pad` of an edge have any (the others keep their value and
// are not touched).  One block per interior row (n, h).
__global__ void __launch_bounds__(kThreads) reflect_fold_kernel(float* __restrict__ g, int H, int W, int pad, int C) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, cg = C >> 2;
  const int h = blockIdx.x, n = blockIdx.y;
  // rows that map to h: h itself, -h (if 1<=h<=pad), 2(H-1)-h (if H-1-pad <= h <= H-2)
  int hs[3], nh = 0;
  hs[nh++] = h;
  if (h >= 1 && h <= pad) hs[nh++] = -h;
  if (h <= H - 2 && h >= H - 1 - pad) hs[nh++] = 2 * (H - 1) - h;
  float* base = g + (long long)n * Hp * Wp * (long long)C;
  const bool whole_row = nh > 1 || 2 * pad + 2 >= W;
  const int ncols = whole_row ? W : 2 * pad;                   // else only columns 1..pad and W-1-pad..W-2 have mirror images
  const int total = ncols * cg;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    const int col = e / cg;
    const int c = (e - col * cg) * 4;
    const int w = whole_row ? col : (col < pad ? 1 + col : W - 1 - 2 * pad + col);
    int ws[3], nw = 0;
    ws[nw++] = w;
    if (w >= 1 && w <= pad) ws[nw++] = -w;
    if (w <= W - 2 && w >= W - 1 - pad) ws[nw++] = 2 * (W - 1) - w;
    float4 acc = {0, 0, 0, 0};
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nw; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(base + ((long long)(hs[i] + pad) * Wp + ws[j] + pad) * C + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    *reinterpret_cast<float4*>(base + ((long long)(h + pad) * Wp + w + pad) * C + c) = acc;
  }
}

// Fold + un-pad + channel slice in one pass: dst(view, H x W) = the interior of the reflect-padded gradient map src(view: H x W at
// offset (pad, pad) of Hp x Wp) plus the border positions that mirror onto each pixel.  Nothing is modified in place, so the
// padded map needs no copy.  One block per row (n, h).
__global__ void __launch_bounds__(kThreads) copy_view_fold_kernel(const float* __restrict__ src, View sv, float* __restrict__ dst, View dv,
                                                                   int C) {
  const int cg = C >> 2, H = sv.H, W = sv.W, pad = sv.ph;
  const int h = blockIdx.x, n = blockIdx.y;
  int hs[3], nh = 0;
  hs[nh++] = h;
  if (h >= 1 && h <= pad) hs[nh++] = -h;
  if (h <= H - 2 && h >= H - 1 - pad) hs[nh++] = 2 * (H - 1) - h;
  const float* sbase = src + (long long)n * sv.Hp * sv.Wp * (long long)sv.ld + sv.coff;
  float* drow = dst + (((long long)n * dv.Hp + h + dv.ph) * dv.Wp + dv.pw) * (long long)dv.ld + dv.coff;
  const int total = W * cg;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    const int w = e / cg;
    const int c = (e - w * cg) * 4;
    int ws[3], nw = 0;
    ws[nw++] = w;
    if (w >= 1 && w <= pad) ws[nw++] = -w;
    if (w <= W - 2 && w >= W - 1 - pad) ws[nw++] = 2 * (W - 1) - w;
    float4 acc = {0, 0, 0, 0};
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nw; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(sbase + ((long long)(hs[i] + pad) * sv.Wp + ws[j] + pad) * sv.ld + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    *reinterpret_cast<float4*>(drow + (long long)w * dv.ld + c) = acc;
  }
}

// Encoder output NHWC (B,F,T,C) -> LSTM sequence (V,B,ld) at column offset, feature index c*F+f,
// with the reference's nearest resample T -> V  (M1/networks.py:131-133; src = floor(i*T/V)).
__global__ void feat_to_seq_kernel(const float* __restrict__ in, int B, int F, int T, int C, float* __restrict__ out, int V, int ld,
                                   int coff) {
  const long long total = (long long)V * B * C * F;
  const float scale = (float)T / (float)V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(e % F);
    long long t = e / F;
    const int c = (int)(t % C); t /= C;
    const int b = (int)(t % B);
    const int v = (int)(t / B);
    const int ts = (V == T) ? v : min((int)floorf(v * scale), T - 1);
    out[((long long)v * B + b) * ld + coff + c * F + f] = in[(((long long)b * F + f) * T + ts) * C + c];
  }
}

__global__ void feat_to_seq_bwd_kernel(const float* __restrict__ gout, int B, int F, int T, int C, float* __restrict__ gin, int V, int ld,
                                       int coff) {
  const long long total = (long long)V * B * C * F;
  const float scale = (float)T / (float)V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(e % F);
    long long t = e / F;
    const int c = (int)(t % C); t /= C;
    const int b = (int)(t % B);
    const int v = (int)(t / B);
    const int ts = (V == T) ? v : min((int)floorf(v * scale), T - 1);
    atomicAdd(gin + (((long long)b * F + f) * T + ts) * C + c, gout[((long long)v * B + b) * ld + coff + c * F + f]);
  }
}

// ----------------------------------------------------------------------------- weight packing
// PyTorch conv weight (Cout,Cin,kh,kw) -> GEMM B operand [Cout][ntaps*CinP], k = tap*CinP + ci (zero for ci >= Cin)
//   mode 0: forward            tap = a*kw + b
//   mode 1: data gradient      rows = Cin, cols = tap'*CoutP + co with tap' the flipped tap (a'=kh-1-a, b'=kw-1-b)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, int CinP, int CoutP, int mode,
                                        float* __restrict__ out) {
  const int ntaps = kh * kw;
  const long long total = mode == 0 ? (long long)Cout * ntaps * CinP : (long long)Cin * ntaps * CoutP;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      const int ci = (int)(e % CinP);
      const long long t = e / CinP;
      const int tap = (int)(t % ntaps);
      const int co = (int)(t / ntaps);
      out[e] = ci < Cin ? w[((long long)co * Cin + ci) * ntaps + tap] : 0.f;
    } else {
      const int co = (int)(e % CoutP);
      const long long t = e / CoutP;
      const int tapf = (int)(t % ntaps);
      const int ci = (int)(t / ntaps);
      const int tap = ntaps - 1 - tapf;
      out[e] = co < Cout ? w[((long long)co * Cin + ci) * ntaps + tap] : 0.f;
    }
  }
}

// Generic strided gather of selected taps into a GEMM operand (one launch per convolution call):
//   out[r][t*KP + k] = w[r*sr + k*sk + tap_off[t]]   (k < K; zero for K <= k < KP), optionally rounded to TF32
struct TapOffsets { int off[49]; };
__global__ void pack_taps_kernel(const float* __restrict__ w, int R, int K, int KP, long long sr, long long sk, int ntaps, TapOffsets taps,
                                 int round, float* __restrict__ out) {
  const long long total = (long long)R * ntaps * KP;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % KP);
    const long long t2 = e / KP;
    const int t = (int)(t2 % ntaps);
    const int r = (int)(t2 / ntaps);
    float v = k < K ? w[r * sr + k * sk + taps.off[t]] : 0.f;
    if (round) v = tf32_rna(v);
    out[e] = v;
  }
}

__global__ void pack_taps_half_kernel(const float* __restrict__ w, int R, int K, int KP, long long sr, long long sk, int ntaps, TapOffsets taps,
                                      __half* __restrict__ out) {
  const long long total = (long long)R * ntaps * KP;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % KP);
    const long long t2 = e / KP;
    const int t = (int)(t2 % ntaps);
    const int r = (int)(t2 / ntaps);
    const int off = taps.off[t];                                  // < 0: a zero tap (pads a folded operand row, see sos_im2col_half)
    out[e] = __float2half_rn(k < K && off >= 0 ? fminf(fmaxf(w[r * sr + k * sk + off], -65504.f), 65504.f) : 0.f);
  }
}

// Taps folded into channels for inputs with TWO real channels (the spectrogram inputs of all three networks): out[n, oh, ow, 2 t + c]
// = x[n, oh + dh_t, ow + dw_t, c] (zero outside the image and for 2 t >= 2 ntaps).  A thread writes 8 halves = 4 taps.
struct FoldTaps { int16_t dh[32], dw[32]; };
__global__ void im2col_half_kernel(const __half* __restrict__ x, int H, int W, int Cp, int OH, int OW, int ntaps, FoldTaps taps,
                                   uint4* __restrict__ out, int g8, long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(e % g8);
    long long pix = e / g8;
    const int ow = (int)(pix % OW); pix /= OW;
    const int oh = (int)(pix % OH);
    const long long n = pix / OH;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = 4 * q + i;
      v[i] = 0u;
      if (t < ntaps) {
        const int ih = oh + taps.dh[t], iw = ow + taps.dw[t];
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v[i] = __ldg(reinterpret_cast<const uint32_t*>(x + ((n * H + ih) * W + iw) * (long long)Cp));
      }
    }
    out[e] = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

// All weight packs of a training step in ONE launch: block (i, j) does chunk j of descriptor i (include/sos_b200.h: sos_pack_desc).
__global__ void pack_taps_half_multi_kernel(const sos_pack_desc* __restrict__ descs) {
  const sos_pack_desc& d = descs[blockIdx.x];
  const float* __restrict__ w = reinterpret_cast<const float*>(d.w);
  __half* __restrict__ out = reinterpret_cast<__half*>(d.out_half);
  const unsigned KP = (unsigned)d.KP, K = (unsigned)d.K, ntaps = (unsigned)d.ntaps;
  const unsigned total = (unsigned)(d.rows * ntaps * KP);          // (a weight tensor: far below 2^32 elements)
  const unsigned span = gridDim.y * blockDim.x;
  // the gathers are latency bound (one scattered 4-byte read per element): four independent ones in flight per thread
  for (unsigned e0 = blockIdx.y * blockDim.x + threadIdx.x; e0 < total; e0 += 4 * span) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned e = e0 + i * span;
      v[i] = 0.f;
      if (e < total) {
        const unsigned k = e % KP, t2 = e / KP, t = t2 % ntaps, r = t2 / ntaps;
        const int off = d.tap_off[t];
        if (k < K && off >= 0) v[i] = __ldg(w + (long long)r * d.row_stride + (long long)k * d.k_stride + off);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned e = e0 + i * span;
      if (e < total) out[e] = __float2half_rn(fminf(fmaxf(v[i], -65504.f), 65504.f));
    }
  }
}

// wgrad result [taps][CoutP?]... -> PyTorch layout.  src is [tap][Cout][CinP] (tap-major), dst (Cout,Cin,kh,kw).
__global__ void unpack_wgrad_kernel(const float* __restrict__ src, int Cout, int Cin, int ntaps, int CinP, float* __restrict__ dst,
                                    int accumulate) {
  const long long total = (long long)Cout * Cin * ntaps;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(e % ntaps);
    const long long t = e / ntaps;
    const int ci = (int)(t % Cin);
    const int co = (int)(t / Cin);
    const float v = src[((long long)tap * Cout + co) * CinP + ci];
    dst[e] = accumulate ? dst[e] + v : v;
  }
}

// wgrad buffer [tap][RP][CP] (padded rows / columns) -> += parameter gradient (R, Cc, taps) in PyTorch's layout.
// CLEAR: the visited source elements are zeroed after they are read (the padded ones are never written by the weight-gradient
// kernel), which leaves the buffer ready for the next weight gradient without a fill launch.
template <bool CLEAR>
__global__ void accumulate_wgrad_kernel(float* __restrict__ src, int ntaps, int RP, int CP, int R, int Cc, float* __restrict__ dst) {
  const long long total = (long long)R * Cc * ntaps;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(e % ntaps);
    const long long t = e / ntaps;
    const int c = (int)(t % Cc);
    const int r = (int)(t / Cc);
    const long long si = ((long long)tap * RP + r) * CP + c;
    dst[e] += src[si];
    if (CLEAR) src[si] = 0.f;
  }
}

}  // namespace

// ============================================================================= C ABI
template <typename TY, typename TDZ>
static int bn_backward_h(const void* dz, const void* y, void* dy_half, long long rows, int C, const float* scale, const float* shift,
                         const float* mean, const float* invstd, int act, const float* slope, float* partial, float* dgamma, float* dbeta,
                         float* dslope, float* m1, float* m2, float* scal, int accumulate, int c_real, const float* in_inv, int G,
                         cudaStream_t stream, bool reduce_done = false) {
  const size_t smem = (size_t)(kThreads / (C / 8)) * 4 * C * sizeof(float);      // [rows per block][4][C] <= 32 KB
  if (reduce_done) {
    // (pass 1 came out of the epilogue of the data-gradient GEMM that produced dz: sos_conv_args::bnr_partial)
  } else if ((act & SOS_ACT_MASK) == 2)
    bn_bwd_reduce_h_kernel<TY, TDZ, true><<<G, kThreads, smem, stream>>>(reinterpret_cast<const TDZ*>(dz), reinterpret_cast<const TY*>(y), rows, C, scale,
                                                                         shift, mean, invstd, act, slope, partial, scal + 2);
  else
    bn_bwd_reduce_h_kernel<TY, TDZ, false><<<G, kThreads, smem, stream>>>(reinterpret_cast<const TDZ*>(dz), reinterpret_cast<const TY*>(y), rows, C, scale,
                                                                          shift, mean, invstd, act, slope, partial, scal + 2);
  SOS_CHECK_LAUNCH("sos_bn_act_backward_half(reduce)");
  bn_bwd_finalize_h_kernel<<<ceil_div(C, kFinCh), kFinCh * kFinRows, 0, stream>>>(partial, G, C, (double)rows, dgamma, dbeta,
                                                                                (act & SOS_ACT_MASK) == 2 ? dslope : nullptr, m1, m2, scale, scal + 2,
                                                                                accumulate, c_real, in_inv);
  SOS_CHECK_LAUNCH("sos_bn_act_backward_half(finalize)");
  const long long tot8 = rows * (C / 8);
  bn_bwd_apply_h_kernel<TY, TDZ><<<grid_mult(grid_for(tot8, kThreads * 2, 148 * 8), C / 8), kThreads, 0, stream>>>(
      reinterpret_cast<const TDZ*>(dz), reinterpret_cast<const TY*>(y), reinterpret_cast<uint4*>(dy_half), (unsigned)tot8, (unsigned)(C / 8), scale, shift,
      mean, invstd, m1, m2, act, slope, scal, (float)((double)rows * (double)C), in_inv);
  SOS_CHECK_LAUNCH("sos_bn_act_backward_half(apply)");
  return SOS_OK;
}

extern "C" {

int sos_icrm_forward(const float* Y, const float* crm, float* rec, int64_t batch, int64_t plane, float a, float b,
                     cudaStream_t stream) {
  SOS_CHECK_ARG(Y && crm && rec && batch > 0 && plane > 0 && a != 0.f, "sos_icrm_forward: bad arguments");
  const long long total = batch * plane;
  icrm_fwd_kernel<<<grid_for(total), kThreads, 0, stream>>>(Y, crm, rec, plane, total, 1.f / a, b);
  SOS_CHECK_LAUNCH("sos_icrm_forward");
  return SOS_OK;
}

int sos_icrm_backward(const float* Y, const float* crm, const float* grad_rec, float* grad_crm, int64_t batch, int64_t plane,
                      float a, cudaStream_t stream) {
  SOS_CHECK_ARG(Y && crm && grad_rec && grad_crm && batch > 0 && plane > 0 && a != 0.f, "sos_icrm_backward: bad arguments");
  const long long total = batch * plane;
  icrm_bwd_kernel<<<grid_for(total), kThreads, 0, stream>>>(Y, crm, grad_rec, grad_crm, plane, total, 1.f / a);
  SOS_CHECK_LAUNCH("sos_icrm_backward");
  return SOS_OK;
}

int sos_mse_fwd_bwd(const float* pred, const float* target, int64_t n, float* loss_sum, float* grad_or_null, float grad_scale,
                    cudaStream_t stream) {
  SOS_CHECK_ARG(pred && target && n > 0, "sos_mse_fwd_bwd: bad arguments");
  mse_kernel<<<grid_for(n, kThreads, 148 * 4), kThreads, 0, stream>>>(pred, target, n, loss_sum, grad_or_null, grad_scale);
  SOS_CHECK_LAUNCH("sos_mse_fwd_bwd");
  return SOS_OK;
}

int sos_bce_logits_fwd_bwd(const float* logits, const float* labels, int64_t n, float* loss_sum, float* grad_or_null,
                           float grad_scale, cudaStream_t stream) {
  SOS_CHECK_ARG(logits && labels && n > 0, "sos_bce_logits_fwd_bwd: bad arguments");
  bce_kernel<<<grid_for(n, kThreads, 148), kThreads, 0, stream>>>(logits, labels, n, loss_sum, grad_or_null, grad_scale);
  SOS_CHECK_LAUNCH("sos_bce_logits_fwd_bwd");
  return SOS_OK;
}

int sos_round_tf32(float* x, int64_t n, cudaStream_t stream) {
  SOS_CHECK_ARG(x && n > 0, "sos_round_tf32: bad arguments");
  round_tf32_kernel<<<grid_for(n), kThreads, 0, stream>>>(x, n);
  SOS_CHECK_LAUNCH("sos_round_tf32");
  return SOS_OK;
}

int sos_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                  float eps, int64_t step, float grad_scale, cudaStream_t stream) {
  SOS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "sos_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n), kThreads, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), beta1, beta2, eps,
                                                    (float)(1.0 / sqrt(bc2)), grad_scale);
  SOS_CHECK_LAUNCH("sos_adam_step");
  return SOS_OK;
}

int sos_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state, float beta1,
                      float beta2, float eps, float grad_scale, cudaStream_t stream) {
  SOS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && state && n > 0 && n % 4 == 0, "sos_adam_step_dev: bad arguments (n must be a multiple of 4)");
  SOS_CHECK_ARG(((uintptr_t)param % 16) == 0 && ((uintptr_t)grad % 16) == 0 && ((uintptr_t)exp_avg % 16) == 0 && ((uintptr_t)exp_avg_sq % 16) == 0,
                "sos_adam_step_dev: buffers must be 16-byte aligned");
  adam_tick_kernel<<<1, 1, 0, stream>>>(state, beta1, beta2);
  adam_dev_kernel<<<grid_for(n / 4), kThreads, 0, stream>>>(reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
                                                            reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), n / 4, state,
                                                            beta1, beta2, eps, grad_scale);
  SOS_CHECK_LAUNCH("sos_adam_step_dev");
  return SOS_OK;
}

// views are passed as 8 ints: H, W, Hp, Wp, ph, pw, ld, coff
static View mk_view(const int32_t* v) { return View{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]}; }
static bool view_dense(const View& v, int C) { return v.Hp == v.H && v.Wp == v.W && v.ph == 0 && v.pw == 0 && v.ld == C && v.coff == 0; }
static bool view_ok(const View& v, int C) {
  return v.H > 0 && v.W > 0 && v.Hp >= v.H + v.ph && v.Wp >= v.W + v.pw && v.ph >= 0 && v.pw >= 0 && v.ld >= v.coff + C &&
         (v.ld % 4) == 0 && (v.coff % 4) == 0;
}

int sos_bn_partial_blocks(int64_t rows, int64_t channels) {
  if (channels < 4 || channels % 4 || channels > 1024) return 0;
  const int rpb = kThreads / (int)(channels / 4);
  long long g = ceil_div_ll(rows, (long long)rpb * 8);
  // ONE wave of the reduction kernels (2 resident blocks per SM): measured at 32 x 256 x 203 x 48 halves, 592 / 296 / 148 blocks
  // take 58.9 / 54.3 / 76.8 us (scripts/bench_bn.py; SOS_BN_G overrides)
  static const int cap = getenv("SOS_BN_G") ? atoi(getenv("SOS_BN_G")) : 2 * sos_num_sms();
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int sos_bn_stats(const float* y, int64_t rows, int64_t channels, float* partial, cudaStream_t stream) {
  const int G = sos_bn_partial_blocks(rows, channels);
  SOS_CHECK_ARG(y && partial && rows > 0 && G > 0, "sos_bn_stats: bad arguments (channels must be a multiple of 4, <= 1024)");
  const int C = (int)channels, rpb = kThreads / (C / 4);
  const size_t smem = (size_t)rpb * 2 * C * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(bn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bn_stats_kernel<<<G, kThreads, smem, stream>>>(y, rows, C, partial);
  SOS_CHECK_LAUNCH("sos_bn_stats");
  return SOS_OK;
}

int sos_bn_finalize(const float* partial, int64_t rows, int64_t channels, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale, float* shift,
                    cudaStream_t stream) {
  const int G = sos_bn_partial_blocks(rows, channels);
  SOS_CHECK_ARG(partial && gamma && beta && mean && invstd && scale && shift && G > 0, "sos_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div((int)channels, kFinCh), kFinCh * kFinRows, 0, stream>>>(partial, G, (int)channels, (double)rows, gamma, beta, eps, momentum,
                                                                      running_mean, running_var, mean, invstd, scale, shift);
  SOS_CHECK_LAUNCH("sos_bn_finalize");
  return SOS_OK;
}

int sos_bn_finalize_partial(const float* partial, int64_t g_rows, int64_t rows, int64_t channels, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                            float* shift, cudaStream_t stream) {
  SOS_CHECK_ARG(partial && gamma && beta && mean && invstd && scale && shift && g_rows > 0 && rows > 0 && channels > 0,
                "sos_bn_finalize_partial: bad arguments");
  bn_finalize_kernel<<<ceil_div((int)channels, kFinCh), kFinCh * kFinRows, 0, stream>>>(partial, (int)g_rows, (int)channels, (double)rows, gamma, beta, eps, momentum,
                                                                     running_mean, running_var, mean, invstd, scale, shift);
  SOS_CHECK_LAUNCH("sos_bn_finalize_partial");
  return SOS_OK;
}

int sos_bn_eval_coeffs(int64_t channels, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, float* scale, float* shift, cudaStream_t stream) {
  SOS_CHECK_ARG(channels > 0 && gamma && beta && running_mean && running_var && scale && shift, "sos_bn_eval_coeffs: bad arguments");
  bn_eval_coeffs_kernel<<<ceil_div((int)channels, 128), 128, 0, stream>>>((int)channels, gamma, beta, running_mean, running_var, eps,
                                                                         scale, shift);
  SOS_CHECK_LAUNCH("sos_bn_eval_coeffs");
  return SOS_OK;
}

int sos_bn_act(const float* y, float* z, const int32_t* z_view, int64_t rows, int64_t channels, const float* scale,
               const float* shift, int act, const float* slope, cudaStream_t stream) {
  SOS_CHECK_ARG(y && z && z_view && scale && shift && rows > 0 && channels >= 4 && channels % 4 == 0, "sos_bn_act: bad arguments");
  const View zv = mk_view(z_view);
  SOS_CHECK_ARG(view_ok(zv, (int)channels) && rows % ((long long)zv.H * zv.W) == 0, "sos_bn_act: inconsistent view");
  SOS_CHECK_ARG((act & SOS_ACT_MASK) != 2 || slope, "sos_bn_act: PReLU needs a slope pointer");
  const long long tot4 = rows * (channels / 4);
  if (view_dense(zv, (int)channels) && tot4 < (1ll << 32))
    bn_act_dense_kernel<<<grid_for(tot4, kThreads * kUnroll, 148 * 8), kThreads, 0, stream>>>(y, z, (unsigned)tot4, (unsigned)(channels / 4), scale,
                                                                                              shift, act, slope);
  else
    bn_act_kernel<<<grid_for(tot4), kThreads, 0, stream>>>(y, z, rows, (int)channels, scale, shift, act, slope, zv);
  SOS_CHECK_LAUNCH("sos_bn_act");
  return SOS_OK;
}

int sos_bn_act_backward(const float* dz, const int32_t* dz_view, const float* y, float* dy, int64_t rows, int64_t channels,
                        const float* scale, const float* shift, const float* mean, const float* invstd, int act, const float* slope,
                        float* partial, float* dgamma, float* dbeta, float* dslope, float* m1, float* m2, cudaStream_t stream) {
  const int G = sos_bn_partial_blocks(rows, channels);
  SOS_CHECK_ARG(dz && dz_view && y && dy && scale && shift && mean && invstd && partial && dgamma && dbeta && m1 && m2 && G > 0,
                "sos_bn_act_backward: bad arguments");
  const View dv = mk_view(dz_view);
  const int C = (int)channels;
  SOS_CHECK_ARG(view_ok(dv, C) && rows % ((long long)dv.H * dv.W) == 0, "sos_bn_act_backward: inconsistent view");
  SOS_CHECK_ARG((act & SOS_ACT_MASK) != 2 || (slope && dslope), "sos_bn_act_backward: PReLU needs slope and dslope");
  const size_t smem = (size_t)3 * C * sizeof(float);
  bn_bwd_reduce_kernel<3><<<G, kThreads, smem, stream>>>(dz, dv, y, rows, C, scale, shift, mean, invstd, act, slope, partial);
  SOS_CHECK_LAUNCH("sos_bn_act_backward(reduce)");
  bn_bwd_finalize_kernel<3><<<ceil_div(C, kFinCh), kFinCh * kFinRows, 0, stream>>>(partial, G, C, (double)rows, dgamma, dbeta, (act & SOS_ACT_MASK) == 2 ? dslope : nullptr, m1, m2,
                                                                 nullptr, nullptr, 0, C);
  SOS_CHECK_LAUNCH("sos_bn_act_backward(finalize)");
  const long long tot4 = rows * (channels / 4);
  if (view_dense(dv, C) && tot4 < (1ll << 32))
    bn_bwd_apply_dense_kernel<0><<<grid_for(tot4, kThreads * kUnroll, 148 * 8), kThreads, 0, stream>>>(
        dz, y, dy, (unsigned)tot4, (unsigned)(channels / 4), scale, shift, mean, invstd, m1, m2, act, slope);
  else
    bn_bwd_apply_kernel<<<grid_for(tot4), kThreads, 0, stream>>>(dz, dv, y, dy, rows, C, scale, shift, mean, invstd, m1, m2, act, slope);
  SOS_CHECK_LAUNCH("sos_bn_act_backward(apply)");
  return SOS_OK;
}

int sos_bn_act_half(const void* y, int y_dtype, void* z_half, int64_t rows, int64_t channels, const float* scale, const float* shift, int act,
                    const float* slope, cudaStream_t stream) {
  SOS_CHECK_ARG(y && z_half && scale && shift && rows > 0 && channels >= 8 && channels % 8 == 0, "sos_bn_act_half: bad arguments (channels % 8)");
  SOS_CHECK_ARG(y_dtype == SOS_DTYPE_TF32 || y_dtype == SOS_DTYPE_F16, "sos_bn_act_half: unknown y type");
  SOS_CHECK_ARG((act & SOS_ACT_MASK) != 2 || slope, "sos_bn_act_half: PReLU needs a slope pointer");
  const long long tot8 = rows * (channels / 8);
  SOS_CHECK_ARG(tot8 < (1ll << 32), "sos_bn_act_half: too many elements");
  const int grid = grid_mult(grid_for(tot8, kThreads * 8, 148 * 8), (int)(channels / 8));
  if (y_dtype == SOS_DTYPE_F16)
    bn_act_h_kernel<__half><<<grid, kThreads, 0, stream>>>(reinterpret_cast<const __half*>(y), reinterpret_cast<uint4*>(z_half), (unsigned)tot8,
                                                           (unsigned)(channels / 8), scale, shift, act, slope);
  else
    bn_act_h_kernel<float><<<grid, kThreads, 0, stream>>>(reinterpret_cast<const float*>(y), reinterpret_cast<uint4*>(z_half), (unsigned)tot8,
                                                          (unsigned)(channels / 8), scale, shift, act, slope);
  SOS_CHECK_LAUNCH("sos_bn_act_half");
  return SOS_OK;
}

int sos_bn_act_backward_half(const void* dz, int dz_dtype, const float* dz_inv_scale, const void* y, int y_dtype, void* dy_half, int64_t rows,
                             int64_t channels, const float* scale, const float* shift, const float* mean, const float* invstd, int act,
                             const float* slope, float* partial, float* dgamma, float* dbeta, float* dslope, float* m1, float* m2, float* scal,
                             int accumulate_param_grads, int64_t real_channels, cudaStream_t stream) {
  const int G = sos_bn_partial_blocks(rows, channels);
  SOS_CHECK_ARG(dz && y && dy_half && scale && shift && mean && invstd && partial && dgamma && dbeta && m1 && m2 && scal && G > 0 &&
                    channels % 8 == 0 && channels <= 256,
                "sos_bn_act_backward_half: bad arguments");
  SOS_CHECK_ARG((dz_dtype == SOS_DTYPE_TF32 || dz_dtype == SOS_DTYPE_F16) && (y_dtype == SOS_DTYPE_TF32 || y_dtype == SOS_DTYPE_F16),
                "sos_bn_act_backward_half: unknown dz / y type");
  SOS_CHECK_ARG((act & SOS_ACT_MASK) != 2 || (slope && dslope), "sos_bn_act_backward_half: PReLU needs slope and dslope");
  SOS_CHECK_ARG(rows * (channels / 8) < (1ll << 32), "sos_bn_act_backward_half: too many elements");
  const int C = (int)channels, cr = (int)(real_channels > 0 ? real_channels : channels);
#define SOS_BN_BWD(TY, TDZ) \
  return bn_backward_h<TY, TDZ>(dz, y, dy_half, rows, C, scale, shift, mean, invstd, act, slope, partial, dgamma, dbeta, dslope, m1, m2, scal, \
                                accumulate_param_grads, cr, dz_inv_scale, G, stream)
  if (y_dtype == SOS_DTYPE_F16) {
    if (dz_dtype == SOS_DTYPE_F16) SOS_BN_BWD(__half, __half);
    SOS_BN_BWD(__half, float);
  }
  if (dz_dtype == SOS_DTYPE_F16) SOS_BN_BWD(float, __half);
  SOS_BN_BWD(float, float);
#undef SOS_BN_BWD
}

int sos_bn_act_backward_half_pre(const void* dz, int dz_dtype, const float* dz_inv_scale, const void* y, int y_dtype, void* dy_half, int64_t rows,
                                 int64_t channels, const float* scale, const float* shift, const float* mean, const float* invstd, int act,
                                 const float* slope, const float* partial, int64_t partial_rows, float* dgamma, float* dbeta, float* dslope, float* m1,
                                 float* m2, float* scal, int accumulate_param_grads, int64_t real_channels, cudaStream_t stream) {
  SOS_CHECK_ARG(dz && y && dy_half && scale && shift && mean && invstd && partial && partial_rows > 0 && dgamma && dbeta && m1 && m2 && scal &&
                    channels % 8 == 0 && channels <= 256,
                "sos_bn_act_backward_half_pre: bad arguments");
  SOS_CHECK_ARG(dz_dtype == SOS_DTYPE_F16 && y_dtype == SOS_DTYPE_F16 && (act & SOS_ACT_MASK) == 1,
                "sos_bn_act_backward_half_pre: half dz / y and ReLU only (the fused reduction of an encoder chain)");
  SOS_CHECK_ARG(rows * (channels / 8) < (1ll << 32), "sos_bn_act_backward_half_pre: too many elements");
  const int C = (int)channels, cr = (int)(real_channels > 0 ? real_channels : channels);
  return bn_backward_h<__half, __half>(dz, y, dy_half, rows, C, scale, shift, mean, invstd, act, slope, const_cast<float*>(partial), dgamma, dbeta,
                                       dslope, m1, m2, scal, accumulate_param_grads, cr, dz_inv_scale, (int)partial_rows, stream, true);
}

int sos_to_half(const float* x, int64_t rows, int64_t cs, void* out_half, int64_t cd, float* scal, cudaStream_t stream) {
  SOS_CHECK_ARG(x && out_half && rows > 0 && cs > 0 && cd >= cs && cd % 8 == 0, "sos_to_half: bad arguments (cd must be a multiple of 8, >= cs)");
  if (scal) {
    sumsq_kernel<<<grid_for(rows * cs, kThreads, 148 * 4), kThreads, 0, stream>>>(x, rows * cs, scal + 2);
    SOS_CHECK_LAUNCH("sos_to_half(sumsq)");
  }
  to_half_kernel<<<grid_for(rows * (cd / 8)), kThreads, 0, stream>>>(x, rows, (int)cs, reinterpret_cast<uint4*>(out_half), (int)cd, scal,
                                                                     (float)((double)rows * (double)cs));
  SOS_CHECK_LAUNCH("sos_to_half");
  return SOS_OK;
}

int sos_affine_act_backward(const float* dz, const int32_t* dz_view, const float* y, float* dy, int64_t rows, int64_t channels,
                            const float* scale, const float* shift, int act, const float* slope, cudaStream_t stream) {
  SOS_CHECK_ARG(dz && dz_view && y && dy && scale && shift && rows > 0 && channels >= 4 && channels % 4 == 0,
                "sos_affine_act_backward: bad arguments");
  const View dv = mk_view(dz_view);
  SOS_CHECK_ARG(view_ok(dv, (int)channels) && rows % ((long long)dv.H * dv.W) == 0, "sos_affine_act_backward: inconsistent view");
  SOS_CHECK_ARG((act & SOS_ACT_MASK) != 2 || slope, "sos_affine_act_backward: PReLU needs a slope pointer");
  const long long tot4 = rows * (channels / 4);
  if (view_dense(dv, (int)channels) && tot4 < (1ll << 32))
    bn_bwd_apply_dense_kernel<1><<<grid_for(tot4, kThreads * kUnroll, 148 * 8), kThreads, 0, stream>>>(
        dz, y, dy, (unsigned)tot4, (unsigned)(channels / 4), scale, shift, nullptr, nullptr, nullptr, nullptr, act, slope);
  else
    affine_act_bwd_kernel<<<grid_for(tot4), kThreads, 0, stream>>>(dz, dv, y, dy, rows, (int)channels, scale, shift, act, slope);
  SOS_CHECK_LAUNCH("sos_affine_act_backward");
  return SOS_OK;
}

int sos_transpose(const float* in, int64_t rows, int64_t cols, float* out, cudaStream_t stream) {
  SOS_CHECK_ARG(in && out && rows > 0 && cols > 0 && ceil_div_ll(rows, 32) <= 65535, "sos_transpose: bad arguments");
  dim3 grid((unsigned)ceil_div_ll(cols, 32), (unsigned)ceil_div_ll(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, out);
  SOS_CHECK_LAUNCH("sos_transpose");
  return SOS_OK;
}

int sos_split_tf32(const float* src, int64_t R, int64_t K, int64_t KP, int64_t stride_r, int64_t stride_k, int64_t k_shift, float* out,
                   int64_t ld_out, int64_t slot_stride, int n_slots, int64_t batch, int64_t src_batch_stride, int64_t out_batch_stride,
                   cudaStream_t stream) {
  SOS_CHECK_ARG(src && out && R > 0 && K > 0 && KP >= K && (n_slots >= 1 && n_slots <= 3), "sos_split_tf32: bad arguments");
  SOS_CHECK_ARG(stride_r == 1 || stride_k == 1, "sos_split_tf32: one of the source strides must be 1");
  SOS_CHECK_ARG(ceil_div_ll(R, 32) <= 65535 && batch >= 1 && batch <= 65535, "sos_split_tf32: too many rows / batches");
  dim3 grid((unsigned)ceil_div_ll(KP, 32), (unsigned)ceil_div_ll(R, 32), (unsigned)batch);
  split_tf32_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, R, K, KP, stride_r, stride_k, k_shift, out, ld_out, slot_stride, n_slots,
                                                      src_batch_stride, out_batch_stride);
  SOS_CHECK_LAUNCH("sos_split_tf32");
  return SOS_OK;
}

int sos_axpy(float* dst, const float* src, int64_t n, float alpha, const float* base_or_null, cudaStream_t stream) {
  SOS_CHECK_ARG(dst && src && n > 0, "sos_axpy: bad arguments");
  axpy_kernel<<<grid_for(n), kThreads, 0, stream>>>(dst, src, n, alpha, base_or_null);
  SOS_CHECK_LAUNCH("sos_axpy");
  return SOS_OK;
}

int sos_seq_map(const float* in, float* out, int64_t T, int64_t B, int64_t C, int64_t ld, int to_map, cudaStream_t stream) {
  SOS_CHECK_ARG(in && out && T > 0 && B > 0 && B <= 65535 && C > 0 && ld >= C, "sos_seq_map: bad arguments");
  dim3 grid((unsigned)ceil_div_ll(T, 32), (unsigned)ceil_div_ll(C, 32), (unsigned)B);
  seq_map_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, out, (int)T, (int)B, (int)C, (int)ld, to_map);
  SOS_CHECK_LAUNCH("sos_seq_map");
  return SOS_OK;
}

int sos_bias_act(float* y, int64_t rows, int64_t cols, int64_t ld, const float* bias, int act, cudaStream_t stream) {
  SOS_CHECK_ARG(y && rows > 0 && cols > 0 && ld >= cols && (act == 0 || act == 1 || act == 3), "sos_bias_act: bad arguments");
  bias_act_kernel<<<grid_for(rows * cols), kThreads, 0, stream>>>(y, rows, (int)cols, ld, bias, act);
  SOS_CHECK_LAUNCH("sos_bias_act");
  return SOS_OK;
}

int sos_bias_act_backward(const float* dy, const float* y, float* dpre, int64_t rows, int64_t cols, int64_t ld, int act,
                          float* dbias_or_null, cudaStream_t stream) {
  SOS_CHECK_ARG(dy && y && (dpre || dbias_or_null) && rows > 0 && cols > 0 && ld >= cols && (act == 0 || act == 1 || act == 3),
                "sos_bias_act_backward: bad arguments");
  long long gy = ceil_div_ll(rows, 8 * 16);
  if (gy > 64) gy = 64;
  dim3 grid((unsigned)ceil_div_ll(cols, 32), (unsigned)gy);
  bias_act_bwd_kernel<<<grid, dim3(32, 8), 0, stream>>>(dy, y, dpre, rows, (int)cols, ld, act, dbias_or_null);
  SOS_CHECK_LAUNCH("sos_bias_act_backward");
  return SOS_OK;
}

int sos_nchw_to_nhwc(const float* x, int64_t batch, int64_t channels, float* out, const int32_t* out_view, int64_t slice_channels,
                     cudaStream_t stream) {
  SOS_CHECK_ARG(x && out && out_view && batch > 0 && channels > 0 && slice_channels >= channels, "sos_nchw_to_nhwc: bad arguments");
  const View ov = mk_view(out_view);
  SOS_CHECK_ARG(view_ok(ov, (int)slice_channels), "sos_nchw_to_nhwc: inconsistent view");
  const long long P = batch * ov.H * ov.W;
  nchw_to_nhwc_kernel<<<grid_for(P * slice_channels), kThreads, 0, stream>>>(x, (int)channels, out, ov, (int)slice_channels, P);
  SOS_CHECK_LAUNCH("sos_nchw_to_nhwc");
  return SOS_OK;
}

int sos_nchw_to_nhwc_half(const float* x, int64_t batch, int64_t channels, int64_t H, int64_t W, void* out_half, int64_t padded_channels,
                          cudaStream_t stream) {
  SOS_CHECK_ARG(x && out_half && batch > 0 && channels > 0 && H > 0 && W > 0 && padded_channels >= channels && padded_channels % 8 == 0,
                "sos_nchw_to_nhwc_half: bad arguments (padded_channels must be a multiple of 8, >= channels)");
  const long long plane = H * W, P = batch * plane;
  nchw_to_nhwc_half_kernel<<<grid_for(P * (padded_channels / 8)), kThreads, 0, stream>>>(x, (int)channels, reinterpret_cast<uint4*>(out_half),
                                                                                        (int)padded_channels, plane, P);
  SOS_CHECK_LAUNCH("sos_nchw_to_nhwc_half");
  return SOS_OK;
}

int sos_accumulate_wgrad(const float* src, int64_t ntaps, int64_t rows_padded, int64_t cols_padded, int64_t rows, int64_t cols, float* dst,
                         cudaStream_t stream) {
  SOS_CHECK_ARG(src && dst && ntaps > 0 && rows > 0 && cols > 0 && rows_padded >= rows && cols_padded >= cols, "sos_accumulate_wgrad: bad arguments");
  accumulate_wgrad_kernel<false><<<grid_for(rows * cols * ntaps), kThreads, 0, stream>>>(const_cast<float*>(src), (int)ntaps, (int)rows_padded,
                                                                                        (int)cols_padded, (int)rows, (int)cols, dst);
  SOS_CHECK_LAUNCH("sos_accumulate_wgrad");
  return SOS_OK;
}

int sos_accumulate_wgrad_clear(float* src, int64_t ntaps, int64_t rows_padded, int64_t cols_padded, int64_t rows, int64_t cols, float* dst,
                               cudaStream_t stream) {
  SOS_CHECK_ARG(src && dst && ntaps > 0 && rows > 0 && cols > 0 && rows_padded >= rows && cols_padded >= cols, "sos_accumulate_wgrad_clear: bad arguments");
  accumulate_wgrad_kernel<true><<<grid_for(rows * cols * ntaps), kThreads, 0, stream>>>(src, (int)ntaps, (int)rows_padded, (int)cols_padded, (int)rows,
                                                                                       (int)cols, dst);
  SOS_CHECK_LAUNCH("sos_accumulate_wgrad_clear");
  return SOS_OK;
}

int sos_nhwc_to_nchw(const float* in, const int32_t* in_view, int64_t batch, int64_t channels, float* out, cudaStream_t stream) {
  SOS_CHECK_ARG(in && in_view && out && batch > 0 && channels > 0, "sos_nhwc_to_nchw: bad arguments");
  const View iv = mk_view(in_view);
  SOS_CHECK_ARG(iv.ld >= iv.coff + channels, "sos_nhwc_to_nchw: inconsistent view");
  const long long P = batch * iv.H * iv.W;
  nhwc_to_nchw_kernel<<<grid_for(P * channels), kThreads, 0, stream>>>(in, iv, out, (int)channels, P);
  SOS_CHECK_LAUNCH("sos_nhwc_to_nchw");
  return SOS_OK;
}

int sos_copy_view(const float* src, const int32_t* src_view, float* dst, const int32_t* dst_view, int64_t batch, int64_t channels,
                  int accumulate, cudaStream_t stream) {
  SOS_CHECK_ARG(src && dst && src_view && dst_view && batch > 0 && channels >= 4 && channels % 4 == 0, "sos_copy_view: bad arguments");
  const View sv = mk_view(src_view), dv = mk_view(dst_view);
  SOS_CHECK_ARG(view_ok(sv, (int)channels) && view_ok(dv, (int)channels), "sos_copy_view: inconsistent view");
  SOS_CHECK_ARG(batch <= 65535, "sos_copy_view: batch > 65535");
  copy_view_kernel<<<dim3((unsigned)dv.H, (unsigned)batch), kThreads, 0, stream>>>(src, sv, dst, dv, (int)channels, accumulate);
  SOS_CHECK_LAUNCH("sos_copy_view");
  return SOS_OK;
}

int sos_copy_view_fold(const float* grad_padded, const int32_t* padded_view, float* dst, const int32_t* dst_view, int64_t batch,
                       int64_t channels, cudaStream_t stream) {
  SOS_CHECK_ARG(grad_padded && dst && padded_view && dst_view && batch > 0 && batch <= 65535 && channels >= 4 && channels % 4 == 0,
                "sos_copy_view_fold: bad arguments");
  const View sv = mk_view(padded_view), dv = mk_view(dst_view);
  SOS_CHECK_ARG(view_ok(sv, (int)channels) && view_ok(dv, (int)channels), "sos_copy_view_fold: inconsistent view");
  SOS_CHECK_ARG(sv.H == dv.H && sv.W == dv.W && sv.ph == sv.pw && sv.ph >= 0 && sv.Hp == sv.H + 2 * sv.ph && sv.Wp == sv.W + 2 * sv.pw &&
                    sv.ph < sv.H && sv.ph < sv.W,
                "sos_copy_view_fold: the source view must be an H x W interior with a symmetric reflect border, the destination H x W");
  copy_view_fold_kernel<<<dim3((unsigned)dv.H, (unsigned)batch), kThreads, 0, stream>>>(grad_padded, sv, dst, dv, (int)channels);
  SOS_CHECK_LAUNCH("sos_copy_view_fold");
  return SOS_OK;
}

int sos_copy_view_backward(const float* grad_dst, const int32_t* dst_view, float* grad_src, const int32_t* src_view, int64_t batch,
                           int64_t channels, cudaStream_t stream) {
  SOS_CHECK_ARG(grad_dst && grad_src && src_view && dst_view && batch > 0 && batch <= 65535 && channels >= 4 && channels % 4 == 0,
                "sos_copy_view_backward: bad arguments");
  const View sv = mk_view(src_view), dv = mk_view(dst_view);
  SOS_CHECK_ARG(view_ok(sv, (int)channels) && view_ok(dv, (int)channels), "sos_copy_view_backward: inconsistent view");
  copy_view_bwd_kernel<<<dim3((unsigned)dv.H, (unsigned)batch), kThreads, 0, stream>>>(grad_dst, dv, grad_src, sv, (int)channels);
  SOS_CHECK_LAUNCH("sos_copy_view_backward");
  return SOS_OK;
}

int sos_reflect_fill(float* buf, int64_t batch, int64_t H, int64_t W, int64_t pad, int64_t channels, cudaStream_t stream) {
  SOS_CHECK_ARG(buf && batch > 0 && pad >= 0 && pad < H && pad < W && channels % 4 == 0, "sos_reflect_fill: bad arguments (pad must be < H, W)");
  if (pad == 0) return SOS_OK;
  SOS_CHECK_ARG(batch <= 65535, "sos_reflect_fill: batch > 65535");
  reflect_fill_kernel<<<dim3((unsigned)(H + 2 * pad), (unsigned)batch), kThreads, 0, stream>>>(buf, (int)H, (int)W, (int)pad, (int)channels);
  SOS_CHECK_LAUNCH("sos_reflect_fill");
  return SOS_OK;
}

int sos_reflect_fold(float* grad_buf, int64_t batch, int64_t H, int64_t W, int64_t pad, int64_t channels, cudaStream_t stream) {
  SOS_CHECK_ARG(grad_buf && batch > 0 && pad >= 0 && pad < H && pad < W && channels % 4 == 0, "sos_reflect_fold: bad arguments");
  if (pad == 0) return SOS_OK;
  // every interior element only reads border cells and itself, and writes itself -> safe in place
  SOS_CHECK_ARG(batch <= 65535, "sos_reflect_fold: batch > 65535");
  reflect_fold_kernel<<<dim3((unsigned)H, (unsigned)batch), kThreads, 0, stream>>>(grad_buf, (int)H, (int)W, (int)pad, (int)channels);
  SOS_CHECK_LAUNCH("sos_reflect_fold");
  return SOS_OK;
}

int sos_feat_to_seq(const float* in, int64_t B, int64_t F, int64_t T, int64_t C, float* out, int64_t V, int64_t ld, int64_t coff,
                    cudaStream_t stream) {
  SOS_CHECK_ARG(in && out && B > 0 && F > 0 && T > 0 && C > 0 && V > 0 && ld >= coff + C * F, "sos_feat_to_seq: bad arguments");
  feat_to_seq_kernel<<<grid_for(V * B * C * F), kThreads, 0, stream>>>(in, (int)B, (int)F, (int)T, (int)C, out, (int)V, (int)ld, (int)coff);
  SOS_CHECK_LAUNCH("sos_feat_to_seq");
  return SOS_OK;
}

int sos_feat_to_seq_backward(const float* grad_out, int64_t B, int64_t F, int64_t T, int64_t C, float* grad_in_zeroed, int64_t V,
                             int64_t ld, int64_t coff, cudaStream_t stream) {
  SOS_CHECK_ARG(grad_out && grad_in_zeroed && B > 0 && F > 0 && T > 0 && C > 0 && V > 0, "sos_feat_to_seq_backward: bad arguments");
  feat_to_seq_bwd_kernel<<<grid_for(V * B * C * F), kThreads, 0, stream>>>(grad_out, (int)B, (int)F, (int)T, (int)C, grad_in_zeroed, (int)V,
                                                                          (int)ld, (int)coff);
  SOS_CHECK_LAUNCH("sos_feat_to_seq_backward");
  return SOS_OK;
}

int sos_pack_conv_weight(const float* w, int64_t Cout, int64_t Cin, int64_t kh, int64_t kw, int64_t CinP, int64_t CoutP, int mode,
                         float* out, cudaStream_t stream) {
  SOS_CHECK_ARG(w && out && Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && CinP >= Cin && CoutP >= Cout && (mode == 0 || mode == 1),
                "sos_pack_conv_weight: bad arguments");
  const long long total = mode == 0 ? Cout * kh * kw * CinP : Cin * kh * kw * CoutP;
  pack_conv_weight_kernel<<<grid_for(total), kThreads, 0, stream>>>(w, (int)Cout, (int)Cin, (int)kh, (int)kw, (int)CinP, (int)CoutP, mode, out);
  SOS_CHECK_LAUNCH("sos_pack_conv_weight");
  return SOS_OK;
}

int sos_pack_taps(const float* w, int64_t rows, int64_t K, int64_t KP, int64_t row_stride, int64_t k_stride, int64_t ntaps,
                  const int32_t* tap_off, int round_tf32, float* out, cudaStream_t stream) {
  SOS_CHECK_ARG(w && out && tap_off && rows > 0 && K > 0 && KP >= K && ntaps > 0 && ntaps <= 49, "sos_pack_taps: bad arguments");
  TapOffsets t;
  for (int i = 0; i < (int)ntaps; ++i) t.off[i] = tap_off[i];
  pack_taps_kernel<<<grid_for(rows * ntaps * KP), kThreads, 0, stream>>>(w, (int)rows, (int)K, (int)KP, row_stride, k_stride, (int)ntaps, t,
                                                                         round_tf32, out);
  SOS_CHECK_LAUNCH("sos_pack_taps");
  return SOS_OK;
}

int sos_pack_taps_half(const float* w, int64_t rows, int64_t K, int64_t KP, int64_t row_stride, int64_t k_stride, int64_t ntaps,
                       const int32_t* tap_off, void* out_half, cudaStream_t stream) {
  SOS_CHECK_ARG(w && out_half && tap_off && rows > 0 && K > 0 && KP >= K && ntaps > 0 && ntaps <= 49, "sos_pack_taps_half: bad arguments");
  TapOffsets t;
  for (int i = 0; i < (int)ntaps; ++i) t.off[i] = tap_off[i];
  pack_taps_half_kernel<<<grid_for(rows * ntaps * KP), kThreads, 0, stream>>>(w, (int)rows, (int)K, (int)KP, row_stride, k_stride, (int)ntaps, t,
                                                                              reinterpret_cast<__half*>(out_half));
  SOS_CHECK_LAUNCH("sos_pack_taps_half");
  return SOS_OK;
}

int sos_im2col_half(const void* x_half, int64_t batch, int64_t H, int64_t W, int64_t channels, int64_t ntaps, const int32_t* tap_dh,
                    const int32_t* tap_dw, int64_t OH, int64_t OW, void* out_half, int64_t out_channels, cudaStream_t stream) {
  SOS_CHECK_ARG(x_half && out_half && tap_dh && tap_dw && batch > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "sos_im2col_half: bad arguments");
  SOS_CHECK_ARG(channels >= 2 && channels % 2 == 0 && ntaps > 0 && ntaps <= 32 && out_channels % 8 == 0 && out_channels >= 2 * ntaps,
                "sos_im2col_half: needs pixels of an even number of halves, at most 32 taps and out_channels >= 2 * ntaps (multiple of 8)");
  FoldTaps t;
  for (int i = 0; i < 32; ++i) {
    const int dh = i < ntaps ? tap_dh[i] : 0, dw = i < ntaps ? tap_dw[i] : 0;
    SOS_CHECK_ARG(dh >= -32768 && dh < 32768 && dw >= -32768 && dw < 32768, "sos_im2col_half: tap offset out of range");
    t.dh[i] = (int16_t)dh;
    t.dw[i] = (int16_t)dw;
  }
  const int g8 = (int)(out_channels / 8);
  const long long total = batch * OH * OW * g8;
  im2col_half_kernel<<<grid_for(total), kThreads, 0, stream>>>(reinterpret_cast<const __half*>(x_half), (int)H, (int)W, (int)channels, (int)OH,
                                                               (int)OW, (int)ntaps, t, reinterpret_cast<uint4*>(out_half), g8, total);
  SOS_CHECK_LAUNCH("sos_im2col_half");
  return SOS_OK;
}

int sos_pack_taps_half_multi(const sos_pack_desc* descs_device, int64_t n, cudaStream_t stream) {
  SOS_CHECK_ARG(descs_device && n > 0 && n <= 65535, "sos_pack_taps_half_multi: bad arguments");
  pack_taps_half_multi_kernel<<<dim3((unsigned)n, 32), kThreads, 0, stream>>>(descs_device);
  SOS_CHECK_LAUNCH("sos_pack_taps_half_multi");
  return SOS_OK;
}

int sos_unpack_wgrad(const float* src, int64_t Cout, int64_t Cin, int64_t ntaps, int64_t CinP, float* dst, int accumulate,
                     cudaStream_t stream) {
  SOS_CHECK_ARG(src && dst && Cout > 0 && Cin > 0 && ntaps > 0 && CinP >= Cin, "sos_unpack_wgrad: bad arguments");
  unpack_wgrad_kernel<<<grid_for(Cout * Cin * ntaps), kThreads, 0, stream>>>(src, (int)Cout, (int)Cin, (int)ntaps, (int)CinP, dst, accumulate);
  SOS_CHECK_LAUNCH("sos_unpack_wgrad");
  return SOS_OK;
}

}  // extern "C"
