// On-device construction of a training item from waveforms, replacing the reference's CPU Dataset workers:
//   add_signals / add_noise_to_audio   M2/tools.py:217-303   (mix clean speech and a noise crop at a given SNR, normalise to `norm`)
//   fast_cRM_sigmoid                   M2/transform.py:36-54,92-94,130-138   (the Dataset's "mask" target, M2/dataset.py:239)
#include "common.cuh"
#include "sos_b200.h"
#include <math.h>

namespace {

constexpr int kMixThreads = 1024;

__device__ __forceinline__ double block_sum_d(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  v = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;                                                       // every thread holds the total
}
__device__ __forceinline__ float block_max_f(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  v = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One block per clip; the clip (2 x L floats) is re-read from L2 in the second and third pass.  fp32 element arithmetic in the
// reference's order (noise / ratio, signal + new_noise, x / scale); the two power sums are accumulated in double.
__global__ void __launch_bounds__(kMixThreads) add_signals_kernel(const float* __restrict__ signal, const float* __restrict__ noise,
                                                                   const float* __restrict__ snr_db, int L, float norm,
                                                                   float* __restrict__ mixed, float* __restrict__ clean,
                                                                   float* __restrict__ full_noise) {
  __shared__ double smd[32];
  __shared__ float smf[32];
  const size_t base = (size_t)blockIdx.x * L;
  const float* s = signal + base;
  const float* n = noise + base;
  double ps = 0.0, pn_ = 0.0;
  for (int i = threadIdx.x; i < L; i += kMixThreads) {
    const float a = s[i], b = n[i];
    ps += (double)(a * a);
    pn_ += (double)(b * b);
  }
  ps = block_sum_d(ps, smd);
  pn_ = block_sum_d(pn_, smd);
  const float sp = (float)ps, np_ = (float)pn_;
  float ratio = 0.f;                                               // 0: the noise is added as it is (M2/tools.py:240-248)
  if (sp != 0.f) {
    const float target = sp / powf(10.f, snr_db[blockIdx.x] / 10.f);
    ratio = sqrtf(np_) / sqrtf(target);
  }
  float mx = 0.f;
  for (int i = threadIdx.x; i < L; i += kMixThreads) {
    const float nn = ratio == 0.f ? n[i] : n[i] / ratio;
    mx = fmaxf(mx, fabsf(s[i] + nn));
  }
  mx = block_max_f(mx, smf);
  const float scale = norm != 0.f ? mx / norm : 0.f;               // 0: no normalisation (M2/tools.py:266-274)
  for (int i = threadIdx.x; i < L; i += kMixThreads) {
    const float a = s[i];
    const float nn = ratio == 0.f ? n[i] : n[i] / ratio;
    const float m = a + nn;
    if (scale != 0.f) {
      mixed[base + i] = m / scale;
      clean[base + i] = a / scale;
      full_noise[base + i] = nn / scale;
    } else {
      mixed[base + i] = m;
      clean[base + i] = a;
      full_noise[base + i] = nn;
    }
  }
}

// crm = sigmoid(a M - b), M = S / Y (complex ratio of clean over mixed, eps 1e-8 in the denominator); (B, 2, plane) layout.
__global__ void crm_fwd_kernel(const float* __restrict__ S, const float* __restrict__ Y, float* __restrict__ crm, long long plane,
                               long long total, float a, float b) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long bi = e / plane, r = e - bi * plane;
    const long long i0 = bi * 2 * plane + r, i1 = i0 + plane;
    const float yr = Y[i0], yi = Y[i1], sr = S[i0], si = S[i1];
    const float den = yr * yr + yi * yi + 1e-8f;
    const float mr = (yr * sr + yi * si) / den, mi = (yr * si - yi * sr) / den;
    crm[i0] = 1.f / (1.f + expf(-a * mr + b));
    crm[i1] = 1.f / (1.f + expf(-a * mi + b));
  }
}

// Segmental SNR (M2/metrics.py:86-175).  One block per clip, a warp per frame: frame f covers samples [f * skip, f * skip + win)
// under the Hann-like window 0.5 (1 - cos(2 pi (n + 1) / (win + 1))); each frame's 10 log10(Es / (En + eps) + off) is clamped to
// [min_snr, max_snr] and the frames are averaged; the overall SNR uses the unwindowed sums.  Sums in double.
__global__ void __launch_bounds__(256) ssnr_kernel(const float* __restrict__ ref, const float* __restrict__ deg, int L, int win, int skip,
                                                   int n_frames, float min_snr, float max_snr, double eps, double off,
                                                   float* __restrict__ overall, float* __restrict__ segmental) {
  __shared__ double smd[32];
  const float* r = ref + (size_t)blockIdx.x * L;
  const float* g = deg + (size_t)blockIdx.x * L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double ps = 0.0, pd = 0.0;
  for (int i = threadIdx.x; i < L; i += 256) {
    const double a = r[i], d = (double)r[i] - (double)g[i];
    ps += a * a;
    pd += d * d;
  }
  ps = block_sum_d(ps, smd);
  pd = block_sum_d(pd, smd);
  double seg = 0.0;
  for (int f = warp; f < n_frames; f += 8) {
    double se = 0.0, ne = 0.0;
    for (int n = lane; n < win; n += 32) {
      const double w = 0.5 * (1.0 - cos(6.283185307179586476925286766559 * (double)(n + 1) / (double)(win + 1)));
      const double c = (double)r[f * skip + n] * w, p = (double)g[f * skip + n] * w;
      se += c * c;
      ne += (c - p) * (c - p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o);
      ne += __shfl_xor_sync(0xffffffffu, ne, o);
    }
    double v = 10.0 * log10(se / (ne + eps) + off);
    v = fmin(fmax(v, (double)min_snr), (double)max_snr);
    if (lane == 0) seg += v;
  }
  seg = block_sum_d(seg, smd);
  if (threadIdx.x == 0) {
    overall[blockIdx.x] = (float)(10.0 * log10(ps / (pd + eps)));
    segmental[blockIdx.x] = n_frames > 0 ? (float)(seg / n_frames) : nanf("");
  }
}

}  // namespace

extern "C" int sos_ssnr(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double win_len_ms, float min_snr,
                        float max_snr, double eps, int shift, float* overall_out, float* segmental_out, cudaStream_t stream) {
  SOS_CHECK_ARG(ref && deg && overall_out && segmental_out && batch > 0 && length > 0 && length < (1ll << 31) && srate > 0 && win_len_ms > 0,
                "sos_ssnr: bad arguments");
  const int win = (int)nearbyint(win_len_ms * (double)srate / 1000.0);          // int(np.round(win_len * srate / 1000)): half to even
  const int skip = win / 4;
  SOS_CHECK_ARG(win >= 4, "sos_ssnr: window of %d samples is too short", win);
  const int n_frames = (int)((double)length / skip - ((double)win / skip));    // int(clean_length / skiprate - winlength / skiprate)
  ssnr_kernel<<<(unsigned)batch, 256, 0, stream>>>(ref, deg, (int)length, win, skip, n_frames < 0 ? 0 : n_frames, min_snr, max_snr, eps,
                                                   shift ? 1.0 : eps, overall_out, segmental_out);
  SOS_CHECK_LAUNCH("sos_ssnr");
  return SOS_OK;
}

extern "C" int sos_add_signals(const float* signal, const float* noise, const float* snr_db, int64_t batch, int64_t length, float norm,
                               float* mixed, float* clean, float* full_noise, cudaStream_t stream) {
  SOS_CHECK_ARG(signal && noise && snr_db && mixed && clean && full_noise && batch > 0 && length > 0 && length < (1ll << 31) && norm >= 0.f,
                "sos_add_signals: bad arguments");
  add_signals_kernel<<<(unsigned)batch, kMixThreads, 0, stream>>>(signal, noise, snr_db, (int)length, norm, mixed, clean, full_noise);
  SOS_CHECK_LAUNCH("sos_add_signals");
  return SOS_OK;
}

extern "C" int sos_crm_forward(const float* clean_spec, const float* mixed_spec, float* crm, int64_t batch, int64_t plane, float a, float b,
                               cudaStream_t stream) {
  SOS_CHECK_ARG(clean_spec && mixed_spec && crm && batch > 0 && plane > 0, "sos_crm_forward: bad arguments");
  const long long total = batch * plane;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  crm_fwd_kernel<<<(unsigned)g, 256, 0, stream>>>(clean_spec, mixed_spec, crm, plane, total, a, b);
  SOS_CHECK_LAUNCH("sos_crm_forward");
  return SOS_OK;
}
