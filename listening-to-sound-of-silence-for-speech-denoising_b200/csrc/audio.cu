// Audio loading / evaluation kernels beside the hot path:
//   sos_resample   librosa.load's resampler = resampy.resample(filter='kaiser_best') (M2/predict.py:303, M1/dataset.py:226):
//                  J. O. Smith band-limited interpolation over a 64-zero-crossing Kaiser-windowed sinc table; one thread per
//                  output sample, taps and accumulation in exactly the published order (the accumulator rounds to fp32 per tap,
//                  as the reference's float32 output array does), so that the result matches tap for tap.
//   sos_wss        weighted spectral slope distance per frame (M2/metrics.py:404-558)
//   sos_llr        log-likelihood ratio per frame from order-P LPC (M2/metrics.py:561-681)
// The two metrics are evaluation-only: one block per (clip, frame), double arithmetic like the reference's numpy.
#include "common.cuh"
#include "sos_b200.h"
#include <math.h>

namespace {

__global__ void resample_kernel(const float* __restrict__ x, long long n_orig, float* __restrict__ y, long long n_out,
                                const double* __restrict__ win, const double* __restrict__ delta, int nwin, int num_table,
                                double sample_ratio, const double* __restrict__ time_reg) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const float* xb = x + (size_t)blockIdx.y * n_orig;
  const double scale = sample_ratio < 1.0 ? sample_ratio : 1.0;
  const int index_step = (int)(scale * num_table);
  const double time_register = time_reg[t];            // (the reference advances it by repeated addition: a host-side cumulative sum)
  const long long n = (long long)time_register;
  double frac = scale * (time_register - (double)n);
  double index_frac = frac * num_table;
  int offset = (int)index_frac;
  double eta = index_frac - offset;
  long long i_max = (nwin - offset) / index_step;
  if (n + 1 < i_max) i_max = n + 1;
  float acc = 0.f;
  for (long long i = 0; i < i_max; ++i) {
    const int idx = offset + (int)i * index_step;
    const double w = win[idx] + eta * delta[idx];
    acc = (float)((double)acc + w * (double)xb[n - i]);
  }
  frac = scale - frac;
  index_frac = frac * num_table;
  offset = (int)index_frac;
  eta = index_frac - offset;
  long long k_max = (nwin - offset) / index_step;
  if (n_orig - n - 1 < k_max) k_max = n_orig - n - 1;
  for (long long k = 0; k < k_max; ++k) {
    const int idx = offset + (int)k * index_step;
    const double w = win[idx] + eta * delta[idx];
    acc = (float)((double)acc + w * (double)xb[n + k + 1]);
  }
  y[(size_t)blockIdx.y * n_out + t] = acc;
}

// ------------------------------------------------------------------------------------------------ WSS / LLR
constexpr int kNumCrit = 25;
__constant__ double c_cent[kNumCrit] = {50., 120, 190, 260, 330, 400, 470, 540, 617.372, 703.378, 798.717, 904.128, 1020.38, 1148.30,
                                        1288.72, 1442.54, 1610.70, 1794.16, 1993.93, 2211.08, 2446.71, 2701.97, 2978.04, 3276.17, 3597.63};
__constant__ double c_bw[kNumCrit] = {70., 70, 70, 70, 70, 70, 70, 77.3724, 86.0056, 95.3398, 105.411, 116.256, 127.914, 140.423,
                                      153.823, 168.154, 183.457, 199.776, 217.153, 235.631, 255.255, 276.072, 298.126, 321.465, 346.136};

// One block (256 threads) per (frame, clip): the two Hann-windowed frames go through ONE complex radix-2 FFT (clean = real part,
// processed = imaginary part; the two spectra separate by conjugate symmetry), then 25 Gaussian critical-band energies each, the
// slopes, nearest-peak weights and the weighted slope distance (Klatt 1982), all as in the reference's loop body.
__global__ void __launch_bounds__(256) wss_kernel(const float* __restrict__ ref, const float* __restrict__ deg, long long length, int winlength,
                                                  int skip, int n_fft, int log2n, double max_freq, double eps, int num_frames,
                                                  double* __restrict__ out) {
  extern __shared__ double sm[];                  // re[n_fft] | im[n_fft] | energies 2 x 25
  double* re = sm;
  double* im = sm + n_fft;
  double* en = sm + 2 * n_fft;
  const int frame = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* r = ref + (size_t)b * length + (size_t)frame * skip;
  const float* d = deg + (size_t)b * length + (size_t)frame * skip;
  for (int i = tid; i < n_fft; i += 256) {
    // bit-reversed load
    unsigned j = __brev((unsigned)i) >> (32 - log2n);
    double a = 0.0, c = 0.0;
    if ((int)j < winlength) {
      const double w = 0.5 * (1.0 - cos(2.0 * 3.14159265358979323846 * ((double)(j + 1) / (double)(winlength + 1))));
      a = (double)r[j] * w;
      c = (double)d[j] * w;
    }
    re[i] = a;
    im[i] = c;
  }
  __syncthreads();
  for (int s = 1; s <= log2n; ++s) {
    const int m = 1 << s, half = m >> 1;
    for (int k = tid; k < n_fft / 2; k += 256) {
      const int grp = k / half, pos = k - grp * half;
      const int i0 = grp * m + pos, i1 = i0 + half;
      double sn, cs;
      sincos(-2.0 * 3.14159265358979323846 * (double)pos / (double)m, &sn, &cs);
      const double tr = re[i1] * cs - im[i1] * sn, ti = re[i1] * sn + im[i1] * cs;
      re[i1] = re[i0] - tr;
      im[i1] = im[i0] - ti;
      re[i0] += tr;
      im[i0] += ti;
    }
    __syncthreads();
  }
  // Z = X + i Y  ->  X[k] = (Z[k] + conj Z[N-k]) / 2,  Y[k] = (Z[k] - conj Z[N-k]) / (2 i)
  const int nby2 = n_fft / 2;
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = warp; c < kNumCrit; c += 8) {
    const double f0 = floor((c_cent[c] / max_freq) * nby2), bw = (c_bw[c] / max_freq) * nby2;
    const double norm = log(c_bw[0]) - log(c_bw[c]), min_factor = exp(-30.0 / (2.0 * 2.303));
    double ec = 0.0, ep = 0.0;
    for (int j = lane; j < nby2; j += 32) {
      const double q = ((double)j - f0) / bw;
      double f = exp(-11.0 * q * q + norm);
      if (!(f > min_factor)) f = 0.0;
      if (f != 0.0) {
        const int jn = (n_fft - j) & (n_fft - 1);
        const double zr = re[j], zi = im[j], wr = re[jn], wi = -im[jn];
        const double xr = 0.5 * (zr + wr), xi = 0.5 * (zi + wi);          // X[j]
        const double yr = 0.5 * (zi - wi), yi = -0.5 * (zr - wr);         // Y[j]
        ec += (xr * xr + xi * xi) * f;
        ep += (yr * yr + yi * yi) * f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ec += __shfl_xor_sync(0xffffffffu, ec, o);
      ep += __shfl_xor_sync(0xffffffffu, ep, o);
    }
    if (lane == 0) {
      en[c] = 10.0 * log10(fmax(ec, eps));
      en[kNumCrit + c] = 10.0 * log10(fmax(ep, eps));
    }
  }
  __syncthreads();
  if (tid == 0) {
    const double* ce = en;
    const double* pe = en + kNumCrit;
    double cs[kNumCrit - 1], ps[kNumCrit - 1];
    for (int i = 0; i < kNumCrit - 1; ++i) { cs[i] = ce[i + 1] - ce[i]; ps[i] = pe[i + 1] - pe[i]; }
    double cmax = ce[0], pmax = pe[0];
    for (int i = 1; i < kNumCrit; ++i) { cmax = fmax(cmax, ce[i]); pmax = fmax(pmax, pe[i]); }
    const double Kmax = 20.0, Klocmax = 1.0;
    double num = 0.0, wsum = 0.0;
    for (int i = 0; i < kNumCrit - 1; ++i) {
      double cpk, ppk;
      int n;
      if (cs[i] > 0) { n = i; while (n < kNumCrit - 1 && cs[n] > 0) ++n; cpk = ce[n - 1]; }
      else { n = i; while (n >= 0 && cs[n] <= 0) --n; cpk = ce[n + 1]; }
      if (ps[i] > 0) { n = i; while (n < kNumCrit - 1 && ps[n] > 0) ++n; ppk = pe[n - 1]; }
      else { n = i; while (n >= 0 && ps[n] <= 0) --n; ppk = pe[n + 1]; }
      const double wc = (Kmax / (Kmax + cmax - ce[i])) * (Klocmax / (Klocmax + cpk - ce[i]));
      const double wp = (Kmax / (Kmax + pmax - pe[i])) * (Klocmax / (Klocmax + ppk - pe[i]));
      const double w = 0.5 * (wc + wp), dsl = cs[i] - ps[i];
      num += w * dsl * dsl;
      wsum += w;
    }
    out[(size_t)b * num_frames + frame] = num / wsum;
  }
}

// One warp per (frame, clip): autocorrelation lags 0..P of the two windowed frames (double), Levinson-Durbin, then
// log(a_p R_c a_p^T / a_c R_c a_c^T) with the float32 casts the reference applies to R and a before the quadratic forms.
__global__ void __launch_bounds__(32) llr_kernel(const float* __restrict__ ref, const float* __restrict__ deg, long long length, int winlength,
                                                 int skip, int P, int num_frames, double* __restrict__ out) {
  extern __shared__ double fr[];                 // clean[winlength] | processed[winlength]
  const int frame = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
  const float* r = ref + (size_t)b * length + (size_t)frame * skip;
  const float* d = deg + (size_t)b * length + (size_t)frame * skip;
  for (int j = lane; j < winlength; j += 32) {
    const double w = 0.5 * (1.0 - cos(2.0 * 3.14159265358979323846 * ((double)(j + 1) / (double)(winlength + 1))));
    fr[j] = (double)r[j] * w;
    fr[winlength + j] = (double)d[j] * w;
  }
  __syncwarp();
  __shared__ double R[2][20];
  for (int sig = 0; sig < 2; ++sig) {
    const double* f = fr + sig * winlength;
    for (int k = 0; k <= P; ++k) {
      double a = 0.0;
      for (int j = lane; j < winlength - k; j += 32) a += f[j] * f[j + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) R[sig][k] = a;
    }
  }
  __syncwarp();
  if (lane == 0) {
    float A[2][20];
    for (int sig = 0; sig < 2; ++sig) {
      double a[20], E[21], past[20];
      const double* Rs = R[sig];
      for (int i = 0; i < P; ++i) a[i] = 1.0;
      E[0] = Rs[0];
      for (int i = 0; i < P; ++i) {
        double sum_term = 0.0;
        for (int j = 0; j < i; ++j) { past[j] = a[j]; sum_term += a[j] * Rs[i - j]; }
        const double rc = (Rs[i + 1] - sum_term) / E[i];
        a[i] = rc;
        for (int j = 0; j < i; ++j) a[j] = past[j] - rc * past[i - 1 - j];
        E[i + 1] = (1.0 - rc * rc) * E[i];
      }
      A[sig][0] = 1.f;
      for (int i = 0; i < P; ++i) A[sig][i + 1] = (float)(-a[i]);
    }
    float Rc[20];
    for (int k = 0; k <= P; ++k) Rc[k] = (float)R[0][k];
    float num = 0.f, den = 0.f;
    for (int i = 0; i <= P; ++i) {
      float tn = 0.f, td = 0.f;
      for (int j = 0; j <= P; ++j) {
        const float rij = Rc[i > j ? i - j : j - i];
        tn += A[1][j] * rij;
        td += A[0][j] * rij;
      }
      num += tn * A[1][i];
      den += td * A[0][i];
    }
    out[(size_t)b * num_frames + frame] = (double)logf(num / den);
  }
}

}  // namespace

extern "C" int sos_resample(const float* x, int64_t batch, int64_t n_in, float* y, int64_t n_out, const double* interp_win,
                            const double* interp_delta, int64_t n_win, int64_t num_table, double sample_ratio, const double* time_register,
                            cudaStream_t stream) {
  SOS_CHECK_ARG(x && y && interp_win && interp_delta && time_register && batch > 0 && batch <= 65535 && n_in > 0 && n_out > 0 && n_win > 1 &&
                    num_table > 0 && sample_ratio > 0,
                "sos_resample: bad arguments");
  const double scale = sample_ratio < 1.0 ? sample_ratio : 1.0;
  SOS_CHECK_ARG((int)(scale * num_table) >= 1, "sos_resample: the filter table is too coarse for this ratio");
  dim3 grid((unsigned)ceil_div_ll(n_out, 128), (unsigned)batch);
  resample_kernel<<<grid, 128, 0, stream>>>(x, n_in, y, n_out, interp_win, interp_delta, (int)n_win, (int)num_table, sample_ratio, time_register);
  SOS_CHECK_LAUNCH("sos_resample");
  return SOS_OK;
}

static int metric_frames(int64_t length, int64_t srate, int* winlength, int* skip) {
  *winlength = (int)lround(30.0 * (double)srate / 1000.0);
  *skip = (int)floor(*winlength / 4.0);
  return (int)((double)length / *skip - ((double)*winlength / *skip));
}

extern "C" int sos_metric_frames(int64_t length, int64_t srate) {
  int w, s;
  return length > 0 && srate >= 1000 ? metric_frames(length, srate, &w, &s) : 0;
}

extern "C" int sos_wss(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double eps, double* dist_out,
                       cudaStream_t stream) {
  SOS_CHECK_ARG(ref && deg && dist_out && batch > 0 && batch <= 65535 && srate >= 1000, "sos_wss: bad arguments");
  int winlength, skip;
  const int nf = metric_frames(length, srate, &winlength, &skip);
  SOS_CHECK_ARG(nf > 0, "sos_wss: clip too short (no 30 ms frame)");
  int log2n = 0;
  while ((1 << log2n) < 2 * winlength) ++log2n;
  const int n_fft = 1 << log2n;
  const size_t smem = (size_t)(2 * n_fft + 2 * kNumCrit) * sizeof(double);
  SOS_CHECK_ARG(smem <= 200 * 1024, "sos_wss: frame too long for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(wss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  wss_kernel<<<dim3((unsigned)nf, (unsigned)batch), 256, smem, stream>>>(ref, deg, length, winlength, skip, n_fft, log2n, (double)srate / 2.0, eps, nf,
                                                                       dist_out);
  SOS_CHECK_LAUNCH("sos_wss");
  return SOS_OK;
}

extern "C" int sos_llr(const float* ref, const float* deg, int64_t batch, int64_t length, int64_t srate, double* dist_out, cudaStream_t stream) {
  SOS_CHECK_ARG(ref && deg && dist_out && batch > 0 && batch <= 65535 && srate >= 1000, "sos_llr: bad arguments");
  int winlength, skip;
  const int nf = metric_frames(length, srate, &winlength, &skip);
  SOS_CHECK_ARG(nf > 0, "sos_llr: clip too short (no 30 ms frame)");
  const int P = srate < 10000 ? 10 : 16;
  const size_t smem = (size_t)2 * winlength * sizeof(double);
  SOS_CHECK_ARG(smem <= 40 * 1024, "sos_llr: frame too long");
  llr_kernel<<<dim3((unsigned)nf, (unsigned)batch), 32, smem, stream>>>(ref, deg, length, winlength, skip, P, nf, dist_out);
  SOS_CHECK_LAUNCH("sos_llr");
  return SOS_OK;
}
