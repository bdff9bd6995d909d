// Silent-interval gate shared by the STFT frame builders and the standalone gate kernel.
#pragma once
#include <stdint.h>

// The reference's bit-string -> sample mask (1 = silent), M2/tools.py:340-362.
// frame_lo[i] = int(i * ratio) computed on the host with the reference's own
// float expression; frame i writes [frame_lo[i], frame_lo[i+1]-1) and the
// one-sample gap frame_lo[i+1]-1 (and any tail) stays 0.  The second pass
// flips every run of equal values shorter than 5 samples (runs are taken on
// the pre-flip values), which fills the gap between two silent frames and
// handles clips truncated in the middle of a frame.
static __device__ __forceinline__ int frame_of(int s, int nb, const int* __restrict__ frame_lo, float inv_ratio) {
  int i = (int)((float)s * inv_ratio);
  if (i > nb) i = nb;
  while (i > 0 && frame_lo[i] > s) --i;
  while (i < nb && frame_lo[i + 1] <= s) ++i;
  return i;                                        // nb means "after the last frame"
}

static __device__ __forceinline__ int preflip(int s, const uint8_t* __restrict__ bits, int nb, const int* __restrict__ frame_lo,
                                       float inv_ratio) {
  const int i = frame_of(s, nb, frame_lo, inv_ratio);
  if (i >= nb) return 0;
  return (s < frame_lo[i + 1] - 1 && bits[i] == 0) ? 1 : 0;
}

static __device__ __noinline__ float sample_mask(int s, int L, const uint8_t* __restrict__ bits, int nb,
                                             const int* __restrict__ frame_lo, float inv_ratio) {
  const int i = frame_of(s, nb, frame_lo, inv_ratio);
  if (i < nb) {                                    // fast path: deep inside a frame
    const int lo = frame_lo[i], hi = frame_lo[i + 1] - 1;
    if (s - lo >= 4 && hi - s > 4 && L - s > 4) return bits[i] == 0 ? 1.f : 0.f;
  }
  const int v = preflip(s, bits, nb, frame_lo, inv_ratio);
  int a = 0, b = 0;
  while (a < 4 && s - a - 1 >= 0 && preflip(s - a - 1, bits, nb, frame_lo, inv_ratio) == v) ++a;
  while (b < 4 && s + b + 1 < L && preflip(s + b + 1, bits, nb, frame_lo, inv_ratio) == v) ++b;
  return (a + 1 + b < 5) ? (float)(1 - v) : (float)v;
}

