// STFT on the 5th-gen tensor cores at fp32-grade accuracy (reference: M2/transform.py:188-193 -> librosa.stft(y, 510, 158, 400)).
//
// The transform is a real DFT of a Hann(400) window centred in a 510-point frame.  With c = 158 t the frame centre and
// d = -200..199 the offset from it, the window is even in d and the DFT phase at the centre is (-1)^k, so
//     re[k] =  (-1)^k  sum_{d=0..199} w_d cos(2 pi k d / 510) e[d],   e[0] = x[c],  e[d] = x[c+d] + x[c-d]
//     im[k] = -(-1)^k  sum_{d=1..199} w_d sin(2 pi k d / 510) o[d],               o[d] = x[c+d] - x[c-d]
// -- two real 200 x 256 products per frame, half the multiply-adds of the 400 x 512 form.  They run as tcgen05 kind::tf32
// MMAs with both operands split x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and hi*hi + lo*hi + hi*lo accumulated in fp32
// (error ~2^-21 of scale, i.e. fp32 grade; the tables are split on the host from float64).
//
// CTA = one tile of 128 consecutive frames of the (clip, frame) index, 320 threads:
//   warp 0     TMA: the table chunks (128 bins x 32 offsets, hi and lo) of each stage through a 3-slot ring
//   warp 1     MMA issuer: accumulators re | im = 2 x 256 TMEM columns
//   warps 2-9  frame builders: a warp takes one frame row at a time (lane = offset: two coalesced 128-byte reads), folds +
//              (optionally) gates + splits, and writes the row of the re-stage AND the im-stage tile in the K-major
//              128-byte-swizzled layout the MMA reads; afterwards the same warps drain TMEM straight to the (B, 2, 256, T)
//              output (thread = frame row, so a warp's 32 lanes hold 32 consecutive frames of one bin: coalesced)
// 28 stages (7 offset chunks x {re, im} x two halves of the bins).  The frame tiles of a chunk (e and o, hi and lo: 64 KB) are
// double buffered, so chunk c+1 is built while chunk c is multiplied; after the last stage warps 2-5 drain re, warps 6-9 im.
#include "common.cuh"
#include "gate.cuh"
#include "ptx.cuh"
#include "sos_b200.h"
#include "tc_common.cuh"
#include <math.h>
#include <vector>

namespace {

using namespace ptx;
using namespace tc;

constexpr int kHop = 158, kBins = 256, kHalf = 200;      // kHalf: offsets d = 0..199
constexpr int kChunk = 32, kChunks = 7, kKpad = kChunk * kChunks;   // 224 >= 200
constexpr int kStages = 4 * kChunks;                      // (chunk, re/im, half of the bins)
constexpr int kBuilderWarps = 8;
constexpr int kThreadsStft = 64 + 32 * kBuilderWarps;
constexpr int kNHalf = kBins / 2;                         // MMA N
constexpr uint32_t kATile = 128 * 128;                    // 128 frames x 32 offsets fp32
constexpr uint32_t kABuf = 4 * kATile;                    // e_hi | e_lo | o_hi | o_lo of one chunk
constexpr uint32_t kBTile = kNHalf * 128;                 // 128 bins x 32 offsets fp32
constexpr uint32_t kBSlot = 2 * kBTile;                   // hi | lo
constexpr int kBSlots = 3;

float* g_tab = nullptr;          // [4: cos_hi, cos_lo, sin_hi, sin_lo][256 bins][224 offsets]
CUtensorMap g_tab_map;

inline float tf32_round_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & ~0x1FFFu;                           // round to nearest, ties away (cvt.rna.tf32)
  memcpy(&f, &u, 4);
  return f;
}

int init_stft_tc() {
  if (g_tab) return SOS_OK;
  std::vector<float> tab((size_t)4 * kBins * kKpad, 0.f);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < kBins; ++k) {
    const double sgn = (k & 1) ? -1.0 : 1.0;
    for (int d = 0; d < kHalf; ++d) {
      const double w = 0.5 + 0.5 * cos(two_pi * d / 400.0);                 // hann(400)[200 + d]
      const int ph = (int)(((long long)k * d) % 510);
      const double cv = sgn * w * cos(two_pi * ph / 510.0), sv = -sgn * w * sin(two_pi * ph / 510.0);
      const double vals[2] = {cv, sv};
      for (int m = 0; m < 2; ++m) {
        const float hi = tf32_round_host((float)vals[m]);
        const float lo = tf32_round_host((float)(vals[m] - (double)hi));
        tab[((size_t)(2 * m) * kBins + k) * kKpad + d] = hi;
        tab[((size_t)(2 * m + 1) * kBins + k) * kKpad + d] = lo;
      }
    }
  }
  if (cudaMalloc(&g_tab, tab.size() * 4) != cudaSuccess) {
    g_tab = nullptr;
    sos_set_error("stft: cudaMalloc of the DFT tables failed");
    return SOS_ERR_CUDA;
  }
  cudaMemcpy(g_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice);
  uint64_t dims[2] = {(uint64_t)kKpad, (uint64_t)4 * kBins};
  uint64_t str[2] = {4, (uint64_t)kKpad * 4};
  uint32_t box[2] = {(uint32_t)kChunk, (uint32_t)kNHalf};
  uint32_t es[2] = {1, 1};
  if (int e = encode_map(&g_tab_map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, g_tab, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B, "stft tables")) {
    cudaFree(g_tab);
    g_tab = nullptr;
    return e;
  }
  return SOS_OK;
}

struct StftParams {
  CUtensorMap tab;
  const float* wave;
  float* out;
  const uint8_t* bits;
  const int* frame_lo;
  int L, T, n_frames_total, nb, gate_mode;
  float inv_ratio;
};

__global__ void __launch_bounds__(kThreadsStft, 1) stft_tc_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_base = smem_base;                                       // [2 bufs][e_hi | e_lo | o_hi | o_lo]
  const uint32_t b_base = a_base + 2 * kABuf;                              // [3 slots][hi | lo]
  const uint32_t bar_base = b_base + kBSlots * kBSlot;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t tfull = bar_base + 8u * 12;
  const uint32_t tmem_slot = bar_base + 8u * 13;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tab);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), 32 * kBuilderWarps);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < kBSlots; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== table chunks (TMA)
    for (int i = 0; i < kStages; ++i) {
      const int slot = i % kBSlots, chunk = i >> 2, part = (i >> 1) & 1, nh = i & 1;
      if (i >= kBSlots) mbar_wait(b_empty(slot), ((i / kBSlots) - 1) & 1, 800);
      if (elect_one_sync()) {
        mbar_expect_tx(b_full(slot), kBSlot);
        const uint32_t bdst = b_base + (uint32_t)slot * kBSlot;
        tma_load_2d(bdst, &p.tab, b_full(slot), chunk * kChunk, (2 * part) * kBins + nh * kNHalf);               // hi
        tma_load_2d(bdst + kBTile, &p.tab, b_full(slot), chunk * kChunk, (2 * part + 1) * kBins + nh * kNHalf);  // lo
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_idesc_tf32(128, kNHalf, 0, 0);
    const uint32_t desc_hi = (uint32_t)((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t lbo_bits = (16u >> 4) << 16;
    for (int i = 0; i < kStages; ++i) {
      const int slot = i % kBSlots, chunk = i >> 2, part = (i >> 1) & 1, nh = i & 1, buf = chunk & 1;
      if ((i & 3) == 0) mbar_wait(a_full(buf), (chunk >> 1) & 1, 801);
      mbar_wait(b_full(slot), (i / kBSlots) & 1, 802);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t a0 = a_base + (uint32_t)buf * kABuf + (uint32_t)part * 2 * kATile, b0 = b_base + (uint32_t)slot * kBSlot;
        const uint32_t d = tmem_base + (uint32_t)part * kBins + (uint32_t)nh * kNHalf;
        const int nkk = chunk == kChunks - 1 ? (kHalf - (kChunks - 1) * kChunk) / 8 : kChunk / 8;   // offsets 192..199 only in the last chunk
        // hi*hi, lo*hi, hi*lo
        const uint32_t a_of[3] = {a0, a0 + kATile, a0}, b_of[3] = {b0, b0, b0 + kBTile};
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a_lo = (a_of[term] >> 4) | lbo_bits, b_lo = (b_of[term] >> 4) | lbo_bits;
          for (int kk = 0; kk < nkk; ++kk)
            umma_tf32(d, ((uint64_t)desc_hi << 32) | (a_lo + 2u * kk), ((uint64_t)desc_hi << 32) | (b_lo + 2u * kk), idesc,
                      (chunk | term | kk) != 0);
        }
        umma_commit(b_empty(slot));
        if ((i & 3) == 3) umma_commit(a_empty(buf));
        if (i == kStages - 1) umma_commit(tfull);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== frame builders, then the drain
    const int q = warp & 3;                              // TMEM lane quadrant of this warp
    const int bw = warp - 2;                             // builder warp 0..7
    const int L = p.L;
    // A builder warp takes one frame row at a time: lane j loads x[c + d0 + j] and x[c - d0 - j] (two coalesced 128-byte
    // reads), folds them into e / o, splits hi / lo and writes element j of the row into the four tiles of the chunk's buffer:
    // a warp store covers one whole 128-byte row of the swizzled tile (conflict-free).
    for (int chunk = 0; chunk < kChunks; ++chunk) {
      const int buf = chunk & 1;
      if (chunk >= 2) mbar_wait(a_empty(buf), ((chunk >> 1) - 1) & 1, 803);
      const uint32_t abuf = a_base + (uint32_t)buf * kABuf;
      const int d = chunk * kChunk + lane;
      const uint32_t col = ((uint32_t)(lane >> 2) << 4) | ((uint32_t)(lane & 3) << 2);   // byte offset of element `lane` before the swizzle
      constexpr int kRows = 8;                           // rows per batch: 16 independent loads in flight per lane
      for (int r0 = bw; r0 < 128; r0 += kBuilderWarps * kRows) {
        float xp[kRows], xm[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int f = blockIdx.x * 128 + r0 + kBuilderWarps * u;
          xp[u] = xm[u] = 0.f;
          if (f < p.n_frames_total && d < kHalf) {
            const int b = f / p.T, t = f - b * p.T;
            const float* x = p.wave + (size_t)b * L;
            const int c = t * kHop;
            int ip = c + d, im = c - d;
            if (ip >= L) ip = 2 * (L - 1) - ip;
            if (im < 0) im = -im;
            xp[u] = __ldg(x + ip);
            xm[u] = __ldg(x + im);
          }
        }
        if (p.gate_mode) {                               // the mask logic is branchy and large: one (non-inlined) copy
#pragma unroll
          for (int u = 0; u < kRows; ++u) {
            const int f = blockIdx.x * 128 + r0 + kBuilderWarps * u;
            if (f < p.n_frames_total && d < kHalf) {
              const int b = f / p.T, t = f - b * p.T;
              const int c = t * kHop;
              int ip = c + d, im = c - d;
              if (ip >= L) ip = 2 * (L - 1) - ip;
              if (im < 0) im = -im;
              const uint8_t* bb = p.bits + (size_t)b * p.nb;
              const float mp = sample_mask(ip, L, bb, p.nb, p.frame_lo, p.inv_ratio), mm = sample_mask(im, L, bb, p.nb, p.frame_lo, p.inv_ratio);
              xp[u] *= (p.gate_mode == 1) ? mp : (1.f - mp);
              xm[u] *= (p.gate_mode == 1) ? mm : (1.f - mm);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int r = r0 + kBuilderWarps * u;
          const float e = d == 0 ? xp[u] : xp[u] + xm[u];
          const float o = d == 0 ? 0.f : xp[u] - xm[u];
          const float e_hi = tf32_rna(e), o_hi = tf32_rna(o);
          const float e_lo = tf32_rna(e - e_hi), o_lo = tf32_rna(o - o_hi);
          // K-major 128-byte swizzle: 16-byte chunk index ^= row & 7
          const uint32_t off = abuf + (uint32_t)r * 128u + (col ^ ((uint32_t)(r & 7) << 4));
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(off), "f"(e_hi) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(off + kATile), "f"(e_lo) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(off + 2 * kATile), "f"(o_hi) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(off + 3 * kATile), "f"(o_lo) : "memory");
        }
      }
      fence_proxy_async_smem();                          // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(a_full(buf));
    }
    const int row = q * 32 + lane;
    const int f = blockIdx.x * 128 + row;                // frame index over (clip, frame)
    const bool valid = f < p.n_frames_total;
    const int b = valid ? f / p.T : 0, t = valid ? f - b * p.T : 0;
    const int col_lo = bw < 4 ? 0 : kBins;               // warps 2-5 drain the re columns, warps 6-9 the im columns
    // ---- drain: thread = frame row, registers = bins; a warp stores 32 consecutive frames of one bin
    mbar_wait(tfull, 0, 804);
    tc_fence_after();
    float* o = p.out + ((size_t)b * 2 * kBins) * p.T + t;
    for (int col0 = col_lo; col0 < col_lo + kBins; col0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
      tmem_ld16(taddr, r);
      tmem_ld16(taddr + 16, r + 16);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int u = 0; u < 32; ++u) o[(size_t)(col0 + u) * p.T] = __uint_as_float(r[u]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int sos_stft_tc_init() { return init_stft_tc(); }

int sos_stft_tc_launch(const float* wave, int64_t batch, int64_t length, float* spec_out, const uint8_t* bits, int64_t n_bits,
                       const int32_t* frame_lo, double ratio, int gate_mode, cudaStream_t stream) {
  if (int e = init_stft_tc()) return e;
  static thread_local StftParams p;
  p.tab = g_tab_map;
  p.wave = wave;
  p.out = spec_out;
  p.bits = gate_mode ? bits : nullptr;
  p.frame_lo = frame_lo;
  p.L = (int)length;
  p.T = 1 + (int)(length / kHop);
  SOS_CHECK_ARG(batch * (int64_t)p.T < (1ll << 31), "sos_stft_forward: too many frames");
  p.n_frames_total = (int)(batch * p.T);
  p.nb = (int)n_bits;
  p.gate_mode = gate_mode;
  p.inv_ratio = gate_mode ? (float)(1.0 / ratio) : 0.f;
  const int smem = 1024 + 2 * (int)kABuf + kBSlots * (int)kBSlot + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(stft_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      sos_set_error("sos_stft_forward: cannot raise dynamic shared memory to %d bytes: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return SOS_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = ceil_div(p.n_frames_total, 128);
  stft_tc_kernel<<<grid, kThreadsStft, smem, stream>>>(p);
  SOS_CHECK_LAUNCH("sos_stft_forward");
  return SOS_OK;
}
