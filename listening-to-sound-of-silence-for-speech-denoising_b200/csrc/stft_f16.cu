// STFT toward the HBM roofline: persistent, fully overlapped tensor-core DFT with IEEE-half split operands
// (reference: M2/transform.py:188-193 -> librosa.stft(y, 510, 158, 400); same folded real DFT as stft_tc.cu).
//
//     re[k] =  (-1)^k  sum_{d=0..199} w_d cos(2 pi k d / 510) e[d],   e[0] = x[c],  e[d] = x[c+d] + x[c-d]
//     im[k] = -(-1)^k  sum_{d=1..199} w_d sin(2 pi k d / 510) o[d],               o[d] = x[c+d] - x[c-d]
//
// Why a second kernel.  stft_tc.cu (one CTA per 128-frame tile, TF32 splits) spends 11.7 us per tile in the tensor pipe and does
// nothing else meanwhile: frame builders wait on global loads, the drain runs after the last MMA, 203 tiles on 148 SMs leave a
// 0.37 wave tail -- 12 % of the HBM roofline.  The transform moves 336 KB per tile (84 KB of samples in, 256 KB of spectrogram
// out): an SM's share of the measured 6.55 TB/s moves that in 7.6 us.  Here:
//   * operands are split into IEEE halves, x = hi + lo with hi = half(x), lo = half(x - hi): kind::f16 runs at twice the TF32 rate.
//     The folded samples are pre-scaled by 16 (exact) so that lo of audio-range samples stays in half's normal range; where it does
//     not (|x| < 0.02) lo is a half subnormal with an ABSOLUTE error <= 3e-8, below fp32's own rounding of the result.
//     hi*hi + lo*hi + hi*lo accumulate in fp32 in TMEM; the dropped lo*lo term is 2^-24;
//   * one radix-2 step over the bins (see kNb below) halves the multiply-adds and the table bytes streamed from L2;
//   * a tile runs its real part (TMEM columns 0..255: even | odd offsets), then its imaginary part (columns 256..511): the real part
//     drains while the imaginary MMAs run, the imaginary part while the NEXT tile's real MMAs run -- TMEM is double buffered by the
//     re / im alternation itself, and every frame element is folded and split exactly once;
//   * the tile's samples are staged ONCE in shared memory (cp.async, all in flight; gated there if requested: once per sample, not
//     once per frame and offset), the frame builders then fold / split out of shared memory with no global latency in their loop;
//   * CTAs are persistent over contiguous tile ranges.
// Warp roles (576 threads): warp 0 TMA (table chunks, 4-slot ring), warp 1 MMA issuer, warps 2-9 drain (thread = frame row, a warp
// stores 32 consecutive frames of one bin: coalesced in the (B, 2, 256, T) layout), warps 10-17 frame builders.
// Measured (scripts/bench_transforms.py, L2 flushed): 56 us for 128 signals (19 % of the HBM copy peak; stft_tc.cu: 74 us), 159 us for
// 512 (27 %; 211 us).  Ablations (SOS_STFT_DBG bits: 1 no builder arithmetic, 2 no drain stores, 4 no sample staging) show what is left:
// per tile the MMA + barrier skeleton takes 8.4 us (the N = 128 MMAs read 8 KB of operands per 64 tensor cycles: shared-memory bound),
// the drain's stores 6.8 us (2048 misaligned 128-byte warp stores per tile: T = 203 frames per row, ~6 LSU cycles each), the builders
// 5 us and the staging 3 us, and these ADD instead of overlapping: they share the SM's load/store pipe.  The next step is a TMA-store
// drain through a shared-memory staging tile (needs the 85 KB sample tile to shrink first).
#include "common.cuh"
#include "gate.cuh"
#include "ptx.cuh"
#include "sos_b200.h"
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace {

using namespace ptx;
using namespace tc;

constexpr int kHop = 158, kBins = 256, kHalf = 200;       // kHalf: offsets d = 0..199
// Bin symmetry (one radix-2 step): cos(2 pi (255-k) d / 510) = (-1)^d cos(2 pi k d / 510) (sin: -(-1)^d), so with the offsets split by
// parity, EV[k] = sum over even d, OD[k] = sum over odd d of the table rows of bin k < 128:
//     re[k] = EV + OD,  re[255-k] = OD - EV;     im[k] = EV' + OD',  im[255-k] = EV' - OD'
// -- only 128 table rows, half the multiply-adds and half the table bytes streamed from L2 per tile.
constexpr int kNb = 128;                                   // table rows (bins 0..127) = MMA N
constexpr int kChunkD = 64, kChunks = 4;                   // a chunk = 64 consecutive offsets = 32 K columns per parity (64-byte rows)
constexpr int kKpar = 32 * kChunks;                        // 128 K columns per parity (100 used)
constexpr int kBuilderWarps = 8, kDrainWarps = 8;
constexpr int kThreads = 32 * (2 + kDrainWarps + kBuilderWarps);    // 576
constexpr uint32_t kATile = 128 * 64;                      // 128 frames x 32 K columns, half
constexpr uint32_t kABuf = 4 * kATile;                     // even_hi | even_lo | odd_hi | odd_lo of one (part, chunk)
constexpr int kABufs = 2;
constexpr uint32_t kBTile = kNb * 64;                      // 128 bins x 32 K columns, half
constexpr uint32_t kBSlot = 2 * kBTile;                    // hi | lo of one (part, chunk, parity)
constexpr int kBSlots = 4;
constexpr int kMaxSeg = 4;                                 // clips a 128-frame tile may touch
constexpr int kXFloats = 127 * kHop + kMaxSeg * 408 + 16;  // staged samples of a tile
constexpr float kPreScale = 16.f;

__half* g_tab = nullptr;          // [4: cos_hi, cos_lo, sin_hi, sin_lo][2 parities][128 bins][128 K columns: offset d = 2 i + parity]
CUtensorMap g_tab_map;

int init_stft_f16() {
  if (g_tab) return SOS_OK;
  std::vector<__half> tab((size_t)4 * 2 * kNb * kKpar, __float2half_rn(0.f));
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < kNb; ++k) {
    const double sgn = (k & 1) ? -1.0 : 1.0;
    for (int d = 0; d < kHalf; ++d) {
      const double w = 0.5 + 0.5 * cos(two_pi * d / 400.0);                 // hann(400)[200 + d]
      const int ph = (int)(((long long)k * d) % 510);
      const double cv = sgn * w * cos(two_pi * ph / 510.0), sv = -sgn * w * sin(two_pi * ph / 510.0);
      const double vals[2] = {cv, sv};
      for (int m = 0; m < 2; ++m) {
        const __half hi = __float2half_rn((float)vals[m]);
        const __half lo = __float2half_rn((float)(vals[m] - (double)__half2float(hi)));
        tab[(((size_t)(2 * m) * 2 + (d & 1)) * kNb + k) * kKpar + (d >> 1)] = hi;
        tab[(((size_t)(2 * m + 1) * 2 + (d & 1)) * kNb + k) * kKpar + (d >> 1)] = lo;
      }
    }
  }
  if (cudaMalloc(&g_tab, tab.size() * sizeof(__half)) != cudaSuccess) {
    g_tab = nullptr;
    sos_set_error("stft: cudaMalloc of the half DFT tables failed");
    return SOS_ERR_CUDA;
  }
  cudaMemcpy(g_tab, tab.data(), tab.size() * sizeof(__half), cudaMemcpyHostToDevice);
  uint64_t dims[2] = {(uint64_t)kKpar, (uint64_t)8 * kNb};
  uint64_t str[2] = {2, (uint64_t)kKpar * 2};
  uint32_t box[2] = {32u, (uint32_t)kNb};
  uint32_t es[2] = {1, 1};
  if (int e = encode_map(&g_tab_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g_tab, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_64B, "stft half tables")) {
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
  int L, T, n_frames_total, n_tiles, nb, gate_mode;
  int tf;                          // frames per tile (<= 128, a multiple of 8): chosen so that the tiles fill whole waves of the grid
  int dbg;                         // ablation switches for measurements (SOS_STFT_DBG): 1 no builder arithmetic, 2 no drain stores, 4 no staging
  float inv_ratio;
};

// Samples of the tile's frames, per clip the tile touches: clip b0 + s covers samples [lo, hi) at x_tile + off.
struct Seg { int lo, hi, off; };
__device__ __forceinline__ void tile_segments(const StftParams& p, int tile, int& b0, int& nseg, Seg* seg) {
  const int f0 = tile * p.tf, f1 = min(p.n_frames_total, f0 + p.tf) - 1;
  b0 = f0 / p.T;
  const int b1 = f1 / p.T;
  nseg = min(b1 - b0 + 1, kMaxSeg);
  int off = 0;
  for (int s = 0; s < nseg; ++s) {
    const int b = b0 + s;
    const int ta = (b == b0) ? f0 - b * p.T : 0, tb = (b == b1) ? f1 - b * p.T : p.T - 1;
    // frame t reads x[c - 199 .. c + 199] around c = 158 t, reflected once at the clip ends (librosa center=True, pad_mode='reflect')
    int lo = ta * kHop - 201, hi = tb * kHop + 200;
    if (lo < 0) { hi = max(hi, 201); lo = 0; }
    if (hi > p.L) { lo = min(lo, p.L - 202); hi = p.L; }
    lo = max(lo, 0) & ~3;                                  // (16-byte aligned starts: vector loads when the clip base allows)
    seg[s].lo = lo;
    seg[s].hi = hi;
    seg[s].off = off;
    off += (hi - lo + 3) & ~3;
  }
}

__global__ void __launch_bounds__(kThreads, 1) stft_f16_kernel(const __grid_constant__ StftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_base = smem_base;                                       // [2 bufs][hi | lo]
  const uint32_t b_base = a_base + kABufs * kABuf;                         // [3 slots][hi | lo]
  const uint32_t x_base = b_base + kBSlots * kBSlot;                       // staged samples (fp32)
  const uint32_t bar_base = x_base + kXFloats * 4;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto t_full = [&](int s) { return bar_base + 8u * (12 + s); };
  auto t_empty = [&](int s) { return bar_base + 8u * (14 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* x_tile = reinterpret_cast<float*>(smem_raw + (x_base - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tab);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), kBuilderWarps);                 // one arrival per builder warp (256 arrivals on one barrier serialise)
      mbar_init(a_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kDrainWarps);
    }
    for (int s = 0; s < kBSlots; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
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

  // contiguous tile range of this CTA
  const int i0 = (int)((long long)blockIdx.x * p.n_tiles / gridDim.x), i1 = (int)((long long)(blockIdx.x + 1) * p.n_tiles / gridDim.x);

  if (warp == 0) {
    // ===================================================================== table chunks (TMA)
    int n = 0;
    for (int tile = i0; tile < i1; ++tile)
      for (int part = 0; part < 2; ++part)
        for (int chunk = 0; chunk < kChunks; ++chunk)
          for (int par = 0; par < 2; ++par, ++n) {
            const int slot = n % kBSlots;
            if (n >= kBSlots) mbar_wait(b_empty(slot), ((n / kBSlots) - 1) & 1, 900);
            if (elect_one_sync()) {
              mbar_expect_tx(b_full(slot), kBSlot);
              const uint32_t bdst = b_base + (uint32_t)slot * kBSlot;
              tma_load_2d(bdst, &p.tab, b_full(slot), chunk * 32, ((2 * part) * 2 + par) * kNb);               // hi
              tma_load_2d(bdst + kBTile, &p.tab, b_full(slot), chunk * 32, ((2 * part + 1) * 2 + par) * kNb);  // lo
            }
            __syncwarp();
          }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_idesc_f16(128, kNb, 0, 0);
    const uint32_t desc_hi = (uint32_t)((512u >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);   // SBO 512 (8 rows x 64 B), version 1, SWIZZLE_64B
    const uint32_t lbo_bits = (16u >> 4) << 16;
    int na = 0, nb = 0, ni = 0;
    for (int tile = i0; tile < i1; ++tile, ++ni) {
      for (int part = 0; part < 2; ++part) {
        if (ni >= 1) mbar_wait(t_empty(part), (ni - 1) & 1, 901);          // the previous tile's `part` columns have been drained
        tc_fence_after();
        for (int chunk = 0; chunk < kChunks; ++chunk, ++na) {
          const int abuf = na % kABufs;
          mbar_wait(a_full(abuf), (na / kABufs) & 1, 902);
          for (int par = 0; par < 2; ++par, ++nb) {
            const int slot = nb % kBSlots;
            mbar_wait(b_full(slot), (nb / kBSlots) & 1, 903);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t a0 = a_base + (uint32_t)abuf * kABuf + (uint32_t)par * 2 * kATile, b0 = b_base + (uint32_t)slot * kBSlot;
              const uint32_t d = tmem_base + (uint32_t)part * 256 + (uint32_t)par * kNb;
              const int nkk = chunk == kChunks - 1 ? 1 : 2;      // K columns 96..99 (+ zero padding to 112) only in the last chunk
              const uint32_t a_of[3] = {a0, a0 + kATile, a0}, b_of[3] = {b0, b0, b0 + kBTile};      // hi*hi, lo*hi, hi*lo
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const uint32_t a_lo = (a_of[term] >> 4) | lbo_bits, b_lo = (b_of[term] >> 4) | lbo_bits;
                for (int kk = 0; kk < nkk; ++kk)
                  umma_f16(d, ((uint64_t)desc_hi << 32) | (a_lo + 2u * kk), ((uint64_t)desc_hi << 32) | (b_lo + 2u * kk), idesc,
                           (chunk | term | kk) != 0);
              }
              umma_commit(b_empty(slot));
              if (par == 1) umma_commit(a_empty(abuf));
              if (par == 1 && chunk == kChunks - 1) umma_commit(t_full(part));
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 2 + kDrainWarps) {
    // ===================================================================== drain: thread = frame row, registers = bins
    const int q = warp & 3;                              // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;                      // which 64 of the 128 table bins this warp drains
    const int row = q * 32 + lane;
    int ni = 0;
    for (int tile = i0; tile < i1; ++tile, ++ni) {
      const int f = tile * p.tf + row;
      const bool in_tile = row < p.tf && f < p.n_frames_total;
      const bool valid = in_tile && !(p.dbg & 2);
      const int b = in_tile ? f / p.T : 0, t = in_tile ? f - b * p.T : 0;
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        mbar_wait(t_full(part), ni & 1, 904);
        tc_fence_after();
        float* o = p.out + ((size_t)(b * 2 + part) * kBins) * p.T + t;
        const float s_hi = part == 0 ? -(1.f / kPreScale) : (1.f / kPreScale);     // bin 255 - k: re = OD - EV, im = EV - OD
#pragma unroll 1
        for (int col0 = ch * 64; col0 < ch * 64 + 64; col0 += 16) {
          uint32_t ev[16], od[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 256 + col0);
          if (!(p.dbg & 8)) {
            tmem_ld16(taddr, ev);
            tmem_ld16(taddr + kNb, od);
            tmem_ld_wait();
          }
          if (valid) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float e = __uint_as_float(ev[u]), d = __uint_as_float(od[u]);
              __stcs(o + (size_t)(col0 + u) * p.T, (e + d) * (1.f / kPreScale));
              __stcs(o + (size_t)(kBins - 1 - col0 - u) * p.T, (e - d) * s_hi);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(part));
      }
    }
  } else {
    // ===================================================================== frame builders
    const int bw = warp - (2 + kDrainWarps);             // 0..7: this warp owns rows bw, bw + 8, ... (16 rows)
    const int bt = threadIdx.x - 32 * (2 + kDrainWarps); // 0..255
    const int L = p.L;
    int n = 0;
    for (int tile = i0; tile < i1; ++tile) {
      // ---- stage the tile's samples once (pre-scaled; gated here when requested).  The previous tile's last chunk has been built by
      //      every builder thread before anyone overwrites the samples.
      asm volatile("bar.sync 2, 256;" ::: "memory");
      int b0, nseg;
      Seg seg[kMaxSeg];
      tile_segments(p, tile, b0, nseg, seg);
      for (int s = 0; s < ((p.dbg & 4) ? 0 : nseg); ++s) {
        const float* x = p.wave + (size_t)(b0 + s) * L + seg[s].lo;
        const int cnt = seg[s].hi - seg[s].lo;
        float* dst = x_tile + seg[s].off;
        if (p.gate_mode == 0 && ((uintptr_t)x & 15) == 0) {
          // plain copy, 16 bytes per cp.async, everything in flight at once (the pre-scale is applied when the samples are folded)
          const int n4 = cnt >> 2;
          const uint32_t d32 = x_base + (uint32_t)seg[s].off * 4u;
          for (int i = bt; i < n4; i += 256)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + 16u * i), "l"(__cvta_generic_to_global(x + 4 * i)) : "memory");
          for (int i = 4 * n4 + bt; i < cnt; i += 256) dst[i] = __ldg(x + i);
        } else {
          const uint8_t* bb = p.gate_mode ? p.bits + (size_t)(b0 + s) * p.nb : nullptr;
          for (int i = bt; i < cnt; i += 256 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (i + u * 256 < cnt) v[u] = __ldg(x + i + u * 256);
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (i + u * 256 < cnt) {
                float m = 1.f;
                if (p.gate_mode) {
                  m = sample_mask(seg[s].lo + i + u * 256, L, bb, p.nb, p.frame_lo, p.inv_ratio);
                  if (p.gate_mode != 1) m = 1.f - m;
                }
                dst[i + u * 256] = v[u] * m;
              }
          }
        }
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      // per-row constants of this warp's 16 rows: index of x[c] in the staged samples (-1: no such frame), and of x[0] of its clip
      int rc[16], rz[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int f = tile * p.tf + bw + 8 * j;
        rc[j] = -1;
        rz[j] = 0;
        if (bw + 8 * j < p.tf && f < p.n_frames_total) {
          const int b = f / p.T, t = f - b * p.T;
          Seg sg = seg[0];
#pragma unroll
          for (int s = 1; s < kMaxSeg; ++s)
            if (s == b - b0) sg = seg[s];
          rz[j] = sg.off - sg.lo;
          rc[j] = t * kHop;
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      for (int part = 0; part < 2; ++part)
        for (int chunk = 0; chunk < kChunks; ++chunk, ++n) {
          const int abuf = n % kABufs;
          if (n >= kABufs) mbar_wait(a_empty(abuf), ((n / kABufs) - 1) & 1, 905);
          const uint32_t ab = a_base + (uint32_t)abuf * kABuf;
          if (!(p.dbg & 1)) {
#pragma unroll 1
            for (int h = 0; h < (chunk == kChunks - 1 ? 1 : 2); ++h) {       // offsets 64 chunk + 32 h + lane (the last chunk's MMAs read 192..223 only)
              const int d = chunk * kChunkD + h * 32 + lane;
              const bool d_ok = d < kHalf && (part == 0 || d > 0);
              // element d goes to the parity tile d & 1, K column (d >> 1) & 31: byte (col >> 3) ^ swz chunk, (col & 7) * 2
              const int col = (d >> 1) & 31;
              const uint32_t dst = ab + (uint32_t)(d & 1) * 2 * kATile + (uint32_t)bw * 64u +
                                   ((((uint32_t)(col >> 3)) ^ ((uint32_t)(bw >> 1) & 3u)) << 4 | (uint32_t)(col & 7) * 2u);
              // v = 16 (x[c+d] +- x[c-d]) (16 x[c] for d = 0): branch-free, loads of all 16 rows first
              const float sg = part == 0 ? (d == 0 ? 0.f : kPreScale) : -kPreScale;
              float xp[16], xm[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c = rc[j];
                const bool ok = d_ok && c >= 0;
                int ip = c + d;
                ip = ip >= L ? 2 * (L - 1) - ip : ip;
                const int im = abs(c - d);
                xp[j] = x_tile[ok ? rz[j] + ip : 0];
                xm[j] = x_tile[ok ? rz[j] + im : 0];
              }
              // split, then pair up K columns: offsets d and d + 2 are neighbouring columns of the same parity tile, so lane d takes
              // lane d + 2's halves (one shuffle of the packed hi|lo word) and the lanes with d % 4 < 2 store 32-bit words -- half the
              // shared-memory wavefronts of 16-bit stores
              uint32_t hl[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const bool ok = d_ok && rc[j] >= 0;
                const float v = ok ? fmaf(sg, xm[j], kPreScale * xp[j]) : 0.f;
                const __half hh = __float2half_rn(v);
                hl[j] = (uint32_t)__half_as_ushort(hh) | ((uint32_t)__half_as_ushort(__float2half_rn(v - __half2float(hh))) << 16);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t nx = __shfl_down_sync(0xffffffffu, hl[j], 2);
                if ((lane & 2) == 0) {
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 512u * j), "r"(__byte_perm(hl[j], nx, 0x5410)));            // hi_d | hi_{d+2}
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + kATile + 512u * j), "r"(__byte_perm(hl[j], nx, 0x7632)));  // lo_d | lo_{d+2}
                }
              }
            }
          }
          fence_proxy_async_smem();                          // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(abuf));
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

int sos_stft_f16_init() { return init_stft_f16(); }

// Returns SOS_ERR_UNSUPPORTED (without setting an error) when the clip is too short for this kernel's tiling: the caller then
// runs the TF32 kernel of stft_tc.cu.
int sos_stft_f16_launch(const float* wave, int64_t batch, int64_t length, float* spec_out, const uint8_t* bits, int64_t n_bits,
                        const int32_t* frame_lo, double ratio, int gate_mode, cudaStream_t stream) {
  const int T = 1 + (int)(length / kHop);
  if (128 / T + 2 > kMaxSeg || length < 512) return SOS_ERR_UNSUPPORTED;
  if (int e = init_stft_f16()) return e;
  static thread_local StftParams p;
  p.tab = g_tab_map;
  p.wave = wave;
  p.out = spec_out;
  p.bits = gate_mode ? bits : nullptr;
  p.frame_lo = frame_lo;
  p.L = (int)length;
  p.T = T;
  SOS_CHECK_ARG(batch * (int64_t)p.T < (1ll << 30), "sos_stft_forward: too many frames");
  p.n_frames_total = (int)(batch * p.T);
  {
    // 128-frame tiles leave a partial last wave (203 tiles on 148 SMs at 128 signals: the grid waits for the 55 CTAs with two
    // tiles); shorter tiles in whole waves keep every SM equally busy -- an MMA over unused rows costs less than an idle SM
    const int sms = sos_num_sms();
    const int waves = ceil_div(ceil_div(p.n_frames_total, 128), sms);
    static const int fixed = getenv("SOS_STFT_TF") ? atoi(getenv("SOS_STFT_TF")) : 0;          // A/B aid
    p.tf = fixed > 0 ? fixed : std::min(128, round_up(ceil_div(p.n_frames_total, waves * sms), 8));
    p.tf = std::max(8, std::min(128, p.tf & ~7));
  }
  p.n_tiles = ceil_div(p.n_frames_total, p.tf);
  p.nb = (int)n_bits;
  p.gate_mode = gate_mode;
  p.dbg = getenv("SOS_STFT_DBG") ? atoi(getenv("SOS_STFT_DBG")) : 0;
  p.inv_ratio = gate_mode ? (float)(1.0 / ratio) : 0.f;
  const int smem = 1024 + kABufs * (int)kABuf + kBSlots * (int)kBSlot + kXFloats * 4 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(stft_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      sos_set_error("sos_stft_forward: cannot raise dynamic shared memory to %d bytes: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return SOS_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = std::min(p.n_tiles, sos_num_sms());
  stft_f16_kernel<<<grid, kThreads, smem, stream>>>(p);
  SOS_CHECK_LAUNCH("sos_stft_forward");
  return SOS_OK;
}
