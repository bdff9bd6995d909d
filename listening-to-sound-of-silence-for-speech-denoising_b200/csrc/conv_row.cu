// Row-streaming stacked-tap convolution on tcgen05 for the narrow (48-channel) dilated 5x5 layers of the silent-interval detector
// and of ContextAggNet's encoder_n (M1/networks.py:91-93, M2/networks.py:61-62,72-80; 21 of a training step's convolutions,
// forward and data gradient).
//
// Why a second kernel.  In the tap-list GEMM (conv_tc.cu) every tap is its own MMA with N = Cout = 48: 24 tensor cycles per
// instruction that reads 4 KB of A + 1.5 KB of B from shared memory (44 cycles at 128 B/clk) -- ncu r02: the 48-channel layers
// sit at 93 % shared-memory pipe utilisation and ~700 TFLOP/s.  Here the taps of one kernel ROW (the KW taps along W) share ONE
// A operand: a CTA streams along W for a fixed 128-pixel block of H.  For input column i (pixels h0 .. h0+127 of column w_i,
// staged once with its H halo) and H-tap f the MMA
//
//     D[128 px, (b, co)] += A_i[px + f*dh, ci] . W[f, b][ci, co]          N = KW * Cout = 240
//
// adds column i's contribution to the KW output columns o = i - half .. i + half at once: their accumulators are adjacent
// slots (Cout TMEM columns each) of a ring over TMEM, slot = output counter mod R.  One MMA of N = 240 (120 cycles) reads the
// same 4 KB of A as an N = 48 one: A traffic per FLOP drops 5x and the layer becomes tensor bound.  A column's box is loaded
// once per stream (not once per tile and fast offset: L2 -> smem traffic falls ~10x) and all KH * KW weight tiles stay
// resident in shared memory (150 KB).  Dilation along W is a lattice: a stream visits columns phi, phi + dw, ...; dilation along
// H is a row offset of f * dh inside the staged box: an operand start address that is NOT aligned to the 1024-byte swizzle atom.
// That works with a zero descriptor base offset -- the tensor core XORs the same absolute shared-memory address bits as TMA did
// when it wrote the box (checked on the GPU: setting the base-offset field to (addr >> 7) & 7 gives wrong products).
//
// A slot's first contribution must overwrite (accumulate = 0): on a stream's first column every slot is fresh, afterwards only
// the newest one (o = i + half), which gets its own N = Cout MMA for the first (f, k-step).  Ranges that wrap around the ring
// end are issued as two MMAs.
//
// CTA = 352 threads: warp 0 TMA producer, warps 1 and 6 MMA issuers (one elected thread each, alternating input columns), warps
// 2-5 / 7-10 two epilogue groups (even /
// odd output columns): tcgen05.ld -> BatchNorm statistics | scale / affine / activation -> half -> dense staging -> TMA store.
#include "common.cuh"
#include "ptx.cuh"
#include "sos_b200.h"
#include "tc_common.cuh"
#include <stdlib.h>
#include <map>
#include <mutex>

namespace {

using namespace ptx;
using namespace tc;

constexpr int kThreadsRow = 352;
constexpr int kRowEpiWarps = 8;
constexpr int kMaxTaps = 25;
constexpr int kRing = 10;                  // accumulator slots of Cout = 48 TMEM columns (even: output parity == epilogue group)

struct alignas(64) RowParams {
  CUtensorMap mapA, mapB, mapD;
  int n_items, n_seg, seg_len, n_hb, dwl, W;     // dwl: lattice step along W (the dilation there)
  int KH, KW, half_w, Cin, Cout, cstore;
  int dh_bytes;                  // bytes between the A operands of consecutive H taps: dh * 128
  int h0_off;                    // H coordinate of a box's first row relative to the tile: -(KH-1)/2 * dh
  int box_rows, box_bytes, stage_bytes, n_stages;
  int b_tile_bytes;              // one (f, b) weight tile: Cout x 128 B
  int R;                         // ring slots
  int ksteps;                    // K steps of 16 channels
  int n_issuers;                 // 2, or 1 (A/B switch SOS_ROW_ISSUERS=1)
  uint32_t idesc[8];             // instruction descriptor for N = nb * Cout, nb = 1..KW
  int16_t tapsel[kMaxTaps];      // weight tap index (column block of wk) of smem tile (f, b)
  const float* scale;
  const float* shift;
  int act;
  const float* slope;
  const float* out_scale;
  float* stats;
  int stats_c, n_out;
  // fused BatchNorm-backward reduction of the layer below (sos_conv_args::bnr_*): its raw conv output and coefficients
  const uint8_t* bnr_y;
  long long bnr_sh, bnr_sw, bnr_sn;      // byte strides of bnr_y along H, W, image
  const float *bnr_scale, *bnr_shift, *bnr_mean, *bnr_invstd;
  float* bnr_partial;
  int bnr_c;
};

// shared memory besides the A stages: alignment slack, resident weight tiles, two staging buffers, barriers (512 B), statistics
// ([8 warps][2][Cout] forward sums, or -- with the fused BatchNorm-backward reduction -- [8][3][Cout] + the layer below's coefficient vectors)
__host__ __device__ inline int row_fixed_smem(int n_taps, int Cout, int cstore, bool bnr = false) {
  return 1024 + n_taps * Cout * 128 + 2 * 128 * cstore * 2 + 512 + kRowEpiWarps * (bnr ? 3 : 2) * Cout * 4 + (bnr ? 4 * Cout * 4 : 0);
}

struct RowItem { int n, hb, phi, k0, k1, i0, i1, L; };
__device__ __forceinline__ RowItem decode_row_item(const RowParams& p, int it) {
  RowItem r;
  const int seg = it % p.n_seg; it /= p.n_seg;
  r.phi = it % p.dwl; it /= p.dwl;
  r.hb = it % p.n_hb;
  r.n = it / p.n_hb;
  r.L = (p.W - r.phi + p.dwl - 1) / p.dwl;            // lattice columns of this phase
  r.k0 = seg * p.seg_len;
  r.k1 = min(r.L, r.k0 + p.seg_len);
  r.i0 = max(0, r.k0 - p.half_w);
  r.i1 = min(r.L, r.k1 + p.half_w);
  return r;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kThreadsRow, 1) rowconv_f16_kernel(const __grid_constant__ RowParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_taps = p.KH * p.KW;

  // smem carve-up: [weights: n_taps tiles][A stages][staging: 2 x (128 rows x Cout halves)][barriers][stats]
  const uint32_t b_base = smem_base;
  const uint32_t stages_base = b_base + (uint32_t)n_taps * p.b_tile_bytes;
  const uint32_t staging_base = stages_base + (uint32_t)p.n_stages * p.stage_bytes;
  const uint32_t stg_bytes = 128u * (uint32_t)p.cstore * 2u;
  const uint32_t bar_base = staging_base + 2u * stg_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (32 + s); };
  const uint32_t bfull_bar = bar_base + 8u * 48;
  auto turn_bar = [&](int t) { return bar_base + 8u * (50 + t); };
  const uint32_t tmem_slot = bar_base + 8u * 49;
  float* stats_s = reinterpret_cast<float*>(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    prefetch_tmap(&p.mapD);
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < p.R; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    mbar_init(bfull_bar, 1);
    mbar_init(turn_bar(0), 1);
    mbar_init(turn_bar(1), 1);
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
    // ===================================================================== TMA producer
    if (elect_one_sync()) {
      // all weight tiles, once: smem tile (f, b) <- column block tapsel[f*KW + b] of wk (Cout rows x 64 channels, tail zero-filled)
      mbar_expect_tx(bfull_bar, (uint32_t)n_taps * p.b_tile_bytes);
      for (int t = 0; t < n_taps; ++t) tma_load_2d(b_base + (uint32_t)t * p.b_tile_bytes, &p.mapB, bfull_bar, p.tapsel[t] * p.Cin, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const RowItem r = decode_row_item(p, it);
        if (r.k0 >= r.k1) continue;
        for (int i = r.i0; i < r.i1; ++i) {
          mbar_wait(empty_bar(stage), phase ^ 1, 100);
          mbar_expect_tx(full_bar(stage), (uint32_t)p.box_bytes);
          tma_load_4d(stages_base + (uint32_t)stage * p.stage_bytes, &p.mapA, full_bar(stage), 0, r.hb * 128 + p.h0_off, r.phi + p.dwl * i, r.n);
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ===================================================================== MMA issuers
    // One elected thread.  Per input column it first works out the (at most three) accumulator ranges ("pieces") its MMAs write --
    // slot, width, weight block, overwrite flag -- and then issues KH * ksteps * pieces MMAs with two adds each; the ring size is
    // a compile-time constant so that slot / phase arithmetic has no integer division (the first version of this loop spent ~390
    // cycles of address arithmetic per 120-cycle MMA).
    // Two issuers take the input columns in turn (warp 1 the even, warp 6 the odd ones of this CTA's column sequence): one thread
    // needs ~1500 cycles of barrier polls and bookkeeping per column on top of ~1700 to issue its MMAs, during which the tensor
    // pipe (1800 cycles of work per column) would drain.  The MMAs of consecutive columns accumulate into the same TMEM slots and
    // must enter the tensor pipe in column order: a thread issues only after the other one has issued the previous column (`turn`
    // barriers); each commit then also covers the other thread's earlier MMAs, because the pipe retires in issue order.
    const int T = warp == 1 ? 0 : 1;
    if (T < p.n_issuers && elect_one_sync()) {
      const int n_iss = p.n_issuers;
      int g = 0;                                      // index of the input column in this CTA's sequence
      uint32_t turn_phase = 0;
      int c_base = 0;                                 // output counter of the current stream's first output column
      constexpr int R = kRing;
      const int Cout = p.Cout, KW = p.KW, KH = p.KH, half_w = p.half_w, ksteps = p.ksteps;
      const uint32_t desc_hi = (uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);     // SBO 1024, version 1, SWIZZLE_128B
      const uint32_t lbo_bits = (16u >> 4) << 16;
      const uint32_t a_lo0 = (stages_base >> 4) | lbo_bits, b_lo0 = (b_base >> 4) | lbo_bits;
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, dh16 = (uint32_t)p.dh_bytes >> 4, tile16 = (uint32_t)p.b_tile_bytes >> 4;
      const uint32_t frow16 = (uint32_t)KW * tile16;                          // weight tiles of one H tap
      const uint32_t idesc0 = make_idesc_f16(128, 0, 0, 0), idstep = ((uint32_t)Cout >> 3) << 17;     // instruction descriptor of N = n * Cout
      mbar_wait(bfull_bar, 0, 199);
      tc_fence_after();
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const RowItem r = decode_row_item(p, it);
        if (r.k0 >= r.k1) continue;
        for (int i = r.i0; i < r.i1; ++i, ++g) {
          if (n_iss == 2 && (g & 1) != T) continue;
          const int stage = g % p.n_stages;
          const uint32_t phase = (uint32_t)(g / p.n_stages) & 1u;
          // outputs fed by this input column, as lattice indices [oa, ob]; their weight blocks b = o - (i - half)
          const int oa = max(r.k0, i - half_w), ob = min(r.k1 - 1, i + half_w);
          const bool all_fresh = i == r.i0;
          const bool one_fresh = !all_fresh && (i + half_w <= r.k1 - 1);          // the newest output o = ob = i + half starts here
          // The outputs [oa, ob] occupy consecutive ring slots from s0 on: one MMA of N = len * Cout, or two where the ring wraps.
          // On the first (f, k step) fresh slots are overwritten: all of them on a stream's first column, else only the newest
          // output, which then gets its own N = Cout MMA.
          const int len = ob - oa + 1;
          const int c_oa = c_base + (oa - r.k0);
          const int s0 = c_oa % R;
          const int n1 = min(len, R - s0), n2 = len - n1;
          const uint32_t bq = (uint32_t)(oa - (i - half_w)) * tile16;
          const uint32_t d1 = tmem_base + (uint32_t)(s0 * Cout), d2 = tmem_base;
          const uint32_t id1 = idesc0 + (uint32_t)n1 * idstep, id2 = idesc0 + (uint32_t)n2 * idstep;
          const uint32_t bq2 = bq + (uint32_t)n1 * tile16;
          // fresh slots: wait until the epilogue has drained their previous occupants
          if (all_fresh || one_fresh) {
            for (int o = all_fresh ? oa : ob; o <= ob; ++o) {
              const int c = c_base + (o - r.k0);
              mbar_wait(tempty_bar(c % R), (((uint32_t)(c / R)) & 1u) ^ 1u, 200);
            }
            tc_fence_after();
          }
          mbar_wait(full_bar(stage), phase, 201);
          if (n_iss == 2) {                                   // my turn: the other issuer has issued the previous column
            mbar_wait(turn_bar(T), turn_phase ^ (T == 0 ? 1u : 0u), 202);
            turn_phase ^= 1u;
          }
          tc_fence_after();
          const uint32_t a_col = a_lo0 + (uint32_t)stage * stage16;
          {
            // first (f, k step)
            const uint64_t ad = ((uint64_t)desc_hi << 32) | a_col;
            if (one_fresh) {
              const int lo = len - 1;                                               // old outputs
              const int m1 = min(lo, R - s0), m2 = lo - m1;
              if (m1 > 0) umma_f16(d1, ad, ((uint64_t)desc_hi << 32) | (b_lo0 + bq), idesc0 + (uint32_t)m1 * idstep, 1u);
              if (m2 > 0) umma_f16(d2, ad, ((uint64_t)desc_hi << 32) | (b_lo0 + bq + (uint32_t)m1 * tile16), idesc0 + (uint32_t)m2 * idstep, 1u);
              const int sf = (s0 + lo) % R;
              umma_f16(tmem_base + (uint32_t)(sf * Cout), ad, ((uint64_t)desc_hi << 32) | (b_lo0 + bq + (uint32_t)lo * tile16), idesc0 + idstep, 0u);
            } else {
              const uint32_t acc = all_fresh ? 0u : 1u;
              umma_f16(d1, ad, ((uint64_t)desc_hi << 32) | (b_lo0 + bq), id1, acc);
              if (n2 > 0) umma_f16(d2, ad, ((uint64_t)desc_hi << 32) | (b_lo0 + bq2), id2, acc);
            }
          }
          for (int f = 0; f < KH; ++f) {
            const uint32_t a_f = a_col + (uint32_t)f * dh16, b_f = b_lo0 + (uint32_t)f * frow16;
#pragma unroll 4
            for (int kk = f == 0 ? 1 : 0; kk < ksteps; ++kk) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_f + 2u * kk);
              const uint32_t b_k = b_f + 2u * kk;
              umma_f16(d1, ad, ((uint64_t)desc_hi << 32) | (b_k + bq), id1, 1u);
              if (n2 > 0) umma_f16(d2, ad, ((uint64_t)desc_hi << 32) | (b_k + bq2), id2, 1u);
            }
          }
          umma_commit(empty_bar(stage));                                            // frees the A stage when these MMAs retire
          // outputs that have now received their last contribution
          {
            const bool last_in = i == r.i1 - 1;
            const int da = max(i - half_w, r.k0), db = last_in ? r.k1 - 1 : i - half_w;
            for (int o = da; o <= db; ++o) umma_commit(tfull_bar((c_base + (o - r.k0)) % R));
          }
          if (n_iss == 2) mbar_arrive(turn_bar(1 - T));
        }
        c_base += r.k1 - r.k0;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5 = group 0, 7..10 = group 1)
    const int eg = warp >= 7 ? 1 : 0;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                // accumulator row = pixel h0 + row
    const int ethread = threadIdx.x - (eg ? 224 : 64);
    const int ewarp = eg * 4 + (ethread >> 5);
    const int bar_id = 1 + eg;
    const int Cout = p.Cout;
    constexpr int R = kRing;
    const float slope = ((p.act & SOS_ACT_MASK) == 2 && p.slope) ? *p.slope : 0.f;
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    const bool epi_math = p.shift || p.act || p.out_scale;
    // per-warp running sums: [2][Cout] forward statistics, or [3][Cout] of the fused BatchNorm-backward reduction
    const bool bnr = p.bnr_partial != nullptr;
    const int sstride = (bnr ? 3 : 2) * Cout;
    float* my_stats = stats_s + ewarp * sstride;
    float* coef_s = stats_s + kRowEpiWarps * sstride;            // [4][Cout]: scale, shift, mean, invstd of the layer below
    if (p.stats || bnr) {
      for (int i = lane; i < sstride; i += 32) my_stats[i] = 0.f;
      __syncwarp();
    }
    if (bnr) {
      for (int i = ethread; i < Cout; i += 128) {
        coef_s[i] = __ldg(p.bnr_scale + i);
        coef_s[Cout + i] = __ldg(p.bnr_shift + i);
        coef_s[2 * Cout + i] = __ldg(p.bnr_mean + i);
        coef_s[3 * Cout + i] = __ldg(p.bnr_invstd + i);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
    const uint32_t sbuf = staging_base + (uint32_t)eg * stg_bytes;
    const uint32_t srow = sbuf + (uint32_t)row * (uint32_t)(p.cstore * 2);
    constexpr int n_ec = 3;                        // Cout = 48 (sos_rowconv_eligible)
    int c_base = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const RowItem r = decode_row_item(p, it);
      if (r.k0 >= r.k1) continue;
      bool have_next = false;
      uint4 yn[6];
      for (int k = r.k0; k < r.k1; ++k) {
        const int c = c_base + (k - r.k0);
        if ((c & 1) != eg) continue;                                   // (R is even: a slot always belongs to the same group)
        const int slot = c % R;
        // (fused reduction) this thread's pixel of the layer below's raw output, 6 x 16 bytes: loaded one column of this group ahead
        // (k + 2), so that the ~1.5 us of an HBM read hide behind the column in between
        uint4 yq[6];
        if (bnr) {
          const uint8_t* ybase = p.bnr_y + (long long)r.n * p.bnr_sn + (long long)(r.hb * 128 + row) * p.bnr_sh;
          if (have_next) {
#pragma unroll
            for (int j = 0; j < 6; ++j) yq[j] = yn[j];
          } else {
            const uint4* yp = reinterpret_cast<const uint4*>(ybase + (long long)(r.phi + p.dwl * k) * p.bnr_sw);
#pragma unroll
            for (int j = 0; j < 6; ++j) yq[j] = __ldg(yp + j);
          }
          have_next = k + 2 < r.k1;
          if (have_next) {
            const uint4* yp = reinterpret_cast<const uint4*>(ybase + (long long)(r.phi + p.dwl * (k + 2)) * p.bnr_sw);
#pragma unroll
            for (int j = 0; j < 6; ++j) yn[j] = __ldg(yp + j);
          }
        }
        mbar_wait(tfull_bar(slot), ((uint32_t)(c / R)) & 1u, 300);
        tc_fence_after();
        // the staging buffer is free once the previous store of this group has read it
        if (ethread < 32 && elect_one_sync()) bulk_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int cc = 0; cc < n_ec; ++cc) {
          uint32_t rg[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * Cout + cc * 16), rg);
          tmem_ld_wait();
          if (p.stats) {
            // 16 channels: 15 shuffles per quantity bring lane i the sum of channel (i & 15) over its half-warp's rows, one more
            // adds the two half-warps
            float v[16], w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = __uint_as_float(rg[i]);
              w[i] = v[i] * v[i];
            }
#pragma unroll
            for (int st = 8; st >= 1; st >>= 1) {
              const bool up = (lane & st) != 0;
#pragma unroll
              for (int i = 0; i < st; ++i) {
                const float send_v = up ? v[i] : v[i + st], keep_v = up ? v[i + st] : v[i];
                const float send_w = up ? w[i] : w[i + st], keep_w = up ? w[i + st] : w[i];
                v[i] = keep_v + __shfl_xor_sync(0xffffffffu, send_v, st);
                w[i] = keep_w + __shfl_xor_sync(0xffffffffu, send_w, st);
              }
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
            w[0] += __shfl_xor_sync(0xffffffffu, w[0], 16);
            if (lane < 16) {
              my_stats[cc * 16 + lane] += v[0];
              my_stats[Cout + cc * 16 + lane] += w[0];
            }
          }
          if (epi_math) {
            const int act = p.act & SOS_ACT_MASK;
            const int ch0 = cc * 16;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rg[i]) * oscale;
            if (p.shift) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int cj = ch0 + i;
                const float b = cj < p.n_out ? __ldg(p.shift + cj) : 0.f;
                const float a = (p.scale && cj < p.n_out) ? __ldg(p.scale + cj) : 1.f;
                v[i] = fmaf(v[i], a, b);
              }
            }
            if (act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (act == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;
            } else if (act == 3) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 1.f / (1.f + expf(-v[i]));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) rg[i] = __float_as_uint(v[i]);
          }
          if (bnr) {
            // g = the value as stored (half) where the layer below was active; sums of g, g * xhat, g^2 per channel: the same
            // transpose-reduce as the forward statistics, three quantities
            float v[16], w[16], u[16];
            const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yq[2 * cc]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ch = cc * 16 + i;
              const __half2 hy = *reinterpret_cast<const __half2*>(&yw[i >> 1]);
              const float yv = (i & 1) ? __high2float(hy) : __low2float(hy);
              const float g0 = __half2float(__float2half_rn(__uint_as_float(rg[i])));
              const float pre = fmaf(yv, coef_s[ch], coef_s[Cout + ch]);
              const float g = pre > 0.f ? g0 : 0.f;
              v[i] = g;
              w[i] = g * ((yv - coef_s[2 * Cout + ch]) * coef_s[3 * Cout + ch]);
              u[i] = g * g;
            }
#pragma unroll
            for (int st = 8; st >= 1; st >>= 1) {
              const bool up = (lane & st) != 0;
#pragma unroll
              for (int i = 0; i < st; ++i) {
                const float sv = up ? v[i] : v[i + st], kv = up ? v[i + st] : v[i];
                const float sw = up ? w[i] : w[i + st], kw = up ? w[i + st] : w[i];
                const float su = up ? u[i] : u[i + st], ku = up ? u[i + st] : u[i];
                v[i] = kv + __shfl_xor_sync(0xffffffffu, sv, st);
                w[i] = kw + __shfl_xor_sync(0xffffffffu, sw, st);
                u[i] = ku + __shfl_xor_sync(0xffffffffu, su, st);
              }
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
            w[0] += __shfl_xor_sync(0xffffffffu, w[0], 16);
            u[0] += __shfl_xor_sync(0xffffffffu, u[0], 16);
            if (lane < 16) {
              my_stats[cc * 16 + lane] += v[0];
              my_stats[Cout + cc * 16 + lane] += w[0];
              my_stats[2 * Cout + cc * 16 + lane] += u[0];
            }
          }
          // dense staging row of cstore halves (a 96-byte pitch: two-way bank conflicts between rows r and r + 4)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (cc * 16 + 8 * j < p.cstore) {
              uint32_t h[4];
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) h[kx] = pack_half2(__uint_as_float(rg[8 * j + 2 * kx]), __uint_as_float(rg[8 * j + 2 * kx + 1]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)((cc * 2 + j) << 4)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                           : "memory");
            }
          }
        }
        // the accumulator slot is free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(slot));
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (ethread < 32 && elect_one_sync()) {
          tma_store_4d(&p.mapD, sbuf, 0, r.hb * 128, r.phi + p.dwl * k, r.n);
          bulk_commit();
        }
      }
      c_base += r.k1 - r.k0;
    }
    if (ethread < 32 && elect_one_sync()) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    // one row of BatchNorm partial sums per CTA: the eight epilogue warps' running sums, added up in a fixed order
    const int N = p.Cout;
    float* dst = p.stats + (size_t)blockIdx.x * 2 * p.stats_c;
    for (int c = threadIdx.x; c < 2 * p.stats_c; c += blockDim.x) {
      const int half = c >= p.stats_c ? 1 : 0, ch = c - half * p.stats_c;
      float sum = 0.f;
      if (ch < N)
        for (int w = 0; w < kRowEpiWarps; ++w) sum += stats_s[w * 2 * N + half * N + ch];
      dst[c] = sum;
    }
  }
  if (p.bnr_partial) {
    // one row [4][bnr_c] per CTA: sum g, sum g * xhat, 0 (PReLU slope term: the chains are ReLU), sum g^2
    const int N = p.Cout;
    float* dst = p.bnr_partial + (size_t)blockIdx.x * 4 * p.bnr_c;
    for (int c = threadIdx.x; c < 4 * p.bnr_c; c += blockDim.x) {
      const int qd = c / p.bnr_c, ch = c - qd * p.bnr_c;
      float sum = 0.f;
      if (ch < N && qd != 2) {
        const int src = qd == 3 ? 2 : qd;
        for (int w = 0; w < kRowEpiWarps; ++w) sum += stats_s[w * 3 * N + src * N + ch];
      }
      dst[c] = sum;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct RowPlan {
  RowParams p;
  MapSpec specA, specB, specD;
  const void *baseA = nullptr, *baseB = nullptr, *baseD = nullptr;
  long long d_offset = 0;
  int smem = 0, grid = 0;
};
std::mutex g_row_mutex;
std::map<std::vector<int32_t>, RowPlan*> g_row_plans;

}  // namespace

// Is this call one the row-streaming kernel serves?  (Pure shape logic; conv_tc.cu asks before it plans a tap-list GEMM.)
bool sos_rowconv_eligible(const sos_conv_args& a) {
  static const int off = getenv("SOS_NO_ROWCONV") && atoi(getenv("SOS_NO_ROWCONV")) == 1;       // A/B aid
  if (off || a.force_plan >= 0) return false;
  if (a.x_dtype != SOS_DTYPE_F16 || a.y_dtype != SOS_DTYPE_F16) return false;
  if (a.stride != 1 || a.osh != 1 || a.osw != 1 || a.oph != 0 || a.opw != 0) return false;
  if (a.OH != a.H || a.OW != a.W || a.YH != a.OH || a.YW != a.OW) return false;
  if (a.H % 128 != 0 || a.Cin % 16 != 0 || a.Cin > 64 || a.Cout != 512 / kRing / 16 * 16) return false;        // Cout == 48
  if (a.Cy % 8 != 0 || a.y_coff % 8 != 0 || a.Cy - a.y_coff < a.Cout) return false;
  const int nt = (int)a.ntaps;
  if (nt < 2 || nt > kMaxTaps) return false;
  // taps = a full KH x KW grid, symmetric about 0, uniform dilations
  int min_h = 0, max_h = 0, min_w = 0, max_w = 0;
  for (int t = 0; t < nt; ++t) {
    min_h = std::min(min_h, (int)a.tap_dh[t]); max_h = std::max(max_h, (int)a.tap_dh[t]);
    min_w = std::min(min_w, (int)a.tap_dw[t]); max_w = std::max(max_w, (int)a.tap_dw[t]);
  }
  if (min_h != -max_h || min_w != -max_w || max_w == 0) return false;
  int dh = 0, dw = 0;
  for (int t = 0; t < nt; ++t) {
    if (a.tap_dh[t] > 0) dh = dh ? std::min(dh, (int)a.tap_dh[t]) : (int)a.tap_dh[t];
    if (a.tap_dw[t] > 0) dw = dw ? std::min(dw, (int)a.tap_dw[t]) : (int)a.tap_dw[t];
  }
  const int KH = dh ? 2 * max_h / dh + 1 : 1, KW = 2 * max_w / dw + 1;
  if (KH * KW != nt || (KW & 1) == 0 || KW < 3) return false;
  std::vector<char> seen(nt, 0);
  for (int t = 0; t < nt; ++t) {
    if ((dh && a.tap_dh[t] % dh) || a.tap_dw[t] % dw) return false;
    const int f = dh ? (a.tap_dh[t] + max_h) / dh : 0, s = (a.tap_dw[t] + max_w) / dw;
    if (f < 0 || f >= KH || s < 0 || s >= KW || seen[f * KW + s]) return false;
    seen[f * KW + s] = 1;
  }
  if (KW * a.Cout > 256 || kRing < 2 * KW) return false;
  const int box_rows = 128 + (KH - 1) * dh;
  if (box_rows > 256) return false;
  const int stage = tc::round_up(box_rows * 128, 1024);
  const int cstore = std::min(tc::round_up((int)a.Cout, 8), (int)(a.Cy - a.y_coff));
  if (row_fixed_smem(nt, (int)a.Cout, cstore) + 2 * stage > tc::kSmemLimit) return false;
  if (a.stats_partial && (a.stats_channels <= 0 || a.stats_channels > a.Cy - a.y_coff || a.stats_channels < a.Cout)) return false;
  return true;
}

namespace {

int plan_rowconv(const sos_conv_args& a, RowPlan& out, bool want_bnr) {
  RowParams& p = out.p;
  memset(&p, 0, sizeof(p));
  const int nt = (int)a.ntaps, Cin = (int)a.Cin, Cout = (int)a.Cout;
  int max_h = 0, max_w = 0, dh = 0, dw = 0;
  for (int t = 0; t < nt; ++t) {
    max_h = std::max(max_h, (int)a.tap_dh[t]);
    max_w = std::max(max_w, (int)a.tap_dw[t]);
    if (a.tap_dh[t] > 0) dh = dh ? std::min(dh, (int)a.tap_dh[t]) : (int)a.tap_dh[t];
    if (a.tap_dw[t] > 0) dw = dw ? std::min(dw, (int)a.tap_dw[t]) : (int)a.tap_dw[t];
  }
  const int KH = dh ? 2 * max_h / dh + 1 : 1, KW = 2 * max_w / dw + 1;
  p.KH = KH; p.KW = KW; p.half_w = (KW - 1) / 2;
  p.Cout = Cout;
  p.Cin = Cin;
  p.dwl = dw;
  p.W = (int)a.W;
  p.n_hb = (int)a.H / 128;
  p.dh_bytes = dh * 128;
  p.h0_off = -max_h;
  p.box_rows = 128 + (KH - 1) * dh;
  p.box_bytes = p.box_rows * 128;
  p.stage_bytes = round_up(p.box_bytes, 1024);
  p.b_tile_bytes = Cout * 128;
  p.R = kRing;
  static const int one_iss = getenv("SOS_ROW_ISSUERS") && atoi(getenv("SOS_ROW_ISSUERS")) == 1;
  p.n_issuers = one_iss ? 1 : 2;
  p.ksteps = round_up(Cin, 16) / 16;
  for (int nb = 1; nb <= KW; ++nb) p.idesc[nb] = make_idesc_f16(128, nb * Cout, 0, 0);
  // smem tile (f, b) holds the tap with dh = (f - (KH-1)/2) * dh and W offset s' - half with s' = KW - 1 - b: input column i feeds
  // output o = i - (s' - half), so ascending outputs (= ascending ring slots) are ascending b
  for (int f = 0; f < KH; ++f)
    for (int b = 0; b < KW; ++b) {
      const int want_h = dh ? (f - (KH - 1) / 2) * dh : 0, want_w = ((KW - 1 - b) - p.half_w) * dw;
      int sel = -1;
      for (int t = 0; t < nt; ++t)
        if (a.tap_dh[t] == want_h && a.tap_dw[t] == want_w) sel = t;
      SOS_CHECK_ARG(sel >= 0, "sos_conv2d_tc (row kernel): tap (%d, %d) missing", want_h, want_w);
      p.tapsel[f * KW + b] = (int16_t)sel;
    }
  const int cstore = std::min(round_up(Cout, 8), (int)(a.Cy - a.y_coff));
  p.cstore = cstore;
  p.stats_c = (int)a.stats_channels;
  p.n_out = Cout;

  const int fixed = row_fixed_smem(nt, Cout, cstore, want_bnr);
  p.n_stages = std::min(4, (kSmemLimit - fixed) / p.stage_bytes);
  SOS_CHECK_ARG(p.n_stages >= 2, "sos_conv2d_tc (row kernel): shared memory");
  out.smem = fixed + p.n_stages * p.stage_bytes;

  // streams: (image, H block, lattice phase) x segments of the lattice columns; about 4 waves of work items
  const int sms = sos_num_sms();
  const int streams = (int)a.N * p.n_hb * dw;
  const int Lmax = ((int)a.W + dw - 1) / dw;
  // segments per stream: the count that maximises (static round-robin balance over the SMs) x (share of non-halo input columns)
  int best_seg = 1;
  double best_eff = 0;
  for (int ns = 1; ns <= std::max(1, Lmax / 8); ++ns) {
    const int sl = ceil_div(Lmax, ns);
    const int nsr = ceil_div(Lmax, sl);
    const double waves = (double)streams * nsr / sms;
    // (a segment's 2 * half halo columns cost A traffic and issue slots, not tensor time: about half a column each)
    const double eff = waves / std::ceil(waves) * ((double)sl / (sl + (nsr > 1 ? p.half_w : 0)));
    if (eff > best_eff + 1e-9) { best_eff = eff; best_seg = nsr; }
  }
  p.seg_len = ceil_div(Lmax, best_seg);
  p.n_seg = ceil_div(Lmax, p.seg_len);
  const long long items = (long long)streams * p.n_seg;
  SOS_CHECK_ARG(items < (1ll << 30), "sos_conv2d_tc (row kernel): too many work items");
  p.n_items = (int)items;
  out.grid = (int)std::min<long long>(items, sms);

  // tensor maps: activations (C, H, W, N) with a (64, box_rows, 1, 1) box; weights (ntaps*Cin, Cout) with a (64, Cout) box;
  // output (C, H, W, N) with a (cstore, 128, 1, 1) box, no swizzle
  {
    const uint64_t pix = (uint64_t)Cin * 2;
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)a.H, (uint64_t)a.W, (uint64_t)a.N};
    uint64_t str[4] = {2, pix * a.W, pix, pix * a.H * a.W};
    uint32_t box[4] = {64, (uint32_t)p.box_rows, 1, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    out.specA = make_spec(CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B, "row activations");
    uint64_t bd[2] = {(uint64_t)nt * Cin, (uint64_t)Cout};
    uint64_t bs[2] = {2, (uint64_t)nt * Cin * 2};
    uint32_t bb[2] = {64, (uint32_t)Cout};
    uint32_t be[2] = {1, 1};
    out.specB = make_spec(CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, bd, bs, bb, be, CU_TENSOR_MAP_SWIZZLE_128B, "row weights");
  }
  {
    const uint64_t pix = (uint64_t)a.Cy * 2;
    out.d_offset = (long long)a.y_coff * 2;
    uint64_t dims[4] = {(uint64_t)cstore, (uint64_t)a.H, (uint64_t)a.W, (uint64_t)a.N};
    uint64_t str[4] = {2, pix * a.W, pix, pix * a.H * a.W};
    uint32_t box[4] = {(uint32_t)cstore, 128, 1, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    out.specD = make_spec(CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_NONE, "row output");
  }
  return SOS_OK;
}

}  // namespace

// Called by sos_conv2d_tc (conv_tc.cu) after its argument checks when sos_rowconv_eligible(a).
int sos_rowconv_launch(const sos_conv_args& a, cudaStream_t stream) {
  std::vector<int32_t> key;
  key.reserve(24 + 2 * (size_t)a.ntaps);
  // the fused BatchNorm-backward reduction needs a little more shared memory: honoured where it still leaves two pipeline stages
  bool want_bnr = a.bnr_partial && a.bnr_y && a.bnr_scale && a.bnr_shift && a.bnr_mean && a.bnr_invstd && a.y_coff == 0 && a.bnr_channels >= a.Cout &&
                  !a.stats_partial;
  if (want_bnr) {
    int max_h = 0, dh = 0;
    for (int t = 0; t < a.ntaps; ++t) {
      max_h = std::max(max_h, (int)a.tap_dh[t]);
      if (a.tap_dh[t] > 0) dh = dh ? std::min(dh, (int)a.tap_dh[t]) : (int)a.tap_dh[t];
    }
    const int stage = round_up((128 + 2 * max_h) * 128, 1024);
    const int cstore = std::min(round_up((int)a.Cout, 8), (int)(a.Cy - a.y_coff));
    if (row_fixed_smem((int)a.ntaps, (int)a.Cout, cstore, true) + 2 * stage > kSmemLimit) want_bnr = false;
  }
  const int64_t fields[] = {a.N, a.H, a.W, a.Cin, a.Cout, a.ntaps, a.Cy, a.y_coff, a.stats_partial ? a.stats_channels : -1, want_bnr ? 1 : 0};
  for (int64_t f : fields) key.push_back((int32_t)f);
  for (int t = 0; t < a.ntaps; ++t) { key.push_back(a.tap_dh[t]); key.push_back(a.tap_dw[t]); }

  std::lock_guard<std::mutex> lock(g_row_mutex);
  RowPlan* plan;
  auto it = g_row_plans.find(key);
  if (it != g_row_plans.end()) {
    plan = it->second;
  } else {
    plan = new RowPlan();
    if (int e = plan_rowconv(a, *plan, want_bnr)) { delete plan; return e; }
    g_row_plans.emplace(std::move(key), plan);
  }
  RowParams& p = plan->p;
  p.out_scale = a.out_scale;
  p.scale = a.epi_scale;
  p.shift = a.epi_shift;
  p.act = (int)a.act;
  p.slope = a.slope;
  p.stats = a.stats_partial;
  p.bnr_partial = want_bnr ? a.bnr_partial : nullptr;
  p.bnr_y = reinterpret_cast<const uint8_t*>(a.bnr_y);
  p.bnr_scale = a.bnr_scale; p.bnr_shift = a.bnr_shift; p.bnr_mean = a.bnr_mean; p.bnr_invstd = a.bnr_invstd;
  p.bnr_c = (int)a.bnr_channels;
  p.bnr_sw = (long long)a.Cy * 2;
  p.bnr_sh = p.bnr_sw * a.W;
  p.bnr_sn = p.bnr_sh * a.H;
  const void* baseD = reinterpret_cast<const uint8_t*>(a.y) + plan->d_offset;
  if (plan->baseA != a.x) { if (int e = encode_spec(&p.mapA, plan->specA, a.x)) return e; plan->baseA = a.x; }
  if (plan->baseB != a.wk) { if (int e = encode_spec(&p.mapB, plan->specB, a.wk)) return e; plan->baseB = a.wk; }
  if (plan->baseD != baseD) { if (int e = encode_spec(&p.mapD, plan->specD, baseD)) return e; plan->baseD = baseD; }

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(rowconv_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess) {
      sos_set_error("sos_conv2d_tc (row kernel): cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SOS_ERR_CUDA;
    }
    attr_set = true;
  }
  rowconv_f16_kernel<<<plan->grid, kThreadsRow, plan->smem, stream>>>(p);
  SOS_CHECK_LAUNCH("sos_conv2d_tc (row kernel)");
  if (a.stats_rows_out) *a.stats_rows_out = plan->grid;
  if (a.bnr_rows_out) *a.bnr_rows_out = want_bnr ? plan->grid : 0;
  if (a.plan_out) {
    const int32_t po[8] = {2, 1, p.dwl, 1, p.KH, p.n_stages, p.stage_bytes, plan->grid};      // [0] = 2: row-streaming kernel
    memcpy(a.plan_out, po, sizeof(po));
  }
  return SOS_OK;
}
