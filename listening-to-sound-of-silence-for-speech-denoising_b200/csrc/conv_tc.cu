// Tap-list implicit GEMM on the 5th-gen tensor cores (tcgen05, kind::tf32, fp32 accumulate in TMEM).
//
// Every convolution on the hot path (zero-"same" dilated convs of the SID / ContextAggNet encoders,
// reflect-padded / strided / transposed convs of InpaintNet, data gradients of all of them) and the plain
// GEMMs (LSTM input projection, MLP head) are instances of
//
//     out[n, oh, ow, :] = epilogue( sum_t  in[n, oh*s + dh_t, ow*s + dw_t, :] . W_t^T )
//
// over NHWC fp32 activations.  Out-of-range input pixels read as zero (TMA out-of-bounds fill), so zero
// padding costs nothing; reflect padding is materialised by the producer layer.
//
// CTA = 224 threads, persistent over output tiles:
//   warp 0    TMA producer: per k-step one activation box per sub-tile (+ halo rows along the slow tile axis,
//             shared by every tap that differs only by a slow-axis shift) and one weight box per tap
//   warp 1, 6 MMA issuers (one elected lane each), accumulators double-buffered in TMEM.  With two sub-tiles per CTA tile
//             each issuer owns one sub-tile's accumulator: a TF32 MMA of N = 48 lasts 24 cycles, less than one thread needs
//             to issue it, so two threads feed the tensor pipe; with one sub-tile warp 6 idles
//   warps 2-5 epilogue: tcgen05.ld -> affine/activation -> smem staging -> TMA store (clips the ragged edge)
// An output tile is 128 pixels = SB (slow) x FB (fast) pixels of one image (FB = 8 or 128).  The slow axis may be
// a dilation lattice (pixels q*g + r for fixed phase r), which turns a dilated tap shift into a shift by whole
// 8-row swizzle atoms of the same staged box.
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

constexpr int kThreadsTc = 352;            // warp 0 producer, 1 and 6 MMA issuers, 2-5 and 7-10 the two epilogue groups
constexpr int kEpiWarps = 8;
// (BatchNorm partial sums of the epilogue warps: [8 warps][2][N] floats of shared memory, only when statistics are requested)
constexpr int kMaxProg = 144;             // entries of the deduplicated MMA programs
struct alignas(64) TcParams {
  CUtensorMap mapA, mapB, mapD;
  int pair;                      // CTA-pair mode (cta_group::2): mapB's box holds N / 2 weight rows, tiles are dealt to pairs
  int total_ctiles, n_nblk, tiles_fast_g, tiles_slow, n_phase;
  int S, N, cbe, n_chunks, n_groups;
  int FB, SB, stride;
  int n_issuers;                 // 1 or 2 MMA-issuing warps (2 when S == 2)
  int cin;                       // K elements per tap in the weight matrix
  int ec;                        // epilogue / store chunk width (16 or 32 channels)
  int n_stages, stage_bytes, a_box_bytes, a_box_stride, b_tile_stride, staging_bytes;
  int n_stg;                     // output staging buffers per epilogue group (2..8)
  int n_egroups;                 // epilogue groups in use: 2, or 1 (A/B switch SOS_EPI_GROUPS=1: warps 7-10 idle)
  int tstep;                     // tap index step between a group's sub-taps (bmerge)
  int bmerge;                    // +-1: the weight tiles of a tap group's sub-taps arrive with ONE 3-D TMA load (taps equally spaced in the packed rows)
  int sw;                        // register chunks (ec channels) per TMA store: 2 = half outputs staged as 128-byte rows of 64 channels
  int wstore;                    // 1 (SOS_WARP_STORE=1): every epilogue warp stores its own 32 rows; 0: one store per group and chunk behind a barrier
  int dbg;                       // measurement aid (SOS_EPI_DBG): 1 no TMA stores, 2 no accumulator reads / staging writes
  int layout_type, sbo;
  uint32_t idesc;
  const float* scale;            // optional per-output-channel affine (eval-mode BN / bias) ...
  const float* shift;
  int act;                       // 0 none, 1 relu, 2 prelu
  const float* slope;
  int esz;                       // operand element size: 4 (TF32 in fp32 storage) or 2 (half)
  int b_tile_bytes;              // bytes of one staged weight box
  int y_half;                    // outputs are stored as half (staging rows of ec * 2 bytes)
  const float* out_scale;        // optional device scalar multiplied into every output
  float* stats;                  // optional BatchNorm partial sums [4 * grid][2][stats_c] of the RAW outputs (sum, sum of squares)
  int stats_c, out_fast, lat_slow;
  int n_out;                     // real output channels: the per-channel epilogue vectors hold this many entries
  TapGroup groups[kMaxGroups];
  // MMA program: one entry per tcgen05.mma of a pipeline stage (same for every tile / channel chunk; tap groups with the
  // same structure share one program): x = A descriptor address delta (16-byte units), y = B delta, z = accumulator column
  // offset, w = 1 for the first MMA of a sub-tile's k loop.  Keeps the single issuing thread at a handful of instructions
  // per MMA (one 24/48-cycle MMA per 8 fp32 of K leaves no room for address arithmetic).
  int16_t prog0[kMaxGroups], prog_n[kMaxGroups];
  uint4 prog[kMaxProg];
};

struct TileCoord { int nb, tfg, ts, ph, n; };
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
  TileCoord c;
  c.nb = t % p.n_nblk; t /= p.n_nblk;
  c.tfg = t % p.tiles_fast_g; t /= p.tiles_fast_g;
  c.ts = t % p.tiles_slow; t /= p.tiles_slow;
  c.ph = t % p.n_phase;
  c.n = t / p.n_phase;
  return c;
}

template <int KIND, int PAIR>
__device__ __forceinline__ void tapgemm_body(const TcParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem carve-up: [stages][staging][barriers]
  const uint32_t stages_base = smem_base;
  const uint32_t staging_base = stages_base + (uint32_t)p.n_stages * p.stage_bytes;
  const uint32_t bar_base = staging_base + (uint32_t)p.staging_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (32 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (34 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 36;
  // per epilogue warp: running per-channel sum / sum of squares of this CTA's tiles, [4 warps][2][256]
  float* stats_s = reinterpret_cast<float*>(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // CTA pairs (PAIR): the two CTAs of a cluster work on two tiles at once.  Each stages its own activation boxes and HALF the rows
  // of every weight tile; the even CTA ("leader") issues M = 256 MMAs (cta_group::2) that read both CTAs' shared memory and write
  // both CTAs' TMEM -- the B operand bytes per FLOP, the part of the shared-memory traffic that bounds the wide layers, halve.
  // Loads of both CTAs complete on the leader's `full` barriers; MMA commits arrive on both CTAs' `empty` / `tfull` barriers
  // (multicast); both CTAs' epilogue warps arrive on the leader's `tempty` barriers.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int grid_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;          // CTAs, or CTA pairs
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = PAIR ? (p.total_ctiles + 1) >> 1 : p.total_ctiles;        // tiles, or tile pairs (an odd last tile gets a dummy partner)
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    prefetch_tmap(&p.mapD);
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.n_issuers);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), p.n_issuers);
      mbar_init(tempty_bar(s), PAIR ? 8 : 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0, t_begin = clock64();
    int n_tiles_done = 0;
    for (int cu = unit0; cu < n_units; cu += grid_units, ++n_tiles_done) {
      const int ct = PAIR ? 2 * cu + (int)rank : cu;                 // (ct == total_ctiles: the dummy partner -- image index out of range,
      const TileCoord tc = decode_tile(p, ct);                       //  every box zero-filled, every store clipped)
      const int fast_t = tc.tfg * p.S * p.FB * p.stride, slow_t = tc.ts * p.SB * p.stride, wrow = tc.nb * p.N;
      for (int g = 0; g < p.n_groups; ++g) {
        const TapGroup& grp = p.groups[g];
        const int n_sub = grp.n_sub;
        // bytes that land per stage: this CTA's boxes, or (PAIR, counted on the leader's barrier) both CTAs' activation boxes and weight halves
        const uint32_t tx_bytes = (PAIR ? 2u : 1u) * (uint32_t)p.S * p.a_box_bytes + (uint32_t)n_sub * p.b_tile_bytes;
        const int fast0 = fast_t + grp.d_fast, slow0 = slow_t + grp.d_slow;
        for (int c = 0; c < p.n_chunks; ++c) {
          const long long tw = (p.dbg & 16) ? clock64() : 0;
          mbar_wait(empty_bar(stage), phase ^ 1, 100);
          if (p.dbg & 16) w_empty += clock64() - tw;
          if (elect_one_sync()) {
            const uint32_t sbase = stages_base + (uint32_t)stage * p.stage_bytes;
            const uint32_t bbase = sbase + (uint32_t)p.S * p.a_box_stride;
            if (PAIR) {
              const uint32_t fb = leader_addr(full_bar(stage));
              if (rank == 0) mbar_expect_tx(full_bar(stage), tx_bytes);
              for (int s = 0; s < p.S; ++s)
                tma_load_5d_pair(sbase + (uint32_t)s * p.a_box_stride, &p.mapA, fb, c * p.cbe, fast0 + s * p.FB * p.stride, slow0, tc.ph, tc.n);
              if (p.bmerge) {
                const int tb = grp.tap[p.bmerge > 0 ? 0 : n_sub - 1];
                tma_load_4d_pair(bbase, &p.mapB, fb, c * p.cbe, wrow + (int)rank * (p.N >> 1), tb % p.tstep, tb / p.tstep);
              }
              else
                for (int j = 0; j < n_sub; ++j)           // this CTA's half of the weight rows
                  tma_load_2d_pair(bbase + (uint32_t)j * p.b_tile_stride, &p.mapB, fb, grp.tap[j] * p.cin + c * p.cbe, wrow + (int)rank * (p.N >> 1));
            } else {
              mbar_expect_tx(full_bar(stage), tx_bytes);
              for (int s = 0; s < p.S; ++s)
                tma_load_5d(sbase + (uint32_t)s * p.a_box_stride, &p.mapA, full_bar(stage), c * p.cbe, fast0 + s * p.FB * p.stride, slow0, tc.ph, tc.n);
              if (p.bmerge) {
                const int tb = grp.tap[p.bmerge > 0 ? 0 : n_sub - 1];
                tma_load_4d(bbase, &p.mapB, full_bar(stage), c * p.cbe, wrow, tb % p.tstep, tb / p.tstep);
              }
              else
                for (int j = 0; j < n_sub; ++j)
                  tma_load_2d(bbase + (uint32_t)j * p.b_tile_stride, &p.mapB, full_bar(stage), grp.tap[j] * p.cin + c * p.cbe, wrow);
            }
          }
          __syncwarp();
          if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if ((p.dbg & 16) && blockIdx.x == 0 && lane == 0)
      printf("tapgemm dbg: producer %d tiles, %lld cycles, %lld waiting for empty stages\n", n_tiles_done, clock64() - t_begin, w_empty);
  } else if (warp == 1 || warp == 6) {
    // ===================================================================== MMA issuer(s)
    const int issuer = warp == 1 ? 0 : 1;
    // ONE thread runs the whole issue loop (elected once, outside every loop: per pipeline stage the thread then executes a
    // barrier poll, two address adds, its MMAs and one commit -- no per-stage elect / reconvergence, no integer division).
    // An N = 48 MMA lasts 24 cycles and a stage holds only 5-15 of them, so every instruction of this loop is on the kernel's
    // critical path (ncu r02: 124 instructions / 965 cycles per 5-MMA stage before this form).
    if (issuer < p.n_issuers && rank == 0 && elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // descriptor words: hi = SBO | version | layout (constant); lo = (address >> 4) | LBO, advanced by the program's deltas
      const uint32_t desc_hi = (uint32_t)((p.sbo >> 4) & 0x3FFF) | (1u << 14) | ((uint32_t)(p.layout_type & 7) << 29);
      const uint32_t lbo_bits = (16u >> 4) << 16;
      const uint32_t idesc = p.idesc;
      const uint32_t a_lo0 = (stages_base >> 4) | lbo_bits;
      const uint32_t b_lo0 = ((stages_base + (uint32_t)p.S * p.a_box_stride) >> 4) | lbo_bits;
      const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
      const int n_groups = p.n_groups, n_chunks = p.n_chunks, n_stages = p.n_stages;
      long long w_te = 0, w_full = 0;
      const long long t_begin = clock64();
      for (int cu = unit0; cu < n_units; cu += grid_units) {
        long long tw = (p.dbg & 16) ? clock64() : 0;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1, 200);
        if (p.dbg & 16) w_te += clock64() - tw;
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)acc * 256;
        uint32_t first_mask = 1u;
        for (int g = 0; g < n_groups; ++g) {
          const int per = p.prog_n[g];                         // MMAs per issuer and stage; the program lists sub-tile 0's first
          const int m0 = p.prog0[g] + issuer * per, m1 = m0 + per;
          for (int c = 0; c < n_chunks; ++c) {
            tw = (p.dbg & 16) ? clock64() : 0;
            mbar_wait(full_bar(stage), phase, 201);
            if (p.dbg & 16) w_full += clock64() - tw;
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step, b_lo = b_lo0 + (uint32_t)stage * stage_step;
#pragma unroll 5
            for (int m = m0; m < m1; ++m) {
              const uint4 e = p.prog[m];
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + e.x);
              const uint64_t bd = ((uint64_t)desc_hi << 32) | (b_lo + e.y);
              if (PAIR) umma_f16_pair(d_base + e.z, ad, bd, idesc, (e.w & first_mask) == 0u);
              else umma<KIND>(d_base + e.z, ad, bd, idesc, (e.w & first_mask) == 0u);
            }
            if (PAIR) umma_commit_pair(empty_bar(stage));
            else umma_commit(empty_bar(stage));            // frees the smem stage when these MMAs retire
            first_mask = 0u;
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
        if (PAIR) umma_commit_pair(tfull_bar(acc));
        else umma_commit(tfull_bar(acc));                  // (tracks every MMA this thread issued before it)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if ((p.dbg & 16) && blockIdx.x == 0)
        printf("tapgemm dbg: issuer %d %lld cycles, %lld waiting for a free accumulator, %lld for full stages\n", issuer, clock64() - t_begin, w_te, w_full);
    }
  } else {
    // ===================================================================== epilogue (warps 2..5 = group 0, 7..10 = group 1)
    // Two groups of four warps: group e drains accumulator buffer e, i.e. every other tile of this CTA, with its own staging
    // buffers, named barrier and store queue -- the drains of consecutive tiles overlap.  (With one group, layers with few taps /
    // channels were bound by the latency chain tcgen05.ld -> st.shared -> fence -> barrier -> TMA store of a single chunk.)
    const int eg = warp >= 7 ? 1 : 0;
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;                // accumulator row = pixel within the tile
    const int ethread = threadIdx.x - (eg ? 224 : 64);         // 0..127 within the group
    const int ewarp = eg * 4 + (ethread >> 5);    // 0..7: statistics row of this warp
    const int neg = p.n_egroups;
    int acc = neg == 2 ? eg : 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    const int n_stg = neg == 2 ? p.n_stg : 2 * p.n_stg;
    const uint32_t my_staging = staging_base + (neg == 2 ? (uint32_t)eg * (uint32_t)(p.staging_bytes >> 1) : 0u);
    const float slope = ((p.act & SOS_ACT_MASK) == 2 && p.slope) ? *p.slope : 0.f;
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    const int n_ec = p.N / p.ec;
    const int crow = p.ec * (p.y_half ? 2 : 4);          // bytes of one register chunk in a staging row
    const int erow = crow * p.sw;                        // bytes of one staging row = the store's box width
    float* my_stats = stats_s + ewarp * 2 * p.N;         // this warp's [2][N]
    if (p.stats) {
      for (int i = lane; i < 2 * p.N; i += 32) my_stats[i] = 0.f;
      __syncwarp();
    }
    // pixel of this thread's accumulator row inside the tile (rows are [slow][fast])
    const int row_f = row % p.FB, row_s = row / p.FB;
    const int wfast = (q * 32) % p.FB, wslow = (q * 32) / p.FB;        // this warp's 32 rows inside the tile
    long long w_tfull = 0;
    const long long t_begin = clock64();
    for (int cu = unit0 + eg * grid_units; cu < n_units && eg < neg; cu += neg * grid_units) {
      const int ct = PAIR ? 2 * cu + (int)rank : cu;
      const TileCoord tc = decode_tile(p, ct);
      const long long tw = (p.dbg & 16) ? clock64() : 0;
      mbar_wait(tfull_bar(acc), acc_phase, 300);
      if (p.dbg & 16) w_tfull += clock64() - tw;
      tc_fence_after();
      for (int s = 0; s < p.S; ++s) {
        for (int cc = 0; cc < n_ec; ++cc) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + s * p.N + cc * p.ec);
          uint32_t r[32];
          if (!(p.dbg & 2)) {
            tmem_ld16(taddr, r);
            if (p.ec == 32) tmem_ld16(taddr + 16, r + 16);
            tmem_ld_wait();
          }
          const int ch0 = tc.nb * p.N + cc * p.ec;
          if (p.stats) {
            // BatchNorm statistics of the raw outputs: transpose-reduce over the warp's 32 rows (31 shuffles per quantity),
            // after which lane i holds the total of channel ch0 + i; rows outside the image (ragged tiles) count as zero
            const bool in_img = (tc.tfg * p.S + s) * p.FB + row_f < p.out_fast && tc.ts * p.SB + row_s < p.lat_slow;
            float v[32], w[32];
            if (p.ec == 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                v[i] = in_img ? __uint_as_float(r[i]) : 0.f;
                w[i] = v[i] * v[i];
              }
#pragma unroll
              for (int st = 16; st >= 1; st >>= 1) {
                const bool up = (lane & st) != 0;
#pragma unroll
                for (int i = 0; i < st; ++i) {
                  const float send_v = up ? v[i] : v[i + st], keep_v = up ? v[i + st] : v[i];
                  const float send_w = up ? w[i] : w[i + st], keep_w = up ? w[i + st] : w[i];
                  v[i] = keep_v + __shfl_xor_sync(0xffffffffu, send_v, st);
                  w[i] = keep_w + __shfl_xor_sync(0xffffffffu, send_w, st);
                }
              }
              my_stats[cc * 32 + lane] += v[0];
              my_stats[p.N + cc * 32 + lane] += w[0];
            } else {
              // 16 channels: 15 shuffles per quantity bring lane i the sum of channel (i & 15) over its half-warp's rows, one more
              // adds the two half-warps
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = in_img ? __uint_as_float(r[i]) : 0.f;
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
                my_stats[p.N + cc * 16 + lane] += w[0];
              }
            }
          }
          if (p.shift || p.act || p.out_scale) {
            // branch-free over the 32 accumulator columns (columns >= ec hold don't-care values that are never stored); every
            // condition is uniform and tested once per chunk, not once per element
            const int act = p.act & SOS_ACT_MASK;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * oscale;
            if (p.shift) {
              // per-channel affine (eval-mode BatchNorm) or bias only (scale == NULL); the vectors hold n_out entries
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (4 * j < p.ec) {
                  const int cj = ch0 + 4 * j;
                  float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (cj + 4 <= p.n_out && (p.n_out & 3) == 0) {
                    b = __ldg(reinterpret_cast<const float4*>(p.shift + cj));
                    if (p.scale) a = __ldg(reinterpret_cast<const float4*>(p.scale + cj));
                  } else {
                    float* af = &a.x; float* bf = &b.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      if (cj + k < p.n_out) { bf[k] = __ldg(p.shift + cj + k); if (p.scale) af[k] = __ldg(p.scale + cj + k); }
                  }
                  v[4 * j] = fmaf(v[4 * j], a.x, b.x);
                  v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
                  v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z);
                  v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
                }
              }
            }
            if (act == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (act == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;
            } else if (act == 3) {                       // sigmoid (mask head, M2/networks.py:70)
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 1.f / (1.f + expf(-v[i]));
            }
            if ((p.act & SOS_ACT_ROUND_TF32) != 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = tf32_rna(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
          }
          // (this buffer is free: thread 0 waited for the store that last read it before the previous chunk's barrier)
          const uint32_t sbuf = my_staging + (uint32_t)buf * (128 * erow);     // (128 * erow == the plan's stg_bytes)
          const uint32_t srow = sbuf + (uint32_t)row * erow;
          // staging rows are written in the TMA store's swizzle (128B rows: 16-byte chunk ^= row & 7; 64B rows: chunk ^=
          // (row >> 1) & 3; 32B rows: chunk ^= (row >> 2) & 1), which is also bank-conflict free for a quarter-warp of
          // consecutive rows
          const int xr = erow == 128 ? (row & 7) : (erow == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
          const int sub = p.sw == 2 ? (cc & 1) : 0;                      // which half of a 2-chunk staging row this chunk fills
          const int jb = sub * (crow >> 4);
          if (p.dbg & 2) {
          } else if (p.y_half) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < p.ec / 8) {
                uint32_t h[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = pack_half2(__uint_as_float(r[8 * j + 2 * k]), __uint_as_float(r[8 * j + 2 * k + 1]));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((jb + j) ^ xr) << 4)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                             : "memory");
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < p.ec / 4)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ xr) << 4)), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                             "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                             : "memory");
            }
          }
          if (sub != p.sw - 1 && cc != n_ec - 1) continue;               // (the row's other half comes with the next chunk)
          const int ch_store = ch0 - sub * p.ec;
          fence_proxy_async_smem();
          // Every warp stores ITS 32 rows (a quarter of the tile: box = ec x min(FB, 32) x max(1, 32 / FB) pixels) with its own
          // bulk-group queue: no barrier between the four warps of a group, so the chains tcgen05.ld -> st.shared -> fence -> TMA
          // store of the eight epilogue warps run independently (layers with little MMA work per tile are bound by that chain:
          // ~2000 cycles per 128-row chunk when the four warps met at a barrier and one thread stored for all).  Before the
          // warp-level sync the elected lane makes sure that at most n_stg - 2 of its earlier stores still read their buffers,
          // i.e. the buffer of the NEXT chunk is free by the time the warp writes it.
          if (p.wstore ? elect_one_sync() : (ethread < 32 && elect_one_sync())) {
            switch (n_stg) {                              // (the wait count is an immediate)
              case 2: bulk_wait_read<0>(); break;
              case 3: bulk_wait_read<1>(); break;
              case 4: bulk_wait_read<2>(); break;
              case 5: bulk_wait_read<3>(); break;
              case 6: bulk_wait_read<4>(); break;
              case 7: bulk_wait_read<5>(); break;
              case 8: bulk_wait_read<6>(); break;
              default: bulk_wait_read<7>(); break;
            }
          }
          if (p.wstore) {
            __syncwarp();
            if (!(p.dbg & 1) && elect_one_sync()) {
              tma_store_5d(&p.mapD, sbuf + (uint32_t)(q * 32) * erow, ch_store, (tc.tfg * p.S + s) * p.FB + wfast, tc.ts * p.SB + wslow, tc.ph, tc.n);
              bulk_commit();
            }
          } else {
            asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
            if (ethread < 32 && !(p.dbg & 1) && elect_one_sync()) {
              tma_store_5d(&p.mapD, sbuf, ch_store, (tc.tfg * p.S + s) * p.FB, tc.ts * p.SB, tc.ph, tc.n);
              bulk_commit();
            }
          }
          if (++buf == n_stg) buf = 0;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tempty_bar(acc));
        else mbar_arrive(tempty_bar(acc));
      }
      if (neg == 2) acc_phase ^= 1;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if ((p.dbg & 16) && blockIdx.x == 0 && ethread == 0)
      printf("tapgemm dbg: epilogue group %d %lld cycles, %lld waiting for full accumulators\n", eg, clock64() - t_begin, w_tfull);
    if ((p.wstore || ethread < 32) && elect_one_sync()) bulk_wait<0>();
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();                // (the peer's MMAs read this CTA's shared memory and write its TMEM until the very end)
  else __syncthreads();
  if (p.stats) {
    // one row of BatchNorm partial sums per CTA: the eight epilogue warps' running sums, added up in a fixed order
    const int N = p.N;
    float* dst = p.stats + (size_t)blockIdx.x * 2 * p.stats_c;
    for (int c = threadIdx.x; c < 2 * p.stats_c; c += blockDim.x) {
      const int half = c >= p.stats_c ? 1 : 0, ch = c - half * p.stats_c;
      float sum = 0.f;
      if (ch < N)
        for (int w = 0; w < kEpiWarps; ++w) sum += stats_s[w * 2 * N + half * N + ch];
      dst[c] = sum;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(kThreadsTc, 1) tapgemm_tf32_kernel(const __grid_constant__ TcParams p) { tapgemm_body<0, 0>(p); }
__global__ void __launch_bounds__(kThreadsTc, 1) tapgemm_f16_kernel(const __grid_constant__ TcParams p) { tapgemm_body<1, 0>(p); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTc, 1) tapgemm_f16_pair_kernel(const __grid_constant__ TcParams p) {
  tapgemm_body<1, 1>(p);
}

// ---- plan cache (SURVEY 8b `sos_plan_*`): the planner sweep, the MMA program and the tensor-map geometry depend only on the
// call's shapes / taps / types, never on its pointers.  They are computed once per distinct geometry (~200 per training step,
// identical from step to step) and kept for the life of the process; a call then patches the pointers and re-encodes a tensor
// map only when its base address differs from the one the cached map was encoded for.
struct TcPlan {
  TcParams p;
  MapSpec specA, specB, specD;
  const void *baseA = nullptr, *baseB = nullptr, *baseD = nullptr;
  long long d_offset = 0;          // byte offset of the output map's origin inside y (channel slice + lattice phase)
  int smem = 0, grid = 0, esz = 4;
  int32_t plan_out[8] = {0};
};
std::mutex g_plan_mutex;
std::map<std::vector<int32_t>, TcPlan*> g_plans;
long long g_plan_hits = 0, g_plan_misses = 0;

// Everything that does not depend on the call's pointers: blocking, plan sweep, MMA program, tensor-map geometry, grid.
int plan_conv2d_tc(const sos_conv_args& a, TcPlan& out) {
  const int esz = a.x_dtype == SOS_DTYPE_F16 ? 2 : 4;       // operand element size
  const int kpe = 32 / esz;                                // K elements per MMA
  const int ysz = a.y_dtype == SOS_DTYPE_F16 ? 2 : 4;
  out.esz = esz;

  // ---- output-channel blocking
  const int Cin = (int)a.Cin, Cout = (int)a.Cout;
  const int CinK = round_up(Cin, kpe);                      // K extent per tap as the MMAs see it
  const int Ntot = round_up(Cout, 16);
  int N = Ntot, n_nblk = 1;
  if (Ntot > 256) {
    int best_waste = 1 << 30;
    for (int nb = ceil_div(Ntot, 256); nb <= ceil_div(Ntot, 256) + 4; ++nb) {
      const int n = round_up(ceil_div(Ntot, nb), 16);
      if (n <= 256 && n * nb - Ntot < best_waste) { best_waste = n * nb - Ntot; N = n; n_nblk = nb; }
    }
  }
  // CTA pairs (cta_group::2, M = 256 per MMA): for the wide layers, whose N = Cout-wide B operand dominates the shared-memory traffic
  static const int pair_env = getenv("SOS_PAIR") ? atoi(getenv("SOS_PAIR")) : 1;     // (A/B switch: SOS_PAIR=0 turns the pair kernel off)
  // (only where a tile carries enough MMA work: layers with few taps / channels are bound by their epilogue and stores, and pairing
  //  costs them 5-15 % -- measured: 7x1 96->96, 1x7 2->96, the transposed convolutions)
  const double tile_mma = (double)a.ntaps * (CinK / kpe) * (N / 2.0) * (2 * N <= 256 ? 2 : 1);
  const int pair = (pair_env >= 1 && esz == 2 && n_nblk == 1 && N % 16 == 0 && N >= 48 &&
                    (tile_mma >= 6000.0 || (a.ntaps >= 9 && tile_mma >= 3000.0) || pair_env == 2) && a.force_plan < 0) ? 1 : 0;
  const int Nb = pair ? N / 2 : N;                           // weight rows a CTA stages per tap
  const int ec = (N % 32 == 0) ? 32 : 16;
  const int stg_bytes = 128 * ec * ysz;                     // one output staging buffer (a multiple of 1024: swizzle-aligned)
  const int stats_smem = a.stats_partial ? kEpiWarps * 2 * N * 4 : 0;
  const int avail = kSmemLimit - 1024 - 2 * 2 * stg_bytes - 512 - stats_smem;   // (two epilogue groups x at least 2 staging buffers)

  // ---- choose orientation / box sharing / taps per box / channel chunk / sub-tiles: the cheapest candidate (tensor time vs
  //      L2->smem feed time per output pixel) whose pipeline stage fits at least twice (three times preferred) in shared memory
  Geometry geo{(int)a.ntaps, a.tap_dh, a.tap_dw, (int)a.H, (int)a.W, (int)a.OH, (int)a.OW, (int)a.stride, Cin, Cout,
               a.osh == 1 && a.osw == 1 && a.oph == 0 && a.opw == 0, kMaxSub, true, esz};
  struct Cand { Plan pl; int cbe = 0, S = 0, n_stages = 0, stage_bytes = 0, a_box_bytes = 0; double cost = 1e300, tile_cycles = 0; };
  Cand best;
  auto consider = [&](const Plan& pl, int cbe, int S) {
    if (n_nblk > 1 || S * N > 256) S = 1;
    const int cb = cbe * esz;
    const int a_box = (pl.SB + pl.halo) * pl.FB * cb;
    int max_sub = 1;
    for (auto& g : pl.groups) max_sub = std::max(max_sub, (int)g.n_sub);
    const int stage = S * round_up(a_box, 1024) + max_sub * round_up(Nb * cb, 1024);
    const int n_stages = std::min(8, avail / stage);
    if (n_stages < 2) return;
    const int out_fast = pl.fast_is_w ? (int)a.OW : (int)a.OH, out_slow = pl.fast_is_w ? (int)a.OH : (int)a.OW;
    const int tiles_fast = ceil_div(out_fast, pl.FB);
    const double mma = (double)a.ntaps * (CinK / kpe) * S * (128.0 * N / 256.0);
    const double bytes = ((double)pl.groups.size() * S * (pl.SB + pl.halo) * pl.FB + (double)a.ntaps * Nb) * Cin * (double)esz;
    const int lat_slow = out_slow / pl.g;
    const double util = ((double)lat_slow / (ceil_div(lat_slow, pl.SB) * pl.SB)) * ((double)out_fast / (ceil_div(tiles_fast, S) * S * pl.FB));
    double cost = std::max(mma, bytes / 40.0) / (util * S);
    if (n_stages < 3) cost *= 1.3;
    // Measured (r02, scripts/bench_conv.py): 32-byte operand rows (one L2 sector per TMA row, 32B swizzle) cost ~1.4x, more when
    // several such chunks make up K (48 -> 48 5x5 as 3 x 16 channels: 0.33 ms against 0.25 ms as ONE zero-tailed 64-wide chunk);
    // TMA's out-of-bounds zero fill is cheap for a quarter of the box (48 of 64 channels) but not for half or more of it
    // (16 -> 64 as a half-empty 32-wide chunk: 0.20 ms against 0.15 ms with 16-wide rows; 8 of 32 channels: 0.35-0.5 against 0.2).
    const int n_chunks = ceil_div(CinK, cbe);
    if (cb == 32) cost *= 1.4 + 0.2 * (n_chunks - 1);
    else if (cb == 64) cost *= 1.05;
    if (cbe * n_chunks > Cin) {
      const double oob = 1.0 - (double)Cin / (cbe * n_chunks);
      cost *= 1.0 + 2.0 * oob * oob;
    }
    if (cost < best.cost) { best.pl = pl; best.cbe = cbe; best.S = S; best.n_stages = n_stages; best.stage_bytes = stage; best.a_box_bytes = a_box; best.cost = cost; best.tile_cycles = std::max(mma, bytes / 40.0); }
  };
  static const int force_cbe = getenv("SOS_FORCE_CBE") ? atoi(getenv("SOS_FORCE_CBE")) : 0;     // debugging aid
  auto sweep = [&](bool fw, bool sh) -> bool {
    bool any = false;
    static const int force_ms = getenv("SOS_FORCE_MAXSUB") ? atoi(getenv("SOS_FORCE_MAXSUB")) : 0;   // debugging aid
    for (int ms = sh ? (force_ms > 0 ? std::min(force_ms, kMaxSub) : kMaxSub) : 1; ms >= 1; --ms) {
      geo.max_sub = ms;
      Plan pl;
      if (!build_plan(geo, fw, sh, pl)) continue;
      any = true;
      for (int cbe = 4 * kpe; cbe >= kpe; cbe >>= 1) {
        // a chunk wider than the whole K extent is legal: ONE chunk whose box tail is zero-filled by TMA (activations) or
        // never multiplied (weights); the MMA program then stops after CinK / kpe steps
        if (CinK % cbe && cbe < CinK) continue;   // (a ragged second chunk -- 48 channels as 32 + 16 of 64-byte rows, 5 pipeline stages instead of 2 -- was measured: 3 % slower)
        if (force_cbe > 0 && cbe != force_cbe) continue;
        consider(pl, cbe, 2);
        consider(pl, cbe, 1);
      }
    }
    return any;
  };
  if (a.force_plan >= 0) {            // test hook: bit0 = fast_is_w, bit1 = share
    SOS_CHECK_ARG(sweep((a.force_plan & 1) != 0, (a.force_plan & 2) != 0), "sos_conv2d_tc: forced plan %d not applicable", (int)a.force_plan);
  } else {
    for (int fw = 1; fw >= 0; --fw)
      for (int sh = 1; sh >= 0; --sh) sweep(fw != 0, sh != 0);
  }
  SOS_CHECK_ARG(best.cost < 1e299, "sos_conv2d_tc: no feasible plan (Cin %d Cout %d taps %d)", Cin, Cout, (int)a.ntaps);
  const Plan& pl = best.pl;

  TcParams& p = out.p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.n_nblk = n_nblk;
  p.cbe = best.cbe;
  p.n_chunks = ceil_div(CinK, p.cbe);
  p.cin = Cin;
  p.ec = ec;
  p.FB = pl.FB;
  p.SB = pl.SB;
  p.S = best.S;
  p.n_issuers = best.S == 2 ? 2 : 1;
  p.stride = (int)a.stride;
  p.n_groups = (int)pl.groups.size();
  for (int i = 0; i < p.n_groups; ++i) p.groups[i] = pl.groups[i];
  const int cb = p.cbe * esz;
  p.esz = esz;
  p.b_tile_bytes = N * cb;
  p.y_half = ysz == 2;
  p.layout_type = cb == 128 ? 2 : (cb == 64 ? 4 : 6);
  p.sbo = 8 * cb;
  p.pair = pair;
  p.idesc = esz == 2 ? make_idesc_f16(p.pair ? 256 : 128, N, 0, 0) : make_idesc_tf32(128, N, 0, 0);
  const int box_slow = pl.SB + pl.halo;
  p.a_box_bytes = best.a_box_bytes;
  p.a_box_stride = round_up(p.a_box_bytes, 1024);
  p.b_tile_stride = round_up(Nb * cb, 1024);
  // The producer thread needs ~140 cycles per TMA instruction (role timers, SOS_EPI_DBG=16: on the 96-channel 5x5 layers it is busy
  // 85 % of the time and the issuers wait 16 % of theirs for operands), so the n_sub weight tiles of a stage come with ONE 3-D load
  // where the group's taps are equally spaced in the packed weight rows (k = tap * Cin + ci): dim 2 walks the taps -- upwards; a
  // group whose taps run downwards (data gradients: negated offsets) is loaded from its last tap and its tiles are used in reverse.
  int tstep = 0;
  bool merge = !(getenv("SOS_B_MERGE") && atoi(getenv("SOS_B_MERGE")) == 0) && p.b_tile_stride == Nb * cb;
  for (int gi = 0; gi < p.n_groups && merge; ++gi) {
    const TapGroup& grp = p.groups[gi];
    if (grp.n_sub != p.groups[0].n_sub || grp.n_sub < 2) merge = false;
    for (int j = 1; j < grp.n_sub && merge; ++j) {
      const int dlt = grp.tap[j] - grp.tap[j - 1];
      if (dlt == 0 || (tstep && dlt != tstep)) merge = false;
      tstep = dlt;
    }
  }
  p.bmerge = merge ? (tstep > 0 ? 1 : -1) : 0;
  if (tstep < 0) tstep = -tstep;
  p.tstep = std::max(1, tstep);
  p.stage_bytes = best.stage_bytes;
  p.n_stages = best.n_stages;
  {
    // A TMA store holds its staging buffer for ~2000 cycles (issue -> shared memory read), so a CTA tile of S * N / ec chunk
    // stores needs enough buffers in flight to hide that behind its MMA / load time: layers with few taps or channels (2 -> 96,
    // 64 -> 2 and their gradients) are otherwise bound by the store latency (ncu r02: 1.3 TB/s of stores with 2 buffers).  Extra
    // buffers are taken from the operand pipeline, which keeps at least 2 (3 if it had them) stages.
    const double per_tile = (double)best.S * (N / ec) * 2000.0 / std::max(1.0, best.tile_cycles);
    // The store engine's cost is per ROW of the box (~10 cycles per pixel row whether it holds 64 or 128 bytes: SOS_EPI_DBG
    // ablations), so a store-bound layer with half outputs stages TWO register chunks side by side: 128-byte
    // rows of 64 channels, half the rows, fences and barriers per tile (the last store of 96 channels is clipped by the tensor map).
    static const int wide_env = !(getenv("SOS_WIDE_STORE") && atoi(getenv("SOS_WIDE_STORE")) == 0);
    p.sw = (wide_env && ysz == 2 && ec == 32 && N >= 64 && per_tile >= 3.0) ? 2 : 1;
    const int sbytes = stg_bytes * p.sw;
    int n_stg = std::min(p.sw == 2 ? 4 : 8, std::max(3, (int)std::ceil(per_tile / p.sw / 2) + 1));   // per epilogue group (2 only when shared memory is short:
                                                                                     // consecutive stores of a group then serialise)
    const int budget = kSmemLimit - 1024 - 512 - stats_smem;
    const int keep = per_tile >= 3.0 ? std::min(p.n_stages, 2) : p.n_stages;   // (a third buffer never costs a tensor-bound layer a pipeline stage)
    while (n_stg > 2 && (budget - 2 * n_stg * sbytes) / p.stage_bytes < keep) --n_stg;
    p.n_stages = std::min(p.n_stages, (budget - 2 * n_stg * sbytes) / p.stage_bytes);
    p.n_stg = n_stg;
    p.staging_bytes = 2 * n_stg * sbytes;
    static const int one_group = getenv("SOS_EPI_GROUPS") && atoi(getenv("SOS_EPI_GROUPS")) == 1;     // A/B aid
    p.n_egroups = one_group ? 1 : 2;
    static const int epi_dbg = getenv("SOS_EPI_DBG") ? atoi(getenv("SOS_EPI_DBG")) : 0;
    p.dbg = epi_dbg;
  }
  {
    int n = 0;
    const int kk_per_chunk = std::min(p.cbe, CinK) / kpe;
    for (int gi = 0; gi < p.n_groups; ++gi) {
      const TapGroup& grp = p.groups[gi];
      int same = -1;                                      // an earlier group with the same sub-tap structure shares its program
      for (int gj = 0; gj < gi && same < 0; ++gj)
        if (p.groups[gj].n_sub == grp.n_sub && memcmp(p.groups[gj].a_off, grp.a_off, grp.n_sub) == 0) same = gj;
      if (same >= 0) { p.prog0[gi] = p.prog0[same]; p.prog_n[gi] = p.prog_n[same]; continue; }
      p.prog0[gi] = (int16_t)n;
      for (int s2 = 0; s2 < p.S; ++s2)
        for (int j = 0; j < grp.n_sub; ++j)
          for (int kk = 0; kk < kk_per_chunk; ++kk) {
            const uint32_t a_delta = (uint32_t)s2 * p.a_box_stride + (uint32_t)grp.a_off[j] * (p.FB * cb) + kk * 32;
            const uint32_t b_delta = (uint32_t)(p.bmerge < 0 ? grp.n_sub - 1 - j : j) * p.b_tile_stride + kk * 32;
            SOS_CHECK_ARG(n < kMaxProg, "sos_conv2d_tc: MMA program of %d entries is too long", n);
            p.prog[n++] = make_uint4(a_delta >> 4, b_delta >> 4, (uint32_t)s2 * N, (j == 0 && kk == 0) ? 1u : 0u);
          }
      p.prog_n[gi] = (int16_t)((n - p.prog0[gi]) / p.n_issuers);   // per issuer (sub-tile)
    }
  }
  p.stats_c = (int)a.stats_channels;
  p.n_out = Cout;
  SOS_CHECK_ARG(a.stats_partial == nullptr || (n_nblk == 1 && a.stats_channels > 0 && a.stats_channels <= 256 && a.stats_channels <= a.Cy - a.y_coff),
                "sos_conv2d_tc: fused BatchNorm statistics need raw outputs of at most 256 channels in one channel block");

  // ---- tensor-map geometry.  Dim order: (channel, fast, slow/g, phase(g), image)
  const bool fw = pl.fast_is_w;
  const int g = pl.g;
  const uint64_t pixA = (uint64_t)Cin * esz;
  const CUtensorMapDataType dtA = esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  const uint64_t in_fast = fw ? a.W : a.H, in_slow = fw ? a.H : a.W;
  const uint64_t sA_fast = fw ? pixA : pixA * a.W, sA_slow = fw ? pixA * a.W : pixA;
  const CUtensorMapSwizzle sw = cb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    uint64_t dims[5] = {(uint64_t)Cin, in_fast, in_slow / g, (uint64_t)g, (uint64_t)a.N};
    uint64_t str[5] = {(uint64_t)esz, sA_fast, sA_slow * g, sA_slow, pixA * a.H * a.W};
    uint32_t box[5] = {(uint32_t)p.cbe, (uint32_t)(pl.FB * a.stride), (uint32_t)(box_slow * a.stride), 1, 1};
    uint32_t es[5] = {1, (uint32_t)a.stride, (uint32_t)a.stride, 1, 1};
    SOS_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "sos_conv2d_tc: activation box too large");
    out.specA = make_spec(dtA, 5, dims, str, box, es, sw, "activations");
    const uint32_t nb_box = (uint32_t)(p.pair ? N / 2 : N);                  // (a CTA of a pair stages half the weight rows)
    if (p.bmerge) {
      // (k within the tap, weight row, tap index mod step, tap index / step): every coordinate is bounds-checked by TMA -- a chunk
      // that reaches past the tap's Cin columns is zero filled, never read from the next tap / row / beyond the buffer
      uint64_t bd[4] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)tstep, (uint64_t)((a.ntaps - 1) / tstep + 1)};
      uint64_t bs[4] = {(uint64_t)esz, (uint64_t)a.ntaps * Cin * esz, (uint64_t)Cin * esz, (uint64_t)tstep * Cin * esz};
      uint32_t bb[4] = {(uint32_t)p.cbe, nb_box, 1, (uint32_t)p.groups[0].n_sub};
      uint32_t be[4] = {1, 1, 1, 1};
      out.specB = make_spec(dtA, 4, bd, bs, bb, be, sw, "weights");
    } else {
      uint64_t bd[2] = {(uint64_t)a.ntaps * Cin, (uint64_t)Cout};
      uint64_t bs[2] = {(uint64_t)esz, (uint64_t)a.ntaps * Cin * esz};
      uint32_t bb[2] = {(uint32_t)p.cbe, nb_box};
      uint32_t be[2] = {1, 1};
      out.specB = make_spec(dtA, 2, bd, bs, bb, be, sw, "weights");
    }
  }
  const int out_fast = fw ? (int)a.OW : (int)a.OH, out_slow = fw ? (int)a.OH : (int)a.OW;
  {
    const uint64_t pixY = (uint64_t)a.Cy * ysz;
    const uint64_t rowY = pixY * a.YW;
    // output pixel (oh, ow) lives at (oh*osh + oph, ow*osw + opw)
    const uint64_t sY_h = rowY * a.osh, sY_w = pixY * a.osw;
    out.d_offset = (long long)((a.y_coff + ((uint64_t)a.oph * a.YW + a.opw) * a.Cy) * ysz);
    const int cstore = std::min(round_up(Cout, 16 / ysz), (int)(a.Cy - a.y_coff));
    const uint64_t sY_fast = fw ? sY_w : sY_h, sY_slow = fw ? sY_h : sY_w;
    uint64_t dims[5] = {(uint64_t)cstore, (uint64_t)out_fast, (uint64_t)(out_slow / g), (uint64_t)g, (uint64_t)a.N};
    uint64_t str[5] = {(uint64_t)ysz, sY_fast, sY_slow * g, sY_slow, pixY * a.YH * a.YW};
    // (a store covers ONE epilogue warp's 32 accumulator rows of the FB x SB pixel tile)
    // (per-warp stores: measured no gain in the step, 4x the store instructions -- off unless SOS_WARP_STORE=1)
    static const int warp_store = getenv("SOS_WARP_STORE") && atoi(getenv("SOS_WARP_STORE")) == 1;
    p.wstore = warp_store;
    uint32_t box[5] = {(uint32_t)(p.ec * p.sw), (uint32_t)(warp_store ? std::min(pl.FB, 32) : pl.FB), (uint32_t)(warp_store ? std::max(1, 32 / pl.FB) : pl.SB), 1, 1};
    uint32_t es[5] = {1, 1, 1, 1, 1};
    const int erow = ec * p.sw * ysz;
    out.specD = make_spec(ysz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dims, str, box, es,
                          erow == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (erow == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), "output");
  }
  const int tiles_fast = ceil_div(out_fast, pl.FB);
  p.out_fast = out_fast;
  p.lat_slow = out_slow / g;
  p.tiles_fast_g = ceil_div(tiles_fast, p.S);
  p.tiles_slow = ceil_div(out_slow / g, pl.SB);
  p.n_phase = g;
  const long long total = (long long)a.N * g * p.tiles_slow * p.tiles_fast_g * n_nblk;
  SOS_CHECK_ARG(total < (1ll << 31), "sos_conv2d_tc: too many tiles");
  p.total_ctiles = (int)total;

  out.smem = 1024 + p.n_stages * p.stage_bytes + p.staging_bytes + 512 + stats_smem;
  out.grid = (int)std::min<long long>(total, sos_num_sms());
  if (p.pair) out.grid = -1;                      // (sized at the first launch: whole clusters that are co-resident)
  out.plan_out[0] = pl.fast_is_w;
  out.plan_out[1] = pl.share;
  out.plan_out[2] = pl.g;
  out.plan_out[3] = p.S;
  out.plan_out[4] = p.n_groups;
  out.plan_out[5] = p.n_stages;
  out.plan_out[6] = p.stage_bytes;
  out.plan_out[7] = out.grid;
  return SOS_OK;
}

}  // namespace

extern "C" void sos_plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  if (hits) *hits = g_plan_hits;
  if (misses) *misses = g_plan_misses;
  if (entries) *entries = (int64_t)g_plans.size();
}

// conv_row.cu: the row-streaming stacked-tap kernel for the narrow dilated k x k layers
bool sos_rowconv_eligible(const sos_conv_args& a);
int sos_rowconv_launch(const sos_conv_args& a, cudaStream_t stream);

extern "C" int sos_conv2d_tc(const sos_conv_args* ap, cudaStream_t stream) {
  SOS_CHECK_ARG(ap != nullptr, "sos_conv2d_tc: null args");
  const sos_conv_args& a = *ap;
  SOS_CHECK_ARG(a.x && a.wk && a.y && a.tap_dh && a.tap_dw, "sos_conv2d_tc: null pointer");
  SOS_CHECK_ARG((a.x_dtype == SOS_DTYPE_TF32 || a.x_dtype == SOS_DTYPE_F16) && (a.y_dtype == SOS_DTYPE_TF32 || a.y_dtype == SOS_DTYPE_F16),
                "sos_conv2d_tc: unknown operand / output type");
  const int ysz = a.y_dtype == SOS_DTYPE_F16 ? 2 : 4;
  // (a half map of 8 channels is legal: the 16-channel TMA boxes read the missing K half as zeros)
  SOS_CHECK_ARG(a.N > 0 && a.H > 0 && a.W > 0 && a.Cin >= 8 && a.Cin % 8 == 0, "sos_conv2d_tc: Cin must be a positive multiple of 8 (got %lld)",
                (long long)a.Cin);
  SOS_CHECK_ARG(a.Cout > 0 && a.OH > 0 && a.OW > 0 && a.ntaps > 0 && a.ntaps <= 49, "sos_conv2d_tc: bad output shape / taps");
  SOS_CHECK_ARG(a.stride == 1 || a.stride == 2, "sos_conv2d_tc: stride must be 1 or 2");
  SOS_CHECK_ARG(a.osh >= 1 && a.osw >= 1 && a.oph >= 0 && a.opw >= 0 && a.oph < a.osh && a.opw < a.osw, "sos_conv2d_tc: bad output lattice");
  SOS_CHECK_ARG(a.Cy % (16 / ysz) == 0 && a.y_coff % (16 / ysz) == 0 && a.y_coff < a.Cy, "sos_conv2d_tc: output channels / offset must be multiples of 16 bytes");
  SOS_CHECK_ARG(((uintptr_t)a.x % 16) == 0 && ((uintptr_t)a.y % 16) == 0 && ((uintptr_t)a.wk % 16) == 0, "sos_conv2d_tc: pointers must be 16-byte aligned");
  SOS_CHECK_ARG((a.OH - 1) * a.osh + a.oph < a.YH && (a.OW - 1) * a.osw + a.opw < a.YW, "sos_conv2d_tc: output lattice exceeds the output buffer");
  SOS_CHECK_ARG(a.epi_scale == nullptr || a.epi_shift != nullptr, "sos_conv2d_tc: epi_scale needs epi_shift (a shift alone is a bias)");
  SOS_CHECK_ARG(a.stats_partial == nullptr || (a.epi_shift == nullptr && (a.act & SOS_ACT_MASK) == 0),
                "sos_conv2d_tc: fused BatchNorm statistics are taken of the RAW outputs (no affine / activation in the same call)");

  if (sos_rowconv_eligible(a)) return sos_rowconv_launch(a, stream);
  if (a.bnr_rows_out) *a.bnr_rows_out = 0;             // (the tap GEMM has no fused BatchNorm-backward reduction: the caller runs pass 1 itself)

  std::vector<int32_t> key;
  key.reserve(24 + 2 * (size_t)a.ntaps);
  const int64_t fields[] = {a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.ntaps, a.stride, a.YH, a.YW, a.Cy, a.y_coff, a.osh, a.osw, a.oph, a.opw,
                            a.x_dtype, a.y_dtype, a.force_plan, a.stats_partial ? a.stats_channels : -1};
  for (int64_t f : fields) key.push_back((int32_t)f);
  for (int t = 0; t < a.ntaps; ++t) { key.push_back(a.tap_dh[t]); key.push_back(a.tap_dw[t]); }

  std::lock_guard<std::mutex> lock(g_plan_mutex);     // (held through the launch: the cached parameter block is patched in place)
  TcPlan* plan;
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    plan = it->second;
    ++g_plan_hits;
  } else {
    plan = new TcPlan();
    if (int e = plan_conv2d_tc(a, *plan)) { delete plan; return e; }
    g_plans.emplace(std::move(key), plan);
    ++g_plan_misses;
  }
  TcParams& p = plan->p;
  p.out_scale = a.out_scale;
  p.scale = a.epi_scale;
  p.shift = a.epi_shift;
  p.act = (int)a.act;
  p.slope = a.slope;
  p.stats = a.stats_partial;
  const void* baseD = reinterpret_cast<const uint8_t*>(a.y) + plan->d_offset;
  if (plan->baseA != a.x) { if (int e = encode_spec(&p.mapA, plan->specA, a.x)) return e; plan->baseA = a.x; }
  if (plan->baseB != a.wk) { if (int e = encode_spec(&p.mapB, plan->specB, a.wk)) return e; plan->baseB = a.wk; }
  if (plan->baseD != baseD) { if (int e = encode_spec(&p.mapD, plan->specD, baseD)) return e; plan->baseD = baseD; }

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tapgemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess ||
        cudaFuncSetAttribute(tapgemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess) {
      sos_set_error("sos_conv2d_tc: cannot raise dynamic shared memory to %d bytes: %s", kSmemLimit, cudaGetErrorString(cudaGetLastError()));
      return SOS_ERR_CUDA;
    }
    attr_set = true;
  }
  if (p.pair) {
    static int max_clusters = 0;
    if (max_clusters == 0) {
      if (cudaFuncSetAttribute(tapgemm_f16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess) {
        sos_set_error("sos_conv2d_tc: cannot raise dynamic shared memory of the pair kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return SOS_ERR_CUDA;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sos_num_sms());
      cfg.blockDim = dim3(kThreadsTc);
      cfg.dynamicSmemBytes = kSmemLimit;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, tapgemm_f16_pair_kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sos_num_sms() / 2;
      }
      max_clusters = n;
    }
    if (plan->grid < 0) plan->grid = 2 * (int)std::min<long long>((p.total_ctiles + 1) / 2, max_clusters);
    tapgemm_f16_pair_kernel<<<plan->grid, kThreadsTc, plan->smem, stream>>>(p);
  } else if (plan->esz == 2) tapgemm_f16_kernel<<<plan->grid, kThreadsTc, plan->smem, stream>>>(p);
  else tapgemm_tf32_kernel<<<plan->grid, kThreadsTc, plan->smem, stream>>>(p);
  SOS_CHECK_LAUNCH("sos_conv2d_tc");
  if (a.stats_rows_out) *a.stats_rows_out = plan->grid;
  if (a.plan_out) memcpy(a.plan_out, plan->plan_out, sizeof(plan->plan_out));
  return SOS_OK;
}

extern "C" int sos_conv_stats_rows(void) { return sos_num_sms(); }

// Host-only planner query (no CUDA call): what sos_conv2d_tc would choose for these shapes / taps / types.
extern "C" int sos_conv2d_plan(const sos_conv_args* ap, int32_t* info) {
  SOS_CHECK_ARG(ap != nullptr && info != nullptr && ap->tap_dh && ap->tap_dw, "sos_conv2d_plan: null pointer");
  SOS_CHECK_ARG(ap->ntaps > 0 && ap->ntaps <= 49 && ap->Cin >= 8 && ap->Cin % 8 == 0 && ap->Cout > 0, "sos_conv2d_plan: bad shapes");
  if (ap->N > 0 && ap->H > 0 && ap->W > 0 && sos_rowconv_eligible(*ap)) {      // served by the row-streaming kernel (conv_row.cu)
    memset(info, 0, 16 * sizeof(int32_t));
    info[0] = 2;
    info[10] = (int32_t)ap->Cout;
    return SOS_OK;
  }
  TcPlan plan;
  if (int e = plan_conv2d_tc(*ap, plan)) return e;
  const TcParams& p = plan.p;
  const int32_t v[16] = {plan.plan_out[0], plan.plan_out[1] + 2 * (p.sw == 2) + 4 * (p.bmerge != 0), plan.plan_out[2], p.S, p.n_groups, p.n_stages, p.stage_bytes, plan.grid,
                         p.cbe, p.n_chunks, p.N, p.ec, p.FB, p.SB, p.n_stg + 100 * p.pair, plan.smem};
  memcpy(info, v, sizeof(v));
  return SOS_OK;
}
