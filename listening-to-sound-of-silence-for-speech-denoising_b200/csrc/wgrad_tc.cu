// Weight gradient of the tap-list operator on the 5th-gen tensor cores (tcgen05 kind::tf32, fp32 accumulate in TMEM).
//
//     dw[t][co][ci] += sum_{pixels p} dy[p][co] * x[p*s + off_t][ci]
//
// is a GEMM whose reduction axis is the pixel axis, so both operands are consumed "MN-major": a staged NHWC box
// [pixel][32-channel chunk] is the canonical MN-major tile of 32-bit operands, whose only legal shared-memory layout is
// the 128-byte swizzle with 32-byte atoms (descriptor layout type 1 = SWIZZLE_128B_BASE32B, TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): one 128-byte row per k index, 4 rows per swizzle atom.  Channel counts that are
// not multiples of 32 use the same 32-wide boxes; TMA zero-fills the out-of-range channels.
//   A = dy tile   (M = 128 output channels of one channel block, K = 8 pixels per MMA)
//   B = x box     (N = Cin, same K) -- one box per tap group, shared by the taps that differ by a slow-axis shift
//   D = dw[t] block, one TMEM accumulator (N columns) per tap, up to 512 columns per job
// Work item = (job, pixel slice): a job is a set of tap groups of one output-channel block whose accumulators fit in
// TMEM; the pixel tiles are split into slices so that all SMs are busy and jobs of the same slice run concurrently
// (their operand boxes hit in L2).  Partial sums are merged with fp32 atomics (red.global.add).
//
// Stacked mode (32 < Cout <= 64): an M = 128 MMA would be at least half empty, so accumulator rows 64..127 are given to a
// SECOND tap: the dy box shifted by `delta` pixels along the fast axis (same tensor map, box origin + delta; the row tail reads
// as zero) is staged as M chunks 2-3 next to dy[p] in chunks 0-1.  With the x box of the tap group at fast offset f, rows 0..63 then accumulate
// that group's taps and rows 64..127 the taps of the group at fast offset f - delta: one MMA serves two taps (5 x 5 taps ->
// 15 accumulators instead of 25).  The shifted half sees dy columns [delta, W + delta), so the tile grid is extended to
// negative fast coordinates (zero-filled) to cover columns [0, delta) as well.
//
// CTA = 192 threads: warp 0 TMA producer, warps 1-5 MMA issuers (one elected thread each, accumulator slots dealt round-robin),
// warps 2-5 also drain TMEM at the end of each work item.
#include "common.cuh"
#include "ptx.cuh"
#include "sos_b200.h"
#include "tc_common.cuh"
#include <map>
#include <mutex>

namespace {

using namespace ptx;
using namespace tc;

constexpr int kThreadsWg = 192;
constexpr int kIssuers = 5;                // warps 1..5 each elect one MMA-issuing thread
constexpr int kMaxJobs = 100;
constexpr int kMaxSlots = 64;
constexpr int kMaxItems = 512;

// One TMEM accumulator of a job.  Narrow slot: one tap, N = round16(Cin) columns.  Wide slot (Cin <= 64, half operands): up to four
// taps of ONE tap group that differ by equally spaced slow-axis shifts share the MMA -- the x operand of tap j0 + c is N chunk c (64
// channels) of the same staged box, `lbo` bytes further on, so the descriptor's chunk stride addresses all of them: N = 64 * ntaps,
// the dy operand (4 KB per MMA) is read once for four taps (ncu r02: the 48-channel weight gradients were shared-memory bound at
// 48 wavefronts per 24-cycle N = 48 MMA).  Accumulator rows 0..63 (all 128 when not stacked) receive taps tap_lo[], rows 64..127
// taps tap_hi[] (-1: unused).
struct WgSlot {
  int16_t x_idx, ntaps, col0, n_cols;
  uint32_t b_desc;               // ((byte offset of the x operand inside the stage's x area) >> 4) | (chunk stride >> 4) << 16
  uint32_t idesc;
  int16_t tap_lo[4], tap_hi[4];
};

struct alignas(64) WgParams {
  CUtensorMap mapX, mapDY;
  int stacked, delta, tf_extra;      // delta: fast-axis shift of the stacked dy box; tf_extra: tile columns added at negative fast coordinates
  int n_jobs, n_items, total_tiles;
  int tiles_fast, tiles_slow, n_phase;
  int FB, SB, stride;
  int Cin, Cout, N;
  int cbi, n_ci_chunks, cbo;
  int ksteps, shift_mul;            // 8-pixel k steps per tile; k steps per slow row (tap shifts are whole slow rows)
  int x_box_bytes, x_box_stride, dy_chunk_bytes, dy_chunk_stride, x_off;
  int stage_bytes, n_stages;
  int layout_a, layout_b;
  int kstep_bytes;                  // operand bytes of one k step: (8 | 16 pixels) x 128-byte rows
  int row_bytes;                    // bytes of one slow row of a staged box (FB pixels x 128)
  int m_half_chunks;                // channel chunks per 64 accumulator rows (stacked mode stages the shifted box there)
  uint32_t desc_hi;                 // descriptor high word of the dy operand: SBO | version | layout
  uint32_t desc_hi_b;               // ... of the x operand (union mode: its SBO is the union box's row pitch)
  uint32_t a_inc16, b_inc16;        // descriptor advance per k step (16-byte units)
  uint32_t a_lbo16;                 // chunk stride of the dy operand (16-byte units)
  // Union mode (Cin <= 64, 8-pixel tiles, fast-axis tap span < 32 pixels): ONE x box of FB + span pixels per slow row serves every
  // fast offset -- a tap's operand starts `d_fast` pixels (128-byte rows) into it, an address that is not swizzle-atom aligned, which
  // the tensor core handles (it XORs the absolute address bits TMA used; see conv_row.cu) -- and, stacked, ONE dy box of FB + delta
  // pixels serves both the plain and the shifted half of the M rows.  L2 -> shared-memory traffic per tile: 164 -> 96 KB on the
  // 48-channel 5x5 layers.
  int uni, fbu_x, fbu_dy, xmin_fast, xmin_slow, dy_box_bytes;
  int dbg;                          // SOS_WGRAD_DBG=1: role timers (cycles waiting on barriers) printed by block 0
  const float* out_scale;           // optional device scalar multiplied into the sums
  float* dw;
  int16_t job_coblk[kMaxJobs], job_g0[kMaxJobs], job_ng[kMaxJobs], job_s0[kMaxJobs], job_ns[kMaxJobs];
  // a job's pixel tiles are split into job_nsl[j] slices of job_tps[j] tiles; the slice counts are proportional to the jobs' MMA
  // counts, so that every work item (job, slice) costs about the same.  Items are ordered by their relative position in the
  // tile stream, so that the jobs working on the same pixels run at the same time (their operand boxes then hit in L2).
  int16_t job_nsl[kMaxJobs];
  int32_t job_tps[kMaxJobs];
  int16_t item_job[kMaxItems], item_slice[kMaxItems];
  int16_t job_group[kMaxJobs * 4];   // tap groups whose x boxes a job stages (job_g0 = first index, job_ng = count)
  WgSlot slots[kMaxSlots];
  TapGroup groups[kMaxGroups];
};

struct WgTile { int tf, ts, ph, n; };
__device__ __forceinline__ WgTile decode_wg_tile(const WgParams& p, int t) {
  WgTile c;
  c.tf = t % p.tiles_fast - p.tf_extra; t /= p.tiles_fast;
  c.ts = t % p.tiles_slow; t /= p.tiles_slow;
  c.ph = t % p.n_phase;
  c.n = t / p.n_phase;
  return c;
}

struct WgItem { int job, t0, t1; };
__device__ __forceinline__ WgItem decode_wg_item(const WgParams& p, int wi) {
  const int j = p.item_job[wi], slice = p.item_slice[wi];
  WgItem it;
  it.job = j;
  it.t0 = slice * p.job_tps[j];
  it.t1 = min(p.total_tiles, it.t0 + p.job_tps[j]);
  return it;
}

template <int KIND>
__device__ __forceinline__ void wgrad_body(const WgParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // barriers live at the front (the operand stages may be over-read by the padded MMA rows, never the barriers)
  const uint32_t bar_base = smem_base;
  const uint32_t stages_base = smem_base + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  const uint32_t tfull_bar = bar_base + 8u * 32, tempty_bar = bar_base + 8u * 33;
  const uint32_t tmem_slot = bar_base + 8u * 34;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapX);
    prefetch_tmap(&p.mapDY);
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kIssuers);
    }
    mbar_init(tfull_bar, kIssuers);
    mbar_init(tempty_bar, 4);
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
  const int n_items = p.n_items;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    const long long t_begin = clock64();
    for (int wi = blockIdx.x; wi < n_items; wi += gridDim.x) {
      const WgItem it = decode_wg_item(p, wi);
      const int job = it.job, t0 = it.t0, t1 = it.t1;
      const int coblk = p.job_coblk[job], g0 = p.job_g0[job], ng = p.job_ng[job];
      const int co_here = min(128, p.Cout - coblk * 128);
      const int n_co_chunks = (co_here + p.cbo - 1) / p.cbo;
      const uint32_t tx = p.uni ? (uint32_t)n_co_chunks * p.dy_box_bytes + (uint32_t)p.x_box_bytes
                                : (uint32_t)n_co_chunks * (p.stacked ? 2 : 1) * p.dy_chunk_bytes + (uint32_t)ng * p.n_ci_chunks * p.x_box_bytes;
      for (int tile = t0; tile < t1; ++tile) {
        const WgTile tc = decode_wg_tile(p, tile);
        const long long tw = p.dbg ? clock64() : 0;
        mbar_wait(empty_bar(stage), phase ^ 1, 500);
        if (p.dbg) w_empty += clock64() - tw;
        if (elect_one_sync()) {
          mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sbase = stages_base + (uint32_t)stage * p.stage_bytes;
          if (p.uni) {
            for (int cc = 0; cc < n_co_chunks; ++cc)
              tma_load_5d(sbase + (uint32_t)cc * p.dy_chunk_stride, &p.mapDY, full_bar(stage), coblk * 128 + cc * p.cbo, tc.tf * p.FB, tc.ts * p.SB, tc.ph,
                          tc.n);
            tma_load_5d(sbase + p.x_off, &p.mapX, full_bar(stage), 0, tc.tf * p.FB + p.xmin_fast, tc.ts * p.SB + p.xmin_slow, tc.ph, tc.n);
            goto loaded;
          }
          for (int cc = 0; cc < n_co_chunks; ++cc)
            tma_load_5d(sbase + (uint32_t)cc * p.dy_chunk_stride, &p.mapDY, full_bar(stage), coblk * 128 + cc * p.cbo, tc.tf * p.FB,
                        tc.ts * p.SB, tc.ph, tc.n);
          if (p.stacked)
            for (int cc = 0; cc < n_co_chunks; ++cc)     // M chunks 2, 3: the same channels, `delta` pixels further along the fast axis
              tma_load_5d(sbase + (uint32_t)(p.m_half_chunks + cc) * p.dy_chunk_stride, &p.mapDY, full_bar(stage), cc * p.cbo, tc.tf * p.FB + p.delta,
                          tc.ts * p.SB, tc.ph, tc.n);
          for (int gi = 0; gi < ng; ++gi) {
            const TapGroup& grp = p.groups[p.job_group[g0 + gi]];
            const uint32_t xb = sbase + p.x_off + (uint32_t)gi * p.n_ci_chunks * p.x_box_stride;
            for (int c = 0; c < p.n_ci_chunks; ++c)
              tma_load_5d(xb + (uint32_t)c * p.x_box_stride, &p.mapX, full_bar(stage), c * p.cbi, tc.tf * p.FB * p.stride + grp.d_fast,
                          tc.ts * p.SB * p.stride + grp.d_slow, tc.ph, tc.n);
          }
        loaded:;
        }
        __syncwarp();
        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
      }
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) printf("wgrad dbg: producer %lld cycles, %lld waiting for empty stages\n", clock64() - t_begin, w_empty);
  } else {
    // ===================================================================== MMA issuers (warps 1..5) + drain (warps 2..5)
    // The accumulator slots of a job are independent, so they are dealt round-robin to kIssuers elected threads (one per warp):
    // an N = 48 MMA lasts 24 cycles but costs one thread ~75 cycles of uniform-datapath instructions (ncu r02: the single
    // issuer of the 48-channel layers ran at 620 cycles per 8-MMA slot against 192 cycles of tensor time).  The drain warps are
    // idle while a work item's tiles stream through, so they issue as well.
    const int issuer = warp - 1;                  // 0..4
    const int q = warp & 3;                       // TMEM lane quadrant (drain)
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t a_inc = p.a_inc16, b_inc = p.b_inc16;                       // one k step (8 or 16 pixels) of one chunk
    const uint32_t desc_hi = p.desc_hi, desc_hi_b = p.desc_hi_b;
    const int ksteps = p.ksteps, n_stages = p.n_stages;
    const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
    // descriptor words: hi = SBO | version | layout; lo = (address >> 4) | LBO (chunk pitch); one k step = 8 / 16 pixels
    const uint32_t a_lo_base = (stages_base >> 4) | (p.a_lbo16 << 16);
    const uint32_t b_lo_base = (stages_base + (uint32_t)p.x_off) >> 4;      // (+ the slot's offset and chunk stride)
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    long long w_te = 0, w_full = 0, w_drain = 0;
    const long long t_begin = clock64();
    for (int wi = blockIdx.x; wi < n_items; wi += gridDim.x) {
      const WgItem it = decode_wg_item(p, wi);
      const int job = it.job, t0 = it.t0, t1 = it.t1;
      if (t0 >= t1) continue;
      const int coblk = p.job_coblk[job], s0 = p.job_s0[job], ns = p.job_ns[job];
      if (elect_one_sync()) {
        long long tw = p.dbg ? clock64() : 0;
        mbar_wait(tempty_bar, acc_phase ^ 1, 600);
        if (p.dbg) w_te += clock64() - tw;
        tc_fence_after();
        uint32_t first = 0u;
        int st = stage;
        uint32_t ph = phase;
        for (int tile = t0; tile < t1; ++tile) {
          tw = p.dbg ? clock64() : 0;
          mbar_wait(full_bar(st), ph, 601);
          if (p.dbg) w_full += clock64() - tw;
          tc_fence_after();
          const uint32_t a_lo0 = a_lo_base + (uint32_t)st * stage_step, b_lo0 = b_lo_base + (uint32_t)st * stage_step;
          for (int i = issuer; i < ns; i += kIssuers) {
            const uint32_t d = tmem_base + (uint32_t)p.slots[s0 + i].col0;
            const uint32_t idesc = p.slots[s0 + i].idesc;
            uint32_t a_lo = a_lo0;
            uint32_t b_lo = b_lo0 + p.slots[s0 + i].b_desc;
#pragma unroll 8
            for (int ks = 0; ks < ksteps; ++ks) {
              umma<KIND>(d, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi_b << 32) | b_lo, idesc, first | (uint32_t)ks);
              a_lo += a_inc;
              b_lo += b_inc;
            }
          }
          umma_commit(empty_bar(st));               // (an issuer without slots in this job still arrives)
          first = 1u;
          if (++st == n_stages) { st = 0; ph ^= 1; }
        }
        umma_commit(tfull_bar);
      }
      __syncwarp();
      // every lane tracks the pipeline position the elected lane advanced through
      {
        const int adv = t1 - t0;
        const int tot = stage + adv;
        phase ^= (uint32_t)((tot / n_stages) & 1);
        stage = tot % n_stages;
      }
      if (warp >= 2) {
        // ---- drain: accumulator row q*32 + lane = output channel, and (stacked) which of the slot's two taps
        const int co = p.stacked ? (q & 1) * 32 + lane : coblk * 128 + q * 32 + lane;
        const long long td = p.dbg ? clock64() : 0;
        mbar_wait(tfull_bar, acc_phase, 700);
        tc_fence_after();
        for (int i = 0; i < ns; ++i) {
          const WgSlot& sl = p.slots[s0 + i];
          const bool wide = sl.ntaps > 1;
          for (int c0 = 0; c0 < sl.n_cols; c0 += 16) {
            const int sub = wide ? (c0 >> 6) : 0, ci0 = wide ? (c0 & 63) : c0;
            const int tap = (p.stacked && (q >> 1)) ? sl.tap_hi[sub] : sl.tap_lo[sub];
            if (ci0 >= p.Cin) continue;                                  // (padding columns of a 64-wide chunk)
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sl.col0 + c0), r);
            tmem_ld_wait();
            if (co < p.Cout && tap >= 0) {
              float* dst = p.dw + ((size_t)tap * p.Cout + co) * p.Cin + ci0;
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (ci0 + k < p.Cin) atomicAdd(dst + k, __uint_as_float(r[k]) * oscale);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar);
        if (p.dbg) w_drain += clock64() - td;
      }
      acc_phase ^= 1;
    }
    // (w_te / w_full are the elected lane's; it may differ from lane 0, so every lane prints its own non-zero counters)
    if (p.dbg && blockIdx.x == 0 && (w_full != 0 || (lane == 0 && warp == 2)))
      printf("wgrad dbg: warp %d lane %d: %lld cycles, %lld waiting for the accumulator, %lld for full stages; drain (incl. wait for the last MMA) %lld\n",
             warp, lane, clock64() - t_begin, w_te, w_full, w_drain);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(kThreadsWg, 1) wgrad_tf32_kernel(const __grid_constant__ WgParams p) { wgrad_body<0>(p); }
__global__ void __launch_bounds__(kThreadsWg, 1) wgrad_f16_kernel(const __grid_constant__ WgParams p) { wgrad_body<1>(p); }

struct WgPlan {
  WgParams p;
  MapSpec specX, specDY;
  const void *baseX = nullptr, *baseDY = nullptr;
  int smem = 0, grid = 0, esz = 4;
  int32_t plan_out[8] = {0};
};
std::mutex g_wg_mutex;
std::map<std::vector<int32_t>, WgPlan*> g_wg_plans;

// Everything that does not depend on the call's pointers (see conv_tc.cu: plan cache).
int plan_conv2d_wgrad(const sos_wgrad_args& a, WgPlan& out) {
  const int esz = a.dtype == SOS_DTYPE_F16 ? 2 : 4;
  const int kpix = 32 / esz;                 // pixels (K) per MMA
  const int cchunk = 128 / esz;              // channels per 128-byte row of a staged box
  out.esz = esz;
  Geometry geo{(int)a.ntaps, a.tap_dh, a.tap_dw, (int)a.H, (int)a.W, (int)a.OH, (int)a.OW, (int)a.stride, (int)a.Cin, (int)a.Cout, true,
               std::min(kMaxSub, 512 / round_up((int)a.Cin, 16)), true, esz};
  Plan best;
  for (int fw = 1; fw >= 0; --fw)
    for (int sh = 1; sh >= 0; --sh) {
      Plan pl;
      if (build_plan(geo, fw != 0, sh != 0, pl) && pl.FB <= 16 && pl.cost < best.cost) best = pl;
    }
  if (a.force_plan >= 0) {
    Plan pl;
    SOS_CHECK_ARG(build_plan(geo, (a.force_plan & 1) != 0, (a.force_plan & 2) != 0, pl) && pl.FB <= 16,
                  "sos_conv2d_wgrad: forced plan %d not applicable", (int)a.force_plan);
    best = pl;
  }
  SOS_CHECK_ARG(best.cost < 1e299, "sos_conv2d_wgrad: no feasible plan");
  const Plan& pl = best;

  WgParams& p = out.p;
  memset(&p, 0, sizeof(p));
  const int Cin = (int)a.Cin, Cout = (int)a.Cout;
  p.Cin = Cin;
  p.Cout = Cout;
  p.N = round_up(Cin, 16);
  p.cbi = cchunk;
  p.n_ci_chunks = ceil_div(Cin, cchunk);
  p.cbo = cchunk;
  p.FB = pl.FB;                      // 8, or 16 for short dilation lattices (then SB <= 8)
  p.stride = (int)a.stride;
  // TF32: 128-byte swizzle with 32-byte atoms (4 k rows per atom, SBO 512); half: plain 128-byte swizzle (8 k rows, SBO 1024)
  p.layout_a = p.layout_b = esz == 2 ? 2 : 1;
  p.kstep_bytes = kpix * 128;
  p.row_bytes = p.FB * 128;
  p.m_half_chunks = 64 / cchunk;

  // ---- jobs: tap groups packed into TMEM (512 columns) per output-channel block; at most kMaxJobGroups boxes per stage
  const int n_coblk = ceil_div(Cout, 128);
  const int n_groups = (int)pl.groups.size();
  SOS_CHECK_ARG(n_groups <= kMaxGroups, "sos_conv2d_wgrad: too many tap groups");
  for (int i = 0; i < n_groups; ++i) {
    p.groups[i] = pl.groups[i];
  }
  const int co_blk_ch = std::min(128, Cout);
  const int n_co_chunks_max = ceil_div(co_blk_ch, p.cbo);
  // ---- stacked mode: pair tap groups whose fast offsets differ by `delta` and whose sub-tap structure is identical
  int delta = 0;
  {
    int best_d = 1 << 30;
    for (int i = 0; i < n_groups; ++i)
      for (int j = 0; j < n_groups; ++j) {
        const int d = pl.groups[i].d_fast - pl.groups[j].d_fast;
        if (d > 0 && d < best_d) best_d = d;
      }
    if (best_d < (1 << 30)) delta = best_d;
  }
  const int out_fast_px = pl.fast_is_w ? (int)a.OW : (int)a.OH;
  bool stacked = Cout > 32 && Cout <= 64 && a.stride == 1 && delta > 0 && delta < out_fast_px && (a.force_plan < 0 || a.force_plan < 4);
  std::vector<int> partner(n_groups, -1), is_upper(n_groups, 0);
  if (stacked) {
    bool any = false;
    std::vector<int> order(n_groups);
    for (int i = 0; i < n_groups; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return pl.groups[x].d_fast > pl.groups[y].d_fast; });
    for (int oi = 0; oi < n_groups; ++oi) {
      const int A = order[oi];
      if (is_upper[A]) continue;
      for (int B = 0; B < n_groups; ++B) {
        const TapGroup &ga = pl.groups[A], &gb = pl.groups[B];
        if (B == A || is_upper[B] || partner[B] >= 0 || gb.d_fast != ga.d_fast - delta || gb.d_slow != ga.d_slow || gb.n_sub != ga.n_sub ||
            memcmp(gb.a_off, ga.a_off, ga.n_sub) != 0)
          continue;
        partner[A] = B;
        is_upper[B] = 1;
        any = true;
        break;
      }
    }
    stacked = any;
    if (!stacked) std::fill(is_upper.begin(), is_upper.end(), 0);
  }
  p.stacked = stacked;
  p.delta = stacked ? delta : 0;
  p.tf_extra = stacked ? ceil_div(delta, p.FB) : 0;
  // "lead" groups stage an x box and own accumulator slots; the upper partner of a pair rides in rows 64..127
  std::vector<int> leads;
  for (int i = 0; i < n_groups; ++i)
    if (!is_upper[i]) leads.push_back(i);
  const int n_leads = (int)leads.size();
  // accumulator slots of every lead group: wide slots (up to four equally spaced taps per MMA) where the operands allow
  struct SlotDesc { int li, j0, nt, cols; };
  std::vector<SlotDesc> sd;
  static const int narrow_only = getenv("SOS_WGRAD_NARROW") && atoi(getenv("SOS_WGRAD_NARROW")) == 1;     // A/B aid
  const bool can_wide = esz == 2 && p.n_ci_chunks == 1 && !narrow_only && a.force_plan < 0;
  // union boxes: one x box spanning every group's fast offset (and, stacked, one dy box spanning the shift) -- see WgParams::uni
  int xmin_fast = 0, xmax_fast = 0, xmin_slow = 0, halo_u = 0;
  for (int i = 0; i < n_groups; ++i) {
    xmin_fast = i ? std::min(xmin_fast, (int)pl.groups[i].d_fast) : (int)pl.groups[i].d_fast;
    xmax_fast = i ? std::max(xmax_fast, (int)pl.groups[i].d_fast) : (int)pl.groups[i].d_fast;
    xmin_slow = i ? std::min(xmin_slow, (int)pl.groups[i].d_slow) : (int)pl.groups[i].d_slow;
  }
  for (int i = 0; i < n_groups; ++i)
    for (int j = 0; j < pl.groups[i].n_sub; ++j) halo_u = std::max(halo_u, pl.groups[i].d_slow - xmin_slow + pl.groups[i].a_off[j]);
  static const int no_union = getenv("SOS_WGRAD_NO_UNION") && atoi(getenv("SOS_WGRAD_NO_UNION")) == 1;      // A/B aid
  const int fbu_x = p.FB + (xmax_fast - xmin_fast), fbu_dy = p.FB + (stacked ? delta : 0);
  const bool uni = can_wide && !no_union && pl.share && a.stride == 1 && p.FB == 8 && n_groups > 1 && fbu_x < n_groups * p.FB && fbu_x <= 256 &&
                   (!stacked || n_co_chunks_max == 1);
  p.uni = uni;
  p.fbu_x = fbu_x; p.fbu_dy = fbu_dy; p.xmin_fast = xmin_fast; p.xmin_slow = xmin_slow;
  bool any_wide = false;
  std::vector<int> spacing(n_leads, 0);
  for (int li = 0; li < n_leads; ++li) {
    const TapGroup& ga = pl.groups[leads[li]];
    bool eq = can_wide && ga.n_sub > 1;
    const int da = ga.n_sub > 1 ? ga.a_off[1] - ga.a_off[0] : 0;
    for (int j = 1; j < ga.n_sub && eq; ++j) eq = (ga.a_off[j] - ga.a_off[j - 1]) == da && da > 0;
    spacing[li] = eq ? da : 0;
    if (eq) {
      for (int j = 0; j < ga.n_sub;) {
        const int nt = std::min(4, ga.n_sub - j);
        sd.push_back({li, j, nt, nt == 1 ? p.N : 64 * nt});
        any_wide |= nt > 1;
        j += nt;
      }
    } else {
      for (int j = 0; j < ga.n_sub; ++j) sd.push_back({li, j, 1, p.N});
    }
  }
  for (const SlotDesc& d : sd) SOS_CHECK_ARG(d.cols <= 512, "sos_conv2d_wgrad: accumulator does not fit in TMEM");
  // choose SB (pixels per tile = FB*SB) so that at least 2 stages fit, then how many groups' boxes a stage may hold
  int SB = pl.SB;
  static const int force_sb = getenv("SOS_WGRAD_SB") ? atoi(getenv("SOS_WGRAD_SB")) : 0;     // experiment aid
  if (force_sb > 0 && any_wide && force_sb < SB) SB = force_sb;
  const int avail = kSmemLimit - 2048;
  const int dy_chunks_staged = stacked ? 128 / cchunk : n_co_chunks_max;
  auto stage_bytes_for = [&](int sb, int ng, int* x_off_out, int* reach_out) {
    if (uni) {
      const int dy_stride = round_up(sb * fbu_dy * 128, 1024);
      const int x_off = (stacked ? 1 : n_co_chunks_max) * dy_stride;
      const int used = x_off + round_up((sb + halo_u) * fbu_x * 128, 1024);
      // (the M = 128 rows of a non-stacked dy operand with fewer than 128 channels read one chunk pitch past the staged chunks)
      const int reach_a = stacked ? x_off : (128 / p.cbo) * dy_stride;
      if (x_off_out) *x_off_out = x_off;
      if (reach_out) *reach_out = std::max(reach_a, used);
      return used;
    }
    const int dy_chunk = sb * p.FB * 128;
    const int dy_stride = round_up(dy_chunk, 1024);
    const int x_box = (sb + pl.halo) * p.FB * 128;
    const int x_stride = round_up(x_box, 1024);
    const int x_off = dy_chunks_staged * dy_stride;
    const int used = x_off + ng * p.n_ci_chunks * x_stride;
    // the padded MMA rows (M = 128, N rounded to 16) read past the staged chunks: keep that inside the stage
    const int reach_a = (128 / p.cbo) * dy_stride;
    const int reach_b = x_off + ng * p.n_ci_chunks * x_stride;
    if (x_off_out) *x_off_out = x_off;
    if (reach_out) *reach_out = std::max(reach_a, reach_b);
    return used;
  };
  for (;;) {
    int reach = 0;
    const int used = stage_bytes_for(SB, 1, nullptr, &reach);
    if (2 * used + std::max(0, reach - used) <= avail || SB == 2) break;
    SB /= 2;
  }
  int gcap = uni ? n_leads : 1;                              // groups per job (sharing the staged dy tile) while 3 stages fit
  for (int ng = 2; ng <= std::min(n_leads, 4) && !uni; ++ng) {
    int reach = 0;
    const int used = stage_bytes_for(SB, ng, nullptr, &reach);
    if ((any_wide ? 2 : 3) * used + std::max(0, reach - used) > avail) break;
    gcap = ng;
  }
  // jobs: bins of slots with at most 512 TMEM columns and gcap distinct groups.  Narrow plans keep whole groups together in tap
  // order (as the first implementation did); wide plans are packed first-fit by decreasing width
  struct JobDesc { std::vector<int> groups; std::vector<int> slots; int cols = 0; };
  std::vector<JobDesc> jd;
  {
    std::vector<int> order(sd.size());
    for (size_t i = 0; i < sd.size(); ++i) order[i] = (int)i;
    if (any_wide) std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return sd[x].cols > sd[y].cols; });
    for (int si : order) {
      const SlotDesc& d = sd[si];
      int placed = -1;
      const int first = any_wide ? 0 : std::max(0, (int)jd.size() - 1);      // narrow: only the open (last) job
      for (int j = first; j < (int)jd.size() && placed < 0; ++j) {
        const bool has = std::find(jd[j].groups.begin(), jd[j].groups.end(), d.li) != jd[j].groups.end();
        if (jd[j].cols + d.cols <= 512 && (has || (int)jd[j].groups.size() < gcap)) placed = j;
      }
      if (placed < 0) { jd.emplace_back(); placed = (int)jd.size() - 1; }
      JobDesc& J = jd[placed];
      if (std::find(J.groups.begin(), J.groups.end(), d.li) == J.groups.end()) J.groups.push_back(d.li);
      J.slots.push_back(si);
      J.cols += d.cols;
    }
  }
  int max_ng = 1;
  for (const JobDesc& J : jd) max_ng = std::max(max_ng, (int)J.groups.size());
  int reach = 0;
  const int used = stage_bytes_for(SB, max_ng, &p.x_off, &reach);
  p.SB = SB;
  p.dy_chunk_bytes = SB * p.FB * 128;
  p.dy_chunk_stride = round_up(uni ? SB * fbu_dy * 128 : p.dy_chunk_bytes, 1024);
  p.dy_box_bytes = SB * fbu_dy * 128;
  p.x_box_bytes = uni ? (SB + halo_u) * fbu_x * 128 : (SB + pl.halo) * p.FB * 128;
  // descriptors: SBO = stride between the 8-pixel atoms along K (the next slow row of an 8-pixel-wide tile: the box's row pitch in
  // union mode), advance per k step = kpix pixels
  {
    const uint32_t sbo_plain = esz == 2 ? 1024u : 512u;
    const uint32_t sbo_a = uni ? (uint32_t)fbu_dy * 128u : sbo_plain, sbo_b = uni ? (uint32_t)fbu_x * 128u : sbo_plain;
    p.desc_hi = (sbo_a >> 4) | (1u << 14) | ((uint32_t)p.layout_a << 29);
    p.desc_hi_b = (sbo_b >> 4) | (1u << 14) | ((uint32_t)p.layout_b << 29);
    p.a_inc16 = (uni ? (uint32_t)(kpix / 8) * fbu_dy * 128u : (uint32_t)p.kstep_bytes) >> 4;
    p.b_inc16 = (uni ? (uint32_t)(kpix / 8) * fbu_x * 128u : (uint32_t)p.kstep_bytes) >> 4;
    p.a_lbo16 = ((uni && stacked) ? (uint32_t)delta * 128u : (uint32_t)p.dy_chunk_stride) >> 4;
  }
  SOS_CHECK_ARG((SB * p.FB) % kpix == 0, "sos_conv2d_wgrad: tile of %d x %d pixels is not a multiple of the k step", SB, p.FB);
  p.ksteps = SB * p.FB / kpix;
  p.shift_mul = p.FB / 8;
  p.x_box_stride = round_up(p.x_box_bytes, 1024);
  p.stage_bytes = used;
  const int tail = std::max(0, reach - used);
  p.n_stages = std::min(8, (avail - tail) / used);
  SOS_CHECK_ARG(p.n_stages >= 1, "sos_conv2d_wgrad: stage of %d bytes does not fit in shared memory", used);

  int n_jobs = 0, n_slots = 0, n_jg = 0;
  std::vector<double> job_cost;
  for (int cb = 0; cb < n_coblk; ++cb) {
    for (const JobDesc& J : jd) {
      const int ng = (int)J.groups.size(), nsl = (int)J.slots.size();
      SOS_CHECK_ARG(n_jobs < kMaxJobs && n_slots + nsl <= kMaxSlots && n_jg + ng <= kMaxJobs * 4, "sos_conv2d_wgrad: too many jobs / accumulators");
      p.job_coblk[n_jobs] = (int16_t)cb;
      p.job_g0[n_jobs] = (int16_t)n_jg;
      p.job_ng[n_jobs] = (int16_t)ng;
      p.job_s0[n_jobs] = (int16_t)n_slots;
      p.job_ns[n_jobs] = (int16_t)nsl;
      for (int gi = 0; gi < ng; ++gi) p.job_group[n_jg++] = (int16_t)leads[J.groups[gi]];
      int col = 0;
      double cost = 0;
      for (int si : J.slots) {
        const SlotDesc& d = sd[si];
        const int A = leads[d.li];
        const TapGroup& ga = pl.groups[A];
        const int gi = (int)(std::find(J.groups.begin(), J.groups.end(), d.li) - J.groups.begin());
        WgSlot ws{};
        ws.x_idx = (int16_t)gi;
        ws.ntaps = (int16_t)d.nt;
        ws.col0 = (int16_t)col;
        ws.n_cols = (int16_t)d.cols;
        const uint32_t urow = (uint32_t)fbu_x * 128u;        // union mode: bytes per slow row of the x box
        const uint32_t off = uni ? (uint32_t)(ga.d_slow - xmin_slow + ga.a_off[d.j0]) * urow + (uint32_t)(ga.d_fast - xmin_fast) * 128u
                                 : (uint32_t)gi * p.n_ci_chunks * p.x_box_stride + (uint32_t)ga.a_off[d.j0] * (uint32_t)p.row_bytes;
        const uint32_t lbo = d.nt > 1 ? (uint32_t)spacing[d.li] * (uni ? urow : (uint32_t)p.row_bytes) : (uint32_t)p.x_box_stride;
        SOS_CHECK_ARG((lbo >> 4) < (1u << 14), "sos_conv2d_wgrad: chunk stride too large for the descriptor");
        ws.b_desc = (off >> 4) | ((lbo >> 4) << 16);
        ws.idesc = esz == 2 ? make_idesc_f16(128, d.cols, 1, 1) : make_idesc_tf32(128, d.cols, 1, 1);
        for (int c = 0; c < 4; ++c) {
          ws.tap_lo[c] = (int16_t)(c < d.nt ? ga.tap[d.j0 + c] : -1);
          ws.tap_hi[c] = (int16_t)((c < d.nt && partner[A] >= 0) ? pl.groups[partner[A]].tap[d.j0 + c] : -1);
        }
        p.slots[n_slots++] = ws;
        col += d.cols;
        cost += p.ksteps * std::max(d.cols / 2.0, (4096.0 + 32.0 * round_up(d.cols, 64)) / 128.0);
      }
      job_cost.push_back(cost);
      ++n_jobs;
    }
  }
  p.n_jobs = n_jobs;

  // ---- tensor maps, dim order (channel, fast, slow/g, phase(g), image)
  const bool fw = pl.fast_is_w;
  const int g = pl.g;
  {
    const uint64_t pix = (uint64_t)Cin * esz;
    const uint64_t in_fast = fw ? a.W : a.H, in_slow = fw ? a.H : a.W;
    const uint64_t s_fast = fw ? pix : pix * a.W, s_slow = fw ? pix * a.W : pix;
    uint64_t dims[5] = {(uint64_t)Cin, in_fast, in_slow / g, (uint64_t)g, (uint64_t)a.N};
    uint64_t str[5] = {(uint64_t)esz, s_fast, s_slow * g, s_slow, pix * a.H * a.W};
    uint32_t box[5] = {(uint32_t)p.cbi, (uint32_t)(uni ? fbu_x : p.FB * a.stride), (uint32_t)(uni ? SB + halo_u : (SB + pl.halo) * a.stride), 1, 1};
    uint32_t es[5] = {1, (uint32_t)a.stride, (uint32_t)a.stride, 1, 1};
    SOS_CHECK_ARG(box[2] <= 256, "sos_conv2d_wgrad: activation box too large");
    const CUtensorMapSwizzle sw = esz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    out.specX = make_spec(esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, dims, str, box, es, sw, "wgrad activations");
  }
  const int out_fast = fw ? (int)a.OW : (int)a.OH, out_slow = fw ? (int)a.OH : (int)a.OW;
  {
    const uint64_t pix = (uint64_t)a.Cdy * esz;
    const uint64_t s_fast = fw ? pix : pix * a.OW, s_slow = fw ? pix * a.OW : pix;
    uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)out_fast, (uint64_t)(out_slow / g), (uint64_t)g, (uint64_t)a.N};
    uint64_t str[5] = {(uint64_t)esz, s_fast, s_slow * g, s_slow, pix * a.OH * a.OW};
    uint32_t box[5] = {(uint32_t)p.cbo, (uint32_t)(uni ? fbu_dy : p.FB), (uint32_t)SB, 1, 1};
    uint32_t es[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = esz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    out.specDY = make_spec(esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, dims, str, box, es, sw, "wgrad output grads");
  }
  p.tiles_fast = ceil_div(out_fast, p.FB) + p.tf_extra;
  p.tiles_slow = ceil_div(out_slow / g, SB);
  p.n_phase = g;
  const long long total = (long long)a.N * g * p.tiles_slow * p.tiles_fast;
  SOS_CHECK_ARG(total < (1ll << 30), "sos_conv2d_wgrad: too many tiles");
  p.total_tiles = (int)total;
  const int sms = sos_num_sms();
  {
    // slices per job proportional to its cost per tile = max(tensor / shared-memory time of its MMAs, L2 -> smem time of its
    // boxes), at least one, about one work item per SM in total.  Jobs of (nearly) equal cost get the same slice count, which
    // keeps the items of one slice -- the jobs that read the same pixels -- exactly in step.
    std::vector<double> wgt(n_jobs);
    double total_wd = 0, wmin = 1e300, wmax = 0;
    for (int j = 0; j < n_jobs; ++j) {
      const int co_here = std::min(128, Cout - p.job_coblk[j] * 128);
      const double bytes = (double)ceil_div(co_here, p.cbo) * (stacked ? 2 : 1) * p.dy_chunk_bytes + (double)p.job_ng[j] * p.n_ci_chunks * p.x_box_bytes;
      wgt[j] = std::max(job_cost[j], bytes / 48.0);
      total_wd += wgt[j];
      wmin = std::min(wmin, wgt[j]);
      wmax = std::max(wmax, wgt[j]);
    }
    if (wmax <= 1.15 * wmin) { for (int j = 0; j < n_jobs; ++j) wgt[j] = 1.0; total_wd = n_jobs; }
    int items = 0;
    std::vector<std::pair<double, std::pair<int, int>>> order;
    for (int j = 0; j < n_jobs; ++j) {
      int nsl = std::max(1, std::min((int)total, (int)(sms * wgt[j] / total_wd)));
      const int tps = ceil_div((int)total, nsl);
      nsl = ceil_div((int)total, tps);
      SOS_CHECK_ARG(items + nsl <= kMaxItems, "sos_conv2d_wgrad: too many work items");
      p.job_nsl[j] = (int16_t)nsl;
      p.job_tps[j] = tps;
      for (int sl = 0; sl < nsl; ++sl) order.push_back({(double)sl / nsl, {j, sl}});
      items += nsl;
    }
    std::stable_sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    for (int i = 0; i < items; ++i) { p.item_job[i] = (int16_t)order[i].second.first; p.item_slice[i] = (int16_t)order[i].second.second; }
    p.n_items = items;
  }

  out.smem = 2048 + p.n_stages * p.stage_bytes + tail;
  out.grid = std::min(p.n_items, sms);
  out.plan_out[0] = pl.fast_is_w;
  out.plan_out[1] = pl.share;
  out.plan_out[2] = pl.g;
  out.plan_out[3] = SB;
  out.plan_out[4] = n_jobs;
  out.plan_out[5] = p.n_stages;
  out.plan_out[6] = p.stage_bytes;
  out.plan_out[7] = p.job_nsl[0] + (stacked ? 1000 : 0);
  return SOS_OK;
}

}  // namespace

extern "C" int sos_conv2d_wgrad(const sos_wgrad_args* ap, cudaStream_t stream) {
  SOS_CHECK_ARG(ap != nullptr, "sos_conv2d_wgrad: null args");
  const sos_wgrad_args& a = *ap;
  SOS_CHECK_ARG(a.x && a.dy && a.dw && a.tap_dh && a.tap_dw, "sos_conv2d_wgrad: null pointer");
  SOS_CHECK_ARG(a.N > 0 && a.H > 0 && a.W > 0 && a.OH > 0 && a.OW > 0 && a.ntaps > 0 && a.ntaps <= 49, "sos_conv2d_wgrad: bad shape");
  SOS_CHECK_ARG(a.dtype == SOS_DTYPE_TF32 || a.dtype == SOS_DTYPE_F16, "sos_conv2d_wgrad: unknown operand type");
  const int esz = a.dtype == SOS_DTYPE_F16 ? 2 : 4;
  SOS_CHECK_ARG(a.Cin >= 8 && a.Cin % 8 == 0 && a.Cin <= 256, "sos_conv2d_wgrad: Cin must be a multiple of 8 in [8, 256] (got %lld)",
                (long long)a.Cin);
  SOS_CHECK_ARG(a.Cout >= 8 && a.Cout % 8 == 0 && a.Cout <= 1024, "sos_conv2d_wgrad: Cout must be a multiple of 8 in [8, 1024] (got %lld)",
                (long long)a.Cout);
  SOS_CHECK_ARG(a.Cdy % (16 / esz) == 0 && a.dy_coff % (16 / esz) == 0 && a.dy_coff + a.Cout <= a.Cdy, "sos_conv2d_wgrad: bad dy channel slice");
  SOS_CHECK_ARG(a.stride == 1 || a.stride == 2, "sos_conv2d_wgrad: stride must be 1 or 2");
  SOS_CHECK_ARG(((uintptr_t)a.x % 16) == 0 && ((uintptr_t)a.dy % 16) == 0, "sos_conv2d_wgrad: pointers must be 16-byte aligned");

  std::vector<int32_t> key;
  key.reserve(16 + 2 * (size_t)a.ntaps);
  const int64_t fields[] = {a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.Cdy, a.dy_coff, a.ntaps, a.stride, a.force_plan, a.dtype};
  for (int64_t f : fields) key.push_back((int32_t)f);
  for (int t = 0; t < a.ntaps; ++t) { key.push_back(a.tap_dh[t]); key.push_back(a.tap_dw[t]); }

  std::lock_guard<std::mutex> lock(g_wg_mutex);
  WgPlan* plan;
  auto it = g_wg_plans.find(key);
  if (it != g_wg_plans.end()) {
    plan = it->second;
  } else {
    plan = new WgPlan();
    if (int e = plan_conv2d_wgrad(a, *plan)) { delete plan; return e; }
    g_wg_plans.emplace(std::move(key), plan);
  }
  WgParams& p = plan->p;
  p.out_scale = a.out_scale;
  static const int wg_dbg = getenv("SOS_WGRAD_DBG") ? atoi(getenv("SOS_WGRAD_DBG")) : 0;
  p.dbg = wg_dbg;
  p.dw = a.dw;
  const void* baseDY = reinterpret_cast<const uint8_t*>(a.dy) + a.dy_coff * esz;
  if (plan->baseX != a.x) { if (int e = encode_spec(&p.mapX, plan->specX, a.x)) return e; plan->baseX = a.x; }
  if (plan->baseDY != baseDY) { if (int e = encode_spec(&p.mapDY, plan->specDY, baseDY)) return e; plan->baseDY = baseDY; }

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess ||
        cudaFuncSetAttribute(wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess) {
      sos_set_error("sos_conv2d_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SOS_ERR_CUDA;
    }
    attr_set = true;
  }
  if (plan->esz == 2) wgrad_f16_kernel<<<plan->grid, kThreadsWg, plan->smem, stream>>>(p);
  else wgrad_tf32_kernel<<<plan->grid, kThreadsWg, plan->smem, stream>>>(p);
  SOS_CHECK_LAUNCH("sos_conv2d_wgrad");
  if (a.plan_out) memcpy(a.plan_out, plan->plan_out, sizeof(plan->plan_out));
  return SOS_OK;
}
