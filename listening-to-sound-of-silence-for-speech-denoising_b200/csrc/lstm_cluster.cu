// Bidirectional LSTM recurrence with the blocks of one direction in ONE thread-block cluster (sm_100a: up to 16 CTAs with the
// non-portable size): the step-to-step exchange goes through distributed shared memory and the hardware cluster barrier instead
// of global memory and an atomic-counter barrier (lstm.cu).  Same contract as sos_lstm_forward / sos_lstm_backward; used when the
// batch fits one 32-clip tile and ceil(H / 16) <= 16 (both networks of the hot path: H = 100 -> 7 CTAs, H = 200 -> 13 CTAs).
//
// forward   a CTA owns 16 hidden units (64 gate rows) of one direction and keeps its W_hh rows in shared memory.  Per step:
//           64 x 32 x H product against h_prev (the full vector, shared memory, double buffered), cell update, then the CTA
//           writes its 16 x 32 slice of h into the "next" buffer of EVERY CTA of the cluster (16-byte DSMEM stores) and the
//           cluster barrier publishes it.
// backward  dh_rec[b][k] = sum_r dgx[t_next][b][r] W_hh[r][k].  A CTA owns the SAME 64 rows r (its 16 units' four gates), whose
//           dgx it produced itself in the previous step (kept in shared memory): it computes its partial of dh_rec for ALL H
//           units and scatters the 16-unit slices to their owners (a reduce-scatter through DSMEM); the owner sums the slices.
#include "common.cuh"
#include "sos_b200.h"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kCU = 16;            // hidden units per CTA
constexpr int kCRows = 4 * kCU;    // gate rows per CTA
constexpr int kCB = 32;            // batch tile (whole batch)
constexpr int kHPitch = kCB + 4;   // h_s row pitch (floats): 16-byte aligned rows

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// The per-step cluster barrier as an explicit release / acquire pair.  cooperative_groups' cluster.sync() puts a GPU-scope
// MEMBAR + ERRBAR in front of the barrier (18 % of the forward kernel's stall samples, profiles/r01b_lstm_fwd_cluster_ncu.txt):
// the exchange only needs the distributed-shared-memory stores ordered at cluster scope, and the split lets the step's global
// stores (cell, h, gates: nobody in the cluster reads them) be issued between arrive and wait.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Two IEEE fp32 FMAs in one instruction (sm_100 FFMA2): (d0, d1) += w * (b0, b1).  The recurrent products are issue bound
// (one SM does 410 k FMAs per time step), so halving the FMA instruction count is what shortens a step.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float w, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// gx (T,B,2,4H); w_hh (2,4H,H); out (T,B,2H); gates (T,B,2,4H) activated; cell (T,B,2,H)
// The per-step product (64 gate rows x 32 clips x H) is bound by shared-memory wavefronts, not by FMA issue: a thread therefore
// owns a 4-row x 8-clip register tile (3 LDS.128 per 32 FMAs; W_hh k-major so that 4 rows are one 16-byte load) of one eighth of
// the k range; the eight partial sums are added in the cell update, (unit, clip) per thread, whose cell state and input-projection
// terms (fetched one step ahead) live in registers.
constexpr int kFwdThreads = 512, kKSplit = 8;
__global__ void __launch_bounds__(kFwdThreads, 1) lstm_fwd_cluster_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh, int T, int B,
                                                                          int H, float* __restrict__ out, float* __restrict__ gates,
                                                                          float* __restrict__ cell) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int d = blockIdx.y, nblk = gridDim.x, j0 = blockIdx.x * kCU, tid = threadIdx.x;
  float* w_s = sm;                                   // [H][64]            k-major; row = q * 16 + unit
  float* h_s = w_s + (size_t)H * kCRows;             // [2][H][kHPitch]    h_prev, k-major, clip fastest
  float* g_s = h_s + 2 * H * kHPitch;                // [8][64][kCB + 1]   partial gate pre-activations of the k slices
  float* o_s = g_s + kKSplit * kCRows * (kCB + 1);   // [16][kCB]          this CTA's new h slice (16-byte rows)
  for (int e = tid; e < kCRows * H; e += kFwdThreads) {
    const int r = e / H, k = e - r * H;              // coalesced global read, transposing store (once per launch)
    const int q = r / kCU, j = j0 + (r - q * kCU);
    w_s[k * kCRows + r] = j < H ? w_hh[((size_t)d * 4 * H + (size_t)q * H + j) * H + k] : 0.f;
  }
  for (int e = tid; e < 2 * H * kHPitch; e += kFwdThreads) h_s[e] = 0.f;
  cluster.sync();                                    // every CTA's buffers exist and are zeroed before anyone writes into them
  const int rg = tid & 15, bg = (tid >> 4) & 3, ks = tid >> 6;            // 4 gate rows, 8 clips, k slice
  const int kc = (H + kKSplit - 1) / kKSplit, ka = ks * kc, kb = min(H, ka + kc);
  const int ju = tid & (kCU - 1), bu = tid >> 4, ju_g = j0 + ju;          // cell update: (unit, clip)
  const bool cell_live = ju_g < H && bu < B;
  float c_reg = 0.f;
  auto load_gx = [&](int step, float (&o)[4]) {
    const int t = d == 0 ? step : T - 1 - step;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) o[qq] = cell_live ? __ldg(gx + (((size_t)t * B + bu) * 2 + d) * 4 * H + (size_t)qq * H + ju_g) : 0.f;
  };
  float gxv[4], gxn[4] = {0.f, 0.f, 0.f, 0.f};
  load_gx(0, gxv);
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const float* hc = h_s + (size_t)(step & 1) * H * kHPitch;
    float* hn = h_s + (size_t)((step + 1) & 1) * H * kHPitch;
    if (step + 1 < T) load_gx(step + 1, gxn);        // in flight during this step
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
    if (step > 0) {
      const float* wp = w_s + rg * 4;
      const float* hb = hc + bg * 8;
#pragma unroll 2
      for (int k = ka; k < kb; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(wp + k * kCRows);
        const float4 x = *reinterpret_cast<const float4*>(hb + k * kHPitch), y = *reinterpret_cast<const float4*>(hb + k * kHPitch + 4);
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          ffma2(acc[a][0], acc[a][1], wv[a], x.x, x.y);
          ffma2(acc[a][2], acc[a][3], wv[a], x.z, x.w);
          ffma2(acc[a][4], acc[a][5], wv[a], y.x, y.y);
          ffma2(acc[a][6], acc[a][7], wv[a], y.z, y.w);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int i = 0; i < 8; ++i) g_s[(ks * kCRows + rg * 4 + a) * (kCB + 1) + bg * 8 + i] = acc[a][i];
    __syncthreads();
    float h = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f;
    if (cell_live) {
      float pre[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        float v = gxv[qq];
#pragma unroll
        for (int s2 = 0; s2 < kKSplit; ++s2) v += g_s[(s2 * kCRows + qq * kCU + ju) * (kCB + 1) + bu];
        pre[qq] = v;
      }
      ig = sigm(pre[0]); fg = sigm(pre[1]); gg = tanhf(pre[2]); og = sigm(pre[3]);
      c_reg = fg * c_reg + ig * gg;
      h = og * tanhf(c_reg);
    }
    o_s[ju * kCB + bu] = h;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) gxv[qq] = gxn[qq];
    __syncthreads();
    if (step + 1 < T) {
      // broadcast the slice into the "next" h buffer of every CTA of the direction: 16 rows x 8 float4 per destination
      for (int e = tid; e < nblk * kCU * (kCB / 4); e += kFwdThreads) {
        const int peer = e / (kCU * (kCB / 4)), rem = e - peer * (kCU * (kCB / 4));
        const int row = rem >> 3, v = rem & 7;
        if (j0 + row < H) {
          float* dst = cluster.map_shared_rank(hn, peer) + (size_t)(j0 + row) * kHPitch + v * 4;
          *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(o_s + row * kCB + v * 4);
        }
      }
      cluster_arrive();                              // release: this thread's slice stores are ordered before the barrier
    }
    if (cell_live) {                                 // the step's outputs: global stores issued under the barrier's latency
      cell[(((size_t)t * B + bu) * 2 + d) * H + ju_g] = c_reg;
      out[((size_t)t * B + bu) * 2 * H + (size_t)d * H + ju_g] = h;
      float* gp = gates + (((size_t)t * B + bu) * 2 + d) * 4 * H + ju_g;
      gp[0] = ig;
      gp[H] = fg;
      gp[2 * H] = gg;
      gp[3 * H] = og;
    }
    if (step + 1 < T) cluster_wait();                // acquire: every CTA's slice of the next h is in this CTA's buffer
  }
}

// dout (T,B,2H); gates, cell as written by the forward; dgx (T,B,2,4H) out; dc_ws (B,2,H) scratch.
constexpr int kBwdThreads = 256;     // 256: every (4 k, 8 clips) item sums all 64 gate rows; 512: two row halves combined through smem
constexpr int kRowSplit = kBwdThreads / 256, kPairs = 512 / kBwdThreads;
__global__ void __launch_bounds__(kBwdThreads, 1) lstm_bwd_cluster_kernel(const float* __restrict__ dout, const float* __restrict__ w_hh,
                                                                        const float* __restrict__ gates, const float* __restrict__ cell, int T,
                                                                        int B, int H, float* __restrict__ dgx) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int d = blockIdx.y, nblk = gridDim.x, rank = blockIdx.x, j0 = rank * kCU, tid = threadIdx.x;
  const int Hq = (H + 3) & ~3;                       // k extent rounded to float4
  float* w_s = sm;                                   // [64][Hq]           W_hh rows of this CTA's units, k fastest
  float* dg_s = w_s + kCRows * Hq;                   // [64][kCB]          dgx of the previous step for those rows, clip fastest
  float* part_s = dg_s + kCRows * kCB;               // [2][nblk][16][kCB] partial dh_rec slices received from every CTA
  float* dc_s = part_s + 2 * nblk * kCU * kCB;       // [16][kCB]          dc carried across steps
  float* red_s = dc_s + kCU * kCB;                   // [Hq/4 * 4 items][32]   partial sums of the second row half
  for (int e = tid; e < kCRows * Hq; e += kBwdThreads) {
    const int r = e / Hq, k = e - r * Hq;
    const int q = r / kCU, j = j0 + (r - q * kCU);
    w_s[e] = (j < H && k < H) ? w_hh[((size_t)d * 4 * H + (size_t)q * H + j) * H + k] : 0.f;
  }
  for (int e = tid; e < kCRows * kCB; e += kBwdThreads) dg_s[e] = 0.f;
  for (int e = tid; e < kCU * kCB; e += kBwdThreads) dc_s[e] = 0.f;
  cluster.sync();
  const int R = 4 * H;
  const int ju = tid & (kCU - 1), j = j0 + ju;
  const int hf = tid >> 8, item = tid & 255;         // product: row half, (4 k, 8 clips) item
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? T - 1 - step : step;      // backward walks each direction's time axis in reverse
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = d == 0 ? (t > 0) : (t < T - 1);
    float* pr = part_s + (size_t)(step & 1) * nblk * kCU * kCB;
    // elementwise operands of this step (independent of the other CTAs): in flight during the product
    float pf[kPairs][7];
#pragma unroll
    for (int half = 0; half < kPairs; ++half) {
      const int b = (tid >> 4) + (kBwdThreads / 16) * half;
#pragma unroll
      for (int i = 0; i < 7; ++i) pf[half][i] = 0.f;
      if (j < H && b < B) {
        const size_t gi = (((size_t)t * B + b) * 2 + d) * R + j;
        pf[half][0] = __ldg(gates + gi); pf[half][1] = __ldg(gates + gi + H); pf[half][2] = __ldg(gates + gi + 2 * H);
        pf[half][3] = __ldg(gates + gi + 3 * H);
        pf[half][4] = __ldg(cell + (((size_t)t * B + b) * 2 + d) * H + j);
        pf[half][5] = has_prev ? __ldg(cell + (((size_t)tp * B + b) * 2 + d) * H + j) : 0.f;
        pf[half][6] = __ldg(dout + ((size_t)t * B + b) * 2 * H + (size_t)d * H + j);
      }
    }
    if (step > 0) {
      // partial[b][k] = sum over this CTA's 64 rows of dg_s[r][b] * w_s[r][k], for all k: thread -> (4 consecutive k, 8 clips),
      // written straight into the owner CTA's receive slot [rank][k - owner*16][b]
      const int kq = Hq >> 2;                        // float4 groups along k (kq * 4 <= 256 items: H <= 256)
      const bool live = item < kq * 4;
      const int k4 = item % kq, bg = item / kq;      // bg: 8 clips
      float acc[4][8];
      if (live) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
#pragma unroll 2
        for (int rr = hf * (kCRows / kRowSplit); rr < (hf + 1) * (kCRows / kRowSplit); ++rr) {
          const float4 w = *reinterpret_cast<const float4*>(w_s + rr * Hq + k4 * 4);
          const float4 g0 = *reinterpret_cast<const float4*>(dg_s + rr * kCB + bg * 8), g1 = *reinterpret_cast<const float4*>(dg_s + rr * kCB + bg * 8 + 4);
          const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 8; i += 2) ffma2(acc[a][i], acc[a][i + 1], wv[a], gv[i], gv[i + 1]);
        }
        if (hf == 1) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 8; ++i) red_s[(a * 8 + i) * 256 + item] = acc[a][i];      // item fastest: conflict-free
        }
      }
      if (kRowSplit > 1) __syncthreads();
      if (live && hf == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (kRowSplit > 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[a][i] += red_s[(a * 8 + i) * 256 + item];
          }
          const int k = k4 * 4 + a;
          if (k < H) {
            const int owner = k / kCU, ku = k - owner * kCU;
            float* dst = cluster.map_shared_rank(pr, owner) + ((size_t)rank * kCU + ku) * kCB + bg * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
          }
        }
      }
      cluster_arrive();
      cluster_wait();                                // every CTA's partial slices have arrived
    }
#pragma unroll
    for (int half = 0; half < kPairs; ++half) {
      const int b = (tid >> 4) + (kBwdThreads / 16) * half;
      float dgi = 0.f, dgf = 0.f, dgg = 0.f, dgo = 0.f;
      if (j < H && b < B) {
        float dh_rec = 0.f;
        if (step > 0)
          for (int p = 0; p < nblk; ++p) dh_rec += pr[((size_t)p * kCU + ju) * kCB + b];
        const float ig = pf[half][0], fg = pf[half][1], gg = pf[half][2], og = pf[half][3], c = pf[half][4], cp = pf[half][5];
        const float tc = tanhf(c);
        const float dh = pf[half][6] + dh_rec;
        const float dc = dh * og * (1.f - tc * tc) + dc_s[ju * kCB + b];
        dgi = dc * gg * ig * (1.f - ig);
        dgf = dc * cp * fg * (1.f - fg);
        dgg = dc * ig * (1.f - gg * gg);
        dgo = dh * tc * og * (1.f - og);
        const size_t gi = (((size_t)t * B + b) * 2 + d) * R + j;
        dgx[gi] = dgi;
        dgx[gi + H] = dgf;
        dgx[gi + 2 * H] = dgg;
        dgx[gi + 3 * H] = dgo;
        dc_s[ju * kCB + b] = dc * fg;
      }
      dg_s[(0 * kCU + ju) * kCB + b] = dgi;
      dg_s[(1 * kCU + ju) * kCB + b] = dgf;
      dg_s[(2 * kCU + ju) * kCB + b] = dgg;
      dg_s[(3 * kCU + ju) * kCB + b] = dgo;
    }
    __syncthreads();                                 // dg_s complete before the next step's product
  }
}

bool cluster_enabled() {
  static const bool on = !(getenv("SOS_LSTM_CLUSTER") && atoi(getenv("SOS_LSTM_CLUSTER")) == 0);
  return on;
}

bool g_cluster_broken = false;     // a failed cluster launch (e.g. no GPC can host the cluster) switches back to lstm.cu for good

template <typename... Args>
cudaError_t launch_cluster(void (*kernel)(Args...), int nblk, int threads, size_t smem, cudaStream_t stream, Args... args) {
  static size_t smem_set[64] = {0};               // per kernel instantiation AND device: function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& cur = smem_set[dev & 63];
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    cur = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)nblk, 2, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nblk;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace

// Returns 1 when the cluster kernel was launched, 0 when the shape is outside its range (the caller then runs the global-barrier
// kernel), negative on a launch error.
int sos_lstm_forward_cluster(const float* gx, const float* w_hh, int64_t T, int64_t B, int64_t H, float* out, float* gates_ws, float* cell_ws,
                             cudaStream_t stream) {
  const int nblk = (int)((H + kCU - 1) / kCU);
  if (!cluster_enabled() || g_cluster_broken || B > kCB || nblk > 16) return 0;
  const size_t smem = ((size_t)kCRows * H + 2 * (size_t)H * kHPitch + (size_t)kKSplit * kCRows * (kCB + 1) + (size_t)kCU * kCB) * sizeof(float);
  if (smem > 220 * 1024) return 0;
  cudaError_t e = launch_cluster(lstm_fwd_cluster_kernel, nblk, kFwdThreads, smem, stream, gx, w_hh, (int)T, (int)B, (int)H, out, gates_ws, cell_ws);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_cluster_broken = true;
    return 0;
  }
  return 1;
}

int sos_lstm_backward_cluster(const float* dout, const float* w_hh, const float* gates_ws, const float* cell_ws, int64_t T, int64_t B, int64_t H,
                              float* dgx, cudaStream_t stream) {
  const int nblk = (int)((H + kCU - 1) / kCU);
  if (!cluster_enabled() || g_cluster_broken || B > kCB || nblk > 16) return 0;
  const size_t Hq = (size_t)((H + 3) & ~3);
  const size_t smem = ((size_t)kCRows * Hq + (size_t)kCRows * kCB + 2 * (size_t)nblk * kCU * kCB + (size_t)kCU * kCB + 32 * 256) * sizeof(float);
  if (smem > 220 * 1024) return 0;
  cudaError_t e = launch_cluster(lstm_bwd_cluster_kernel, nblk, kBwdThreads, smem, stream, dout, w_hh, gates_ws, cell_ws, (int)T, (int)B, (int)H, dgx);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_cluster_broken = true;
    return 0;
  }
  return 1;
}
