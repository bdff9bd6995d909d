// Bidirectional single-layer LSTM recurrence (nn.LSTM of M1/networks.py:95,147-148 and M2/networks.py:64,88-89),
// PyTorch gate order i, f, g, o.  The input projection x W_ih^T + b_ih + b_hh is a plain GEMM done by the caller;
// these kernels run the serial part as ONE persistent launch per pass: both directions in the same grid (direction 0
// walks t = 0..T-1, direction 1 walks t = T-1..0), a block owns kUnits hidden units of one direction for the whole
// batch and keeps its 4*kUnits rows of W_hh in shared memory for all T steps; between steps the blocks of a direction
// meet at a global-memory barrier (release/acquire on a counter), so a step costs a barrier plus one small matrix
// product instead of a kernel launch.  The grid (2 * ceil(H / kUnits) blocks, <= 128 for H <= 512) must be co-resident:
// it is launched with one block per SM's worth of shared memory and the host checks it against the SM count.
#include "common.cuh"
#include <atomic>
#include <mutex>
#include "sos_b200.h"
#include <algorithm>

namespace {

constexpr int kUnits = 8;        // hidden units per block
constexpr int kLstmThreads = 256;
constexpr int kBt = 32;          // batch tile held in shared memory

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Barrier among the blocks of one direction (they are co-resident).  `target` = arrivals expected so far.
__device__ __forceinline__ void dir_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

constexpr int kHsPitch = kBt + 4;    // h_s row pitch (floats): keeps float4 alignment, 4-way conflicts on the transposing store only
constexpr int kGsPitch = kBt + 1;    // g_s row pitch of the backward kernel: conflict-free transposing store
constexpr int kBwdChunk = 800;       // gate rows of dgx staged per pass of the backward kernel

// gx (T,B,2,4H); w_hh (2,4H,H); out (T,B,2H); gates (T,B,2,4H) activated; cell (T,B,2,H)
__global__ void __launch_bounds__(kLstmThreads) lstm_fwd_step_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh,
                                                                     int T, int B, int H, float* __restrict__ out,
                                                                     float* __restrict__ gates, float* __restrict__ cell,
                                                                     unsigned int* __restrict__ sync) {
  extern __shared__ __align__(16) float sm[];
  const int d = blockIdx.y;
  const int j0 = blockIdx.x * kUnits;
  const int Hp = H + 1;
  float* h_s = sm;                                   // [H][kHsPitch]  h_prev transposed: k-major, batch fastest
  float* w_s = h_s + H * kHsPitch;                   // [32][H+1]      row r = q*kUnits + jj
  float* g_s = w_s + 32 * Hp;                        // [32][kBt+1]
  const int tid = threadIdx.x;
  for (int e = tid; e < 32 * H; e += kLstmThreads) {
    const int r = e / H, k = e - r * H;
    const int q = r / kUnits, jj = r - q * kUnits;
    const int j = j0 + jj;
    w_s[r * Hp + k] = j < H ? w_hh[((size_t)d * 4 * H + (size_t)q * H + j) * H + k] : 0.f;
  }
  const int r = tid & 31;                            // gate row handled by this thread
  const int bg = tid >> 5;                           // batch group: 4 batches each
  const int q = r / kUnits, jj = r - q * kUnits;
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const int tp = d == 0 ? t - 1 : t + 1;           // previous time step of this direction
    const bool first = step == 0;
    for (int b0 = 0; b0 < B; b0 += kBt) {
      __syncthreads();
      if (!first) {
        // h_prev was written by the other blocks of this direction during the previous step: read through L2
        if ((H & 3) == 0) {
          const int H4 = H >> 2;
          for (int e = tid; e < kBt * H4; e += kLstmThreads) {
            const int bb = e / H4, k = (e - bb * H4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + bb < B) v = __ldcg(reinterpret_cast<const float4*>(out + ((size_t)tp * B + b0 + bb) * 2 * H + (size_t)d * H + k));
            h_s[k * kHsPitch + bb] = v.x;
            h_s[(k + 1) * kHsPitch + bb] = v.y;
            h_s[(k + 2) * kHsPitch + bb] = v.z;
            h_s[(k + 3) * kHsPitch + bb] = v.w;
          }
        } else {
          for (int e = tid; e < kBt * H; e += kLstmThreads) {
            const int bb = e / H, k = e - bb * H;
            h_s[k * kHsPitch + bb] = (b0 + bb < B) ? __ldcg(out + ((size_t)tp * B + b0 + bb) * 2 * H + (size_t)d * H + k) : 0.f;
          }
        }
      }
      // the input-projection term of this thread's four gate values, in flight during the product
      float gxv[4];
      const int j = j0 + jj;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = b0 + bg * 4 + i;
        gxv[i] = (b < B && j < H) ? __ldg(gx + (((size_t)t * B + b) * 2 + d) * 4 * H + (size_t)q * H + j) : 0.f;
      }
      __syncthreads();
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (!first) {
        const float* wr = w_s + r * Hp;
        const float* hb = h_s + bg * 4;
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
          const float w = wr[k];
          const float4 h4 = *reinterpret_cast<const float4*>(hb + k * kHsPitch);
          acc[0] = fmaf(w, h4.x, acc[0]);
          acc[1] = fmaf(w, h4.y, acc[1]);
          acc[2] = fmaf(w, h4.z, acc[2]);
          acc[3] = fmaf(w, h4.w, acc[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) g_s[r * (kBt + 1) + bg * 4 + i] = acc[i] + gxv[i];
      __syncthreads();
      // cell update: thread -> (unit jj2, batch bb)
      const int jj2 = tid & (kUnits - 1), bb = tid >> 3;
      const int j2 = j0 + jj2, b = b0 + bb;
      if (j2 < H && b < B) {
        const float ig = sigmoidf_(g_s[(0 * kUnits + jj2) * (kBt + 1) + bb]);
        const float fg = sigmoidf_(g_s[(1 * kUnits + jj2) * (kBt + 1) + bb]);
        const float gg = tanhf(g_s[(2 * kUnits + jj2) * (kBt + 1) + bb]);
        const float og = sigmoidf_(g_s[(3 * kUnits + jj2) * (kBt + 1) + bb]);
        const float cp = first ? 0.f : cell[(((size_t)tp * B + b) * 2 + d) * H + j2];      // this thread's own value
        const float c = fg * cp + ig * gg;
        const float h = og * tanhf(c);
        cell[(((size_t)t * B + b) * 2 + d) * H + j2] = c;
        out[((size_t)t * B + b) * 2 * H + (size_t)d * H + j2] = h;
        float* gp = gates + (((size_t)t * B + b) * 2 + d) * 4 * H + j2;
        gp[0] = ig;
        gp[H] = fg;
        gp[2 * H] = gg;
        gp[3 * H] = og;
      }
    }
    if (step + 1 < T) dir_barrier(sync + d, (unsigned)(step + 1) * gridDim.x);
  }
}

// Backward through time.  dgx (T,B,2,4H) receives pre-activation gate gradients; dh_ws is unused storage kept for
// ABI stability; dc_ws (B,2,H) carries dc across steps (first written at step 0; each element is private to one thread).
// The recurrent term dh_rec[b][k] = sum_r dgx[t_next][b][d][r] * w_hh[d][r][k] reads the dgx rows every block of the
// direction wrote in the previous step, hence the barrier between steps.  Per batch tile it is a 32 x 8 x 4H product:
// the 8 warps split the gate rows, lane = clip, 8 accumulators (the block's hidden units) per thread, then the warps'
// partial tiles are summed through shared memory.
__global__ void __launch_bounds__(kLstmThreads) lstm_bwd_step_kernel(const float* __restrict__ dout, const float* __restrict__ w_hh,
                                                                     const float* __restrict__ gates, const float* __restrict__ cell,
                                                                     int T, int B, int H, float* __restrict__ dgx,
                                                                     float* __restrict__ dc_ws, unsigned int* __restrict__ sync) {
  extern __shared__ __align__(16) float sm[];
  const int d = blockIdx.y;
  const int k0 = blockIdx.x * kUnits;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = 4 * H;
  const int RC = R < kBwdChunk ? R : kBwdChunk;      // gate rows staged per pass
  float* w_s = sm;                                   // [R][kUnits]      w_hh[d][r][k0 + u]
  float* g_s = w_s + (size_t)R * kUnits;             // [RC][kGsPitch]   dgx[tn] rows of this batch tile, clip fastest
  float* red_s = g_s + (size_t)RC * kGsPitch;        // [8 warps][kUnits][kBt]
  const int kk = tid & (kUnits - 1), bb = tid >> 3;  // elementwise part: thread -> (unit kk, batch bb)
  for (int e = tid; e < kUnits * R; e += kLstmThreads) {
    const int r = e / kUnits, u = e - r * kUnits;    // coalescing is poor (stride H) but this runs once per launch
    w_s[e] = (k0 + u < H) ? w_hh[((size_t)d * R + r) * H + k0 + u] : 0.f;
  }
  // The elementwise operands of a step (saved gates, cell states, dout) do not depend on the other blocks: for the first
  // batch tile they are fetched BEFORE the inter-step barrier so that their DRAM latency overlaps the wait.
  float pf[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto fetch = [&](int step, int b0, float (&o)[7]) {
    const int t = d == 0 ? T - 1 - step : step;
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = d == 0 ? (t > 0) : (t < T - 1);
    const int j = k0 + kk, b = b0 + bb;
    if (j < H && b < B) {
      const size_t gi = (((size_t)t * B + b) * 2 + d) * R + j;
      o[0] = __ldg(gates + gi); o[1] = __ldg(gates + gi + H); o[2] = __ldg(gates + gi + 2 * H); o[3] = __ldg(gates + gi + 3 * H);
      o[4] = __ldg(cell + (((size_t)t * B + b) * 2 + d) * H + j);
      o[5] = has_prev ? __ldg(cell + (((size_t)tp * B + b) * 2 + d) * H + j) : 0.f;
      o[6] = __ldg(dout + ((size_t)t * B + b) * 2 * H + (size_t)d * H + j);
    }
  };
  fetch(0, 0, pf);
  for (int step = 0; step < T; ++step) {
    // backward walks each direction's time axis in reverse
    const int t = d == 0 ? T - 1 - step : step;
    const int tn = d == 0 ? t + 1 : t - 1;           // the step processed just before (later in the recurrence)
    const bool first = step == 0;
    for (int b0 = 0; b0 < B; b0 += kBt) {
      float cur[7];
      if (b0 == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) cur[i] = pf[i];
      } else {
        fetch(step, b0, cur);
      }
      float dh_rec = 0.f;
      if (!first) {
        float acc[kUnits];
#pragma unroll
        for (int u = 0; u < kUnits; ++u) acc[u] = 0.f;
        for (int r0 = 0; r0 < R; r0 += RC) {
          const int rc = min(RC, R - r0);
          __syncthreads();
          if (((rc | r0) & 3) == 0) {
            const int rc4 = rc >> 2;
            for (int e = tid; e < kBt * rc4; e += kLstmThreads) {
              const int b2 = e / rc4, rr = (e - b2 * rc4) * 4;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (b0 + b2 < B) v = __ldcg(reinterpret_cast<const float4*>(dgx + (((size_t)tn * B + b0 + b2) * 2 + d) * R + r0 + rr));
              g_s[rr * kGsPitch + b2] = v.x;
              g_s[(rr + 1) * kGsPitch + b2] = v.y;
              g_s[(rr + 2) * kGsPitch + b2] = v.z;
              g_s[(rr + 3) * kGsPitch + b2] = v.w;
            }
          } else {
            for (int e = tid; e < kBt * rc; e += kLstmThreads) {
              const int b2 = e / rc, rr = e - b2 * rc;
              g_s[rr * kGsPitch + b2] = (b0 + b2 < B) ? __ldcg(dgx + (((size_t)tn * B + b0 + b2) * 2 + d) * R + r0 + rr) : 0.f;
            }
          }
          __syncthreads();
          const int per = (rc + 7) >> 3;
          const int ra = warp * per, rb = min(rc, ra + per);
#pragma unroll 4
          for (int rr = ra; rr < rb; ++rr) {
            const float g = g_s[rr * kGsPitch + lane];
            const float4 w0 = *reinterpret_cast<const float4*>(w_s + (size_t)(r0 + rr) * kUnits);
            const float4 w1 = *reinterpret_cast<const float4*>(w_s + (size_t)(r0 + rr) * kUnits + 4);
            acc[0] = fmaf(g, w0.x, acc[0]); acc[1] = fmaf(g, w0.y, acc[1]); acc[2] = fmaf(g, w0.z, acc[2]); acc[3] = fmaf(g, w0.w, acc[3]);
            acc[4] = fmaf(g, w1.x, acc[4]); acc[5] = fmaf(g, w1.y, acc[5]); acc[6] = fmaf(g, w1.z, acc[6]); acc[7] = fmaf(g, w1.w, acc[7]);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnits; ++u) red_s[(warp * kUnits + u) * kBt + lane] = acc[u];
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; ++w) dh_rec += red_s[(w * kUnits + kk) * kBt + bb];
      }
      const int j = k0 + kk, b = b0 + bb;
      if (j < H && b < B) {
        const size_t gi = (((size_t)t * B + b) * 2 + d) * R + j;
        const float ig = cur[0], fg = cur[1], gg = cur[2], og = cur[3], c = cur[4], cp = cur[5];
        const float tc = tanhf(c);
        const float dh = cur[6] + dh_rec;
        const size_t ci = ((size_t)b * 2 + d) * H + j;
        const float dc = dh * og * (1.f - tc * tc) + (first ? 0.f : dc_ws[ci]);
        dgx[gi] = dc * gg * ig * (1.f - ig);
        dgx[gi + H] = dc * cp * fg * (1.f - fg);
        dgx[gi + 2 * H] = dc * ig * (1.f - gg * gg);
        dgx[gi + 3 * H] = dh * tc * og * (1.f - og);
        dc_ws[ci] = dc * fg;
      }
    }
    if (step + 1 < T) {
      fetch(step + 1, 0, pf);
      dir_barrier(sync + d, (unsigned)(step + 1) * gridDim.x);
    }
  }
}

}  // namespace

// Barrier counters: a ring of slots so that launches on different streams do not share one (2 counters per launch).
static unsigned int* g_sync = nullptr;
static std::atomic<unsigned> g_sync_next{0};
static std::mutex g_sync_mutex;
constexpr int kSyncSlots = 256;
static unsigned int* next_sync_slot(cudaStream_t stream) {
  {
    std::lock_guard<std::mutex> lock(g_sync_mutex);
    if (!g_sync && cudaMalloc(&g_sync, kSyncSlots * 2 * sizeof(unsigned int)) != cudaSuccess) {
      g_sync = nullptr;
      return nullptr;
    }
  }
  unsigned int* p = g_sync + 2 * (g_sync_next.fetch_add(1) % kSyncSlots);
  if (cudaMemsetAsync(p, 0, 2 * sizeof(unsigned int), stream) != cudaSuccess) return nullptr;
  return p;
}

int sos_lstm_forward_cluster(const float* gx, const float* w_hh, int64_t T, int64_t B, int64_t H, float* out, float* gates_ws, float* cell_ws,
                             cudaStream_t stream);
int sos_lstm_backward_cluster(const float* dout, const float* w_hh, const float* gates_ws, const float* cell_ws, int64_t T, int64_t B, int64_t H,
                              float* dgx, cudaStream_t stream);

extern "C" int sos_lstm_forward(const float* gx, const float* w_hh, int64_t T, int64_t B, int64_t H, float* out, float* gates_ws,
                                float* cell_ws, cudaStream_t stream) {
  SOS_CHECK_ARG(gx && w_hh && out && gates_ws && cell_ws && T > 0 && B > 0 && H > 0 && H <= 512, "sos_lstm_forward: bad arguments");
  if (int c = sos_lstm_forward_cluster(gx, w_hh, T, B, H, out, gates_ws, cell_ws, stream)) return c < 0 ? c : SOS_OK;   // cluster + DSMEM path
  const size_t smem = ((size_t)H * kHsPitch + (size_t)32 * (H + 1) + 32 * (kBt + 1)) * sizeof(float);
  SOS_CHECK_ARG(smem <= 200 * 1024, "sos_lstm_forward: hidden size too large for shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(lstm_fwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid(ceil_div((int)H, kUnits), 2);
  SOS_CHECK_ARG((int)(grid.x * grid.y) <= sos_num_sms(), "sos_lstm_forward: the persistent grid (%d blocks) must fit the device's %d SMs",
                (int)(grid.x * grid.y), sos_num_sms());
  unsigned int* sync = next_sync_slot(stream);
  if (!sync) {
    sos_set_error("sos_lstm_forward: cannot allocate the barrier counters");
    return SOS_ERR_CUDA;
  }
  // the kernel's blocks meet at a global-memory barrier between time steps: a COOPERATIVE launch, so that the driver guarantees
  // their co-residency (or refuses) whatever else runs on the device (another stream's weight gradients, NCCL, the other agent)
  int Ti = (int)T, Bi = (int)B, Hi = (int)H;
  void* args[] = {(void*)&gx, (void*)&w_hh, (void*)&Ti, (void*)&Bi, (void*)&Hi, (void*)&out, (void*)&gates_ws, (void*)&cell_ws, (void*)&sync};
  if (cudaLaunchCooperativeKernel((const void*)lstm_fwd_step_kernel, grid, dim3(kLstmThreads), args, smem, stream) != cudaSuccess) {
    sos_set_error("sos_lstm_forward: cooperative launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SOS_ERR_CUDA;
  }
  return SOS_OK;
}

extern "C" int sos_lstm_backward(const float* dout, const float* w_hh, const float* out, const float* gates_ws,
                                 const float* cell_ws, int64_t T, int64_t B, int64_t H, float* dgx, float* dh_ws, float* dc_ws,
                                 cudaStream_t stream) {
  (void)out;
  (void)dh_ws;
  SOS_CHECK_ARG(dout && w_hh && gates_ws && cell_ws && dgx && dc_ws && T > 0 && B > 0 && H > 0 && H <= 512,
                "sos_lstm_backward: bad arguments");
  if (int c = sos_lstm_backward_cluster(dout, w_hh, gates_ws, cell_ws, T, B, H, dgx, stream)) return c < 0 ? c : SOS_OK;
  const size_t smem = ((size_t)kUnits * 4 * H + (size_t)std::min<int64_t>(4 * H, kBwdChunk) * kGsPitch + 8 * kUnits * kBt) * sizeof(float);
  SOS_CHECK_ARG(smem <= 200 * 1024, "sos_lstm_backward: hidden size too large for shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid(ceil_div((int)H, kUnits), 2);
  SOS_CHECK_ARG((int)(grid.x * grid.y) <= sos_num_sms(), "sos_lstm_backward: the persistent grid (%d blocks) must fit the device's %d SMs",
                (int)(grid.x * grid.y), sos_num_sms());
  unsigned int* sync = next_sync_slot(stream);
  if (!sync) {
    sos_set_error("sos_lstm_backward: cannot allocate the barrier counters");
    return SOS_ERR_CUDA;
  }
  int Ti = (int)T, Bi = (int)B, Hi = (int)H;
  void* args[] = {(void*)&dout, (void*)&w_hh, (void*)&gates_ws, (void*)&cell_ws, (void*)&Ti, (void*)&Bi, (void*)&Hi, (void*)&dgx, (void*)&dc_ws, (void*)&sync};
  if (cudaLaunchCooperativeKernel((const void*)lstm_bwd_step_kernel, grid, dim3(kLstmThreads), args, smem, stream) != cudaSuccess) {
    sos_set_error("sos_lstm_backward: cooperative launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SOS_ERR_CUDA;
  }
  return SOS_OK;
}
