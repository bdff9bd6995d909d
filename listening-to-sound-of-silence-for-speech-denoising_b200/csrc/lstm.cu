// Bidirectional single-layer LSTM recurrence (nn.LSTM of M1/networks.py:95,147-148 and M2/networks.py:64,88-89),
// PyTorch gate order i, f, g, o.  The input projection x W_ih^T + b_ih + b_hh is a plain GEMM done by the caller;
// these kernels run the serial part: one launch per time step, both directions in the same launch
// (direction 0 walks t = 0..T-1, direction 1 walks t = T-1..0), so the only global synchronisation is the
// kernel boundary.  Per step a block owns kUnits hidden units of one direction for the whole batch and keeps
// its 4*kUnits rows of W_hh in shared memory.
#include "common.cuh"
#include "sos_b200.h"

namespace {

constexpr int kUnits = 8;        // hidden units per block
constexpr int kLstmThreads = 256;
constexpr int kBt = 32;          // batch tile held in shared memory

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gx (T,B,2,4H); w_hh (2,4H,H); out (T,B,2H); gates (T,B,2,4H) activated; cell (T,B,2,H)
__global__ void __launch_bounds__(kLstmThreads) lstm_fwd_step_kernel(const float* __restrict__ gx, const float* __restrict__ w_hh,
                                                                     int T, int B, int H, int step, float* __restrict__ out,
                                                                     float* __restrict__ gates, float* __restrict__ cell) {
  extern __shared__ float sm[];
  const int d = blockIdx.y;
  const int j0 = blockIdx.x * kUnits;
  const int t = d == 0 ? step : T - 1 - step;
  const int tp = d == 0 ? t - 1 : t + 1;             // previous time step of this direction
  const bool first = step == 0;
  const int Hp = H + 1;
  float* w_s = sm;                                   // [32][H+1]   row r = q*kUnits + jj
  float* h_s = w_s + 32 * Hp;                        // [kBt][H]
  float* g_s = h_s + kBt * H;                        // [32][kBt+1]
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
  for (int b0 = 0; b0 < B; b0 += kBt) {
    __syncthreads();
    if (!first) {
      for (int e = tid; e < kBt * H; e += kLstmThreads) {
        const int bb = e / H, k = e - bb * H;
        h_s[e] = (b0 + bb < B) ? out[((size_t)tp * B + b0 + bb) * 2 * H + (size_t)d * H + k] : 0.f;
      }
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (!first) {
      const float* wr = w_s + r * Hp;
      const float* hb = h_s + (bg * 4) * H;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float w = wr[k];
        acc[0] = fmaf(w, hb[k], acc[0]);
        acc[1] = fmaf(w, hb[H + k], acc[1]);
        acc[2] = fmaf(w, hb[2 * H + k], acc[2]);
        acc[3] = fmaf(w, hb[3 * H + k], acc[3]);
      }
    }
    const int j = j0 + jj;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + bg * 4 + i;
      float v = 0.f;
      if (b < B && j < H) v = acc[i] + gx[(((size_t)t * B + b) * 2 + d) * 4 * H + (size_t)q * H + j];
      g_s[r * (kBt + 1) + bg * 4 + i] = v;
    }
    __syncthreads();
    // cell update: thread -> (unit jj2, batch bb)
    const int jj2 = tid & (kUnits - 1), bb = tid >> 3;
    const int j2 = j0 + jj2, b = b0 + bb;
    if (j2 < H && b < B) {
      const float ig = sigmoidf_(g_s[(0 * kUnits + jj2) * (kBt + 1) + bb]);
      const float fg = sigmoidf_(g_s[(1 * kUnits + jj2) * (kBt + 1) + bb]);
      const float gg = tanhf(g_s[(2 * kUnits + jj2) * (kBt + 1) + bb]);
      const float og = sigmoidf_(g_s[(3 * kUnits + jj2) * (kBt + 1) + bb]);
      const float cp = first ? 0.f : cell[(((size_t)tp * B + b) * 2 + d) * H + j2];
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
}

// One backward step.  dgx (T,B,2,4H) receives pre-activation gate gradients; dh_ws is unused storage kept for
// ABI stability; dc_ws (B,2,H) carries dc across steps (zeroed by the first step).
// The recurrent term dh_rec[b][k] = sum_r dgx[t_next][b][d][r] * w_hh[d][r][k] is recomputed from the previous
// launch's dgx, so no intra-launch grid synchronisation is needed.
__global__ void __launch_bounds__(kLstmThreads) lstm_bwd_step_kernel(const float* __restrict__ dout, const float* __restrict__ w_hh,
                                                                     const float* __restrict__ gates, const float* __restrict__ cell,
                                                                     int T, int B, int H, int step, float* __restrict__ dgx,
                                                                     float* __restrict__ dc_ws) {
  extern __shared__ float sm[];
  const int d = blockIdx.y;
  const int k0 = blockIdx.x * kUnits;
  // backward walks each direction's time axis in reverse
  const int t = d == 0 ? T - 1 - step : step;
  const int tn = d == 0 ? t + 1 : t - 1;             // the step processed by the previous launch (later in recurrence)
  const int tp = d == 0 ? t - 1 : t + 1;             // earlier step in recurrence (for c_prev)
  const bool first = step == 0;
  const bool has_prev = d == 0 ? (t > 0) : (t < T - 1);
  const int tid = threadIdx.x;
  const int R = 4 * H;
  float* w_s = sm;                                   // [kUnits][R]  w_hh[d][r][k0+kk] transposed slice
  float* g_s = w_s + kUnits * R;                     // [kBt][H+1]   one gate chunk of dgx[tn]
  const int kk = tid & (kUnits - 1), bb = tid >> 3;  // thread -> (unit kk, batch bb)
  if (!first) {
    for (int e = tid; e < kUnits * R; e += kLstmThreads) {
      const int r = e / kUnits, u = e - r * kUnits;  // coalescing is poor (stride H) but the slice is small and L2 resident
      w_s[u * R + r] = (k0 + u < H) ? w_hh[((size_t)d * R + r) * H + k0 + u] : 0.f;
    }
  }
  for (int b0 = 0; b0 < B; b0 += kBt) {
    float dh_rec = 0.f;
    if (!first) {
      for (int qc = 0; qc < 4; ++qc) {
        __syncthreads();
        for (int e = tid; e < kBt * H; e += kLstmThreads) {
          const int b2 = e / H, rr = e - b2 * H;
          g_s[b2 * (H + 1) + rr] = (b0 + b2 < B) ? dgx[(((size_t)tn * B + b0 + b2) * 2 + d) * R + (size_t)qc * H + rr] : 0.f;
        }
        __syncthreads();
        const float* wr = w_s + kk * R + qc * H;
        const float* gr = g_s + bb * (H + 1);
#pragma unroll 4
        for (int rr = 0; rr < H; ++rr) dh_rec = fmaf(gr[rr], wr[rr], dh_rec);
      }
    }
    const int j = k0 + kk, b = b0 + bb;
    if (j < H && b < B) {
      const size_t gi = (((size_t)t * B + b) * 2 + d) * R + j;
      const float ig = gates[gi], fg = gates[gi + H], gg = gates[gi + 2 * H], og = gates[gi + 3 * H];
      const float c = cell[(((size_t)t * B + b) * 2 + d) * H + j];
      const float cp = has_prev ? cell[(((size_t)tp * B + b) * 2 + d) * H + j] : 0.f;
      const float tc = tanhf(c);
      const float dh = dout[((size_t)t * B + b) * 2 * H + (size_t)d * H + j] + dh_rec;
      const size_t ci = ((size_t)b * 2 + d) * H + j;
      const float dc = dh * og * (1.f - tc * tc) + (first ? 0.f : dc_ws[ci]);
      dgx[gi] = dc * gg * ig * (1.f - ig);
      dgx[gi + H] = dc * cp * fg * (1.f - fg);
      dgx[gi + 2 * H] = dc * ig * (1.f - gg * gg);
      dgx[gi + 3 * H] = dh * tc * og * (1.f - og);
      dc_ws[ci] = dc * fg;
    }
  }
}

}  // namespace

extern "C" int sos_lstm_forward(const float* gx, const float* w_hh, int64_t T, int64_t B, int64_t H, float* out, float* gates_ws,
                                float* cell_ws, cudaStream_t stream) {
  SOS_CHECK_ARG(gx && w_hh && out && gates_ws && cell_ws && T > 0 && B > 0 && H > 0 && H <= 512, "sos_lstm_forward: bad arguments");
  const size_t smem = ((size_t)32 * (H + 1) + (size_t)kBt * H + 32 * (kBt + 1)) * sizeof(float);
  SOS_CHECK_ARG(smem <= 200 * 1024, "sos_lstm_forward: hidden size too large for shared memory");
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(lstm_fwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid(ceil_div((int)H, kUnits), 2);
  for (int s = 0; s < (int)T; ++s)
    lstm_fwd_step_kernel<<<grid, kLstmThreads, smem, stream>>>(gx, w_hh, (int)T, (int)B, (int)H, s, out, gates_ws, cell_ws);
  SOS_CHECK_LAUNCH("sos_lstm_forward");
  return SOS_OK;
}

extern "C" int sos_lstm_backward(const float* dout, const float* w_hh, const float* out, const float* gates_ws,
                                 const float* cell_ws, int64_t T, int64_t B, int64_t H, float* dgx, float* dh_ws, float* dc_ws,
                                 cudaStream_t stream) {
  (void)out;
  (void)dh_ws;
  SOS_CHECK_ARG(dout && w_hh && gates_ws && cell_ws && dgx && dc_ws && T > 0 && B > 0 && H > 0 && H <= 512,
                "sos_lstm_backward: bad arguments");
  const size_t smem = ((size_t)kUnits * 4 * H + (size_t)kBt * (H + 1)) * sizeof(float);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid(ceil_div((int)H, kUnits), 2);
  for (int s = 0; s < (int)T; ++s)
    lstm_bwd_step_kernel<<<grid, kLstmThreads, smem, stream>>>(dout, w_hh, gates_ws, cell_ws, (int)T, (int)B, (int)H, s, dgx, dc_ws);
  SOS_CHECK_LAUNCH("sos_lstm_backward");
  return SOS_OK;
}
