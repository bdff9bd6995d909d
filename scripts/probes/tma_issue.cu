// How expensive is it to ISSUE a TMA load, and does issuing from several lanes / warps help?  (sm_100a)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
struct Params {
  CUtensorMap map2, map5;
  const CUtensorMap* gmap2;     // copy in global memory
  int variant, rounds, per_round, rows;   // per_round loads of (rows x 128 B) each
  long long* cycles;
};
// variants: 0 = one thread issues all loads of a round (2D map, param space)
//           1 = same, tensormap in global memory
//           2 = lanes 0..per_round-1 of warp 0 issue one load each
//           3 = lane 0 of warps 0..per_round-1 issue one load each
//           4 = one thread, 5D map (box {32, 8, rows/8, 1, 1})
//           5 = one thread, no prefetch.tensormap (others prefetch)
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base, data0 = base + 1024;
  const uint32_t box = p.rows * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.variant != 5) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.map2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&p.map5) : "memory");
    }
  }
  __syncthreads();
  long long t0 = clock64(), t_issue = 0;
  const CUtensorMap* m2 = p.variant == 1 ? p.gmap2 : &p.map2;
  for (int r = 0; r < p.rounds + 2; ++r) {
    const int h = r & 1;
    const uint32_t bar = bar0 + 8 * h;
    const uint32_t dst0 = data0 + h * p.per_round * box;
    if (r >= 2) mbar_wait(bar, ((r >> 1) - 1) & 1);            // everyone waits (keeps the roles in step)
    __syncthreads();
    if (r < p.rounds) {
      if (threadIdx.x == 0) mbar_expect_tx(bar, box * p.per_round);
      __syncthreads();
      long long ti = clock64();
      const int row0 = ((r * p.per_round) % 64) * p.rows;
      if (p.variant == 0 || p.variant == 1 || p.variant == 5) {
        if (threadIdx.x == 0) for (int k = 0; k < p.per_round; ++k) tma_load_2d(dst0 + k * box, m2, bar, 0, row0 + k * p.rows);
      } else if (p.variant == 2) {
        if (warp == 0 && lane < p.per_round) tma_load_2d(dst0 + lane * box, m2, bar, 0, row0 + lane * p.rows);
      } else if (p.variant == 3) {
        if (lane == 0 && warp < p.per_round) tma_load_2d(dst0 + warp * box, m2, bar, 0, row0 + warp * p.rows);
      } else if (p.variant == 4) {
        if (threadIdx.x == 0) for (int k = 0; k < p.per_round; ++k) tma_load_5d(dst0 + k * box, &p.map5, bar, 0, 0, (row0 + k * p.rows) / 8, 0, 0);
      }
      if (threadIdx.x == 0) t_issue += clock64() - ti;
    }
  }
  if (threadIdx.x == 0) { p.cycles[blockIdx.x] = clock64() - t0; p.cycles[gridDim.x + blockIdx.x] = t_issue; }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  char* buf; CK(cudaMalloc(&buf, 64 << 20)); CK(cudaMemset(buf, 1, 64 << 20));
  long long* cyc; CK(cudaMalloc(&cyc, 2 * sms * 8));
  CUtensorMap* gmap; CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  for (int rows : {8, 32, 96}) {
    for (int per_round : {1, 4, 8}) {
      if (2 * per_round * rows * 128 > 200 * 1024) continue;
      Params p; memset(&p, 0, sizeof(p));
      p.rows = rows; p.per_round = per_round; p.rounds = 512; p.cycles = cyc; p.gmap2 = gmap;
      {
        cuuint64_t gd[2] = {32, 65536}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {32, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        if (enc(&p.map2, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc2 failed\n"); return 1; }
        cuuint64_t gd5[5] = {32, 8, 8192, 1, 1}; cuuint64_t gs5[4] = {128, 1024, 8192 * 1024, 8192 * 1024}; cuuint32_t bx5[5] = {32, 8, (cuuint32_t)(rows / 8), 1, 1}; cuuint32_t es5[5] = {1, 1, 1, 1, 1};
        if (enc(&p.map5, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, buf, gd5, gs5, bx5, es5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc5 failed\n"); return 1; }
        CK(cudaMemcpy(gmap, &p.map2, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
      }
      for (int variant = 0; variant < 6; ++variant) {
        if (variant == 3 && per_round > 8) continue;
        p.variant = variant;
        for (int grid : {1, sms}) {
          const int smem = 2048 + 2 * per_round * rows * 128;
          probe<<<grid, 256, smem>>>(p); CK(cudaDeviceSynchronize());
          probe<<<grid, 256, smem>>>(p); CK(cudaDeviceSynchronize());
          std::vector<long long> h(2 * grid);
          CK(cudaMemcpy(h.data(), cyc, 2 * grid * 8, cudaMemcpyDeviceToHost));
          double avg = 0, iss = 0; for (int i = 0; i < grid; ++i) { avg += h[i]; iss += h[grid + i]; } avg /= grid; iss /= grid;
          const double loads = (double)p.rounds * per_round;
          printf("rows %3d x128B  per_round %d  variant %d  grid %3d: %7.1f cyc/load total, %6.1f cyc/load issue, %6.1f B/cyc/SM\n", rows, per_round, variant, grid, avg / loads, iss / loads, loads * rows * 128 / avg);
        }
      }
    }
  }
  return 0;
}
