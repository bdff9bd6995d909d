// TMA throughput probe (sm_100a): how fast can one SM / the whole chip stream boxes from L2/HBM into shared memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcuda
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
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Params {
  CUtensorMap map;
  int mode;          // 0 tensor 3d, 1 bulk 1d
  int box_bytes;     // bytes per load
  int depth;         // loads in flight
  int iters;
  int distinct;      // 1: each CTA walks its own region, 0: all CTAs read the same boxes
  int n1, n2;        // tensor coordinates extents in boxes (dim1, dim2) for walking
  int box1, box2;
  const char* base;  // for bulk
  long long region_bytes;
  long long* cycles;
};

__global__ void __launch_bounds__(32, 1) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base;  // 64 barriers
  const uint32_t data0 = base + 1024;
  const uint32_t slot = (p.box_bytes + 1023) & ~1023;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.depth; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    const int cta = p.distinct ? blockIdx.x : 0;
    // two half rings: issue a whole half (depth/2 loads on ONE barrier), then wait for the other half
    const int half = p.depth / 2;
    int it = 0;
    long long t_issue = 0;
    for (int round = 0; round * half < p.iters + half; ++round) {
      const int h = round & 1;
      if (round >= 2) mbar_wait(bar0 + 8 * h, ((round >> 1) - 1) & 1);
      if (round * half < p.iters) {
        long long ti = clock64();
        mbar_expect_tx(bar0 + 8 * h, p.box_bytes * half);
        for (int k = 0; k < half; ++k, ++it) {
          const int s = h * half + k;
          if (p.mode == 0) {
            const int b = it % (p.n1 * p.n2);
            tma_load_3d(data0 + s * slot, &p.map, bar0 + 8 * h, 0, (b % p.n1) * p.box1, (cta * p.n2 + b / p.n1) * p.box2);
          } else {
            const long long off = ((long long)it * p.box_bytes) & (p.region_bytes - 1);
            bulk_load_1d(data0 + s * slot, p.base + (long long)cta * p.region_bytes + off, p.box_bytes, bar0 + 8 * h);
          }
        }
        t_issue += clock64() - ti;
      }
    }
    p.cycles[gridDim.x + blockIdx.x] = t_issue;
    p.cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  const size_t total = 12ull << 30;
  char* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  long long* cyc;
  CK(cudaMalloc(&cyc, 2 * sms * 8));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  struct Case { const char* name; int mode, rowb, box1, box2, pix_stride, depth, distinct; };
  // tensor: dim0 = rowb bytes of a pixel (pix_stride bytes apart), dim1 = pixels along W (box1), dim2 = rows along H (box2; W pixels per row = 256)
  std::vector<Case> cases = {
      {"A 128B x 8 x 20 (stride 384) d4 distinct", 0, 128, 8, 20, 384, 4, 1},
      {"A 128B x 8 x 20 (stride 384) d8 distinct", 0, 128, 8, 20, 384, 8, 1},
      {"A  64B x 8 x 20 (stride 384) d8 distinct", 0, 64, 8, 20, 384, 8, 1},
      {"A 128B x 20 x 20 (stride 384) d4 distinct", 0, 128, 20, 20, 384, 4, 1},
      {"A 128B x 8 x 20 (stride 128 planar) d8 distinct", 0, 128, 8, 20, 128, 8, 1},
      {"A 128B x 32 x 20 (stride 128 planar) d4 distinct", 0, 128, 32, 20, 128, 4, 1},
      {"B 128B x 96 rows (stride 9600) d8 shared", 0, 128, 96, 1, 9600, 8, 0},
      {"B 128B x 96 rows (stride 9600) d8 distinct", 0, 128, 96, 1, 9600, 8, 1},
      {"B  64B x 96 rows (stride 9600) d8 shared", 0, 64, 96, 1, 9600, 8, 0},
      {"bulk1d 12KB d8 shared", 1, 0, 12288, 0, 0, 8, 0},
      {"bulk1d 12KB d8 distinct", 1, 0, 12288, 0, 0, 8, 1},
      {"bulk1d 12KB d4 shared", 1, 0, 12288, 0, 0, 4, 0},
      {"bulk1d 48KB d4 distinct", 1, 0, 49152, 0, 0, 4, 1},
      {"bulk1d 4KB d16 shared", 1, 0, 4096, 0, 0, 16, 0},
  };
  for (auto& c : cases) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.mode = c.mode; p.depth = c.depth; p.distinct = c.distinct; p.cycles = cyc; p.base = buf;
    long long per_cta;
    if (c.mode == 0) {
      const int W = 256;                 // pixels per image row
      // tensor dims: {rowb/4 floats, W pixels, H rows}; H big enough for all CTAs
      p.box_bytes = c.rowb * c.box1 * c.box2;
      p.box1 = c.box1; p.box2 = c.box2;
      p.n1 = W / c.box1; p.n2 = 8;      // each CTA walks n1 x n2 boxes (re-walks -> L2 hits)
      const long long rows_total = (long long)sms * p.n2 * c.box2;
      cuuint64_t gd[3] = {(cuuint64_t)(c.rowb / 4), (cuuint64_t)W, (cuuint64_t)rows_total};
      cuuint64_t gs[2] = {(cuuint64_t)c.pix_stride, (cuuint64_t)c.pix_stride * W};
      if (c.pix_stride == 9600) { gd[1] = 96 * 8; gs[1] = 9600ull * 96 * 8; p.n1 = 8; p.n2 = 1; gd[2] = sms; }
      cuuint32_t bx[3] = {(cuuint32_t)(c.rowb / 4), (cuuint32_t)c.box1, (cuuint32_t)c.box2};
      cuuint32_t es[3] = {1, 1, 1};
      if ((long long)gs[1] * (long long)gd[2] > (long long)total) { printf("%s: too big\n", c.name); continue; }
      CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       c.rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
      per_cta = (long long)p.n1 * p.n2 * p.box_bytes;
    } else {
      p.box_bytes = c.box1;
      p.region_bytes = 1 << 20;
      per_cta = p.region_bytes;
    }
    p.iters = 2048;
    const int smem = 1024 + 1024 + c.depth * ((p.box_bytes + 1023) & ~1023);
    if (smem > 227 * 1024) { printf("%s: smem too big\n", c.name); continue; }
    for (int grid : {sms, 1}) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<grid, 32, smem>>>(p);      // warm L2
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    probe<<<grid, 32, smem>>>(p);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(2 * grid);
    CK(cudaMemcpy(h.data(), cyc, 2 * grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0, iss = 0; for (int i = 0; i < grid; ++i) { avg += h[i]; iss += h[grid + i]; } avg /= grid; iss /= grid;
    const double bytes = (double)p.iters * p.box_bytes;
    printf("%-50s grid %3d box %6d B: %6.1f B/cyc/SM, %7.1f cyc/load (issue %5.1f), chip %6.2f TB/s\n", c.name, grid, p.box_bytes, bytes / avg, avg / p.iters, iss / p.iters,
           bytes * grid / (ms * 1e-3) / 1e12);
    }
  }
  // single-SM run of two cases to separate per-SM engine limits from chip (L2) limits
  return 0;
}
