// Host-side planning helpers shared by the tensor-core tap GEMM (conv_tc.cu) and its weight gradient (wgrad_tc.cu):
// tap grouping (taps that differ only by a shift along the slow tile axis share one staged activation box),
// dilation lattices and TMA tensor-map encoding.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace tc {

constexpr int kMaxGroups = 49;
constexpr int kMaxSub = 7;
constexpr int kSmemLimit = 227 * 1024;

struct TapGroup {
  int16_t d_fast, d_slow;      // coordinate offsets of the activation box (tensor elements)
  int8_t n_sub;                // taps served by this box
  int8_t a_off[kMaxSub];       // slow-axis shift of each tap, in units of FB rows
  int16_t tap[kMaxSub];        // weight tap index (k offset = tap * cin)
};

// Geometry of one tap-list operator, independent of which of the three GEMMs (fwd / dgrad / wgrad) is run.
struct Geometry {
  int ntaps;
  const int32_t* tap_dh;
  const int32_t* tap_dw;
  int H, W, OH, OW, stride;
  int Cin, Cout;
  bool lattice_out;            // output addressed densely (no sub-pixel scatter) -> dilation lattices allowed
  int max_sub;                 // most taps one staged box may serve (<= kMaxSub)
  bool allow_wide = false;     // 16 x 8 (fast x slow) tiles for short dilation lattices (forward / dgrad kernel only)
  int esz = 4;                 // operand element size in bytes: 4 (TF32 in fp32 storage) or 2 (half)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* estr, CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    sos_set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return SOS_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i + 1];
  CUresult r = fn(m, dt, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sos_set_error("cuTensorMapEncodeTiled(%s) failed with %d: rank %d dims [%llu %llu %llu %llu %llu] strides [%llu %llu %llu %llu] box [%u %u %u %u %u]",
                  what, (int)r, rank, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)(rank > 2 ? gd[2] : 0),
                  (unsigned long long)(rank > 3 ? gd[3] : 0), (unsigned long long)(rank > 4 ? gd[4] : 0), (unsigned long long)gs[0],
                  (unsigned long long)(rank > 2 ? gs[1] : 0), (unsigned long long)(rank > 3 ? gs[2] : 0),
                  (unsigned long long)(rank > 4 ? gs[3] : 0), bx[0], bx[1], rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0,
                  rank > 4 ? bx[4] : 0);
    return SOS_ERR_CUDA;
  }
  return SOS_OK;
}

// Geometry of a tensor map without its base address: cached per plan, encoded again only when the base pointer changes.
struct MapSpec {
  CUtensorMapDataType dt;
  int rank;
  uint64_t dims[5], strides[5];
  uint32_t box[5], es[5];
  CUtensorMapSwizzle sw;
  const char* what;
};
inline MapSpec make_spec(CUtensorMapDataType dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                         const uint32_t* estr, CUtensorMapSwizzle sw, const char* what) {
  MapSpec s{};
  s.dt = dt; s.rank = rank; s.sw = sw; s.what = what;
  for (int i = 0; i < rank; ++i) { s.dims[i] = dims[i]; s.strides[i] = strides_bytes[i]; s.box[i] = box[i]; s.es[i] = estr[i]; }
  return s;
}
inline int encode_spec(CUtensorMap* m, const MapSpec& s, const void* base) {
  return encode_map(m, s.dt, s.rank, base, s.dims, s.strides, s.box, s.es, s.sw, s.what);
}

inline int gcd_i(int a, int b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { int t = a % b; a = b; b = t; } return a; }
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Plan {
  bool fast_is_w = true;
  int g = 1;                 // slow-axis lattice step
  bool share = false;
  int FB = 8, SB = 16;
  int S = 1;
  std::vector<TapGroup> groups;
  int halo = 0;              // extra slow rows in the activation box
  double cost = 1e300;
};

// Build the tap groups for one orientation / sharing mode; returns false if not applicable.
inline bool build_plan(const Geometry& a, bool fast_is_w, bool share, Plan& pl) {
  const int nt = (int)a.ntaps;
  const int in_slow = fast_is_w ? (int)a.H : (int)a.W, out_slow = fast_is_w ? (int)a.OH : (int)a.OW;
  const int out_fast = fast_is_w ? (int)a.OW : (int)a.OH;
  pl.fast_is_w = fast_is_w;
  pl.share = share;
  pl.groups.clear();
  pl.g = 1;
  pl.halo = 0;
  pl.FB = 8;
  pl.SB = 16;
  if (out_slow == 1) {          // plain GEMM rows: one 128-row strip along the fast axis
    pl.FB = 128;
    pl.SB = 1;
  }
  std::vector<int> foff(nt), soff(nt);
  for (int t = 0; t < nt; ++t) {
    foff[t] = fast_is_w ? a.tap_dw[t] : a.tap_dh[t];
    soff[t] = fast_is_w ? a.tap_dh[t] : a.tap_dw[t];
  }
  if (!share) {
    if (nt > kMaxGroups) return false;
    for (int t = 0; t < nt; ++t) {
      TapGroup gq{};
      gq.d_fast = (int16_t)foff[t];
      gq.d_slow = (int16_t)soff[t];
      gq.n_sub = 1;
      gq.a_off[0] = 0;
      gq.tap[0] = (int16_t)t;
      pl.groups.push_back(gq);
    }
  } else {
    if (a.stride != 1 || pl.SB == 1) return false;
    // lattice step = gcd of all slow offsets (they must all be multiples of it)
    int g = 0;
    for (int t = 0; t < nt; ++t) g = gcd_i(g, soff[t]);
    if (g == 0) g = 1;
    const bool lattice_out = a.lattice_out;
    if (g > 1 && (!lattice_out || in_slow % g != 0 || out_slow % g != 0)) return false;
    pl.g = g;
    if (a.allow_wide && out_slow / g <= 8 && out_fast >= 16) {   // a lattice of <= 8 rows would half-fill a 16-row tile
      pl.FB = 16;
      pl.SB = 8;
    }
    // group taps by fast offset
    std::vector<int> order(nt);
    for (int t = 0; t < nt; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return foff[x] != foff[y] ? foff[x] < foff[y] : soff[x] < soff[y]; });
    size_t i = 0;
    while (i < order.size()) {
      size_t j = i;
      const int base = soff[order[i]] / g;
      TapGroup gq{};
      gq.d_fast = (int16_t)foff[order[i]];
      gq.d_slow = (int16_t)base;
      while (j < order.size() && foff[order[j]] == foff[order[i]] && gq.n_sub < std::min(kMaxSub, a.max_sub) && soff[order[j]] / g - base <= 6) {
        gq.a_off[gq.n_sub] = (int8_t)(soff[order[j]] / g - base);
        gq.tap[gq.n_sub] = (int16_t)order[j];
        pl.halo = std::max(pl.halo, (int)gq.a_off[gq.n_sub]);
        ++gq.n_sub;
        ++j;
      }
      pl.groups.push_back(gq);
      i = j;
    }
    if ((int)pl.groups.size() > kMaxGroups) return false;
  }
  // sub-tiles sharing the weight boxes
  const int N = std::min(256, round_up((int)a.Cout, 16));
  const int tiles_fast = ceil_div(out_fast, pl.FB);
  pl.S = (2 * N <= 256 && tiles_fast >= 2) ? 2 : 1;
  // cost model: per output pixel, max(tensor time, L2 feed time), divided by tile utilisation
  const int kpe = 32 / a.esz;   // K elements per MMA (32 operand bytes per row)
  const int cbe = a.Cin % (4 * kpe) == 0 ? 4 * kpe : (a.Cin % (2 * kpe) == 0 ? 2 * kpe : kpe);
  const int n_chunks = (int)a.Cin / cbe;
  const double mma = (double)nt * n_chunks * (cbe / kpe) * pl.S * (128.0 * N / 256.0);
  const double bytes = ((double)pl.groups.size() * pl.S * (pl.SB + pl.halo) * pl.FB + (double)nt * N) * cbe * (double)a.esz * n_chunks;
  const double t = std::max(mma, bytes / 40.0);
  const int lat_slow = out_slow / pl.g;
  const double util = ((double)lat_slow / (ceil_div(lat_slow, pl.SB) * pl.SB)) *
                      ((double)out_fast / (ceil_div(tiles_fast, pl.S) * pl.S * pl.FB));
  pl.cost = t / (util * pl.S);
  return true;
}


}  // namespace tc
