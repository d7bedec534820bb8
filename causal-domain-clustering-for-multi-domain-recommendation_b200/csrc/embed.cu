// Sparse-field embedding: gather (a1), deterministic sorted-segment backward (a2) and the fused
// reference-exact dense Adam sweep (a2+a7+a17).  HBM-bound integer/byte work: 128-bit accesses,
// E/4 lanes per row so a warp touches 32/(E/4) consecutive rows of w/m/v (fully coalesced sweep).
#include "common.cuh"
#include <cub/cub.cuh>
#include <type_traits>

namespace cdcmdr {

constexpr int kShortSeg = 64;        // segments up to this many entries are summed inline, in order
constexpr int kLongThreads = 256;    // one CTA per PIECE of a long segment
constexpr int kPiece = 512;          // long segments are cut into pieces of this many entries (a Zipf-hot id or the single
                                     // domain id of a batch is a segment of up to B entries: one CTA per segment serialises)
constexpr int kRegPartials = 2048;

struct EmbedPlan {                   // lives at the start of the caller-provided workspace
  int64_t n, V, n_long_max;
  int E_max;
  size_t off_keys_in, off_vals_in, off_keys, off_vals, off_uniq, off_cnt, off_start, off_nuniq,
      off_seg_of_row, off_long_slot, off_nlong, off_long_sum, off_long_seg, off_long_off, off_cub, off_reg, off_gsum;
  size_t cub_bytes, total;
};

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static int sort_bits(int64_t V) { int b = 1; while ((int64_t(1) << b) <= V) ++b; return b; }  // keys in [0, V]

static size_t cub_temp_bytes(int64_t n, int64_t V) {
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, sort_bits(V));
  cub::DeviceRunLengthEncode::Encode(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                     (int32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 1024;
}

static EmbedPlan make_layout(int64_t n, int64_t V, int E_max) {
  EmbedPlan p{};
  p.n = n; p.V = V; p.E_max = E_max; p.n_long_max = n / kShortSeg + n / kPiece + 2;   // pieces <= n/kPiece + #long segments
  size_t o = align256(sizeof(EmbedPlan));
  auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
  p.off_keys_in = take(n * 4); p.off_vals_in = take(n * 4);
  p.off_keys = take(n * 4);    p.off_vals = take(n * 4);
  p.off_uniq = take(n * 4);    p.off_cnt = take(n * 4);   p.off_start = take((n + 1) * 4);
  p.off_nuniq = take(16);      p.off_seg_of_row = take((size_t)V * 4);
  p.off_long_slot = take(n * 4); p.off_nlong = take(16);
  p.off_long_sum = take((size_t)p.n_long_max * E_max * 4);
  p.off_long_seg = take((size_t)p.n_long_max * 4);
  p.off_long_off = take((size_t)p.n_long_max * 4);
  p.off_reg = take(kRegPartials * 8);
  p.off_gsum = take((size_t)n * E_max * 4);          // per-segment gradient sums of the dense-exact update (<= n segments)
  p.cub_bytes = cub_temp_bytes(n, V);
  p.off_cub = take(p.cub_bytes);
  p.total = o;
  return p;
}

template <typename T> static T* at(const void* base, size_t off) { return (T*)((char*)base + off); }

// ------------------------------------------------------------------------------------------ gather
// Random 64..256-byte row reads.  ILP > 1 makes a thread issue the loads of several independent (sample, field) rows before it
// converts / stores them; measured on B200 at C4 it did not help (34 us at ILP 1 -> 44 us at ILP 4: the grid already keeps
// 64 warps per SM in flight), so the launcher uses ILP 1.
template <int VEC, int ILP>
__global__ void __launch_bounds__(256)
embed_gather_kernel(const int32_t* __restrict__ x, const int64_t* __restrict__ offsets,
                    const float* __restrict__ table, float* __restrict__ out_f32,
                    uint16_t* __restrict__ out_bf16, int64_t ld_bf16, int64_t B, int F, int E,
                    int64_t V, int* __restrict__ oob) {
  const int lanes = E / VEC;                       // lanes per (b,f) row
  const int64_t total = B * (int64_t)F * lanes;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * ILP) {
    int64_t bf[ILP], row[ILP]; int q[ILP]; bool ok[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      const int64_t i = i0 + u * stride;
      int64_t q64 = 0; bf[u] = 0;
      if (i < total) split_idx(i, lanes, bf[u], q64);
      q[u] = (int)q64;
      int64_t bb, f64;
      split_idx(bf[u], F, bb, f64);
      row[u] = i < total ? (int64_t)__ldg(x + bf[u]) + __ldg(offsets + f64) : -1;
      ok[u] = row[u] >= 0 && row[u] < V;
    }
    float v[ILP][VEC];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (ok[u]) {
        if (VEC == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(table + row[u] * E) + q[u]);
          v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w;
        } else {
          v[u][0] = __ldg(table + row[u] * E + q[u]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total) continue;
      if (!ok[u] && oob) *oob = 1;
      int64_t b, f64;
      split_idx(bf[u], F, b, f64);
      const int64_t col = f64 * E + q[u] * VEC;
      if (out_f32) {
        float* o = out_f32 + b * (int64_t)F * E + col;
        if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(v[u][0], v[u][1 % VEC], v[u][2 % VEC], v[u][3 % VEC]);
        else o[0] = v[u][0];
      }
      if (out_bf16) {
        uint16_t* o = out_bf16 + b * ld_bf16 + col;
        if (VEC == 4) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[u][0], v[u][1 % VEC]), pack_bf16x2(v[u][2 % VEC], v[u][3 % VEC]));
        else o[0] = f32_to_bf16(v[u][0]);
      }
    }
  }
}

// One thread per (sample, field) row: the index is read once per row, then the row's EV 128-bit loads are issued back to back
// (EV x 16 bytes in flight per thread - the lane-per-quad kernel above has 16 and runs out of resident threads: 2048 x 148
// threads x 16 B = 4.8 MB in flight bound it to ~4 TB/s on random 64-byte rows).  Same bytes, same order: bit-exact.
template <int EV>
__global__ void __launch_bounds__(256)
embed_gather_rows_kernel(const int32_t* __restrict__ x, const int64_t* __restrict__ offsets, const float* __restrict__ table,
                         float* __restrict__ out_f32, uint16_t* __restrict__ out_bf16, int64_t ld_bf16, int64_t B, int F,
                         int64_t V, int* __restrict__ oob) {
  constexpr int E = 4 * EV;
  const int64_t total = B * (int64_t)F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, f;
    split_idx(i, F, b, f);
    const int64_t row = (int64_t)__ldg(x + i) + __ldg(offsets + f);
    const bool ok = row >= 0 && row < V;
    float4 t[EV];
    const float4* src = reinterpret_cast<const float4*>(table + (ok ? row : 0) * E);
#pragma unroll
    for (int k = 0; k < EV; ++k) t[k] = __ldg(src + k);
    if (!ok) {
      if (oob) *oob = 1;
#pragma unroll
      for (int k = 0; k < EV; ++k) t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(out_f32 + i * E);
#pragma unroll
      for (int k = 0; k < EV; ++k) o[k] = t[k];
    }
    if (out_bf16) {
      uint4* o = reinterpret_cast<uint4*>(out_bf16 + b * ld_bf16 + f * E);
#pragma unroll
      for (int k = 0; k < EV / 2; ++k)
        o[k] = make_uint4(pack_bf16x2(t[2 * k].x, t[2 * k].y), pack_bf16x2(t[2 * k].z, t[2 * k].w),
                          pack_bf16x2(t[2 * k + 1].x, t[2 * k + 1].y), pack_bf16x2(t[2 * k + 1].z, t[2 * k + 1].w));
    }
  }
}

// Row-range sharded table over NVLink peer memory (SURVEY 8e / BASELINE configs[4]: 500 M x 64 rows over 8 GPUs): rank r owns rows
// [r*rows_per, (r+1)*rows_per) of the concatenated table (layer.py:140) in its own HBM and nothing else; every shard is mapped into
// every process, so the lookup is ONE kernel that reads each row where it lives - no index exchange, no row exchange, no packing,
// static shapes whatever the id distribution.  E/4 lanes per row (a 256-byte row of E = 64 is one contiguous 16-lane load: remote
// reads move whole 32-byte sectors), index loaded once per row group and broadcast by shuffle.  Same bytes as a local gather:
// bit-exact.
template <int LANES>
__global__ void __launch_bounds__(256)
embed_gather_peer_kernel(const int32_t* __restrict__ x, const int64_t* __restrict__ offsets, const float* const* __restrict__ shards,
                         int64_t rows_per, float* __restrict__ out_f32, uint16_t* __restrict__ out_bf16, int64_t ld_bf16, int64_t B,
                         int F, int E, int64_t V, int* __restrict__ oob) {
  const int64_t total = B * (int64_t)F * LANES;
  const int q = threadIdx.x % LANES;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bf = i / LANES;
    int64_t b, f;
    split_idx(bf, F, b, f);
    const int64_t row = (int64_t)__ldg(x + bf) + __ldg(offsets + f);
    const bool ok = row >= 0 && row < V;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const int64_t owner = row / rows_per;
      const float* src = shards[owner] + (row - owner * rows_per) * E + 4 * q;
      t = *reinterpret_cast<const float4*>(src);            // peer memory: plain (coherent) load, not the read-only path
    } else if (oob) {
      *oob = 1;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + bf * E + 4 * q) = t;
    if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + b * ld_bf16 + f * E + 4 * q) = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
  }
}

// ------------------------------------------------------------------------------------------ plan
__global__ void plan_keys_kernel(const int32_t* __restrict__ x, const int64_t* __restrict__ offsets, int64_t n, int F,
                                 int64_t V, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t bi, fi;
    split_idx(i, F, bi, fi);
    int64_t row = (int64_t)x[i] + offsets[fi];
    keys[i] = (row >= 0 && row < V) ? (uint32_t)row : (uint32_t)V;   // out-of-range rows sort last and are ignored
    vals[i] = (int32_t)i;
  }
}

__global__ void plan_segments_kernel(const uint32_t* __restrict__ uniq, const int32_t* __restrict__ cnt,
                                     const int32_t* __restrict__ nuniq, int64_t V, int32_t* __restrict__ seg_of_row,
                                     int32_t* __restrict__ long_slot, int32_t* __restrict__ nlong,
                                     int32_t* __restrict__ long_seg, int32_t* __restrict__ long_off) {
  const int nu = *nuniq;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += gridDim.x * blockDim.x) {
    const uint32_t r = uniq[i];
    if ((int64_t)r < V) seg_of_row[r] = i;
    int slot = -1;
    if (cnt[i] > kShortSeg && (int64_t)r < V) {
      // the slot NUMBER depends on atomic order, the sums stored there do not
      const int np = (cnt[i] + kPiece - 1) / kPiece;
      slot = atomicAdd(nlong, np);
      for (int j = 0; j < np; ++j) { long_seg[slot + j] = i; long_off[slot + j] = j * kPiece; }
    }
    long_slot[i] = slot;
  }
}

// gradient rows arrive as fp32 (the model program's own input gradient) or as bf16 (the replicas' row-gradient exchange, widened here
// instead of by a separate cast pass over the whole inbox): four consecutive elements of either as a float4
template <typename GT> __device__ __forceinline__ float4 ld_grad4(const GT* p);
template <> __device__ __forceinline__ float4 ld_grad4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld_grad4<uint16_t>(const uint16_t* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
template <typename GT> __device__ __forceinline__ float ld_grad1(const GT* p);
template <> __device__ __forceinline__ float ld_grad1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_grad1<uint16_t>(const uint16_t* p) { return bf16_to_f32(*p); }

// one CTA per piece of a long segment: lane-group j sums entries j, j+G, j+2G, ... in order; fixed smem tree afterwards
template <int VEC, typename GT>
__global__ void __launch_bounds__(kLongThreads)
long_segment_kernel(const GT* __restrict__ grad_out, int64_t ldg, int F, int E, const int32_t* __restrict__ vals,
                    const int32_t* __restrict__ start, const int32_t* __restrict__ cnt, const int32_t* __restrict__ long_seg,
                    const int32_t* __restrict__ long_off, const int32_t* __restrict__ nlong, float* __restrict__ long_sum) {
  extern __shared__ float sm[];                    // [groups][E]
  const int lanes = E / VEC;
  const int groups = kLongThreads / lanes;
  const int g = threadIdx.x / lanes, q = threadIdx.x % lanes;
  for (int ls = blockIdx.x; ls < *nlong; ls += gridDim.x) {
    const int seg = long_seg[ls], off = long_off[ls];
    const int s0 = start[seg] + off, c = min(kPiece, cnt[seg] - off);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (g < groups) {
      for (int i = g; i < c; i += groups) {
        const int pos = vals[s0 + i];
        const GT* src = grad_out + (int64_t)(pos / F) * ldg + (int64_t)(pos % F) * E + q * VEC;
        if (VEC == 4) { const float4 t = ld_grad4<GT>(src); acc[0] += t.x; acc[1 % VEC] += t.y; acc[2 % VEC] += t.z; acc[3 % VEC] += t.w; }
        else acc[0] += ld_grad1<GT>(src);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) sm[g * E + q * VEC + j] = acc[j];
    }
    __syncthreads();
    for (int half = 1; half < groups; half <<= 1) {      // groups is a power of two for E in {4,8,16,32,64,...}; generic otherwise
      if (g < groups && (g % (2 * half)) == 0 && g + half < groups) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) sm[g * E + q * VEC + j] += sm[(g + half) * E + q * VEC + j];
      }
      __syncthreads();
    }
    if (g == 0) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) long_sum[(int64_t)ls * E + q * VEC + j] = sm[q * VEC + j];
    }
    __syncthreads();
  }
}

template <int VEC, typename GT>
__device__ __forceinline__ void segment_sum(float (&acc)[VEC], const GT* __restrict__ grad_out, int64_t ldg, int F, int E,
                                            int q, int seg, const int32_t* __restrict__ vals, const int32_t* __restrict__ start,
                                            const int32_t* __restrict__ cnt, const int32_t* __restrict__ long_slot,
                                            const float* __restrict__ long_sum) {
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (seg < 0) return;
  const int c = cnt[seg];
  if (c > kShortSeg) {                                   // pieces of the segment, in order
    const float* src = long_sum + (int64_t)long_slot[seg] * E + q * VEC;
    const int np = (c + kPiece - 1) / kPiece;
    for (int i = 0; i < np; ++i, src += E) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += src[j];
    }
    return;
  }
  const int s0 = start[seg];
  if (VEC == 4) {
    // four entries per round: the four index loads, then the four row loads, are independent of each other - a segment of <= 4
    // entries (most of them) costs three dependent round trips instead of two per entry.  The adds keep the entry order.
    for (int i = 0; i < c; i += 4) {
      int pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pos[u] = (i + u < c) ? __ldg(vals + s0 + i + u) : -1;
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos[u] >= 0) {
          const uint32_t b = (uint32_t)pos[u] / (uint32_t)F, f = (uint32_t)pos[u] - b * (uint32_t)F;
          t[u] = ld_grad4<GT>(grad_out + (int64_t)b * ldg + (int64_t)f * E + q * VEC);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (pos[u] >= 0) { acc[0] += t[u].x; acc[1 % VEC] += t[u].y; acc[2 % VEC] += t[u].z; acc[3 % VEC] += t[u].w; }
    }
    return;
  }
  for (int i = 0; i < c; ++i) {
    const int pos = vals[s0 + i];
    const GT* src = grad_out + (int64_t)(pos / F) * ldg + (int64_t)(pos % F) * E + q * VEC;
    acc[0] += ld_grad1<GT>(src);
  }
}

template <int VEC>
__global__ void embed_dense_grad_kernel(const float* __restrict__ grad_out, int64_t ldg, int F, int E, int64_t V,
                                        const int32_t* __restrict__ seg_of_row, const int32_t* __restrict__ vals,
                                        const int32_t* __restrict__ start, const int32_t* __restrict__ cnt,
                                        const int32_t* __restrict__ long_slot, const float* __restrict__ long_sum,
                                        float* __restrict__ grad_table) {
  const int lanes = E / VEC;
  const int64_t total = V * lanes;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, q64;
    split_idx(i, lanes, r, q64);
    const int q = (int)q64;
    float acc[VEC];
    segment_sum<VEC, float>(acc, grad_out, ldg, F, E, q, seg_of_row[r], vals, start, cnt, long_slot, long_sum);
    float* o = grad_table + r * E + q * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else o[0] = acc[0];
  }
}

struct AdamK { float lr_t, b1, b2, eps, wd, l2x2, bc2_sqrt; };
__device__ __forceinline__ AdamK load_adam(const cdcmdr_step_state_t* st, float l2) {
  AdamK k; k.lr_t = st->lr_t; k.b1 = st->beta1; k.b2 = st->beta2; k.eps = st->eps; k.wd = st->weight_decay;
  k.l2x2 = 2.f * l2; k.bc2_sqrt = st->bc2_sqrt; return k;
}

__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, const AdamK& k) {
  g = g + k.l2x2 * w;                    // gradient of l2 * sum(w^2)                 layer.py:106-108
  g = g + k.wd * w;                      // Adam weight_decay (L2 style)              run.py:720
  m = m + (1.f - k.b1) * (g - m);        // exp_avg.lerp_(grad, 1-beta1)
  v = k.b2 * v + (1.f - k.b2) * g * g;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
  w = w - k.lr_t * (m / denom);          // param.addcdiv_(exp_avg, denom, value=-lr/bc1)
}

template <int VEC, bool DENSE>
__global__ void __launch_bounds__(256, 4)
embed_adam_kernel(const float* __restrict__ grad_out, int64_t ldg, int F, int E, int64_t V,
                  const int32_t* __restrict__ seg_of_row, const uint32_t* __restrict__ uniq, const int32_t* __restrict__ nuniq,
                  const int32_t* __restrict__ vals, const int32_t* __restrict__ start, const int32_t* __restrict__ cnt,
                  const int32_t* __restrict__ long_slot, const float* __restrict__ long_sum,
                  float* __restrict__ table, float* __restrict__ mom, float* __restrict__ var,
                  const cdcmdr_step_state_t* __restrict__ st, float l2, double* __restrict__ reg_partials) {
  const AdamK k = load_adam(st, l2);
  const int lanes = E / VEC;
  const int64_t rows = DENSE ? V : (int64_t)(*nuniq);
  const int64_t total = rows * lanes;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double sq = 0.0, sqn = 0.0;                      // sum of squares of the rows before / (lazy mode) after the update
  if (VEC == 4) {
    // The sweep is bound by memory latency (ncu: long-scoreboard stalls, 29 % of DRAM bandwidth at one row per thread in flight):
    // every thread keeps TWO independent rows in flight - all six 128-bit loads are issued before either row is touched.
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
      int64_t r[2], o[2]; int q[2], seg[2]; bool ok[2];
      float4 w[2], m[2], v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * stride;
        ok[u] = i < total;
        int64_t q64 = 0; r[u] = 0;
        if (ok[u]) split_idx(i, lanes, r[u], q64);
        q[u] = (int)q64;
        seg[u] = -1;
        if (ok[u]) {
          if (DENSE) seg[u] = seg_of_row[r[u]];
          else { seg[u] = (int)r[u]; r[u] = uniq[seg[u]]; ok[u] = r[u] < V; }
        }
        o[u] = r[u] * E + q[u] * VEC;
        if (ok[u]) {
          w[u] = *reinterpret_cast<const float4*>(table + o[u]); m[u] = *reinterpret_cast<const float4*>(mom + o[u]);
          v[u] = *reinterpret_cast<const float4*>(var + o[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
        float acc[VEC];
        segment_sum<VEC, float>(acc, grad_out, ldg, F, E, q[u], seg[u], vals, start, cnt, long_slot, long_sum);
        sq += (double)w[u].x * w[u].x + (double)w[u].y * w[u].y + (double)w[u].z * w[u].z + (double)w[u].w * w[u].w;
        adam_elem(w[u].x, m[u].x, v[u].x, acc[0], k); adam_elem(w[u].y, m[u].y, v[u].y, acc[1 % VEC], k);
        adam_elem(w[u].z, m[u].z, v[u].z, acc[2 % VEC], k); adam_elem(w[u].w, m[u].w, v[u].w, acc[3 % VEC], k);
        if (!DENSE) sqn += (double)w[u].x * w[u].x + (double)w[u].y * w[u].y + (double)w[u].z * w[u].z + (double)w[u].w * w[u].w;
        *reinterpret_cast<float4*>(table + o[u]) = w[u]; *reinterpret_cast<float4*>(mom + o[u]) = m[u];
        *reinterpret_cast<float4*>(var + o[u]) = v[u];
      }
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
      int64_t r, q64; int seg;
      split_idx(i, lanes, r, q64);
      const int q = (int)q64;
      if (DENSE) { seg = seg_of_row[r]; }
      else { seg = (int)r; r = uniq[seg]; if (r >= V) continue; }
      float acc[VEC];
      segment_sum<VEC, float>(acc, grad_out, ldg, F, E, q, seg, vals, start, cnt, long_slot, long_sum);
      const int64_t o = r * E + q * VEC;
      float w = table[o], m = mom[o], v = var[o];
      sq += (double)w * w;
      adam_elem(w, m, v, acc[0], k);
      if (!DENSE) sqn += (double)w * w;
      table[o] = w; mom[o] = m; var[o] = v;
    }
  }
  if (reg_partials) {                              // deterministic: fixed grid, fixed in-block tree
    __shared__ double red[8];
    if (!DENSE) sq = sqn - sq;                     // lazy mode: the CHANGE of the touched rows' sum of squares
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; reg_partials[blockIdx.x] = t; }
  }
}

// Dense-exact update in two passes.  Pass 1 (this kernel): one thread per (touched row, 4 columns) sums the row's segment into a
// compact [n_unique, E] buffer - every thread has work, the dependent vals -> grad_out loads of different segments overlap.
// Pass 2 (embed_adam_dense4_kernel) is then a pure stream over table / m / v: the only dependent load is seg_of_row -> gsum row,
// and consecutive touched rows read consecutive gsum rows.  The fused single pass interleaved the segment walks with the sweep:
// ncu showed long-scoreboard stalls at 29 % of DRAM bandwidth (every warp waited for its slowest segment).
template <int VEC, typename GT>
__global__ void __launch_bounds__(256)
embed_segsum_kernel(const GT* __restrict__ grad_out, int64_t ldg, int F, int E, const int32_t* __restrict__ nuniq,
                    const int32_t* __restrict__ vals, const int32_t* __restrict__ start, const int32_t* __restrict__ cnt,
                    const int32_t* __restrict__ long_slot, const float* __restrict__ long_sum, float* __restrict__ gsum) {
  const int lanes = E / VEC;
  const int64_t total = (int64_t)(*nuniq) * lanes;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t seg, q64;
    split_idx(i, lanes, seg, q64);
    const int q = (int)q64;
    float acc[VEC];
    segment_sum<VEC, GT>(acc, grad_out, ldg, F, E, q, (int)seg, vals, start, cnt, long_slot, long_sum);
    float* o = gsum + seg * E + q * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else o[0] = acc[0];
  }
}

__global__ void __launch_bounds__(256, 4)
embed_adam_dense4_kernel(int E, int64_t V, const int32_t* __restrict__ seg_of_row, const float* __restrict__ gsum,
                         float* __restrict__ table, float* __restrict__ mom, float* __restrict__ var,
                         const cdcmdr_step_state_t* __restrict__ st, float l2, double* __restrict__ reg_partials) {
  const AdamK k = load_adam(st, l2);
  const int lanes = E / 4;
  const int64_t total = V * lanes;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double sq = 0.0;
#pragma unroll 2
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t r, q64;
    split_idx(i, lanes, r, q64);
    const int64_t o = r * E + q64 * 4;
    const int seg = __ldg(seg_of_row + r);
    float4 w = *reinterpret_cast<const float4*>(table + o), m = *reinterpret_cast<const float4*>(mom + o);
    float4 v = *reinterpret_cast<const float4*>(var + o);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (seg >= 0) g = __ldg(reinterpret_cast<const float4*>(gsum + (int64_t)seg * E + q64 * 4));
    sq += (double)w.x * w.x + (double)w.y * w.y + (double)w.z * w.z + (double)w.w * w.w;
    adam_elem(w.x, m.x, v.x, g.x, k); adam_elem(w.y, m.y, v.y, g.y, k);
    adam_elem(w.z, m.z, v.z, g.z, k); adam_elem(w.w, m.w, v.w, g.w, k);
    *reinterpret_cast<float4*>(table + o) = w; *reinterpret_cast<float4*>(mom + o) = m; *reinterpret_cast<float4*>(var + o) = v;
  }
  if (reg_partials) {                              // deterministic: fixed grid, fixed in-block tree
    __shared__ double red[8];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; reg_partials[blockIdx.x] = t; }
  }
}

__global__ void reg_finalize_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double t = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += partials[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w]; *out = s; }
}

// lazy regulariser: *before = running total (the value the step's loss uses), then running += sum of the partial deltas
__global__ void reg_running_kernel(const double* __restrict__ partials, int n, double* __restrict__ running, double* __restrict__ before) {
  __shared__ double red[32];
  double t = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += partials[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    const double r = *running;
    if (before) *before = r;
    *running = r + s;
  }
}

static int grid_for(int64_t work, int threads, int max_ctas_per_sm = 8) {
  int64_t g = ceil_div(work, threads);
  const int64_t cap = (int64_t)kNumSMs * max_ctas_per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" int cdcmdr_embed_gather_fwd(const int32_t* x, const int64_t* offsets, const float* table, float* out_f32,
                                       uint16_t* out_bf16, int64_t ld_bf16, int64_t B, int F, int E, int64_t V,
                                       int* oob_flag, cdcmdr_stream_t s) {
  CDC_REQUIRE(B >= 0 && F > 0 && E > 0 && V > 0, "bad gather shape");
  if (B == 0) return 0;
  CDC_REQUIRE(out_f32 || out_bf16, "gather needs an output");
  const bool rows_ok = (E == 16 || E == 32 || E == 64) && ((uintptr_t)table % 16) == 0 && (!out_f32 || ((uintptr_t)out_f32 % 16) == 0) &&
                       (!out_bf16 || (ld_bf16 % 8 == 0 && ((uintptr_t)out_bf16 % 16) == 0));
  if (rows_ok) {
    const int grid = grid_for(B * F, 256);
    if (E == 16) embed_gather_rows_kernel<4><<<grid, 256, 0, to_stream(s)>>>(x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, V, oob_flag);
    else if (E == 32) embed_gather_rows_kernel<8><<<grid, 256, 0, to_stream(s)>>>(x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, V, oob_flag);
    else embed_gather_rows_kernel<16><<<grid, 256, 0, to_stream(s)>>>(x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, V, oob_flag);
  } else if (E % 4 == 0 && (!out_bf16 || ld_bf16 % 4 == 0)) {
    const int64_t work = B * F * (E / 4);
    embed_gather_kernel<4, 1><<<grid_for(work, 256), 256, 0, to_stream(s)>>>(x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, E, V, oob_flag);
  } else {
    const int64_t work = B * F * (int64_t)E;
    embed_gather_kernel<1, 1><<<grid_for(work, 256), 256, 0, to_stream(s)>>>(x, offsets, table, out_f32, out_bf16, ld_bf16, B, F, E, V, oob_flag);
  }
  CDC_LAUNCHED();
  return 0;
}

extern "C" size_t cdcmdr_embed_plan_bytes(int64_t n_idx, int64_t V, int E_max) {
  if (n_idx <= 0) n_idx = 1;
  return make_layout(n_idx, V, E_max).total;
}

extern "C" int cdcmdr_embed_plan_build(const int32_t* x, const int64_t* offsets, int64_t B, int F, int64_t V, int E_max,
                                       void* plan, size_t plan_bytes, cdcmdr_stream_t s) {
  const int64_t n = B * F;
  CDC_REQUIRE(n > 0 && n < (int64_t(1) << 31) && V < (int64_t(1) << 31) - 1, "plan size out of int32 range");
  EmbedPlan L = make_layout(n, V, E_max);
  CDC_REQUIRE(plan_bytes >= L.total, "plan workspace too small");
  cudaStream_t st = to_stream(s);
  uint32_t* keys_in = at<uint32_t>(plan, L.off_keys_in); int32_t* vals_in = at<int32_t>(plan, L.off_vals_in);
  uint32_t* keys = at<uint32_t>(plan, L.off_keys); int32_t* vals = at<int32_t>(plan, L.off_vals);
  uint32_t* uniq = at<uint32_t>(plan, L.off_uniq); int32_t* cnt = at<int32_t>(plan, L.off_cnt);
  int32_t* start = at<int32_t>(plan, L.off_start); int32_t* nuniq = at<int32_t>(plan, L.off_nuniq);
  int32_t* seg_of_row = at<int32_t>(plan, L.off_seg_of_row);
  plan_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, offsets, n, F, V, keys_in, vals_in);
  CDC_LAUNCHED();
  size_t tb = L.cub_bytes;
  void* tmp = at<void>(plan, L.off_cub);
  CDC_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys, vals_in, vals, (int)n, 0, sort_bits(V), st));
  tb = L.cub_bytes;
  CDC_CHECK(cudaMemsetAsync(cnt, 0, n * 4, st));
  CDC_CHECK(cub::DeviceRunLengthEncode::Encode(tmp, tb, keys, uniq, cnt, nuniq, (int)n, st));
  tb = L.cub_bytes;
  CDC_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, start, (int)n, st));
  g_launches.fetch_add(4, std::memory_order_relaxed);   // CUB: sort passes + RLE + scan (lower bound)
  CDC_CHECK(cudaMemsetAsync(seg_of_row, 0xFF, (size_t)V * 4, st));
  CDC_CHECK(cudaMemsetAsync(at<void>(plan, L.off_nlong), 0, 16, st));
  plan_segments_kernel<<<grid_for(n, 256), 256, 0, st>>>(uniq, cnt, nuniq, V, seg_of_row, at<int32_t>(plan, L.off_long_slot),
                                                         at<int32_t>(plan, L.off_nlong), at<int32_t>(plan, L.off_long_seg),
                                                         at<int32_t>(plan, L.off_long_off));
  CDC_LAUNCHED();
  return 0;
}

namespace cdcmdr {
// host copy of the layout is recomputed from (B*F, V, E_max) - the plan header on the device is informational
template <typename GT>
static int launch_long(const GT* grad_out, int64_t ldg, const void* plan, const EmbedPlan& L, int F, int E, cudaStream_t st) {
  const int lanes = (E % 4 == 0) ? E / 4 : E;
  CDC_REQUIRE(lanes <= kLongThreads, "embed_dim too large for the long-segment kernel");
  const size_t smem = (size_t)(kLongThreads / lanes) * E * sizeof(float);
  CDC_REQUIRE(smem <= 48 * 1024, "embed_dim too large for the long-segment kernel");
  const int grid = (int)(L.n_long_max < 8 * kNumSMs ? L.n_long_max : 8 * kNumSMs);
  if (E % 4 == 0)
    long_segment_kernel<4, GT><<<grid, kLongThreads, smem, st>>>(grad_out, ldg, F, E, at<int32_t>(plan, L.off_vals), at<int32_t>(plan, L.off_start),
        at<int32_t>(plan, L.off_cnt), at<int32_t>(plan, L.off_long_seg), at<int32_t>(plan, L.off_long_off), at<int32_t>(plan, L.off_nlong),
        at<float>(plan, L.off_long_sum));
  else
    long_segment_kernel<1, GT><<<grid, kLongThreads, smem, st>>>(grad_out, ldg, F, E, at<int32_t>(plan, L.off_vals), at<int32_t>(plan, L.off_start),
        at<int32_t>(plan, L.off_cnt), at<int32_t>(plan, L.off_long_seg), at<int32_t>(plan, L.off_long_off), at<int32_t>(plan, L.off_nlong),
        at<float>(plan, L.off_long_sum));
  CDC_LAUNCHED();
  return 0;
}
}  // namespace cdcmdr

extern "C" int cdcmdr_embed_bwd_dense(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F, int E,
                                      int64_t V, float* grad_table, cdcmdr_stream_t s) {
  CDC_REQUIRE(E <= E_max, "E exceeds the plan's E_max");
  EmbedPlan L = make_layout(B * F, V, E_max);
  cudaStream_t st = to_stream(s);
  if (int rc = launch_long(grad_out, ldg, plan, L, F, E, st)) return rc;
  const int lanes = (E % 4 == 0) ? E / 4 : E;
  const int grid = grid_for(V * lanes, 256);
#define ARGS grad_out, ldg, F, E, V, at<int32_t>(plan, L.off_seg_of_row), at<int32_t>(plan, L.off_vals), at<int32_t>(plan, L.off_start), \
             at<int32_t>(plan, L.off_cnt), at<int32_t>(plan, L.off_long_slot), at<float>(plan, L.off_long_sum), grad_table
  if (E % 4 == 0) embed_dense_grad_kernel<4><<<grid, 256, 0, st>>>(ARGS);
  else embed_dense_grad_kernel<1><<<grid, 256, 0, st>>>(ARGS);
#undef ARGS
  CDC_LAUNCHED();
  return 0;
}

template <typename GT>
static int embed_adam_impl(bool dense, const GT* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F, int E,
                           int64_t V, float* table, float* m, float* v, float l2, const cdcmdr_step_state_t* h, double* reg_sumsq,
                           cdcmdr_stream_t s, double* reg_running = nullptr) {
  CDC_REQUIRE(E <= E_max, "E exceeds the plan's E_max");
  CDC_REQUIRE(h, "Adam needs the device step state");
  EmbedPlan L = make_layout(B * F, V, E_max);
  cudaStream_t st = to_stream(s);
  if (int rc = launch_long(grad_out, ldg, plan, L, F, E, st)) return rc;
  const int lanes = (E % 4 == 0) ? E / 4 : E;
  const int64_t rows = dense ? V : B * F;
  int grid = grid_for(rows * lanes, 256);
  if (grid > kRegPartials) grid = kRegPartials;
  double* partials = ((reg_sumsq && dense) || (reg_running && !dense)) ? at<double>(plan, L.off_reg) : nullptr;
#define ARGS grad_out, ldg, F, E, V, at<int32_t>(plan, L.off_seg_of_row), at<uint32_t>(plan, L.off_uniq), at<int32_t>(plan, L.off_nuniq), \
             at<int32_t>(plan, L.off_vals), at<int32_t>(plan, L.off_start), at<int32_t>(plan, L.off_cnt),                                  \
             at<int32_t>(plan, L.off_long_slot), at<float>(plan, L.off_long_sum), table, m, v, h, l2, partials
  if (dense && E % 4 == 0) {
    float* gsum = at<float>(plan, L.off_gsum);
    embed_segsum_kernel<4, GT><<<grid_for(B * F * lanes, 256), 256, 0, st>>>(grad_out, ldg, F, E, at<int32_t>(plan, L.off_nuniq),
        at<int32_t>(plan, L.off_vals), at<int32_t>(plan, L.off_start), at<int32_t>(plan, L.off_cnt), at<int32_t>(plan, L.off_long_slot),
        at<float>(plan, L.off_long_sum), gsum);
    CDC_LAUNCHED();
    embed_adam_dense4_kernel<<<grid, 256, 0, st>>>(E, V, at<int32_t>(plan, L.off_seg_of_row), gsum, table, m, v, h, l2, partials);
  } else if constexpr (std::is_same<GT, float>::value) {
    if (dense) embed_adam_kernel<1, true><<<grid, 256, 0, st>>>(ARGS);
    else if (E % 4 == 0) embed_adam_kernel<4, false><<<grid, 256, 0, st>>>(ARGS);
    else embed_adam_kernel<1, false><<<grid, 256, 0, st>>>(ARGS);
  } else {
    CDC_REQUIRE(false, "bf16 row gradients: only the dense-exact update with embed_dim % 4 == 0 reads them directly");
  }
#undef ARGS
  CDC_LAUNCHED();
  if (partials && !dense) {
    reg_running_kernel<<<1, 256, 0, st>>>(partials, grid, reg_running, reg_sumsq);
    CDC_LAUNCHED();
  } else if (partials) {
    reg_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, reg_sumsq);
    CDC_LAUNCHED();
  }
  return 0;
}

extern "C" int cdcmdr_embed_bwd_adam_dense_exact(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                                 int E, int64_t V, float* table, float* m, float* v, float l2, const cdcmdr_step_state_t* h,
                                                 double* reg_sumsq, cdcmdr_stream_t s) {
  return embed_adam_impl(true, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, h, reg_sumsq, s);
}
// the same update from bf16 row gradients (what the replicas' gradient exchange delivers): the segment sums widen them on the fly
extern "C" int cdcmdr_embed_bwd_adam_dense_exact_g16(const uint16_t* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                                     int E, int64_t V, float* table, float* m, float* v, float l2,
                                                     const cdcmdr_step_state_t* h, double* reg_sumsq, cdcmdr_stream_t s) {
  CDC_REQUIRE(E % 4 == 0 && ldg % 4 == 0 && ((uintptr_t)grad_out % 8) == 0, "bf16 row gradients: embed_dim and pitch must be multiples of 4");
  return embed_adam_impl<uint16_t>(true, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, h, reg_sumsq, s);
}
extern "C" int cdcmdr_embed_bwd_adam_sparse_lazy(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                                 int E, int64_t V, float* table, float* m, float* v, float l2, const cdcmdr_step_state_t* h,
                                                 cdcmdr_stream_t s) {
  return embed_adam_impl(false, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, h, nullptr, s);
}

// sparse_lazy with an incrementally maintained regulariser: a full-table sum of squares per step is 16 GB of reads per GPU at the
// 500 M x 64 scale.  reg_running (device double) holds sum(w^2) over the table (the caller initialises it once with
// cdcmdr_reg_l2_sum); this call writes its value BEFORE the update to reg_before (what the step's loss uses, may be NULL) and adds
// the change of the touched rows.
extern "C" int cdcmdr_embed_bwd_adam_sparse_lazy_reg(const float* grad_out, int64_t ldg, const void* plan, int E_max, int64_t B, int F,
                                                     int E, int64_t V, float* table, float* m, float* v, float l2,
                                                     const cdcmdr_step_state_t* h, double* reg_running, double* reg_before,
                                                     cdcmdr_stream_t s) {
  CDC_REQUIRE(reg_running, "lazy regulariser: running total missing");
  return embed_adam_impl(false, grad_out, ldg, plan, E_max, B, F, E, V, table, m, v, l2, h, reg_before, s, reg_running);
}

extern "C" int cdcmdr_embed_gather_peer(const int32_t* x, const int64_t* offsets, const float* const* shards, int64_t rows_per,
                                        float* out_f32, uint16_t* out_bf16, int64_t ld_bf16, int64_t B, int F, int E, int64_t V,
                                        int* oob_flag, cdcmdr_stream_t s) {
  CDC_REQUIRE(B >= 0 && F > 0 && E > 0 && V > 0 && rows_per > 0 && shards, "bad peer gather arguments");
  if (B == 0) return 0;
  CDC_REQUIRE(out_f32 || out_bf16, "gather needs an output");
  CDC_REQUIRE(E % 4 == 0 && (E == 4 || E == 8 || E == 16 || E == 32 || E == 64 || E == 128), "peer gather: embed_dim must be 4..128, a power of two");
  CDC_REQUIRE((!out_f32 || ((uintptr_t)out_f32 % 16) == 0) && (!out_bf16 || (ld_bf16 % 4 == 0 && ((uintptr_t)out_bf16 % 8) == 0)),
              "peer gather: misaligned output");
  const int lanes = E / 4;
  const int grid = grid_for(B * F * lanes, 256, 16);
  cudaStream_t st = to_stream(s);
#define PG(L) embed_gather_peer_kernel<L><<<grid, 256, 0, st>>>(x, offsets, shards, rows_per, out_f32, out_bf16, ld_bf16, B, F, E, V, oob_flag)
  switch (lanes) {
    case 1: PG(1); break; case 2: PG(2); break; case 4: PG(4); break; case 8: PG(8); break; case 16: PG(16); break; default: PG(32); break;
  }
#undef PG
  CDC_LAUNCHED();
  return 0;
}
