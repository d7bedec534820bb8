// Cross-network elementwise stages (a10/a11/a12), small row softmax / tanh, and per-domain routing (a14/a15):
// stable partition of rows by group id + row permutation.  Integer work is bit-exact by construction.
#include "common.cuh"

namespace cdcmdr {

static int grid_1d(int64_t work, int threads, int max_per_sm = 8) {
  int64_t g = ceil_div(work, threads);
  const int64_t cap = (int64_t)kNumSMs * max_per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

// ------------------------------------------------------------------------------------------ cross v1 / v2
__global__ void cross_fuse_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x, const float* __restrict__ xw, int xw_cols,
                                      const float* __restrict__ b, float* __restrict__ out, int64_t B, int64_t D) {
  const int64_t total = B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, d; split_idx(i, D, r, d);
    const float w = xw_cols == 1 ? xw[r] : xw[i];
    out[i] = x0[i] * w + b[d] + x[i];
  }
}

// v2: dxw[b,d] = dout*x0 ; dx0_acc += dout*xw
__global__ void cross_fuse_bwd_v2_kernel(const float* __restrict__ x0, const float* __restrict__ xw, const float* __restrict__ dout,
                                         float* __restrict__ dx0_acc, float* __restrict__ dxw, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = dout[i];
    dxw[i] = g * x0[i];
    dx0_acc[i] += g * xw[i];
  }
}

// v1: one warp per row: dxw[b] = sum_d dout*x0 ; dx0_acc += dout*xw[b]
__global__ void __launch_bounds__(256)
cross_fuse_bwd_v1_kernel(const float* __restrict__ x0, const float* __restrict__ xw, const float* __restrict__ dout,
                         float* __restrict__ dx0_acc, float* __restrict__ dxw, int64_t B, int64_t D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < B; r += (int64_t)gridDim.x * 8) {
    const float w = xw[r];
    float acc = 0.f;
    for (int64_t d = lane; d < D; d += 32) {
      const float g = dout[r * D + d];
      acc = fmaf(g, x0[r * D + d], acc);
      dx0_acc[r * D + d] += g * w;
    }
    acc = warp_sum(acc);
    if (lane == 0) dxw[r] = acc;
  }
}

// ------------------------------------------------------------------------------------------ cross mix
__global__ void crossmix_combine_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x, const float* __restrict__ u,
                                            const float* __restrict__ g, const float* __restrict__ bias, float* __restrict__ out,
                                            int64_t B, int64_t D, int n_exp) {
  const int64_t total = B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, d; split_idx(i, D, r, d);
    const float bd = bias[d], x0v = x0[i];
    float acc = 0.f;
    for (int e = 0; e < n_exp; ++e) acc = fmaf(g[r * n_exp + e], x0v * (u[(int64_t)e * total + i] + bd), acc);
    out[i] = x[i] + acc;
  }
}

// one warp per row; dgate needs a row reduction
__global__ void __launch_bounds__(256)
crossmix_combine_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ u, const float* __restrict__ g,
                            const float* __restrict__ bias, const float* __restrict__ dout, float* __restrict__ du,
                            float* __restrict__ dgate, float* __restrict__ dx0_acc, int64_t B, int64_t D, int n_exp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = B * D;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < B; r += (int64_t)gridDim.x * 8) {
    for (int e = 0; e < n_exp; ++e) {
      const float ge = g[r * n_exp + e];
      float acc = 0.f;
      for (int64_t d = lane; d < D; d += 32) {
        const int64_t i = r * D + d;
        const float go = dout[i], x0v = x0[i], ub = u[(int64_t)e * total + i] + bias[d];
        du[(int64_t)e * total + i] = go * ge * x0v;
        acc = fmaf(go * x0v, ub, acc);
        dx0_acc[i] += go * ge * ub;
      }
      acc = warp_sum(acc);
      if (lane == 0) dgate[r * n_exp + e] = acc;
    }
  }
}

__global__ void tanh_fwd_kernel(float* __restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ y, float* __restrict__ dy, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = y[i];
    dy[i] = dy[i] * (1.f - t * t);
  }
}

__global__ void softmax_rows_fwd_kernel(const float* __restrict__ z, int64_t ldz, float* __restrict__ p, int64_t ldp, int64_t B, int n) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    const float* zr = z + r * ldz;
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j) mx = fmaxf(mx, zr[j]);
    float sum = 0.f;
    for (int j = 0; j < n; ++j) sum += expf(zr[j] - mx);
    for (int j = 0; j < n; ++j) p[r * ldp + j] = expf(zr[j] - mx) / sum;
  }
}
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ p, int64_t ldp, const float* __restrict__ dp, int64_t lddp,
                                        float* __restrict__ dz, int64_t lddz, int64_t B, int n) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < B; r += (int64_t)gridDim.x * blockDim.x) {
    float dot = 0.f;
    for (int j = 0; j < n; ++j) dot = fmaf(p[r * ldp + j], dp[r * lddp + j], dot);
    for (int j = 0; j < n; ++j) dz[r * lddz + j] = p[r * ldp + j] * (dp[r * lddp + j] - dot);
  }
}

// ------------------------------------------------------------------------------------------ routing
constexpr int kRouteChunk = 1024;    // rows per block = 32 warp slots of 32 rows
constexpr int kRouteMaxGroups = 256;

// stage 1: per-chunk histogram  hist[chunk][g]
__global__ void __launch_bounds__(256)
route_hist_kernel(const int64_t* __restrict__ group, int64_t B, int n_group, int32_t* __restrict__ hist) {
  __shared__ int cnt[kRouteMaxGroups];
  for (int i = threadIdx.x; i < n_group; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * kRouteChunk;
  for (int i = threadIdx.x; i < kRouteChunk; i += blockDim.x) {
    const int64_t r = r0 + i;
    if (r < B) { const int64_t g = group[r]; if (g >= 0 && g < n_group) atomicAdd(&cnt[(int)g], 1); }   // integer atomics: exact
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_group; i += blockDim.x) hist[(int64_t)blockIdx.x * n_group + i] = cnt[i];
}

// stage 2 (one block): counts, group_start, and base[chunk][g] = group_start[g] + sum_{chunk' < chunk} hist[chunk'][g]
__global__ void route_scan_kernel(const int32_t* __restrict__ hist, int n_chunks, int n_group, int32_t* __restrict__ base,
                                  int32_t* __restrict__ counts, int32_t* __restrict__ group_start) {
  __shared__ int tot[kRouteMaxGroups];
  __shared__ int start[kRouteMaxGroups + 1];
  const int g = threadIdx.x;
  if (g < n_group) {
    int run = 0;
    for (int c = 0; c < n_chunks; ++c) { base[(int64_t)c * n_group + g] = run; run += hist[(int64_t)c * n_group + g]; }
    tot[g] = run; counts[g] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < n_group; ++i) { start[i] = run; run += tot[i]; }
    start[n_group] = run;
    for (int i = 0; i <= n_group; ++i) group_start[i] = start[i];
  }
  __syncthreads();
  if (g < n_group) {
    const int s0 = start[g];
    for (int c = 0; c < n_chunks; ++c) base[(int64_t)c * n_group + g] += s0;
  }
}

// stage 3: stable scatter.  Each warp handles 32-row slots; rank inside a slot by match_any, across slots by a
// per-group serial prefix over the chunk's 32 slots.
__global__ void __launch_bounds__(256)
route_scatter_kernel(const int64_t* __restrict__ group, int64_t B, int n_group, const int32_t* __restrict__ base, int32_t* __restrict__ perm) {
  __shared__ int slot_cnt[32][kRouteMaxGroups + 1];   // +1 breaks the 256-stride bank pattern
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * (kRouteMaxGroups + 1); i += blockDim.x) (&slot_cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * kRouteChunk;
  int my_g[4], my_rank[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int slot = it * 8 + warp;
    const int64_t r = r0 + slot * 32 + lane;
    int g = -1;
    if (r < B) { const int64_t gg = group[r]; if (gg >= 0 && gg < n_group) g = (int)gg; }
    const unsigned peers = __match_any_sync(0xffffffffu, g);
    my_g[it] = g;
    my_rank[it] = __popc(peers & ((1u << lane) - 1u));
    if (g >= 0 && my_rank[it] == 0) slot_cnt[slot][g] = __popc(peers);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < n_group; g += blockDim.x) {    // exclusive prefix over the 32 slots
    int run = base[(int64_t)blockIdx.x * n_group + g];
    for (int sl = 0; sl < 32; ++sl) { const int c = slot_cnt[sl][g]; slot_cnt[sl][g] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int slot = it * 8 + warp;
    const int64_t r = r0 + slot * 32 + lane;
    if (my_g[it] >= 0) perm[slot_cnt[slot][my_g[it]] + my_rank[it]] = (int32_t)r;
  }
}

template <typename V>
__global__ void permute_rows_kernel(const V* __restrict__ src, int64_t lds, const int32_t* __restrict__ perm, int64_t n, int64_t vec_cols,
                                    V* __restrict__ dst, int64_t ldd, int scatter) {
  const int64_t total = n * vec_cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, vec_cols, r, c);
    const int64_t p = perm ? (int64_t)perm[r] : r;          // perm == NULL: plain strided 2-D copy
    if (scatter) dst[p * ldd + c] = src[r * lds + c];
    else dst[r * ldd + c] = src[p * lds + c];
  }
}

template <typename V>
__global__ void copy2d_batched_kernel(const V* __restrict__ src, int64_t src_bs, int64_t lds, V* __restrict__ dst, int64_t dst_bs, int64_t ldd,
                                      int64_t batches, int64_t rows, int64_t cols) {
  const int64_t per = rows * cols, total = batches * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, rc, r, c;
    split_idx(i, per, b, rc);
    split_idx(rc, cols, r, c);
    dst[b * dst_bs + r * ldd + c] = src[b * src_bs + r * lds + c];
  }
}

// per-segment BCE means: block (seg, chunk) sums its rows in double, a second kernel adds the chunks of a segment in order
constexpr int kBceChunks = 64;
__global__ void __launch_bounds__(256)
bce_segments_partial_kernel(const float* __restrict__ pred, int64_t ld_pred, const void* __restrict__ target, int target_is_f32,
                            const int64_t* __restrict__ seg_start, double* __restrict__ partial) {
  const int seg = blockIdx.x, chunk = blockIdx.y;
  const int64_t r0 = seg_start[seg], r1 = seg_start[seg + 1];
  const int64_t per = ceil_div(r1 - r0 > 0 ? r1 - r0 : 1, (int64_t)kBceChunks);
  const int64_t c0 = r0 + chunk * per, c1 = (c0 + per < r1) ? c0 + per : r1;
  double acc = 0.0;
  for (int64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
    const float p = pred[i * ld_pred];
    const float t = target_is_f32 ? static_cast<const float*>(target)[i] : (float)static_cast<const int16_t*>(target)[i];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    acc += (double)(-(t * lp + (1.f - t) * l1p));
  }
  __shared__ double red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0;
    for (int w = 0; w < 8; ++w) tsum += red[w];
    partial[(int64_t)seg * kBceChunks + chunk] = tsum;
  }
}
__global__ void bce_segments_finalize_kernel(const double* __restrict__ partial, const int64_t* __restrict__ seg_start, int n_seg,
                                             float* __restrict__ out_mean) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= n_seg) return;
  double t = 0;
  for (int c = 0; c < kBceChunks; ++c) t += partial[(int64_t)seg * kBceChunks + c];
  const int64_t n = seg_start[seg + 1] - seg_start[seg];
  out_mean[seg] = n > 0 ? (float)(t / (double)n) : __int_as_float(0x7fc00000);
}

__global__ void domain_to_group_kernel(const int32_t* __restrict__ x, int64_t B, int F, int domain_idx,
                                       const int64_t* __restrict__ d2g, int n_domain, int64_t* __restrict__ groups) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const int d = x[b * F + domain_idx];
    groups[b] = (d >= 0 && d < n_domain) ? d2g[d] : -1;
  }
}

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" int cdcmdr_cross_fuse_fwd(const float* x0, const float* x, const float* xw, int xw_cols, const float* b, float* out,
                                     int64_t B, int64_t D, cdcmdr_stream_t s) {
  CDC_REQUIRE(xw_cols == 1 || xw_cols == D, "xw must be [B,1] or [B,D]");
  if (B <= 0 || D <= 0) return 0;
  cross_fuse_fwd_kernel<<<grid_1d(B * D, 256), 256, 0, to_stream(s)>>>(x0, x, xw, xw_cols, b, out, B, D);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_cross_fuse_bwd(const float* x0, const float* xw, int xw_cols, const float* dout, float* dx0_acc, float* dxw,
                                     int64_t B, int64_t D, cdcmdr_stream_t s) {
  CDC_REQUIRE(xw_cols == 1 || xw_cols == D, "xw must be [B,1] or [B,D]");
  if (B <= 0 || D <= 0) return 0;
  if (xw_cols == 1 && D != 1) cross_fuse_bwd_v1_kernel<<<grid_1d(B * 32, 256), 256, 0, to_stream(s)>>>(x0, xw, dout, dx0_acc, dxw, B, D);
  else cross_fuse_bwd_v2_kernel<<<grid_1d(B * D, 256), 256, 0, to_stream(s)>>>(x0, xw, dout, dx0_acc, dxw, B * D);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_crossmix_combine_fwd(const float* x0, const float* x, const float* u, const float* g, const float* bias, float* out,
                                           int64_t B, int64_t D, int n_exp, cdcmdr_stream_t s) {
  if (B <= 0 || D <= 0) return 0;
  crossmix_combine_fwd_kernel<<<grid_1d(B * D, 256), 256, 0, to_stream(s)>>>(x0, x, u, g, bias, out, B, D, n_exp);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_crossmix_combine_bwd(const float* x0, const float* u, const float* g, const float* bias, const float* dout, float* du,
                                           float* dgate, float* dx0_acc, int64_t B, int64_t D, int n_exp, cdcmdr_stream_t s) {
  if (B <= 0 || D <= 0) return 0;
  crossmix_combine_bwd_kernel<<<grid_1d(B * 32, 256), 256, 0, to_stream(s)>>>(x0, u, g, bias, dout, du, dgate, dx0_acc, B, D, n_exp);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_tanh_fwd(float* x, int64_t n, cdcmdr_stream_t s) {
  if (n <= 0) return 0;
  tanh_fwd_kernel<<<grid_1d(n, 256), 256, 0, to_stream(s)>>>(x, n);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_tanh_bwd(const float* y, float* dy, int64_t n, cdcmdr_stream_t s) {
  if (n <= 0) return 0;
  tanh_bwd_kernel<<<grid_1d(n, 256), 256, 0, to_stream(s)>>>(y, dy, n);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_softmax_rows_fwd(const float* z, int64_t ldz, float* p, int64_t ldp, int64_t B, int n, cdcmdr_stream_t s) {
  CDC_REQUIRE(n >= 1 && n <= 64, "softmax width must be in [1,64]");
  if (B <= 0) return 0;
  softmax_rows_fwd_kernel<<<grid_1d(B, 256), 256, 0, to_stream(s)>>>(z, ldz, p, ldp, B, n);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_softmax_rows_bwd(const float* p, int64_t ldp, const float* dp, int64_t lddp, float* dz, int64_t lddz, int64_t B, int n,
                                       cdcmdr_stream_t s) {
  CDC_REQUIRE(n >= 1 && n <= 64, "softmax width must be in [1,64]");
  if (B <= 0) return 0;
  softmax_rows_bwd_kernel<<<grid_1d(B, 256), 256, 0, to_stream(s)>>>(p, ldp, dp, lddp, dz, lddz, B, n);
  CDC_LAUNCHED();
  return 0;
}

extern "C" size_t cdcmdr_route_scratch_bytes(int64_t B, int n_group) {
  const int64_t chunks = ceil_div(B > 0 ? B : 1, kRouteChunk);
  return (size_t)(2 * chunks * (n_group > 0 ? n_group : 1)) * sizeof(int32_t) + 256;
}

extern "C" int cdcmdr_route_partition(const int64_t* group, int64_t B, int n_group, int32_t* perm, int32_t* counts, int32_t* group_start,
                                      void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(n_group >= 1 && n_group <= kRouteMaxGroups, "n_group must be in [1,256]");
  CDC_REQUIRE(B >= 0 && B < (int64_t(1) << 31) && perm && counts && group_start && scratch, "bad routing arguments");
  cudaStream_t st = to_stream(s);
  const int n_chunks = (int)ceil_div(B > 0 ? B : 1, kRouteChunk);
  int32_t* hist = (int32_t*)scratch;
  int32_t* base = hist + (int64_t)n_chunks * n_group;
  route_hist_kernel<<<n_chunks, 256, 0, st>>>(group, B, n_group, hist);
  CDC_LAUNCHED();
  route_scan_kernel<<<1, kRouteMaxGroups, 0, st>>>(hist, n_chunks, n_group, base, counts, group_start);
  CDC_LAUNCHED();
  if (B > 0) {
    route_scatter_kernel<<<n_chunks, 256, 0, st>>>(group, B, n_group, base, perm);
    CDC_LAUNCHED();
  }
  return 0;
}

extern "C" int cdcmdr_permute_rows(const void* src, int64_t lds, const int32_t* perm, int64_t n, int64_t cols, int elt_bytes, void* dst,
                                   int64_t ldd, int scatter, cdcmdr_stream_t s) {
  CDC_REQUIRE(elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8, "element size must be 2, 4 or 8 bytes");
  if (n <= 0 || cols <= 0) return 0;
  cudaStream_t st = to_stream(s);
  const int64_t row_bytes = cols * elt_bytes, lds_b = lds * elt_bytes, ldd_b = ldd * elt_bytes;
  if (row_bytes % 16 == 0 && lds_b % 16 == 0 && ldd_b % 16 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    permute_rows_kernel<uint4><<<grid_1d(n * (row_bytes / 16), 256), 256, 0, st>>>((const uint4*)src, lds_b / 16, perm, n, row_bytes / 16,
                                                                                  (uint4*)dst, ldd_b / 16, scatter);
  } else if (elt_bytes == 2) {
    permute_rows_kernel<uint16_t><<<grid_1d(n * cols, 256), 256, 0, st>>>((const uint16_t*)src, lds, perm, n, cols, (uint16_t*)dst, ldd, scatter);
  } else if (elt_bytes == 4) {
    permute_rows_kernel<uint32_t><<<grid_1d(n * cols, 256), 256, 0, st>>>((const uint32_t*)src, lds, perm, n, cols, (uint32_t*)dst, ldd, scatter);
  } else {
    permute_rows_kernel<uint64_t><<<grid_1d(n * cols, 256), 256, 0, st>>>((const uint64_t*)src, lds, perm, n, cols, (uint64_t*)dst, ldd, scatter);
  }
  CDC_LAUNCHED();
  return 0;
}

extern "C" size_t cdcmdr_bce_segments_scratch_bytes(int n_seg) { return (size_t)(n_seg > 0 ? n_seg : 1) * kBceChunks * sizeof(double); }

extern "C" int cdcmdr_bce_segments(const float* pred, int64_t ld_pred, const void* target, int target_is_f32, const int64_t* seg_start,
                                   int n_seg, float* out_mean, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(n_seg >= 0 && n_seg <= 65535, "bad segment count");
  if (n_seg == 0) return 0;
  CDC_REQUIRE(pred && target && seg_start && out_mean && scratch && ld_pred >= 1, "bad bce_segments arguments");
  cudaStream_t st = to_stream(s);
  bce_segments_partial_kernel<<<dim3((unsigned)n_seg, kBceChunks), 256, 0, st>>>(pred, ld_pred, target, target_is_f32, seg_start,
                                                                                 (double*)scratch);
  CDC_LAUNCHED();
  bce_segments_finalize_kernel<<<(n_seg + 255) / 256, 256, 0, st>>>((const double*)scratch, seg_start, n_seg, out_mean);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_copy2d_batched(const void* src, int64_t src_bs, int64_t lds, void* dst, int64_t dst_bs, int64_t ldd, int64_t batches,
                                     int64_t rows, int64_t cols, int elt_bytes, cdcmdr_stream_t s) {
  CDC_REQUIRE(elt_bytes == 2 || elt_bytes == 4, "element size must be 2 or 4 bytes");
  if (batches <= 0 || rows <= 0 || cols <= 0) return 0;
  CDC_REQUIRE(src && dst, "null copy operand");
  // 128-bit path when every byte stride, the row length and both bases are multiples of 16 (the exchange blocks of the sharded table)
  const int64_t eb = elt_bytes;
  if ((cols * eb) % 16 == 0 && (lds * eb) % 16 == 0 && (ldd * eb) % 16 == 0 && (src_bs * eb) % 16 == 0 && (dst_bs * eb) % 16 == 0 &&
      (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    const int64_t vc = cols * eb / 16;
    copy2d_batched_kernel<uint4><<<grid_1d(batches * rows * vc, 256), 256, 0, to_stream(s)>>>(
        (const uint4*)src, src_bs * eb / 16, lds * eb / 16, (uint4*)dst, dst_bs * eb / 16, ldd * eb / 16, batches, rows, vc);
    CDC_LAUNCHED();
    return 0;
  }
  const int grid = grid_1d(batches * rows * cols, 256);
  if (elt_bytes == 4)
    copy2d_batched_kernel<uint32_t><<<grid, 256, 0, to_stream(s)>>>((const uint32_t*)src, src_bs, lds, (uint32_t*)dst, dst_bs, ldd, batches, rows, cols);
  else
    copy2d_batched_kernel<uint16_t><<<grid, 256, 0, to_stream(s)>>>((const uint16_t*)src, src_bs, lds, (uint16_t*)dst, dst_bs, ldd, batches, rows, cols);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_domain_to_group(const int32_t* x, int64_t B, int F, int domain_idx, const int64_t* domain2group, int n_domain,
                                      int64_t* groups, cdcmdr_stream_t s) {
  CDC_REQUIRE(domain_idx >= 0 && domain_idx < F && n_domain >= 1, "bad domain index");
  if (B <= 0) return 0;
  domain_to_group_kernel<<<grid_1d(B, 256), 256, 0, to_stream(s)>>>(x, B, F, domain_idx, domain2group, n_domain, groups);
  CDC_LAUNCHED();
  return 0;
}
