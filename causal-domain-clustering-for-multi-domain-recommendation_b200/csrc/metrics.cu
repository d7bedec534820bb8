// N4: the evaluation metrics of Run.test / evaluate_multi_domain (reference run.py:647-711: scikit-learn
// roc_auc_score + log_loss over the whole validation set and per domain, after a .cpu().numpy() of every batch) on the device.
//
// AUC of a set = (sum of the MIDRANKS of its positives - n_pos (n_pos + 1) / 2) / (n_pos * n_neg): the Mann-Whitney statistic with
// ties counted one half, which is what the trapezoidal ROC integral of roc_auc_score evaluates to.  All sets at once:
//   key = (domain << 32) | order-preserving bits of the fp32 prediction  ->  one radix sort of (key, sample index)
//   run heads (key != previous key) -> inclusive sum = run id per position; run starts scattered by the heads
//   domain starts = lower bound of (d << 32) in the sorted keys
//   per position: midrank inside its domain = (run start - domain start) + (run length + 1) / 2
//   per domain: fixed-shape two-stage reduction (chunk partials in double, then the chunks in order) of
//               {midranks of positives, positives, log-loss terms} - deterministic, no atomics.
// log loss as sklearn: p clipped to [eps, 1 - eps] with eps = FLT_EPSILON (predictions are fp32), -mean(y log p + (1-y) log(1-p)).
// A set with one class only has no AUC: NaN for both metrics (the reference's `except ValueError` branch, run.py:699-703).
#include "common.cuh"
#include <cub/cub.cuh>
#include <float.h>

namespace cdcmdr {

constexpr int kMetricChunks = 64;                  // chunk partials per domain

__device__ __forceinline__ uint32_t sortable_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <typename TD>
__global__ void metric_keys_kernel(const float* __restrict__ pred, const TD* __restrict__ domain, int64_t n, int n_domain,
                                   unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    long long d = domain ? (long long)domain[i] : 0;
    if (d < 0 || d >= n_domain) d = n_domain;                  // out-of-range domains sort last and are ignored
    keys[i] = ((unsigned long long)d << 32) | sortable_f32(pred[i]);
    vals[i] = (int32_t)i;
  }
}

__global__ void metric_heads_kernel(const unsigned long long* __restrict__ keys, int64_t n, int32_t* __restrict__ head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// run_start[r] = first position of run r (r = runid - 1); run_start[num_runs] = n
__global__ void metric_runs_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ runid, int64_t n,
                                   int32_t* __restrict__ run_start) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (head[i]) run_start[runid[i] - 1] = (int32_t)i;
    if (i == n - 1) run_start[runid[i]] = (int32_t)n;
  }
}

__global__ void metric_domain_starts_kernel(const unsigned long long* __restrict__ keys, int64_t n, int n_domain, int64_t* __restrict__ dstart) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > n_domain) return;
  const unsigned long long want = (unsigned long long)d << 32;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  dstart[d] = lo;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  __syncthreads();
  return t;                                                    // valid in thread 0
}

// grid (kMetricChunks, n_domain): chunk c of domain d -> partial[(d * kMetricChunks + c) * 3 + {rank sum of positives, positives, loss sum}]
template <typename TY>
__global__ void __launch_bounds__(256)
metric_partials_kernel(const float* __restrict__ pred, const TY* __restrict__ target, const int32_t* __restrict__ vals,
                       const int32_t* __restrict__ runid, const int32_t* __restrict__ run_start, const int64_t* __restrict__ dstart,
                       double* __restrict__ partial) {
  __shared__ double red[8];
  const int d = blockIdx.y, c = blockIdx.x;
  const int64_t s = dstart[d], e = dstart[d + 1];
  const int64_t len = e - s, per = (len + kMetricChunks - 1) / kMetricChunks;
  const int64_t j0 = s + c * per, j1 = (j0 + per < e) ? j0 + per : e;
  double rs = 0, np = 0, ls = 0;
  for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    const int32_t i = vals[j];
    const int32_t r = runid[j] - 1;
    const int64_t rs0 = run_start[r], rc = run_start[r + 1] - rs0;
    const double midrank = (double)(rs0 - s) + 0.5 * (double)(rc + 1);
    const float y = (float)target[i];
    float p = pred[i];
    p = fminf(fmaxf(p, FLT_EPSILON), 1.f - FLT_EPSILON);
    if (y > 0.f) { rs += midrank; np += 1.0; ls -= log((double)p); }
    else ls -= log(1.0 - (double)p);
  }
  const double a = block_sum_d(rs, red), b = block_sum_d(np, red), l = block_sum_d(ls, red);
  if (threadIdx.x == 0) {
    double* o = partial + ((int64_t)d * kMetricChunks + c) * 3;
    o[0] = a; o[1] = b; o[2] = l;
  }
}

// out[d*4 + {auc, logloss, n_pos, n}]
__global__ void metric_finalize_kernel(const double* __restrict__ partial, const int64_t* __restrict__ dstart, int n_domain, double* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_domain) return;
  double rs = 0, np = 0, ls = 0;
  for (int c = 0; c < kMetricChunks; ++c) {
    const double* p = partial + ((int64_t)d * kMetricChunks + c) * 3;
    rs += p[0]; np += p[1]; ls += p[2];
  }
  const double n = (double)(dstart[d + 1] - dstart[d]), nn = n - np;
  const double nanv = __longlong_as_double(0x7ff8000000000000ll);
  const bool ok = np > 0 && nn > 0;
  out[d * 4 + 0] = ok ? (rs - np * (np + 1.0) * 0.5) / (np * nn) : nanv;
  out[d * 4 + 1] = ok ? ls / n : nanv;
  out[d * 4 + 2] = np;
  out[d * 4 + 3] = n;
}

struct MetricLayout { size_t keys_in, keys, vals_in, vals, head, runid, run_start, dstart, partial, cub, cub_bytes, total; };

static MetricLayout metric_layout(int64_t n, int n_domain) {
  MetricLayout L{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
  L.keys_in = take(n * 8); L.keys = take(n * 8); L.vals_in = take(n * 4); L.vals = take(n * 4);
  L.head = take(n * 4); L.runid = take(n * 4); L.run_start = take((n + 1) * 4); L.dstart = take((size_t)(n_domain + 2) * 8);
  L.partial = take((size_t)n_domain * kMetricChunks * 3 * 8);
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  L.cub_bytes = a > b ? a : b;
  L.cub = take(L.cub_bytes);
  L.total = o;
  return L;
}

template <typename T> static inline T* mat(void* base, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(base) + off); }

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" size_t cdcmdr_auc_logloss_scratch_bytes(int64_t n, int32_t n_domain) {
  if (n <= 0) n = 1;
  if (n_domain < 1) n_domain = 1;
  return metric_layout(n, n_domain).total;
}

extern "C" int cdcmdr_auc_logloss(const float* pred, const void* target, int target_is_f32, const void* domain, int domain_is_i64,
                                  int64_t n, int32_t n_domain, double* out, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(n_domain >= 1 && out && scratch, "metrics: bad arguments");
  CDC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) - 1, "metrics: sample count out of int32 range");
  cudaStream_t st = to_stream(s);
  if (n == 0) {
    CDC_CHECK(cudaMemsetAsync(out, 0xFF, (size_t)n_domain * 4 * sizeof(double), st));   // all-ones bit pattern = NaN
    return 0;
  }
  CDC_REQUIRE(pred && target, "metrics: null operand");
  const MetricLayout L = metric_layout(n, n_domain);
  auto* keys_in = mat<unsigned long long>(scratch, L.keys_in); auto* keys = mat<unsigned long long>(scratch, L.keys);
  auto* vals_in = mat<int32_t>(scratch, L.vals_in); auto* vals = mat<int32_t>(scratch, L.vals);
  auto* head = mat<int32_t>(scratch, L.head); auto* runid = mat<int32_t>(scratch, L.runid);
  auto* run_start = mat<int32_t>(scratch, L.run_start); auto* dstart = mat<int64_t>(scratch, L.dstart);
  auto* partial = mat<double>(scratch, L.partial);
  int64_t g = ceil_div(n, 256);
  if (g > 8 * kNumSMs) g = 8 * kNumSMs;
  if (domain && domain_is_i64) metric_keys_kernel<int64_t><<<(int)g, 256, 0, st>>>(pred, (const int64_t*)domain, n, n_domain, keys_in, vals_in);
  else metric_keys_kernel<int32_t><<<(int)g, 256, 0, st>>>(pred, (const int32_t*)domain, n, n_domain, keys_in, vals_in);
  CDC_LAUNCHED();
  int dbits = 1;
  while ((1 << dbits) <= n_domain) ++dbits;
  size_t tb = L.cub_bytes;
  CDC_CHECK(cub::DeviceRadixSort::SortPairs(mat<void>(scratch, L.cub), tb, keys_in, keys, vals_in, vals, (int)n, 0, 32 + dbits, st));
  metric_heads_kernel<<<(int)g, 256, 0, st>>>(keys, n, head);
  CDC_LAUNCHED();
  tb = L.cub_bytes;
  CDC_CHECK(cub::DeviceScan::InclusiveSum(mat<void>(scratch, L.cub), tb, head, runid, (int)n, st));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  metric_runs_kernel<<<(int)g, 256, 0, st>>>(head, runid, n, run_start);
  CDC_LAUNCHED();
  metric_domain_starts_kernel<<<(unsigned)ceil_div(n_domain + 1, 128), 128, 0, st>>>(keys, n, n_domain, dstart);
  CDC_LAUNCHED();
  if (target_is_f32)
    metric_partials_kernel<float><<<dim3(kMetricChunks, (unsigned)n_domain), 256, 0, st>>>(pred, (const float*)target, vals, runid, run_start, dstart, partial);
  else
    metric_partials_kernel<int16_t><<<dim3(kMetricChunks, (unsigned)n_domain), 256, 0, st>>>(pred, (const int16_t*)target, vals, runid, run_start, dstart, partial);
  CDC_LAUNCHED();
  metric_finalize_kernel<<<(unsigned)ceil_div(n_domain, 128), 128, 0, st>>>(partial, dstart, n_domain, out);
  CDC_LAUNCHED();
  return 0;
}
