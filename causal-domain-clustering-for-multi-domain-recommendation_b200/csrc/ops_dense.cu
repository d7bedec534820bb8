// HBM-bound stages of the dense path: gate softmax + expert mix (a5/a9), BatchNorm (+ReLU, dropout) (a4/a8/a14),
// sigmoid + tower selection + BCE (a8/a15/a17), L2 regulariser (a7), fused Adam over the parameter arena (a17),
// column sums (bias gradients) and casts.  Every reduction is two-stage with a fixed partition -> deterministic.
#include "common.cuh"

namespace cdcmdr {

// ------------------------------------------------------------------------------------------ vector helpers
template <typename T, int VEC> struct VecIO;
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <> struct VecIO<uint16_t, 4> {
  static __device__ __forceinline__ void load(const uint16_t* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(uint16_t* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
};
template <> struct VecIO<uint16_t, 1> {
  static __device__ __forceinline__ void load(const uint16_t* p, float (&v)[1]) { v[0] = bf16_to_f32(*p); }
  static __device__ __forceinline__ void store(uint16_t* p, const float (&v)[1]) { *p = f32_to_bf16(v[0]); }
};

static int grid_1d(int64_t work, int threads, int max_per_sm = 8) {
  int64_t g = ceil_div(work, threads);
  const int64_t cap = (int64_t)kNumSMs * max_per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

constexpr int kMixStageSmem = 96 * 1024;    // per-CTA cap for the row-staging gate-mix kernels (8 warps, one row each)
constexpr int kMixSmemPerSM = 220 * 1024;
constexpr int kReducePartials = 1024;       // fixed upper bound on stage-1 blocks of scalar reductions

__device__ __forceinline__ double block_sum_256(double v) {   // deterministic: warp tree + ordered warp sum
  __shared__ double red[8];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0) for (int w = 0; w < 8; ++w) t += red[w];
  return t;                                                     // valid on thread 0
}

__global__ void reduce_finalize_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  double t = 0;
  for (int i = threadIdx.x; i < n; i += 256) t += partials[i];
  t = block_sum_256(t);
  if (threadIdx.x == 0) *out = t;
}

// ------------------------------------------------------------------------------------------ gate mix
struct MixK { int n_gates, n_experts, h, max_sel; const int32_t* gate_col; const int32_t* gate_n; const int32_t* gate_sel; };

// STAGE: the warp first copies its row of H (and dOut in the backward) into shared memory with 128-bit loads - all of
// them independent, so enough bytes are in flight to hide HBM latency - and the mixing loops then read shared memory.
// Without it every (gate, expert) pair issues its own short global load behind runtime loop bounds.
__device__ __forceinline__ void stage_row(void* dst, const void* src, int bytes, int lane) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = lane; i < bytes / 16; i += 32) d4[i] = __ldg(s4 + i);
}

template <typename T, int VEC, bool STAGE>
__global__ void __launch_bounds__(256)
gate_mix_fwd_kernel(MixK d, const T* __restrict__ H, int64_t ldh, const float* __restrict__ logits, int64_t ldl,
                    T* __restrict__ out, int64_t ldo, float* __restrict__ probs, int64_t B, int stage_off) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ int s_col[32], s_n[32], s_sel[1024];
  const int np = d.n_gates * d.max_sel;
  for (int i = threadIdx.x; i < d.n_gates; i += blockDim.x) { s_col[i] = d.gate_col[i]; s_n[i] = d.gate_n[i]; }
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_sel[i] = d.gate_sel[i] * d.h;      // element offset of the expert's block
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = dyn + warp * np;
  const int row_elems = d.n_experts * d.h;
  T* Hs = reinterpret_cast<T*>(reinterpret_cast<char*>(dyn) + stage_off) + (size_t)warp * row_elems;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < B; row += (int64_t)gridDim.x * 8) {
    if (STAGE) stage_row(Hs, H + row * ldh, row_elems * (int)sizeof(T), lane);
    const T* Hrow = STAGE ? Hs : H + row * ldh;
    if (lane < d.n_gates) {
      const int n = s_n[lane];
      const float* lg = logits + row * ldl + s_col[lane];
      float mx = -INFINITY;
      for (int s = 0; s < n; ++s) mx = fmaxf(mx, lg[s]);
      float sum = 0.f;
      for (int s = 0; s < n; ++s) { const float e = expf(lg[s] - mx); sp[lane * d.max_sel + s] = e; sum += e; }
      float* pr = probs + row * np + lane * d.max_sel;
      for (int s = 0; s < d.max_sel; ++s) {
        const float p = s < n ? sp[lane * d.max_sel + s] / sum : 0.f;
        sp[lane * d.max_sel + s] = p; pr[s] = p;
      }
    }
    __syncwarp();
    for (int c = lane * VEC; c < d.h; c += 32 * VEC) {
      for (int j = 0; j < d.n_gates; ++j) {
        float acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
        const int n = s_n[j];
        for (int s = 0; s < n; ++s) {
          float hv[VEC];
          VecIO<T, VEC>::load(Hrow + s_sel[j * d.max_sel + s] + c, hv);
          const float p = sp[j * d.max_sel + s];
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[q] = fmaf(p, hv[q], acc[q]);
        }
        VecIO<T, VEC>::store(out + row * ldo + j * d.h + c, acc);
      }
    }
    __syncwarp();
  }
}

template <typename T, int VEC, bool STAGE>
__global__ void __launch_bounds__(256)
gate_mix_bwd_kernel(MixK d, const T* __restrict__ H, int64_t ldh, const float* __restrict__ probs,
                    const T* __restrict__ dOut, int64_t ldo, T* __restrict__ dH, int64_t lddh, float relu_scale,
                    float* __restrict__ dlogits, int64_t lddl, int64_t B, int stage_off) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ int s_col[32], s_n[32], s_sel[1024], s_inv_cnt[64], s_inv[64 * 32], s_inv_o[64 * 32];
  const int np = d.n_gates * d.max_sel;
  for (int i = threadIdx.x; i < d.n_gates; i += blockDim.x) { s_col[i] = d.gate_col[i]; s_n[i] = d.gate_n[i]; }
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_sel[i] = d.gate_sel[i];
  __syncthreads();
  if (threadIdx.x < d.n_experts) {                       // inverse map expert -> (gate, slot) pairs, gate-ascending
    int cnt = 0;
    for (int j = 0; j < d.n_gates; ++j)
      for (int s = 0; s < s_n[j]; ++s)
        if (s_sel[j * d.max_sel + s] == (int)threadIdx.x) { s_inv[threadIdx.x * 32 + cnt] = j * d.max_sel + s; s_inv_o[threadIdx.x * 32 + cnt] = j * d.h; ++cnt; }
    s_inv_cnt[threadIdx.x] = cnt;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_sel[i] *= d.h;                     // element offset of the expert's block
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = dyn + (warp * 2) * np;
  float* sd = sp + np;
  const int h_elems = d.n_experts * d.h, o_elems = d.n_gates * d.h;
  T* Hs = reinterpret_cast<T*>(reinterpret_cast<char*>(dyn) + stage_off) + (size_t)warp * (h_elems + o_elems);
  T* Os = Hs + h_elems;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < B; row += (int64_t)gridDim.x * 8) {
    if (STAGE) {
      stage_row(Hs, H + row * ldh, h_elems * (int)sizeof(T), lane);
      stage_row(Os, dOut + row * ldo, o_elems * (int)sizeof(T), lane);
    }
    const T* Hrow = STAGE ? Hs : H + row * ldh;
    const T* Orow = STAGE ? Os : dOut + row * ldo;
    for (int k = lane; k < np; k += 32) sp[k] = probs[row * np + k];
    __syncwarp();
    for (int j = 0; j < d.n_gates; ++j) {
      const int n = s_n[j];
      for (int s = 0; s < n; ++s) {
        float part = 0.f;
        for (int c = lane * VEC; c < d.h; c += 32 * VEC) {
          float dv[VEC], hv[VEC];
          VecIO<T, VEC>::load(Orow + j * d.h + c, dv);
          VecIO<T, VEC>::load(Hrow + s_sel[j * d.max_sel + s] + c, hv);
#pragma unroll
          for (int q = 0; q < VEC; ++q) part = fmaf(dv[q], hv[q], part);
        }
        part = warp_sum(part);
        if (lane == 0) sd[j * d.max_sel + s] = part;
      }
    }
    __syncwarp();
    if (lane < d.n_gates) {
      const int n = s_n[lane];
      float dot = 0.f;
      for (int s = 0; s < n; ++s) dot = fmaf(sp[lane * d.max_sel + s], sd[lane * d.max_sel + s], dot);
      float* dz = dlogits + row * lddl + s_col[lane];
      for (int s = 0; s < n; ++s) dz[s] = sp[lane * d.max_sel + s] * (sd[lane * d.max_sel + s] - dot);
    }
    for (int c = lane * VEC; c < d.h; c += 32 * VEC) {
      for (int e = 0; e < d.n_experts; ++e) {
        float acc[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
        const int cnt = s_inv_cnt[e];
        for (int i = 0; i < cnt; ++i) {
          const int k = s_inv[e * 32 + i];
          float dv[VEC];
          VecIO<T, VEC>::load(Orow + s_inv_o[e * 32 + i] + c, dv);
          const float p = sp[k];
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[q] = fmaf(p, dv[q], acc[q]);
        }
        if (relu_scale > 0.f) {
          float hv[VEC];
          VecIO<T, VEC>::load(Hrow + e * d.h + c, hv);
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[q] = hv[q] > 0.f ? acc[q] * relu_scale : 0.f;
        }
        VecIO<T, VEC>::store(dH + row * lddh + e * d.h + c, acc);
      }
    }
    __syncwarp();
  }
}

// Row-per-warp backward for the common shape h == 32*VEC (every lane owns VEC columns of each expert block, no inner
// column loop) and at most 32 (gate, expert) pairs.  The generic kernel above spends ~70 instructions per pair (unrolled
// column-loop prologues, 64-bit index math, a 10-instruction warp reduction per pair; ncu: 3.5 k instructions per row,
// issue-bound).  Here: offsets precomputed per pair, one partial per lane per pair into shared memory, and ONE transposed
// reduction (lane p sums the 32 partials of pair p) instead of 26 butterfly reductions.
template <int VEC> struct VecS;
template <> struct VecS<4> {
  static __device__ __forceinline__ void load(const uint16_t* p, float (&v)[4]) { VecIO<uint16_t, 4>::load(p, v); }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { VecIO<float, 4>::load(p, v); }
  static __device__ __forceinline__ void store(uint16_t* p, const float (&v)[4]) { VecIO<uint16_t, 4>::store(p, v); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { VecIO<float, 4>::store(p, v); }
};
template <> struct VecS<2> {
  static __device__ __forceinline__ void load(const uint16_t* p, float (&v)[2]) {
    const uint32_t t = *reinterpret_cast<const uint32_t*>(p); v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u);
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
  static __device__ __forceinline__ void store(uint16_t* p, const float (&v)[2]) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};

constexpr int kMixFastPairs = 32;

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
gate_mix_bwd_row_kernel(MixK d, const T* __restrict__ H, int64_t ldh, const float* __restrict__ probs,
                        const T* __restrict__ dOut, int64_t ldo, T* __restrict__ dH, int64_t lddh, float relu_scale,
                        float* __restrict__ dlogits, int64_t lddl, int64_t B, int stage_off) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ int s_col[32], s_n[32], s_first[32], p_h[kMixFastPairs], p_o[kMixFastPairs], p_slot[kMixFastPairs];
  __shared__ int s_inv_cnt[64], s_inv_slot[64 * 8], s_inv_o[64 * 8], s_npairs;
  const int np = d.n_gates * d.max_sel, h = d.h;
  if (threadIdx.x == 0) {                                // compact list of valid (gate, slot) pairs, gate-major
    int n = 0;
    for (int j = 0; j < d.n_gates; ++j) {
      s_col[j] = d.gate_col[j]; s_n[j] = d.gate_n[j]; s_first[j] = n;
      for (int s2 = 0; s2 < d.gate_n[j]; ++s2, ++n) { p_h[n] = d.gate_sel[j * d.max_sel + s2] * h; p_o[n] = j * h; p_slot[n] = j * d.max_sel + s2; }
    }
    s_npairs = n;
  }
  __syncthreads();
  if (threadIdx.x < d.n_experts) {                       // inverse map expert -> pairs that mix it (at most 8 per expert here)
    int cnt = 0;
    for (int p = 0; p < s_npairs; ++p)
      if (p_h[p] == (int)threadIdx.x * h) { s_inv_slot[threadIdx.x * 8 + cnt] = p_slot[p]; s_inv_o[threadIdx.x * 8 + cnt] = p_o[p]; ++cnt; }
    s_inv_cnt[threadIdx.x] = cnt;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = s_npairs;
  // per warp: probabilities [np] | pair dots by slot [np] | partials [32 pairs][33]
  float* sp = dyn + warp * (2 * np + kMixFastPairs * 33);
  float* sd = sp + np;
  float* part = sd + np;
  const int h_elems = d.n_experts * h, o_elems = d.n_gates * h;
  T* Hs = reinterpret_cast<T*>(reinterpret_cast<char*>(dyn) + stage_off) + (size_t)warp * (h_elems + o_elems);
  T* Os = Hs + h_elems;
  const int lc = lane * VEC;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < B; row += (int64_t)gridDim.x * 8) {
    stage_row(Hs, H + row * ldh, h_elems * (int)sizeof(T), lane);
    stage_row(Os, dOut + row * ldo, o_elems * (int)sizeof(T), lane);
    for (int k = lane; k < np; k += 32) sp[k] = probs[row * np + k];
    __syncwarp();
    for (int p = 0; p < n_pairs; ++p) {                  // partial <dOut_j, H_e> over this lane's VEC columns
      float dv[VEC], hv[VEC];
      VecS<VEC>::load(Os + p_o[p] + lc, dv);
      VecS<VEC>::load(Hs + p_h[p] + lc, hv);
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < VEC; ++q) t = fmaf(dv[q], hv[q], t);
      part[p * 33 + lane] = t;
    }
    __syncwarp();
    if (lane < n_pairs) {                                // fixed-order sum of the 32 partials of pair `lane`
      float t = 0.f;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) t += part[lane * 33 + l];
      sd[p_slot[lane]] = t;
    }
    __syncwarp();
    if (lane < d.n_gates) {                              // softmax backward: dz = p * (dp - <p, dp>)
      const int n = s_n[lane], base = lane * d.max_sel;
      float dot = 0.f;
      for (int s2 = 0; s2 < n; ++s2) dot = fmaf(sp[base + s2], sd[base + s2], dot);
      float* dz = dlogits + row * lddl + s_col[lane];
      for (int s2 = 0; s2 < n; ++s2) dz[s2] = sp[base + s2] * (sd[base + s2] - dot);
    }
    for (int e = 0; e < d.n_experts; ++e) {              // dH_e = (sum_{pairs mixing e} p * dOut_gate) * relu'(H_e)
      float acc[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
      const int cnt = s_inv_cnt[e];
      for (int i = 0; i < cnt; ++i) {
        float dv[VEC];
        VecS<VEC>::load(Os + s_inv_o[e * 8 + i] + lc, dv);
        const float pr = sp[s_inv_slot[e * 8 + i]];
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = fmaf(pr, dv[q], acc[q]);
      }
      if (relu_scale > 0.f) {
        float hv[VEC];
        VecS<VEC>::load(Hs + e * h + lc, hv);
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = hv[q] > 0.f ? acc[q] * relu_scale : 0.f;
      }
      VecS<VEC>::store(dH + row * lddh + e * h + lc, acc);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ column reductions
constexpr int kMaxChunks = 64;

static int pick_chunks(int64_t B, int64_t C) {
  const int64_t col_tiles = ceil_div(C, 32);
  int64_t ch = ceil_div(2 * kNumSMs, col_tiles);
  if (ch > kMaxChunks) ch = kMaxChunks;
  const int64_t by_rows = ceil_div(B, 8);
  if (ch > by_rows) ch = by_rows;
  return ch < 1 ? 1 : (int)ch;
}

// F::operator()(row, col, acc[NV]) accumulates one element's contribution.  Block = (32 cols) x (8 row lanes).
template <int NV, typename F>
__global__ void __launch_bounds__(256)
col_partial_kernel(int64_t B, int64_t C, int chunks, double* __restrict__ partial, F f) {
  __shared__ double sm[NV][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int chunk = blockIdx.y;
  const int64_t rpc = ceil_div(B, chunks);
  const int64_t r0 = chunk * rpc, r1 = (r0 + rpc < B) ? r0 + rpc : B;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  if (c < C) for (int64_t r = r0 + ty; r < r1; r += 8) f(r, c, acc);
#pragma unroll
  for (int v = 0; v < NV; ++v) sm[v][ty][tx] = acc[v];
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double t = 0;
      for (int y = 0; y < 8; ++y) t += sm[v][y][tx];
      partial[((int64_t)v * chunks + chunk) * C + c] = t;
    }
  }
}

template <typename T> struct ColSumF {
  const T* X; int64_t ld;
  __device__ __forceinline__ void operator()(int64_t r, int64_t c, double (&a)[1]) const { a[0] += (double)ld_act<T>(X + r * ld + c); }
};

// Bandwidth-shaped column sum (bias gradients of the big activation-gradient matrices): every thread owns VEC = 16 bytes
// of consecutive columns and walks its chunk's rows with 128-bit loads (a warp reads 512 contiguous bytes per row), fp32
// partial per thread over <= a few hundred rows, then double across the 8 row lanes and across chunks (fixed order).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ X, int64_t ld, int64_t B, int64_t C, int chunks, double* __restrict__ partial) {
  __shared__ float sm[8][32 * VEC + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  const int chunk = blockIdx.y;
  const int64_t rpc = ceil_div(B, chunks);
  const int64_t r0 = chunk * rpc, r1 = (r0 + rpc < B) ? r0 + rpc : B;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  if (c0 < C) {
    const T* p = X + c0;
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + r * ld));
      if constexpr (VEC == 8) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[2 * k] += __uint_as_float(w[k] << 16); acc[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u); }
      } else {
        acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y); acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) sm[ty][tx * VEC + v] = acc[v];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VEC; i += 256) {
    const int64_t c = (int64_t)blockIdx.x * 32 * VEC + i;
    if (c < C) {
      double t = 0;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += (double)sm[y][i];
      partial[(int64_t)chunk * C + c] = t;
    }
  }
}

__global__ void colsum_finalize_kernel(const double* __restrict__ partial, int chunks, int64_t C, float* __restrict__ out, int accumulate) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t = 0;
  for (int k = 0; k < chunks; ++k) t += partial[(int64_t)k * C + c];
  out[c] = accumulate ? out[c] + (float)t : (float)t;
}

// Same sum for many chunks (narrow matrices are cut into up to kColsumMaxChunks row chunks so that the partial kernel fills the
// machine): block = 32 columns x 32 chunk lanes, lane y adds chunks y, y+32, ... and the 32 lane sums are added in lane order.
__global__ void __launch_bounds__(1024)
colsum_finalize_wide_kernel(const double* __restrict__ partial, int chunks, int64_t C, float* __restrict__ out, int accumulate) {
  __shared__ double sm[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  double t = 0;
  if (c < C) {
#pragma unroll 4
    for (int k = ty; k < chunks; k += 32) t += partial[(int64_t)k * C + c];
  }
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < C) {
    double u = 0;
#pragma unroll
    for (int y = 0; y < 32; ++y) u += sm[y][tx];
    out[c] = accumulate ? out[c] + (float)u : (float)u;
  }
}

// ------------------------------------------------------------------------------------------ batch norm
struct BnStatF {
  const float* Z; int64_t ldz;
  __device__ __forceinline__ void operator()(int64_t r, int64_t c, double (&a)[2]) const {
    const double x = (double)Z[r * ldz + c]; a[0] += x; a[1] += x * x;
  }
};

constexpr float kBnEps = 1e-5f, kBnMom = 0.1f;

__global__ void bn_fwd_finalize_kernel(const double* __restrict__ partial, int chunks, int64_t B, int64_t C, int train,
                                       float* __restrict__ rmean, float* __restrict__ rvar,
                                       float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!train) {
    save_mean[c] = rmean[c];
    save_invstd[c] = 1.0f / sqrtf(rvar[c] + kBnEps);
    return;
  }
  double s = 0, q = 0;
  for (int k = 0; k < chunks; ++k) { s += partial[(int64_t)k * C + c]; q += partial[((int64_t)chunks + k) * C + c]; }
  const double mean = s / (double)B;
  double var = q / (double)B - mean * mean;
  if (var < 0) var = 0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)kBnEps));
  if (rmean) {
    const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
    rmean[c] = (1.f - kBnMom) * rmean[c] + kBnMom * (float)mean;
    rvar[c] = (1.f - kBnMom) * rvar[c] + kBnMom * (float)unb;
  }
}

template <typename T>
__global__ void bn_apply_kernel(const float* __restrict__ Z, int64_t ldz, T* __restrict__ A, int64_t lda, int64_t B, int64_t C,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ gamma2, const float* __restrict__ beta2,
                                const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  const int64_t total = B * C;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u;
  const uint32_t thr = drop_thr16(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, C, r, c);
    float g = gamma[c], b = beta[c];
    if (gamma2) { g *= gamma2[c]; b += beta2[c]; }
    float v = (Z[r * ldz + c] - mean[c]) * invstd[c] * g + b;
    if (relu) v = fmaxf(v, 0.f);
    if (drop_p > 0.f) v = drop_keep(s0, thr, (uint32_t)r, (uint32_t)c) ? v * keep_scale : 0.f;
    st_act<T>(A + r * lda + c, v);
  }
}

// Two per-column sums at once, 4 columns per thread with 128-bit loads.  Block = 8 column quads (32 columns) x 32 row lanes, so
// that narrow matrices (the towers' 64..256 columns) still spread over several hundred blocks; fp32 partial per thread over its
// rows, double across row lanes and chunks (fixed order).
template <typename F>
__global__ void __launch_bounds__(256)
col_partial2_vec4_kernel(int64_t B, int64_t C, int chunks, double* __restrict__ partial, F f) {
  __shared__ float sm[2][32][33];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int64_t c0 = ((int64_t)blockIdx.x * 8 + tx) * 4;
  const int chunk = blockIdx.y;
  const int64_t rpc = ceil_div(B, chunks);
  const int64_t r0 = chunk * rpc, r1 = (r0 + rpc < B) ? r0 + rpc : B;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 < C) {
    typename F::Cols cols(f, c0);
#pragma unroll 2
    for (int64_t r = r0 + ty; r < r1; r += 32) f.add4(cols, r, c0, a, b);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { sm[0][ty][tx * 4 + q] = a[q]; sm[1][ty][tx * 4 + q] = b[q]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int v = threadIdx.x >> 5, i = threadIdx.x & 31;        // 2 statistics x 32 columns
    const int64_t c = (int64_t)blockIdx.x * 32 + i;
    if (c < C) {
      double t = 0;
#pragma unroll 8
      for (int y = 0; y < 32; ++y) t += (double)sm[v][y][i];
      partial[((int64_t)v * chunks + chunk) * C + c] = t;
    }
  }
}

struct BnStatV4 {
  const float* Z; int64_t ldz;
  struct Cols { __device__ Cols(const BnStatV4&, int64_t) {} };
  __device__ __forceinline__ void add4(const Cols&, int64_t r, int64_t c, float (&a)[4], float (&b)[4]) const {
    float z[4];
    VecIO<float, 4>::load(Z + r * ldz + c, z);
#pragma unroll
    for (int q = 0; q < 4; ++q) { a[q] += z[q]; b[q] = fmaf(z[q], z[q], b[q]); }
  }
};

template <typename T, typename TD> struct BnBwdStatV4 {
  const float* Z; int64_t ldz; const T* A; int64_t lda; const TD* dA; int64_t ldda;
  const float* mean; const float* invstd; int relu; float keep_scale;
  struct Cols {
    float m[4], iv[4];
    __device__ Cols(const BnBwdStatV4& f, int64_t c) { VecIO<float, 4>::load(f.mean + c, m); VecIO<float, 4>::load(f.invstd + c, iv); }
  };
  __device__ __forceinline__ void add4(const Cols& k, int64_t r, int64_t c, float (&a)[4], float (&b)[4]) const {
    float dy[4], z[4];
    VecIO<TD, 4>::load(dA + r * ldda + c, dy);
    if (relu) {
      float av[4];
      VecIO<T, 4>::load(A + r * lda + c, av);
#pragma unroll
      for (int q = 0; q < 4; ++q) dy[q] = av[q] > 0.f ? dy[q] * keep_scale : 0.f;
    }
    VecIO<float, 4>::load(Z + r * ldz + c, z);
#pragma unroll
    for (int q = 0; q < 4; ++q) { a[q] += dy[q]; b[q] = fmaf(dy[q], (z[q] - k.m[q]) * k.iv[q], b[q]); }
  }
};

static int pick_chunks_vec4(int64_t B, int64_t C) {
  const int64_t col_blocks = ceil_div(C, 32);
  int64_t ch = ceil_div(4 * kNumSMs, col_blocks);
  if (ch > kMaxChunks) ch = kMaxChunks;
  if (ch > B / 128) ch = B / 128;
  return ch < 1 ? 1 : (int)ch;
}

// 4 columns per thread, 128-bit accesses, 32-bit index arithmetic (one division per 4 elements): the towers' [B, 64..256]
// activations are small enough that the scalar kernel's 64-bit div/mod per element dominated its run time.
template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_vec4_kernel(const float* __restrict__ Z, int64_t ldz, T* __restrict__ A, int64_t lda, uint32_t total, uint32_t C4,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ gamma2, const float* __restrict__ beta2,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                     float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u;
  const uint32_t thr = drop_thr16(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / C4, c = (i - r * C4) * 4u;
    float z[4], g[4], b[4], m[4], iv[4], v[4];
    VecIO<float, 4>::load(Z + (int64_t)r * ldz + c, z);
    VecIO<float, 4>::load(gamma + c, g); VecIO<float, 4>::load(beta + c, b);
    VecIO<float, 4>::load(mean + c, m); VecIO<float, 4>::load(invstd + c, iv);
    if (gamma2) {
      float g2[4], b2[4];
      VecIO<float, 4>::load(gamma2 + c, g2); VecIO<float, 4>::load(beta2 + c, b2);
#pragma unroll
      for (int q = 0; q < 4; ++q) { g[q] *= g2[q]; b[q] += b2[q]; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = (z[q] - m[q]) * iv[q] * g[q] + b[q];
      if (relu) v[q] = fmaxf(v[q], 0.f);
    }
    if (drop_p > 0.f) {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = drop_keep(s0, thr, r, c + q) ? v[q] * keep_scale : 0.f;
    }
    VecIO<T, 4>::store(A + (int64_t)r * lda + c, v);
  }
}

template <typename T, typename TD, typename TZ>
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec4_kernel(const float* __restrict__ Z, int64_t ldz, const T* __restrict__ A, int64_t lda,
                         const TD* __restrict__ dA, int64_t ldda, TZ* __restrict__ dZ, int64_t lddz,
                         uint32_t total, uint32_t C4, int64_t C, int64_t n_total, const float* __restrict__ gamma,
                         const float* __restrict__ gamma2, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ sums, int relu, float keep_scale, int train) {
  const float inv_n = 1.f / (float)n_total;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / C4, c = (i - r * C4) * 4u;
    float dy[4], g[4], iv[4], dz[4];
    VecIO<TD, 4>::load(dA + (int64_t)r * ldda + c, dy);
    if (relu) {
      float a[4];
      VecIO<T, 4>::load(A + (int64_t)r * lda + c, a);
#pragma unroll
      for (int q = 0; q < 4; ++q) dy[q] = a[q] > 0.f ? dy[q] * keep_scale : 0.f;
    }
    VecIO<float, 4>::load(gamma + c, g); VecIO<float, 4>::load(invstd + c, iv);
    if (gamma2) {
      float g2[4];
      VecIO<float, 4>::load(gamma2 + c, g2);
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] *= g2[q];
    }
    if (train) {
      float z[4], m[4], s1[4], s2[4];
      VecIO<float, 4>::load(Z + (int64_t)r * ldz + c, z); VecIO<float, 4>::load(mean + c, m);
      VecIO<float, 4>::load(sums + c, s1); VecIO<float, 4>::load(sums + C + c, s2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float xh = (z[q] - m[q]) * iv[q];
        dz[q] = g[q] * iv[q] * (dy[q] - s1[q] * inv_n - xh * s2[q] * inv_n);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) dz[q] = dy[q] * g[q] * iv[q];
    }
    VecIO<TZ, 4>::store(dZ + (int64_t)r * lddz + c, dz);
  }
}

template <typename T, typename TD> struct BnBwdStatF {
  const float* Z; int64_t ldz; const T* A; int64_t lda; const TD* dA; int64_t ldda;
  const float* mean; const float* invstd; int relu; float keep_scale;
  __device__ __forceinline__ void operator()(int64_t r, int64_t c, double (&a)[2]) const {
    float dy = ld_act<TD>(dA + r * ldda + c);
    if (relu) dy = ld_act<T>(A + r * lda + c) > 0.f ? dy * keep_scale : 0.f;
    const float xh = (Z[r * ldz + c] - mean[c]) * invstd[c];
    a[0] += (double)dy; a[1] += (double)dy * (double)xh;
  }
};

// sums2[v*C + c] = sum_k partial[(v*chunks + k)*C + c], v in {0, 1}: the per-feature sums of one rank
// block = 32 (v, c) entries x 8 chunk lanes: lane y adds chunks y, y+8, ..., the 8 lane sums are added in lane order (fixed order;
// a single thread per entry walked the <= 64 chunks serially: 12 us for a few KB)
__global__ void __launch_bounds__(256)
chunks_to_sums_kernel(const double* __restrict__ partial, int chunks, int64_t C, double* __restrict__ sums2) {
  __shared__ double sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + tx;
  double t = 0;
  if (i < 2 * C) {
    const int64_t v = i / C, c = i % C;
#pragma unroll 4
    for (int k = ty; k < chunks; k += 8) t += partial[((int64_t)v * chunks + k) * C + c];
  }
  sm[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && i < 2 * C) {
    double u = 0;
#pragma unroll
    for (int y = 0; y < 8; ++y) u += sm[y][tx];
    sums2[i] = u;
  }
}

// sums[0][c] = sum dy, sums[1][c] = sum dy*xhat  (fp32 copies kept in `sums` for the apply kernel)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partial, int chunks, int64_t C, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int k = 0; k < chunks; ++k) { s += partial[(int64_t)k * C + c]; q += partial[((int64_t)chunks + k) * C + c]; }
  sums[c] = (float)s; sums[C + c] = (float)q;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
}

template <typename T, typename TD, typename TZ>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ Z, int64_t ldz, const T* __restrict__ A, int64_t lda,
                                    const TD* __restrict__ dA, int64_t ldda, TZ* __restrict__ dZ, int64_t lddz,
                                    int64_t B, int64_t n_total, int64_t C, const float* __restrict__ gamma, const float* __restrict__ gamma2,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ sums, int relu, float keep_scale, int train) {
  const int64_t total = B * C;
  const float inv_n = 1.f / (float)n_total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, C, r, c);
    float dy = ld_act<TD>(dA + r * ldda + c);
    if (relu) dy = ld_act<T>(A + r * lda + c) > 0.f ? dy * keep_scale : 0.f;
    float g = gamma[c];
    if (gamma2) g *= gamma2[c];
    float dz;
    if (train) {
      const float xh = (Z[r * ldz + c] - mean[c]) * invstd[c];
      dz = g * invstd[c] * (dy - sums[c] * inv_n - xh * sums[C + c] * inv_n);
    } else {
      dz = dy * g * invstd[c];
    }
    st_act<TZ>(dZ + r * lddz + c, dz);
  }
}

// ------------------------------------------------------------------------------------------ row dot (Linear(d, 1) heads)
// logit[b, g] = A[b, g*d : (g+1)*d] . w[g] + bias[g]     (tower output layers, layer.py:192-193)
template <typename T>
__global__ void __launch_bounds__(256)
rowdot_fwd_kernel(const T* __restrict__ A, int64_t lda, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int64_t ldo, int64_t B, int G, int d) {
  const int64_t total = B * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, g64;
    split_idx(i, G, b, g64);
    const int g = (int)g64;
    const T* a = A + b * lda + (int64_t)g * d;
    const float* wg = w + (int64_t)g * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(ld_act<T>(a + k), wg[k], acc);
    out[b * ldo + g] = acc + (bias ? bias[g] : 0.f);
  }
}

// long rows (the DCN / DCNv2 heads and the wide linear over d = F*E inputs): one warp per (row, group), lanes stride the row - a
// thread walking 832 consecutive floats of its own row reads 32 different lines per warp instruction (124 us at 16 384 x 832)
template <typename T>
__global__ void __launch_bounds__(256)
rowdot_fwd_wide_kernel(const T* __restrict__ A, int64_t lda, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ out, int64_t ldo, int64_t B, int G, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t total = B * G;
  for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * 8) {
    int64_t b, g64;
    split_idx(i, G, b, g64);
    const int g = (int)g64;
    const T* a = A + b * lda + (int64_t)g * d;
    const float* wg = w + (int64_t)g * d;
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) acc = fmaf(ld_act<T>(a + k), wg[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[b * ldo + g] = acc + (bias ? bias[g] : 0.f);
  }
}

// dA[b, g*d + k] = dlogit[b, g] * w[g, k]
__global__ void rowdot_bwd_x_kernel(const float* __restrict__ dl, int64_t ldl, const float* __restrict__ w, float* __restrict__ dA,
                                    int64_t ldda, int64_t B, int G, int d) {
  const int64_t C = (int64_t)G * d, total = B * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, c; split_idx(i, C, b, c);
    dA[b * ldda + c] = dl[b * ldl + (uint32_t)c / (uint32_t)d] * w[c];
  }
}

template <typename T> struct RowdotWF {
  const T* A; int64_t lda; const float* dl; int64_t ldl; int d;
  __device__ __forceinline__ void operator()(int64_t r, int64_t c, double (&a)[1]) const {
    a[0] += (double)(ld_act<T>(A + r * lda + c) * dl[r * ldl + c / d]);
  }
};

// ------------------------------------------------------------------------------------------ sigmoid / select / BCE
template <typename TT>
__global__ void __launch_bounds__(256)
sigmoid_select_bce_kernel(const float* __restrict__ logits, const float* __restrict__ lin, int64_t ld_lin, int64_t B, int T,
                          int mode, const int64_t* __restrict__ sel, int col, const TT* __restrict__ target,
                          float* __restrict__ pred, float* __restrict__ psel, double* __restrict__ partials,
                          float* __restrict__ dlogits, float* __restrict__ dlin, int64_t ld_dlin, float inv_batch) {
  double loss = 0.0;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float l0 = lin ? lin[b * ld_lin] : 0.f;
    int c = col;
    if (mode == 0) c = (int)sel[b];
    float ps = 0.f, ys = 0.f;
    for (int t = 0; t < T; ++t) {
      const float y = 1.f / (1.f + expf(-(logits[b * T + t] + l0)));
      pred[b * T + t] = y;
      if (mode == 2) ps += y;
      else if (t == c) { ps = y; ys = y; }
    }
    if (mode == 2) ps = ps / (float)T;
    // a selection outside [0, T) (a domain id outside the table gives -1 from domain_to_group): the reference's gather raises; here
    // the row's prediction and the step's loss become NaN - loud in step_losses() - and the row contributes no gradient
    const bool bad_sel = mode == 0 && (unsigned)c >= (unsigned)T;
    if (bad_sel) ps = __int_as_float(0x7fc00000);
    if (psel) psel[b] = ps;
    if (target) {
      const float tg = (float)target[b];
      const float lp = fmaxf(logf(ps), -100.f), l1p = fmaxf(log1pf(-ps), -100.f);
      loss += bad_sel ? (double)ps : (double)(-(tg * lp + (1.f - tg) * l1p));
      if (dlogits) {
        const float dps = (ps - tg) / fmaxf((1.f - ps) * ps, 1e-12f) * inv_batch;
        float dsum = 0.f;
        for (int t = 0; t < T; ++t) {
          float dz;
          if (mode == 2) { const float y = pred[b * T + t]; dz = dps / (float)T * y * (1.f - y); }
          else dz = (t == c && !bad_sel) ? dps * ys * (1.f - ys) : 0.f;
          dlogits[b * T + t] = dz;
          dsum += dz;
        }
        if (dlin) dlin[b * ld_dlin] = dsum;
      }
    }
  }
  if (partials) {
    loss = block_sum_256(loss);
    if (threadIdx.x == 0) partials[blockIdx.x] = loss;
  }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ dpred, float* __restrict__ dl,
                                   float* __restrict__ dlin, int64_t ld_dlin, int64_t B, int T) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float dsum = 0.f;
    for (int t = 0; t < T; ++t) {
      const float y = pred[b * T + t];
      const float dz = dpred[b * T + t] * y * (1.f - y);
      dl[b * T + t] = dz;
      dsum += dz;
    }
    if (dlin) dlin[b * ld_dlin] = dsum;
  }
}

// ------------------------------------------------------------------------------------------ regulariser / Adam
__global__ void __launch_bounds__(256)
reg_l2_partial_kernel(const float* __restrict__ w, const float* __restrict__ coef, float coef_scalar, int64_t n,
                      double* __restrict__ partials) {
  double t = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float c = coef ? coef[i] : coef_scalar;
    if (c != 0.f) t += (double)c * (double)w[i] * (double)w[i];
  }
  t = block_sum_256(t);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void reg_l2_grad_kernel(const float* __restrict__ w, const float* __restrict__ coef, float coef_scalar, float scale,
                                   float* __restrict__ grad, int accumulate, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = scale * 2.f * (coef ? coef[i] : coef_scalar) * w[i];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}
__global__ void relu_mask_kernel(const float* __restrict__ dA, int64_t ldda, const float* __restrict__ A, int64_t lda, float* __restrict__ out,
                                 int64_t ldo, int64_t rows, int64_t cols, float scale) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    out[r * ldo + c] = A[r * lda + c] > 0.f ? dA[r * ldda + c] * scale : 0.f;
  }
}

// mixed-dtype variant (bf16 path on a one-row batch): each operand is fp32 or bf16 by flag
__global__ void relu_mask_any_kernel(const void* __restrict__ dA, int64_t ldda, int da_bf16, const void* __restrict__ A, int64_t lda, int a_bf16,
                                     void* __restrict__ out, int64_t ldo, int out_bf16, int64_t rows, int64_t cols, float scale) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    const float a = a_bf16 ? bf16_to_f32(((const uint16_t*)A)[r * lda + c]) : ((const float*)A)[r * lda + c];
    const float d = da_bf16 ? bf16_to_f32(((const uint16_t*)dA)[r * ldda + c]) : ((const float*)dA)[r * ldda + c];
    const float v = a > 0.f ? d * scale : 0.f;
    if (out_bf16) ((uint16_t*)out)[r * ldo + c] = f32_to_bf16(v); else ((float*)out)[r * ldo + c] = v;
  }
}

__global__ void adam_dense_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                  const float* __restrict__ l2coef, const uint8_t* __restrict__ present, int64_t n,
                                  const cdcmdr_step_state_t* __restrict__ st) {
  const float lr_t = st->lr_t, b1 = st->beta1, b2 = st->beta2, eps = st->eps, wd = st->weight_decay, bc2 = st->bc2_sqrt;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (present && !present[i]) continue;
    float p = w[i];
    float g = grad[i];
    if (l2coef) g = g + 2.f * l2coef[i] * p;
    g = g + wd * p;
    float mi = m[i], vi = v[i];
    mi = mi + (1.f - b1) * (g - mi);
    vi = b2 * vi + (1.f - b2) * g * g;
    const float denom = sqrtf(vi) / bc2 + eps;
    p = p - lr_t * (mi / denom);
    w[i] = p; m[i] = mi; v[i] = vi;
  }
}

// ------------------------------------------------------------------------------------------ casts / elementwise
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, int64_t lds, uint16_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    dst[r * ldd + c] = f32_to_bf16(src[r * lds + c]);
  }
}
// four elements per thread (128-bit fp32 side, 64-bit bf16 side): the scalar kernels above spend a 64-bit division per element
__global__ void __launch_bounds__(256)
cast_f32_bf16_vec4_kernel(const float* __restrict__ src, int64_t lds, uint16_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols4) {
  const int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols4, r, c);
    const float4 t = __ldg(reinterpret_cast<const float4*>(src + r * lds) + c);
    *reinterpret_cast<uint2*>(dst + r * ldd + 4 * c) = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
  }
}
__global__ void __launch_bounds__(256)
cast_bf16_f32_vec4_kernel(const uint16_t* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols4,
                          int accumulate) {
  const int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols4, r, c);
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src + r * lds) + c);
    float4 t = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                           __uint_as_float(u.y & 0xffff0000u));
    float4* o = reinterpret_cast<float4*>(dst + r * ldd) + c;
    if (accumulate) { const float4 a = *o; t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w; }
    *o = t;
  }
}
__global__ void cast_bf16_f32_kernel(const uint16_t* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, int accumulate) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    const float v = bf16_to_f32(src[r * lds + c]);
    dst[r * ldd + c] = accumulate ? dst[r * ldd + c] + v : v;
  }
}
__global__ void ewise_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n, int op) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (op == 0) v = a[i] * b[i];
    else if (op == 1) v = a[i] + b[i];
    else if (op == 2) v = out[i] + a[i] * b[i];
    else v = out[i] + a[i];
    out[i] = v;
  }
}
// STAR's star-topology factors over all T towers in one launch (star.py:91-92, 100-101): a, out are [G, n] blocks, b is either one
// shared [n] block (broadcast over the towers) or a [G, n] block.
//   op 0: out[g, i] = a[g, i] * b[i]          op 1: out[g, i] = a[g, i] + b[i]
//   op 2: out[i]    = sum_g a[g, i] * b[g, i]  op 3: out[i]   += sum_g a[g, i]      op 4: out[g, i] += a[g, i]
// (fixed g order: deterministic)
__global__ void ewise_group_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n, int G, int op) {
  if (op == 0 || op == 1 || op == 4) {
    const int64_t total = n * G;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      int64_t g, j; split_idx(i, n, g, j);
      out[i] = op == 0 ? a[i] * b[j] : (op == 1 ? a[i] + b[j] : out[i] + a[i]);
    }
    return;
  }
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += op == 2 ? a[g * n + j] * b[g * n + j] : a[g * n + j];
    out[j] = op == 2 ? acc : out[j] + acc;
  }
}
__global__ void add2d_kernel(const float* __restrict__ a, int64_t lda, float* __restrict__ out, int64_t ldo, int64_t rows, int64_t cols, int accumulate) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    const float v = a[r * lda + c];
    out[r * ldo + c] = accumulate ? out[r * ldo + c] + v : v;
  }
}

}  // namespace cdcmdr

using namespace cdcmdr;

// ============================================================================================ C-ABI
namespace cdcmdr {
// gate_mix_mma.cu: bf16 rows on mma.sync; -1 = shape / layout not covered
int gate_mix_fwd_mma(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* logits, int64_t ldl, void* out, int64_t ldo,
                     float* probs, int64_t B, int is_bf16, cudaStream_t st);
int gate_mix_bwd_mma(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* probs, const void* dOut, int64_t ldo, void* dH,
                     int64_t lddh, float relu_scale, float* dlogits, int64_t lddl, int64_t B, int is_bf16, cudaStream_t st);
}
// CDCMDR_MIX_SIMT=1 keeps the SIMT gate-mix kernels for bf16 rows too (A/B measurements)
static bool mix_simt_only() {
  static const bool v = [] { const char* e = getenv("CDCMDR_MIX_SIMT"); return e && e[0] == '1'; }();
  return v;
}

static int check_mix(const cdcmdr_mix_desc_t* d) {
  CDC_REQUIRE(d && d->n_gates >= 1 && d->n_gates <= 32 && d->max_sel >= 1 && d->max_sel <= 32 && d->n_experts >= 1 &&
              d->n_experts <= 64 && d->h >= 1, "bad gate-mix descriptor");
  return 0;
}

extern "C" int cdcmdr_gate_mix_fwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* logits, int64_t ldl,
                                   void* out, int64_t ldo, float* probs, int64_t B, int is_bf16, cdcmdr_stream_t s) {
  if (int rc = check_mix(d)) return rc;
  if (B == 0) return 0;
  if (!mix_simt_only()) {
    const int rc = gate_mix_fwd_mma(d, H, ldh, logits, ldl, out, ldo, probs, B, is_bf16, to_stream(s));
    if (rc >= 0) return rc;
  }
  MixK k{d->n_gates, d->n_experts, d->h, d->max_sel, d->gate_col, d->gate_n, d->gate_sel};
  size_t smem = (size_t)8 * d->n_gates * d->max_sel * sizeof(float);
  CDC_REQUIRE(smem <= 32 * 1024, "gate-mix descriptor too large");
  const bool vec = d->h % 4 == 0 && ldh % 4 == 0 && ldo % 4 == 0 &&
                   ((uintptr_t)H % (is_bf16 ? 8 : 16) == 0) && ((uintptr_t)out % (is_bf16 ? 8 : 16) == 0);
  const size_t elt = is_bf16 ? 2 : 4;
  const size_t row_bytes = (size_t)d->n_experts * d->h * elt;
  const int stage_off = (int)((smem + 15) & ~(size_t)15);
  const bool stage = vec && B >= 64 && row_bytes % 16 == 0 && (ldh * elt) % 16 == 0 && ((uintptr_t)H % 16) == 0 &&
                     stage_off + 8 * row_bytes <= kMixStageSmem;
  if (stage) smem = stage_off + 8 * row_bytes;
  const int grid = grid_1d(B * 32, 256, stage ? (int)(kMixSmemPerSM / smem > 8 ? 8 : (kMixSmemPerSM / smem < 1 ? 1 : kMixSmemPerSM / smem)) : 8);
  cudaStream_t st = to_stream(s);
#define MIXF(T, V, S) do { if (S) CDC_CHECK(cudaFuncSetAttribute(gate_mix_fwd_kernel<T, V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixStageSmem)); \
    gate_mix_fwd_kernel<T, V, S><<<grid, 256, smem, st>>>(k, (const T*)H, ldh, logits, ldl, (T*)out, ldo, probs, B, stage_off); } while (0)
  if (is_bf16) { if (stage) MIXF(uint16_t, 4, true); else if (vec) MIXF(uint16_t, 4, false); else MIXF(uint16_t, 1, false); }
  else { if (stage) MIXF(float, 4, true); else if (vec) MIXF(float, 4, false); else MIXF(float, 1, false); }
#undef MIXF
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_gate_mix_bwd(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* probs,
                                   const void* dOut, int64_t ldo, void* dH, int64_t lddh, float relu_scale,
                                   float* dlogits, int64_t lddl, int64_t B, int is_bf16, cdcmdr_stream_t s) {
  if (int rc = check_mix(d)) return rc;
  if (B == 0) return 0;
  if (!mix_simt_only()) {
    const int rc = gate_mix_bwd_mma(d, H, ldh, probs, dOut, ldo, dH, lddh, relu_scale, dlogits, lddl, B, is_bf16, to_stream(s));
    if (rc >= 0) return rc;
  }
  MixK k{d->n_gates, d->n_experts, d->h, d->max_sel, d->gate_col, d->gate_n, d->gate_sel};
  size_t smem = (size_t)16 * d->n_gates * d->max_sel * sizeof(float);
  CDC_REQUIRE(smem <= 32 * 1024, "gate-mix descriptor too large");
  const int al = is_bf16 ? 8 : 16;
  const bool vec = d->h % 4 == 0 && ldh % 4 == 0 && ldo % 4 == 0 && lddh % 4 == 0 && ((uintptr_t)H % al == 0) &&
                   ((uintptr_t)dOut % al == 0) && ((uintptr_t)dH % al == 0);
  const size_t elt = is_bf16 ? 2 : 4;
  const size_t h_bytes = (size_t)d->n_experts * d->h * elt, o_bytes = (size_t)d->n_gates * d->h * elt;
  const int stage_off = (int)((smem + 15) & ~(size_t)15);
  const bool stage = vec && B >= 64 && h_bytes % 16 == 0 && o_bytes % 16 == 0 && (ldh * elt) % 16 == 0 && (ldo * elt) % 16 == 0 &&
                     ((uintptr_t)H % 16) == 0 && ((uintptr_t)dOut % 16) == 0 && stage_off + 8 * (h_bytes + o_bytes) <= kMixStageSmem;
  if (stage) smem = stage_off + 8 * (h_bytes + o_bytes);
  const int grid = grid_1d(B * 32, 256, stage ? (int)(kMixSmemPerSM / smem > 8 ? 8 : (kMixSmemPerSM / smem < 1 ? 1 : kMixSmemPerSM / smem)) : 8);
  cudaStream_t st = to_stream(s);
  {
    // fast row kernel: every lane owns VEC columns of each expert block (h == 32*VEC) and <= 32 (gate, expert) pairs, each expert
    // mixed by <= 8 gates.  Pair counts come from the descriptor's host mirror: the caller's n_gates*max_sel bound is enough.
    const int vecw = d->h == 128 ? 4 : (d->h == 64 ? 2 : 0);
    const size_t per_warp = (size_t)(2 * d->n_gates * d->max_sel + kMixFastPairs * 33) * sizeof(float);
    const int soff = (int)((8 * per_warp + 15) & ~(size_t)15);
    const size_t fsmem = soff + 8 * (h_bytes + o_bytes);
    if (stage && vecw && d->n_gates * d->max_sel <= 64 && d->n_gates <= 8 && fsmem <= kMixStageSmem && d->n_pairs > 0 && d->n_pairs <= kMixFastPairs) {
      const int per_sm = (int)(kMixSmemPerSM / fsmem);
      const int fgrid = grid_1d(B * 32, 256, per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm));
#define MIXR(T, V) do { CDC_CHECK(cudaFuncSetAttribute(gate_mix_bwd_row_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixStageSmem)); \
      gate_mix_bwd_row_kernel<T, V><<<fgrid, 256, fsmem, st>>>(k, (const T*)H, ldh, probs, (const T*)dOut, ldo, (T*)dH, lddh, relu_scale, dlogits, lddl, B, soff); } while (0)
      if (is_bf16) { if (vecw == 4) MIXR(uint16_t, 4); else MIXR(uint16_t, 2); }
      else { if (vecw == 4) MIXR(float, 4); else MIXR(float, 2); }
#undef MIXR
      CDC_LAUNCHED();
      return 0;
    }
  }
#define MIXB(T, V, S) do { if (S) CDC_CHECK(cudaFuncSetAttribute(gate_mix_bwd_kernel<T, V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixStageSmem)); \
    gate_mix_bwd_kernel<T, V, S><<<grid, 256, smem, st>>>(k, (const T*)H, ldh, probs, (const T*)dOut, ldo, (T*)dH, lddh, relu_scale, dlogits, lddl, B, stage_off); } while (0)
  if (is_bf16) { if (stage) MIXB(uint16_t, 4, true); else if (vec) MIXB(uint16_t, 4, false); else MIXB(uint16_t, 1, false); }
  else { if (stage) MIXB(float, 4, true); else if (vec) MIXB(float, 4, false); else MIXB(float, 1, false); }
#undef MIXB
  CDC_LAUNCHED();
  return 0;
}

// row chunks of the bandwidth-shaped column sum: enough blocks to fill the machine even when the matrix is one column block wide
constexpr int kColsumMaxChunks = 1024;
static int colsum_chunks(int64_t B, int64_t col_blocks) {
  int64_t ch = ceil_div(4 * kNumSMs, col_blocks);
  if (ch > kColsumMaxChunks) ch = kColsumMaxChunks;
  if (ch > B / 64) ch = B / 64;
  return ch < 1 ? 1 : (int)ch;
}
extern "C" size_t cdcmdr_colsum_scratch_bytes(int64_t C) {
  const int64_t c = C > 0 ? C : 1;
  // the widest vector (8 bf16) gives the fewest column blocks; a ragged C is summed as its whole vectors (>= C-7 columns) + a tail
  int64_t ch = colsum_chunks((int64_t)1 << 40, ceil_div(c > 8 ? c - 7 : 1, 256));
  if (ch < kMaxChunks) ch = kMaxChunks;
  return (size_t)ch * (size_t)(c + 8) * sizeof(double);
}
extern "C" size_t cdcmdr_bn_scratch_bytes(int64_t C) {
  const size_t c = (size_t)(C > 0 ? C : 1);
  return 2 * (size_t)kMaxChunks * c * sizeof(double) + 2 * c * sizeof(float) + 512 + 2 * c * sizeof(double);
}
extern "C" size_t cdcmdr_reduce_scratch_bytes(void) { return kReducePartials * sizeof(double); }

extern "C" int cdcmdr_colsum(const void* X, int64_t ld, int is_bf16, int64_t B, int64_t C, float* out, int accumulate,
                             void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(B >= 0 && C >= 0 && scratch, "bad colsum arguments");
  if (C == 0) return 0;
  cudaStream_t st = to_stream(s);
  double* partial = (double*)scratch;
  const int vec = is_bf16 ? 8 : 4;
  const int64_t c_pad = ceil_div(C, vec) * vec;
  if (B >= 1024 && C % vec != 0 && C > 4 * vec && ld % vec == 0 && ld < c_pad && ((uintptr_t)X % 16) == 0) {
    // whole vectors through the bandwidth-shaped kernel, the few tail columns through the generic one
    const int64_t Cm = C - C % vec;
    if (int rc = cdcmdr_colsum(X, ld, is_bf16, B, Cm, out, accumulate, scratch, s)) return rc;
    const void* Xt = is_bf16 ? (const void*)((const uint16_t*)X + Cm) : (const void*)((const float*)X + Cm);
    return cdcmdr_colsum(Xt, ld, is_bf16, B, C - Cm, out + Cm, accumulate, scratch, s);
  }
  // a last vector that hangs over C reads the row's padding columns (ld >= c_pad); their sums are never written
  if (B >= 1024 && ld >= c_pad && ld % vec == 0 && ((uintptr_t)X % 16) == 0) {
    const int64_t col_blocks = ceil_div(C, 32 * vec);
    const int ch = colsum_chunks(B, col_blocks);
    dim3 vgrid((unsigned)col_blocks, (unsigned)ch);
    if (is_bf16) colsum_vec_kernel<uint16_t, 8><<<vgrid, 256, 0, st>>>((const uint16_t*)X, ld, B, C, ch, partial);
    else colsum_vec_kernel<float, 4><<<vgrid, 256, 0, st>>>((const float*)X, ld, B, C, ch, partial);
    CDC_LAUNCHED();
    if (ch > 32) colsum_finalize_wide_kernel<<<(unsigned)ceil_div(C, 32), 1024, 0, st>>>(partial, ch, C, out, accumulate);
    else colsum_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(partial, ch, C, out, accumulate);
    CDC_LAUNCHED();
    return 0;
  }
  const int chunks = pick_chunks(B > 0 ? B : 1, C);
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks);
  if (is_bf16) col_partial_kernel<1><<<grid, 256, 0, st>>>(B, C, chunks, partial, ColSumF<uint16_t>{(const uint16_t*)X, ld});
  else col_partial_kernel<1><<<grid, 256, 0, st>>>(B, C, chunks, partial, ColSumF<float>{(const float*)X, ld});
  CDC_LAUNCHED();
  colsum_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(partial, chunks, C, out, accumulate);
  CDC_LAUNCHED();
  return 0;
}

// ---- BatchNorm in two stages, so that data-parallel ranks can all-reduce the per-feature sums in between.
// scratch layout: [2*kMaxChunks*C doubles: per-chunk partials][2*C floats: backward sums][pad to 8][2*C doubles: stage sums]
static double* bn_stage_sums(void* scratch, int64_t C) {
  size_t off = 2 * (size_t)kMaxChunks * (size_t)C * sizeof(double) + 2 * (size_t)C * sizeof(float);
  off = (off + 255) & ~(size_t)255;
  return (double*)((char*)scratch + off);
}

extern "C" int cdcmdr_bn_fwd_stats(const float* Z, int64_t ldz, int64_t B, int64_t C, double* sums, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(Z && sums && scratch, "bad batch-norm arguments");
  if (C == 0) return 0;
  cudaStream_t st = to_stream(s);
  if (B == 0) { CDC_CHECK(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(double), st)); return 0; }
  int chunks = pick_chunks(B, C);
  double* partial = (double*)scratch;
  if (B >= 256 && C % 4 == 0 && ldz % 4 == 0 && ((uintptr_t)Z % 16) == 0) {
    chunks = pick_chunks_vec4(B, C);
    dim3 vgrid((unsigned)ceil_div(C, 32), (unsigned)chunks);
    col_partial2_vec4_kernel<<<vgrid, 256, 0, st>>>(B, C, chunks, partial, BnStatV4{Z, ldz});
  } else {
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks);
    col_partial_kernel<2><<<grid, 256, 0, st>>>(B, C, chunks, partial, BnStatF{Z, ldz});
  }
  CDC_LAUNCHED();
  chunks_to_sums_kernel<<<(unsigned)ceil_div(2 * C, 32), 256, 0, st>>>(partial, chunks, C, sums);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_bn_fwd_apply(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, void* A, int64_t lda_, int a_is_bf16,
                                   int64_t B, int64_t n_total, int64_t C, const double* sums, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && p->gamma && p->beta && p->save_mean && p->save_invstd, "bad batch-norm arguments");
  CDC_REQUIRE(p->train || (p->running_mean && p->running_var), "eval batch-norm needs running statistics");
  CDC_REQUIRE(!p->train || sums, "train batch-norm needs the stage-1 sums");
  CDC_REQUIRE(p->drop_p <= 0.f || (p->seed_dev && p->relu), "batch-norm dropout needs relu and a device seed");
  CDC_REQUIRE((p->gamma2 == nullptr) == (p->beta2 == nullptr), "gamma2/beta2 must come together");
  if (C == 0) return 0;
  cudaStream_t st = to_stream(s);
  bn_fwd_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(sums, 1, n_total, C, p->train, p->running_mean, p->running_var,
                                                                    p->save_mean, p->save_invstd);
  CDC_LAUNCHED();
  if (B == 0) return 0;
  const int g = grid_1d(B * C, 256);
  const int al = a_is_bf16 ? 8 : 16;
  auto a16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
  if (C % 4 == 0 && ldz % 4 == 0 && lda_ % 4 == 0 && B * (C / 4) < ((int64_t)1 << 32) && a16(Z) && ((uintptr_t)A % al) == 0 && a16(p->gamma) &&
      a16(p->beta) && a16(p->save_mean) && a16(p->save_invstd) && (!p->gamma2 || (a16(p->gamma2) && a16(p->beta2)))) {
    const uint32_t total = (uint32_t)(B * (C / 4));
    const int gv = grid_1d(total, 256);
    if (a_is_bf16)
      bn_apply_vec4_kernel<uint16_t><<<gv, 256, 0, st>>>(Z, ldz, (uint16_t*)A, lda_, total, (uint32_t)(C / 4), p->gamma, p->beta, p->gamma2,
                                                         p->beta2, p->save_mean, p->save_invstd, p->relu, p->drop_p, p->seed_dev, p->salt);
    else
      bn_apply_vec4_kernel<float><<<gv, 256, 0, st>>>(Z, ldz, (float*)A, lda_, total, (uint32_t)(C / 4), p->gamma, p->beta, p->gamma2,
                                                      p->beta2, p->save_mean, p->save_invstd, p->relu, p->drop_p, p->seed_dev, p->salt);
    CDC_LAUNCHED();
    return 0;
  }
  if (a_is_bf16)
    bn_apply_kernel<uint16_t><<<g, 256, 0, st>>>(Z, ldz, (uint16_t*)A, lda_, B, C, p->gamma, p->beta, p->gamma2, p->beta2, p->save_mean,
                                                 p->save_invstd, p->relu, p->drop_p, p->seed_dev, p->salt);
  else
    bn_apply_kernel<float><<<g, 256, 0, st>>>(Z, ldz, (float*)A, lda_, B, C, p->gamma, p->beta, p->gamma2, p->beta2, p->save_mean,
                                              p->save_invstd, p->relu, p->drop_p, p->seed_dev, p->salt);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_bn_fwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, void* A, int64_t lda_, int a_is_bf16,
                             int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && scratch, "bad batch-norm arguments");
  if (B == 0 || C == 0) return 0;
  double* sums = bn_stage_sums(scratch, C);
  if (p->train) { if (int rc = cdcmdr_bn_fwd_stats(Z, ldz, B, C, sums, scratch, s)) return rc; }
  return cdcmdr_bn_fwd_apply(p, Z, ldz, A, lda_, a_is_bf16, B, B, C, sums, s);
}

// stage 1 of the backward: LOCAL sums over the rows (sums[0..C) = sum dy, sums[C..2C) = sum dy*xhat) and the parameter
// gradients they are (dgamma = sum dy*xhat, dbeta = sum dy; the data-parallel all-reduce of the gradient arena adds ranks)
extern "C" int cdcmdr_bn_bwd_stats(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                                   const void* dA, int64_t ldda, int da_is_bf16, float* dgamma, float* dbeta, int accumulate,
                                   int64_t B, int64_t C, double* sums, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && p->gamma && p->save_mean && p->save_invstd && scratch && sums, "bad batch-norm arguments");
  CDC_REQUIRE(!p->relu || A, "relu backward needs the forward output");
  if (C == 0) return 0;
  cudaStream_t st = to_stream(s);
  if (B == 0) {
    CDC_CHECK(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(double), st));
    if (!accumulate) {
      if (dgamma) CDC_CHECK(cudaMemsetAsync(dgamma, 0, (size_t)C * sizeof(float), st));
      if (dbeta) CDC_CHECK(cudaMemsetAsync(dbeta, 0, (size_t)C * sizeof(float), st));
    }
    return 0;
  }
  CDC_REQUIRE(dA, "bad batch-norm arguments");
  int chunks = pick_chunks(B, C);
  double* partial = (double*)scratch;
  float* fsums = (float*)((char*)scratch + 2 * (size_t)kMaxChunks * (size_t)C * sizeof(double));
  const float keep_scale = p->drop_p > 0.f ? 1.f / (1.f - p->drop_p) : 1.f;
  const int combo = (a_is_bf16 ? 2 : 0) | (da_is_bf16 ? 1 : 0);
  CDC_REQUIRE(combo == 0 || combo == 2 || combo == 3, "batch-norm backward dtypes: (A, dA) must be (fp32, fp32), (bf16, fp32) or (bf16, bf16)");
  auto al = [](const void* q, int is_bf16) { return ((uintptr_t)q % (is_bf16 ? 8 : 16)) == 0; };
  if (B >= 256 && C % 4 == 0 && ldz % 4 == 0 && ldda % 4 == 0 && (!p->relu || lda_ % 4 == 0) && al(Z, 0) && al(dA, da_is_bf16) &&
      (!p->relu || al(A, a_is_bf16)) && al(p->save_mean, 0) && al(p->save_invstd, 0)) {
    chunks = pick_chunks_vec4(B, C);
    dim3 vgrid((unsigned)ceil_div(C, 32), (unsigned)chunks);
#define BNSV(TA, TD) col_partial2_vec4_kernel<<<vgrid, 256, 0, st>>>(B, C, chunks, partial, \
      BnBwdStatV4<TA, TD>{Z, ldz, (const TA*)A, lda_, (const TD*)dA, ldda, p->save_mean, p->save_invstd, p->relu, keep_scale})
    if (combo == 0) BNSV(float, float); else if (combo == 2) BNSV(uint16_t, float); else BNSV(uint16_t, uint16_t);
#undef BNSV
  } else {
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks);
#define BNS(TA, TD) col_partial_kernel<2><<<grid, 256, 0, st>>>(B, C, chunks, partial, \
      BnBwdStatF<TA, TD>{Z, ldz, (const TA*)A, lda_, (const TD*)dA, ldda, p->save_mean, p->save_invstd, p->relu, keep_scale})
    if (combo == 0) BNS(float, float); else if (combo == 2) BNS(uint16_t, float); else BNS(uint16_t, uint16_t);
#undef BNS
  }
  CDC_LAUNCHED();
  chunks_to_sums_kernel<<<(unsigned)ceil_div(2 * C, 32), 256, 0, st>>>(partial, chunks, C, sums);
  CDC_LAUNCHED();
  bn_bwd_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(sums, 1, C, fsums, dgamma, dbeta, accumulate);
  CDC_LAUNCHED();
  return 0;
}

// stage 2: dZ for the local rows from sums over ALL n_total rows
extern "C" int cdcmdr_bn_bwd_apply(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                                   const void* dA, int64_t ldda, int da_is_bf16, void* dZ, int64_t lddz, int dz_is_bf16,
                                   int64_t B, int64_t n_total, int64_t C, const double* sums, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && p->gamma && p->save_mean && p->save_invstd && scratch && sums, "bad batch-norm arguments");
  CDC_REQUIRE(!p->relu || A, "relu backward needs the forward output");
  if (B == 0 || C == 0) return 0;
  CDC_REQUIRE(dA && dZ, "bad batch-norm arguments");
  cudaStream_t st = to_stream(s);
  float* fsums = (float*)((char*)scratch + 2 * (size_t)kMaxChunks * (size_t)C * sizeof(double));
  const float keep_scale = p->drop_p > 0.f ? 1.f / (1.f - p->drop_p) : 1.f;
  const int combo = (a_is_bf16 ? 4 : 0) | (da_is_bf16 ? 2 : 0) | (dz_is_bf16 ? 1 : 0);
  CDC_REQUIRE(combo == 0 || combo == 5 || combo == 7, "batch-norm backward dtypes: (A, dA, dZ) must be fp32x3, (bf16, fp32, bf16) or bf16x3");
  bn_bwd_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(sums, 1, C, fsums, nullptr, nullptr, 0);
  CDC_LAUNCHED();
  const int g = grid_1d(B * C, 256);
  {
    auto al = [](const void* q, int is_bf16) { return ((uintptr_t)q % (is_bf16 ? 8 : 16)) == 0; };
    auto a16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
    if (C % 4 == 0 && ldz % 4 == 0 && ldda % 4 == 0 && lddz % 4 == 0 && (!p->relu || lda_ % 4 == 0) && B * (C / 4) < ((int64_t)1 << 32) && a16(Z) &&
        al(dA, da_is_bf16) && al(dZ, dz_is_bf16) && (!p->relu || al(A, a_is_bf16)) && a16(p->gamma) && a16(p->save_mean) && a16(p->save_invstd) &&
        a16(fsums) && (!p->gamma2 || a16(p->gamma2))) {
      const uint32_t total = (uint32_t)(B * (C / 4));
      const int gv = grid_1d(total, 256);
#define BNV(TA, TD, TZ) bn_bwd_apply_vec4_kernel<TA, TD, TZ><<<gv, 256, 0, st>>>(Z, ldz, (const TA*)A, lda_, (const TD*)dA, ldda, (TZ*)dZ, lddz, \
      total, (uint32_t)(C / 4), C, n_total, p->gamma, p->gamma2, p->save_mean, p->save_invstd, fsums, p->relu, keep_scale, p->train)
      if (combo == 0) BNV(float, float, float); else if (combo == 5) BNV(uint16_t, float, uint16_t); else BNV(uint16_t, uint16_t, uint16_t);
#undef BNV
      CDC_LAUNCHED();
      return 0;
    }
  }
#define BNB(TA, TD, TZ) bn_bwd_apply_kernel<TA, TD, TZ><<<g, 256, 0, st>>>(Z, ldz, (const TA*)A, lda_, (const TD*)dA, ldda, (TZ*)dZ, lddz, \
      B, n_total, C, p->gamma, p->gamma2, p->save_mean, p->save_invstd, fsums, p->relu, keep_scale, p->train)
  if (combo == 0) BNB(float, float, float); else if (combo == 5) BNB(uint16_t, float, uint16_t); else BNB(uint16_t, uint16_t, uint16_t);
#undef BNB
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_bn_bwd(const cdcmdr_bn_t* p, const float* Z, int64_t ldz, const void* A, int64_t lda_, int a_is_bf16,
                             const void* dA, int64_t ldda, int da_is_bf16, void* dZ, int64_t lddz, int dz_is_bf16, float* dgamma, float* dbeta,
                             int accumulate, int64_t B, int64_t C, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && scratch && dA && dZ, "bad batch-norm arguments");
  if (B == 0 || C == 0) return 0;
  double* sums = bn_stage_sums(scratch, C);
  if (int rc = cdcmdr_bn_bwd_stats(p, Z, ldz, A, lda_, a_is_bf16, dA, ldda, da_is_bf16, dgamma, dbeta, accumulate, B, C, sums, scratch, s)) return rc;
  return cdcmdr_bn_bwd_apply(p, Z, ldz, A, lda_, a_is_bf16, dA, ldda, da_is_bf16, dZ, lddz, dz_is_bf16, B, B, C, sums, scratch, s);
}

extern "C" int cdcmdr_rowdot_fwd(const void* A, int64_t lda_, int a_is_bf16, const float* w, const float* bias, float* out, int64_t ldo,
                                 int64_t B, int G, int d, cdcmdr_stream_t s) {
  CDC_REQUIRE(A && w && out && G >= 1 && d >= 1, "bad rowdot arguments");
  if (B <= 0) return 0;
  if (d >= 128) {
    int64_t gw = ceil_div(B * G, 8);
    if (gw > 16 * kNumSMs) gw = 16 * kNumSMs;
    if (a_is_bf16) rowdot_fwd_wide_kernel<uint16_t><<<(int)gw, 256, 0, to_stream(s)>>>((const uint16_t*)A, lda_, w, bias, out, ldo, B, G, d);
    else rowdot_fwd_wide_kernel<float><<<(int)gw, 256, 0, to_stream(s)>>>((const float*)A, lda_, w, bias, out, ldo, B, G, d);
    CDC_LAUNCHED();
    return 0;
  }
  const int g = grid_1d(B * G, 256);
  if (a_is_bf16) rowdot_fwd_kernel<uint16_t><<<g, 256, 0, to_stream(s)>>>((const uint16_t*)A, lda_, w, bias, out, ldo, B, G, d);
  else rowdot_fwd_kernel<float><<<g, 256, 0, to_stream(s)>>>((const float*)A, lda_, w, bias, out, ldo, B, G, d);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_rowdot_bwd(const void* A, int64_t lda_, int a_is_bf16, const float* w, const float* dlogit, int64_t ldl,
                                 float* dA, int64_t ldda, float* dW, float* dbias, int64_t B, int G, int d, void* scratch,
                                 cdcmdr_stream_t s) {
  CDC_REQUIRE(A && w && dlogit && G >= 1 && d >= 1 && scratch, "bad rowdot arguments");
  if (B <= 0) return 0;
  cudaStream_t st = to_stream(s);
  const int64_t C = (int64_t)G * d;
  if (dA) {
    rowdot_bwd_x_kernel<<<grid_1d(B * C, 256), 256, 0, st>>>(dlogit, ldl, w, dA, ldda, B, G, d);
    CDC_LAUNCHED();
  }
  if (dW) {
    const int chunks = pick_chunks(B, C);
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks);
    double* partial = (double*)scratch;
    if (a_is_bf16) col_partial_kernel<1><<<grid, 256, 0, st>>>(B, C, chunks, partial, RowdotWF<uint16_t>{(const uint16_t*)A, lda_, dlogit, ldl, d});
    else col_partial_kernel<1><<<grid, 256, 0, st>>>(B, C, chunks, partial, RowdotWF<float>{(const float*)A, lda_, dlogit, ldl, d});
    CDC_LAUNCHED();
    colsum_finalize_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(partial, chunks, C, dW, 0);
    CDC_LAUNCHED();
  }
  if (dbias) return cdcmdr_colsum(dlogit, ldl, 0, B, G, dbias, 0, scratch, s);
  return 0;
}

extern "C" int cdcmdr_sigmoid_select_bce(const float* logits, const float* lin, int64_t ld_lin, int64_t B, int32_t T, int32_t mode,
                                         const int64_t* sel, int32_t col, const void* target, int target_is_f32,
                                         float* pred, float* psel, double* loss_sum, float* dlogits,
                                         float* dlin, int64_t ld_dlin, float inv_batch, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(logits && pred && T >= 1 && mode >= 0 && mode <= 3, "bad sigmoid/BCE arguments");
  CDC_REQUIRE(mode != 0 || sel, "mode 0 needs per-sample columns");
  CDC_REQUIRE(mode != 1 || (col >= 0 && col < T), "mode 1 column out of range");
  CDC_REQUIRE(mode != 3 || !target, "mode 3 is forward only");
  CDC_REQUIRE(!target || (loss_sum && scratch), "loss needs an output and scratch");
  CDC_REQUIRE(!dlin || dlogits, "dlin needs dlogits");
  if (B == 0) return 0;
  cudaStream_t st = to_stream(s);
  int grid = grid_1d(B, 256);
  if (grid > kReducePartials) grid = kReducePartials;
  double* partials = target ? (double*)scratch : nullptr;
  if (mode == 3) col = -1;
  if (target_is_f32)
    sigmoid_select_bce_kernel<float><<<grid, 256, 0, st>>>(logits, lin, ld_lin, B, T, mode, sel, col, (const float*)target, pred, psel,
                                                           partials, dlogits, dlin, ld_dlin, inv_batch);
  else
    sigmoid_select_bce_kernel<int16_t><<<grid, 256, 0, st>>>(logits, lin, ld_lin, B, T, mode, sel, col, (const int16_t*)target, pred, psel,
                                                             partials, dlogits, dlin, ld_dlin, inv_batch);
  CDC_LAUNCHED();
  if (target) {
    reduce_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, loss_sum);
    CDC_LAUNCHED();
  }
  return 0;
}

extern "C" int cdcmdr_sigmoid_bwd(const float* pred, const float* dpred, float* dlogits, float* dlin, int64_t ld_dlin,
                                  int64_t B, int32_t T, cdcmdr_stream_t s) {
  if (B <= 0 || T <= 0) return 0;
  sigmoid_bwd_kernel<<<grid_1d(B, 256), 256, 0, to_stream(s)>>>(pred, dpred, dlogits, dlin, ld_dlin, B, T);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_reg_l2_sum(const float* w, const float* coef, float coef_scalar, int64_t n, double* out_sum, void* scratch,
                                 cdcmdr_stream_t s) {
  CDC_REQUIRE(w && out_sum && scratch && n >= 0, "bad reg_l2_sum arguments");
  cudaStream_t st = to_stream(s);
  int grid = grid_1d(n > 0 ? n : 1, 256);
  if (grid > kReducePartials) grid = kReducePartials;
  reg_l2_partial_kernel<<<grid, 256, 0, st>>>(w, coef, coef_scalar, n, (double*)scratch);
  CDC_LAUNCHED();
  reduce_finalize_kernel<<<1, 256, 0, st>>>((const double*)scratch, grid, out_sum);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_reg_l2_grad(const float* w, const float* coef, float coef_scalar, float scale, float* grad, int accumulate,
                                  int64_t n, cdcmdr_stream_t s) {
  if (n <= 0) return 0;
  reg_l2_grad_kernel<<<grid_1d(n, 256), 256, 0, to_stream(s)>>>(w, coef, coef_scalar, scale, grad, accumulate, n);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_relu_mask_f32(const float* dA, int64_t ldda, const float* A, int64_t lda_, float* out, int64_t ldo, int64_t rows,
                                    int64_t cols, float scale, cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  relu_mask_kernel<<<grid_1d(rows * cols, 256), 256, 0, to_stream(s)>>>(dA, ldda, A, lda_, out, ldo, rows, cols, scale);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_relu_mask(const void* dA, int64_t ldda, int da_bf16, const void* A, int64_t lda_, int a_bf16, void* out, int64_t ldo,
                                int out_bf16, int64_t rows, int64_t cols, float scale, cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  CDC_REQUIRE(dA && A && out, "relu_mask: null operand");
  relu_mask_any_kernel<<<grid_1d(rows * cols, 256), 256, 0, to_stream(s)>>>(dA, ldda, da_bf16, A, lda_, a_bf16, out, ldo, out_bf16, rows, cols, scale);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_adam_dense(float* w, const float* grad, float* m, float* v, const float* l2coef, const uint8_t* present,
                                 int64_t n, const cdcmdr_step_state_t* st, cdcmdr_stream_t s) {
  CDC_REQUIRE(w && grad && m && v && st, "bad Adam arguments");
  if (n <= 0) return 0;
  adam_dense_kernel<<<grid_1d(n, 256), 256, 0, to_stream(s)>>>(w, grad, m, v, l2coef, present, n, st);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_cast_f32_bf16(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols, cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0)
    cast_f32_bf16_vec4_kernel<<<grid_1d(rows * (cols / 4), 256), 256, 0, to_stream(s)>>>(src, lds, dst, ldd, rows, cols / 4);
  else
    cast_f32_bf16_kernel<<<grid_1d(rows * cols, 256), 256, 0, to_stream(s)>>>(src, lds, dst, ldd, rows, cols);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_cast_bf16_f32(const uint16_t* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols, int accumulate,
                                    cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src % 8) == 0 && ((uintptr_t)dst % 16) == 0)
    cast_bf16_f32_vec4_kernel<<<grid_1d(rows * (cols / 4), 256), 256, 0, to_stream(s)>>>(src, lds, dst, ldd, rows, cols / 4, accumulate);
  else
    cast_bf16_f32_kernel<<<grid_1d(rows * cols, 256), 256, 0, to_stream(s)>>>(src, lds, dst, ldd, rows, cols, accumulate);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_ewise_f32(const float* a, const float* b, float* out, int64_t n, int op, cdcmdr_stream_t s) {
  CDC_REQUIRE(op >= 0 && op <= 3, "bad elementwise op");
  if (n <= 0) return 0;
  ewise_kernel<<<grid_1d(n, 256), 256, 0, to_stream(s)>>>(a, b, out, n, op);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_ewise_group_f32(const float* a, const float* b, float* out, int64_t n, int G, int op, cdcmdr_stream_t s) {
  CDC_REQUIRE(op >= 0 && op <= 4 && G >= 1, "bad grouped elementwise op");
  if (n <= 0) return 0;
  const int64_t work = (op == 2 || op == 3) ? n : n * G;
  ewise_group_kernel<<<grid_1d(work, 256), 256, 0, to_stream(s)>>>(a, b, out, n, G, op);
  CDC_LAUNCHED();
  return 0;
}
extern "C" int cdcmdr_add2d_f32(const float* a, int64_t lda, float* out, int64_t ldo, int64_t rows, int64_t cols, int accumulate,
                                cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  add2d_kernel<<<grid_1d(rows * cols, 256), 256, 0, to_stream(s)>>>(a, lda, out, ldo, rows, cols, accumulate);
  CDC_LAUNCHED();
  return 0;
}
