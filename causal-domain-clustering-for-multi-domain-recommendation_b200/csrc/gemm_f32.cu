// Strided / grouped fp32 GEMM on CUDA cores: the exact-parity path for every nn.Linear / F.linear /
// torch.matmul call site of the reference (layer.py:119,185,193,275,336; star.py:93,102).  64x64x16 tiles,
// 4x4 register micro-tiles, fused bias / ReLU / mask / dropout / accumulate epilogue, deterministic split-K.
#include "common.cuh"

namespace cdcmdr {

constexpr int BM = 64, BN = 64, BK = 16;

struct EpiF32 {
  float* C; int64_t c_rs, c_gs;
  const float* bias; int64_t bias_gs;
  int act;
  const float* mask; int64_t mask_rs, mask_gs; float mask_scale;
  float drop_p; const uint64_t* seed_dev; uint32_t salt;
  int accumulate;
};

__device__ __forceinline__ void epilogue_store(const EpiF32& e, int g, int64_t m, int64_t n, float v) {
  if (e.bias) v += e.bias[g * e.bias_gs + n];
  if (e.act == 1) v = fmaxf(v, 0.f);
  if (e.mask) v = (e.mask[g * e.mask_gs + m * e.mask_rs + n] > 0.f) ? v * e.mask_scale : 0.f;
  if (e.drop_p > 0.f) {
    const bool keep = drop_keep(drop_s0(*e.seed_dev, e.salt), drop_thr16(e.drop_p), (uint32_t)m, (uint32_t)(g * e.c_gs + n));
    v = keep ? v * (1.f / (1.f - e.drop_p)) : 0.f;
  }
  float* c = e.C + g * e.c_gs + m * e.c_rs + n;
  *c = e.accumulate ? (*c + v) : v;
}

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bt, int64_t M, int64_t N, int64_t K,
                int64_t a_rs, int64_t a_cs, int64_t b_rs, int64_t b_cs, int64_t a_gs, int64_t b_gs,
                int split_k, float* __restrict__ partial, EpiF32 epi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int g = blockIdx.z / split_k, z = blockIdx.z % split_k;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kchunk = ceil_div(ceil_div(K, split_k), BK) * BK;
  const int64_t k_lo = z * kchunk, k_hi = (k_lo + kchunk < K) ? k_lo + kchunk : K;
  const float* Ag = A + g * a_gs;
  const float* Bg = Bt + g * b_gs;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool a_kfast = (a_cs == 1), b_kfast = (b_cs == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_lo; k0 < k_hi; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_hi) ? Ag[m * a_rs + k * a_cs] : 0.f;
      int nn;
      if (b_kfast) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      const int64_t n = n0 + nn; const int64_t kb = k0 + kk;
      Bs[kk][nn] = (n < N && kb < k_hi) ? Bg[n * b_rs + kb * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (split_k > 1) partial[(((int64_t)z * gridDim.z / split_k + g) * M + m) * N + n] = acc[i][j];
      else epilogue_store(epi, g, m, n, acc[i][j]);
    }
  }
}

__global__ void splitk_epilogue_kernel(const float* __restrict__ partial, int split_k, int G, int64_t M, int64_t N, EpiF32 epi) {
  const int64_t total = (int64_t)G * M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < split_k; ++z) v += partial[(int64_t)z * total + i];
    const int64_t n = i % N, m = (i / N) % M; const int g = (int)(i / (N * M));
    epilogue_store(epi, g, m, n, v);
  }
}

}  // namespace cdcmdr
using namespace cdcmdr;

extern "C" int cdcmdr_gemm_f32(const cdcmdr_gemm_f32_t* p, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && p->G >= 1 && p->M >= 0 && p->N >= 0 && p->K >= 0, "bad gemm shape");
  if (p->M == 0 || p->N == 0) return 0;
  const int split = p->split_k > 1 ? p->split_k : 1;
  CDC_REQUIRE(split == 1 || p->workspace, "split-K needs a workspace");
  CDC_REQUIRE(p->drop_p <= 0.f || p->seed_dev, "dropout needs a device seed");
  EpiF32 e{p->C, p->c_rs, p->c_gs, p->bias, p->bias_gs, p->act, p->mask, p->mask_rs, p->mask_gs, p->mask_scale,
           p->drop_p, p->seed_dev, p->salt, p->accumulate};
  dim3 grid((unsigned)ceil_div(p->N, BN), (unsigned)ceil_div(p->M, BM), (unsigned)(p->G * split));
  CDC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm grid too large");
  gemm_f32_kernel<<<grid, 256, 0, to_stream(s)>>>(p->A, p->Bt, p->M, p->N, p->K, p->a_rs, p->a_cs, p->b_rs, p->b_cs, p->a_gs, p->b_gs,
                                                  split, p->workspace, e);
  CDC_LAUNCHED();
  if (split > 1) {
    const int64_t total = (int64_t)p->G * p->M * p->N;
    int grid2 = (int)(ceil_div(total, 256) < 8 * kNumSMs ? ceil_div(total, 256) : 8 * kNumSMs);
    splitk_epilogue_kernel<<<grid2, 256, 0, to_stream(s)>>>(p->workspace, split, p->G, p->M, p->N, e);
    CDC_LAUNCHED();
  }
  return 0;
}
