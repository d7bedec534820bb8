// N3: the field self-attention block of BaseModel.atten_forward (reference model/layer.py:58-84): the attention
// core of nn.MultiheadAttention over the L = field_num tokens of one sample, and the ReLU + Linear(F*A -> 1) head that turns the
// block's output into the scalar added to every tower logit.  The projections around the core (atten_embedding, in_proj,
// out_proj, V_res_embedding) are plain Linear layers over the [B*L, A] token matrix and run on the GEMM entry points.
//
// L <= 32 tokens, so one (sample, head) pair is a warp-sized problem: a warp stages Q, K, V [L, dh] (and dO in the backward) in
// its own shared-memory slice (row stride dh+1: the score loop reads one row per lane), lane j owns score column j, the softmax
// reductions are warp shuffles, and the products against V / K / Q put lanes on the head dimension.  fp32 throughout.
#include "common.cuh"

namespace cdcmdr {

constexpr int kAttnWarps = 4;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// dropout on the attention weights (nn.MultiheadAttention(dropout=p), train mode): element (pair, i, j) -> hash row pair*L+i, col j
__device__ __forceinline__ float attn_keep_scale(float drop_p, uint32_t s0, uint32_t thr, uint32_t row, uint32_t col) {
  if (drop_p <= 0.f) return 1.f;
  return drop_keep(s0, thr, row, col) ? 1.f / (1.f - drop_p) : 0.f;
}

__global__ void __launch_bounds__(kAttnWarps * 32)
attn_fwd_kernel(const float* __restrict__ qkv, int64_t ld, float* __restrict__ out, int64_t ldo, float* __restrict__ probs,
                int64_t B, int L, int H, int dh, float scale, float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = dh + 1, A = H * dh;
  float* Q = smem + (size_t)warp * (3 * L * ds + L * L);
  float* K = Q + L * ds;
  float* V = K + L * ds;
  float* P = V + L * ds;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * kAttnWarps + warp; pr < pairs; pr += (int64_t)gridDim.x * kAttnWarps) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const float* base = qkv + b * L * ld + h * dh;
    for (int e = lane; e < L * dh; e += 32) {
      const int l = e / dh, d = e - l * dh;
      const float* row = base + (int64_t)l * ld + d;
      Q[l * ds + d] = row[0];
      K[l * ds + d] = row[A];
      V[l * ds + d] = row[2 * A];
    }
    __syncwarp();
    for (int i = 0; i < L; ++i) {                              // lane j: score (i, j)
      float s = -INFINITY;
      if (lane < L) {
        float acc = 0.f;
        for (int d = 0; d < dh; ++d) acc = fmaf(Q[i * ds + d], K[lane * ds + d], acc);
        s = acc * scale;
      }
      const float mx = warp_max(s);
      const float ex = lane < L ? expf(s - mx) : 0.f;
      const float sum = warp_sum(ex);
      if (lane < L) {
        const float p = ex / sum;
        if (probs) probs[(pr * L + i) * L + lane] = p;
        P[i * L + lane] = p * attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)lane);
      }
    }
    __syncwarp();
    float* obase = out + b * L * ldo + h * dh;
    for (int d = lane; d < dh; d += 32) {                      // lane d: O[i, d] = sum_j P[i, j] V[j, d]
      for (int i = 0; i < L; ++i) {
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(P[i * L + j], V[j * ds + d], acc);
        obase[(int64_t)i * ldo + d] = acc;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kAttnWarps * 32)
attn_bwd_kernel(const float* __restrict__ qkv, int64_t ld, const float* __restrict__ probs, const float* __restrict__ dout, int64_t lddo,
                float* __restrict__ dqkv, int64_t lddq, int64_t B, int L, int H, int dh, float scale, float drop_p,
                const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = dh + 1, A = H * dh;
  float* Q = smem + (size_t)warp * (4 * L * ds + 2 * L * L);
  float* K = Q + L * ds;
  float* V = K + L * ds;
  float* dO = V + L * ds;
  float* Pd = dO + L * ds;                                     // dropped + rescaled probabilities (what multiplied V)
  float* dS = Pd + L * L;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * kAttnWarps + warp; pr < pairs; pr += (int64_t)gridDim.x * kAttnWarps) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const float* base = qkv + b * L * ld + h * dh;
    const float* dobase = dout + b * L * lddo + h * dh;
    for (int e = lane; e < L * dh; e += 32) {
      const int l = e / dh, d = e - l * dh;
      const float* row = base + (int64_t)l * ld + d;
      Q[l * ds + d] = row[0];
      K[l * ds + d] = row[A];
      V[l * ds + d] = row[2 * A];
      dO[l * ds + d] = dobase[(int64_t)l * lddo + d];
    }
    __syncwarp();
    for (int i = 0; i < L; ++i) {                              // lane j: dP(i, j) = dO[i] . V[j]; softmax backward along j
      float p = 0.f, dp = 0.f, ks = 1.f;
      if (lane < L) {
        p = probs[(pr * L + i) * L + lane];
        ks = attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)lane);
        float acc = 0.f;
        for (int d = 0; d < dh; ++d) acc = fmaf(dO[i * ds + d], V[lane * ds + d], acc);
        dp = acc * ks;                                         // gradient w.r.t. the softmax output
      }
      const float dot = warp_sum(p * dp);
      if (lane < L) {
        Pd[i * L + lane] = p * ks;
        dS[i * L + lane] = p * (dp - dot) * scale;             // gradient w.r.t. q.k (the 1/sqrt(dh) folded in)
      }
    }
    __syncwarp();
    float* dbase = dqkv + b * L * lddq + h * dh;
    for (int d = lane; d < dh; d += 32) {
      for (int i = 0; i < L; ++i) {                            // dQ[i, d] = sum_j dS[i, j] K[j, d]
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(dS[i * L + j], K[j * ds + d], acc);
        dbase[(int64_t)i * lddq + d] = acc;
      }
      for (int j = 0; j < L; ++j) {                            // dK[j, d] = sum_i dS[i, j] Q[i, d] ; dV[j, d] = sum_i Pd[i, j] dO[i, d]
        float ak = 0.f, av = 0.f;
        for (int i = 0; i < L; ++i) {
          ak = fmaf(dS[i * L + j], Q[i * ds + d], ak);
          av = fmaf(Pd[i * L + j], dO[i * ds + d], av);
        }
        dbase[(int64_t)j * lddq + A + d] = ak;
        dbase[(int64_t)j * lddq + 2 * A + d] = av;
      }
    }
    __syncwarp();
  }
}

// head: lin[b] (+)= sum_j relu(z[b, j]) * w[j]            (F.relu -> view(B, F*A) -> atten_linear, layer.py:82-83)
__global__ void __launch_bounds__(256)
attn_pool_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w, float* __restrict__ lin, int64_t ld_lin, int accumulate,
                     int64_t B, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* zr = z + row * n;
  float acc = 0.f;
  for (int64_t j = lane; j < n; j += 32) acc = fmaf(fmaxf(zr[j], 0.f), w[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) lin[row * ld_lin] = accumulate ? lin[row * ld_lin] + acc : acc;
}

// backward of the head: dz[b, j] = dlin[b] * w[j] * [z > 0] ; partial[chunk, j] = sum over the chunk's rows of dlin[b] * relu(z[b, j])
constexpr int kPoolRowsPerBlock = 8;
__global__ void __launch_bounds__(32 * kPoolRowsPerBlock)
attn_pool_bwd_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ dlin, int64_t ld_dlin,
                     float* __restrict__ dz, double* __restrict__ partial, int64_t B, int64_t n, int64_t rows_per_chunk) {
  __shared__ double red[kPoolRowsPerBlock][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = (r0 + rows_per_chunk < B) ? r0 + rows_per_chunk : B;
  double acc = 0.0;
  if (j < n) {
    const float wj = w[j];
    for (int64_t r = r0 + ty; r < r1; r += kPoolRowsPerBlock) {
      const float zv = z[r * n + j], g = dlin[r * ld_dlin];
      dz[r * n + j] = zv > 0.f ? g * wj : 0.f;
      acc += (double)(g * fmaxf(zv, 0.f));
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < n) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPoolRowsPerBlock; ++k) s += red[k][tx];
    partial[(int64_t)blockIdx.y * n + j] = s;
  }
}

__global__ void attn_pool_bwd_finalize_kernel(const double* __restrict__ partial, int chunks, int64_t n, float* __restrict__ dw) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(int64_t)c * n + j];
  dw[j] = (float)s;
}

// ------------------------------------------------------------------------------------------ bf16 token path
// The tensor-core path keeps the whole block's token matrices in bf16 (the projections run on the tcgen05 GEMM): the core reads
// bf16 q / k / v, keeps its arithmetic in fp32, writes bf16, and stores NO probabilities - the backward recomputes the L x L
// softmax from q and k (34 kFLOP per (sample, head), against a 4*L*L-byte round trip through HBM per layer).
__device__ __forceinline__ void attn_stage_bf16(const uint16_t* __restrict__ base, int64_t ld, int col0, float* __restrict__ dst, int L, int dh,
                                                int ds, int lane) {
  const int hp = dh >> 1;                                      // bf16 pairs per row (dh is even on this path)
  for (int e = lane; e < L * hp; e += 32) {
    const int l = e / hp, d2 = e - l * hp;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(base + (int64_t)l * ld + col0 + 2 * d2);
    dst[l * ds + 2 * d2] = __uint_as_float(w << 16);
    dst[l * ds + 2 * d2 + 1] = __uint_as_float(w & 0xFFFF0000u);
  }
}

// softmax row i of the (sample, head) pair `pr` from staged Q / K: returns this lane's probability p(i, lane) (0 for lane >= L)
__device__ __forceinline__ float attn_prob_row(const float* __restrict__ Q, const float* __restrict__ K, int i, int L, int dh, int ds, float scale,
                                               int lane) {
  float s = -INFINITY;
  if (lane < L) {
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(Q[i * ds + d], K[lane * ds + d], acc);
    s = acc * scale;
  }
  const float mx = warp_max(s);
  const float ex = lane < L ? expf(s - mx) : 0.f;
  const float sum = warp_sum(ex);
  return lane < L ? ex / sum : 0.f;
}

__global__ void __launch_bounds__(kAttnWarps * 32)
attn_fwd_bf16_kernel(const uint16_t* __restrict__ qkv, int64_t ld, uint16_t* __restrict__ out, int64_t ldo, int64_t B, int L, int H, int dh,
                     float scale, float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = dh + 1, A = H * dh;
  float* Q = smem + (size_t)warp * (3 * L * ds + L * L);
  float* K = Q + L * ds;
  float* V = K + L * ds;
  float* P = V + L * ds;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * kAttnWarps + warp; pr < pairs; pr += (int64_t)gridDim.x * kAttnWarps) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const uint16_t* base = qkv + b * L * ld;
    attn_stage_bf16(base, ld, h * dh, Q, L, dh, ds, lane);
    attn_stage_bf16(base, ld, A + h * dh, K, L, dh, ds, lane);
    attn_stage_bf16(base, ld, 2 * A + h * dh, V, L, dh, ds, lane);
    __syncwarp();
    for (int i = 0; i < L; ++i) {
      const float p = attn_prob_row(Q, K, i, L, dh, ds, scale, lane);
      if (lane < L) P[i * L + lane] = p * attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)lane);
    }
    __syncwarp();
    uint16_t* obase = out + b * L * ldo + h * dh;
    for (int d = lane; d < dh; d += 32) {
      for (int i = 0; i < L; ++i) {
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(P[i * L + j], V[j * ds + d], acc);
        obase[(int64_t)i * ldo + d] = f32_to_bf16(acc);
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kAttnWarps * 32)
attn_bwd_bf16_kernel(const uint16_t* __restrict__ qkv, int64_t ld, const uint16_t* __restrict__ dout, int64_t lddo, uint16_t* __restrict__ dqkv,
                     int64_t lddq, int64_t B, int L, int H, int dh, float scale, float drop_p, const uint64_t* __restrict__ seed_dev,
                     uint32_t salt) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ds = dh + 1, A = H * dh;
  float* Q = smem + (size_t)warp * (4 * L * ds + 2 * L * L);
  float* K = Q + L * ds;
  float* V = K + L * ds;
  float* dO = V + L * ds;
  float* Pd = dO + L * ds;
  float* dS = Pd + L * L;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * kAttnWarps + warp; pr < pairs; pr += (int64_t)gridDim.x * kAttnWarps) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const uint16_t* base = qkv + b * L * ld;
    attn_stage_bf16(base, ld, h * dh, Q, L, dh, ds, lane);
    attn_stage_bf16(base, ld, A + h * dh, K, L, dh, ds, lane);
    attn_stage_bf16(base, ld, 2 * A + h * dh, V, L, dh, ds, lane);
    attn_stage_bf16(dout + b * L * lddo, lddo, h * dh, dO, L, dh, ds, lane);
    __syncwarp();
    for (int i = 0; i < L; ++i) {
      const float p = attn_prob_row(Q, K, i, L, dh, ds, scale, lane);     // recomputed, not stored by the forward
      float dp = 0.f, ks = 1.f;
      if (lane < L) {
        ks = attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)lane);
        float acc = 0.f;
        for (int d = 0; d < dh; ++d) acc = fmaf(dO[i * ds + d], V[lane * ds + d], acc);
        dp = acc * ks;
      }
      const float dot = warp_sum(p * dp);
      if (lane < L) {
        Pd[i * L + lane] = p * ks;
        dS[i * L + lane] = p * (dp - dot) * scale;
      }
    }
    __syncwarp();
    uint16_t* dbase = dqkv + b * L * lddq + h * dh;
    for (int d = lane; d < dh; d += 32) {
      for (int i = 0; i < L; ++i) {
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(dS[i * L + j], K[j * ds + d], acc);
        dbase[(int64_t)i * lddq + d] = f32_to_bf16(acc);
      }
      for (int j = 0; j < L; ++j) {
        float ak = 0.f, av = 0.f;
        for (int i = 0; i < L; ++i) {
          ak = fmaf(dS[i * L + j], Q[i * ds + d], ak);
          av = fmaf(Pd[i * L + j], dO[i * ds + d], av);
        }
        dbase[(int64_t)j * lddq + A + d] = f32_to_bf16(ak);
        dbase[(int64_t)j * lddq + 2 * A + d] = f32_to_bf16(av);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ bf16 core on mma.sync
// One warp per (sample, head); L <= 32 tokens and dh in {16, 32, 64} make every product a handful of m16n8k16 tensor-core
// instructions on registers: Q, K, V (and dO) are loaded once as 8x8 bf16 blocks - thread (g = lane/4, t = lane%4) holds elements
// (row g, columns 2t, 2t+1) of each block, one 32-bit load - which is at once the A-fragment layout, the B-fragment layout of the
// TRANSPOSED operand, and the packed C-fragment layout; where a product needs the other orientation (P V, dS K, dS^T Q, P^T dO)
// the block goes through movmatrix.trans.  Scores, softmax, dropout and the softmax backward stay fp32 in the accumulator
// fragments (row reductions = 8 local values + two quad shuffles).  The SIMT kernels above spent their time in ~2 k shared-memory
// loads per pair (2.6 ms forward / 6.0 ms backward per layer at 65 536 x 23 tokens x 2 heads); this is ~100 instructions per pair.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t blk_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// the [32 rows, DH columns] bf16 matrix at `base` (row pitch ld) as 8x8 blocks blk[rb][cb]; rows >= L read as zero
template <int DH>
__device__ __forceinline__ void load_blocks(const uint16_t* __restrict__ base, int64_t ld, int L, int g, int t, uint32_t (&blk)[4][DH / 8]) {
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) {
    const int row = 8 * rb + g;
#pragma unroll
    for (int cb = 0; cb < DH / 8; ++cb)
      blk[rb][cb] = row < L ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)row * ld + 8 * cb + 2 * t)) : 0u;
  }
}
// accumulator tiles acc[mt][nt] (rows 16mt + g (+8), columns 8nt + 2t, +1) -> bf16 rows of `base`
template <int DH>
__device__ __forceinline__ void store_tiles(uint16_t* __restrict__ base, int64_t ld, int L, int g, int t, const float (&acc)[2][DH / 8][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      if (r0 < L) *reinterpret_cast<uint32_t*>(base + (int64_t)r0 * ld + 8 * nt + 2 * t) = pack_bf16x2(acc[mt][nt][0], acc[mt][nt][1]);
      if (r1 < L) *reinterpret_cast<uint32_t*>(base + (int64_t)r1 * ld + 8 * nt + 2 * t) = pack_bf16x2(acc[mt][nt][2], acc[mt][nt][3]);
    }
}
// S = scale * Q K^T -> row softmax over the L valid columns (fp32, in the accumulator fragments p[mt][nt][4])
template <int DH>
__device__ __forceinline__ void softmax_tiles(const uint32_t (&Qb)[4][DH / 8], const uint32_t (&Kb)[4][DH / 8], int L, float scale, int t,
                                              float (&p)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      p[mt][nt][0] = p[mt][nt][1] = p[mt][nt][2] = p[mt][nt][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < DH / 16; ++kt)
        mma_bf16_16816(p[mt][nt], Qb[2 * mt][2 * kt], Qb[2 * mt + 1][2 * kt], Qb[2 * mt][2 * kt + 1], Qb[2 * mt + 1][2 * kt + 1],
                       Kb[nt][2 * kt], Kb[nt][2 * kt + 1]);
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {                          // row 16mt + g + 8hh: its 32 columns live in this quad
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = 8 * nt + 2 * t + c;
          float v = p[mt][nt][2 * hh + c] * scale;
          v = j < L ? v : -INFINITY;
          p[mt][nt][2 * hh + c] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float e = expf(p[mt][nt][2 * hh + c] - mx);      // exp(-inf) = 0 for the masked columns
          p[mt][nt][2 * hh + c] = e;
          sum += e;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) p[mt][nt][2 * hh + c] *= inv;
    }
}

template <int DH>
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(const uint16_t* __restrict__ qkv, int64_t ld, uint16_t* __restrict__ out, int64_t ldo, int64_t B, int L, int H, float scale,
                    float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int A = H * DH;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); pr < pairs; pr += (int64_t)gridDim.x * 4) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const uint16_t* base = qkv + b * L * ld + h * DH;
    uint32_t Qb[4][DH / 8], Kb[4][DH / 8], Vb[4][DH / 8];
    load_blocks<DH>(base, ld, L, g, t, Qb);
    load_blocks<DH>(base + A, ld, L, g, t, Kb);
    load_blocks<DH>(base + 2 * A, ld, L, g, t, Vb);
    float p[2][4][4];
    softmax_tiles<DH>(Qb, Kb, L, scale, t, p);
    uint32_t Pb[4][4];                                         // dropped, rescaled probabilities as bf16 blocks [row block][column block]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 16 * mt + g + 8 * (e >> 1), j = 8 * nt + 2 * t + (e & 1);
          v[e] = p[mt][nt][e] * attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)j);
        }
        Pb[2 * mt][nt] = pack_bf16x2(v[0], v[1]);
        Pb[2 * mt + 1][nt] = pack_bf16x2(v[2], v[3]);
      }
    float o[2][DH / 8][4];
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      uint32_t vt[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) vt[rb] = blk_trans(Vb[rb][nt]);       // B[k = token j][n = d]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          mma_bf16_16816(o[mt][nt], Pb[2 * mt][2 * kt], Pb[2 * mt + 1][2 * kt], Pb[2 * mt][2 * kt + 1], Pb[2 * mt + 1][2 * kt + 1],
                         vt[2 * kt], vt[2 * kt + 1]);
      }
    }
    store_tiles<DH>(out + b * L * ldo + h * DH, ldo, L, g, t, o);
  }
}

template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_mma_kernel(const uint16_t* __restrict__ qkv, int64_t ld, const uint16_t* __restrict__ dout, int64_t lddo, uint16_t* __restrict__ dqkv,
                    int64_t lddq, int64_t B, int L, int H, float scale, float drop_p, const uint64_t* __restrict__ seed_dev, uint32_t salt) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int A = H * DH;
  const uint32_t s0 = drop_p > 0.f ? drop_s0(*seed_dev, salt) : 0u, thr = drop_thr16(drop_p);
  const int64_t pairs = B * H;
  for (int64_t pr = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); pr < pairs; pr += (int64_t)gridDim.x * 4) {
    const int64_t b = pr / H;
    const int h = (int)(pr - b * H);
    const uint16_t* base = qkv + b * L * ld + h * DH;
    uint16_t* dbase = dqkv + b * L * lddq + h * DH;
    uint32_t Qb[4][DH / 8], Kb[4][DH / 8], Vb[4][DH / 8], Gb[4][DH / 8];
    load_blocks<DH>(base, ld, L, g, t, Qb);
    load_blocks<DH>(base + A, ld, L, g, t, Kb);
    load_blocks<DH>(base + 2 * A, ld, L, g, t, Vb);
    load_blocks<DH>(dout + b * L * lddo + h * DH, lddo, L, g, t, Gb);
    float p[2][4][4];
    softmax_tiles<DH>(Qb, Kb, L, scale, t, p);                  // recomputed: the forward stores no probabilities
    // dP = dO V^T (gradient w.r.t. the dropped probabilities), then the softmax backward row by row
    uint32_t Sb[4][4], Pb[4][4];                               // dS and the dropped probabilities as bf16 blocks [i block][j block]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < DH / 16; ++kt)
          mma_bf16_16816(dp[nt], Gb[2 * mt][2 * kt], Gb[2 * mt + 1][2 * kt], Gb[2 * mt][2 * kt + 1], Gb[2 * mt + 1][2 * kt + 1],
                         Vb[nt][2 * kt], Vb[nt][2 * kt + 1]);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int i = 16 * mt + g + 8 * hh;
        float ks[4][2], dot = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            ks[nt][c] = attn_keep_scale(drop_p, s0, thr, (uint32_t)(pr * L + i), (uint32_t)(8 * nt + 2 * t + c));
            dp[nt][2 * hh + c] *= ks[nt][c];
            dot = fmaf(p[mt][nt][2 * hh + c], dp[nt][2 * hh + c], dot);
          }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float ds0 = p[mt][nt][2 * hh] * (dp[nt][2 * hh] - dot) * scale, ds1 = p[mt][nt][2 * hh + 1] * (dp[nt][2 * hh + 1] - dot) * scale;
          Sb[2 * mt + hh][nt] = pack_bf16x2(ds0, ds1);
          Pb[2 * mt + hh][nt] = pack_bf16x2(p[mt][nt][2 * hh] * ks[nt][0], p[mt][nt][2 * hh + 1] * ks[nt][1]);
        }
      }
    }
    float acc[2][DH / 8][4];
    // dQ = dS K
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      uint32_t kt_[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) kt_[rb] = blk_trans(Kb[rb][nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          mma_bf16_16816(acc[mt][nt], Sb[2 * mt][2 * kt], Sb[2 * mt + 1][2 * kt], Sb[2 * mt][2 * kt + 1], Sb[2 * mt + 1][2 * kt + 1],
                         kt_[2 * kt], kt_[2 * kt + 1]);
      }
    }
    store_tiles<DH>(dbase, lddq, L, g, t, acc);
    // dK = dS^T Q and dV = Pd^T dO: the transposed [j block][i block] operands are movmatrix images of the blocks above
    uint32_t St[4][4], Pt[4][4];
#pragma unroll
    for (int jb = 0; jb < 4; ++jb)
#pragma unroll
      for (int ib = 0; ib < 4; ++ib) { St[jb][ib] = blk_trans(Sb[ib][jb]); Pt[jb][ib] = blk_trans(Pb[ib][jb]); }
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      uint32_t qt[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) qt[rb] = blk_trans(Qb[rb][nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          mma_bf16_16816(acc[mt][nt], St[2 * mt][2 * kt], St[2 * mt + 1][2 * kt], St[2 * mt][2 * kt + 1], St[2 * mt + 1][2 * kt + 1],
                         qt[2 * kt], qt[2 * kt + 1]);
      }
    }
    store_tiles<DH>(dbase + A, lddq, L, g, t, acc);
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      uint32_t gt[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) gt[rb] = blk_trans(Gb[rb][nt]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          mma_bf16_16816(acc[mt][nt], Pt[2 * mt][2 * kt], Pt[2 * mt + 1][2 * kt], Pt[2 * mt][2 * kt + 1], Pt[2 * mt + 1][2 * kt + 1],
                         gt[2 * kt], gt[2 * kt + 1]);
      }
    }
    store_tiles<DH>(dbase + 2 * A, lddq, L, g, t, acc);
  }
}

static inline bool attn_mma_ok(int dh, int64_t ld, int64_t ld2, int H) {
  return (dh == 16 || dh == 32 || dh == 64) && ld % 2 == 0 && ld2 % 2 == 0 && ((int64_t)H * dh) % 2 == 0;
}

// head on bf16 tokens: one warp per sample, 8 bf16 per 128-bit load
__global__ void __launch_bounds__(256)
attn_pool_fwd_bf16_kernel(const uint16_t* __restrict__ z, const float* __restrict__ w, float* __restrict__ lin, int64_t ld_lin, int accumulate,
                          int64_t B, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const uint16_t* zr = z + row * n;
  float acc = 0.f;
  for (int64_t j = lane; j < n; j += 32) acc = fmaf(fmaxf(bf16_to_f32(zr[j]), 0.f), w[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) lin[row * ld_lin] = accumulate ? lin[row * ld_lin] + acc : acc;
}

__global__ void __launch_bounds__(32 * kPoolRowsPerBlock)
attn_pool_bwd_bf16_kernel(const uint16_t* __restrict__ z, const float* __restrict__ w, const float* __restrict__ dlin, int64_t ld_dlin,
                          uint16_t* __restrict__ dz, double* __restrict__ partial, int64_t B, int64_t n, int64_t rows_per_chunk) {
  __shared__ double red[kPoolRowsPerBlock][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = (r0 + rows_per_chunk < B) ? r0 + rows_per_chunk : B;
  double acc = 0.0;
  if (j < n) {
    const float wj = w[j];
    for (int64_t r = r0 + ty; r < r1; r += kPoolRowsPerBlock) {
      const float zv = bf16_to_f32(z[r * n + j]), g = dlin[r * ld_dlin];
      dz[r * n + j] = f32_to_bf16(zv > 0.f ? g * wj : 0.f);
      acc += (double)(g * fmaxf(zv, 0.f));
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < n) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPoolRowsPerBlock; ++k) s += red[k][tx];
    partial[(int64_t)blockIdx.y * n + j] = s;
  }
}

static inline int pool_chunks(int64_t B) {
  const int64_t c = ceil_div(B, 64);
  return (int)(c < 1 ? 1 : (c > 256 ? 256 : c));
}

static inline int attn_grid(int64_t pairs) {
  const int64_t blocks = ceil_div(pairs, kAttnWarps);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" int cdcmdr_attn_fwd(const float* qkv, int64_t ld, float* out, int64_t ldo, float* probs, int64_t B, int L, int H, int dh,
                               float scale, float drop_p, const uint64_t* seed_dev, uint32_t salt, cdcmdr_stream_t s) {
  CDC_REQUIRE(L >= 1 && L <= 32, "attention runs over at most 32 field tokens");
  CDC_REQUIRE(H >= 1 && dh >= 1 && dh <= 256, "bad head geometry");
  CDC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev), "bad attention dropout");
  if (B <= 0) return 0;
  CDC_REQUIRE(qkv && out && ld >= 3 * (int64_t)H * dh && ldo >= (int64_t)H * dh, "bad attention operands");
  const size_t smem = (size_t)kAttnWarps * (3 * (size_t)L * (dh + 1) + (size_t)L * L) * sizeof(float);
  CDC_REQUIRE(smem <= 200 * 1024, "attention tile does not fit shared memory");
  CDC_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<<<attn_grid(B * H), kAttnWarps * 32, smem, to_stream(s)>>>(qkv, ld, out, ldo, probs, B, L, H, dh, scale, drop_p, seed_dev,
                                                                            salt);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_attn_bwd(const float* qkv, int64_t ld, const float* probs, const float* dout, int64_t lddo, float* dqkv,
                               int64_t lddq, int64_t B, int L, int H, int dh, float scale, float drop_p, const uint64_t* seed_dev,
                               uint32_t salt, cdcmdr_stream_t s) {
  CDC_REQUIRE(L >= 1 && L <= 32, "attention runs over at most 32 field tokens");
  CDC_REQUIRE(H >= 1 && dh >= 1 && dh <= 256, "bad head geometry");
  CDC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev), "bad attention dropout");
  if (B <= 0) return 0;
  CDC_REQUIRE(qkv && probs && dout && dqkv && ld >= 3 * (int64_t)H * dh && lddq >= 3 * (int64_t)H * dh && lddo >= (int64_t)H * dh,
              "bad attention operands");
  const size_t smem = (size_t)kAttnWarps * (4 * (size_t)L * (dh + 1) + 2 * (size_t)L * L) * sizeof(float);
  CDC_REQUIRE(smem <= 200 * 1024, "attention tile does not fit shared memory");
  CDC_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_kernel<<<attn_grid(B * H), kAttnWarps * 32, smem, to_stream(s)>>>(qkv, ld, probs, dout, lddo, dqkv, lddq, B, L, H, dh, scale,
                                                                            drop_p, seed_dev, salt);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_attn_pool_fwd(const float* z, const float* w, float* lin, int64_t ld_lin, int accumulate, int64_t B, int64_t n,
                                    cdcmdr_stream_t s) {
  if (B <= 0) return 0;
  CDC_REQUIRE(z && w && lin && n >= 1 && ld_lin >= 1, "bad attention head operands");
  attn_pool_fwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, to_stream(s)>>>(z, w, lin, ld_lin, accumulate, B, n);
  CDC_LAUNCHED();
  return 0;
}

extern "C" size_t cdcmdr_attn_pool_scratch_bytes(int64_t B, int64_t n) {
  return (size_t)pool_chunks(B) * (size_t)(n > 0 ? n : 1) * sizeof(double);
}

extern "C" int cdcmdr_attn_pool_bwd(const float* z, const float* w, const float* dlin, int64_t ld_dlin, float* dz, float* dw, int64_t B,
                                    int64_t n, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(n >= 1, "bad attention head width");
  CDC_REQUIRE(dw && scratch, "bad attention head operands");
  if (B <= 0) {
    CDC_CHECK(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), to_stream(s)));
    return 0;
  }
  CDC_REQUIRE(z && w && dlin && dz && ld_dlin >= 1, "bad attention head operands");
  const int chunks = pool_chunks(B);
  const int64_t rows_per_chunk = ceil_div(B, chunks);
  attn_pool_bwd_kernel<<<dim3((unsigned)ceil_div(n, 32), (unsigned)chunks), 32 * kPoolRowsPerBlock, 0, to_stream(s)>>>(
      z, w, dlin, ld_dlin, dz, (double*)scratch, B, n, rows_per_chunk);
  CDC_LAUNCHED();
  attn_pool_bwd_finalize_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, to_stream(s)>>>((const double*)scratch, chunks, n, dw);
  CDC_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ bf16 token path entry points
extern "C" int cdcmdr_attn_fwd_bf16(const uint16_t* qkv, int64_t ld, uint16_t* out, int64_t ldo, int64_t B, int L, int H, int dh, float scale,
                                    float drop_p, const uint64_t* seed_dev, uint32_t salt, cdcmdr_stream_t s) {
  CDC_REQUIRE(L >= 1 && L <= 32, "attention runs over at most 32 field tokens");
  CDC_REQUIRE(H >= 1 && dh >= 2 && dh <= 256 && dh % 2 == 0, "bf16 attention needs an even head width");
  CDC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev), "bad attention dropout");
  if (B <= 0) return 0;
  CDC_REQUIRE(qkv && out && ld >= 3 * (int64_t)H * dh && ldo >= (int64_t)H * dh && ld % 2 == 0 && ((uintptr_t)qkv & 3) == 0,
              "bad attention operands");
  if (attn_mma_ok(dh, ld, ldo, H) && ((uintptr_t)out & 3) == 0) {
    const int grid = attn_grid(B * H);
    if (dh == 16) attn_fwd_mma_kernel<16><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, out, ldo, B, L, H, scale, drop_p, seed_dev, salt);
    else if (dh == 32) attn_fwd_mma_kernel<32><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, out, ldo, B, L, H, scale, drop_p, seed_dev, salt);
    else attn_fwd_mma_kernel<64><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, out, ldo, B, L, H, scale, drop_p, seed_dev, salt);
    CDC_LAUNCHED();
    return 0;
  }
  const size_t smem = (size_t)kAttnWarps * (3 * (size_t)L * (dh + 1) + (size_t)L * L) * sizeof(float);
  CDC_REQUIRE(smem <= 200 * 1024, "attention tile does not fit shared memory");
  CDC_CHECK(cudaFuncSetAttribute(attn_fwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_bf16_kernel<<<attn_grid(B * H), kAttnWarps * 32, smem, to_stream(s)>>>(qkv, ld, out, ldo, B, L, H, dh, scale, drop_p, seed_dev, salt);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_attn_bwd_bf16(const uint16_t* qkv, int64_t ld, const uint16_t* dout, int64_t lddo, uint16_t* dqkv, int64_t lddq, int64_t B,
                                    int L, int H, int dh, float scale, float drop_p, const uint64_t* seed_dev, uint32_t salt,
                                    cdcmdr_stream_t s) {
  CDC_REQUIRE(L >= 1 && L <= 32, "attention runs over at most 32 field tokens");
  CDC_REQUIRE(H >= 1 && dh >= 2 && dh <= 256 && dh % 2 == 0, "bf16 attention needs an even head width");
  CDC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev), "bad attention dropout");
  if (B <= 0) return 0;
  CDC_REQUIRE(qkv && dout && dqkv && ld >= 3 * (int64_t)H * dh && lddq >= 3 * (int64_t)H * dh && lddo >= (int64_t)H * dh && ld % 2 == 0 &&
              lddo % 2 == 0 && ((uintptr_t)qkv & 3) == 0 && ((uintptr_t)dout & 3) == 0, "bad attention operands");
  if (attn_mma_ok(dh, ld, lddo, H) && lddq % 2 == 0 && ((uintptr_t)dqkv & 3) == 0) {
    const int grid = attn_grid(B * H);
    if (dh == 16) attn_bwd_mma_kernel<16><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, dout, lddo, dqkv, lddq, B, L, H, scale, drop_p, seed_dev, salt);
    else if (dh == 32) attn_bwd_mma_kernel<32><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, dout, lddo, dqkv, lddq, B, L, H, scale, drop_p, seed_dev, salt);
    else attn_bwd_mma_kernel<64><<<grid, 128, 0, to_stream(s)>>>(qkv, ld, dout, lddo, dqkv, lddq, B, L, H, scale, drop_p, seed_dev, salt);
    CDC_LAUNCHED();
    return 0;
  }
  const size_t smem = (size_t)kAttnWarps * (4 * (size_t)L * (dh + 1) + 2 * (size_t)L * L) * sizeof(float);
  CDC_REQUIRE(smem <= 200 * 1024, "attention tile does not fit shared memory");
  CDC_CHECK(cudaFuncSetAttribute(attn_bwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_bf16_kernel<<<attn_grid(B * H), kAttnWarps * 32, smem, to_stream(s)>>>(qkv, ld, dout, lddo, dqkv, lddq, B, L, H, dh, scale, drop_p,
                                                                                 seed_dev, salt);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_attn_pool_fwd_bf16(const uint16_t* z, const float* w, float* lin, int64_t ld_lin, int accumulate, int64_t B, int64_t n,
                                         cdcmdr_stream_t s) {
  if (B <= 0) return 0;
  CDC_REQUIRE(z && w && lin && n >= 1 && ld_lin >= 1, "bad attention head operands");
  attn_pool_fwd_bf16_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, to_stream(s)>>>(z, w, lin, ld_lin, accumulate, B, n);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_attn_pool_bwd_bf16(const uint16_t* z, const float* w, const float* dlin, int64_t ld_dlin, uint16_t* dz, float* dw,
                                         int64_t B, int64_t n, void* scratch, cdcmdr_stream_t s) {
  CDC_REQUIRE(n >= 1, "bad attention head width");
  CDC_REQUIRE(dw && scratch, "bad attention head operands");
  if (B <= 0) {
    CDC_CHECK(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), to_stream(s)));
    return 0;
  }
  CDC_REQUIRE(z && w && dlin && dz && ld_dlin >= 1, "bad attention head operands");
  const int chunks = pool_chunks(B);
  const int64_t rows_per_chunk = ceil_div(B, chunks);
  attn_pool_bwd_bf16_kernel<<<dim3((unsigned)ceil_div(n, 32), (unsigned)chunks), 32 * kPoolRowsPerBlock, 0, to_stream(s)>>>(
      z, w, dlin, ld_dlin, dz, (double*)scratch, B, n, rows_per_chunk);
  CDC_LAUNCHED();
  attn_pool_bwd_finalize_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, to_stream(s)>>>((const double*)scratch, chunks, n, dw);
  CDC_LAUNCHED();
  return 0;
}
