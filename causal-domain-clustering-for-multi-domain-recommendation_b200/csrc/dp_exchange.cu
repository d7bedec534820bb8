// the embedding exchange of the data-parallel step, fused into the kernels that produce / consume the exchanged bytes,
// over NVLink peer memory.  The table is sharded at field boundaries (owner o holds the rows of fields [fbound[o], fbound[o+1]));
// every rank maps every other rank's exchange buffers (torch's symmetric-memory allocator supplies the mappings - plumbing).
//
//   forward    cdcmdr_dp_push_ids      my [B, F] indices, column block of owner o  ->  o's recv_ids[(rank*B + b), j]      (peer stores)
//              cdcmdr_peer_barrier
//              cdcmdr_dp_gather_push   owner: table row of every received index -> the REQUESTER's X[b, (f0+j)*E ...]     (local reads,
//              cdcmdr_peer_barrier              fp32 or bf16 on the way out                                                 peer stores)
//   backward   cdcmdr_dp_push_grads    my dX[b, owner o's columns] (fp32) -> o's grad_recv[(rank*B + b), :] fp32 or bf16  (peer stores)
//              cdcmdr_peer_barrier
//
// which replaces: pack + NCCL all-to-all (ids), gather into a send buffer + NCCL all-to-all (rows) + unpack, cast + NCCL all-to-all
// (row gradients) of the round-1 path - 3 collectives and 8 packing / cast launches per step.  Stores, not loads, cross NVLink: a
// store is posted (no round trip), and the consumer finds the bytes in its own HBM / L2.
//
// cdcmdr_peer_barrier: flags[r] = rank r's block of [n_slots][world] uint64 (zero-initialised, symmetric).  A call on slot s
// release-stores the slot's next sequence number into flags[peer][s][rank] of every peer (system scope: everything this GPU wrote
// before - earlier kernels of the stream included - is visible to whoever acquires the flag) and spins until all `world` entries of
// its own block carry it.  Sequence numbers only grow, so a slot can be reused by the next step without a reset; the counters live
// in device memory (seqs[n_slots], private to the rank), so a captured CUDA graph replays the exchange without the host.
#include "common.cuh"

namespace cdcmdr {

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32)
peer_barrier_kernel(unsigned long long* const* __restrict__ flags, int rank, int world, int slot, unsigned long long* __restrict__ seqs) {
  const unsigned long long seq = seqs[slot] + 1ull;
  __syncwarp();
  __threadfence_system();
  for (int r = threadIdx.x; r < world; r += 32) st_release_sys_u64(flags[r] + (int64_t)slot * world + rank, seq);
  for (int r = threadIdx.x; r < world; r += 32) {
    const unsigned long long* mine = flags[rank] + (int64_t)slot * world + r;
    while (ld_acquire_sys_u64(mine) < seq) { }
  }
  __syncwarp();
  if (threadIdx.x == 0) seqs[slot] = seq;
}

// blockIdx.y = owner.  Contiguous stores into the owner's buffer: element idx = b*nf + j of my B x nf block.
__global__ void __launch_bounds__(256)
dp_push_ids_kernel(const int32_t* __restrict__ x, int64_t B, int F, int32_t* const* __restrict__ recv, const int32_t* __restrict__ fbound,
                   int rank) {
  const int o = blockIdx.y;
  const int f0 = fbound[o], nf = fbound[o + 1] - f0;
  if (nf <= 0) return;
  const int64_t total = B * nf;
  int32_t* dst = recv[o] + (int64_t)rank * total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, j;
    split_idx(i, nf, b, j);
    dst[i] = __ldg(x + b * F + f0 + j);
  }
}

template <typename OutT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void put(float* o, const float4& t) { *reinterpret_cast<float4*>(o) = t; }
};
template <> struct Store4<uint16_t> {
  static __device__ __forceinline__ void put(uint16_t* o, const float4& t) {
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
  }
};

// 16 bytes of the destination per thread, consecutive threads on consecutive chunks: what crosses NVLink are runs of whole
// 32-byte sectors (a sample's nf*E elements are contiguous at the destination), not 8-byte pieces - the first cut of these kernels
// stored 8 bytes per thread at a 32-byte stride and reached 160 GB/s on the remote half.
template <typename OutT> struct Chunk;                        // kEl elements = 16 bytes of OutT
template <> struct Chunk<float> {
  static constexpr int kEl = 4;
  static __device__ __forceinline__ void copy(float* dst, const float* src, bool ok) {
    const float4 t = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(dst) = t;
  }
};
template <> struct Chunk<uint16_t> {
  static constexpr int kEl = 8;
  static __device__ __forceinline__ void copy(uint16_t* dst, const float* src, bool ok) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ok) { a = reinterpret_cast<const float4*>(src)[0]; b = reinterpret_cast<const float4*>(src)[1]; }
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
};

// Owner side.  One thread per 16-byte chunk of a received (requester p, sample b)'s nf*E output elements: the index of the chunk's
// row (re-read by the E/kEl threads of a row: an L1 hit), 16 or 32 bytes of the table row from the local shard, one 16-byte store
// into the requester's X.
template <typename OutT>
__global__ void __launch_bounds__(256)
dp_gather_push_kernel(const int32_t* __restrict__ recv_ids, const int64_t* __restrict__ off_local, const float* __restrict__ shard,
                      int64_t Vl, OutT* const* __restrict__ xs, int64_t ldx, int col0, int64_t B, int nf, int E, int world,
                      int packed, int* __restrict__ oob) {
  constexpr int kEl = Chunk<OutT>::kEl;
  const int cpr = E / kEl;                                    // chunks per row
  const int64_t cps = (int64_t)nf * cpr;                      // chunks per sample
  const int64_t total = (int64_t)world * B * cps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pb, c, p, b;
    split_idx(i, cps, pb, c);
    split_idx(pb, B, p, b);
    const int j = (int)(c / cpr), h = (int)(c - (int64_t)j * cpr);
    const int64_t row = (int64_t)__ldg(recv_ids + pb * nf + j) + __ldg(off_local + j);
    const bool ok = row >= 0 && row < Vl;
    if (!ok && oob) *oob = 1;
    // packed: this owner's [B, nf*E] block is contiguous at the requester (element offset B*col0) - whole warps store 512
    // contiguous bytes; otherwise straight into the requester's [B, ldx] matrix (runs of nf*E elements)
    OutT* dst = packed ? xs[p] + B * col0 + (b * cps + c) * kEl : xs[p] + b * ldx + col0 + c * kEl;
    Chunk<OutT>::copy(dst, shard + (ok ? row : 0) * E + h * kEl, ok);
  }
}

// row-per-thread form for widths a 16-byte bf16 chunk does not divide (E = 4)
template <int EV, typename OutT>
__global__ void __launch_bounds__(256)
dp_gather_push_rows_kernel(const int32_t* __restrict__ recv_ids, const int64_t* __restrict__ off_local, const float* __restrict__ shard,
                           int64_t Vl, OutT* const* __restrict__ xs, int64_t ldx, int col0, int64_t B, int nf, int world,
                           int packed, int* __restrict__ oob) {
  constexpr int E = 4 * EV;
  const int64_t total = (int64_t)world * B * nf;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pb, j, p, b;
    split_idx(i, nf, pb, j);
    split_idx(pb, B, p, b);
    const int64_t row = (int64_t)__ldg(recv_ids + i) + __ldg(off_local + j);
    const bool ok = row >= 0 && row < Vl;
    float4 t[EV];
    const float4* src = reinterpret_cast<const float4*>(shard + (ok ? row : 0) * E);
#pragma unroll
    for (int k = 0; k < EV; ++k) t[k] = src[k];
    if (!ok) {
      if (oob) *oob = 1;
#pragma unroll
      for (int k = 0; k < EV; ++k) t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    OutT* o = packed ? xs[p] + B * col0 + (b * nf + j) * E : xs[p] + b * ldx + col0 + j * E;
#pragma unroll
    for (int k = 0; k < EV; ++k) Store4<OutT>::put(o + 4 * k, t[k]);
  }
}

// blockIdx.y = owner.  My dX[b, f0*E .. f1*E) -> the owner's grad_recv[(rank*B + b), 0 .. nf*E): contiguous at the destination,
// 16 bytes of it per thread.
template <typename OutT>
__global__ void __launch_bounds__(256)
dp_push_grads_kernel(const float* __restrict__ dX, int64_t ldg, int64_t B, int E, OutT* const* __restrict__ grecv,
                     const int32_t* __restrict__ fbound, int rank) {
  constexpr int kEl = Chunk<OutT>::kEl;
  const int o = blockIdx.y;
  const int f0 = fbound[o], nf = fbound[o + 1] - f0;
  if (nf <= 0) return;
  const int64_t q = (int64_t)nf * E / kEl;                     // chunks per sample
  const int64_t total = B * q;
  OutT* dst = grecv[o] + (int64_t)rank * total * kEl;
  const float* src = dX + (int64_t)f0 * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, c;
    split_idx(i, q, b, c);
    Chunk<OutT>::copy(dst + kEl * i, src + b * ldg + c * kEl, true);
  }
}

// narrow form (E = 4 with bf16 output: 8 bytes per thread)
template <typename OutT>
__global__ void __launch_bounds__(256)
dp_push_grads4_kernel(const float* __restrict__ dX, int64_t ldg, int64_t B, int E, OutT* const* __restrict__ grecv,
                      const int32_t* __restrict__ fbound, int rank) {
  const int o = blockIdx.y;
  const int f0 = fbound[o], nf = fbound[o + 1] - f0;
  if (nf <= 0) return;
  const int64_t q = (int64_t)nf * E / 4;
  const int64_t total = B * q;
  OutT* dst = grecv[o] + (int64_t)rank * total * 4;
  const float* src = dX + (int64_t)f0 * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, c;
    split_idx(i, q, b, c);
    Store4<OutT>::put(dst + 4 * i, __ldg(reinterpret_cast<const float4*>(src + b * ldg) + c));
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

extern "C" int cdcmdr_peer_barrier(uint64_t* const* peer_flags, int rank, int world, int slot, int n_slots, uint64_t* seqs,
                                   cdcmdr_stream_t s) {
  CDC_REQUIRE(peer_flags && seqs, "peer barrier: null argument");
  CDC_REQUIRE(world >= 1 && world <= 256 && rank >= 0 && rank < world && slot >= 0 && slot < n_slots, "peer barrier: bad rank / world / slot");
  peer_barrier_kernel<<<1, 32, 0, to_stream(s)>>>(reinterpret_cast<unsigned long long* const*>(peer_flags), rank, world, slot,
                                                  reinterpret_cast<unsigned long long*>(seqs));
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_dp_push_ids(const int32_t* x, int64_t B, int F, int32_t* const* recv_ids, const int32_t* fbound, int rank, int world,
                                  cdcmdr_stream_t s) {
  if (B == 0) return 0;
  CDC_REQUIRE(x && recv_ids && fbound && B > 0 && F > 0 && world >= 1 && rank >= 0 && rank < world, "dp_push_ids: bad arguments");
  dim3 grid((unsigned)grid_for(B * ((F + world - 1) / world + 1), 256, 4), (unsigned)world);
  dp_push_ids_kernel<<<grid, 256, 0, to_stream(s)>>>(x, B, F, recv_ids, fbound, rank);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_dp_gather_push(const int32_t* recv_ids, const int64_t* off_local, const float* shard, int64_t Vl, void* const* xs,
                                     int out_bf16, int64_t ldx, int col0, int64_t B, int nf, int E, int world, int packed,
                                     int* oob_flag, cdcmdr_stream_t s) {
  if (B == 0 || nf == 0) return 0;
  CDC_REQUIRE(recv_ids && off_local && shard && xs && Vl > 0 && B > 0 && nf > 0 && world >= 1, "dp_gather_push: bad arguments");
  CDC_REQUIRE(E == 4 || E == 8 || E == 16 || E == 32 || E == 64, "dp_gather_push: embed_dim must be 4, 8, 16, 32 or 64");
  CDC_REQUIRE((packed || ldx % 4 == 0) && col0 % 4 == 0, "dp_gather_push: misaligned destination");
  cudaStream_t st = to_stream(s);
  if (out_bf16 ? (E % 8 == 0 && (packed || ldx % 8 == 0) && col0 % 8 == 0) : true) {
    const int grid = grid_for((int64_t)world * B * nf * (E / (out_bf16 ? 8 : 4)), 256, 8);
    if (out_bf16) dp_gather_push_kernel<uint16_t><<<grid, 256, 0, st>>>(recv_ids, off_local, shard, Vl, reinterpret_cast<uint16_t* const*>(xs),
                                                                        ldx, col0, B, nf, E, world, packed, oob_flag);
    else dp_gather_push_kernel<float><<<grid, 256, 0, st>>>(recv_ids, off_local, shard, Vl, reinterpret_cast<float* const*>(xs), ldx, col0,
                                                            B, nf, E, world, packed, oob_flag);
    CDC_LAUNCHED();
    return 0;
  }
  const int grid = grid_for((int64_t)world * B * nf, 256, 16);
#define GP(EV)                                                                                                                      \
  dp_gather_push_rows_kernel<EV, uint16_t><<<grid, 256, 0, st>>>(recv_ids, off_local, shard, Vl, reinterpret_cast<uint16_t* const*>(xs), \
                                                                 ldx, col0, B, nf, world, packed, oob_flag)
  switch (E / 4) {
    case 1: GP(1); break; case 2: GP(2); break; case 4: GP(4); break; case 8: GP(8); break; default: GP(16); break;
  }
#undef GP
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_dp_push_grads(const float* dX, int64_t ldg, int64_t B, int F, int E, void* const* grad_recv, int out_bf16,
                                    const int32_t* fbound, int rank, int world, cdcmdr_stream_t s) {
  if (B == 0) return 0;
  CDC_REQUIRE(dX && grad_recv && fbound && B > 0 && F > 0 && world >= 1 && rank >= 0 && rank < world, "dp_push_grads: bad arguments");
  CDC_REQUIRE(E % 4 == 0 && ldg % 4 == 0 && ((uintptr_t)dX % 16) == 0, "dp_push_grads: embed_dim / pitch must be multiples of 4 floats");
  dim3 grid((unsigned)grid_for(B * (((int64_t)(F + world - 1) / world + 1) * E / 4), 256, 8), (unsigned)world);
  cudaStream_t st = to_stream(s);
  if (out_bf16 && E % 8 == 0)
    dp_push_grads_kernel<uint16_t><<<grid, 256, 0, st>>>(dX, ldg, B, E, reinterpret_cast<uint16_t* const*>(grad_recv), fbound, rank);
  else if (out_bf16)
    dp_push_grads4_kernel<uint16_t><<<grid, 256, 0, st>>>(dX, ldg, B, E, reinterpret_cast<uint16_t* const*>(grad_recv), fbound, rank);
  else
    dp_push_grads_kernel<float><<<grid, 256, 0, st>>>(dX, ldg, B, E, reinterpret_cast<float* const*>(grad_recv), fbound, rank);
  CDC_LAUNCHED();
  return 0;
}
