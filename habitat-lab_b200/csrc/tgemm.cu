// hb200 -- TF32 tensor-core GEMM (tcgen05 kind::tf32) on fp32 operands for the dense layers:
// visual_fc, the LSTM input projections and their data / weight gradients.
//
//   C[M,N] (f32) (+)= A[M,K] * B[K,N] (+ bias[N]) (ReLU)
//   A(m,k) = a[m*a_ms + k*a_ks],  B(k,n) = b[k*b_ks + n*b_ns]   (one of the two strides of each == 1)
//
// Precision: operands are fp32 in memory; the tensor core reads them as TF32 (10-bit mantissa,
// the reference's own cuDNN-RNN precision on CUDA), accumulation is fp32 in TMEM.
//
// Operand tiles are copied with 16-byte cp.async (4 floats) straight into the no-swizzle UMMA
// layouts (descriptor encodings pinned by hb200_umma_gemm_probe):
//   K-major  (k contiguous):  vector (row, k4) at k4*rows*16 + row*16          LBO = rows*16, SBO = 128
//   MN-major (mn contiguous): vector (mn4, k)  at (k>>3)*LBO + mn4*128 + (k&7)*16,  SBO = 128 (next 4 mn),
//                             LBO = (rows/4)*128 (next 8 k)
// One MMA covers K = 8 (32 bytes); a chunk is K = 32 (4 MMAs); 3-stage cp.async ring as in conv.cu.
// Weight gradients (K = frames, tiny M x N tile grid) are split over K with fp32 atomics.
#include <cuda.h>
#include "common.cuh"
#include <stdlib.h>
#include <map>
#include <mutex>
#include "umma.cuh"

namespace hb200 {
void count_launch(int n);
using namespace umma;

constexpr int GT_M = 128, GT_K = 32, GT_STAGES = 3;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

struct TgemmArgs {
  const float* a; long long a_ms, a_ks;
  const float* b; long long b_ks, b_ns;
  float* c; long long ldc;
  const float* bias;
  int M, N, K, k_per_split, accumulate, relu;
  // deterministic split-K for skinny problems (the actor's 64-row batches): every split stores its partial tile to
  // ws[split][M][N]; the last CTA of a tile to arrive (ticket counter) sums the splits in order and runs the epilogue
  float* ws;
  int* tickets;
  int vec4;   // c, ldc (and bias) allow 16-byte accesses
};

// load one [ROWS x 32] operand tile (rows = m or n, zero-filled out of range)
template <int MNMAJOR>
__device__ __forceinline__ void tg_load(const float* __restrict__ p, long long s_mn, long long s_k, int mn0, int MN,
                                        int k0, int k_end, uint32_t sdst, int rows) {
  if (!MNMAJOR) {
    // K-major: vector = 4 consecutive k of one row; 8 vectors per row per chunk
    // 128-byte swizzle: 8 consecutive threads fetch the 8 x 16 B of one row (a full 128-byte line) and write one
    // swizzled shared-memory row (conflict free); the row-per-thread mapping of the no-swizzle layout touched 32
    // different lines per request and used half of every 32-byte sector.
    for (int v = threadIdx.x; v < rows * 8; v += 128) {
      const int row = v >> 3, k4 = v & 7;
      const int gm = mn0 + row, gk = k0 + k4 * 4;
      const bool ok = gm < MN && gk + 3 < k_end;
      const float* g = ok ? p + (long long)gm * s_mn + gk : p;
      cp_async16(sdst + (uint32_t)((row >> 3) << 10) + (uint32_t)((row & 7) << 7) + (uint32_t)((k4 ^ (row & 7)) << 4),
                 g, ok);
    }
  } else {
    // MN-major: vector = 4 consecutive mn at one k; rows/4 vectors per k
    const int r4 = rows >> 2;
    for (int v = threadIdx.x; v < r4 * GT_K; v += 128) {
      // 8 consecutive threads = the 8 k rows of one core matrix (128 contiguous smem bytes), the next
      // 8 threads the next 4 mn (adjacent 16 B in global memory)
      const int kl = v & 7, mb = (v >> 3) % r4, k = (v / (8 * r4)) * 8 + kl;
      const int gm = mn0 + mb * 4, gk = k0 + k;
      const bool ok = gm + 3 < MN && gk < k_end;
      const float* g = ok ? p + (long long)gk * s_k + gm : p;
      cp_async16(sdst + (uint32_t)(k >> 3) * (r4 * 128) + (uint32_t)mb * 128 + (uint32_t)(k & 7) * 16, g, ok);
    }
  }
}

// TMEM accumulator tile -> C: shared by the cp.async and the TMA kernels.  Plain launches add bias / accumulate / ReLU
// here; split-K launches either add their partial with fp32 atomics (weight gradients) or park it in the workspace
// for the ticketed, in-order reduction of the skinny path.
template <int BN>
__device__ __forceinline__ void tg_epilogue(const TgemmArgs& a, uint32_t tmem_base, int m0, int n0) {
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int m = m0 + tid;
  if (a.ws) {
    __shared__ int s_last;
    float* wrow = a.ws + ((size_t)blockIdx.z * a.M + m) * a.N + n0;
#pragma unroll 1
    for (int col0 = 0; col0 < BN; col0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + col0, r);
      tmem_ld_wait();
      if (m < a.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (n0 + col0 + j < a.N)   // N % 4 == 0
            __stcg(reinterpret_cast<float4*>(wrow + col0 + j),
                   make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                               __uint_as_float(r[j + 3])));
        }
      }
    }
    fence_before_sync();
    __threadfence();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) s_last = (atomicAdd(&a.tickets[tile], 1) == (int)gridDim.z - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int rows = min(GT_M, a.M - m0), cols4 = min(BN, a.N - n0) >> 2;
    for (int i = tid; i < rows * cols4; i += 128) {
      const int r = i / cols4, c = (i - r * cols4) << 2;
      const size_t off = (size_t)(m0 + r) * a.N + n0 + c;
      float4 acc = __ldcg(reinterpret_cast<const float4*>(a.ws + off));
#pragma unroll 8
      for (int z = 1; z < (int)gridDim.z; ++z) {
        const float4 p = __ldcg(reinterpret_cast<const float4*>(a.ws + (size_t)z * a.M * a.N + off));
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      float v[4] = {acc.x, acc.y, acc.z, acc.w};
      float* dst = a.c + (long long)(m0 + r) * a.ldc + n0 + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (a.bias) v[j] += a.bias[n0 + c + j];
        if (a.accumulate) v[j] += dst[j];
        if (a.relu) v[j] = fmaxf(v[j], 0.f);
        dst[j] = v[j];
      }
    }
    if (tid == 0) a.tickets[tile] = 0;   // ready for the next launch on this stream
    return;
  }
#pragma unroll 1
  for (int col0 = 0; col0 < BN; col0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + col0, r);
    tmem_ld_wait();
    if (m < a.M && gridDim.z == 1 && a.vec4) {
      // one thread owns 128 contiguous bytes of its row: float4 stores (N % 4 == 0: a group of 4 is in or out)
      float* dst = a.c + (long long)m * a.ldc + n0 + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = n0 + col0 + j;
        if (n >= a.N) continue;
        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                               __uint_as_float(r[j + 3]));
        if (a.bias) {
          const float4 b4 = *reinterpret_cast<const float4*>(a.bias + n);
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        }
        if (a.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(dst + j);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(dst + j) = v;
      }
    } else if (m < a.M && a.vec4) {
      // split-K partial: 16-byte reductions (red.global.add.v4.f32), a quarter of the scalar atomics' L2 operations
      float* dst = a.c + (long long)m * a.ldc + n0 + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = n0 + col0 + j;
        if (n >= a.N) continue;
        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                               __uint_as_float(r[j + 3]));
        if (a.bias && blockIdx.z == 0) {
          const float4 b4 = *reinterpret_cast<const float4*>(a.bias + n);
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        }
        atomicAdd(reinterpret_cast<float4*>(dst + j), v);
      }
    } else if (m < a.M) {
      float* dst = a.c + (long long)m * a.ldc + n0 + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + col0 + j;
        if (n >= a.N) continue;
        float v = __uint_as_float(r[j]);
        if (gridDim.z > 1) {
          if (a.bias && blockIdx.z == 0) v += a.bias[n];
          atomicAdd(dst + j, v);
        } else {
          if (a.bias) v += a.bias[n];
          if (a.accumulate) v += dst[j];
          if (a.relu) v = fmaxf(v, 0.f);
          dst[j] = v;
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// NST = cp.async ring depth: 3 for the learner's many-CTA launches, deeper for the skinny (few CTAs, latency-bound) path
template <int BN, int A_MN, int B_MN, int NST = GT_STAGES>
__global__ void __launch_bounds__(128) tgemm_kernel(const TgemmArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar[NST];
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t kABytes = GT_M * GT_K * 4, kBBytes = BN * GT_K * 4, kStage = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * GT_M, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * a.k_per_split;
  const int k_end = min(a.K, k_begin + a.k_per_split);
  const int nchunks = (k_end - k_begin + GT_K - 1) / GT_K;
  if (nchunks <= 0) return;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&mma_bar[s], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc = make_idesc_tf32(GT_M, BN, A_MN, B_MN);

  auto load_chunk = [&](int c, int st) {
    const uint32_t sa = sbase + st * kStage, sb = sa + kABytes;
    const int k0 = k_begin + c * GT_K;
    tg_load<A_MN>(a.a, a.a_ms, a.a_ks, m0, a.M, k0, k_end, sa, GT_M);
    tg_load<B_MN>(a.b, a.b_ns, a.b_ks, n0, a.N, k0, k_end, sb, BN);
  };
#pragma unroll
  for (int c = 0; c < NST - 1; ++c) {
    if (c < nchunks) load_chunk(c, c);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    const int st = c % NST;
    cp_async_wait<NST - 2>();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t sa = sbase + st * kStage, sb = sa + kABytes;
#pragma unroll
      for (int kk = 0; kk < GT_K / 8; ++kk) {
        // K-major: the two 16-byte k4 vectors of this K=8 step are LBO = rows*16 apart
        // MN-major: one 8-k block per step, blocks LBO' = (rows/4)*128 apart; SBO = 128 between 4-mn vectors
        const uint64_t da = A_MN ? make_smem_desc(sa + kk * (GT_M / 4) * 128, (GT_M / 4) * 128, 128, kNoSwizzle)
                                 : make_smem_desc(sa + kk * 32, 16, 1024, kSwizzle128B);
        const uint64_t db = B_MN ? make_smem_desc(sb + kk * (BN / 4) * 128, (BN / 4) * 128, 128, kNoSwizzle)
                                 : make_smem_desc(sb + kk * 32, 16, 1024, kSwizzle128B);
        mma_tf32_ss(tmem_base, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
      }
      mma_commit(&mma_bar[st]);
    }
    const int nc = c + NST - 1;
    if (nc < nchunks) {
      if (c >= 1) mbar_wait(&mma_bar[(c - 1) % NST], ((c - 1) / NST) & 1);
      load_chunk(nc, nc % NST);
    }
    cp_async_commit();
  }
  mbar_wait(&mma_bar[(nchunks - 1) % NST], ((nchunks - 1) / NST) & 1);
  fence_after_sync();

  tg_epilogue<BN>(a, tmem_base, m0, n0);
}

// ---- TMA-fed variant ---------------------------------------------------------------------------------------------------
// The cp.async kernel above spends ~4000 cycles per K chunk on 2048 LDGSTS + their addresses (the tensor core needs 256):
// 54 TFLOP/s on the learner's dense layers.  Here one thread issues two TMA box loads per chunk ([32 k] x [128 | BN rows],
// SWIZZLE_128B = the K-major operand layout, out-of-range rows / k zero-filled by the TMA unit), one thread issues the
// MMAs, and the ring is ordered by full (transaction-count) / empty (tcgen05.commit) barriers.
template <int BN, int NST>
__global__ void __launch_bounds__(128) tgemm_tma_kernel(const TgemmArgs a, const __grid_constant__ CUtensorMap tmap_a,
                                                        const __grid_constant__ CUtensorMap tmap_b) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST], empty_bar[NST], done_bar;
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t kABytes = GT_M * GT_K * 4, kBBytes = BN * GT_K * 4, kStage = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  const CUtensorMap* const pa = &tmap_a;   // param-space addresses, taken in the kernel body
  const CUtensorMap* const pb = &tmap_b;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * GT_M, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * a.k_per_split;
  const int k_end = min(a.K, k_begin + a.k_per_split);
  const int nchunks = (k_end - k_begin + GT_K - 1) / GT_K;
  if (nchunks <= 0) return;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc = make_idesc_tf32(GT_M, BN, 0, 0);
  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % NST;
        if (c >= NST) mbar_wait(&empty_bar[st], ((c / NST) - 1) & 1);
        const uint32_t sa = sbase + st * kStage;
        mbar_expect_tx(&full_bar[st], kStage);
        tma_load_2d(sa, pa, &full_bar[st], k_begin + c * GT_K, m0);
        tma_load_2d(sa + kABytes, pb, &full_bar[st], k_begin + c * GT_K, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % NST;
        mbar_wait(&full_bar[st], (c / NST) & 1);
        fence_after_sync();
        const uint32_t sa = sbase + st * kStage, sb = sa + kABytes;
#pragma unroll
        for (int kk = 0; kk < GT_K / 8; ++kk)
          mma_tf32_ss(tmem_base, make_smem_desc(sa + kk * 32, 16, 1024, kSwizzle128B),
                      make_smem_desc(sb + kk * 32, 16, 1024, kSwizzle128B), idesc, (c > 0 || kk > 0) ? 1u : 0u);
        mma_commit(&empty_bar[st]);
      }
      mma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0);
  fence_after_sync();
  tg_epilogue<BN>(a, tmem_base, m0, n0);
}
}  // namespace hb200

using namespace hb200;

// workspace of the skinny split-K path: one per stream (launches on a stream are ordered; two streams never share one),
// and one per CUDA-graph capture (keyed by the capture id: the graph keeps using its buffers after the capturing stream
// has gone back to the pool).  Under capture the allocation runs in relaxed capture mode -- cudaMalloc is not a stream
// operation -- and the ticket memset becomes the graph's first node (tickets are zero between launches anyway).
static int skinny_workspace(cudaStream_t st, size_t floats, int tiles, float** ws, int** tickets) {
  struct Ws { float* buf = nullptr; size_t cap = 0; int* tickets = nullptr; };
  static std::mutex mu;
  static std::map<unsigned long long, Ws> table;
  constexpr int kTickets = 1024;
  HB_CHECK_ARG(tiles <= kTickets, "tgemm: %d tiles exceed the ticket table", tiles);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  unsigned long long cap_id = 0;
  HB_CUDA(cudaStreamGetCaptureInfo(st, &cs, &cap_id));
  const bool capturing = cs == cudaStreamCaptureStatusActive;
  const unsigned long long key = capturing ? ((1ull << 63) | cap_id) : (unsigned long long)(uintptr_t)st;
  std::lock_guard<std::mutex> lock(mu);
  Ws& w = table[key];
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  if (capturing) HB_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
  cudaError_t e = cudaSuccess;
  if (!w.tickets) {
    e = cudaMalloc(&w.tickets, kTickets * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(w.tickets, 0, kTickets * sizeof(int), st);
  }
  if (e == cudaSuccess && w.cap < floats) {
    if (w.buf && !capturing) {   // earlier launches of a capture keep their pointer: never freed there
      e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) e = cudaFree(w.buf);
    }
    w.buf = nullptr; w.cap = 0;
    const size_t cap = floats < (size_t)(4 << 20) ? (size_t)(4 << 20) : floats;
    if (e == cudaSuccess) e = cudaMalloc(&w.buf, cap * sizeof(float));
    if (e == cudaSuccess) w.cap = cap;
  }
  if (capturing) cudaThreadExchangeStreamCaptureMode(&mode);
  HB_CUDA(e);
  *ws = w.buf; *tickets = w.tickets;
  return HB200_OK;
}

typedef CUresult (*TgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TgEncodeFn tg_encode_fn() {
  static TgEncodeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (TgEncodeFn)p;
  return fn;
}
// row-major fp32 matrix [rows, k] with row pitch ld floats -> boxes of [box_rows] x [32 k] in the 128-byte swizzle
static int tg_tensor_map(CUtensorMap* tm, const float* p, long long rows, long long k, long long ld, int box_rows) {
  TgEncodeFn enc = tg_encode_fn();
  if (!enc) {
    set_last_error("tgemm: cuTensorMapEncodeTiled is not available from this driver");
    return HB200_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)GT_K, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("tgemm: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return HB200_ERR_CUDA;
  }
  return HB200_OK;
}
static int g_tgemm_tma = getenv("HB200_NO_TGEMM_TMA") ? 0 : 1;

template <int BN, int NST>
static int launch_tgemm_tma(const TgemmArgs& g, dim3 grid, cudaStream_t st) {
  CUtensorMap ta, tb;
  int rc = tg_tensor_map(&ta, g.a, g.M, g.K, g.a_ms, GT_M);
  if (rc) return rc;
  rc = tg_tensor_map(&tb, g.b, g.N, g.K, g.b_ns, BN);
  if (rc) return rc;
  const size_t smem = (size_t)NST * (GT_M * GT_K * 4 + BN * GT_K * 4) + 1024;
  auto kern = tgemm_tma_kernel<BN, NST>;
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 128, smem, st>>>(g, ta, tb);
  HB_LAUNCH_OK();
  count_launch(1);
  return HB200_OK;
}

extern "C" int hb200_set_tgemm_tma(int on) { g_tgemm_tma = on ? 1 : 0; return HB200_OK; }
extern "C" int hb200_get_tgemm_tma(void) { return g_tgemm_tma; }

extern "C" int hb200_tgemm(const float* a, long long a_ms, long long a_ks, const float* b, long long b_ks,
                           long long b_ns, float* c, long long ldc, const float* bias, int m, int n, int k,
                           int accumulate, int relu, hb200_stream_t stream) {
  HB_CHECK_ARG(a && b && c && m > 0 && n > 0 && k > 0, "tgemm: bad args");
  // MN-major fp32 operands would need the SWIZZLE_128B_BASE32B layouts (the plain interleaved MN-major
  // descriptor is 16-bit only: verified wrong on hardware) -> K-major only, callers transpose
  HB_CHECK_ARG(a_ks == 1 && b_ks == 1, "tgemm: both operands must be K-major (a_ks == 1, b_ks == 1)");
  const int a_mn = (a_ks == 1) ? 0 : 1;   // k contiguous -> K-major
  const int b_mn = (b_ks == 1) ? 0 : 1;
  // 16-byte vector loads: leading dimensions / pointers must be multiples of 4 floats
  const long long lda = a_mn ? a_ks : a_ms, ldb = b_mn ? b_ks : b_ns;
  HB_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0,
               "tgemm: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  HB_CHECK_ARG(k % 4 == 0 && (!a_mn || m % 4 == 0) && (!b_mn || n % 4 == 0), "tgemm: k (and mn-major extents) must be multiples of 4");
  int BN = n >= 256 ? 256 : (n >= 128 ? 128 : (n >= 64 ? 64 : 32));
  {
    // N tile 128 instead of 256 when that is what fills the 148 SMs (2 CTAs per SM fit at 96 KB of stages)
    static const int forced = getenv("HB200_TGEMM_BN") ? atoi(getenv("HB200_TGEMM_BN")) : 0;
    if (forced == 32 || forced == 64 || forced == 128 || forced == 256) {
      if (forced <= BN) BN = forced;
    } else if (BN == 256 && (long long)cdiv(n, 256) * cdiv(m, GT_M) < 2 * kNumSMs) {
      BN = 128;
    }
  }
  HB_CHECK_ARG(n % 4 == 0, "tgemm: n must be a multiple of 4");
  TgemmArgs g;
  g.a = a; g.a_ms = a_ms; g.a_ks = a_ks; g.b = b; g.b_ks = b_ks; g.b_ns = b_ns; g.c = c; g.ldc = ldc; g.bias = bias;
  g.M = m; g.N = n; g.K = k; g.accumulate = accumulate; g.relu = relu;
  g.ws = nullptr; g.tickets = nullptr;
  g.vec4 = (ldc % 4 == 0 && ((uintptr_t)c & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0)) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= GT_M && k >= 256 && cdiv(n, BN) < kNumSMs / 2 && !a_mn && !b_mn) {
    // one row tile (the actor: 64 frames): a handful of CTAs would each walk the whole K, one DRAM latency per
    // 32-wide chunk (77 us for visual_fc at K = 2048).  32-wide N tiles x K splits of >= 4 chunks put one CTA on every SM,
    // each with a deep cp.async ring; partial tiles meet in an L2-resident workspace and the last CTA of each tile
    // reduces them in split order (8 KB per split).
    BN = 32;
    constexpr int kDeep = 6;   // 6 x (16 KB + 4 KB) stages
    const int nt = cdiv(n, BN);
    int sp = kNumSMs / nt;
    if (sp > k / 128) sp = k / 128;
    if (sp > 32) sp = 32;
    if (sp < 1) sp = 1;
    const int kps = ((cdiv(k, sp) + GT_K - 1) / GT_K) * GT_K;
    sp = cdiv(k, kps);
    if (sp > 1) {
      int rc = skinny_workspace(st, (size_t)sp * m * n, nt, &g.ws, &g.tickets);
      if (rc) return rc;
    }
    g.k_per_split = kps;
    dim3 grid(nt, 1, sp);
    if (g_tgemm_tma) return launch_tgemm_tma<32, kDeep>(g, grid, st);
    const size_t smem = (size_t)kDeep * (GT_M * GT_K * 4 + 32 * GT_K * 4) + 1024;
    auto kern = tgemm_kernel<32, 0, 0, kDeep>;
    HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, st>>>(g);
    HB_LAUNCH_OK();
    count_launch(1);
    return HB200_OK;
  }
  const long long tiles = (long long)cdiv(n, BN) * cdiv(m, GT_M);
  int splits = 1;
  if (accumulate && !relu && tiles < kNumSMs && k >= 1024) {
    splits = (int)((2 * kNumSMs + tiles - 1) / tiles);
    if (splits > k / 256) splits = k / 256;
    if (splits < 1) splits = 1;
  }
  g.k_per_split = ((cdiv(k, splits) + GT_K - 1) / GT_K) * GT_K;
  splits = cdiv(k, g.k_per_split);
  dim3 grid(cdiv(n, BN), cdiv(m, GT_M), splits);
  if (g_tgemm_tma && !a_mn && !b_mn) {
    // ring depth: two CTAs per SM (3 x 32 KB stages) when there are enough tiles for that, else one CTA with a deep ring
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    switch (BN) {
      case 32: return launch_tgemm_tma<32, 4>(g, grid, st);
      case 64: return launch_tgemm_tma<64, 4>(g, grid, st);
      case 128: return ctas > kNumSMs ? launch_tgemm_tma<128, 3>(g, grid, st) : launch_tgemm_tma<128, 6>(g, grid, st);
      default: return launch_tgemm_tma<256, 4>(g, grid, st);
    }
  }
#define HB_TG(bn, AM, BM)                                                                          \
  {                                                                                                \
    const size_t smem = (size_t)GT_STAGES * (GT_M * GT_K * 4 + bn * GT_K * 4) + 1024;              \
    auto kern = tgemm_kernel<bn, AM, BM>;                                                          \
    HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    kern<<<grid, 128, smem, st>>>(g);                                                              \
  }
#define HB_TG_BN(AM, BM)                 \
  switch (BN) {                          \
    case 32: HB_TG(32, AM, BM); break;   \
    case 64: HB_TG(64, AM, BM); break;   \
    case 128: HB_TG(128, AM, BM); break; \
    default: HB_TG(256, AM, BM); break;  \
  }
  if (!a_mn && !b_mn) { HB_TG_BN(0, 0) }
  else if (!a_mn && b_mn) { HB_TG_BN(0, 1) }
  else if (a_mn && !b_mn) { HB_TG_BN(1, 0) }
  else { HB_TG_BN(1, 1) }
#undef HB_TG_BN
#undef HB_TG
  HB_LAUNCH_OK();
  count_launch(1);
  return HB200_OK;
}
