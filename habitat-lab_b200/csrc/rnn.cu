// hb200 -- masked LSTM recurrence, one launch per (layer, time step), full fp32.
//   h_in = h_{t-1} * m_t ; c_in = c_{t-1} * m_t          (mask resets the state BEFORE the step,
//   gates = xproj_t + h_in W_hh^T                          HB/rl/models/rnn_state_encoder.py:301-316)
//   i,f,g,o = sig,sig,tanh,sig ; c = f*c_in + i*g ; h = o*tanh(c)
// The input projection (x W_ih^T + b_ih + b_hh) for ALL T*N frames is one hb200_sgemm call; only
// the truly sequential h W_hh^T part lives here.  This replaces the PackedSequence index
// machinery (rnn_state_encoder.py:35-277): a masked recurrence needs nothing but `masks`.
#include "common.cuh"

namespace hb200 {
void count_launch(int n);

constexpr int kUnits = 4;  // hidden units per block -> 16 gate rows

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// grid = H / kUnits blocks, 256 threads (8 warps).  smem: W rows [16][H].
template <int NJ>  // H = 32 * NJ
__global__ void __launch_bounds__(256)
lstm_step_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh,
                     const float* __restrict__ b_hh, const uint8_t* __restrict__ masks, const float* __restrict__ h_prev,
                     long long hp_stride, const float* __restrict__ c_prev, long long cp_stride,
                     float* __restrict__ h, float* __restrict__ c, float* __restrict__ gates_out, int n) {
  constexpr int H = NJ * 32;
  extern __shared__ float sw[];  // [16][H]
  const int u0 = blockIdx.x * kUnits;
  for (int i = threadIdx.x; i < 16 * H; i += blockDim.x) {
    const int r = i / H, k = i - r * H;
    const int gate = r >> 2, u = r & 3;  // row r = gate*4 + u
    sw[i] = w_hh[((size_t)gate * H + u0 + u) * H + k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < n; s += 8) {
    const float m = masks[s] ? 1.f : 0.f;
    float hv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) hv[j] = h_prev[(size_t)s * hp_stride + lane + 32 * j] * m;
    float dot[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc = fmaf(hv[j], sw[r * H + lane + 32 * j], acc);
      dot[r] = warp_sum(acc);
    }
    // lanes 0..3 finish one hidden unit each (static register indexing via select)
    float gi = 0, gf = 0, gg = 0, go = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane == u) { gi = dot[u]; gf = dot[4 + u]; gg = dot[8 + u]; go = dot[12 + u]; }
    if (lane < kUnits) {
      const int col = u0 + lane;
      const float* xp = xproj + (size_t)s * 4 * H;
      if (b_hh) { gi += b_hh[col]; gf += b_hh[H + col]; gg += b_hh[2 * H + col]; go += b_hh[3 * H + col]; }
      const float i_ = sigmoidf_(gi + xp[col]);
      const float f_ = sigmoidf_(gf + xp[H + col]);
      const float g_ = tanhf(gg + xp[2 * H + col]);
      const float o_ = sigmoidf_(go + xp[3 * H + col]);
      const float cin = c_prev[(size_t)s * cp_stride + col] * m;
      const float cn = f_ * cin + i_ * g_;
      const float hn = o_ * tanhf(cn);
      c[(size_t)s * H + col] = cn;
      h[(size_t)s * H + col] = hn;
      if (gates_out) {
        float* go_ = gates_out + (size_t)s * 4 * H;
        go_[col] = i_; go_[H + col] = f_; go_[2 * H + col] = g_; go_[3 * H + col] = o_;
      }
    }
  }
}

// pointwise part of the backward step; also zeroes dh_prev for the matmul kernel's atomics
__global__ void lstm_step_bwd_pointwise_kernel(const float* __restrict__ dh_out, const float* __restrict__ dh_rec,
                                               const float* __restrict__ dc_rec, const float* __restrict__ gates,
                                               const float* __restrict__ c, const float* __restrict__ c_prev,
                                               long long cp_stride, const uint8_t* __restrict__ masks,
                                               float* __restrict__ dgates, float* __restrict__ dh_prev,
                                               float* __restrict__ dc_prev, int n, int H) {
  const int total = n * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i / H, k = i - s * H;
    const float m = masks[s] ? 1.f : 0.f;
    const float* gt = gates + (size_t)s * 4 * H;
    const float i_ = gt[k], f_ = gt[H + k], g_ = gt[2 * H + k], o_ = gt[3 * H + k];
    float dh = dh_out ? dh_out[i] : 0.f;
    if (dh_rec) dh += dh_rec[i];
    const float tc = tanhf(c[i]);
    float dc = dh * o_ * (1.f - tc * tc);
    if (dc_rec) dc += dc_rec[i];
    const float cin = c_prev[(size_t)s * cp_stride + k] * m;
    float* dg = dgates + (size_t)s * 4 * H;
    dg[k] = dc * g_ * i_ * (1.f - i_);
    dg[H + k] = dc * cin * f_ * (1.f - f_);
    dg[2 * H + k] = dc * i_ * (1.f - g_ * g_);
    dg[3 * H + k] = dh * tc * o_ * (1.f - o_);
    dc_prev[i] = dc * f_ * m;
    dh_prev[i] = 0.f;
  }
}

// dh_prev[s,k] += m_s * sum_{r in slab} dgates[s,r] * W_hh[r,k];  grid (H/32, R/rows_per_block)
template <int NS>  // sequences per warp
__global__ void __launch_bounds__(256)
lstm_step_bwd_matmul_kernel(const float* __restrict__ dgates, const float* __restrict__ w_hh,
                            const uint8_t* __restrict__ masks, float* __restrict__ dh_prev, int n, int H,
                            int rows_per_block) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = r0 + rows_per_block;
  const int R = 4 * H;
  for (int sb = warp * NS; sb < n; sb += 8 * NS) {
    float acc[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) acc[q] = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float w = w_hh[(size_t)r * H + k];
#pragma unroll
      for (int q = 0; q < NS; ++q) {
        const int s = sb + q;
        if (s < n) acc[q] = fmaf(dgates[(size_t)s * R + r], w, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int s = sb + q;
      if (s < n && masks[s]) atomicAdd(&dh_prev[(size_t)s * H + k], acc[q]);
    }
  }
}
}  // namespace hb200

using namespace hb200;

extern "C" int hb200_lstm_step_fwd(const float* xproj, const float* w_hh, const float* b_hh, const uint8_t* masks,
                                   const float* h_prev, long long h_prev_stride, const float* c_prev,
                                   long long c_prev_stride, float* h, float* c, float* gates_out, int n,
                                   int hidden, hb200_stream_t stream) {
  HB_CHECK_ARG(xproj && w_hh && masks && h_prev && c_prev && h && c && n > 0, "lstm_step_fwd: bad args");
  HB_CHECK_ARG(hidden % 32 == 0 && hidden >= 32 && hidden <= 512 && hidden % kUnits == 0,
               "lstm_step_fwd: hidden=%d unsupported (multiple of 32, <= 512)", hidden);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * 16 * hidden;
  const int grid = hidden / kUnits;
#define HB_LSTM(NJ)                                                                                       \
  {                                                                                                       \
    auto kern = lstm_step_fwd_kernel<NJ>;                                                                 \
    if (smem > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 256, smem, st>>>(xproj, w_hh, b_hh, masks, h_prev, h_prev_stride, c_prev, c_prev_stride, h, c,  \
                                  gates_out, n);                                                          \
  }
  switch (hidden / 32) {
    case 1: HB_LSTM(1); break;
    case 2: HB_LSTM(2); break;
    case 4: HB_LSTM(4); break;
    case 8: HB_LSTM(8); break;
    case 16: HB_LSTM(16); break;
    default:
      set_last_error("lstm_step_fwd: hidden=%d unsupported (32,64,128,256,512)", hidden);
      return HB200_ERR_UNSUPPORTED;
  }
#undef HB_LSTM
  HB_LAUNCH_OK();
  count_launch(1);
  return HB200_OK;
}

extern "C" int hb200_lstm_step_bwd(const float* dh_out, const float* dh_rec, const float* dc_rec,
                                   const float* gates, const float* c, const float* c_prev,
                                   long long c_prev_stride, const float* w_hh, const uint8_t* masks,
                                   float* dgates, float* dh_prev, float* dc_prev, int n, int hidden,
                                   hb200_stream_t stream) {
  HB_CHECK_ARG(gates && c && c_prev && w_hh && masks && dgates && dh_prev && dc_prev && n > 0,
               "lstm_step_bwd: bad args");
  HB_CHECK_ARG(hidden % 32 == 0, "lstm_step_bwd: hidden must be a multiple of 32");
  cudaStream_t st = (cudaStream_t)stream;
  const int total = n * hidden;
  lstm_step_bwd_pointwise_kernel<<<cdiv(total, 256), 256, 0, st>>>(dh_out, dh_rec, dc_rec, gates, c, c_prev,
                                                                   c_prev_stride, masks, dgates, dh_prev,
                                                                   dc_prev, n, hidden);
  HB_LAUNCH_OK();
  const int R = 4 * hidden;
  int rpb = 256;
  while (R % rpb) rpb >>= 1;
  dim3 grid(hidden / 32, R / rpb);
  lstm_step_bwd_matmul_kernel<4><<<grid, 256, 0, st>>>(dgates, w_hh, masks, dh_prev, n, hidden, rpb);
  HB_LAUNCH_OK();
  count_launch(2);
  return HB200_OK;
}
