// Flat-buffer optimizer kernels: fused AdamW (device-side grad scale / NaN-skip), Σg², moment pruning with an
// exact histogram (radix) quantile.  Reference behaviour: torch.optim.AdamW (torchrun_main.py:666, 814),
// clip_grad_norm_ (:805-808), training_utils.py:150-170, 354-361 (random / magnitude pruning of Adam moments).
#include "common.cuh"
#include "kernels.h"

namespace rb {

template <typename T>
__device__ __forceinline__ float ld_f(const T* p, long long i);
template <>
__device__ __forceinline__ float ld_f<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ld_f<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T>
__device__ __forceinline__ void st_f(T* p, long long i, float v);
template <>
__device__ __forceinline__ void st_f<float>(float* p, long long i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void st_f<bf16>(bf16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// ============================================================================================ AdamW
// 8 elements per thread per iteration; params are bf16; grads / moments are bf16 or fp32.
template <typename GT, typename ST>
__global__ void __launch_bounds__(256) adamw_kernel(bf16* __restrict__ p, const GT* __restrict__ g, ST* __restrict__ m, ST* __restrict__ v,
                                                    long long n, float lr, float b1, float b2, float eps, float wd, float bc1_inv,
                                                    float bc2_rsqrt, const float* __restrict__ gscale_ptr, float gscale_host,
                                                    const float* __restrict__ skip_ptr, const float* __restrict__ step_dev) {
  if (skip_ptr != nullptr && *skip_ptr != 0.f) return;
  if (step_dev != nullptr) {  // device-resident step count (advances only on updates that are not skipped)
    const float t = fmaxf(*step_dev, 1.f);
    bc1_inv = 1.f / (1.f - powf(b1, t));
    bc2_rsqrt = 1.f / sqrtf(1.f - powf(b2, t));
  }
  const float gs = gscale_host * (gscale_ptr ? *gscale_ptr : 1.f);
  if (!(fabsf(gs) < INFINITY)) return;  // non-finite gradient norm (the scale is poisoned by the caller): skip the update
  const float decay = 1.f - lr * wd;
  const float step_size = lr * bc1_inv;
  const long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float pf[8];
    unpack8(reinterpret_cast<const bf16x8*>(p)[i], pf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long e = i * 8 + j;
      const float gr = ld_f(g, e) * gs;
      const float mm = b1 * ld_f(m, e) + (1.f - b1) * gr;
      const float vv = b2 * ld_f(v, e) + (1.f - b2) * gr * gr;
      st_f(m, e, mm);
      st_f(v, e, vv);
      const float denom = sqrtf(vv) * bc2_rsqrt + eps;
      pf[j] = pf[j] * decay - step_size * (mm / denom);
    }
    reinterpret_cast<bf16x8*>(p)[i] = pack8(pf);
  }
}

void adamw_flat(void* param, const void* grad, bool grad_f32, void* exp_avg, void* exp_avg_sq, bool state_f32, long long n, float lr,
                float beta1, float beta2, float eps, float weight_decay, int step, const float* grad_scale, float grad_scale_host,
                const float* skip, const float* step_dev, cudaStream_t s) {
  if (n % 8) throw std::runtime_error("adamw: n must be a multiple of 8");
  const float bc1_inv = 1.f / (1.f - powf(beta1, (float)step));
  const float bc2_rsqrt = 1.f / sqrtf(1.f - powf(beta2, (float)step));
  const int grid = (int)std::min<long long>((n / 8 + 255) / 256, (long long)num_sms() * 8);
  bf16* p = (bf16*)param;
#define LAUNCH(GT, ST)                                                                                                         \
  adamw_kernel<GT, ST><<<grid, 256, 0, s>>>(p, (const GT*)grad, (ST*)exp_avg, (ST*)exp_avg_sq, n, lr, beta1, beta2, eps,       \
                                            weight_decay, bc1_inv, bc2_rsqrt, grad_scale, grad_scale_host, skip, step_dev)
  if (grad_f32 && state_f32) LAUNCH(float, float);
  else if (grad_f32) LAUNCH(float, bf16);
  else if (state_f32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
  RB_CHECK_LAUNCH("adamw_flat");
}

// ============================================================================================ Σ x²
// Deterministic two-stage reduction (block partials in a fixed order, no atomics): replicas that hold identical
// gradients must compute bit-identical norms, otherwise their clip coefficients — and then their parameters — drift.
constexpr int kSumsqBlocks = 1024;
template <typename T>
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const T* __restrict__ x, long long n, float* __restrict__ partials) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float f = ld_f(x, i);
    acc += f * f;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int nb, float* __restrict__ out) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partials[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) *out += acc;
}
void sumsq(const void* x, bool is_f32, long long n, float* out, cudaStream_t s) {
  static float* partials[16] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (partials[dev & 15] == nullptr) check(cudaMalloc(&partials[dev & 15], kSumsqBlocks * sizeof(float)), "cudaMalloc(sumsq partials)");
  float* ws = partials[dev & 15];
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)kSumsqBlocks);
  if (is_f32) sumsq_partial_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, ws);
  else sumsq_partial_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, n, ws);
  RB_CHECK_LAUNCH("sumsq_partial");
  sumsq_final_kernel<<<1, 256, 0, s>>>(ws, grid, out);
  RB_CHECK_LAUNCH("sumsq_final");
}

// ============================================================================================ pruning
template <typename T>
__global__ void __launch_bounds__(256) random_prune_kernel(T* __restrict__ x, long long n, uint32_t thr24, uint32_t seed, long long col0) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!keep_bit(seed, 0u, (uint32_t)(i + col0), thr24)) st_f(x, i, 0.f);
}
void random_prune(void* x, bool is_f32, long long n, float ratio, uint32_t seed, long long col_offset, cudaStream_t s) {
  const uint32_t thr = (uint32_t)llroundf(ratio * 16777216.f);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  if (is_f32) random_prune_kernel<float><<<grid, 256, 0, s>>>((float*)x, n, thr, seed, col_offset);
  else random_prune_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)x, n, thr, seed, col_offset);
  RB_CHECK_LAUNCH("random_prune");
}

template <typename T>
__global__ void __launch_bounds__(256) threshold_prune_kernel(T* __restrict__ x, long long n, const float* __restrict__ thr) {
  const float t = *thr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(fabsf(ld_f(x, i)) > t)) st_f(x, i, 0.f);
}
void threshold_prune(void* x, bool is_f32, long long n, const float* thr, cudaStream_t s) {
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  if (is_f32) threshold_prune_kernel<float><<<grid, 256, 0, s>>>((float*)x, n, thr);
  else threshold_prune_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)x, n, thr);
  RB_CHECK_LAUNCH("threshold_prune");
}

// Exact quantile of |x| by two-level radix select on the fp32 bit pattern (|x| >= 0 so bit order == value order):
//   pass 1: histogram of the top 16 bits                       -> bin holding rank k (for k = lo and k = hi)
//   pass 2: histogram of the low 16 bits inside those bins     -> exact order statistics v_lo, v_hi
//   thr = v_lo + frac * (v_hi - v_lo)   (torch.quantile 'linear'), rounded to bf16 when the data is bf16
struct QuantileWs {
  unsigned int hist1[65536];
  unsigned int hist2[2][65536];
  unsigned int bin[2];
  unsigned int rank_in_bin[2];
};
size_t magnitude_quantile_workspace_bytes() { return sizeof(QuantileWs); }

template <typename T>
__global__ void __launch_bounds__(256) hist1_kernel(const T* __restrict__ x, long long n, QuantileWs* ws) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&ws->hist1[__float_as_uint(fabsf(ld_f(x, i))) >> 16], 1u);
}
__global__ void __launch_bounds__(1024) scan1_kernel(QuantileWs* ws, long long k_lo, long long k_hi) {
  // single block: each thread owns 64 consecutive bins
  __shared__ unsigned long long partial[1024];
  unsigned long long local = 0;
  const int b0 = threadIdx.x * 64;
  for (int b = 0; b < 64; ++b) local += ws->hist1[b0 + b];
  partial[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; ++t) {
      unsigned long long c = partial[t];
      partial[t] = run;
      run += c;
    }
  }
  __syncthreads();
  unsigned long long run = partial[threadIdx.x];
  for (int b = 0; b < 64; ++b) {
    const unsigned long long c = ws->hist1[b0 + b];
    for (int q = 0; q < 2; ++q) {
      const unsigned long long k = q == 0 ? (unsigned long long)k_lo : (unsigned long long)k_hi;
      if (k >= run && k < run + c) {
        ws->bin[q] = b0 + b;
        ws->rank_in_bin[q] = (unsigned int)(k - run);
      }
    }
    run += c;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) hist2_kernel(const T* __restrict__ x, long long n, QuantileWs* ws) {
  const unsigned int b_lo = ws->bin[0], b_hi = ws->bin[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int u = __float_as_uint(fabsf(ld_f(x, i)));
    const unsigned int top = u >> 16;
    if (top == b_lo) atomicAdd(&ws->hist2[0][u & 0xFFFFu], 1u);
    if (top == b_hi) atomicAdd(&ws->hist2[1][u & 0xFFFFu], 1u);
  }
}
__global__ void __launch_bounds__(1024) scan2_kernel(QuantileWs* ws, float frac, int round_bf16, float* thr) {
  __shared__ unsigned int partial[1024];
  __shared__ float vals[2];
  for (int q = 0; q < 2; ++q) {
    unsigned int local = 0;
    const int b0 = threadIdx.x * 64;
    for (int b = 0; b < 64; ++b) local += ws->hist2[q][b0 + b];
    partial[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int run = 0;
      for (int t = 0; t < 1024; ++t) {
        unsigned int c = partial[t];
        partial[t] = run;
        run += c;
      }
    }
    __syncthreads();
    unsigned int run = partial[threadIdx.x];
    const unsigned int k = ws->rank_in_bin[q];
    for (int b = 0; b < 64; ++b) {
      const unsigned int c = ws->hist2[q][b0 + b];
      if (k >= run && k < run + c) vals[q] = __uint_as_float((ws->bin[q] << 16) | (unsigned int)(b0 + b));
      run += c;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float t = vals[0] + frac * (vals[1] - vals[0]);
    if (round_bf16) t = bf16_round(t);
    *thr = t;
  }
}

void magnitude_quantile(const void* x, bool is_f32, long long n, float ratio, float* thr, void* workspace, cudaStream_t s) {
  QuantileWs* ws = reinterpret_cast<QuantileWs*>(workspace);
  check(cudaMemsetAsync(ws, 0, sizeof(QuantileWs), s), "memset(quantile ws)");
  const double pos = (double)ratio * (double)(n - 1);
  const long long lo = (long long)pos;
  const long long hi = std::min<long long>(lo + 1, n - 1);
  const float frac = (float)(pos - (double)lo);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  if (is_f32) hist1_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, ws);
  else hist1_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, n, ws);
  RB_CHECK_LAUNCH("quantile_hist1");
  scan1_kernel<<<1, 1024, 0, s>>>(ws, lo, hi);
  RB_CHECK_LAUNCH("quantile_scan1");
  if (is_f32) hist2_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, ws);
  else hist2_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)x, n, ws);
  RB_CHECK_LAUNCH("quantile_hist2");
  scan2_kernel<<<1, 1024, 0, s>>>(ws, frac, is_f32 ? 0 : 1, thr);
  RB_CHECK_LAUNCH("quantile_scan2");
}

}  // namespace rb
