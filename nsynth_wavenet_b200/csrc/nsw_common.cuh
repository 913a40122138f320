// Shared host/device helpers for libnsw_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/nsw.h"

namespace nsw {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define NSW_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ::nsw::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                       cudaGetErrorString(e_));                                           \
      return NSW_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define NSW_CHECK(cond, code, ...)                                                        \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::nsw::set_error(__VA_ARGS__);                                                      \
      return (code);                                                                      \
    }                                                                                     \
  } while (0)

#define NSW_TRY(expr)                                                                     \
  do {                                                                                    \
    int rc_ = (expr);                                                                     \
    if (rc_ != NSW_OK) return rc_;                                                        \
  } while (0)

// ---- named-tensor lookup (TF variable names; EMA-shadow suffix tolerated) ----
class TensorMap {
 public:
  TensorMap(const nsw_tensor* t, int n);
  // returns NULL (and sets the error string) if missing or if numel mismatches
  const float* get(const std::string& name, int64_t expect_numel) const;
  bool has(const std::string& name) const { return map_.count(name) != 0; }

 private:
  std::map<std::string, const nsw_tensor*> map_;
};

// ---- device buffer that frees itself ----
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int ensure(size_t n) {
    if (n <= bytes) return NSW_OK;
    release();
    NSW_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return NSW_OK;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

int upload(DevBuf& buf, const void* host, size_t bytes);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel, device), whatever
// device is current, and safely from several threads (`done` is one bit per device ordinal).
int ensure_dynamic_smem(const void* kernel, int bytes, std::atomic<uint64_t>& done);

// ---- fp16-range guard of the split-precision (fp16 hi + fp16 lo) engines ------------------------------------
// Every activation that becomes a tensor-core operand is rounded to fp16 (max 65504): a value beyond that turns
// into inf and the result into garbage WITHOUT any fault (fminf / fmaxf in the epilogues even swallow the NaNs).
// Each conversion site therefore tracks the largest magnitude it split and, when it is out of range, raises one
// flag word in mapped pinned host memory (zero cost on the host: it is read after a synchronise that happens
// anyway).  The *_host entry points and nsw_range_status() turn it into NSW_ERANGE.
typedef int (*range_setter_fn)(unsigned int* device_visible_flag);
void register_range_setter(range_setter_fn fn);
int range_guard_init(int device);  // idempotent per device; called by every *_create
// after a synchronise: NSW_OK, or NSW_ERANGE (and the flag is cleared) if any split saw |v| > 65504 or a NaN
int range_check(const char* who);

// deconv layer geometry shared by the IAF and fastgen handles
//   out[o] = sum_q x[m-q] * K[r + q*s],  o + p = m*s + r   (masked.py:235-291)
struct DeconvGeom {
  int k, s, cin, cout, ntaps, p;
  int a_off;  // first input frame of GEMM row m is m + a_off
  int mclip(int L) const { return (s * L - 1 + p) / s + 1; }  // GEMM rows per clip
};

#ifdef __CUDACC__
// ------------------------------ device helpers ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// packed 2-wide fp32 FMA (SASS FFMA2): d.{x,y} += a.{x,y} * s
__device__ __forceinline__ void ffma2_s(unsigned long long& d, unsigned long long a, float s) {
  unsigned long long b;
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(s));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ unsigned long long lds_u64(const void* p) {
  unsigned long long r;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(smem_u32(p)));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  uint32_t sz = valid ? 16u : 0u;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float softplusf_acc(float v) {
  // log(1 + e^v), stable on both tails (tf.nn.softplus)
  return fmaxf(v, 0.0f) + log1pf(expf(-fabsf(v)));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NSW_ACT_LEAKY_RELU) return v >= 0.0f ? v : 0.4f * v;
  if (act == NSW_ACT_RELU) return fmaxf(v, 0.0f);
  return tanhf(v);
}

static __device__ unsigned int* nsw_tu_range_ptr = nullptr;  // one instance per translation unit and device
__device__ __forceinline__ void range_track(uint32_t& mx, float v) {
#ifndef NSW_NO_RANGE_GUARD  // (timing experiments only)
  mx = max(mx, __float_as_uint(v) & 0x7fffffffu);  // NaN and inf compare above every finite magnitude
#endif
}
// one FMNMX per value for the epilogues that sit on a kernel's critical chain (the residual stream of the flow / layer
// kernels).  fmaxf ignores NaNs, which is enough there: every value those epilogues split is a finite combination of
// values that passed an integer-tracked site (start conv, deconv / conditioning epilogues) unless some term overflowed
// to inf first, and inf is caught.
__device__ __forceinline__ void range_track_fast(float& mx, float v) {
#ifndef NSW_NO_RANGE_GUARD
  mx = fmaxf(mx, fabsf(v));
#endif
}
__device__ __forceinline__ void range_commit(uint32_t mx);
__device__ __forceinline__ void range_commit_fast(float mx) { range_commit(__float_as_uint(mx)); }
__device__ __forceinline__ void range_commit(uint32_t mx) {
  if (mx > 0x477FE000u /* 65504.0f */) {
    unsigned int* p = nsw_tu_range_ptr;
    if (p) *p = 1u;
  }
}
// one per .cu that splits activations: hands that unit's kernels the flag word
#define NSW_RANGE_GUARD_TU(tag)                                                                       \
  namespace {                                                                                         \
  int range_setter_##tag(unsigned int* flag) {                                                        \
    NSW_CUDA(cudaMemcpyToSymbol(::nsw::nsw_tu_range_ptr, &flag, sizeof(flag)));                              \
    return NSW_OK;                                                                                    \
  }                                                                                                   \
  struct RangeReg_##tag {                                                                             \
    RangeReg_##tag() { ::nsw::register_range_setter(&range_setter_##tag); }                           \
  } range_reg_##tag;                                                                                  \
  }

// integer code q in [-Q/2, Q/2) -> audio: utils.inv_cast_quantize (utils.py:157-159, :167-169) or, with mu-law,
// utils.inv_mu_law (utils.py:108-122, :125-139)
__device__ __forceinline__ float inv_quant_dev(float q, float Q, int use_mu_law) {
  if (!use_mu_law) return q / (Q * 0.5f);
  const float mu = 255.0f;
  float out = (q + 0.5f) * 2.0f / (mu + 1.0f);
  const float sgn = out > 0.0f ? 1.0f : (out < 0.0f ? -1.0f : 0.0f);
  out = sgn / mu * (powf(1.0f + mu, fabsf(out)) - 1.0f);
  return q == 0.0f ? q : out;
}
// ParallelWavenet._clip_quant_scale (parallel_wavenet.py:348-359)
__device__ __forceinline__ float clip_quant_scale_dev(float x, float Q, int use_mu_law) {
  x = fminf(fmaxf(x, -1.0f), 1.0f - 2.0f / Q);
  return inv_quant_dev(floorf(x * Q * 0.5f) /* utils.cast_quantize, utils.py:142-154 */, Q, use_mu_law);
}

// utils.mu_law (utils.py:72-87) scaled the way Wavenet / Fastgen feed it (wavenet.py:411-414):
// floor(sign(x) * log(1 + 255|x|) / log(256) * 128) / (Q/2)
__device__ __forceinline__ float mu_law_scaled_dev(float x, float Q) {
  const float sgn = x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f);
  const float out = sgn * logf(1.0f + 255.0f * fabsf(x)) / 5.545177444479562f;  // np.log(256)
  return floorf(out * 128.0f) / (Q * 0.5f);
}

// Philox4x32-10 (counter-based; one call gives 4 x 32 random bits)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// uniform in [1e-5, 1 - 1e-5] like tf.random_uniform(minval=1e-5, maxval=1-1e-5)
__device__ __forceinline__ float u01_clipped(uint32_t r) {
  float u = (static_cast<float>(r >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  return 1e-5f + u * (1.0f - 2e-5f);
}
#endif  // __CUDACC__

}  // namespace nsw
