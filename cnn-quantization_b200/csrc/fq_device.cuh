// fqb200 device-side building blocks (sm_100a).  Header-only, included by fqb200.cu.
//
// Design notes:
//  * the hot path is HBM-bound elementwise + reduction work on fp32 NCHW tensors: no tensor cores.
//  * every global access to the tensor is a 128-bit vector (scalar only when H*W is not a multiple of 4),
//    issued UNROLL-deep per thread so ~64 KB per SM are in flight.
//  * one persistent cooperative kernel walks the tensor up to three times (statistics, deviations, apply);
//    phases are separated by a hand-written grid barrier whose last-arriving CTA runs the O(C) parameter solve.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fqb {

constexpr int kThreads = 512;           // CTA size of every kernel here
constexpr int kWarps = kThreads / 32;
constexpr int kCtasPerSm = 2;           // __launch_bounds__(512, 2): <= 64 registers per thread
constexpr int kUnroll = 4;              // independent 128-bit loads in flight per thread

// ------------------------------------------------------------------------------------------------
// global memory access
// ------------------------------------------------------------------------------------------------
// Tensor reads: coherent (the buffer may be written in place later in the same launch), no L1 allocation
// (each byte is used once per phase; L2 - 126 MB - is what carries reuse between phases).
__device__ __forceinline__ float4 ld_tensor(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_tensor(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Tensor writes: streaming (evict-first) so the output does not push the input out of L2.
__device__ __forceinline__ void st_tensor(float4* p, const float4& v) { __stcs(p, v); }
__device__ __forceinline__ void st_tensor(float* p, float v) { __stcs(p, v); }

// Workspace traffic between CTAs goes through L2 only.
template <typename T>
__device__ __forceinline__ T ld_ws(const T* p) { return __ldcg(p); }
template <typename T>
__device__ __forceinline__ void st_ws(T* p, T v) { __stcg(p, v); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// NaN-propagating min / max (torch.clamp / torch.min / torch.max semantics) and C fminf/fmaxf
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// ------------------------------------------------------------------------------------------------
// exact fp32 division by a CTA-uniform divisor
// ------------------------------------------------------------------------------------------------
// q = RN(x / s) with r = RN(1/s) precomputed once per item: one FMUL + two FFMA instead of the ~10
// instruction MUFU.RCP/Newton/FCHK sequence nvcc emits per element.  (Markstein: with r the correctly
// rounded reciprocal and the remainder formed exactly by FMA, the corrected quotient is the correctly
// rounded one; tests/test_gpu_parity.py::test_division_is_ieee checks it against __fdiv_rn on 2^28 pairs.)
// Outside the exponent window where the remainder is exact (huge / inf / NaN quotients) the value is
// clamped away by the caller, so the uncorrected product is returned.
struct Divisor {
  float s;     // divisor
  float r;     // RN(1/s)
  bool fast;   // s in a range where the 3-instruction sequence is exact
};
__device__ __forceinline__ Divisor make_divisor(float s) {
  Divisor d;
  d.s = s;
  d.r = __frcp_rn(s);
  float a = fabsf(s);
  d.fast = (a > 1e-30f) && (a < 1e30f);
  return d;
}
__device__ __forceinline__ float div_exact(float x, const Divisor& d) {
  if (!d.fast) return __fdiv_rn(x, d.s);  // CTA-uniform branch
  float q0 = __fmul_rn(x, d.r);
  float rem = __fmaf_rn(-q0, d.s, x);
  float q = __fmaf_rn(rem, d.r, q0);
  // |q0| tiny: rem may have underflowed, but then |x/s| << 0.5 and any value that small rounds the same way
  // after "+ zero_point"; |q0| huge or non-finite: clamped by the caller.
  return (fabsf(q0) < 1e30f) ? q : q0;
}

// round-half-even of t in [0, 2^22): two full-rate FADDs instead of FRND
__device__ __forceinline__ float rint_small_nonneg(float t) {
  const float magic = 8388608.0f;  // 2^23
  return __fsub_rn(__fadd_rn(t, magic), magic);
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
struct OpAdd {
  template <typename T>
  __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMin {
  __device__ __forceinline__ float operator()(float a, float b) const { return fminf(a, b); }
};
struct OpMax {
  __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};

// Block-wide reduction, result broadcast to every thread; fixed combination order (deterministic).
// `scratch` holds kWarps elements of T in shared memory.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch) {
  v = warp_reduce(v, op);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = scratch[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) r = op(r, scratch[w]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// grid barrier with a leader section
// ------------------------------------------------------------------------------------------------
// Three words in the workspace, zero between launches.  `arrive` counts CTA arrivals monotonically
// (epoch e completes at e*gridDim.x), `release` publishes the last completed epoch, `exited` lets the
// last CTA to leave zero all three again.  Co-residency comes from cudaLaunchCooperativeKernel.
struct GridSync {
  unsigned arrive;
  unsigned release;
  unsigned exited;
  unsigned pad;
};

// All threads of all CTAs call this.  Returns true in exactly one CTA (the last to arrive) WITHOUT
// waiting: that CTA runs the serial section and then calls grid_release().  Every other CTA returns
// false only after the leader has released the epoch.
__device__ __forceinline__ bool grid_arrive(GridSync* gs, unsigned& epoch, int* sh_flag) {
  ++epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // publish this CTA's workspace writes (cumulative over the bar.sync above)
    unsigned prev = atomicAdd(&gs->arrive, 1u);
    *sh_flag = (prev + 1u == epoch * gridDim.x) ? 1 : 0;
  }
  __syncthreads();
  const bool lead = (*sh_flag != 0);
  if (lead) {
    __threadfence();  // acquire side: see every other CTA's writes
  } else {
    if (threadIdx.x == 0) {
      while (ld_acquire_u32(&gs->release) < epoch) __nanosleep(100);
      __threadfence();
    }
    __syncthreads();
  }
  return lead;
}
__device__ __forceinline__ void grid_release(GridSync* gs, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release_u32(&gs->release, epoch);
  }
}
__device__ __forceinline__ void grid_exit(GridSync* gs) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(&gs->exited, 1u);
    if (prev + 1u == gridDim.x) {  // everyone is past the last wait: safe to re-arm
      gs->arrive = 0u;
      gs->release = 0u;
      __threadfence();
      gs->exited = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// geometry: the tensor as [outer][groups][inner], walked per (group, part) item
// ------------------------------------------------------------------------------------------------
struct Geometry {
  unsigned groups;          // G
  unsigned parts;           // P: items per group
  unsigned inner_v;         // inner / VEC
  unsigned step_q, step_r;  // kThreads / inner_v, kThreads % inner_v
  unsigned long long group_v;    // outer * inner_v: vectors per group
  unsigned long long row_pitch;  // groups * inner_v: vectors between consecutive outer slices of a group
  unsigned long long items;      // G * P
};

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

// Walk item (g, p): thread t visits vectors t, t+S, t+2S, ... of the item's range inside group g and calls
// body(value, offset_in_vectors).  Loads are issued kUnroll at a time before any is consumed.
template <int VEC, typename Body>
__device__ __forceinline__ void walk_item(const Geometry& geo, const float* __restrict__ base, unsigned long long item,
                                          Body&& body) {
  using V = typename VecT<VEC>::type;
  const V* src = reinterpret_cast<const V*>(base);
  const unsigned g = static_cast<unsigned>(item % geo.groups);
  const unsigned p = static_cast<unsigned>(item / geo.groups);
  const unsigned long long vb = (geo.group_v * p) / geo.parts;
  const unsigned long long ve = (geo.group_v * (p + 1ull)) / geo.parts;
  const unsigned len = static_cast<unsigned>(ve - vb);
  unsigned k = threadIdx.x;
  if (k >= len) return;
  // cursor of the first vector of this thread: row a, position j within the row
  const unsigned long long v0 = vb + k;
  const unsigned long long a0 = v0 / geo.inner_v;
  unsigned j = static_cast<unsigned>(v0 - a0 * geo.inner_v);
  unsigned long long off = a0 * geo.row_pitch + static_cast<unsigned long long>(g) * geo.inner_v + j;
  const unsigned long long adv = static_cast<unsigned long long>(geo.step_q) * geo.row_pitch + geo.step_r;
  const unsigned long long wrap = geo.row_pitch - geo.inner_v;
  auto advance = [&]() {
    j += geo.step_r;
    off += adv;
    if (j >= geo.inner_v) {
      j -= geo.inner_v;
      off += wrap;
    }
  };
  constexpr unsigned S = kThreads;
  while (k + (kUnroll - 1) * S < len) {
    unsigned long long o[kUnroll];
    V x[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      o[u] = off;
      advance();
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) x[u] = ld_tensor(src + o[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) body(x[u], o[u]);
    k += kUnroll * S;
  }
  while (k < len) {
    V x = ld_tensor(src + off);
    body(x, off);
    advance();
    k += S;
  }
}

}  // namespace fqb
