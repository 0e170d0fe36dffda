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
// FAST is the hoisted, CTA-uniform `d.fast`: the caller picks the loop version once per segment.
template <bool FAST>
__device__ __forceinline__ float div_exact(float x, const Divisor& d) {
  if (!FAST) return __fdiv_rn(x, d.s);
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
// geometry: the tensor as [outer][groups][inner]; slabs x balanced contiguous chunks
// ------------------------------------------------------------------------------------------------
// The vectors of every group are enumerated in (outer, inner) order: v in [0, group_v).  They are cut into P
// "slabs" (slab p = v in [group_v*p/P, group_v*(p+1)/P), i.e. a range of the batch dimension); a slab is one
// contiguous region of memory.  Inside slab p the index space (g, v) - group-major - is split into gridDim.x equal
// contiguous chunks, one per CTA: every CTA streams the same number of bytes whatever C, N, H*W are, in at most a
// couple of long segments (a chunk is cut only at group boundaries).  All CTAs work on the same slab at the same
// time, so the union of their accesses stays inside a few hundred MB (DRAM-page / TLB friendly on GB-sized tensors).
// Each (slab, CTA, group) segment yields one partial at slot p * (grid + G) + c + g (unique: consecutive CTAs share
// at most one group).
struct Geometry {
  unsigned groups;          // G
  unsigned slabs;           // P
  unsigned inner_v;         // inner / VEC
  unsigned step_q, step_r;  // kThreads / inner_v, kThreads % inner_v
  unsigned red_lanes;       // lanes per group in the leader's partial reductions (power of two <= 32)
  unsigned long long group_v;    // outer * inner_v: vectors per group
  unsigned long long row_pitch;  // groups * inner_v: vectors between consecutive outer slices of a group
};

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

struct Slab {
  unsigned long long v0;   // first vector (within a group) of the slab
  unsigned long long len;  // vectors per group in the slab
  unsigned long long tot;  // groups * len
};
__device__ __forceinline__ Slab slab_of(const Geometry& geo, unsigned p) {
  Slab s;
  s.v0 = geo.group_v * p / geo.slabs;
  s.len = geo.group_v * (p + 1ull) / geo.slabs - s.v0;
  s.tot = s.len * geo.groups;
  return s;
}
// the CTA whose chunk of the slab contains slab-local index w
__device__ __forceinline__ unsigned cta_of(const Slab& s, unsigned long long w) {
  return static_cast<unsigned>(((w + 1ull) * gridDim.x - 1ull) / s.tot);
}
__device__ __forceinline__ size_t partial_slot(const Geometry& geo, unsigned p, unsigned c, unsigned g) {
  return static_cast<size_t>(p) * (gridDim.x + geo.groups) + c + g;
}

// Visit the segments (g, slab p, first vector within the group, length) of this CTA; `reverse` walks the slabs and
// the segments inside a chunk back to front.
template <typename Fn>
__device__ __forceinline__ void for_each_segment(const Geometry& geo, bool reverse, Fn&& fn) {
  for (unsigned i = 0; i < geo.slabs; ++i) {
    const unsigned p = reverse ? geo.slabs - 1u - i : i;
    const Slab s = slab_of(geo, p);
    const unsigned long long w0 = s.tot * blockIdx.x / gridDim.x, w1 = s.tot * (blockIdx.x + 1ull) / gridDim.x;
    if (w1 <= w0) continue;
    const unsigned g_first = static_cast<unsigned>(w0 / s.len);
    const unsigned g_last = static_cast<unsigned>((w1 - 1ull) / s.len);
    for (unsigned k = 0; k <= g_last - g_first; ++k) {
      const unsigned g = reverse ? g_last - k : g_first + k;
      const unsigned long long gbase = static_cast<unsigned long long>(g) * s.len;
      const unsigned long long vb = (g == g_first) ? w0 - gbase : 0ull;
      const unsigned long long ve = (g == g_last) ? w1 - gbase : s.len;
      fn(g, p, s.v0 + vb, static_cast<unsigned>(ve - vb));
    }
  }
}

// Walk `len` vectors of group g starting at vector vb (REV: from the last one backwards, so that a phase walked in
// the opposite direction starts on what the previous phase touched last and still finds it in L2).  Thread t visits
// vectors t, t+S, t+2S, ...; U independent loads are issued (bounds-predicated) before any is consumed.
// The (row, column) cursor advances by adds only: one 64-bit division per segment.  Offsets are 32-bit vector
// indices from the tensor base (the host refuses tensors of 2^32 vectors = 64 GB and more): registers are what
// limits the number of loads in flight.
template <int VEC, int U, bool REV, typename Body>
__device__ __forceinline__ void walk_segment(const Geometry& geo, const float* base, unsigned g, unsigned long long vb,
                                             unsigned len, Body&& body) {
  using V = typename VecT<VEC>::type;
  const V* src = reinterpret_cast<const V*>(base);
  if (threadIdx.x >= len) return;
  const unsigned long long v0 = REV ? vb + (len - 1u - threadIdx.x) : vb + threadIdx.x;
  const unsigned long long a0 = v0 / geo.inner_v;
  unsigned j = static_cast<unsigned>(v0 - a0 * geo.inner_v);
  unsigned off = static_cast<unsigned>(a0 * geo.row_pitch + static_cast<unsigned long long>(g) * geo.inner_v + j);
  const unsigned adv = static_cast<unsigned>(geo.step_q * geo.row_pitch + geo.step_r);
  const unsigned wrap = static_cast<unsigned>(geo.row_pitch - geo.inner_v);
  const unsigned inner_v = geo.inner_v, step_r = geo.step_r;
  auto advance = [&]() {
    if (!REV) {
      j += step_r;
      off += adv;
      if (j >= inner_v) {
        j -= inner_v;
        off += wrap;
      }
    } else {
      off -= adv;
      if (j < step_r) {
        j += inner_v;
        off -= wrap;
      }
      j -= step_r;
    }
  };
  constexpr unsigned S = kThreads;
  // vectors left for this thread's lane of the stride-S sequence: ceil((len - tid) / S)
  unsigned left = (len - threadIdx.x + S - 1u) / S;
  while (left >= static_cast<unsigned>(U)) {  // full batches: no predicates
    unsigned o[U];
    V x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      o[u] = off;
      x[u] = ld_tensor(src + off);
      advance();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) body(x[u], o[u]);
    left -= U;
  }
  if (left) {  // one predicated batch for the tail
    unsigned o[U];
    V x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      o[u] = off;
      if (static_cast<unsigned>(u) < left) x[u] = ld_tensor(src + off);
      advance();
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (static_cast<unsigned>(u) < left) body(x[u], o[u]);
  }
}

// ------------------------------------------------------------------------------------------------
// asynchronous variant of walk_segment for the 128-bit path: a per-thread ring of cp.async (LDGSTS) copies
// ------------------------------------------------------------------------------------------------
// Registers cap how many loads a thread can keep in flight (64 registers -> 4 x 128 bit).  cp.async writes the
// loaded vector straight into shared memory, so the ring can be D deep at no register cost: with D = 8 and 32 warps
// per SM, 128 KB per SM are in flight - enough to cover HBM latency at full bandwidth.  Every thread reads back only
// its own slots, so no CTA barrier is involved; completion is tracked with cp.async groups (one per vector, in order).
constexpr int kRingDepth = 8;
constexpr int kRingBytes = kRingDepth * kThreads * 16;  // dynamic shared memory per CTA

__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(unsigned smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr) : "memory");
  return v;
}

template <bool REV, typename Body>
__device__ __forceinline__ void walk_segment_async(const Geometry& geo, const float* base, unsigned g, unsigned long long vb,
                                                   unsigned len, unsigned ring /* shared-space address of this thread's slot 0 */,
                                                   Body&& body) {
  constexpr int D = kRingDepth;
  constexpr unsigned S = kThreads;
  const float4* src = reinterpret_cast<const float4*>(base);
  if (threadIdx.x >= len) return;
  const unsigned long long v0 = REV ? vb + (len - 1u - threadIdx.x) : vb + threadIdx.x;
  const unsigned long long a0 = v0 / geo.inner_v;
  unsigned j = static_cast<unsigned>(v0 - a0 * geo.inner_v);
  unsigned off = static_cast<unsigned>(a0 * geo.row_pitch + static_cast<unsigned long long>(g) * geo.inner_v + j);
  const unsigned adv = static_cast<unsigned>(geo.step_q * geo.row_pitch + geo.step_r);
  const unsigned wrap = static_cast<unsigned>(geo.row_pitch - geo.inner_v);
  const unsigned inner_v = geo.inner_v, step_r = geo.step_r;
  auto advance = [&]() {
    if (!REV) {
      j += step_r;
      off += adv;
      if (j >= inner_v) {
        j -= inner_v;
        off += wrap;
      }
    } else {
      off -= adv;
      if (j < step_r) {
        j += inner_v;
        off -= wrap;
      }
      j -= step_r;
    }
  };
  const unsigned left = (len - threadIdx.x + S - 1u) / S;  // vectors this thread visits
  unsigned o[D];                                            // offsets of the vectors in flight (stores need them)
  // prologue: fill the ring (always D groups so that the group arithmetic below is uniform)
#pragma unroll
  for (int d = 0; d < D; ++d) {
    o[d] = off;
    if (static_cast<unsigned>(d) < left) {
      cp_async16(ring + d * (S * 16u), src + off);
      advance();
    }
    cp_async_commit();
  }
  unsigned i = 0;
  // steady state: D vectors per trip; slot d holds vector i + d
  while (i + D <= left) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      cp_async_wait<D - 1>();  // the oldest group (this slot) has landed
      const float4 x = lds128(ring + d * (S * 16u));
      const unsigned od = o[d];
      o[d] = off;
      if (i + D + d < left) {  // refill the slot with the vector D ahead
        cp_async16(ring + d * (S * 16u), src + off);
        advance();
      }
      cp_async_commit();
      body(x, od);
    }
    i += D;
  }
  // drain: fewer than D left, all already in flight
  cp_async_wait<0>();
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (i + d < left) {
      const float4 x = lds128(ring + d * (S * 16u));
      body(x, o[d]);
    }
  }
}

}  // namespace fqb
