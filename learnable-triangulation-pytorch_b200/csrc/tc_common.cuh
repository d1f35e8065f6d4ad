// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, conv_tc_fold.cu): mbarrier, TMA, TMEM, UMMA.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace lt {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a CONVERGED warp (elect.sync): keeps the surrounding control flow warp-uniform so that the compiler
// keeps descriptors / loop state in uniform registers instead of serialising through divergence waterfalls
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (error returned to the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int kEpiThreads = 256;   // 8 epilogue warps: two per TMEM lane quadrant, each taking 16 of a block's 32 channels
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// byte offset of 16-byte chunk c16 of row `row` inside a [rows][128 B] tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int c16) { return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 contiguous bytes (16 fp16 values of one voxel row) in one request
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// 32 contiguous bytes through the read-only path (two 16-byte requests)
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// two floats -> packed split-fp16 (hi pair, lo pair); element `a` lands in the low half-word (lower address)
__device__ __forceinline__ void split_s32x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(b), "f"(a));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
  const float ra = (a - hf.x) * kLoScale, rb = (b - hf.y) * kLoScale;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(rb), "f"(ra));
}

// 16 channels (half `half` of a 32-channel block) of staged row `srow`: [rows][128 B] tile, 128B swizzle.
// split-fp16: hi halves in chunks 0..3, lo halves in chunks 4..7 (8 channels per chunk); fp32: 4 channels per chunk.
__device__ __forceinline__ void epi_load16(uint32_t base, int srow, int half, int out_format, float (&r)[16]) {
  if (out_format == LT_FMT_F32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 q = lds128(base + sw128_off(srow, half * 4 + c));
      r[c * 4] = __uint_as_float(q.x); r[c * 4 + 1] = __uint_as_float(q.y);
      r[c * 4 + 2] = __uint_as_float(q.z); r[c * 4 + 3] = __uint_as_float(q.w);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 qh = lds128(base + sw128_off(srow, half * 2 + c)), ql = lds128(base + sw128_off(srow, 4 + half * 2 + c));
      const __half2* hh = reinterpret_cast<const __half2*>(&qh);
      const __half2* ll = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(hh[e]), b = __half22float2(ll[e]);
        r[c * 8 + e * 2] = fmaf(b.x, kLoInv, a.x);
        r[c * 8 + e * 2 + 1] = fmaf(b.y, kLoInv, a.y);
      }
    }
  }
}
__device__ __forceinline__ void epi_store16(uint32_t base, int srow, int half, int out_format, const float (&v)[16]) {
  if (out_format == LT_FMT_F32) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(base + sw128_off(srow, half * 4 + c), make_uint4(__float_as_uint(v[c * 4]), __float_as_uint(v[c * 4 + 1]),
                                                               __float_as_uint(v[c * 4 + 2]), __float_as_uint(v[c * 4 + 3])));
  } else {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint4 qh, ql;
      split_s32x2(v[c * 8], v[c * 8 + 1], qh.x, ql.x);
      split_s32x2(v[c * 8 + 2], v[c * 8 + 3], qh.y, ql.y);
      split_s32x2(v[c * 8 + 4], v[c * 8 + 5], qh.z, ql.z);
      split_s32x2(v[c * 8 + 6], v[c * 8 + 7], qh.w, ql.w);
      sts128(base + sw128_off(srow, half * 2 + c), qh);
      sts128(base + sw128_off(srow, 4 + half * 2 + c), ql);
    }
  }
}
// scale/shift, residual, ReLU on 16 channels starting at channel index co
__device__ __forceinline__ void epi_affine16(float (&v)[16], const float* __restrict__ scale, const float* __restrict__ shift, int co) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + co + j));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + co + j));
    v[j] = fmaf(v[j], sc.x, sh.x); v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
    v[j + 2] = fmaf(v[j + 2], sc.z, sh.z); v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
  }
}
__device__ __forceinline__ void epi_activate16(float (&v)[16], const float (&r)[16], int residual, int relu) {
  if (residual == LT_RES_BEFORE_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += r[j];
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (residual == LT_RES_AFTER_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += r[j];
  }
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms
// 1024 bytes apart): start address >> 4 | LBO=1 | SBO=1024>>4 | version=1 | layout=SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, 64-byte-swizzled descriptor (rows of 64 bytes, 8-row atoms 512 bytes apart): layout = SWIZZLE_64B (4)
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=fp16 (format 0), both K-major, M=128, N=n
__host__ __device__ inline uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


// ---- host helpers (conv_tc.cu) ----
int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, const uint32_t* estrides, int swizzle128, int f32 = 0);   // swizzle128: 1 = 128B, 2 = 64B, 0 = none

}  // namespace lt
