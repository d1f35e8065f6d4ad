// PTX wrappers for CTA-pair (cta_group::2) tcgen05 kernels: cluster addressing, pair-wide TMEM allocation, M = 256 MMAs issued by
// the leader CTA, multicast commits, TMA loads that signal a barrier in the peer CTA.  Used by conv_pair.cu and conv_fold_pair.cu.
#pragma once
#include "tc_common.cuh"

namespace lt {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// default semantics (.release.cta): a single SYNCS.ARRIVE.  The .release.cluster form costs MEMBAR.ALL.GPU + ERRBAR, i.e. it waits
// for every outstanding global store of the thread (the direct-store epilogue has 64 sectors in flight per lane); what has to be
// ordered here is the tcgen05.ld of the accumulator, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already do.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// scale/shift (16 channels from co) out of shared memory: broadcast LDS instead of LDG behind the store traffic
__device__ __forceinline__ void epi_affine16_smem(float (&v)[16], const float* scale, const float* shift, int co) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 sc = *reinterpret_cast<const float4*>(scale + co + j);
    const float4 sh = *reinterpret_cast<const float4*>(shift + co + j);
    v[j] = fmaf(v[j], sc.x, sh.x); v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
    v[j + 2] = fmaf(v[j + 2], sc.z, sh.z); v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs have completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the peer (leader) CTA
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// kind::f16 instruction descriptor for the pair: D = f32, A = B = fp16, K-major, M = 256, N = n
__device__ __forceinline__ uint32_t make_idesc_f16_m256(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

}  // namespace lt
