// Micro-probe: tcgen05.mma issue cost for cta_group::1 vs cta_group::2 (SS mode, kind::f16, K = 16 per instruction),
// plus a semantic check of the 2-CTA operand split (which CTA supplies which B rows / D columns).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/cta2_probe tools/cta2_probe.cu && /tmp/cta2_probe
//
// Output: one line per (cta_group, N pattern, grid) with cycles per instruction measured by the issuing thread around a
// chain of back-to-back MMAs that ends in a tcgen05.commit -> mbarrier wait.  The model being tested:
//   cta_group::1:  T = (128 + N) / 2 cycles      (A rows + B rows fetched from shared memory at 64 B/clk, 32 B per row)
//   cta_group::2:  T = (128 + N/2) / 2 cycles    (each SM fetches its 128 A rows and HALF of the B rows), floor N/2.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {   // bounded: false on timeout
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000LL) return false;
  return true;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::f16: D=f32 (bit 4), A=B=fp16, K-major both, M (>>4) at bit 24, N (>>3) at bit 17
__device__ __forceinline__ uint32_t make_idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Result {
  long long cycles;      // issue -> completion of the whole chain (leader CTA)
  int n_instr;
  float d[2][4];         // semantic check: D[row 5][col 0], [col N/2 - 1], [col N/2], [col N - 1] for CTA rank 0 / 1
  int timeout;
};

// pattern: 0 = chain of N=n1 MMAs; 1 = alternate N=n1 / N=n2 (the product's hi*[hi;lo] + lo*hi pair)
template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(int n1, int n2, int pattern, int chain, Result* res) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __half* A = reinterpret_cast<__half*>(smem);              // 128 rows x 64 fp16 (128B swizzle) = 16 KB
  __half* Bm = reinterpret_cast<__half*>(smem + 16384);     // 256 rows x 32 fp16 (64B swizzle)  = 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // constant tiles (swizzle-invariant): A = 1 (rank 0) / 3 (rank 1); B = 1 (rank 0) / 2 (rank 1)
  const __half av = __float2half(rank == 0 ? 1.0f : 3.0f), bv = __float2half(rank == 0 ? 1.0f : 2.0f);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) { A[i] = av; Bm[i] = bv; }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc<CG>(slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;

  long long t0 = 0, t1 = 0;
  int timeout = 0;
  if (rank == 0 && warp == 1) {
    const uint64_t ad = make_sw128_desc(smem_u32(A)), bd = make_sw64_desc(smem_u32(Bm));
    const uint32_t id1 = make_idesc(128 * CG, n1), id2 = make_idesc(128 * CG, n2);
    if (lane == 0) {
      t0 = clock64();
      for (int i = 0; i < chain; ++i) {
        if (pattern == 0) umma<CG>(tmem, ad, bd, id1, i > 0);
        else { umma<CG>(tmem, ad, bd, id1, i > 0); umma<CG>(tmem + 256, ad + 4, bd, id2, i > 0); }
      }
      commit<CG>(bar);
    }
    __syncwarp();
  }
  // everyone (both CTAs) waits for the commit (multicast for CG == 2)
  if (!mbar_wait(bar, 0)) timeout = 1;
  if (rank == 0 && warp == 1 && lane == 0) t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // semantic check (block 0 / cluster 0 only): row 5 of this CTA's accumulator, four columns
  if (blockIdx.x < CG && warp == 0) {
    const int cols[4] = {0, n1 / 2 - 1, n1 / 2, n1 - 1};
    for (int k = 0; k < 4; ++k) {
      uint32_t v;
      tmem_ld1(tmem + (uint32_t)cols[k], v);   // warp 0 -> lanes 0..31; lane 5 = row 5
      if (lane == 5) res->d[rank][k] = __uint_as_float(v);
    }
  }
  if (blockIdx.x == 0 && warp == 1 && lane == 0) { res->cycles = t1 - t0; res->n_instr = chain * (pattern == 0 ? 1 : 2); }
  if (timeout && threadIdx.x == 0) atomicExch(&res->timeout, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) tmem_dealloc<CG>(tmem, 512);
}

template <int CG>
static void run(int n1, int n2, int pattern, int chain, int grid, Result* dres) {
  CK(cudaMemset(dres, 0, sizeof(Result)));
  const size_t smem = 32768 + 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {   // second run is warm
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<CG>, n1, n2, pattern, chain, dres));
    CK(cudaDeviceSynchronize());
  }
  Result r;
  CK(cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost));
  const double per = (double)r.cycles / r.n_instr;
  const double model = pattern == 0 ? (CG == 1 ? (128 + n1) / 2.0 : (128 + n1 / 2.0) / 2.0)
                                    : (CG == 1 ? ((128 + n1) + (128 + n2)) / 4.0 : ((128 + n1 / 2.0) + (128 + n2 / 2.0)) / 4.0);
  printf("cta_group=%d grid=%3d pattern=%d N=%3d/%3d chain=%d: %.1f cyc/instr (model %.1f, math floor %.1f)%s\n", CG, grid, pattern, n1,
         pattern ? n2 : 0, chain, per, model, pattern == 0 ? n1 / 2.0 : (n1 + n2) / 4.0, r.timeout ? "  TIMEOUT" : "");
  if (pattern == 0 && chain == 1)
    printf("   semantic: cta0 D[5][0,N/2-1,N/2,N-1] = %.0f %.0f %.0f %.0f   cta1 = %.0f %.0f %.0f %.0f\n", r.d[0][0], r.d[0][1], r.d[0][2],
           r.d[0][3], r.d[1][0], r.d[1][1], r.d[1][2], r.d[1][3]);
}

// ------------------------------------------------------------------------------------------------
// Mode "ring": the conv kernel's issue pattern with DISTINCT operands per chunk -- a ring of STAGES (A 16 KB + B 16 KB)
// tiles, four MMAs per chunk (A_hi x [B_hi;B_lo] N=2Nt, A_lo x B_hi N=Nt, twice), no TMA (static shared memory), the
// issue loop unrolled by the ring depth.  variant (cta_group::2 only): 0 = second MMA N=Nt into a third accumulator,
// 1 = second MMA N=2Nt ("zero-row" B tile).  Reports cycles per chunk of the issuing CTA.
// ------------------------------------------------------------------------------------------------
template <int CG, int STAGES>
__global__ void __launch_bounds__(128, 1) ring_kernel(int nt, int variant, int chunks, Result* res) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + STAGES * 32768 + 64);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < STAGES * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc<CG>(slot, CG == 2 ? 512 : 256);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  long long t0 = 0, t1 = 0;
  int timeout = 0;
  if (rank == 0 && warp == 1) {
    const int n1 = 2 * nt, n2 = (CG == 2 && variant == 1) ? 2 * nt : nt;
    const uint32_t id1 = make_idesc(128 * CG, n1), id2 = make_idesc(128 * CG, n2);
    const uint32_t d1 = tmem, d2 = (CG == 2 && variant == 0) ? tmem + 2 * nt : tmem + nt;
    if (lane == 0) {
      t0 = clock64();
      for (int q = 0; q < chunks; q += STAGES) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
          const uint32_t a_addr = smem_u32(smem + s * 32768);
          const uint64_t ad = make_sw128_desc(a_addr), bd = make_sw64_desc(a_addr + 16384);
          const uint32_t acc = (q | s) ? 1u : 0u;
          umma<CG>(d1, ad, bd, id1, acc);
          umma<CG>(d2, ad + 4, bd, id2, acc);
          umma<CG>(d1, ad + 2, bd + 2, id1, 1);
          umma<CG>(d2, ad + 6, bd + 2, id2, 1);
        }
      }
      commit<CG>(bar);
    }
    __syncwarp();
  }
  if (!mbar_wait(bar, 0)) timeout = 1;
  if (rank == 0 && warp == 1 && lane == 0) t1 = clock64();
  if (blockIdx.x == 0 && warp == 1 && lane == 0) { res->cycles = t1 - t0; res->n_instr = chunks; }
  if (timeout && threadIdx.x == 0) atomicExch(&res->timeout, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) tmem_dealloc<CG>(tmem, CG == 2 ? 512 : 256);
}

template <int CG, int STAGES>
static void run_ring(int nt, int variant, int grid, int ctas_per_sm, Result* dres) {
  CK(cudaMemset(dres, 0, sizeof(Result)));
  const size_t smem = (size_t)STAGES * 32768 + 128 + 1024;
  CK(cudaFuncSetAttribute(ring_kernel<CG, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int chunks = 256;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, ring_kernel<CG, STAGES>, nt, variant, chunks, dres));
    CK(cudaDeviceSynchronize());
  }
  Result r;
  CK(cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost));
  const int n2 = (CG == 2 && variant == 1) ? 2 * nt : nt;
  printf("ring cta_group=%d stages=%d Nt=%3d (MMA N=%3d + N=%3d) grid=%3d (%d CTA/SM): %.1f cyc/chunk (math floor %.1f; 1-CTA fetch model %.1f)%s\n",
         CG, STAGES, nt, 2 * nt, n2, grid, ctas_per_sm, (double)r.cycles / r.n_instr, 2.0 * (2 * nt + n2) / 2.0,
         2.0 * ((128 + 2 * nt) + (128 + nt)) / 2.0, r.timeout ? "  TIMEOUT" : "");
}

// ------------------------------------------------------------------------------------------------
// Mode "feed": L2 -> shared-memory bandwidth of 1-D TMA bulk copies (16 KB each, DEPTH in flight per CTA) out of a
// window that fits in L2: the ceiling for kernels that re-read operand tiles from L2 (implicit-GEMM taps).
// ------------------------------------------------------------------------------------------------
template <int DEPTH>
__global__ void __launch_bounds__(64, 1) feed_kernel(const uint8_t* buf, size_t window, int copies, unsigned long long* total_cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DEPTH * 16384);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    size_t off = ((size_t)blockIdx.x * 1234567u * 16384u) % window;
    for (int i = 0; i < copies + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) { if (!mbar_wait(&bars[s], ((i / DEPTH) - 1) & 1)) break; }
      if (i < copies) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(16384u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + s * 16384)), "l"(buf + off), "r"(16384u), "r"(smem_u32(&bars[s])) : "memory");
        off += 16384u * 37u;
        if (off >= window) off -= window;
      }
    }
    atomicAdd(total_cycles, (unsigned long long)(clock64() - t0));
  }
}

template <int DEPTH>
static void run_feed(const uint8_t* buf, size_t window, int grid, unsigned long long* dcyc) {
  const size_t smem = (size_t)DEPTH * 16384 + 256 + 1024;
  CK(cudaFuncSetAttribute(feed_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int copies = 2048;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(dcyc, 0, 8));
    CK(cudaEventRecord(e0));
    feed_kernel<DEPTH><<<grid, 64, smem>>>(buf, window, copies, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double bytes = (double)grid * copies * 16384.0;
  printf("feed depth=%d window=%zu MB grid=%3d: %.1f GB/s aggregate (%.1f GB/s per CTA)\n", DEPTH, window >> 20, grid, bytes / best / 1e6,
         bytes / best / 1e6 / grid);
}

// ------------------------------------------------------------------------------------------------
// Mode "issue": what does it cost to ISSUE one chunk (mbarrier wait that is already satisfied + descriptor arithmetic +
// election + four tiny MMAs + commit), for the loop shapes used by the conv kernels?  N = 16 so that the tensor pipe itself
// is never the limit.  The kw-folded kernel was bound by exactly this.
//   style 0: converged warp, runtime ring slot (modulo by increment), one election per iteration        (conv_tc today)
//   style 1: one lane runs the whole loop                                                               (divergent: waterfalls)
//   style 2: converged warp, four iterations unrolled with compile-time slot offsets, one election per four iterations
//   style 3: as 0 with the slot computed by integer modulo of a runtime stage count                     (conv_tc before r01e)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__global__ void __launch_bounds__(128, 1) issue_kernel(int style, int stages, int stage_bytes, int iters, Result* res) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* ready = reinterpret_cast<uint64_t*>(smem + 4 * 32768);     // never armed: waiting for parity 1 returns at once
  uint64_t* sink = ready + 1;                                          // receives the commits
  uint64_t* done = ready + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(ready + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(ready, 1); mbar_init(sink, (1u << 20) - 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc<1>(slot, 64);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t id = make_idesc(128, 16);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    if (style == 1) {
      if (lane == 0) {
        int s = 0;
        for (int q = 0; q < iters; ++q) {
          mbar_wait(ready, 1);
          const uint32_t a = base + (uint32_t)(s * stage_bytes);
          const uint64_t ad = make_sw128_desc(a), bd = make_sw64_desc(a + 16384);
          umma<1>(tmem, ad, bd, id, q > 0); umma<1>(tmem + 16, ad + 4, bd, id, q > 0); umma<1>(tmem, ad + 2, bd + 2, id, 1); umma<1>(tmem + 16, ad + 6, bd + 2, id, 1);
          commit<1>(sink);
          if (++s == stages) s = 0;
        }
      }
      __syncwarp();
    } else if (style == 2) {
      for (int q = 0; q < iters; q += 4) {
        mbar_wait(ready, 1);
        if (elect_one_sync()) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t a = base + (uint32_t)(u * 32768);
            const uint64_t ad = make_sw128_desc(a), bd = make_sw64_desc(a + 16384);
            umma<1>(tmem, ad, bd, id, (q | u) ? 1u : 0u); umma<1>(tmem + 16, ad + 4, bd, id, (q | u) ? 1u : 0u); umma<1>(tmem, ad + 2, bd + 2, id, 1); umma<1>(tmem + 16, ad + 6, bd + 2, id, 1);
            commit<1>(sink);
          }
        }
        __syncwarp();
      }
    } else {
      int s = 0;
      for (int q = 0; q < iters; ++q) {
        if (style == 3) s = q % stages;
        mbar_wait(ready, 1);
        const uint32_t a = base + (uint32_t)(s * stage_bytes);
        const uint64_t ad = make_sw128_desc(a), bd = make_sw64_desc(a + 16384);
        if (elect_one_sync()) {
          umma<1>(tmem, ad, bd, id, q > 0); umma<1>(tmem + 16, ad + 4, bd, id, q > 0); umma<1>(tmem, ad + 2, bd + 2, id, 1); umma<1>(tmem + 16, ad + 6, bd + 2, id, 1);
          commit<1>(sink);
        }
        __syncwarp();
        if (style == 0 && ++s == stages) s = 0;
      }
    }
    if (lane == 0) { commit<1>(done); }
    __syncwarp();
    mbar_wait(done, 0);
    if (lane == 0 && blockIdx.x == 0) { res->cycles = clock64() - t0; res->n_instr = iters; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 64);
}

static void run_issue(int style, Result* dres) {
  CK(cudaMemset(dres, 0, sizeof(Result)));
  const size_t smem = 4 * 32768 + 256 + 1024;
  CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    issue_kernel<<<1, 128, smem>>>(style, 4, 32768, 1024, dres);
    CK(cudaDeviceSynchronize());
  }
  Result r;
  CK(cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost));
  printf("issue style=%d: %.1f cycles per chunk (4 MMAs N=16 + commit; math ~32)\n", style, (double)r.cycles / r.n_instr);
}

int main() {
  Result* dres;
  CK(cudaMalloc(&dres, sizeof(Result)));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs: %d\n", sms);
  // semantic check: one K=16 MMA, N = 128.  Expected if "B rows [0, N/2) come from CTA 0 and [N/2, N) from CTA 1":
  //   cta0 (A = 1): 16 16 32 32     cta1 (A = 3): 48 48 96 96        (cta_group::1: 16 16 16 16)
  run<1>(128, 0, 0, 1, 1, dres);
  run<2>(128, 0, 0, 1, 2, dres);
  const int ns[] = {32, 64, 96, 128, 192, 256};
  for (int n : ns) { run<1>(n, 0, 0, 256, 1, dres); run<2>(n, 0, 0, 256, 2, dres); }
  for (int n : ns) { run<1>(n, 0, 0, 256, sms, dres); run<2>(n, 0, 0, 256, sms & ~1, dres); }
  // product patterns: (2Nt, Nt) pairs
  const int nts[] = {32, 64, 96, 128};
  for (int nt : nts) { run<1>(2 * nt, nt, 1, 128, sms, dres); run<2>(2 * nt, nt, 1, 128, sms & ~1, dres); run<2>(2 * nt, 2 * nt, 1, 128, sms & ~1, dres); }
  // product issue pattern with distinct operands
  const int rnts[] = {64, 96, 128};
  for (int nt : rnts) {
    run_ring<1, 4>(nt, 0, sms, 1, dres);
    run_ring<2, 4>(nt, 0, sms & ~1, 1, dres);
    run_ring<2, 4>(nt, 1, sms & ~1, 1, dres);
  }
  run_ring<1, 3>(128, 0, 2 * sms, 2, dres);      // two independent CTAs per SM sharing the tensor pipe
  run_ring<1, 3>(64, 0, 2 * sms, 2, dres);
  // L2 -> shared memory feed rate
  {
    uint8_t* buf; unsigned long long* dcyc;
    const size_t window = (size_t)48 << 20;
    CK(cudaMalloc(&buf, window + (1 << 20)));
    CK(cudaMemset(buf, 1, window + (1 << 20)));
    CK(cudaMalloc(&dcyc, 8));
    run_feed<4>(buf, window, sms, dcyc);
    run_feed<8>(buf, window, sms, dcyc);
    run_feed<4>(buf, window, 2 * sms, dcyc);
    run_feed<8>(buf, (size_t)16 << 20, sms, dcyc);
    run_feed<12>(buf, (size_t)16 << 20, sms, dcyc);
  }
  for (int st = 0; st < 4; ++st) run_issue(st, dres);
  printf("done\n");
  return 0;
}
