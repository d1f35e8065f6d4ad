// kw-folded tensor-core convolution for the narrow (Cin = 32) stride-1 cubic layers of the V2V net at
// full resolution (v2v.py:146 7^3 32->16, :24-28 3^3 32->32) -- the layers that dominate the step.
//
// Why: with 32 input channels an implicit-GEMM tap carries only K = 32, so a generic "one shifted A tile per
// tap" kernel re-streams 128 x 128 B from L2 for every tap and is pinned to the L2->smem bandwidth, and a 128 x N
// UMMA with N = 16/32 is bound by the shared-memory read of A (32 cycles) rather than by the tensor pipe.
// Here the kw taps are folded into the MMA's N dimension:
//     D[(line, xi)][kw*NC + co] = sum_{kd,kh,ci} in[z+kd-p][y+line+kh-p][x0-p+xi][ci] * W[kd][kh][kw][ci][co]
//     out[(line, xo)][co]       = sum_kw D[(line, xo+kw)][kw*NC + co]          (shift-add in the epilogue)
// so one un-shifted line tile feeds N = K*NC = 96 / 112 columns (MMA is math-bound, not A-read-bound),
// a z-slab of (LINES+K-1) lines is loaded ONCE per kd and reused for all kh via the descriptor start address, and
// the 3^3 weights (108 KB) stay resident in shared memory for the whole persistent CTA.  Full-width mode (W = 16 / 32 / 64:
// the M tile spans whole lines) keeps every row; otherwise a 16- / 32-row x window yields 16-K+1 / 32-K+1 outputs per line.
//
// Products: activations are split-fp16 rows [32 hi | 32 lo]; the weights of a (kd,kh) step are stored as a
// 2*NF-row, 64-byte-swizzled tile  [hi rows (kw,co) ; lo rows (kw,co)] x 32 channels, so that per 16-wide K slice
//     MMA1: A_hi x [B_hi ; B_lo]  (N = 2*NF)  -> [D1 | D2]     (hi*hi and hi*lo in one pass over A_hi)
//     MMA2: A_lo x  B_hi          (N =   NF)  ->       D2
// i.e. two MMAs instead of three (a single-CTA SS-mode UMMA costs ~(128 + N)/2 cycles of operand fetch).
//
// Persistent CTAs (one per SM), warp roles as in conv_tc.cu; TMEM holds two accumulator stages of 2*NF columns
// (hi*hi and the cross terms) so the epilogue of tile i overlaps the MMAs of tile i+1; the epilogue
// shift-adds with warp shuffles, applies scale/shift/residual/ReLU, compacts the valid rows into a 128B-swizzled
// staging tile and stores it with TMA (which also clips partial tiles).
#include "tc_common.cuh"
#include "pair_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace lt {

struct FoldParams {
  int N, D, H, W;
  int K, pad;
  int NC, NF, OWt;
  int WX, LINES, wx_shift;   // x window of the M tile (16 or 32 rows per line) and lines per tile (128 / WX)
  int xwins, yblks;
  long tiles;
  int slab_bytes, a_slots;
  int b_resident, b_slots, b_bytes;
  int stage_rows;        // LINES * OWt
  int fullw;             // 1: the M tile spans whole lines (WX == W, OWt == W): no x halo, the kw shift-add zero-fills at the line ends
  int x_halo;            // x offset of the M tile's first row relative to its first output (pad, or 0 in full-width mode)
  int pair;              // CTA-pair variant (see conv_fold_kernel<true>): b_bytes is the per-CTA share of a (kd, kh) weight tile
  int off_xch;           // full-width, W > 32: edge rows exchanged between the epilogue warps of a line (floats, [half][quad][side][pad][pad][16])
  int out_format, relu, residual;
  int direct;            // full-width, split-fp16 output: rows go from registers to global memory (no staging tile, no TMA store / residual load)
  const void* res;
  void* out;
  const float* scale;
  const float* shift;
  // smem offsets (bytes from the 1024-aligned base)
  int off_b, off_a, off_out, off_res, off_bar;
  unsigned long long* prof;   // optional [16] cycle counters (LT_FOLD_PROF=1): time spent waiting per role / barrier
  int dbg;               // bring-up knobs (LT_FOLD_DBG): 1 = hi*hi MMAs only, 2 = skip epilogue math/stores, 3 = load slabs once
  int fast_issue;        // 1: single-lane MMA issue loop with the kh taps unrolled (fold_issue_loop), 0: generic loop
};

__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, unsigned long long& acc, bool on) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MMA issue loop with K compile-time: the whole (converged) warp runs it so that slab / weight descriptors stay in uniform
// registers; per kd one slab wait, then -- resident weights -- ONE election covering K x 4 tcgen05.mma whose descriptors
// differ from the slab / weight base by compile-time constants (no per-tap address arithmetic, parameter reloads, modulo).
// Measured on B200 (LT_FOLD_PROF): the generic loop spent ~590 cycles per (kd, kh) step issuing 4 MMAs whose execution
// takes 288 (3^3) / 336 (7^3) cycles and never waited on TMA or the epilogue: the kernel was bound by its own issue latency.
template <bool PAIR>
__device__ __forceinline__ void fold_mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma2_f16(d, ad, bd, idesc, acc); else umma_f16(d, ad, bd, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void fold_commit(uint64_t* bar) {
  if (PAIR) umma2_commit_mc(bar); else umma_commit(bar);
}

template <int K, bool PAIR>
__device__ __forceinline__ void fold_issue_loop(const FoldParams& p, uint32_t tmem_base, uint32_t a_base, uint32_t b_base, uint64_t* a_full,
                                                uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty, uint64_t* acc_full,
                                                uint64_t* acc_empty) {
  const uint32_t idesc = PAIR ? make_idesc_f16_m256(p.NF) : make_idesc_f16(p.NF), idesc2 = PAIR ? make_idesc_f16_m256(2 * p.NF) : make_idesc_f16(2 * p.NF);
  // pair: the A_lo x B_hi product reads the second region of a CTA's weight tile (this CTA's half of the B_hi rows), NF rows in
  const uint64_t b2 = PAIR ? (uint64_t)((uint32_t)(p.NF * 64) >> 4) : 0ull;
  const long tile0 = PAIR ? (long)(blockIdx.x >> 1) : (long)blockIdx.x, tstep = PAIR ? (long)(gridDim.x >> 1) : (long)gridDim.x;
  const long ntile = PAIR ? (p.tiles + 1) / 2 : p.tiles;
  const uint32_t line_step = (uint32_t)(p.WX * 128) >> 4;     // descriptor address units (16 B) per slab line
  const uint32_t b_bytes = (uint32_t)p.b_bytes, b_step = b_bytes >> 4;
  const uint32_t slab_bytes = (uint32_t)p.slab_bytes, a_slots = (uint32_t)p.a_slots, b_slots = (uint32_t)p.b_slots;
  const uint32_t nf = (uint32_t)p.NF;
  const bool resident = p.b_resident != 0;
  if (resident) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
  uint32_t slot = 0, a_ph = 0, bs = 0, b_ph = 0, it = 0;
  const bool prof = p.prof != nullptr;
  unsigned long long w_acc = 0, w_af = 0, w_bf = 0;
  const long long tstart = prof ? clock64() : 0;
  for (long tile = tile0; tile < ntile; tile += tstep, ++it) {
    const uint32_t as = it & 1u;
    mbar_wait_prof(&acc_empty[as], ((it >> 1) & 1u) ^ 1u, w_acc, prof);
    tc_fence_after();
    const uint32_t d1 = tmem_base + as * 2u * nf, d2 = d1 + nf;
#pragma unroll 1
    for (int kd = 0; kd < K; ++kd) {
      mbar_wait_prof(&a_full[slot], a_ph, w_af, prof);
      tc_fence_after();
      const uint64_t ad0 = make_sw128_desc(a_base + slot * slab_bytes);
      const uint32_t first = kd ? 1u : 0u;
      if (resident) {
        const uint64_t bd0 = make_sw64_desc(b_base + (uint32_t)(kd * K) * b_bytes);   // weights of (kd, kh = 0)
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < K; ++kh) {
            const uint64_t ad = ad0 + (uint64_t)((uint32_t)kh * line_step), bd = bd0 + (uint64_t)((uint32_t)kh * b_step);
            fold_mma<PAIR>(d1, ad, bd, idesc2, kh == 0 ? first : 1u);   // A_hi(s0) x [B_hi;B_lo](s0) -> [D1|D2]
            fold_mma<PAIR>(d2, ad + 4, bd + b2, idesc, 1);               // A_lo(s0) x B_hi(s0)        -> D2
            fold_mma<PAIR>(d1, ad + 2, bd + 2, idesc2, 1);               // slice 1
            fold_mma<PAIR>(d2, ad + 6, bd + b2 + 2, idesc, 1);
          }
          fold_commit<PAIR>(&a_empty[slot]);
          if (kd == K - 1) fold_commit<PAIR>(&acc_full[as]);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          mbar_wait_prof(&b_full[bs], b_ph, w_bf, prof);
          tc_fence_after();
          const uint64_t ad = ad0 + (uint64_t)((uint32_t)kh * line_step), bd = make_sw64_desc(b_base + bs * b_bytes);
          if (elect_one()) {
            fold_mma<PAIR>(d1, ad, bd, idesc2, kh == 0 ? first : 1u);
            fold_mma<PAIR>(d2, ad + 4, bd + b2, idesc, 1);
            fold_mma<PAIR>(d1, ad + 2, bd + 2, idesc2, 1);
            fold_mma<PAIR>(d2, ad + 6, bd + b2 + 2, idesc, 1);
            fold_commit<PAIR>(&b_empty[bs]);
            if (kh == K - 1) fold_commit<PAIR>(&a_empty[slot]);
            if (kh == K - 1 && kd == K - 1) fold_commit<PAIR>(&acc_full[as]);
          }
          __syncwarp();
          if (++bs == b_slots) { bs = 0; b_ph ^= 1u; }
        }
      }
      if (++slot == a_slots) { slot = 0; a_ph ^= 1u; }
    }
  }
  if (prof && (threadIdx.x & 31) == 0) {
    atomicAdd(&p.prof[3], w_acc); atomicAdd(&p.prof[4], w_af); atomicAdd(&p.prof[5], w_bf);
    atomicAdd(&p.prof[6], (unsigned long long)(clock64() - tstart));
  }
}

// PAIR = true: two CTAs of a cluster (one TPC) run the (kd, kh) steps of their two M tiles as ONE M = 256 tcgen05.mma.cta_group::2
// sequence issued by the leader.  Why (ncu, profiles/r02c_fold_summary.md): with M = 128 a K slice reads A 4 KB + [B_hi;B_lo] 6 KB and
// A_lo 4 KB + B_hi 3 KB of shared memory for 144 cycles of math -- 118 B/clk against the 128 B/clk the tensor core can fetch: the
// operand-fetch path was 75 % busy and the math pipe 54 %.  In a pair each CTA fetches its own A tile but only HALF of every B
// operand (a cta_group::2 MMA takes B rows [0, N/2) from CTA 0 and [N/2, N) from CTA 1): 87 B/clk.  Per-CTA weight tile of a
// (kd, kh) step: rank 0 = [B_hi (NF rows) ; B_hi rows [0, NF/2)], rank 1 = [B_lo (NF rows) ; B_hi rows [NF/2, NF)] (64-byte rows).
// Barriers: a_full / b_full / acc_empty live in the leader (the peer's TMA loads and epilogue warps signal them remotely),
// a_empty / b_empty / acc_full are per CTA and receive the leader's multicast commits.
template <bool PAIR>
__global__ void __launch_bounds__(320, 1) conv_fold_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmOut,
                                                           const __grid_constant__ CUtensorMap tmRes, const FoldParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_smem = smem + p.off_b;
  uint8_t* a_smem = smem + p.off_a;
  uint8_t* out_stage = smem + p.off_out;
  uint8_t* res_stage = smem + p.off_res;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* a_full = bars;                 // [a_slots]
  uint64_t* a_empty = a_full + p.a_slots;  // [a_slots]
  uint64_t* b_full = a_empty + p.a_slots;  // [b_slots] (streamed) or [1] (resident: all weights)
  uint64_t* b_empty = b_full + p.b_slots;  // [b_slots]
  uint64_t* acc_full = b_empty + p.b_slots;  // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* res_full = acc_empty + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int taps2 = p.K * p.K;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // this CTA's tiles: tile0, tile0 + tstep, ...  (pair k of P walks the tile pairs k, k + P, ...; rank r takes tile 2 * pair + r; an odd
  // tile count leaves the last pair's rank 1 with tile == p.tiles, which decodes to n == N: loads zero-fill, stores are clipped)
  const long tile0 = PAIR ? (long)(blockIdx.x >> 1) * 2 + rank : (long)blockIdx.x;
  const long tstep = PAIR ? (long)(gridDim.x >> 1) * 2 : (long)gridDim.x;
  const long tile_end = PAIR ? ((p.tiles + 1) / 2) * 2 : p.tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.b_slots; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 16 : kEpiThreads); }
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmOut); }
  if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, 512u); else tmem_alloc(tmem_slot, 512u); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // both CTAs' barriers initialised and TMEM allocated before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loops; one elected lane issues) =================
    // pair: "full" barriers are the leader's; its producer expects both CTAs' bytes, each CTA's loads signal it (cluster address)
    const uint32_t a_full0 = PAIR ? map_to_cta(smem_u32(&a_full[0]), 0) : 0u, b_full0 = PAIR ? map_to_cta(smem_u32(&b_full[0]), 0) : 0u;
    const uint32_t tx_mul = PAIR ? 2u : 1u;
    const int b_rows = PAIR ? p.NF + p.NF / 2 : 2 * p.NF;      // rows of one (kd, kh) weight tile of this CTA
    if (p.b_resident) {
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(&b_full[0], tx_mul * (uint32_t)(taps2 * p.b_bytes));
        for (int t = 0; t < taps2; ++t) {
          if (PAIR) tma2_load_2d(b_smem + (size_t)t * p.b_bytes, &tmB, b_full0, 0, (t * 2 + (int)rank) * b_rows);
          else tma_load_2d(b_smem + (size_t)t * p.b_bytes, &tmB, &b_full[0], 0, t * b_rows);
        }
      }
      __syncwarp();
    }
    uint32_t pa = 0, pb = 0;
    uint32_t slot = 0, a_ph = 0, bsl = 0, b_ph = 0;     // ring slots / phases advanced incrementally (no modulo per step)
    const bool prof = p.prof != nullptr;
    unsigned long long w_ae = 0, w_be = 0;
    const long long tstart = clock64();
    for (long tile = tile0; tile < tile_end; tile += tstep) {
      long t = tile;
      const int xw = (int)(t % p.xwins); t /= p.xwins;
      const int yb = (int)(t % p.yblks); t /= p.yblks;
      const int z = (int)(t % p.D);
      const int n = (int)(t / p.D);
      for (int kd = 0; kd < p.K; ++kd) {
        mbar_wait_prof(&a_empty[slot], a_ph ^ 1u, w_ae, prof);
        if (elect_one()) {
          if (!PAIR && p.dbg == 3 && pa >= (uint32_t)p.a_slots) {
            mbar_arrive(&a_full[slot]);          // debug: reuse stale slab contents, no TMA traffic
          } else {
            if (rank == 0) mbar_expect_tx(&a_full[slot], tx_mul * (uint32_t)p.slab_bytes);
            if (PAIR) tma2_load_5d(a_smem + (size_t)slot * p.slab_bytes, &tmA, a_full0 + slot * 8u, 0, xw * p.OWt - p.x_halo, yb * p.LINES - p.pad,
                                   z + kd - p.pad, n);
            else tma_load_5d(a_smem + (size_t)slot * p.slab_bytes, &tmA, &a_full[slot], 0, xw * p.OWt - p.x_halo, yb * p.LINES - p.pad,
                             z + kd - p.pad, n);
          }
        }
        __syncwarp();
        ++pa;
        if (++slot == (uint32_t)p.a_slots) { slot = 0; a_ph ^= 1u; }
        if (!p.b_resident) {
          for (int kh = 0; kh < p.K; ++kh) {
            const uint32_t bs = bsl;
            mbar_wait_prof(&b_empty[bs], b_ph ^ 1u, w_be, prof);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&b_full[bs], tx_mul * (uint32_t)p.b_bytes);
              if (PAIR) tma2_load_2d(b_smem + (size_t)bs * p.b_bytes, &tmB, b_full0 + bs * 8u, 0, ((kd * p.K + kh) * 2 + (int)rank) * b_rows);
              else tma_load_2d(b_smem + (size_t)bs * p.b_bytes, &tmB, &b_full[bs], 0, (kd * p.K + kh) * b_rows);
            }
            __syncwarp();
            ++pb;
            if (++bsl == (uint32_t)p.b_slots) { bsl = 0; b_ph ^= 1u; }
          }
        }
      }
    }
    if (prof && lane == 0) { atomicAdd(&p.prof[0], w_ae); atomicAdd(&p.prof[1], w_be); atomicAdd(&p.prof[2], (unsigned long long)(clock64() - tstart)); }
  } else if (warp == 1 && (PAIR || p.fast_issue)) {
    // ================= MMA issuer, fast path: converged warp, kh taps unrolled, one election per slab (pair: leader CTA only) ==========
    if (rank == 0) {
      if (p.K == 3) fold_issue_loop<3, PAIR>(p, tmem_base, smem_u32(a_smem), smem_u32(b_smem), a_full, a_empty, b_full, b_empty, acc_full, acc_empty);
      else fold_issue_loop<7, PAIR>(p, tmem_base, smem_u32(a_smem), smem_u32(b_smem), a_full, a_empty, b_full, b_empty, acc_full, acc_empty);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loops; one elected lane issues) =================
    const uint32_t idesc = make_idesc_f16(p.NF), idesc2 = make_idesc_f16(2 * p.NF);
    if (p.b_resident) { mbar_wait(&b_full[0], 0); }
    uint32_t pa = 0, pb = 0, it = 0;
    const bool prof = p.prof != nullptr;
    unsigned long long w_acc = 0, w_af = 0, w_bf = 0;
    const long long tstart = clock64();
    for (long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u;
      mbar_wait_prof(&acc_empty[as], ((it >> 1) & 1u) ^ 1u, w_acc, prof);
      tc_fence_after();
      const uint32_t d1 = tmem_base + as * 2u * (uint32_t)p.NF;
      const uint32_t d2 = d1 + (uint32_t)p.NF;
      for (int kd = 0; kd < p.K; ++kd) {
        const uint32_t slot = pa % p.a_slots;
        mbar_wait_prof(&a_full[slot], (pa / p.a_slots) & 1u, w_af, prof);
        tc_fence_after();
        const uint32_t slab = smem_u32(a_smem + (size_t)slot * p.slab_bytes);
        for (int kh = 0; kh < p.K; ++kh) {
          uint32_t b_addr;
          uint32_t bs = 0;
          if (p.b_resident) {
            b_addr = smem_u32(b_smem + (size_t)(kd * p.K + kh) * p.b_bytes);
          } else {
            bs = pb % p.b_slots;
            mbar_wait_prof(&b_full[bs], (pb / p.b_slots) & 1u, w_bf, prof);
            tc_fence_after();
            b_addr = smem_u32(b_smem + (size_t)bs * p.b_bytes);
          }
          const uint64_t ad = make_sw128_desc(slab + (uint32_t)(kh * p.WX) * 128u);   // line kh of the slab: WX rows x 128 B
          const uint64_t bd = make_sw64_desc(b_addr);                          // 2*NF rows x 64 B: [hi rows ; lo rows]
          const uint32_t first = (kd | kh) ? 1u : 0u;
          if (elect_one()) {
            // A: hi slices at +0,+32 B, lo slices at +64,+96 B of each 128-byte row; B: slices at +0,+32 B of each 64-byte row
            umma_f16(d1, ad, bd, idesc2, first);          // A_hi(s0) x [B_hi;B_lo](s0) -> [D1|D2]
            umma_f16(d2, ad + 4, bd, idesc, 1);           // A_lo(s0) x B_hi(s0)        -> D2
            umma_f16(d1, ad + 2, bd + 2, idesc2, 1);      // slice 1
            umma_f16(d2, ad + 6, bd + 2, idesc, 1);
            if (!p.b_resident) umma_commit(&b_empty[bs]);
            if (kh == p.K - 1) umma_commit(&a_empty[slot]);
            if (kh == p.K - 1 && kd == p.K - 1) umma_commit(&acc_full[as]);
          }
          __syncwarp();
          if (!p.b_resident) ++pb;
        }
        ++pa;
      }
    }
    if (prof && lane == 0) { atomicAdd(&p.prof[3], w_acc); atomicAdd(&p.prof[4], w_af); atomicAdd(&p.prof[5], w_bf); atomicAdd(&p.prof[6], (unsigned long long)(clock64() - tstart)); }
  } else {
    // ================= epilogue (warps 2..9: two per TMEM lane quadrant, 16 output channels each) =================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int line = row >> p.wx_shift, xi = row & (p.WX - 1);
    const bool keep = xi < p.OWt;
    const int srow = line * p.OWt + xi;        // compacted staging row
    const bool leader = threadIdx.x == 64;
    const bool has_cols = half * 16 < p.NC;    // NC = 16: the upper half only writes the zero padding channels
    const uint32_t stage_bytes = (uint32_t)p.stage_rows * 128u;
    float* const xch_base = p.off_xch ? reinterpret_cast<float*>(smem + p.off_xch) : nullptr;
    const int xch_floats = 256 * p.pad * p.pad;     // one buffer: [half][quad][side][pad][pad][16]
    const bool right_nb = (((quad + 1) * 32) & (p.WX - 1)) != 0, left_nb = ((quad * 32) & (p.WX - 1)) != 0;   // neighbour warp in the same line
    uint32_t it = 0;
    const bool prof = p.prof != nullptr;
    unsigned long long w_accf = 0, w_res = 0;
    const long long tstart = clock64();
    const uint32_t acc_empty0 = PAIR ? map_to_cta(smem_u32(&acc_empty[0]), 0) : 0u;
    for (long tile = tile0; tile < tile_end; tile += tstep, ++it) {
      long t = tile;
      const int xw = (int)(t % p.xwins); t /= p.xwins;
      const int yb = (int)(t % p.yblks); t /= p.yblks;
      const int z = (int)(t % p.D);
      const int n = (int)(t / p.D);
      const uint32_t as = it & 1u;
      float* const xch = xch_base ? xch_base + ((p.direct && (it & 1u)) ? xch_floats : 0) : nullptr;
      // direct mode: this thread's voxel row (16 channels = 32 B of high halves + 32 B of low halves) goes straight from registers
      // to global memory and its residual comes straight from global memory, requested here, a whole accumulator wait ahead.
      // (The staged path queues its TMA store behind every slab load the producer has already issued -- the SM's TMA queue is
      // FIFO -- and the next tile cannot touch the staging buffer before that store has drained: ~3600 cycles per tile measured
      // against 2592 cycles of MMA work for the 3^3 layers.)
      uint32_t rh[8], rl[8];
      uint8_t* gdst = nullptr;
      bool gvalid = false;
      if (p.direct) {
        const int y = yb * p.LINES + line;
        gvalid = y < p.H && n < p.N;
        const long voxel = (((long)n * p.D + z) * p.H + y) * p.W + xi;
        gdst = reinterpret_cast<uint8_t*>(p.out) + voxel * 128 + half * 32;
        if (p.residual != LT_RES_NONE && gvalid) {
          const uint8_t* gsrc = reinterpret_cast<const uint8_t*>(p.res) + voxel * 128 + half * 32;
          ldg256(gsrc, rh);
          ldg256(gsrc + 64, rl);
        }
      }
      if (leader && p.residual != LT_RES_NONE && !p.direct) {
        bulk_wait_read<0>();                    // staging buffer is shared: the previous tile's store must have drained it
        mbar_expect_tx(res_full, stage_bytes);
        tma_load_5d(res_stage, &tmRes, res_full, 0, xw * p.OWt, yb * p.LINES, z, n);
      }
      mbar_wait_prof(&acc_full[as], (it >> 1) & 1u, w_accf, prof);
      tc_fence_after();
      float v[16], r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0f;
      if (has_cols) {
        const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 2u * (uint32_t)p.NF + (uint32_t)(half * 16);
        for (int kw = 0; kw < p.K; ++kw) {
          uint32_t t1[16], t2[16];
          tmem_ld16_nowait(tb + (uint32_t)(kw * p.NC), t1);
          tmem_ld16_nowait(tb + (uint32_t)(p.NF + kw * p.NC), t2);
          tmem_wait_ld();
          if (!p.fullw) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float d = fmaf(__uint_as_float(t2[j]), kLoInv, __uint_as_float(t1[j]));
              v[j] += __shfl_down_sync(0xffffffffu, d, kw);    // row (line, xi + kw) -> output (line, xi)
            }
          } else {
            // full-width tile: out(line, xi) += D(line, xi + s)[kw], s = kw - pad; rows outside [0, W) are the zero padding
            const int sft = kw - p.pad, src = lane + sft;
            const bool ok = (unsigned)(xi + sft) < (unsigned)p.WX && (unsigned)src < 32u;
            float d[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              d[j] = fmaf(__uint_as_float(t2[j]), kLoInv, __uint_as_float(t1[j]));
              const float t = __shfl_sync(0xffffffffu, d[j], src & 31);
              v[j] += ok ? t : 0.0f;
            }
            if (xch) {
              // lines wider than a warp: rows whose source sits in the neighbouring warp of the same line go through smem.
              // low-edge row r' = lane publishes the blocks kw in (pad + r', 2 pad]; high-edge row r'' = 31 - lane the blocks [0, pad - r'')
              int slot = -1;
              if (lane < p.pad && kw > p.pad + lane) slot = (0 * p.pad + lane) * p.pad + (kw - p.pad - 1);
              if (lane >= 32 - p.pad && kw < p.pad - (31 - lane)) slot = (1 * p.pad + (31 - lane)) * p.pad + kw;
              if (slot >= 0) {
                float* dst = xch + ((size_t)((half * 4 + quad) * 2 * p.pad * p.pad + slot) << 4);
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(d[j], d[j + 1], d[j + 2], d[j + 3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      if (PAIR) {                               // accumulator stage drained (one arrival per warp, on the leader's barrier)
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acc_empty0 + as * 8u);
      } else {
        mbar_arrive(&acc_empty[as]);            // the MMA warp may start tile it+2
      }
      if (p.dbg == 2) continue;
      if (leader && !p.direct) bulk_wait_read<0>();   // previous tile's store has finished reading the staging buffer
      if (!p.direct || xch) epi_bar_sync();     // staging buffer free / every warp's edge rows are published
      if (xch && has_cols) {
        const int pp = p.pad * p.pad;
        if (right_nb && lane >= 32 - p.pad) {   // sources in the next warp's low edge
          for (int sft = 32 - lane; sft <= p.pad; ++sft) {
            const float* src = xch + ((size_t)((half * 4 + quad + 1) * 2 * pp + (lane + sft - 32) * p.pad + (sft - 1)) << 4);
#pragma unroll
            for (int j = 0; j < 16; j += 4) { const float4 q = *reinterpret_cast<const float4*>(src + j); v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w; }
          }
        }
        if (left_nb && lane < p.pad) {          // sources in the previous warp's high edge
          for (int sft = -p.pad; sft <= -(lane + 1); ++sft) {
            const float* src = xch + ((size_t)((half * 4 + quad - 1) * 2 * pp + pp + (-1 - lane - sft) * p.pad + (sft + p.pad)) << 4);
#pragma unroll
            for (int j = 0; j < 16; j += 4) { const float4 q = *reinterpret_cast<const float4*>(src + j); v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w; }
          }
        }
      }
      epi_affine16(v, p.scale, p.shift, half * 16);
      if (p.direct) {
        if (p.residual != LT_RES_NONE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&rh[j])), b = __half22float2(*reinterpret_cast<const __half2*>(&rl[j]));
            r[2 * j] = gvalid ? fmaf(b.x, kLoInv, a.x) : 0.0f;
            r[2 * j + 1] = gvalid ? fmaf(b.y, kLoInv, a.y) : 0.0f;
          }
        }
        epi_activate16(v, r, p.residual, p.relu);
        if (gvalid) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_s32x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
          stg256(gdst, hi);
          stg256(gdst + 64, lo);
        }
        continue;                               // (the edge-row exchange is double-buffered by tile parity: no second barrier)
      }
      if (p.residual != LT_RES_NONE) {
        mbar_wait_prof(res_full, it & 1u, w_res, prof);
        if (keep) epi_load16(smem_u32(res_stage), srow, half, p.out_format, r);
      }
      epi_activate16(v, r, p.residual, p.relu);
      // the residual tile and the output tile share one staging buffer: every thread overwrites exactly the 16-channel half row it
      // has just read, so no barrier is needed between the read and the write
      if (keep) epi_store16(smem_u32(out_stage), srow, half, p.out_format, v);
      fence_proxy_async();
      epi_bar_sync();
      if (leader) {
        tma_store_5d(&tmOut, out_stage, 0, xw * p.OWt, yb * p.LINES, z, n);
        bulk_commit();
      }
    }
    if (leader) bulk_wait<0>();
    if (prof && threadIdx.x == 64) { atomicAdd(&p.prof[7], w_accf); atomicAdd(&p.prof[8], w_res); atomicAdd(&p.prof[9], (unsigned long long)(clock64() - tstart)); }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer may still read this CTA's weight half / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, 512u); else tmem_dealloc(tmem_base, 512u);
  }
}

// fp32 [K^3 taps (kd,kh,kw)][32][Cout] -> fp16 [kd][kh][2][kw*NC + co][32]: hi rows then lo rows, 64 bytes per row
__global__ void __launch_bounds__(256) pack_fold_weights_kernel(const float* __restrict__ w, sh_t* __restrict__ out, int K, int Cout, int NC) {
  const int NF = K * NC;
  const long total = (long)K * K * NF * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % 32);
    long r = i / 32;
    const int col = (int)(r % NF); r /= NF;
    const int kh = (int)(r % K);
    const int kd = (int)(r / K);
    const int kw = col / NC, co = col % NC;
    const float v = (co < Cout) ? w[((((long)kd * K + kh) * K + kw) * 32 + ci) * Cout + co] : 0.0f;
    sh_t hi, lo;
    split_s32(v, hi, lo);
    sh_t* tile = out + ((long)kd * K + kh) * 2 * NF * 32;
    tile[(long)col * 32 + ci] = hi;
    tile[((long)NF + col) * 32 + ci] = lo;
  }
}

// CTA-pair layout (conv_fold_kernel<true>): fp16 [kd][kh][rank][NF + NF/2 rows][32], 64 bytes per row:
//   rank 0: B_hi rows (kw, co) [0, NF), then B_hi rows [0, NF/2);   rank 1: B_lo rows [0, NF), then B_hi rows [NF/2, NF)
__global__ void __launch_bounds__(256) pack_fold_pair_weights_kernel(const float* __restrict__ w, sh_t* __restrict__ out, int K, int Cout, int NC) {
  const int NF = K * NC, R = NF + NF / 2;
  const long total = (long)K * K * 2 * R * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % 32);
    long r = i / 32;
    const int row = (int)(r % R); r /= R;
    const int rank = (int)(r % 2); r /= 2;
    const int kh = (int)(r % K);
    const int kd = (int)(r / K);
    const bool second = row >= NF;                       // second region: this CTA's half of the B_hi rows
    const int col = second ? (row - NF) + rank * (NF / 2) : row;
    const bool want_lo = !second && rank == 1;
    const int kw = col / NC, co = col % NC;
    const float v = (co < Cout) ? w[((((long)kd * K + kh) * K + kw) * 32 + ci) * Cout + co] : 0.0f;
    sh_t hi, lo;
    split_s32(v, hi, lo);
    out[i] = want_lo ? lo : hi;
  }
}

int conv_fold_supported(const lt_conv_desc* d) {
  const bool cubic = d->KD == d->KH && d->KH == d->KW && (d->KW == 3 || d->KW == 7);
  const int p = d->KW / 2;
  const int NC = (d->Cout + 15) & ~15;
  return cubic && d->Cin == 32 && NC <= 32 && d->KW * NC <= 128 && 2 * d->KW * NC <= 256 && d->sd == 1 && d->sh == 1 && d->sw == 1 && d->pd == p &&
         d->ph == p && d->pw == p && d->OD == d->ID && d->OH == d->IH && d->OW == d->IW && d->osd == 1 && d->osh == 1 &&
         d->osw == 1 && d->ood == 0 && d->ooh == 0 && d->oow == 0 && d->FD == d->OD && d->FH == d->OH && d->FW == d->OW &&
         d->FC == 32 && d->in_format == LT_FMT_S32 && d->IW >= 16;
}

static inline int up1024(int v) { return (v + 1023) & ~1023; }

int conv_fold_fwd(const lt_conv_desc* d, const void* in, const void* weight, const float* scale, const float* shift,
                  const void* residual, void* out, void* stream) {
  LT_REQUIRE(conv_fold_supported(d), "conv_fold: unsupported layer shape");
  FoldParams p;
  p.N = d->N; p.D = d->ID; p.H = d->IH; p.W = d->IW;
  p.K = d->KW; p.pad = d->KW / 2;
  p.NC = (d->Cout + 15) & ~15;
  p.NF = p.K * p.NC;
  // x window.  Full-width mode (W = 16 / 32 / 64): the M tile spans whole lines, every MMA row is an output position and the
  // kw shift-add of the epilogue zero-fills at the line ends (the rows beyond them are the convolution's zero padding) -- no
  // wasted rows.  Otherwise 16- or 32-row windows that keep 16-K+1 / 32-K+1 outputs per line (whichever wastes fewer rows).
  p.fullw = (opts().fold_fullw && (p.W == 16 || p.W == 32 || p.W == 64)) ? 1 : 0;
  if (p.fullw) {
    p.WX = p.W;
  } else {
    p.WX = 16;
    if (p.W >= 32 && ceil_div(p.W, 32 - p.K + 1) * 32 < ceil_div(p.W, 16 - p.K + 1) * 16) p.WX = 32;
  }
  p.wx_shift = p.WX == 64 ? 6 : (p.WX == 32 ? 5 : 4);
  p.LINES = 128 / p.WX;
  p.OWt = p.fullw ? p.WX : p.WX - p.K + 1;
  p.x_halo = p.fullw ? 0 : p.pad;
  p.xwins = ceil_div(p.W, p.OWt);
  p.yblks = ceil_div(p.H, p.LINES);
  p.tiles = (long)p.N * p.D * p.yblks * p.xwins;
  const int slab_lines = p.LINES + p.K - 1;
  p.slab_bytes = p.WX * slab_lines * 128;
  // CTA pairs (fold_pair = 2: always; 1: when the one-CTA kernel would have to stream its weights) if there are at least two tiles;
  // debug knock-outs stay on the one-CTA kernel
  const bool single_resident = p.K * p.K * p.NF * 128 <= 112 * 1024;
  p.pair = ((opts().fold_pair == 2 || (opts().fold_pair == 1 && !single_resident)) && p.tiles >= 2 && sm_count() >= 2 &&
            (opts().fold_debug & 15) == 0) ? 1 : 0;
  p.b_bytes = p.pair ? p.NF * 96 : p.NF * 128;     // per-CTA share of a (kd, kh) weight tile: 1.5 NF or 2 NF rows of 64 bytes
  p.b_resident = (p.K * p.K * p.b_bytes <= 112 * 1024) ? 1 : 0;
  p.stage_rows = p.LINES * p.OWt;
  p.out_format = d->out_format; p.relu = d->relu; p.residual = d->residual;
  p.scale = scale; p.shift = shift;
  p.res = residual; p.out = out;
  p.direct = (p.fullw && d->out_format == LT_FMT_S32 && opts().fold_direct) ? 1 : 0;
  p.dbg = opts().fold_debug & 15;
  static unsigned long long* prof_buf = nullptr;
  const bool want_prof = (opts().fold_debug & 16) != 0;
  if (want_prof && !prof_buf) cudaMalloc(&prof_buf, 16 * sizeof(unsigned long long));
  p.prof = want_prof ? prof_buf : nullptr;
  const int fast_issue = opts().fold_fast_issue;
  p.fast_issue = (fast_issue && p.dbg == 0) ? 1 : 0;
  if (want_prof) cudaMemsetAsync(prof_buf, 0, 16 * sizeof(unsigned long long), (cudaStream_t)stream);
  // shared memory: weights (resident: all K^2 (kd, kh) tiles; streamed: a ring of b_slots tiles that hides the L2 latency of a
  // (kd, kh) step), the slab ring (as many slots as fit, at most 5), one staging tile, the edge-row exchange
  const int xch_bytes = (p.fullw && p.WX > 32) ? 1024 * p.pad * p.pad * (p.direct ? 2 : 1) : 0;
  const int stage_bytes_out = p.direct ? 0 : up1024(p.stage_rows * 128);
  const int budget = 227 * 1024 - 1024 - stage_bytes_out - xch_bytes - 512;
  if (p.b_resident) {
    p.b_slots = 1;
  } else {
    p.b_slots = p.WX == 16 ? 8 : 6;
    while (p.b_slots > 3 && budget - up1024(p.b_slots * p.b_bytes) < 2 * up1024(p.slab_bytes)) --p.b_slots;
  }
  const int b_region = p.b_resident ? p.K * p.K * p.b_bytes : p.b_slots * p.b_bytes;
  p.a_slots = (budget - up1024(b_region)) / p.slab_bytes;
  if (p.a_slots > 5) p.a_slots = 5;
  if (!p.b_resident && p.a_slots > 3) p.a_slots = 3;
  LT_REQUIRE(p.a_slots >= 2, "conv_fold: shared memory budget exceeded (slab %d B, weights %d B)", p.slab_bytes, b_region);
  p.off_b = 0;
  p.off_a = up1024(b_region);
  p.off_out = p.off_a + up1024(p.a_slots * p.slab_bytes);
  p.off_res = p.off_out;   // residual tile and output tile share one staging buffer (each thread overwrites the half row it has read)
  p.off_xch = xch_bytes ? p.off_out + stage_bytes_out : 0;
  p.off_bar = p.off_out + stage_bytes_out + xch_bytes;
  const size_t smem = (size_t)p.off_bar + (2 * p.a_slots + 2 * p.b_slots + 5) * 8 + 16 + 1024;
  LT_REQUIRE(smem <= 227 * 1024, "conv_fold: shared memory budget exceeded (%zu)", smem);

  CUtensorMap tmA, tmB, tmOut, tmRes;
  {
    const uint64_t rowb = 128;  // 32 channels split-fp16
    const uint64_t dims[5] = {64, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N};
    const uint64_t str[4] = {rowb, rowb * p.W, rowb * p.W * p.H, rowb * p.W * p.H * p.D};
    const uint32_t bx[5] = {64, (uint32_t)p.WX, (uint32_t)slab_lines, 1, 1};
    int rc = make_map(&tmA, in, 5, dims, str, bx, nullptr, 1);
    if (rc) return rc;
  }
  {
    // the packed filter holds both layouts: [one-CTA: K^2 x 2 NF rows | pair: K^2 x 2 ranks x 1.5 NF rows], 64-byte rows, 64B swizzle
    const size_t single_bytes = (size_t)p.K * p.K * 2 * p.NF * 64;
    const uint8_t* wbase = reinterpret_cast<const uint8_t*>(weight) + (p.pair ? single_bytes : 0);
    const uint32_t rows = (uint32_t)(p.pair ? p.NF + p.NF / 2 : 2 * p.NF);
    const uint64_t dims[2] = {32, (uint64_t)p.K * p.K * (p.pair ? 2 : 1) * rows};
    const uint64_t str[1] = {64};
    const uint32_t bx[2] = {32, rows};
    int rc = make_map(&tmB, wbase, 2, dims, str, bx, nullptr, 2);
    if (rc) return rc;
  }
  {
    const int f32 = d->out_format == LT_FMT_F32;
    const uint64_t rowb = 128;
    const uint64_t dims[5] = {(uint64_t)(f32 ? 32 : 64), (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.D, (uint64_t)p.N};
    const uint64_t str[4] = {rowb, rowb * p.W, rowb * p.W * p.H, rowb * p.W * p.H * p.D};
    const uint32_t bx[5] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)p.OWt, (uint32_t)p.LINES, 1, 1};
    int rc = make_map(&tmOut, out, 5, dims, str, bx, nullptr, 1, f32);
    if (rc) return rc;
    tmRes = tmOut;
    if (d->residual != LT_RES_NONE) {
      rc = make_map(&tmRes, residual, 5, dims, str, bx, nullptr, 1, f32);
      if (rc) return rc;
    }
  }
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(conv_fold_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_fold_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return fail(LT_ERR_CUDA, "conv_fold: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  long grid;
  if (p.pair) {
    const long pairs = (p.tiles + 1) / 2, P = sm_count() / 2;
    grid = 2 * (pairs < P ? pairs : P);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_fold_kernel<true>, tmA, tmB, tmOut, tmRes, p);
    if (e != cudaSuccess) return fail(LT_ERR_CUDA, "conv_fold_kernel<pair>: %s", cudaGetErrorString(e));
  } else {
    grid = p.tiles < sm_count() ? p.tiles : sm_count();
    conv_fold_kernel<false><<<(unsigned)grid, 320, smem, (cudaStream_t)stream>>>(tmA, tmB, tmOut, tmRes, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LT_ERR_CUDA, "conv_fold_kernel: %s", cudaGetErrorString(e));
  if (want_prof) {   // debug only: synchronises
    unsigned long long h[16];
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const double g = (double)grid;
    fprintf(stderr, "[fold prof K=%d tiles/cta=%.0f] producer: total %.0f wait a_empty %.0f b_empty %.0f | mma: total %.0f wait acc_empty %.0f a_full %.0f b_full %.0f | epi: total %.0f wait acc_full %.0f res %.0f (cycles per CTA)\n",
            p.K, (double)p.tiles / g, h[2] / g, h[0] / g, h[1] / g, h[6] / g, h[3] / g, h[4] / g, h[5] / g, h[9] / g, h[7] / g, h[8] / g);
  }
  return LT_OK;
}

}  // namespace lt

using namespace lt;

extern "C" size_t lt_conv_fold_weight_bytes(int K, int Cout) {
  const int NC = (Cout + 15) & ~15;
  const size_t NF = (size_t)K * NC;
  return (size_t)K * K * 2 * NF * 64 + (size_t)K * K * 3 * NF * 64;     // one-CTA layout + CTA-pair layout
}

extern "C" int lt_conv_fold_pack_weights(const float* w_tap_ci_co, void* packed, int K, int Cout, void* stream) {
  LT_REQUIRE(w_tap_ci_co && packed && (K == 3 || K == 7) && Cout > 0 && Cout <= 32, "conv_fold_pack_weights: bad arguments");
  const int NC = (Cout + 15) & ~15;
  const long total = (long)K * K * K * NC * 32;
  long blocks = (total + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  pack_fold_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_tap_ci_co, reinterpret_cast<sh_t*>(packed), K, Cout, NC);
  LT_CHECK_LAUNCH("pack_fold_weights_kernel");
  sh_t* pair_base = reinterpret_cast<sh_t*>(reinterpret_cast<uint8_t*>(packed) + (size_t)K * K * 2 * K * NC * 64);
  pack_fold_pair_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_tap_ci_co, pair_base, K, Cout, NC);
  LT_CHECK_LAUNCH("pack_fold_pair_weights_kernel");
  return LT_OK;
}
