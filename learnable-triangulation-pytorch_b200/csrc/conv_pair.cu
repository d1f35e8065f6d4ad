// CTA-pair tensor-core implicit-GEMM convolution for sm_100a (tcgen05 cta_group::2).
//
// Why: the one-CTA-per-tile kernel (conv_tc.cu) moves 32 KB of operands per 128x128x32 chunk and is bound by the SM's TMA
// ingest (~57 B/clk for 128-byte rows, ~44 B/clk for the 5-D activation boxes; tools/tma_probe.cu, profiles/r02_probes.md),
// not by the tensor pipe.  Here two CTAs of a cluster (one TPC) compute a 256 x Nt tile together: each CTA loads its own
// 128-position activation tile and HALF of the weight tile, one thread of the leader issues M = 256 MMAs that read both
// halves, and each CTA's accumulator (its 128 rows x Nt fp32 columns) lives in its own TMEM.  At Nt = 256 a CTA ingests
// 32 KB per 768 math cycles (42 B/clk) instead of 32 KB per 384.
//
// Precision: split-fp16 operands with UNSCALED low parts (common.cuh), three products per 16-wide K slice.  hi*hi accumulates into
// D1, hi*lo + lo*hi into D2 (Nt columns further), and the epilogue adds them in fp32 RN (lt_options.pair_two_acc = 1, default): tcgen05
// truncates every accumulation step, so the big accumulator must see as few steps as possible (pair_plan()).  2 x Nt columns per
// stage: two stages for Nt = 128, ONE for Nt = 256 (then the epilogue of tile i does not overlap the main loop of tile i+1).
// pair_two_acc = 0: all three products into one accumulator of Nt columns, two stages.  Weights: [tap][Cin/32][CoutP rows][32 hi | 32 lo]
// fp16 = 128-byte rows, 128B swizzle, so the K slices of both operands are descriptor offsets (+0,+2 hi; +4,+6 lo, in 16-byte units).
//
// Persistent: grid = 2 x min(#SM / 2, tiles); pair k walks tiles k, k + P, ...  (tile = (pair of M tiles, N tile), N fastest
// so that concurrently running pairs share the activation tile in L2).
//
// Warp roles per CTA (608 threads): warp 0 TMA producer (both CTAs), warp 1 TMEM allocator + MMA issuer (leader CTA only),
// warps 2..17 two epilogue groups of 8 warps (both CTAs, each draining its own TMEM lanes), warp 18 residual producer.
// Staged epilogue (float32 outputs; group 0 only): tcgen05.ld -> scale/shift (+ residual tile that arrived by TMA) -> ReLU ->
// repack into a swizzled smem tile -> TMA store (also implements the stride-phase mapping).
//
// Barriers (same smem offsets in both CTAs):
//   full[s]       leader's copy is used: 1 arrival (leader producer's expect_tx of both CTAs' bytes) + complete_tx of all four
//                 TMA loads (the non-leader's loads signal the leader's barrier through the .cta_group::2 form)
//   empty[s]      per CTA; tcgen05.commit multicast from the leader frees the slot in both CTAs
//   acc_full[a]   per CTA; commit multicast when a tile's last MMA has completed
//   acc_empty[a]  leader's copy is used: 16 (staged) / 32 (direct) arrivals = epilogue warps x 2 CTAs (remote mbarrier.arrive from the peer)
//   res_full[b] / res_empty[b]  per CTA: residual staging buffer b filled by TMA / read by the 8 warps of the group that owns it
//
// Direct-store epilogue (split-fp16 outputs): no CTA-level barrier at all.  The per-channel scale / shift sit in shared memory
// (loaded once per CTA), a dedicated warp (18) streams the residual tiles through the four staging buffers under the full/empty
// barriers, and every epilogue warp runs tcgen05.ld -> affine -> residual -> ReLU -> split -> st.global on its own.  Measured
// with ncu source counters on the 1x1 256 -> 1024 layer at 24x24 (profiles/r02c_pair_epilogue.md): the previous epilogue spent 27 %
// of its stall samples on the scale / shift LDGs (L1-miss latency + LSU queue throttle behind the 32-sector row stores), 10 % in the
// per-block bar.sync that recycled the residual buffer and 9 % in the MEMBAR.GPU + ERRBAR of a release.cluster remote arrive.
#include "conv_tc_params.cuh"
#include "pair_common.cuh"

namespace lt {

struct PairExtra {
  int n_tiles;
  long m_tiles, m_pairs, total_tiles;
  int stage_bytes, off_out, off_res, off_bar;
  int b_rows;   // weight rows per chunk = CoutP
  int res_bufs; // residual staging buffers (power of two): 16 KB blocks requested this many blocks ahead
  int direct_out;   // split-fp16 output rows are stored straight from registers
  int off_aff, aff_n;   // per-channel scale [aff_n] and shift [aff_n] copied to shared memory at kernel start
  int two_acc;          // 1: hi*hi -> D1, the two cross products -> D2 (Nt columns further); the epilogue adds them in fp32 RN
  int acc_stages;       // accumulator stages in TMEM: 512 / (Nt * (1 + two_acc)), at most 2
  unsigned long long* prof;   // optional [16] cycle counters (lt_options.pair_prof): time each role spends waiting, summed over CTAs
};

__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long& acc, bool on) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

constexpr int kPairEpiWarps = 8;       // per group
constexpr int kPairThreads = 64 + 2 * kPairEpiWarps * 32 + 32;   // producer + MMA warp + two epilogue groups of 8 warps + residual producer
constexpr int kPairResWarp = 2 + 2 * kPairEpiWarps;
__device__ __forceinline__ void epi_bar_sync_g(int grp) {
  if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}
constexpr int kMaxResBufs = 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ OutMaps tmOut, const __grid_constant__ OutMaps tmRes, const TcParams p,
                 const PairExtra x) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + x.off_out;   // 2 x 16 KB
  uint8_t* res_stage = smem + x.off_res;   // res_bufs x 16 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + x.off_bar);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_full = empty + p.stages;   // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* res_full = acc_empty + 2;      // [kMaxResBufs]
  uint64_t* res_empty = res_full + kMaxResBufs;   // [kMaxResBufs] (direct-store epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + kMaxResBufs);
  float* s_scale = reinterpret_cast<float*>(smem + x.off_aff);
  float* s_shift = s_scale + x.aff_n;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nchunks = p.KD * p.KH * p.KW * p.CB;
  const int stage_bytes = x.stage_bytes;
  const bool prof_on = x.prof != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kPairEpiWarps * (x.direct_out ? 2 : 1)); }
    for (int i = 0; i < kMaxResBufs; ++i) { mbar_init(&res_full[i], 1); mbar_init(&res_empty[i], kPairEpiWarps); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < x.aff_n; i += kPairThreads) { s_scale[i] = __ldg(p.scale + i); s_shift[i] = __ldg(p.shift + i); }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmOut.m[0]); }
  if (warp == 1) tmem_alloc2(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // both CTAs' barriers initialised and TMEM allocated before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // persistent schedule: pair k of P walks tiles k, k + P, ...; tile -> (M-tile pair, N tile), N fastest
  const long pair_idx = (long)(blockIdx.x >> 1), n_pairs = (long)(gridDim.x >> 1);
  auto decode = [&](long tile, int& ow0, int& oh0, int& od0, int& nb0, int& n0) {
    n0 = (int)(tile % x.n_tiles) * p.Nt;
    long t = (tile / x.n_tiles) * 2 + rank;      // this CTA's 128-position M tile; t >= m_tiles decodes to nb0 >= N (all OOB)
    ow0 = (int)(t % p.tw) * p.bw; t /= p.tw;
    oh0 = (int)(t % p.th) * p.bh; t /= p.th;
    od0 = (int)(t % p.td) * p.bd; t /= p.td;
    nb0 = (int)t * p.bn;
  };

  // residual tiles: block c of this CTA's tile sequence (c runs across tiles) lives in staging buffer c % RB
  const int nblk = p.Nt >> 5;
  const int esz = (p.out_format == LT_FMT_F32) ? 1 : 2;
  const bool has_res = p.residual != LT_RES_NONE;
  const long my_tiles = (x.total_tiles > pair_idx) ? (x.total_tiles - 1 - pair_idx) / n_pairs + 1 : 0;
  const long total_blocks = my_tiles * nblk;
  // The residual tiles stream through RB 16 KB buffers indexed by a block counter that runs across tiles: block c lives in
  // buffer c % RB and is requested RB blocks ahead (a K-short 1x1 expand layer reads as many residual bytes as operand bytes:
  // with two buffers its epilogue ran at the latency of one TMA round trip per block)
  const int RB = x.res_bufs, rb_shift = (RB == 4) ? 2 : 1;
  auto issue_res = [&](long c) {   // one thread: residual block c of this CTA's tile sequence
    const long tile_c = pair_idx + (c / nblk) * n_pairs;
    const int blk = (int)(c % nblk);
    int a0, a1, a2, a3, an;
    decode(tile_c, a0, a1, a2, a3, an);
    const int buf = (int)(c & (RB - 1));
    int ch = an + blk * 32, mi = 0;
    if (p.n_maps > 1) { mi = ch / p.oc; ch -= mi * p.oc; }
    mbar_expect_tx(&res_full[buf], 16384u);
    tma_load_5d(res_stage + buf * 16384, &tmRes.m[mi], &res_full[buf], ch * esz, a0, a1, a2, a3);
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own activation tile + own half of the weight tile) =================
    uint32_t rs = 0, rph = 0;
    unsigned long long w_empty = 0;
    const long long t_start = prof_on ? clock64() : 0;
    const uint32_t full0 = map_to_cta(smem_u32(&full[0]), 0);   // leader's full[] in cluster address space
    const int b_half = p.Nt >> 1;
    for (long tile = pair_idx; tile < x.total_tiles; tile += n_pairs) {
      int ow0, oh0, od0, nb0, n0;
      decode(tile, ow0, oh0, od0, nb0, n0);
      const int ax = ow0 * p.sw - p.pw, ay = oh0 * p.sh - p.ph, az = od0 * p.sd - p.pd;
      const int brow0 = n0 + (int)rank * b_half;
      int cb = 0, kw = 0, kh = 0, kd = 0;
      for (int q = 0; q < nchunks; ++q) {
        mbar_wait_t(&empty[rs], rph ^ 1u, w_empty, prof_on);
        uint8_t* a_dst = smem + (size_t)rs * stage_bytes;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full[rs], 2u * (uint32_t)stage_bytes);
          const uint32_t bar = full0 + rs * 8u;
          tma2_load_5d(a_dst, &tmA, bar, cb * 64, ax + kw, ay + kh, az + kd, nb0);
          tma2_load_2d(a_dst + kATileBytes, &tmB, bar, 0, q * x.b_rows + brow0);
        }
        __syncwarp();
        if (++cb == p.CB) { cb = 0; if (++kw == p.KW) { kw = 0; if (++kh == p.KH) { kh = 0; ++kd; } } }
        if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1u; }
      }
    }
    if (prof_on && lane == 0) { atomicAdd(&x.prof[0 + 8 * rank], w_empty); atomicAdd(&x.prof[1 + 8 * rank], (unsigned long long)(clock64() - t_start)); }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only; M = 256 spans both CTAs) =================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_f16_m256(p.Nt);
      uint32_t rs = 0, rph = 0, it = 0;
      unsigned long long w_acc = 0, w_full = 0;
      const long long t_start = prof_on ? clock64() : 0;
      const uint64_t ad0 = make_sw128_desc(smem_u32(smem));
      const uint64_t bd0 = make_sw128_desc(smem_u32(smem) + kATileBytes);
      const uint64_t sdelta = (uint64_t)(stage_bytes >> 4);
      for (long tile = pair_idx; tile < x.total_tiles; tile += n_pairs, ++it) {
        const uint32_t as = x.acc_stages == 2 ? (it & 1u) : 0u, aph = x.acc_stages == 2 ? ((it >> 1) & 1u) : (it & 1u);
        mbar_wait_t(&acc_empty[as], aph ^ 1u, w_acc, prof_on);
        tc_fence_after();
        const uint32_t d = tmem_base + as * (uint32_t)(p.Nt << x.two_acc);
        const uint32_t d2 = x.two_acc ? d + (uint32_t)p.Nt : d;     // cross products (see the header comment on accumulation)
        for (int q = 0; q < nchunks; ++q) {
          mbar_wait_t(&full[rs], rph, w_full, prof_on);
          tc_fence_after();
          const uint64_t ad = ad0 + sdelta * rs, bd = bd0 + sdelta * rs;
          if (elect_one()) {
            // rows = [32 hi | 32 lo] fp16: slices hi0 +0, hi1 +2, lo0 +4, lo1 +6 (16-byte units)
            umma2_f16(d, ad, bd, idesc, q == 0 ? 0u : 1u);   // hi * hi
            umma2_f16(d, ad + 2, bd + 2, idesc, 1u);
            umma2_f16(d2, ad, bd + 4, idesc, (q == 0 && x.two_acc) ? 0u : 1u);   // hi * lo
            umma2_f16(d2, ad + 2, bd + 6, idesc, 1u);
            umma2_f16(d2, ad + 4, bd, idesc, 1u);            // lo * hi
            umma2_f16(d2, ad + 6, bd + 2, idesc, 1u);
            umma2_commit_mc(&empty[rs]);
            if (q == nchunks - 1) umma2_commit_mc(&acc_full[as]);
          }
          __syncwarp();
          if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1u; }
        }
      }
      if (prof_on && lane == 0) { atomicAdd(&x.prof[2], w_acc); atomicAdd(&x.prof[3], w_full); atomicAdd(&x.prof[4], (unsigned long long)(clock64() - t_start)); }
    }
  } else if (warp == kPairResWarp) {
    // ================= residual producer (direct-store epilogue): block c -> buffer c % RB once its previous reader group is done ==========
    if (x.direct_out && has_res) {
      for (long c = 0; c < total_blocks; ++c) {
        mbar_wait(&res_empty[c & (RB - 1)], (uint32_t)(((c >> rb_shift) & 1) ^ 1));
        if (elect_one()) issue_res(c);
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (both CTAs; each CTA drains its own 128 TMEM lanes) =================
    // Two groups of 8 warps (2..9, 10..17).  Direct-store mode: group g takes the 32-channel blocks i = g, g + 2, ... of a tile
    // (measured with one group: ~1700 cycles per block of dependent tcgen05.ld -> FFMA -> cvt -> store work at two warps per
    // scheduler = 13.6k cycles per 256-wide tile against a 6-10k cycle main loop of a K = 256 layer: epilogue-bound).
    // Staged mode (float32 outputs): group 0 only.
    const int grp = (warp - 2) >> 3;
    if (grp == 1 && !x.direct_out) goto teardown;
    const int quad = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = quad * 32 + lane;
    const bool leader = threadIdx.x == 64 + grp * 256;
    const uint32_t acc_empty0 = map_to_cta(smem_u32(&acc_empty[0]), 0);
    if (leader && has_res && !x.direct_out)
      for (long c = 0; c < RB && c < total_blocks; ++c) issue_res(c);
    uint32_t it = 0;
    long c = 0;
    unsigned long long w_accf = 0, w_res = 0;
    const long long t_start = prof_on ? clock64() : 0;
    for (long tile = pair_idx; tile < x.total_tiles; tile += n_pairs, ++it) {
      int ow0, oh0, od0, nb0, n0;
      decode(tile, ow0, oh0, od0, nb0, n0);
      const uint32_t as = x.acc_stages == 2 ? (it & 1u) : 0u, aph = x.acc_stages == 2 ? ((it >> 1) & 1u) : (it & 1u);
      mbar_wait_t(&acc_full[as], aph, w_accf, prof_on);
      tc_fence_after();
      const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)(p.Nt << x.two_acc);
      if (x.direct_out) {
        // ---- direct epilogue (split-fp16 outputs): this thread's voxel row goes from registers to global memory as 2 x 32
        // bytes per 32-channel block (high halves, low halves).  No staging tile, no TMA store: the TMA queue of the SM is
        // FIFO, so a store issued here would wait behind every operand load the producer has already queued (up to 160 KB,
        // ~3000 cycles) and throttle the epilogue of K-short layers to one block per queue latency (measured: 1x1 256 -> 1024
        // with residual at 24x24: 60 us staged).  Residual tiles still arrive by TMA, four blocks ahead.
        int r_ = row;
        const int dw = r_ % p.bw; r_ /= p.bw;
        const int dh = r_ % p.bh; r_ /= p.bh;
        const int dd = r_ % p.bd; r_ /= p.bd;
        const int ow = ow0 + dw, oh = oh0 + dh, od = od0 + dd, nb = nb0 + r_;
        const bool valid = ow < p.OW && oh < p.OH && od < p.OD && nb < p.N;
        const long opix = (((long)nb * p.FD + (od * p.osd + p.ood)) * p.FH + (oh * p.osh + p.ooh)) * p.FW + (ow * p.osw + p.oow);
        for (int i = grp; i < nblk; i += 2) {
          const long cb = c + i;                 // block counter across this CTA's tiles
          const int rbuf = (int)(cb & (RB - 1));
          float v[16], r[16];
          {
            uint32_t t1[16];
            if (x.two_acc) {
              uint32_t t2[16];
              tmem_ld16_nowait(tlane + (uint32_t)(i * 32 + half * 16), t1);
              tmem_ld16_nowait(tlane + (uint32_t)(p.Nt + i * 32 + half * 16), t2);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(t1[j]) + __uint_as_float(t2[j]);
            } else {
              tmem_ld16(tlane + (uint32_t)(i * 32 + half * 16), t1);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(t1[j]);
            }
          }
          if (i + 2 >= nblk) {                   // this group's last block of the tile: its part of the accumulator is drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc_empty0 + as * 8u);
          }
          epi_affine16_smem(v, s_scale, s_shift, n0 + i * 32 + half * 16);
          if (has_res) {
            mbar_wait_t(&res_full[rbuf], (uint32_t)((cb >> rb_shift) & 1), w_res, prof_on);
            epi_load16(smem_u32(res_stage + rbuf * 16384), row, half, LT_FMT_S32, r);
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&res_empty[rbuf]);   // this warp is done with res_stage[rbuf]
          }
          epi_activate16(v, r, p.residual, p.relu);
          if (valid) {
            int ch = n0 + i * 32;
            long pix = opix;
            if (p.n_maps > 1) {
              const int mi = ch / p.oc;
              ch -= mi * p.oc;
              pix += ((long)(mi / (p.gh * p.gw)) * p.FH + (mi / p.gw) % p.gh) * p.FW + mi % p.gw;
            }
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + (pix * p.FC + ch) * 4 + half * 32;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split_s32x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            stg256(dst, hi);
            stg256(dst + 64, lo);
          }
        }
        c += nblk;
        continue;
      }
      for (int i = 0; i < nblk; ++i, ++c) {
        const int buf = (int)(c & 1);              // output staging buffer
        const int rbuf = (int)(c & (RB - 1));      // residual staging buffer
        float v[16], r[16];
        {
          uint32_t t1[16];
          if (x.two_acc) {
            uint32_t t2[16];
            tmem_ld16_nowait(tlane + (uint32_t)(i * 32 + half * 16), t1);
            tmem_ld16_nowait(tlane + (uint32_t)(p.Nt + i * 32 + half * 16), t2);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(t1[j]) + __uint_as_float(t2[j]);
          } else {
            tmem_ld16(tlane + (uint32_t)(i * 32 + half * 16), t1);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(t1[j]);
          }
        }
        if (i == nblk - 1) {               // accumulator stage fully read by this warp: release it to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(acc_empty0 + as * 8u);
        }
        epi_affine16_smem(v, s_scale, s_shift, n0 + i * 32 + half * 16);
        if (has_res) {
          mbar_wait(&res_full[rbuf], (uint32_t)((c >> rb_shift) & 1));
          epi_load16(smem_u32(res_stage + rbuf * 16384), row, half, p.out_format, r);
        }
        epi_activate16(v, r, p.residual, p.relu);
        if (leader) bulk_wait_read<1>();      // the store that last read out_stage[buf] (two blocks ago) is done with it
        epi_bar_sync();                       // also: every thread has finished reading res_stage[rbuf]
        epi_store16(smem_u32(out_stage + buf * 16384), row, half, p.out_format, v);
        fence_proxy_async();
        epi_bar_sync();
        if (leader) {
          int ch = n0 + i * 32, mi = 0;
          if (p.n_maps > 1) { mi = ch / p.oc; ch -= mi * p.oc; }
          tma_store_5d(&tmOut.m[mi], out_stage + buf * 16384, ch * esz, ow0, oh0, od0, nb0);
          bulk_commit();
          if (has_res && c + RB < total_blocks) issue_res(c + RB);
        }
      }
    }
    if (leader) bulk_wait<0>();
    if (prof_on && threadIdx.x == 64) { atomicAdd(&x.prof[5 + 8 * rank], w_accf); atomicAdd(&x.prof[6 + 8 * rank], w_res); atomicAdd(&x.prof[7 + 8 * rank], (unsigned long long)(clock64() - t_start)); }
  }

teardown:
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer may still read this CTA's operand half / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
bool pair_plan(const lt_conv_desc* d, const TcParams& p, int CoutP, PairPlan* plan) {
  if (!p.tma_epi || p.terms != 3 || CoutP % 128 != 0) return false;
  const long m_tiles = (long)p.tw * p.th * p.td * p.tn;
  const long m_pairs = (m_tiles + 1) / 2;
  const int nchunks = p.KD * p.KH * p.KW * p.CB;
  const int P = sm_count() / 2;
  // pick the N tile that minimises (waves) x (cycles per chunk): math 3*Nt, TMA ingest ~ (16 KB + Nt*64 B) / 50 B/clk,
  // issue ~ 450 cycles per chunk
  double best = 1e300;
  int best_nt = 0;
  for (int nt = 256; nt >= 128; nt >>= 1) {
    if (CoutP % nt) continue;
    if (opts().pair_nt != 0 && opts().pair_nt != nt && CoutP % opts().pair_nt == 0) continue;   // A/B override
    const long tiles = m_pairs * (CoutP / nt);
    const double waves = (double)((tiles + P - 1) / P);
    const double per_chunk = fmax(fmax(3.0 * nt, (16384.0 + nt * 64.0) / 50.0), 450.0);
    const double epi = 250.0 * (nt / 32);     // exposed only when the main loop is shorter
    const double t = waves * fmax(per_chunk * nchunks, epi) + epi;
    if (t < best) { best = t; best_nt = nt; }
  }
  if (!best_nt) return false;
  // small grids: latency-bound; the one-CTA kernel has the cheaper prologue and a split-K path that spreads the K loop over
  // the SMs (measured: 2x2 taps 2048 -> 256 at 12x12 = 36 pair tiles with 256 chunks each: 70 us here, 59 us there)
  const long best_tiles = m_pairs * (CoutP / best_nt);
  if (best_tiles * 4 < P || (best_tiles * 2 <= P && nchunks >= 64)) return false;
  // Accumulation accuracy.  tcgen05 adds every MMA into the fp32 accumulator with TRUNCATION (tools/accum_probe.py: zero-mean operands,
  // rms error 0.027 K x 2^-24 of the result scale, i.e. proportional to the number of MMA steps instead of its square root).  With one
  // accumulator the three products of a 16-channel slice are three such steps on the big accumulator; keeping the two cross products
  // (2^-11 of the main term) in their own accumulator leaves one.  Price: twice the TMEM columns -- with Nt = 256 only ONE accumulator
  // stage fits, so the epilogue of a tile no longer overlaps the main loop of the next (+0.3-0.4 ms per step, mostly on the 1x1
  // 256 -> 1024 layers).  Measured at config #2, B = 8 against the CPU oracle (profiles/r02i_accumulation.md): volumes 1.41e-3 with one
  // accumulator (contract 1e-3 missed), 7.9e-4 with two.
  plan->two_acc = opts().pair_two_acc ? 1 : 0;
  plan->acc_stages = (best_nt << plan->two_acc) * 2 <= 512 ? 2 : 1;
  plan->Nt = best_nt;
  plan->n_tiles = CoutP / best_nt;
  plan->m_tiles = m_tiles;
  plan->m_pairs = m_pairs;
  const int stage_bytes = kATileBytes + best_nt * 64;
  // shared memory: operand ring + 2 output staging tiles + (residual layers) 4 residual staging tiles, 16 KB each
  plan->res_bufs = d->residual != LT_RES_NONE ? kMaxResBufs : 0;
  plan->direct_out = (d->out_format == LT_FMT_S32 && opts().pair_direct_out) ? 1 : 0;
  int stages = (227 * 1024 - 1024 - (plan->direct_out ? 0 : 32768) - plan->res_bufs * 16384 - 8 * CoutP - 512) / stage_bytes;
  if (stages > 8) stages = 8;
  if (opts().pair_stages >= 2 && opts().pair_stages < stages) stages = opts().pair_stages;       // A/B override
  plan->stages = stages;
  const long tiles = m_pairs * plan->n_tiles;
  plan->grid = 2u * (unsigned)(tiles < P ? tiles : P);
  return true;
}

int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmOut, const OutMaps& tmRes, TcParams& p,
                const PairPlan& plan, int CoutP, cudaStream_t st) {
  PairExtra x;
  p.Nt = plan.Nt;
  p.stages = plan.stages;
  p.splits = 1; p.ws = nullptr; p.ws_ld = 0;
  x.n_tiles = plan.n_tiles;
  x.m_tiles = plan.m_tiles;
  x.m_pairs = plan.m_pairs;
  x.total_tiles = plan.m_pairs * plan.n_tiles;
  x.stage_bytes = kATileBytes + plan.Nt * 64;
  x.b_rows = CoutP;
  const int ring = plan.stages * x.stage_bytes;
  x.off_out = (ring + 1023) & ~1023;
  x.direct_out = plan.direct_out;
  x.two_acc = plan.two_acc;
  x.acc_stages = plan.acc_stages;
  x.off_res = x.off_out + (plan.direct_out ? 0 : 32768);
  x.off_aff = x.off_res + plan.res_bufs * 16384;      // scale [CoutP] | shift [CoutP]
  x.aff_n = CoutP;
  x.off_bar = x.off_aff + 8 * CoutP;
  x.res_bufs = plan.res_bufs ? plan.res_bufs : 2;
  const size_t smem = (size_t)x.off_bar + (2 * plan.stages + 4 + 2 * kMaxResBufs) * 8 + 16 + 1024;
  if (smem > 227 * 1024) return fail(LT_ERR_INVALID, "conv_pair: shared memory budget exceeded (%zu)", smem);
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return fail(LT_ERR_CUDA, "conv_pair: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  x.prof = nullptr;
  static unsigned long long* prof_buf = nullptr;
  const bool want_prof = opts().pair_prof != 0;
  if (want_prof) {
    if (!prof_buf) cudaMalloc(&prof_buf, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(prof_buf, 0, 16 * sizeof(unsigned long long), st);
    x.prof = prof_buf;
  }
  conv_pair_kernel<<<plan.grid, kPairThreads, smem, st>>>(tmA, tmB, tmOut, tmRes, p, x);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LT_ERR_CUDA, "conv_pair_kernel: %s", cudaGetErrorString(e));
  if (want_prof) {   // debug only: synchronises
    unsigned long long h[16];
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const double g = (double)(plan.grid / 2);
    fprintf(stderr, "[pair prof Nt=%d chunks=%d tiles/pair=%.1f stages=%d] leader: producer wait empty %.0f of %.0f | mma wait acc_empty %.0f full %.0f of %.0f | "
            "epi wait acc_full %.0f res %.0f of %.0f || peer: producer wait empty %.0f of %.0f | epi wait acc_full %.0f res %.0f of %.0f (cycles per CTA)\n",
            p.Nt, p.KD * p.KH * p.KW * p.CB, (double)x.total_tiles / g, p.stages, h[0] / g, h[1] / g, h[2] / g, h[3] / g, h[4] / g, h[5] / g, h[6] / g,
            h[7] / g, h[8] / g, h[9] / g, h[13] / g, h[14] / g, h[15] / g);
  }
  return LT_OK;
}

// ---- weight packing: fp32 [taps][Cin][Cout] -> fp16 [taps][Cin/32][CoutP][32 hi | 32 lo] (128-byte rows) ----------
__global__ void __launch_bounds__(256) pack_weights_pair_kernel(const float* __restrict__ w, sh_t* __restrict__ out,
                                                                int taps, int Cin, int Cout, int CoutP) {
  const int CB = Cin / 32;
  const long total = (long)taps * CB * CoutP * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % 32);
    long r = i / 32;
    const int n = (int)(r % CoutP); r /= CoutP;
    const int cb = (int)(r % CB);
    const int tap = (int)(r / CB);
    const float v = (n < Cout) ? w[((long)tap * Cin + cb * 32 + j) * Cout + n] : 0.0f;
    sh_t hi, lo;
    split_s32(v, hi, lo);
    sh_t* rowp = out + ((((long)tap * CB + cb) * CoutP + n) << 6);
    rowp[j] = hi;
    rowp[32 + j] = lo;
  }
}

}  // namespace lt

using namespace lt;

extern "C" size_t lt_conv_pair_weight_bytes(int taps, int Cin, int Cout) {
  const int CoutP = (Cout + 127) & ~127;
  return (size_t)taps * (Cin / 32) * CoutP * 128;
}

extern "C" int lt_conv_pair_pack_weights(const float* w, void* packed, int taps, int Cin, int Cout, void* stream) {
  LT_REQUIRE(w && packed, "conv_pair_pack_weights: null pointer");
  LT_REQUIRE(Cin % 32 == 0 && taps > 0 && Cout > 0, "conv_pair_pack_weights: bad sizes");
  const int CoutP = (Cout + 127) & ~127;
  const long total = (long)taps * (Cin / 32) * CoutP * 32;
  long blocks = (total + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  pack_weights_pair_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<sh_t*>(packed), taps, Cin, Cout, CoutP);
  LT_CHECK_LAUNCH("pack_weights_pair_kernel");
  return LT_OK;
}
