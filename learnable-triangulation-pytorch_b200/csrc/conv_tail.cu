// Fused tail of the V2V network (reference v2v.py:154-160, 168-169): back_layers[1] (1x1x1 conv 32->32 + BN + ReLU),
// back_layers[2] (same) and output_layer (1x1x1 conv 32->J, bias) as ONE kernel.
//
// Unfused, each of the three point-wise layers reads and writes the whole 64^3 x 32-channel volume (2 x 268 MB per layer at
// B = 8): ~0.41 ms of pure HBM round trips for 5 GFLOP.  Here a CTA streams 128-voxel tiles of the input once (TMA, 16 KB),
// chains the three 128 x 32 x 32 GEMMs on the tensor cores (tcgen05, 3-term split-fp16 products, fp32 accumulators in TMEM)
// with the two hidden activations going TMEM -> registers (scale / shift / ReLU / split) -> a swizzled shared-memory tile that
// is the next GEMM's A operand, and writes only the logits (float32, `FC` floats per voxel: 17 joints rounded up to 20).
// HBM traffic: 128 B in + 80 B out per voxel instead of 3 x 256 B.
//
// Per tile the chain GEMM1 -> act -> GEMM2 -> act -> GEMM3 -> store is serial (~2000 cycles of latency); throughput comes from
// three co-resident CTAs per SM (76 KB of shared memory, 128 TMEM columns each) and the TMA prefetch of the next input tile.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 activation / store warps
// (one TMEM lane quadrant = 32 voxels each; a thread owns one voxel row).
#include "conv_tc_params.cuh"
#include "softargmax_common.cuh"

namespace lt {

struct TailParams {
  const float* scale1; const float* shift1;   // [32] folded BN of back_layers[1]
  const float* scale2; const float* shift2;   // [32] back_layers[2]
  const float* scale3; const float* bias3;    // [32] output layer: 1 / (filter pre-scale) and bias (zero padded)
  float* logits;                              // [rows][FC]
  long rows;
  long tiles;
  int FC;
  // fused statistics pass of the volumetric soft-argmax (op.py:84-96; null coord = off): rows = B x nvox, nvox % 128 == 0
  const float* coord;                         // [rows][3]
  float* partial;                             // [B][gridDim.x][J][5] online-softmax partials (max, sum e, sum e x, sum e y, sum e z)
  int B, J, tiles_per_sample, softmax;
  float mult;
};

constexpr int kTailThreads = 192;
constexpr int kTailStages = 2;
constexpr int kTailWBytes = 32 * 128;        // one 32 x [32 hi | 32 lo] weight tile
// smem: A ring | H tile | W1 W2 W3 | barriers
constexpr int kTailOffH = kTailStages * kATileBytes;
constexpr int kTailOffW = kTailOffH + kATileBytes;
constexpr int kTailOffBar = kTailOffW + 3 * kTailWBytes;
// fused statistics: per activation warp a [32 rows][FC <= 20 floats] logit tile + [32][4] coordinates; the [4][32][5] merge scratch of
// a flush aliases the first warp's tile (three CTAs of 74 880 bytes + 1 KB each still fit the SM's 228 KB)
constexpr int kTailStatMaxFC = 20;
constexpr int kTailOffStat = kTailOffBar + 128;
constexpr int kTailStatWarpBytes = 32 * kTailStatMaxFC * 4 + 32 * 16;
constexpr int kTailStatBytes = 4 * kTailStatWarpBytes;
static_assert(4 * 32 * 5 * 4 <= 32 * kTailStatMaxFC * 4, "merge scratch must fit the first warp's logit tile");
constexpr int kTailSmem = kTailOffBar + 128 + 1024;
constexpr int kTailSmemStats = kTailOffStat + kTailStatBytes + 1024;

__device__ __forceinline__ void tail_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

__device__ __forceinline__ void issue_gemm32(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) {
  // rows = [32 hi | 32 lo] fp16: K slices hi0 +0, hi1 +2, lo0 +4, lo1 +6 (16-byte units); three products into one accumulator
  umma_f16(d, ad, bd, idesc, 0u);
  umma_f16(d, ad + 2, bd + 2, idesc, 1u);
  umma_f16(d, ad, bd + 4, idesc, 1u);
  umma_f16(d, ad + 2, bd + 6, idesc, 1u);
  umma_f16(d, ad + 4, bd, idesc, 1u);
  umma_f16(d, ad + 6, bd + 2, idesc, 1u);
}

// STATS: 0 = logits only, 1 = + softmax statistics, 2 = + ReLU ("volume_softmax: false") statistics.  The logits a warp has just
// produced (32 voxel rows x J joints, one row per lane) are transposed through a warp-private shared-memory tile so that lane j
// folds joint j of the 32 rows into its online-softmax state (4 rows per step: one rescale + four ex2); the state lives in five
// registers per lane for the whole kernel and is written per (sample, CTA) -- the logits are never re-read for the statistics.
template <int STATS>
__global__ void __launch_bounds__(kTailThreads, 3) v2v_tail_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                    const __grid_constant__ CUtensorMap tmW1,
                                                                    const __grid_constant__ CUtensorMap tmW2,
                                                                    const __grid_constant__ CUtensorMap tmW3, const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* h_tile = smem + kTailOffH;
  uint8_t* w_tiles = smem + kTailOffW;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kTailOffBar);   // [2]
  uint64_t* a_empty = a_full + kTailStages;                             // [2]
  uint64_t* d_full = a_empty + kTailStages;                             // [3] accumulator k complete
  uint64_t* h_full = d_full + 3;                                        // [2] hidden tile k written
  uint64_t* w_full = h_full + 2;                                        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTailStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int i = 0; i < 3; ++i) mbar_init(&d_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&h_full[i], 4);     // one arrival per activation warp
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); }
  if (warp == 1) tmem_alloc(tmem_slot, 128u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer: weights once, then the input tiles =================
    if (elect_one()) {
      mbar_expect_tx(w_full, 3u * kTailWBytes);
      tma_load_2d(w_tiles, &tmW1, w_full, 0, 0);
      tma_load_2d(w_tiles + kTailWBytes, &tmW2, w_full, 0, 0);
      tma_load_2d(w_tiles + 2 * kTailWBytes, &tmW3, w_full, 0, 0);
    }
    __syncwarp();
    uint32_t s = 0, ph = 0;
    for (long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      mbar_wait(&a_empty[s], ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&a_full[s], (uint32_t)kATileBytes);
        tma_load_2d(smem + s * kATileBytes, &tmX, &a_full[s], 0, (int)(t * 128));
      }
      __syncwarp();
      if (++s == kTailStages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_f16(32);
    const uint64_t a0 = make_sw128_desc(smem_u32(smem));
    const uint64_t hd = make_sw128_desc(smem_u32(h_tile));
    const uint64_t w1 = make_sw128_desc(smem_u32(w_tiles)), w2 = make_sw128_desc(smem_u32(w_tiles + kTailWBytes)),
                   w3 = make_sw128_desc(smem_u32(w_tiles + 2 * kTailWBytes));
    mbar_wait(w_full, 0);
    uint32_t s = 0, ph = 0, tp = 0;
    for (long t = blockIdx.x; t < p.tiles; t += gridDim.x, tp ^= 1u) {
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        issue_gemm32(tmem_base, a0 + (uint64_t)(s * (kATileBytes >> 4)), w1, idesc);
        umma_commit(&a_empty[s]);
        umma_commit(&d_full[0]);
      }
      __syncwarp();
      mbar_wait(&h_full[0], tp);
      tc_fence_after();
      if (elect_one()) {
        issue_gemm32(tmem_base + 32u, hd, w2, idesc);
        umma_commit(&d_full[1]);
      }
      __syncwarp();
      mbar_wait(&h_full[1], tp);
      tc_fence_after();
      if (elect_one()) {
        issue_gemm32(tmem_base + 64u, hd, w3, idesc);
        umma_commit(&d_full[2]);
      }
      __syncwarp();
      if (++s == kTailStages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ================= activation / store warps (2..5): thread = one voxel row of the tile =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t hbase = smem_u32(h_tile);
    uint32_t tp = 0;
    // fused statistics state
    constexpr bool SM = STATS == 1;
    const int aw = warp - 2;                      // activation warp 0..3
    float* lg_s = reinterpret_cast<float*>(smem + kTailOffStat + aw * kTailStatWarpBytes);     // [32][FC]
    float4* cd_s = reinterpret_cast<float4*>(lg_s + 32 * kTailStatMaxFC);                       // [32] (x, y, z, -)
    float* merge_s = reinterpret_cast<float*>(smem + kTailOffStat);                            // [4][32][5], aliases warp 0's logit tile
    SoftState st;
    st_init(st, SM);
    int cur_b = -1;
    auto flush = [&](int b) {      // CTA merge of the four warps' states -> partial[b][blockIdx.x][j], then reset (CTA-uniform call sites)
      tail_bar_sync();             // every warp has consumed its last tile (the scratch aliases the first warp's)
      float* my = merge_s + (aw * 32 + lane) * 5;
      my[0] = st.m; my[1] = st.d; my[2] = st.sx; my[3] = st.sy; my[4] = st.sz;
      st_init(st, SM);
      tail_bar_sync();
      if (aw == 0 && lane < p.J) {
        SoftState a{merge_s[lane * 5], merge_s[lane * 5 + 1], merge_s[lane * 5 + 2], merge_s[lane * 5 + 3], merge_s[lane * 5 + 4]};
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float* o = merge_s + (w * 32 + lane) * 5;
          SoftState t{o[0], o[1], o[2], o[3], o[4]};
          st_merge(a, t, SM);
        }
        float* dst = p.partial + (((long)b * gridDim.x + blockIdx.x) * p.J + lane) * 5;
        dst[0] = a.m; dst[1] = a.d; dst[2] = a.sx; dst[3] = a.sy; dst[4] = a.sz;
      }
      tail_bar_sync();
    };
    if (STATS) {
      // samples in which this CTA owns no tile still need an (identity) partial for the merge
      for (int i = threadIdx.x - 64; i < p.B * p.J; i += 128) {
        const int b = i / p.J, j = i % p.J;
        const long lo = (long)b * p.tiles_per_sample, G = gridDim.x;
        const long f0 = lo + (((long)blockIdx.x - lo) % G + G) % G;     // first tile >= lo owned by this CTA
        if (!(f0 < lo + p.tiles_per_sample)) {
          float* dst = p.partial + (((long)b * G + blockIdx.x) * p.J + j) * 5;
          dst[0] = SM ? -INFINITY : 0.0f; dst[1] = 0.f; dst[2] = 0.f; dst[3] = 0.f; dst[4] = 0.f;
        }
      }
    }
    // statistics of the tile in lg_s / cd_s (rows [r_lo, r_hi) of this warp's 32): lane j folds joint j, four rows per step.
    // They are computed ONE TILE LATE, in the two bubbles of the next tile's chain (after handing a hidden tile to the MMA warp the
    // activation warp would otherwise just wait for the next accumulator), so the fused pass adds no latency to the GEMM chain.
    auto stats_rows = [&](int r_lo, int r_hi) {
      if (lane < p.J) {
        for (int r0 = r_lo; r0 < r_hi; r0 += 4) {
          float l[4], x[4], y[4], z[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            l[k] = lg_s[(r0 + k) * p.FC + lane] * p.mult;
            const float4 c = cd_s[r0 + k];
            x[k] = c.x; y[k] = c.y; z[k] = c.z;
          }
          st_push4<SM>(st, l, x, y, z);
        }
      }
    };
    bool pending = false;          // lg_s / cd_s hold a tile (of sample cur_b) whose statistics are not folded in yet
    for (long t = blockIdx.x; t < p.tiles; t += gridDim.x, tp ^= 1u) {
#pragma unroll
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(&d_full[layer], tp);
        tc_fence_after();
        uint32_t t0[16], t1[16];
        tmem_ld16_nowait(tlane + (uint32_t)(layer * 32), t0);
        tmem_ld16_nowait(tlane + (uint32_t)(layer * 32 + 16), t1);
        tmem_wait_ld();
        float v0[16], v1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { v0[j] = __uint_as_float(t0[j]); v1[j] = __uint_as_float(t1[j]); }
        const float* sc = layer == 0 ? p.scale1 : p.scale2;
        const float* sh = layer == 0 ? p.shift1 : p.shift2;
        epi_affine16(v0, sc, sh, 0);
        epi_affine16(v1, sc, sh, 16);
#pragma unroll
        for (int j = 0; j < 16; ++j) { v0[j] = fmaxf(v0[j], 0.f); v1[j] = fmaxf(v1[j], 0.f); }
        // the previous reader of the hidden tile (GEMM layer+1 of this tile's predecessor step) has completed: d_full[layer]
        // of THIS step is committed after it in issue order
        epi_store16(hbase, row, 0, LT_FMT_S32, v0);
        epi_store16(hbase, row, 1, LT_FMT_S32, v1);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(&h_full[layer]);
        if (STATS && pending) stats_rows(layer * 16, layer * 16 + 16);
      }
      if (STATS) {
        // the pending tile is folded in; a sample boundary between it and this tile closes the sample's partial (CTA-uniform)
        const int b = (int)(t / p.tiles_per_sample);
        __syncwarp();
        if (b != cur_b) {
          if (cur_b >= 0) flush(cur_b);
          cur_b = b;
        }
      }
      mbar_wait(&d_full[2], tp);
      tc_fence_after();
      {
        uint32_t t0[16], t1[16];
        tmem_ld16_nowait(tlane + 64u, t0);
        tmem_ld16_nowait(tlane + 80u, t1);
        tmem_wait_ld();
        const long vox = t * 128 + row;
        float3 cxyz = make_float3(0.f, 0.f, 0.f);
        if (STATS && vox < p.rows) { const float* cp = p.coord + vox * 3; cxyz = make_float3(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2)); }
        if (vox < p.rows) {
          float* dst = p.logits + vox * p.FC;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale3 + j));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias3 + j));
            const float4 o = make_float4(fmaf(__uint_as_float(t0[j]), a.x, b.x), fmaf(__uint_as_float(t0[j + 1]), a.y, b.y),
                                         fmaf(__uint_as_float(t0[j + 2]), a.z, b.z), fmaf(__uint_as_float(t0[j + 3]), a.w, b.w));
            *reinterpret_cast<float4*>(dst + j) = o;
            if (STATS) *reinterpret_cast<float4*>(lg_s + lane * p.FC + j) = o;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (16 + j < p.FC) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale3 + 16 + j));
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias3 + 16 + j));
              const float4 o = make_float4(fmaf(__uint_as_float(t1[j]), a.x, b.x), fmaf(__uint_as_float(t1[j + 1]), a.y, b.y),
                                           fmaf(__uint_as_float(t1[j + 2]), a.z, b.z), fmaf(__uint_as_float(t1[j + 3]), a.w, b.w));
              *reinterpret_cast<float4*>(dst + 16 + j) = o;
              if (STATS) *reinterpret_cast<float4*>(lg_s + lane * p.FC + 16 + j) = o;
            }
          }
        }
        if (STATS) {
          // rows = B x nvox with nvox % 128 == 0: every tile is full and lies inside one sample
          cd_s[lane] = make_float4(cxyz.x, cxyz.y, cxyz.z, 0.f);
          __syncwarp();
          pending = true;
        }
      }
      tc_fence_before();   // accumulator reads of this tile are ordered before the next tile's h_full arrivals
    }
    if (STATS) {
      if (pending) { stats_rows(0, 32); __syncwarp(); }
      if (cur_b >= 0) flush(cur_b);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128u);
  }
}

}  // namespace lt

using namespace lt;

// shared launcher: stats = 0 (logits only), 1 (softmax statistics), 2 (ReLU statistics); returns the grid size through *grid_out
static int launch_tail(const void* x, const void* w1, const void* w2, const void* w3, lt::TailParams& p, int stats, int* grid_out, void* stream) {
  using namespace lt;
  CUtensorMap tmX, tmW[3];
  {
    const uint64_t dims[2] = {64, (uint64_t)p.rows};
    const uint64_t str[1] = {128};
    const uint32_t bx[2] = {64, 128};
    int rc = make_map(&tmX, x, 2, dims, str, bx, nullptr, 1);
    if (rc) return rc;
  }
  const void* ws[3] = {w1, w2, w3};
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[2] = {64, 128};
    const uint64_t str[1] = {128};
    const uint32_t bx[2] = {64, 32};
    int rc = make_map(&tmW[i], ws[i], 2, dims, str, bx, nullptr, 1);
    if (rc) return rc;
  }
  p.tiles = (p.rows + 127) / 128;
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(v2v_tail_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(v2v_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemStats);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(v2v_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemStats);
    if (e != cudaSuccess) return fail(LT_ERR_CUDA, "v2v_tail: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  long grid = 3L * sm_count();
  if (grid > p.tiles) grid = p.tiles;
  if (grid_out) *grid_out = (int)grid;
  cudaStream_t st = (cudaStream_t)stream;
  if (stats == 1) v2v_tail_kernel<1><<<(unsigned)grid, kTailThreads, kTailSmemStats, st>>>(tmX, tmW[0], tmW[1], tmW[2], p);
  else if (stats == 2) v2v_tail_kernel<2><<<(unsigned)grid, kTailThreads, kTailSmemStats, st>>>(tmX, tmW[0], tmW[1], tmW[2], p);
  else v2v_tail_kernel<0><<<(unsigned)grid, kTailThreads, kTailSmem, st>>>(tmX, tmW[0], tmW[1], tmW[2], p);
  LT_CHECK_LAUNCH("v2v_tail_kernel");
  return LT_OK;
}

// x: split-fp16 rows [rows][32 hi | 32 lo]; w1/w2/w3: lt_conv_pair_pack_weights(taps = 1, Cin = 32, Cout = 32 / 32 / J) buffers
// (rows padded to 128; the first 32 are used); scale/shift: folded BN of the two hidden layers; scale3 / bias3 [32]: output affine
// (scale3 = 1 / filter pre-scale, bias zero padded);
// logits float32 [rows][FC], FC % 4 == 0, J <= FC <= 32.
extern "C" int lt_v2v_tail_fwd(const void* x, const void* w1, const void* w2, const void* w3, const float* scale1, const float* shift1,
                               const float* scale2, const float* shift2, const float* scale3, const float* bias3, float* logits, long rows,
                               int FC, void* stream) {
  LT_REQUIRE(x && w1 && w2 && w3 && scale1 && shift1 && scale2 && shift2 && scale3 && bias3 && logits, "v2v_tail: null pointer");
  LT_REQUIRE(rows > 0 && rows < (1L << 31) && FC % 4 == 0 && FC >= 4 && FC <= 32, "v2v_tail: bad sizes (rows=%ld FC=%d)", rows, FC);
  TailParams p;
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2; p.scale3 = scale3; p.bias3 = bias3;
  p.logits = logits; p.rows = rows; p.FC = FC;
  p.coord = nullptr; p.partial = nullptr; p.B = 0; p.J = 0; p.tiles_per_sample = 1; p.softmax = 0; p.mult = 1.0f;
  return launch_tail(x, w1, w2, w3, p, 0, nullptr, stream);
}

// Same kernel with the statistics pass of the volumetric soft-argmax (op.py:84-96) fused into the epilogue that produces the logits:
// rows = B x nvox (nvox % 128 == 0, FC <= 20), coord [B][nvox][3]; `workspace` (lt_softargmax3d_workspace_bytes) receives the
// online-softmax partials [B][*n_partials][J][5]; lt_softargmax3d_finish_fwd(..., G = *n_partials, ...) then merges them into the key
// points and writes the normalised volumes.  softmax: 1 = softmax, 0 = ReLU ("volume_softmax: false").
extern "C" int lt_v2v_tail_stats_fwd(const void* x, const void* w1, const void* w2, const void* w3, const float* scale1, const float* shift1,
                                     const float* scale2, const float* shift2, const float* scale3, const float* bias3, float* logits, int B,
                                     long nvox, int FC, const float* coord, int J, float multiplier, int softmax, void* workspace,
                                     size_t workspace_bytes, int* n_partials, void* stream) {
  LT_REQUIRE(x && w1 && w2 && w3 && scale1 && shift1 && scale2 && shift2 && scale3 && bias3 && logits && coord && workspace && n_partials,
             "v2v_tail_stats: null pointer");
  LT_REQUIRE(B > 0 && nvox > 0 && nvox % 128 == 0 && (long)B * nvox < (1L << 31), "v2v_tail_stats: bad sizes (B=%d nvox=%ld)", B, nvox);
  LT_REQUIRE(FC % 4 == 0 && FC >= 4 && FC <= kTailStatMaxFC && J > 0 && J <= FC, "v2v_tail_stats: need J <= FC <= %d, FC %% 4 == 0", kTailStatMaxFC);
  LT_REQUIRE(softmax == 0 || softmax == 1, "v2v_tail_stats: mode must be 0 (ReLU) or 1 (softmax)");
  LT_REQUIRE(workspace_bytes >= lt_softargmax3d_workspace_bytes(B, J, nvox), "v2v_tail_stats: workspace too small");
  TailParams p;
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2; p.scale3 = scale3; p.bias3 = bias3;
  p.logits = logits; p.rows = (long)B * nvox; p.FC = FC;
  p.coord = coord; p.partial = reinterpret_cast<float*>(workspace); p.B = B; p.J = J; p.tiles_per_sample = (int)(nvox / 128);
  p.softmax = softmax; p.mult = multiplier;
  return launch_tail(x, w1, w2, w3, p, softmax ? 1 : 2, n_partials, stream);
}
