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

namespace lt {

struct TailParams {
  const float* scale1; const float* shift1;   // [32] folded BN of back_layers[1]
  const float* scale2; const float* shift2;   // [32] back_layers[2]
  const float* scale3; const float* bias3;    // [32] output layer: 1 / (filter pre-scale) and bias (zero padded)
  float* logits;                              // [rows][FC]
  long rows;
  long tiles;
  int FC;
};

constexpr int kTailThreads = 192;
constexpr int kTailStages = 2;
constexpr int kTailWBytes = 32 * 128;        // one 32 x [32 hi | 32 lo] weight tile
// smem: A ring | H tile | W1 W2 W3 | barriers
constexpr int kTailOffH = kTailStages * kATileBytes;
constexpr int kTailOffW = kTailOffH + kATileBytes;
constexpr int kTailOffBar = kTailOffW + 3 * kTailWBytes;
constexpr int kTailSmem = kTailOffBar + 128 + 1024;

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
      }
      mbar_wait(&d_full[2], tp);
      tc_fence_after();
      {
        uint32_t t0[16], t1[16];
        tmem_ld16_nowait(tlane + 64u, t0);
        tmem_ld16_nowait(tlane + 80u, t1);
        tmem_wait_ld();
        const long vox = t * 128 + row;
        if (vox < p.rows) {
          float* dst = p.logits + vox * p.FC;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale3 + j));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias3 + j));
            *reinterpret_cast<float4*>(dst + j) = make_float4(fmaf(__uint_as_float(t0[j]), a.x, b.x), fmaf(__uint_as_float(t0[j + 1]), a.y, b.y),
                                                             fmaf(__uint_as_float(t0[j + 2]), a.z, b.z), fmaf(__uint_as_float(t0[j + 3]), a.w, b.w));
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (16 + j < p.FC) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale3 + 16 + j));
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias3 + 16 + j));
              *reinterpret_cast<float4*>(dst + 16 + j) = make_float4(fmaf(__uint_as_float(t1[j]), a.x, b.x), fmaf(__uint_as_float(t1[j + 1]), a.y, b.y),
                                                                    fmaf(__uint_as_float(t1[j + 2]), a.z, b.z), fmaf(__uint_as_float(t1[j + 3]), a.w, b.w));
            }
          }
        }
      }
      tc_fence_before();   // accumulator reads of this tile are ordered before the next tile's h_full arrivals
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

// x: split-fp16 rows [rows][32 hi | 32 lo]; w1/w2/w3: lt_conv_pair_pack_weights(taps = 1, Cin = 32, Cout = 32 / 32 / J) buffers
// (rows padded to 128; the first 32 are used); scale/shift: folded BN of the two hidden layers; scale3 / bias3 [32]: output affine
// (scale3 = 1 / filter pre-scale, bias zero padded);
// logits float32 [rows][FC], FC % 4 == 0, J <= FC <= 32.
extern "C" int lt_v2v_tail_fwd(const void* x, const void* w1, const void* w2, const void* w3, const float* scale1, const float* shift1,
                               const float* scale2, const float* shift2, const float* scale3, const float* bias3, float* logits, long rows,
                               int FC, void* stream) {
  LT_REQUIRE(x && w1 && w2 && w3 && scale1 && shift1 && scale2 && shift2 && scale3 && bias3 && logits, "v2v_tail: null pointer");
  LT_REQUIRE(rows > 0 && rows < (1L << 31) && FC % 4 == 0 && FC >= 4 && FC <= 32, "v2v_tail: bad sizes (rows=%ld FC=%d)", rows, FC);
  CUtensorMap tmX, tmW[3];
  {
    const uint64_t dims[2] = {64, (uint64_t)rows};
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
  TailParams p;
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2; p.scale3 = scale3; p.bias3 = bias3;
  p.logits = logits; p.rows = rows; p.tiles = (rows + 127) / 128; p.FC = FC;
  static DeviceOnce configured;
  if (configured.first()) {
    cudaError_t e = cudaFuncSetAttribute(v2v_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmem);
    if (e != cudaSuccess) return fail(LT_ERR_CUDA, "v2v_tail: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  long grid = 3L * sm_count();
  if (grid > p.tiles) grid = p.tiles;
  v2v_tail_kernel<<<(unsigned)grid, kTailThreads, kTailSmem, (cudaStream_t)stream>>>(tmX, tmW[0], tmW[1], tmW[2], p);
  LT_CHECK_LAUNCH("v2v_tail_kernel");
  return LT_OK;
}
