// TMA tensor-load feed probe (round-2 planning): how fast can one SM ingest the conv kernels' operand tiles from L2?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_probe tools/tma_probe.cu -lcuda && /tmp/tma_probe
//
// tools/cta2_probe.cu ("feed" mode) measured 64 B/clk per SM (125 GB/s, 18.5 TB/s aggregate) for 16 KB 1-D bulk copies.
// The conv kernels load TENSOR tiles instead: the A tile is a 5-D box of 128 rows x 128 B (128B swizzle), the B tile a 2-D
// box of 256 rows x 64 B (64B swizzle), and every CTA of a layer walks the SAME weight tiles in the same order.  This
// probe times exactly those shapes, 16 KB per request, DEPTH requests in flight per CTA, 1 or 2 CTAs per SM:
//   b64      2-D {32 fp16, rows}, box {32, 256}, SWIZZLE_64B       (today's B tile)
//   b128     2-D {64 fp16, rows}, box {64, 128}, SWIZZLE_128B      (B tile with two K chunks per row)
//   a5d      5-D {512, 24, 24, 1, 32}, box {64, 8, 8, 1, 2}, SWIZZLE_128B, tap-shifted coordinates (3x3 conv A tile)
//   mode suffix "s": every CTA reads the same tile sequence (weights); otherwise CTA-private tiles.
// If b64 runs at half the bytes/clk of b128, B tiles cost as many TMA row requests as bytes suggest and the 128-byte
// B layout is worth 1.3-1.5x on the feed-bound conv layers.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
  return reinterpret_cast<EncodeTiledFn>(ptr);
}
static void make_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, int swz) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) return false;
  }
}

// mode 0: 2-D loads at coordinates (0, row); mode 1: 5-D conv A tile
template <int DEPTH>
__global__ void __launch_bounds__(64, 2) tma_feed_kernel(const __grid_constant__ CUtensorMap map, int mode, int shared_seq, int tiles_in_window,
                                                         int rows_per_tile, int copies) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DEPTH * 16384);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int t = shared_seq ? 0 : (int)((blockIdx.x * 977u) % (unsigned)tiles_in_window);
  for (int i = 0; i < copies + DEPTH; ++i) {
    const int s = i % DEPTH;
    if (i >= DEPTH && !mbar_wait(&bars[s], ((i / DEPTH) - 1) & 1)) break;
    if (i < copies) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(16384u) : "memory");
      if (mode == 0) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem + s * 16384)), "l"(&map), "r"(smem_u32(&bars[s])), "r"(0), "r"(t * rows_per_tile) : "memory");
      } else {
        // tile t -> (channel block cb of 8, tap kw/kh of 3x3, 3x3 spatial tiles, image pair): the 3x3 conv's A walk
        const int cb = t & 7, tap = (t >> 3) % 9, sp = (t / 72) % 9, np = (t / 648) % 16;
        const int kw = tap % 3, kh = tap / 3, ow0 = (sp % 3) * 8, oh0 = (sp / 3) * 8;
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(smem_u32(smem + s * 16384)), "l"(&map), "r"(smem_u32(&bars[s])), "r"(cb * 64), "r"(ow0 - 1 + kw), "r"(oh0 - 1 + kh),
                       "r"(0), "r"(np * 2) : "memory");
      }
      if (++t >= tiles_in_window) t = 0;
    }
  }
}

template <int DEPTH>
static void run(const char* name, const CUtensorMap& map, int mode, int shared_seq, int tiles, int rows_per_tile, int grid) {
  const size_t smem = (size_t)DEPTH * 16384 + 256 + 1024;
  CK(cudaFuncSetAttribute(tma_feed_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int copies = 2048;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    tma_feed_kernel<DEPTH><<<grid, 64, smem>>>(map, mode, shared_seq, tiles, rows_per_tile, copies);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double bytes = (double)grid * copies * 16384.0;
  printf("%-6s%s depth=%d grid=%3d: %8.1f GB/s aggregate, %6.1f GB/s per CTA\n", name, shared_seq ? " (same tiles)" : " (private)   ", DEPTH, grid,
         bytes / best / 1e6, bytes / best / 1e6 / grid);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t bytes = (size_t)40 << 20;                  // 40 MB window: L2 resident
  void* buf;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0, bytes));
  CUtensorMap m64, m128, m5d;
  {
    const uint64_t rows = bytes / 64, dims[2] = {32, rows}, str[1] = {64};
    const uint32_t box[2] = {32, 256};
    make_map(&m64, buf, 2, dims, str, box, 64);
  }
  {
    const uint64_t rows = bytes / 128, dims[2] = {64, rows}, str[1] = {128};
    const uint32_t box[2] = {64, 128};
    make_map(&m128, buf, 2, dims, str, box, 128);
  }
  {
    // Cin = 256 split-fp16 activations: 512 fp16 = 1 KB per position, 24 x 24 x 32 images = 18.9 MB
    const uint64_t dims[5] = {512, 24, 24, 1, 32}, str[4] = {1024, 1024 * 24, 1024 * 24 * 24, 1024 * 24 * 24};
    const uint32_t box[5] = {64, 8, 8, 1, 2};
    make_map(&m5d, buf, 5, dims, str, box, 128);
  }
  const int tiles = (int)(bytes / 16384);
  printf("SMs: %d\n", sms);
  for (int grid : {sms, 2 * sms}) {
    run<4>("b64", m64, 0, 0, tiles, 256, grid);
    run<4>("b128", m128, 0, 0, tiles, 128, grid);
    run<4>("b64", m64, 0, 1, 144, 256, grid);          // 144 weight tiles = a 3x3 256->256 layer's B walk (2.4 MB)
    run<4>("b128", m128, 0, 1, 144, 128, grid);
    run<4>("a5d", m5d, 1, 0, 72 * 9 * 16, 0, grid);
    run<8>("b64", m64, 0, 0, tiles, 256, grid);
    run<8>("a5d", m5d, 1, 0, 72 * 9 * 16, 0, grid);
  }
  printf("done\n");
  return 0;
}
