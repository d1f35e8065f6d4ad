// Small kernels of the algebraic-triangulation path and the confidence heads:
//   - global-average-pool + 3-layer MLP + sigmoid tail of GlobalAveragePoolingHead (pose_resnet.py:163-174)
//   - normalisation of per-view confidences (triangulation.py:173-174, :268-269)
//   - confidence-weighted DLT triangulation (multiview.py:141-183), one thread per (sample, joint)
#include "common.cuh"

namespace lt {

// One CTA per image: mean over P positions of C0 channels, then Linear(C0,H1)+ReLU, Linear(H1,H2)+ReLU,
// Linear(H2,NO)+Sigmoid.  Weights are row-major [out][in] float32 (nn.Linear layout).  Dynamic smem: C0+H1+H2 floats.
__global__ void __launch_bounds__(256) gap_mlp3_kernel(const void* __restrict__ in, int format, int P, int C0, int H1, int H2, int NO,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       const float* __restrict__ w3, const float* __restrict__ b3,
                                                       float* __restrict__ out) {
  extern __shared__ float sm[];
  float* x0 = sm;
  float* x1 = x0 + C0;
  float* x2 = x1 + H1;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C0; c += blockDim.x) {
    float acc = 0.0f;
    for (int q = 0; q < P; ++q) {
      const long pix = (long)n * P + q;
      if (format == LT_FMT_F32) acc += reinterpret_cast<const float*>(in)[pix * C0 + c];
      else {
        const sh_t* row = reinterpret_cast<const sh_t*>(in) + pix * 2 * C0;
        acc += join_s32(row[s32_off(c)], row[s32_off(c) + 32]);
      }
    }
    x0[c] = acc / (float)P;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < H1; o += blockDim.x) {
    float acc = b1[o];
    for (int i = 0; i < C0; ++i) acc = fmaf(w1[(long)o * C0 + i], x0[i], acc);
    x1[o] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < H2; o += blockDim.x) {
    float acc = b2[o];
    for (int i = 0; i < H1; ++i) acc = fmaf(w2[(long)o * H1 + i], x1[i], acc);
    x2[o] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < NO; o += blockDim.x) {
    float acc = b3[o];
    for (int i = 0; i < H2; ++i) acc = fmaf(w3[(long)o * H2 + i], x2[i], acc);
    out[(long)n * NO + o] = 1.0f / (1.0f + expf(-acc));
  }
}

// conf[B][V][C] /= sum over views; += eps
__global__ void view_normalize_kernel(float* __restrict__ conf, int B, int V, int C, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float s = 0.0f;
  for (int v = 0; v < V; ++v) s += conf[((long)b * V + v) * C + c];
  for (int v = 0; v < V; ++v) conf[((long)b * V + v) * C + c] = conf[((long)b * V + v) * C + c] / s + eps;
}

// Weighted DLT (Hartley & Zisserman 12.2): rows c*(x*P[2] - P[0]), c*(y*P[2] - P[1]); the solution is the right singular
// vector of the smallest singular value of A (2V x 4) = eigenvector of A^T A for its smallest eigenvalue.  A^T A and a cyclic
// Jacobi eigen-solve run in float64 (forming A^T A squares the condition number, fp32 would not do); the reference's
// `-vh[:, 3]` sign cancels in the dehomogenisation.
__global__ void __launch_bounds__(128) triangulate_dlt_kernel(const float* __restrict__ proj, const float* __restrict__ kp2d,
                                                              const float* __restrict__ conf, float* __restrict__ out, int B, int V, int J) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * J) return;
  const int b = idx / J, j = idx % J;
  double M[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) M[r][c] = 0.0;
  for (int v = 0; v < V; ++v) {
    const float* P = proj + ((long)b * V + v) * 12;
    const float x = kp2d[(((long)b * V + v) * J + j) * 2], y = kp2d[(((long)b * V + v) * J + j) * 2 + 1];
    const float cf = conf ? conf[((long)b * V + v) * J + j] : 1.0f;
    double r0[4], r1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // the reference forms these rows in float32 (multiview.py:159-161); keep that rounding, accumulate in float64
      r0[c] = (double)((P[8 + c] * x - P[c]) * cf);
      r1[c] = (double)((P[8 + c] * y - P[4 + c]) * cf);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) M[r][c] += r0[r] * r0[c] + r1[r] * r1[c];
  }
  double E[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) off += M[p][q] * M[p][q];
    if (off < 1e-300) break;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        if (M[p][q] == 0.0) continue;
        const double theta = (M[q][q] - M[p][p]) / (2.0 * M[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const double a = M[k][p], bb = M[k][q]; M[k][p] = c * a - s * bb; M[k][q] = s * a + c * bb; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const double a = M[p][k], bb = M[q][k]; M[p][k] = c * a - s * bb; M[q][k] = s * a + c * bb; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const double a = E[k][p], bb = E[k][q]; E[k][p] = c * a - s * bb; E[k][q] = s * a + c * bb; }
      }
    }
  }
  int m = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if (M[k][k] < M[m][m]) m = k;
  const double w = E[3][m];
  out[(long)idx * 3 + 0] = (float)(E[0][m] / w);
  out[(long)idx * 3 + 1] = (float)(E[1][m] / w);
  out[(long)idx * 3 + 2] = (float)(E[2][m] / w);
}

}  // namespace lt

using namespace lt;

extern "C" int lt_gap_mlp3_fwd(const void* in, int format, int N, int P, int C0, int H1, int H2, int NO, const float* w1,
                               const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, float* out,
                               void* stream) {
  LT_REQUIRE(in && w1 && b1 && w2 && b2 && w3 && b3 && out, "gap_mlp3: null pointer");
  LT_REQUIRE(N > 0 && P > 0 && C0 > 0 && H1 > 0 && H2 > 0 && NO > 0, "gap_mlp3: bad sizes");
  LT_REQUIRE(format == LT_FMT_F32 || C0 % 32 == 0, "gap_mlp3: split-fp16 input needs C0 %% 32 == 0");
  const size_t smem = (size_t)(C0 + H1 + H2) * sizeof(float);
  LT_REQUIRE(smem <= 48 * 1024, "gap_mlp3: hidden sizes too large");
  gap_mlp3_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(in, format, P, C0, H1, H2, NO, w1, b1, w2, b2, w3, b3, out);
  LT_CHECK_LAUNCH("gap_mlp3_kernel");
  return LT_OK;
}

extern "C" int lt_view_normalize_fwd(float* conf, int B, int V, int C, float eps, void* stream) {
  LT_REQUIRE(conf && B > 0 && V > 0 && C > 0, "view_normalize: bad arguments");
  view_normalize_kernel<<<ceil_div((long)B * C, 128), 128, 0, (cudaStream_t)stream>>>(conf, B, V, C, eps);
  LT_CHECK_LAUNCH("view_normalize_kernel");
  return LT_OK;
}

extern "C" int lt_triangulate_dlt_fwd(const float* proj, const float* keypoints_2d, const float* confidences, float* out, int B,
                                      int V, int J, void* stream) {
  LT_REQUIRE(proj && keypoints_2d && out && B > 0 && V > 0 && J > 0, "triangulate_dlt: bad arguments");
  triangulate_dlt_kernel<<<ceil_div((long)B * J, 128), 128, 0, (cudaStream_t)stream>>>(proj, keypoints_2d, confidences, out, B, V, J);
  LT_CHECK_LAUNCH("triangulate_dlt_kernel");
  return LT_OK;
}
