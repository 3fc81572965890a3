// sb_ubench.cu — FP32 FMA-pipe peak microbenchmark: the roofline denominator of the fused train step
// (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks). Three issue patterns, all with the
// accumulator as the addend and 32 independent chains per thread, 8 blocks x 256 threads per SM:
//   0: scalar FFMA            acc = a * m[j] + acc
//   1: packed FFMA2           acc2 = a2 * m2[j] + acc2        (fma.rn.f32x2)
//   2: packed FFMA2, constant acc2 = c2[j] * m2[j] + acc2     (the pattern of the prediction step)
// Design probes (not roofline denominators; reported in the same unit, 2 x lane-operations per second):
//   3: warp shuffles only     m[j].x = shfl_xor(m[j].x, 1)    (32 lane-operations per SHFL)
//   4: pattern 2 with one SHFL after every 4 FFMA2 (only the FFMA2 are counted): do shuffles cost FMA issue slots?
#include "sb_common.cuh"

namespace sb {

namespace {

constexpr int kThreads = 256;
constexpr int kChains = 16;  // float2 accumulators per thread (32 scalar chains)

__constant__ float2 c_peak[kChains];

template <int VARIANT>
__global__ void __launch_bounds__(kThreads) peak_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        int iters) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  float2 acc[kChains], m[kChains];
#pragma unroll
  for (int j = 0; j < kChains; ++j) {
    acc[j] = make_float2(0.f, 0.f);
    m[j] = make_float2(in[(tid + j) & 1023], in[(tid + 2 * j + 1) & 1023]);
  }
  const float2 a = make_float2(in[tid & 1023], in[(tid + 7) & 1023]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int j = 0; j < kChains; ++j) {
        if (VARIANT == 0) {
          acc[j].x = fmaf(a.x, m[j].x, acc[j].x);
          acc[j].y = fmaf(a.y, m[j].y, acc[j].y);
        } else if (VARIANT == 1) {
          acc[j] = __ffma2_rn(a, m[j], acc[j]);
        } else if (VARIANT == 3) {
          m[j].x = __shfl_xor_sync(0xffffffffu, m[j].x, 1);
          m[j].y = __shfl_xor_sync(0xffffffffu, m[j].y, 2);
        } else if (VARIANT == 4) {
          acc[j] = __ffma2_rn(c_peak[j], m[j], acc[j]);
          if (j % 4 == 3) m[(j + 8) % kChains].x = __shfl_xor_sync(0xffffffffu, m[(j + 8) % kChains].x, 1);
        } else {
          acc[j] = __ffma2_rn(c_peak[j], m[j], acc[j]);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kChains; ++j) s += acc[j].x + acc[j].y + m[j].x + m[j].y;
  out[tid] = s;
}

}  // namespace

int fp32_peak(int variant, int iters, double* tflops_host, cudaStream_t s) {
  int dev = 0, sms = 0;
  SB_CUDA_TRY(cudaGetDevice(&dev));
  SB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8;
  float *in = nullptr, *out = nullptr;
  SB_CUDA_TRY(cudaMalloc(&in, 1024 * sizeof(float)));
  SB_CUDA_TRY(cudaMalloc(&out, (size_t)blocks * kThreads * sizeof(float)));
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 1e-3f * (float)((i * 37) % 101) - 0.05f;
  float2 hc[kChains];
  for (int j = 0; j < kChains; ++j) hc[j] = make_float2(1e-3f * j, -1e-3f * j);
  SB_CUDA_TRY(cudaMemcpyAsync(in, h, sizeof(h), cudaMemcpyHostToDevice, s));
  SB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_peak, hc, sizeof(hc), 0, cudaMemcpyHostToDevice, s));
  cudaEvent_t e0, e1;
  SB_CUDA_TRY(cudaEventCreate(&e0));
  SB_CUDA_TRY(cudaEventCreate(&e1));
  auto launch = [&](int it) {
    if (variant == 0) peak_kernel<0><<<blocks, kThreads, 0, s>>>(in, out, it);
    else if (variant == 1) peak_kernel<1><<<blocks, kThreads, 0, s>>>(in, out, it);
    else if (variant == 3) peak_kernel<3><<<blocks, kThreads, 0, s>>>(in, out, it);
    else if (variant == 4) peak_kernel<4><<<blocks, kThreads, 0, s>>>(in, out, it);
    else peak_kernel<2><<<blocks, kThreads, 0, s>>>(in, out, it);
  };
  launch(iters / 8 + 1);  // warm-up
  float best_ms = 1e30f;
  for (int r = 0; r < 3; ++r) {
    SB_CUDA_TRY(cudaEventRecord(e0, s));
    launch(iters);
    SB_CUDA_TRY(cudaEventRecord(e1, s));
    SB_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    SB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) best_ms = ms;
  }
  cudaError_t le = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(in);
  cudaFree(out);
  if (le != cudaSuccess) return cuda_fail(le, "peak_kernel");
  const double fmas = (double)blocks * kThreads * (double)iters * 4.0 * kChains * 2.0;  // lane-FMAs
  *tflops_host = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
  return SB_OK;
}

}  // namespace sb
