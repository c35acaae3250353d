// Microbenchmark: FP32 CUDA-core issue peak on B200 (sm_100a).
// Measures scalar FFMA, packed FFMA2 (fma.rn.f32x2), SHFL.BFLY and DFMA throughput.
// The FFMA2 figure is the roofline denominator for the gate kernels (MEASURED_PEAKS.json has none).
#include <cstdio>
#include <string>
#include <cuda_runtime.h>

#define ILP 16
template <int MODE>
__global__ void __launch_bounds__(256) k_peak(float* out, int iters, float a, float b) {
  float x[ILP]; float2 y[ILP]; double z[ILP/2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 1.f); }
#pragma unroll
  for (int i = 0; i < ILP/2; ++i) z[i] = x[i];
  float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) y[i] = __ffma2_rn(y[i], a2, b2);
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 7));
      } else if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < ILP/2; ++i) z[i] = fma(z[i], (double)a, (double)b);
      } else if (MODE == 4) {  // 1 SHFL per 8 FFMA2 mix
#pragma unroll
        for (int i = 0; i < ILP; ++i) y[i] = __ffma2_rn(y[i], a2, b2);
        x[r] = __shfl_xor_sync(0xffffffffu, x[r], 1);
        x[r+8] = __shfl_xor_sync(0xffffffffu, x[r+8], 2);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i].x + y[i].y;
#pragma unroll
  for (int i = 0; i < ILP/2; ++i) s += (float)z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, double flop_per_inner, int iters, int blocks_per_sm) {
  int dev = 0, sms = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms * blocks_per_sm;
  float* out; cudaMalloc(&out, (size_t)grid * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k_peak<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_peak<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double inner = (double)grid * 256 * iters * 8.0;
  double rate = inner * flop_per_inner / (best * 1e-3);
  printf("{\"bench\":\"%s\",\"ms\":%.4f,\"rate_per_s\":%.6e,\"blocks_per_sm\":%d,\"sms\":%d}\n", name, best, rate, blocks_per_sm, sms);
  cudaFree(out);
  return rate;
}

int main(int argc, char** argv) {
  cudaError_t e = cudaFree(0);
  if (e != cudaSuccess) { printf("no gpu: %s\n", cudaGetErrorString(e)); return 1; }
  const bool quick = argc > 1 && std::string(argv[1]) == "--quick";
  if (quick) {  // roofline denominator only (bench.py)
    for (int bps = 4; bps <= 8; bps *= 2) {
      run<0>("ffma_flops", ILP * 2.0, 4000, bps);
      run<1>("ffma2_flops", ILP * 4.0, 4000, bps);
      run<3>("dfma_flops", ILP / 2 * 2.0, 1000, bps);      // FP64 FMA peak: roofline denominator of the complex128 kernels
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
  }
  for (int bps = 2; bps <= 8; bps *= 2) {
    run<0>("ffma_flops", ILP * 2.0, 4000, bps);
    run<1>("ffma2_flops", ILP * 4.0, 4000, bps);
    run<2>("shfl_lane_ops", ILP * 1.0, 2000, bps);
    run<3>("dfma_flops", ILP / 2 * 2.0, 1000, bps);
    run<4>("ffma2_plus_shfl_flops", ILP * 4.0, 4000, bps);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return 0;
}
