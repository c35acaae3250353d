// Microbenchmark: does a packed FFMA2 (two passes through the FP32 pipe) leave an issue slot for an independent
// instruction of another pipe?  Streams of 16 independent FFMA2 chains, alone and with k extra instructions per FFMA2
// (integer LOP3 / IADD3, FSEL, LDS, SHFL, scalar FFMA), at 4 and 2 resident warps per scheduler.
// time(mix) == time(FFMA2 alone) means the extra instruction issues in the shadow of the packed one.
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 16
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a, float b, int sel) {
  __shared__ float sm[1024];
  float2 y[ILP]; unsigned u[ILP]; float x[ILP];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ILP; ++i) { y[i] = make_float2(threadIdx.x * 1e-3f + i, i + 1.f); u[i] = threadIdx.x * 7 + i; x[i] = i * 0.5f; }
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE < 9) y[i] = __ffma2_rn(y[i], a2, b2);
        // asm volatile: the compiler must neither fold nor drop the extra instructions
        if (MODE == 1) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel)); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel)); }
        if (MODE == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel));
        if (MODE == 3) asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p; }" : "+f"(x[i]) : "f"(a), "r"(sel));
        if (MODE == 4) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(x[i]) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + i * 33) & 1023))));
        if (MODE == 5) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 7));
        if (MODE == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        if (MODE == 7 && (i & 1)) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 7));
        if (MODE == 8) asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "r"(sel + i));
        if (MODE == 9) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel)); }
        if (MODE == 10) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); }
        if (MODE == 11) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel)); }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += y[i].x + y[i].y + x[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int iters, int warps_per_sched) {
  int dev = 0, sms = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one CTA of 256 threads = 2 warps per scheduler; blocks per SM set the residency
  const int bps = warps_per_sched / 2;
  const int grid = sms * bps;
  float* out; cudaMalloc(&out, (size_t)grid * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k_mix<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f, 5);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_mix<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // cycles per FFMA2 slot and scheduler (1.965 GHz assumed; relative numbers matter)
  const double slots = (double)iters * 4 * ILP * warps_per_sched;
  printf("%-28s warps/sched %d  %8.3f ms  %.3f cycles per (FFMA2 + extra) per scheduler\n", name, warps_per_sched, best,
         best * 1e-3 * 1.965e9 / slots);
  cudaFree(out);
}

int main() {
  if (cudaFree(0) != cudaSuccess) { printf("no gpu\n"); return 1; }
  for (int w = 2; w <= 4; w += 2) {
    run<0>("FFMA2 alone", 4000, w);
    run<2>("FFMA2 + LOP3", 4000, w);
    run<1>("FFMA2 + LOP3 + IADD3", 4000, w);
    run<3>("FFMA2 + select", 4000, w);
    run<4>("FFMA2 + LDS", 4000, w);
    run<5>("FFMA2 + SHFL", 4000, w);
    run<7>("FFMA2 + SHFL/2", 4000, w);
    run<6>("FFMA2 + FFMA", 4000, w);
    run<8>("FFMA2 + MOV", 4000, w);
    run<9>("FFMA + LOP3 (no packed)", 4000, w);
    run<10>("FFMA alone", 4000, w);
    run<11>("LOP3 alone", 4000, w);
  }
  return 0;
}
