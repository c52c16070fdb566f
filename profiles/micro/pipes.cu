// Micro-benchmark of the SM pipes the edge kernel's CUDA-core roles live on (B200, sm_100a):
// cycles per warp-instruction per SM sub-partition for MUFU.EX2 / MUFU.RCP / F2FP pack / packed FFMA2 and the SiLU sequence,
// with 1..4 warps per sub-partition and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int U = 8, ITERS = 2048;

template <int MODE>
__global__ void bench(float* out, long long* cyc, float seed) {
  float v[U];
#pragma unroll
  for (int i = 0; i < U; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      if (MODE == 1) v[i] = rcp(v[i]);
      if (MODE == 2) v[i] = rcp(1.0f + ex2(v[i]));                                   // 2 MUFU + 1 FADD
      if (MODE == 3) { __half2 h = __floats2half2_rn(v[i], v[i] + 1.0f); v[i] = __low2float(h) + __high2float(h); }   // 1 F2FP + unpack + add
      if (MODE == 4) v[i] = v[i] * rcp(1.0f + ex2(v[i] * -1.4426950408889634f));     // SiLU: 2 FMUL, FADD, 2 MUFU
    }
    if (MODE == 5) {                                                                 // packed FFMA2 only (4 per iteration)
      u64 a = pk2(v[0], v[1]), b = pk2(v[2], v[3]), c = pk2(v[4], v[5]), d = pk2(v[6], v[7]);
      a = fma2(a, b, c); b = fma2(b, c, d); c = fma2(c, d, a); d = fma2(d, a, b);
      upk2(a, v[0], v[1]); upk2(b, v[2], v[3]); upk2(c, v[4], v[5]); upk2(d, v[6], v[7]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_instr) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  printf("%-28s", name);
  for (int wps = 1; wps <= 4; ++wps) {          // warps per sub-partition
    const int threads = 128 * wps;
    bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
    bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // warp-instructions of interest issued per sub-partition: wps warps x ITERS x per_iter_instr
    printf("  wps=%d: %6.2f cyc/inst", wps, avg / ((double)wps * ITERS * per_iter_instr));
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", U);
  run<1>("MUFU.RCP", U);
  run<2>("rcp(1+ex2) [per MUFU]", 2 * U);
  run<3>("F2FP.F16 pack [per F2FP]", U);
  run<4>("SiLU [per SiLU]", U);
  run<5>("FFMA2 [per FFMA2]", 4);
  return 0;
}
