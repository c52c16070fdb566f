// Micro-benchmark #2 of the SM sub-partition pipes used by the edge kernel's producers / epilogue (B200, sm_100a):
// which instructions share the 16-lane XU pipe with MUFU, and what the packed-fp32 / conversion / shared-memory ops cost.
// cycles per warp-instruction per sub-partition, 1..4 warps per sub-partition, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned f2fp(float a, float b) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float h2f_lo(unsigned h) { float y; asm volatile("{.reg .b16 l, u; mov.b32 {l, u}, %1; cvt.f32.f16 %0, l;}" : "=f"(y) : "r"(h)); return y; }
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int U = 8, ITERS = 2048;
__shared__ float sbuf[4096];

template <int MODE>
__global__ void bench(float* out, long long* cyc, float seed) {
  float v[U];
  u64 p[U];
#pragma unroll
  for (int i = 0; i < U; ++i) { v[i] = seed + 0.001f * (threadIdx.x + i); p[i] = pk2(v[i], v[i] * 0.5f); }
  const u64 c1 = pk2(0.999f, 1.001f), c2 = pk2(1e-3f, -1e-3f);
  float* sp = sbuf + (threadIdx.x & 255) * 4;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) v[i] = __uint_as_float(f2fp(v[i], v[i]));                       // F2FP only
      if (MODE == 1) v[i] = h2f_lo(__float_as_uint(v[i]));                           // half -> float (HADD2.F32) only
      if (MODE == 2) { v[i] = ex2(v[i]); p[i] = fma2(p[i], c1, c2); }                // 1 MUFU + 1 FFMA2
      if (MODE == 3) { v[i] = ex2(v[i]); p[i] = pk2(__uint_as_float(f2fp(v[i], v[i])), 0.f); }   // 1 MUFU + 1 F2FP (dependent)
      if (MODE == 4) p[i] = fma2(p[i], c1, c2);                                      // FFMA2, 8 independent chains
      if (MODE == 5) v[i] = fma1(v[i], 0.999f, 1e-3f);                               // FFMA, 8 independent chains
      if (MODE == 6) p[i] = mul2(p[i], c1);                                          // FMUL2
      if (MODE == 7) { float a, b; upk2(p[i], a, b); *reinterpret_cast<float2*>(sp + 1024 * (i & 3)) = make_float2(a, b); p[i] = fma2(p[i], c1, c2); }  // STS.64 + FFMA2
      if (MODE == 8) { float4 x; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"((unsigned)__cvta_generic_to_shared(sp + 1024 * (i & 3)))); v[i] += x.x + x.w; }     // LDS.128 (+2 FADD)
      if (MODE == 9) { v[i] = ex2(v[i]); v[(i + 4) & 7] = __uint_as_float(f2fp(v[(i + 4) & 7], v[(i + 4) & 7])); }  // MUFU + independent F2FP
      if (MODE == 10) { v[i] = ex2(v[i]); v[(i + 4) & 7] = h2f_lo(__float_as_uint(v[(i + 4) & 7])); }                // MUFU + independent HADD2.F32
      if (MODE == 11) { v[i] = ex2(v[i]); p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c2, c1); p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c2, c1); }  // 1 MUFU + 4 FFMA2
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) { float a, b; upk2(p[i], a, b); s += v[i] + a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  printf("%-44s", name);
  for (int wps = 1; wps <= 4; ++wps) {
    const int threads = 128 * wps;
    bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
    bench<MODE><<<148, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("  wps=%d: %6.2f", wps, avg / ((double)wps * ITERS * U));        // cycles per loop-body instance per sub-partition
  }
  printf("   cyc per body\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("F2FP.F16.F32.PACK_AB");
  run<1>("HADD2.F32 (half -> float)");
  run<4>("FFMA2 (independent)");
  run<5>("FFMA (independent)");
  run<6>("FMUL2");
  run<2>("MUFU.EX2 + FFMA2");
  run<11>("MUFU.EX2 + 4 FFMA2");
  run<3>("MUFU.EX2 + dependent F2FP");
  run<9>("MUFU.EX2 + independent F2FP");
  run<10>("MUFU.EX2 + independent HADD2.F32");
  run<7>("STS.64 + FFMA2");
  run<8>("LDS.128 + 2 FADD");
  return 0;
}
