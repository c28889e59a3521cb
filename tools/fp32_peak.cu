// fp32_peak.cu -- measures what the FP32 pipe of this GPU sustains for (a) scalar FFMA with three register operands,
// (b) packed fma.rn.f32x2 (FFMA2, sm_100+), so that bench.py's FP32 roofline denominator can be judged against a
// measured figure instead of the nominal SMs x 128 x 2 x clock.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 8;

__global__ void __launch_bounds__(256) k_scalar(float* out, float a, float b) {
  float x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x * 1e-3f + i;
  float aa = a + threadIdx.x * 1e-9f, bb = b + threadIdx.x * 1e-9f;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x[i] = fmaf(x[i], aa, bb);   // three register operands
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_packed(float* out, float a, float b) {
  unsigned long long x[CHAINS], aa, bb;
  float a0 = a + threadIdx.x * 1e-9f;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a0), "f"(a0));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    float v = threadIdx.x * 1e-3f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(v), "f"(v + 0.5f));
  }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  float* out;
  cudaMalloc(&out, size_t(blocks) * threads * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; which++) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
      cudaEventRecord(e0);
      if (which == 0) k_scalar<<<blocks, threads>>>(out, 0.999f, 0.001f);
      else k_packed<<<blocks, threads>>>(out, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double flops = double(blocks) * threads * ITERS * CHAINS * 2.0 * (which ? 2.0 : 1.0);
    printf("{\"kernel\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"sms\": %d, \"clock_mhz\": %d}\n",
           which ? "fma.rn.f32x2 (FFMA2)" : "scalar FFMA, 3 register operands", best, flops / best / 1e9, p.multiProcessorCount,
           p.clockRate / 1000);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
