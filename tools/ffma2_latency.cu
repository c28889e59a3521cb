// ffma2_latency.cu -- issue / latency behaviour of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a at the
// occupancy of the tick kernels (2 warps per scheduler): cycles per instruction for dependent chains of ILP 1..8.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_latency tools/ffma2_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int ILP, bool PACKED>
__global__ void k(float* out, long long* cyc, float a, float b) {
  float2 x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f - i);
  const float2 aa = make_float2(a, a + 1e-9f * threadIdx.x), bb = make_float2(b, b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (PACKED) x[i] = __ffma2_rn(x[i], aa, bb);
      else x[i].x = fmaf(x[i].x, aa.y, bb.x);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, bool PACKED> void run(int threads, float* out, long long* dcyc) {
  k<ILP, PACKED><<<148, threads>>>(out, dcyc, 0.999f, 0.001f);
  k<ILP, PACKED><<<148, threads>>>(out, dcyc, 0.999f, 0.001f);
  long long c;
  cudaMemcpy(&c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
  const double per_instr = double(c) / (double(ITERS) * ILP);
  printf("{\"op\": \"%s\", \"ilp\": %d, \"warps_per_scheduler\": %d, \"cycles_per_instr_per_warp\": %.3f, \"instr_per_cycle_per_scheduler\": %.3f}\n",
         PACKED ? "FFMA2" : "FFMA", ILP, threads / 128, per_instr, (threads / 128) / per_instr);
}

int main() {
  float* out; long long* dcyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&dcyc, sizeof(long long));
  for (int threads : {128, 256, 512}) {
    run<1, false>(threads, out, dcyc); run<1, true>(threads, out, dcyc);
    run<2, false>(threads, out, dcyc); run<2, true>(threads, out, dcyc);
    run<3, false>(threads, out, dcyc); run<3, true>(threads, out, dcyc);
    run<4, false>(threads, out, dcyc); run<4, true>(threads, out, dcyc);
    run<8, false>(threads, out, dcyc); run<8, true>(threads, out, dcyc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
