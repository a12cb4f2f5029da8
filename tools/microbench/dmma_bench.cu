// dmma_bench.cu -- fp64 throughput of mma.sync.m8n8k4.f64 (DMMA) against plain DFMA on this GPU.
// Decides whether the per-stage (factor-table) S1 kernels should run their batched mat-mats on the tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ACC>
__global__ void k_dmma(double *out, int iters, double a0, double b0) {
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// independent operands per accumulator (no shared source registers between back-to-back DMMAs)
template <int ACC>
__global__ void k_dmma_ind(double *out, int iters, double a0, double b0) {
  double c[ACC][2], a[ACC], b[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) {
    c[i][0] = c[i][1] = 0.0;
    a[i] = a0 + threadIdx.x + i;
    b[i] = b0 - threadIdx.x - i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
  double c[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps >= 8 ? 256 : warps * 32, blocks = sms * (warps * 32 / threads);
    float ms;
    k_dmma<8><<<blocks, threads>>>(out, 100, 1.0, 1.0);
    cudaEventRecord(e0);
    k_dmma<8><<<blocks, threads>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl_mma = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
    printf("DMMA m8n8k4  warps/SM %2d: %.2f TFLOP/s\n", warps, fl_mma / ms / 1e9);
    k_dfma<8><<<blocks, threads>>>(out, 100, 1.0, 1.0);
    cudaEventRecord(e0);
    k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl_fma = 2.0 * 8.0 * iters * (double)blocks * threads;
    printf("DFMA         warps/SM %2d: %.2f TFLOP/s\n", warps, fl_fma / ms / 1e9);
  }
  for (int warps = 8; warps <= 32; warps *= 2) {
    const int threads = 256, blocks = sms * (warps * 32 / threads);
    float ms;
    k_dmma_ind<16><<<blocks, threads>>>(out, 100, 1.0, 1.0);
    cudaEventRecord(e0);
    k_dmma_ind<16><<<blocks, threads>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA m8n8k4, 16 independent operand sets, warps/SM %2d: %.2f TFLOP/s\n", warps,
           2.0 * 8 * 8 * 4 * 16.0 * iters * (double)blocks * (threads / 32) / ms / 1e9);
    k_dmma_ind<4><<<blocks, threads>>>(out, 100, 1.0, 1.0);
    cudaEventRecord(e0);
    k_dmma_ind<4><<<blocks, threads>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA m8n8k4,  4 independent operand sets, warps/SM %2d: %.2f TFLOP/s\n", warps,
           2.0 * 8 * 8 * 4 * 4.0 * iters * (double)blocks * (threads / 32) / ms / 1e9);
  }
  printf("%s, %d SMs, %d MHz\n", p.name, sms, p.clockRate / 1000);
  return 0;
}
