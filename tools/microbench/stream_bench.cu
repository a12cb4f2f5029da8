// Microbenchmark: how fast can a B200 stream family-sized (20-60 KB) contiguous chunks from HBM into shared memory?
//   mode 0: one chunk per CTA, cp.async.bulk, CTA exits after consuming it (occupancy limited by smem)
//   mode 1: persistent CTAs with a STAGES-deep ring of cp.async.bulk buffers
//   mode 2: plain coalesced 16-byte loads, grid-stride, no shared memory
// usage: stream_bench <chunk_bytes> <total_GB> <mode> <stages> <threads> [ctas_per_sm]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          s32(b)),
      "r"(ph)
      : "memory");
}

__global__ void k_single(const double *src, long nchunks, int chunk_doubles, double *out) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t *bar = (uint64_t *)raw;
  double *buf = (double *)(raw + 16);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __syncthreads();
  double acc = 0;
  uint32_t ph = 0;
  for (long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    if (threadIdx.x == 0) { mbar_expect(bar, chunk_doubles * 8); bulk(buf, src + c * chunk_doubles, chunk_doubles * 8, bar); }
    mbar_wait(bar, ph); ph ^= 1;
    for (int i = threadIdx.x; i < chunk_doubles; i += blockDim.x * 8) acc += buf[i];
    __syncthreads();
  }
  if (acc == 123.456) out[0] = acc;
}

template <int STAGES>
__global__ void k_ring(const double *src, long nchunks, int chunk_doubles, double *out) {
  extern __shared__ __align__(16) unsigned char raw[];
  uint64_t *bar = (uint64_t *)raw;  // STAGES barriers
  double *buf = (double *)(raw + 64);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const long first = blockIdx.x, step = gridDim.x;
  long issued = 0;
  if (threadIdx.x == 0)
    for (int s = 0; s < STAGES; ++s) {
      const long c = first + issued * step;
      if (c < nchunks) { mbar_expect(bar + s, chunk_doubles * 8); bulk(buf + (long)s * chunk_doubles, src + c * chunk_doubles, chunk_doubles * 8, bar + s); }
      ++issued;
    }
  double acc = 0;
  long k = 0;
  for (long c = first; c < nchunks; c += step, ++k) {
    const int s = k % STAGES;
    const uint32_t ph = (k / STAGES) & 1;
    mbar_wait(bar + s, ph);
    const double *b = buf + (long)s * chunk_doubles;
    for (int i = threadIdx.x; i < chunk_doubles; i += blockDim.x * 8) acc += b[i];
    __syncthreads();
    if (threadIdx.x == 0) {
      const long cn = first + (k + STAGES) * step;
      if (cn < nchunks) { mbar_expect(bar + s, chunk_doubles * 8); bulk(buf + (long)s * chunk_doubles, src + cn * chunk_doubles, chunk_doubles * 8, bar + s); }
    }
  }
  if (acc == 123.456) out[0] = acc;
}

__global__ void k_ldg(const double2 *src, long n2, double *out) {
  double acc = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n2; i += 8 * stride) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
  }
  for (; i < n2; i += stride) { double2 v = __ldg(src + i); acc += v.x + v.y; }
  if (acc == 123.456) out[0] = acc;
}

int main(int argc, char **argv) {
  const int chunk = atoi(argv[1]);
  const double gb = atof(argv[2]);
  const int mode = atoi(argv[3]), stages = atoi(argv[4]), threads = atoi(argv[5]);
  const int cps = argc > 6 ? atoi(argv[6]) : 0;
  const int cd = chunk / 8;
  const long nchunks = (long)(gb * 1e9 / chunk);
  double *src, *out;
  cudaMalloc(&src, nchunks * (long)chunk);
  cudaMalloc(&out, 8);
  cudaMemset(src, 0, nchunks * (long)chunk);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) {
      const size_t smem = 16 + chunk;
      cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const long grid = cps > 0 ? (long)sms * cps : nchunks;
      k_single<<<(unsigned)grid, threads, smem>>>(src, nchunks, cd, out);
    } else if (mode == 1) {
      const size_t smem = 64 + (size_t)stages * chunk;
      const long grid = (long)sms * (cps > 0 ? cps : 1);
#define RING(S) { cudaFuncSetAttribute(k_ring<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k_ring<S><<<(unsigned)grid, threads, smem>>>(src, nchunks, cd, out); }
      if (stages == 2) RING(2) else if (stages == 3) RING(3) else if (stages == 4) RING(4) else if (stages == 6) RING(6) else RING(8)
    } else {
      k_ldg<<<sms * (cps > 0 ? cps : 8), threads>>>((const double2 *)src, nchunks * (long)chunk / 16, out);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  printf("chunk %d B  mode %d stages %d threads %d cps %d : %.3f ms  %.1f GB/s  (%s)\n", chunk, mode, stages, threads, cps,
         best, nchunks * (double)chunk / best / 1e6, cudaGetErrorString(err));
  return 0;
}
