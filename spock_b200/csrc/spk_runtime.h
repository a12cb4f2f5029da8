// spk_runtime.h -- thin layer over the CUDA runtime used by the solver.
//
// The product library is always compiled by nvcc for sm_100a (SPK_HOST_EMU undefined).
// SPK_HOST_EMU exists only for tests/hostemu: it compiles the very same kernels as serial host code
// (one "thread" per block, one lane per group) so that index arithmetic and solver orchestration can be
// unit-tested in a container without a GPU.  That build is never linked into libspock_b200.so and is not
// reachable from the C ABI; the product has no CPU fallback.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#ifndef SPK_HOST_EMU
// ------------------------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>

#define SPK_HD __host__ __device__ __forceinline__
#define SPK_D __device__ __forceinline__
#define SPK_CONSTEXPR_HD __host__ __device__ constexpr
#define SPK_KERNEL __global__
#define SPK_SMEM_DECL(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; type *name = reinterpret_cast<type *>(name##_raw)
#define SPK_SYNCTHREADS() __syncthreads()
#define SPK_SYNCWARP() __syncwarp()
#define SPK_EMU 0

typedef cudaStream_t spk_stream_t;
typedef cudaEvent_t spk_event_t;

#define SPK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// the same launch chained to the previous one on the stream by programmatic dependent launch (the kernel must execute
// griddepcontrol.wait before it touches anything the previous kernel wrote); `pdl` false = plain stream order
template <class... KArgs, class... Args>
static inline cudaError_t spk_launch_pdl(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define SPK_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) (void)spk_launch_pdl((pdl), kernel, (grid), (block), (smem), (stream), __VA_ARGS__)

#else
// ------------------------------------------------------------------------------------------------ host emulation (tests only)
#include <cmath>
#include <vector>

#define SPK_HD inline
#define SPK_D inline
#define SPK_CONSTEXPR_HD constexpr
#define SPK_KERNEL static
#define SPK_SYNCTHREADS() ((void)0)
#define SPK_SYNCWARP() ((void)0)
#define SPK_EMU 1

struct spk_dim3 {
  unsigned x, y, z;
  spk_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef spk_dim3 dim3;
extern thread_local spk_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local std::vector<unsigned char> spk_emu_smem;
#define SPK_SMEM_DECL(type, name) type *name = reinterpret_cast<type *>(spk_emu_smem.data())

typedef int spk_stream_t;
typedef double spk_event_t;  // wall-clock ms
struct double2 { double x, y; };

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

// every emulated launch runs block-by-block with ONE thread per block: kernels are written so that any
// power-of-two block size (including 1) is correct
#define SPK_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
  do {                                                                         \
    spk_dim3 g_ = (grid), b_ = (block);                                        \
    if (spk_emu_smem.size() < (size_t)(smem) + 64) spk_emu_smem.resize((size_t)(smem) + 64); \
    gridDim = g_;                                                              \
    blockDim = b_;                                                             \
    for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                                \
      blockIdx = spk_dim3(bx_, 0, 0);                                          \
      for (unsigned tx_ = 0; tx_ < b_.x; ++tx_) {                              \
        threadIdx = spk_dim3(tx_, 0, 0);                                       \
        kernel(__VA_ARGS__);                                                   \
      }                                                                        \
    }                                                                          \
  } while (0)

#define SPK_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...) SPK_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)

#endif

// ------------------------------------------------------------------------------------------------ shared
#define SPK_MAX_CONES 8
#define SPK_MAX_LS 64   // ny + 2d upper bound for the per-thread S2 kernel
#define SPK_MAX_LEVELS 64
