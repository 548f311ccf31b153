/* Fake CUDA runtime for the CPU emulation of the reference (test scaffolding, NOT product
 * code). cudaMalloc zero-fills and pads by 4 MB, which turns the reference's
 * out-of-bounds depth read (tsdf.cu:2114) into "reads 0.0f". */
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include "emu_shim.h"

emu_idx3 emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;

extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n + (4u << 20)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceReset(void) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emu"; }
}
void __cudaSafeCall(cudaError err, const char* file, const int line) {
  if (err != cudaSuccess) { fprintf(stderr, "emu cudaSafeCall failed at %s:%d\n", file, line); abort(); }
}
