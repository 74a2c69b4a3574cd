// dev microbenchmark (round 2): random 32-byte sector loads from ONE CTA -- what bounds a phase of the merge
// scheduler?  Each thread issues U independent 256-bit loads at random sectors per round, then the block syncs.
// Prints cycles per round and cycles per sector for several footprints / U / CTA counts.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(32) S { unsigned v[8]; };
template <int U>
__global__ void __launch_bounds__(512, 1) k(const S* buf, unsigned nsec, int rounds, unsigned long long* out, unsigned* sink) {
  unsigned idx = 12345u + 7919u * threadIdx.x + 104729u * blockIdx.x;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < rounds; r++) {
    unsigned a[U];
#pragma unroll
    for (int u = 0; u < U; u++) { idx = idx * 1664525u + 1013904223u; a[u] = (idx >> 4) % nsec; }
    S s[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(s[u].v[0]), "=r"(s[u].v[1]), "=r"(s[u].v[2]), "=r"(s[u].v[3]), "=r"(s[u].v[4]), "=r"(s[u].v[5]), "=r"(s[u].v[6]), "=r"(s[u].v[7]) : "l"(buf + a[u]));
#pragma unroll
    for (int u = 0; u < U; u++) acc += s[u].v[0] ^ s[u].v[7];
    idx += acc & 1u;
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0xdeadbeefu) *sink = acc;
}
int main() {
  size_t bytes = 4ull << 30;
  S* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  unsigned long long* out; cudaMallocManaged(&out, 1024 * 8);
  unsigned* sink; cudaMalloc(&sink, 4);
  const int rounds = 2000;
  for (size_t mb : {32ull, 110ull, 1360ull, 4096ull}) {
    unsigned nsec = (unsigned)(mb * 1048576ull / 32);
    for (int ctas : {1, 2, 120}) {
      for (int U : {1, 2, 4, 8}) {
        for (int threads : {512, 128}) {
          if (U == 1) k<1><<<ctas, threads>>>(buf, nsec, rounds, out, sink);
          if (U == 2) k<2><<<ctas, threads>>>(buf, nsec, rounds, out, sink);
          if (U == 4) k<4><<<ctas, threads>>>(buf, nsec, rounds, out, sink);
          if (U == 8) k<8><<<ctas, threads>>>(buf, nsec, rounds, out, sink);
          cudaDeviceSynchronize();
          double cyc = (double)out[0] / rounds;
          printf("footprint %5zu MB ctas %3d threads %3d U %d: %8.0f cycles/round  %6.2f cycles/sector/SM\n", mb, ctas, threads, U, cyc, cyc / (threads * U));
        }
      }
    }
  }
  return 0;
}
