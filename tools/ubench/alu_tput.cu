// Development aid: issue throughput of the integer SIMD min/max family on sm_100a (which pipe bounds k_fast?).
// Build + run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/alu_tput tools/ubench/alu_tput.cu && /tmp/alu_tput
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u;
  uint32_t b = blockIdx.x * 97u + 13u, c = threadIdx.x ^ 0x55aa55aau;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP == 0) a[i] = __vimax3_u16x2(a[i], b, c);
        if (OP == 1) a[i] = __vmaxu2(a[i], b);
        if (OP == 2) a[i] = (uint32_t)max((int)a[i], (int)b) + 1u;  // IMNMX + IADD (keeps it from folding)
        if (OP == 3) a[i] = (uint32_t)__vimax3_s32((int)a[i], (int)b, (int)c);
        if (OP == 4) a[i] = __byte_perm(a[i], b, c);
        if (OP == 5) a[i] = (a[i] ^ b) & c | (a[i] >> 3);
        if (OP == 6) {
          __half2 h = __hmax2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b));
          a[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (OP == 7) a[i] = a[i] * b + c;  // IMAD (fma pipe)
        if (OP == 8) a[i] = __vimin3_u16x2(__vimax3_u16x2(a[i], b, c), b, c);
        if (OP == 9) a[i] = __dp4a(a[i], b, c);
        if (OP == 10) a[i] = __vimax3_u16x2(a[i], b, c) * 3u + c;  // VIMNMX3 + IMAD interleaved (two pipes)
        if (OP == 11) {  // HFMA2.RELU: max(a, b) = relu(a - b) + b costs two of these on the fma pipe
          __half2 h = __hfma2_relu(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b),
                                   *reinterpret_cast<__half2*>(&c));
          a[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (OP == 12) {  // independent chains: even registers VIMNMX3 (alu pipe), odd registers HFMA2.RELU (fma pipe)
          if (i & 1) {
            __half2 h = __hfma2_relu(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b),
                                     *reinterpret_cast<__half2*>(&c));
            a[i] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            a[i] = __vimax3_u16x2(a[i], b, c);
          }
        }
        if (OP == 13) {  // HMNMX2 and VIMNMX3 on independent chains: same pipe or not?
          if (i & 1) {
            __half2 h = __hmax2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&b));
            a[i] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            a[i] = __vimax3_u16x2(a[i], b, c);
          }
        }
        if (OP == 14) a[i] = __viaddmax_u16x2(a[i], b, c);
        if (OP == 15) a[i] = __vmaxu4(a[i], b);
      }
      b += 0x00010001u;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int ops_per_inner) {
  uint32_t* d;
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  cudaMalloc(&d, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(d, 10);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * threads * iters * 64.0 * ops_per_inner;  // thread-level ops
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, n / (ms * 1e-3) / (clk * 1e3) / 148.0,
         clk / 1000);
  cudaFree(d);
}

int main() {
  run<0>("VIMNMX3.U16x2", 1);
  run<1>("vmaxu2 (2-input u16x2)", 1);
  run<2>("IMNMX.S32 + IADD", 2);
  run<3>("VIMNMX3.S32", 1);
  run<4>("PRMT", 1);
  run<5>("LOP3+SHF mix", 1);
  run<6>("HMNMX2 (half2 max)", 1);
  run<7>("IMAD", 1);
  run<8>("VIMNMX3 x2 dependent", 2);
  run<9>("IDP.4A", 1);
  run<10>("VIMNMX3 + IMAD (two pipes)", 2);
  run<11>("HFMA2.RELU", 1);
  run<12>("VIMNMX3 | HFMA2.RELU independent", 1);
  run<13>("VIMNMX3 | HMNMX2 independent", 1);
  run<14>("VIADDMNMX.U16x2", 1);
  run<15>("vmaxu4 (byte SIMD)", 1);
  return 0;
}
