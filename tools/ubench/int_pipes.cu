// Integer-instruction throughput on sm_100a: which of the ops the DP cell uses share the ALU pipe, and at what rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
template <int OP>
__global__ void __launch_bounds__(256) k(int* out, int b, int c, int one, int m) {
  int a[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = threadIdx.x + i * b;
  for (int it = 0; it < ITERS; it++) {
    const int bi = b + it, mi = m ^ it, ci = c ^ (it << 3);   // loop-varying operands, shared by the chains
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == 0) a[i] = a[i] + bi;                                  // add
      if (OP == 1) a[i] = a[i] * one + bi;                            // IMAD
      if (OP == 2) a[i] = (a[i] & mi) | ci;                            // LOP3
      if (OP == 3) a[i] = max(a[i] + 1, bi);                            // needs an add too; see OP 12
      if (OP == 4) a[i] = max(max(a[i] ^ 1, bi), ci);                  // max3 (+xor)
      if (OP == 5) a[i] = a[i] > bi ? ci : a[i] + 1;                   // ISETP + SEL + add
      if (OP == 6) a[i] = max(a[i] + bi, ci);                          // VIADDMNMX
      if (OP == 7) { a[i] = a[i] * one + bi; a[i] = (a[i] & mi) | ci; }  // IMAD + LOP3
      if (OP == 8) a[i] = __funnelshift_r(a[i], bi, ci);               // SHF
      if (OP == 9) a[i] = __byte_perm(a[i], bi, ci);                   // PRMT
      if (OP == 10) a[i] = min(a[i] ^ 1, mi);                             // pure VIMNMX chain (value saturates, still issued)
      if (OP == 11) { a[i] = (a[i] & mi) | ci; a[i] = max(a[i], bi); }   // LOP3 + VIMNMX
      if (OP == 12) { a[i] = a[i] * one + bi; a[i] = max(a[i], ci); }   // IMAD + VIMNMX
      if (OP == 13) { a[i] = a[i] * one + bi; a[i] = a[i] > ci ? a[i] : mi; } // IMAD + ISETP + SEL
      if (OP == 14) { if (a[i] & ci) a[i] = a[i] * one + bi; }          // LOP3.P + predicated IMAD
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int n_instr, int* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int grid = sm * 8;
  k<OP><<<grid, 256>>>(d, 3, 5, 1, 0x7ffffff0);
  cudaEventRecord(e0);
  k<OP><<<grid, 256>>>(d, 3, 5, 1, 0x7ffffff0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 * ITERS * CH * n_instr;
  const double per_clk_sm = warp_instr / (ms * 1e-3 * clk * 1e3) / sm;
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM (of the %d counted ops)  %6.1f lanes/clk/SM\n", name, ms, per_clk_sm, n_instr, per_clk_sm * 32);
}
int main() {
  int* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 2);
  run<0>("add", 1, d); run<1>("IMAD", 1, d); run<2>("LOP3", 1, d); run<3>("VIADDMNMX(a+1,b)", 1, d); run<4>("xor+VIMNMX3", 2, d);
  run<5>("ISETP+SEL+add", 3, d); run<6>("VIADDMNMX", 1, d); run<7>("IMAD+LOP3", 2, d); run<8>("SHF", 1, d); run<9>("PRMT", 1, d);
  run<10>("xor+VIMNMX", 2, d); run<11>("LOP3+VIMNMX", 2, d); run<12>("IMAD+VIMNMX", 2, d); run<13>("IMAD+ISETP+SEL", 3, d); run<14>("LOP3.P+@P IMAD", 2, d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
