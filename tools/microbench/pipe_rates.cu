// Throughput of the integer instructions the step kernel is made of, per SM sub-partition, on the device at hand.
// Each thread runs 8 independent dependency chains of one instruction kind; 8 warps per SMSP keep the pipes full.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kIters = 4096, kChains = 8;

template <int kOp>
__global__ void __launch_bounds__(256) rate_kernel(uint32_t* out, uint32_t seed) {
  uint32_t a[kChains], b[kChains];
  double d[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) { a[c] = seed + threadIdx.x * 9u + c; b[c] = seed * 3u + c; d[c] = double(a[c]); }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
      if (kOp == 0) {          // IMAD.WIDE.U32: 32x32 -> 64, both halves consumed by the next multiply of the chain
        uint64_t p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[c]), "r"(0xD2511F53u));
        a[c] = uint32_t(p) + 1u; b[c] ^= uint32_t(p >> 32);
        // the two extra ops (add, xor) are subtracted using the kOp 2 / 3 rates; see main()
      } else if (kOp == 1) {   // IMAD (mad.lo)
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(0xD2511F53u), "r"(b[c]));
      } else if (kOp == 2) {   // LOP3 (xor of three)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
      } else if (kOp == 3) {   // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
      } else if (kOp == 4) {   // IMAD.HI
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(0xCD9E8D57u));
      } else if (kOp == 5) {   // DADD
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.0000001));
      } else if (kOp == 6) {   // bare mul.wide, low half only feeds the chain (hi accumulated with xor every 4th)
        uint64_t p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[c]), "r"(0xD2511F53u));
        a[c] = uint32_t(p) ^ uint32_t(p >> 32);
      } else if (kOp == 7) {   // one Philox round on (a, b) pairs: 2 IMAD.WIDE + 2 LOP3
        uint64_t p0, p1;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a[c]), "r"(0xD2511F53u));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(b[c]), "r"(0xCD9E8D57u));
        a[c] = uint32_t(p1 >> 32) ^ uint32_t(p0) ^ seed;
        b[c] = uint32_t(p0 >> 32) ^ uint32_t(p1) ^ (seed + 1u);
      } else if (kOp == 8) {   // SHF
        asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(7u));
      } else if (kOp == 9) {   // ISETP + SEL
        asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc ^= a[c] ^ b[c] ^ uint32_t(__double2loint(d[c]));
  if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <int kOp>
double run(const char* name, double ops_per_iter, int sms, double clock_hz, uint32_t* out) {
  const int blocks = sms * 8;   // 8 CTAs x 8 warps per SM = 16 warps per SMSP
  rate_kernel<kOp><<<blocks, 256>>>(out, 12345u);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<kOp><<<blocks, 256>>>(out, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_ops = double(blocks) * 8 * kIters * kChains * ops_per_iter;
  const double per_smsp_clk = warp_ops / (ms * 1e-3 * clock_hz) / (sms * 4);
  printf("%-34s %8.3f ms   %6.3f warp-instr/clk/SMSP (of the listed kind%s)\n", name, ms, per_smsp_clk,
         ops_per_iter > 1 ? "s, summed" : "");
  return per_smsp_clk;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double hz = khz * 1e3;
  printf("%s, %d SMs, %.0f MHz (nominal max; rates below assume it)\n", p.name, p.multiProcessorCount, hz / 1e6);
  uint32_t* out;
  cudaMalloc(&out, 4096);
  run<1>("IMAD (mad.lo.u32)", 1, p.multiProcessorCount, hz, out);
  run<2>("LOP3 (3-input xor)", 1, p.multiProcessorCount, hz, out);
  run<3>("IADD", 1, p.multiProcessorCount, hz, out);
  run<8>("SHF", 1, p.multiProcessorCount, hz, out);
  run<9>("ISETP + SEL", 2, p.multiProcessorCount, hz, out);
  run<4>("IMAD.HI (mul.hi.u32)", 1, p.multiProcessorCount, hz, out);
  run<6>("IMAD.WIDE.U32 + LOP3", 2, p.multiProcessorCount, hz, out);
  run<0>("IMAD.WIDE.U32 + IADD + LOP3", 3, p.multiProcessorCount, hz, out);
  run<7>("Philox round (2 IMAD.WIDE + 2 LOP3)", 4, p.multiProcessorCount, hz, out);
  run<5>("DADD", 1, p.multiProcessorCount, hz, out);
  return 0;
}
