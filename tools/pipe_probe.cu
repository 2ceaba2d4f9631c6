// Micro-benchmark: issue throughput of the instruction classes the LDPC decoder can be built from, on this device.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CHAINS 8
#define REPS 16

template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint32_t b, uint32_t c)
{
  if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
  if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));
  if (OP == 5) asm volatile("min.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 6) asm volatile("min.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 7) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 9) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 10) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.b32 %0, %2, %0, p;}" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 11) a = __viaddmin_s32(a, b, c);
  if (OP == 12) a = __viaddmin_s16x2(a, b, c);
  if (OP == 13) a = __vimax3_s16x2(a, b, c);
  if (OP == 14) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 15) asm volatile("set.gt.u32.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 16) asm volatile("prmt.b32 %0, %0, %1, 0xba98;" : "+r"(a) : "r"(b));
  if (OP == 17) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 18) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if (OP == 19) asm volatile("sub.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if (OP == 20) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a));
  if (OP == 21) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
}

template <int OP_A, int OP_B>
__global__ void __launch_bounds__(256) probe(uint32_t* out, int iters)
{
  uint32_t a[CHAINS];
  uint32_t b = threadIdx.x * 2654435761u + 12345u, c = blockIdx.x * 40503u + 777u;
#pragma unroll
  for (int k = 0; k != CHAINS; ++k) a[k] = b * (k + 3);
  for (int i = 0; i != iters; ++i) {
#pragma unroll
    for (int r = 0; r != REPS; ++r) {
#pragma unroll
      for (int k = 0; k != CHAINS; ++k) {
        if (OP_B < 0 || ((k + r) & 1) == 0) step<OP_A>(a[k], b, c);
        else step<(OP_B < 0 ? 0 : OP_B)>(a[k], b, c);
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k != CHAINS; ++k) r ^= a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP_A, int OP_B>
void run(const char* name, uint32_t* d_out, int sms)
{
  const int blocks = sms * 8, threads = 256, iters = 1000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep != 4; ++rep) {
    cudaEventRecord(e0);
    probe<OP_A, OP_B><<<blocks, threads>>>(d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * CHAINS * REPS;
  printf("%-28s %7.2f Tera lane-instr/s   (%.3f ms)\n", name, ops / (best * 1e-3) / 1e12, best);
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  uint32_t* d;
  cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  int s = p.multiProcessorCount;
  run<0, -1>("lop3", d, s);
  run<1, -1>("add.u32", d, s);
  run<19, -1>("sub.u32", d, s);
  run<2, -1>("mad.lo.u32 (imad)", d, s);
  run<3, -1>("prmt", d, s);
  run<16, -1>("prmt sign-replicate", d, s);
  run<4, -1>("shf.l.wrap", d, s);
  run<20, -1>("shl imm", d, s);
  run<5, -1>("min.s32", d, s);
  run<6, -1>("min.s16x2", d, s);
  run<21, -1>("max.u16x2", d, s);
  run<17, -1>("add.s16x2", d, s);
  run<18, -1>("vabsdiff4", d, s);
  run<7, -1>("vabsdiff4.add", d, s);
  run<8, -1>("fma.f16x2", d, s);
  run<14, -1>("add.f16x2", d, s);
  run<9, -1>("min.f16x2", d, s);
  run<15, -1>("set.gt.f16x2", d, s);
  run<10, -1>("setp+selp (2 instr/step)", d, s);
  run<11, -1>("viaddmin_s32", d, s);
  run<12, -1>("viaddmin_s16x2", d, s);
  run<13, -1>("vimax3_s16x2", d, s);
  run<0, 2>("lop3 + imad", d, s);
  run<0, 1>("lop3 + add", d, s);
  run<0, 3>("lop3 + prmt", d, s);
  run<0, 8>("lop3 + fma.f16x2", d, s);
  run<3, 2>("prmt + imad", d, s);
  run<6, 2>("min.s16x2 + imad", d, s);
  run<6, 0>("min.s16x2 + lop3", d, s);
  run<18, 0>("vabsdiff4 + lop3", d, s);
  run<18, 2>("vabsdiff4 + imad", d, s);
  run<1, 2>("add + imad", d, s);
  run<8, 2>("fma.f16x2 + imad", d, s);
  run<9, 0>("min.f16x2 + lop3", d, s);
  run<12, 0>("viaddmin_s16x2 + lop3", d, s);
  return 0;
}
