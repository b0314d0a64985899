// Issue rates of the instructions the attention softmax warps are made of (MUFU.EX2 f32 / f16 / bf16, FFMA register and
// immediate form, F2FP pack), per SM sub-partition, on the GPU it runs on. Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates scripts/microbench/pipe_rates.cu && /tmp/pipe_rates
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
__global__ void k_f32(float* o, int n) { float x = threadIdx.x * 1e-3f; float a0=x,a1=x+1,a2=x+2,a3=x+3,a4=x+4,a5=x+5,a6=x+6,a7=x+7;
  for (int i = 0; i < n; ++i) {
#define E(a) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    E(a0) E(a1) E(a2) E(a3) E(a4) E(a5) E(a6) E(a7) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7; }
__global__ void k_h2(unsigned* o, int n) { unsigned a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  for (int i = 0; i < n; ++i) {
#define H(a) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
    H(a0) H(a1) H(a2) H(a3) H(a4) H(a5) H(a6) H(a7) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7; }
__global__ void k_b2(unsigned* o, int n) { unsigned a0=threadIdx.x,a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  for (int i = 0; i < n; ++i) {
#define B(a) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a));
    B(a0) B(a1) B(a2) B(a3) B(a4) B(a5) B(a6) B(a7) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7; }
__global__ void k_ffma(float* o, int n) { float x = threadIdx.x * 1e-3f; float a0=x,a1=x+1,a2=x+2,a3=x+3,a4=x+4,a5=x+5,a6=x+6,a7=x+7; float b = x * 0.5f, c = x + 0.25f;
  for (int i = 0; i < n; ++i) {
#define F(a) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    F(a0) F(a1) F(a2) F(a3) F(a4) F(a5) F(a6) F(a7) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7; }
__global__ void k_ffma_imm(float* o, int n) { float x = threadIdx.x * 1e-3f; float a0=x,a1=x+1,a2=x+2,a3=x+3,a4=x+4,a5=x+5,a6=x+6,a7=x+7; float b = x * 0.5f;
  for (int i = 0; i < n; ++i) {
#define G(a) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a) : "f"(b));
    G(a0) G(a1) G(a2) G(a3) G(a4) G(a5) G(a6) G(a7) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7; }
__global__ void k_f2fp(unsigned* o, int n) { float x = threadIdx.x * 1e-3f; float a0=x,a1=x+1,a2=x+2,a3=x+3; unsigned r0=0,r1=0,r2=0,r3=0,r4=0,r5=0,r6=0,r7=0;
  for (int i = 0; i < n; ++i) {
#define P(r,a,b) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); a = __uint_as_float(r);
    P(r0,a0,a1) P(r1,a1,a2) P(r2,a2,a3) P(r3,a3,a0) P(r4,a0,a2) P(r5,a1,a3) P(r6,a2,a0) P(r7,a3,a1) }
  o[blockIdx.x * blockDim.x + threadIdx.x] = r0+r1+r2+r3+r4+r5+r6+r7; }
template <class K, class T> void run(const char* name, K k, T* o, int warps_per_sm, double elems_per_instr) {
  int n = 4096; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, warps_per_sm * 32>>>(o, 16); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<148, warps_per_sm * 32>>>(o, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double instr = (double)n * 8 * warps_per_sm;  // warp-instructions per SM
  printf("%-10s warps/SM %2d: %.3f ms, %.2f warp-instr/us/SM  (at %.2f GHz nominal: %.2f clk per warp-instr per SMSP)\n", name, warps_per_sm, ms,
         instr / (ms * 1e3), clk * 1e-6, (ms * 1e-3 * clk * 1e3) / (instr / 4));
}
int main() { float* o; cudaMalloc(&o, 1 << 24);
  for (int w : {4, 8, 16}) { run("ex2.f32", k_f32, o, w, 1); run("ex2.f16x2", k_h2, (unsigned*)o, w, 2); run("ex2.bf16x2", k_b2, (unsigned*)o, w, 2);
    run("ffma", k_ffma, o, w, 1); run("ffma.imm", k_ffma_imm, o, w, 1); run("f2fp.bf16", k_f2fp, (unsigned*)o, w, 1); }
  return 0; }
