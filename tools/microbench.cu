// Micro-benchmarks behind the decoder's roofline figures: what one B200 SM actually sustains for
//   * shared-memory loads (LDS.32 / LDS.64 / LDS.128, conflict free)      -> the "128 B/clk/SM" denominator of roofline.frac
//   * MUFU.RCP                                                             -> 16 per clk per SM ?
//   * ALU-pipe ops (LOP3, FMNMX, FSEL)                                      -> 64 per clk per SM (half rate) ?
//   * FMA-pipe ops (FFMA, packed FFMA2)                                     -> 128 per clk per SM ?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int T = 1024, ITERS = 4096;

template <int W>  // W = words per load: 1, 2, 4
__global__ void __launch_bounds__(T, 2) lds_kernel(float *out) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += T) sm[i] = (float)i;
  __syncthreads();
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  // every lane reads its own W consecutive words: conflict free; 8 independent loads per iteration
  // (the offset walks through the 32 KB window with a stride of 1056 words so that no two loads of the loop share an
  // address — ptxas merges or hoists them otherwise — while a warp still reads 32 x W consecutive words)
  const int base = (threadIdx.x * W) & 8191;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int a = (base + (it * 8 + u) * 1056) & (8191 & ~(W - 1));
      if (W == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + a))); acc0 += v; }
      if (W == 2) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"((unsigned)__cvta_generic_to_shared(sm + a))); acc0 += v; acc1 += w; }
      if (W == 4) { float v, w, x, y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"((unsigned)__cvta_generic_to_shared(sm + a))); acc0 += v; acc1 += w; acc2 += x; acc3 += y; }
    }
  }
  out[blockIdx.x * T + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// stores, and the decoder's mix (one load + one store per word, both conflict free, addresses walking the window)
template <int W, bool MIX>
__global__ void __launch_bounds__(T, 2) sts_kernel(float *out) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += T) sm[i] = (float)i;
  __syncthreads();
  float acc = 0.f;
  const int base = (threadIdx.x * W) & 8191;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int a = (base + (it * 8 + u) * 1056) & (8191 & ~(W - 1));
      const unsigned sa = (unsigned)__cvta_generic_to_shared(sm + a);
      float v = acc + (float)u;
      if (MIX) {
        if (W == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa));
        if (W == 2) { float w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(sa)); v += w; }
        acc += v;
      }
      if (W == 1) asm volatile("st.shared.f32 [%0], %1;" :: "r"(sa), "f"(v) : "memory");
      if (W == 2) asm volatile("st.shared.v2.f32 [%0], {%1,%1};" :: "r"(sa), "f"(v) : "memory");
      if (W == 4) asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" :: "r"(sa), "f"(v) : "memory");
    }
  }
  out[blockIdx.x * T + threadIdx.x] = acc + sm[threadIdx.x];
}

enum Op { OP_RCP, OP_LOP3, OP_FMNMX, OP_FSEL, OP_FFMA, OP_FFMA2 };
template <int OP>
__global__ void __launch_bounds__(T, 2) alu_kernel(float *out, float seed) {
  float a[8];
  uint32_t b[8];
#pragma unroll
  for (int u = 0; u < 8; u++) { a[u] = seed + threadIdx.x * 1e-3f + u; b[u] = threadIdx.x * 2654435761u + u; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (OP == OP_RCP) asm volatile("{.reg .f32 t; add.f32 t, %0, 0f3FC00000; rcp.approx.ftz.f32 %0, t;}" : "+f"(a[u]));  // + 1 FADD (FMA pipe)
      if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, 0x9e3779b9, 0x96;" : "+r"(b[u]) : "r"(b[(u + 1) & 7]));
      if (OP == OP_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[u]) : "f"(a[(u + 1) & 7]));
      if (OP == OP_FSEL) asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %0; selp.f32 %0, %2, %0, p;}" : "+f"(a[u]) : "f"(a[(u + 1) & 7]), "f"(a[(u + 2) & 7]));  // FSETP + FSEL
      if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[u]) : "f"(seed));
      if (OP == OP_FFMA2) asm volatile("{.reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0,%1}, x;}" : "+f"(a[u]), "+f"(a[(u + 4) & 7]) : "f"(seed));
    }
  }
  float s = 0.f; uint32_t x = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) { s += a[u]; x ^= b[u]; }
  out[blockIdx.x * T + threadIdx.x] = s + (float)x;
}

template <class F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = pr.multiProcessorCount, grid = sms * 2;
  const double ghz = clk_khz * 1e-6;
  float *out; CK(cudaMalloc(&out, sizeof(float) * grid * T));
  printf("%s: %d SMs, max SM clock %.3f GHz (per-clock figures assume the max clock; the driver's clock sample is in bench.py)\n", pr.name, sms, ghz);
  const double loads = (double)grid * T * ITERS * 8;
#define LDS(W) { CK(cudaFuncSetAttribute(lds_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768)); \
    float ms = time_ms([&] { lds_kernel<W><<<grid, T, 32768>>>(out); }); CK(cudaGetLastError()); \
    double bps = loads * W * 4 / (ms * 1e-3); \
    printf("LDS.%-3d  %8.1f GB/s  = %6.1f B/clk/SM\n", 32 * W, bps * 1e-9, bps / sms / (ghz * 1e9)); }
  LDS(1) LDS(2) LDS(4)
#define STS(W, MIX, name) { CK(cudaFuncSetAttribute(sts_kernel<W, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768)); \
    float ms = time_ms([&] { sts_kernel<W, MIX><<<grid, T, 32768>>>(out); }); CK(cudaGetLastError()); \
    double bps = loads * W * 4 * (MIX ? 2 : 1) / (ms * 1e-3); \
    printf("%-16s %8.1f GB/s  = %6.1f B/clk/SM\n", name, bps * 1e-9, bps / sms / (ghz * 1e9)); }
  STS(1, false, "STS.32") STS(2, false, "STS.64") STS(4, false, "STS.128") STS(1, true, "LDS.32+STS.32") STS(2, true, "LDS.64+STS.64")
  const double ops = (double)grid * T * ITERS * 8;
#define ALU(OP, name, mult) { float ms = time_ms([&] { alu_kernel<OP><<<grid, T>>>(out, 1.0001f); }); CK(cudaGetLastError()); \
    double ps = ops * mult / (ms * 1e-3); printf("%-10s %8.2f T lane-ops/s = %6.1f per clk per SM\n", name, ps * 1e-12, ps / sms / (ghz * 1e9)); }
  ALU(OP_RCP, "MUFU.RCP", 1) ALU(OP_LOP3, "LOP3", 1) ALU(OP_FMNMX, "FMNMX", 1) ALU(OP_FSEL, "FSETP+FSEL", 2) ALU(OP_FFMA, "FFMA", 1) ALU(OP_FFMA2, "FFMA2", 1)
  printf("(FFMA2: one instruction = 2 fused multiply-adds per lane; the line counts instructions x lanes)\n");
  return 0;
}
