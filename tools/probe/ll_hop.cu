// Micro-probe: what does one cross-CTA hop cost on B200?  (tools/probe, not product code)
//   pingpong: CTA 0 and CTA k bounce an {epoch,value} word N times -> round trip / 2 = one store->poll hop
//   fanin:    P producer warps (spread over CTAs) each store one word; C consumer CTAs poll all P words, then all CTAs repeat
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned v, unsigned ep) {
  unsigned long long w = ((unsigned long long)ep << 32) | v;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_peek(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__global__ void pingpong(unsigned long long* buf, int iters, int peer, long long* out) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      ll_store(buf, i, i);
      while ((unsigned)(ll_peek(buf + 16) >> 32) != (unsigned)i) {}
    }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) {
      while ((unsigned)(ll_peek(buf) >> 32) != (unsigned)i) {}
      ll_store(buf + 16, i, i);
    }
  }
}
// every CTA: wait for all P words of round r (written by warps gw < P, blocked mapping), then (if it owns tasks) write round r+1
__global__ void __launch_bounds__(512, 1) fanin(unsigned long long* buf, int P, int C, int rounds, long long* out) {
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, gw = blockIdx.x * 16 + wid;
  __shared__ float xs[2048];
  long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    unsigned long long* cur = buf + (size_t)r * 1024;  // one slot set per round: nobody can overwrite a word a slower CTA still waits for
    if (gw < P && lane == 0) ll_store(cur + gw, r, r);
    if (blockIdx.x < C) {
      for (int k = tid; k < P; k += 512) {
        unsigned long long w;
        do { w = ll_peek(cur + k); } while ((unsigned)(w >> 32) != (unsigned)r);
        xs[k] = (float)(unsigned)w;
        if (xs[k] < 0) out[1] = 1;
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0 && tid == 0) out[0] = clock64() - t0;
}
__global__ void __launch_bounds__(512, 1) syncs(int n, long long* out, float* sink) {
  long long t0 = clock64();
  float v = threadIdx.x;
  for (int i = 0; i < n; ++i) {
    __syncthreads();
    v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
  }
  if (threadIdx.x == 0) out[0] = clock64() - t0;
  sink[threadIdx.x] = v;
}
int main() {
  unsigned long long* buf; long long* out; float* sink;
  cudaMalloc(&buf, 8 << 20); cudaMemset(buf, 0, 8 << 20); cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  long long h;
  for (int peer : {1, 2, 37, 74, 147}) {
    cudaMemset(buf, 0, 1 << 20);
    int iters = 2000; void* args[] = {&buf, &iters, &peer, &out};
    cudaLaunchCooperativeKernel((void*)pingpong, dim3(148), dim3(32), args, 0, 0);
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    fflush(stdout); printf("pingpong cta0<->cta%d: %.0f cycles per hop (%.0f ns at %d kHz)\n", peer, h / 2.0 / iters, h / 2.0 / iters * 1e6 / clk, clk);
  }
  for (int P : {288, 768}) for (int C : {1, 18, 48, 148}) {
    cudaMemset(buf, 0, 8 << 20);
    int rounds = 500; void* args[] = {&buf, &P, &C, &rounds, &out};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)fanin, dim3(148), dim3(512), args, 0, 0);
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    fflush(stdout); printf("fanin P=%d words, %d polling CTAs: %.0f cycles per round (%.0f ns) %s\n", P, C, (double)h / rounds, (double)h / rounds * 1e6 / clk, cudaGetErrorString(e));
  }
  { int n = 1000; void* args[] = {&n, &out, &sink};
    cudaLaunchKernel((void*)syncs, dim3(1), dim3(512), args, 0, 0); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("syncthreads + 5-step shuffle tree, 512 threads: %.0f cycles\n", (double)h / n); }
  return 0;
}
