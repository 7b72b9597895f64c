// pdl_gap.cu — what one dependent-kernel boundary costs inside a CUDA graph on B200, with ordinary edges and with programmatic
// dependent launch (griddepcontrol.wait at the top of the consumer, launch_dependents at the top of the producer).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_gap tools/probe/pdl_gap.cu ; run: /tmp/pdl_gap
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <bool PDL>
__global__ void k_step(const float* __restrict__ in, float* __restrict__ out, int n, int work) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = in[i];
    for (int k = 0; k < work; ++k) v = v * 1.0001f + 0.5f;
    out[i] = v;
  }
}

template <bool PDL>
static int run(const char* what, int ctas, int work, int smem) {
  const int n = ctas * 256, NODES = 200;
  float *a, *b;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  CK(cudaFuncSetAttribute(k_step<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < NODES; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
    CK(cudaLaunchKernelEx(&cfg, k_step<PDL>, in, out, n, work));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-10s ctas %4d work %5d smem %6d: %.2f us per kernel node\n", what, ctas, work, smem, ms * 1e3f / (20 * NODES));
  fflush(stdout);
  return 0;
}

int main() {
  for (int smem : {0, 100 * 1024, 200 * 1024})
    for (int ctas : {40, 148, 1024})
      for (int work : {0, 2000}) {
        if (run<false>("plain", ctas, work, smem)) return 1;
        if (run<true>("pdl", ctas, work, smem)) return 1;
      }
  return 0;
}
