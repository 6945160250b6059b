#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* ctr, cudaGraphConditionalHandle h) {
  int v = atomicAdd(ctr, 1);
  cudaGraphSetConditional(h, v + 1 < 5);
}
int main() {
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
  cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
  cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &cp);
  printf("add %d\n", (int)e);
  cudaGraph_t bg = cp.conditional.phGraph_out[0];
  cudaStream_t s; cudaStreamCreate(&s);
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  e = cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal);
  body<<<1,1,0,s>>>(d, h);
  cudaGraph_t out; e = cudaStreamEndCapture(s, &out);
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", (int)e);
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  int hv; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost); printf("iterations %d\n", hv);
}
