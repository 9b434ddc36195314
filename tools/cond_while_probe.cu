#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* it, cudaGraphConditionalHandle h, int maxit) {
    int v = ++(*it);
    cudaGraphSetConditional(h, v < maxit ? 1u : 0u);
}
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int* it; cudaMalloc(&it, 4); cudaMemset(it, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(st, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    printf("begin: %s\n", cudaGetErrorString(e));
    body<<<1,1,0,st>>>(it, h, 7);
    e = cudaStreamEndCapture(st, nullptr);
    printf("end: %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
    cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    int hv; cudaMemcpy(&hv, it, 4, cudaMemcpyDeviceToHost); printf("iterations %d\n", hv);
}
