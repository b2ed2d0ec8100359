#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *cnt, cudaGraphConditionalHandle h) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int c = ++*cnt;
        cudaGraphSetConditional(h, c < 5 ? 1u : 0u);
    }
}
int main() {
    int *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; 
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t bg = p.conditional.phGraph_out[0];
    cudaStream_t s; cudaStreamCreate(&s);
    cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal);
    body<<<1, 32, 0, s>>>(d, h);
    cudaStreamEndCapture(s, nullptr);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
    e = cudaGraphLaunch(ex, s); cudaStreamSynchronize(s); printf("launch: %s\n", cudaGetErrorString(e));
    int hc; cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost); printf("count %d\n", hc);
}
