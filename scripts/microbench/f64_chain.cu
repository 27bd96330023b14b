// Dependent f64 add-chain latency on the device (one warp): the floor of a reference-order f64 re-score.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f64_chain f64_chain.cu && ./f64_chain
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(const double* x, double* out, long long* cyc, int n, int mode) {
    double a0 = x[0], a1 = x[1];
    const double y = x[2], z = x[3];
    long long t0 = clock64();
    if (mode == 0) {            // one DADD chain
#pragma unroll 16
        for (int i = 0; i < n; ++i) a0 = __dadd_rn(a0, y);
    } else if (mode == 1) {     // DADD chain fed by independent DMULs (dot product in reference order)
#pragma unroll 16
        for (int i = 0; i < n; ++i) a0 = __dadd_rn(a0, __dmul_rn(y + i, z));
    } else if (mode == 2) {     // two independent chains (cosine: dot and sum of squares)
#pragma unroll 16
        for (int i = 0; i < n; ++i) { a0 = __dadd_rn(a0, __dmul_rn(y + i, z)); a1 = __dadd_rn(a1, __dmul_rn(y + i, y + i)); }
    } else {                    // fp32 FADD chain for comparison
        float f = static_cast<float>(a0), g = static_cast<float>(y);
#pragma unroll 16
        for (int i = 0; i < n; ++i) f = __fadd_rn(f, g);
        a0 = f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
}
int main() {
    double hx[4] = {1.0, 2.0, 1e-9, 3.0}, *dx, *dout; long long* dc; long long hc[1];
    cudaMalloc(&dx, 32); cudaMalloc(&dout, 8 * 1024); cudaMalloc(&dc, 8 * 32);
    cudaMemcpy(dx, hx, 32, cudaMemcpyHostToDevice);
    const int n = 4096;
    const char* names[4] = {"DADD chain", "DADD chain + independent DMUL", "two DADD chains + DMULs", "FADD chain"};
    for (int threads : {32, 96, 1024})
        for (int mode = 0; mode < 4; ++mode) {
            chain<<<1, threads>>>(dx, dout, dc, n, mode); chain<<<1, threads>>>(dx, dout, dc, n, mode);
            cudaDeviceSynchronize(); cudaMemcpy(hc, dc, 8, cudaMemcpyDeviceToHost);
            printf("threads=%4d  %-32s %.1f cycles per element\n", threads, names[mode], double(hc[0]) / n);
        }
    return cudaGetLastError() != cudaSuccess;
}
