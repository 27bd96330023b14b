// l1_half2.cu — issue-rate probe for the batched L1 (manhattan) tile kernel (csrc/batch_scan.cu):
// the fp32 inner loop (FADD + FADD|.| per pair-element, 8x8 register tile, operands from shared memory)
// against a packed-f16 variant (HSUB2 + HADD2|.| per TWO elements, flushed into fp32 every 16 elements).
// Answers one question before the kernel is written: does the packed form raise pair-elements/s on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/l1_half2.bin scripts/microbench/l1_half2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int LD = 129;

template <int TI, int TJ, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) probe(const float* __restrict__ in, float* out, int iters) {
    __shared__ float s_x[33 * LD];   // one spare row: the loop offsets reads by (it & 7)
    __shared__ float s_q[33 * LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < 33 * LD; i += 256) { s_x[i] = in[i % (32 * LD)]; s_q[i] = in[i % (32 * LD) + 32 * LD]; }
    __syncthreads();
    float acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) {
                float xv[TI], qv[TJ];
#pragma unroll
                for (int i = 0; i < TI; ++i) xv[i] = s_x[kk * LD + tx + 16 * i + (it & 7)];
#pragma unroll
                for (int j = 0; j < TJ; ++j) qv[j] = s_q[kk * LD + (ty + 16 * j) % 128 + (it & 7)];
#pragma unroll
                for (int i = 0; i < TI; ++i)
#pragma unroll
                    for (int j = 0; j < TJ; ++j) acc[i][j] += fabsf(xv[i] - qv[j]);
            }
        }
    } else {
        const __half2* hx = reinterpret_cast<const __half2*>(s_x);
        const __half2* hq = reinterpret_cast<const __half2*>(s_q);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                __half2 a2[TI][TJ];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int k2 = part * 8 + kk;
                    __half2 xv[TI], qv[TJ];
#pragma unroll
                    for (int i = 0; i < TI; ++i) xv[i] = hx[k2 * LD + tx + 16 * i + (it & 7)];   // `it`-dependent: nothing hoists out of the loop
#pragma unroll
                    for (int j = 0; j < TJ; ++j) qv[j] = hq[k2 * LD + (ty + 16 * j) % 128 + (it & 7)];
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) {
                            const __half2 d = __habs2(__hsub2(xv[i], qv[j]));
                            a2[i][j] = kk == 0 ? d : __hadd2(a2[i][j], d);
                        }
                }
#pragma unroll
                for (int i = 0; i < TI; ++i)
#pragma unroll
                    for (int j = 0; j < TJ; ++j) {
                        if (MODE == 1) {
                            const float2 f = __half22float2(a2[i][j]);
                            acc[i][j] += f.x;
                            acc[i][j] += f.y;
                        } else {
                            const __half h = __hadd(__low2half(a2[i][j]), __high2half(a2[i][j]));
                            acc[i][j] += __half2float(h);
                        }
                    }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) s += acc[i][j];
    out[blockIdx.x * 256 + tid] = s;
}

template <int TI, int TJ, int MODE, int MINB>
void run(const char* name, float* in, float* out) {
    const int iters = 1500, grid = 148 * MINB * 4;
    probe<TI, TJ, MODE, MINB><<<grid, 256>>>(in, out, 10);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<TI, TJ, MODE, MINB><<<grid, 256>>>(in, out, iters);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) {
        printf("{\"variant\": \"%s\", \"err\": \"%s\"}\n", name, cudaGetErrorString(cudaGetLastError()));
        exit(1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double elems = double(grid) * 256 * iters * 32 * TI * TJ;   // pair-elements
    printf("{\"variant\": \"%s\", \"tile\": \"%dx%d\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"pair_elements_per_s\": %.4e, \"err\": \"%s\"}\n",
           name, TI, TJ, MINB, ms, elems / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 2 * 32 * LD * 4);
    cudaMalloc(&out, 148 * 16 * 256 * 4);
    {   // distinct finite f16 pairs / small positive floats in every word
        unsigned* h = new unsigned[2 * 32 * LD];
        for (int i = 0; i < 2 * 32 * LD; ++i) h[i] = 0x38003400u + (unsigned(i * 2654435761u) & 0x03FF03FFu);
        cudaMemcpy(in, h, 2 * 32 * LD * 4, cudaMemcpyHostToDevice);
        delete[] h;
    }
    run<8, 8, 0, 2>("fp32 FADD + FADD|.|", in, out);
    run<8, 8, 1, 1>("f16x2 HSUB2 + HADD2|.|, flush 2 cvt + 2 FADD / 16 el", in, out);
    run<8, 8, 2, 1>("f16x2 HSUB2 + HADD2|.|, flush HADD + cvt + FADD / 16 el", in, out);
    run<8, 4, 1, 2>("f16x2, flush 2 cvt + 2 FADD / 16 el", in, out);
    run<8, 4, 2, 2>("f16x2, flush HADD + cvt + FADD / 16 el", in, out);
    return 0;
}
