// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition on sm_100a (one or two warps per sub-partition), alone and in the
// softmax instruction mix (FFMA + EX2 + FADD + pack).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// the packed form: two exponentials per lane and instruction (inputs packed by cvt.rn.bf16x2.f32, the result is already the bf16x2 the MMA wants)
__global__ void k2(float* out, long long* cyc, float a, float b, int iters, int pack_inputs) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; i++) s[i] = a * (float)(i + threadIdx.x) * 1e-3f;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            unsigned xin;
            if (pack_inputs) { __nv_bfloat162 v = __floats2bfloat162_rn(fmaf(s[i], a, b), fmaf(s[i + 1], a, b)); xin = *reinterpret_cast<unsigned*>(&v); }
            else xin = __float_as_uint(s[i]);
            const unsigned e = ex2_bf16x2(xin);
            acc ^= e;
            s[i] = __uint_as_float(e << 16) - 1.0f; s[i + 1] = __uint_as_float(e & 0xffff0000u) - 1.0f;
        }
    }
    long long t1 = clock64();
    float r = __uint_as_float(acc);
    for (int i = 0; i < 64; i++) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b, int iters) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 64; i++) s[i] = a * (float)(i + threadIdx.x) * 1e-3f;
    float l0 = 0.f, l1 = 0.f;
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            if (MODE == 0) { s[i] = ex2(s[i]); s[i + 1] = ex2(s[i + 1]); }
            else if (MODE == 3) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[i])); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[i + 1])); }
            else if (MODE == 4) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[i + 1])); }
            else {
                float e0 = ex2(fmaf(s[i], a, b)), e1 = ex2(fmaf(s[i + 1], a, b));
                l0 += e0; l1 += e1;
                if (MODE == 2) { __nv_bfloat162 v = __floats2bfloat162_rn(e0, e1); acc ^= *reinterpret_cast<unsigned*>(&v); }
                s[i] = e0 - 1.0f; s[i + 1] = e1 - 1.0f;
            }
        }
    }
    long long t1 = clock64();
    float r = l0 + l1 + __uint_as_float(acc);
    for (int i = 0; i < 64; i++) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int iters = 200;
    for (int mode = 0; mode < 5; mode++)
        for (int threads : {128, 256, 512}) {
            long long c = 0;
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters);
                if (mode == 1) k<1><<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters);
                if (mode == 2) k<2><<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters);
                if (mode == 3) k<3><<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters);
                if (mode == 4) k<4><<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double per_sub = (double)iters * 64 * (threads / 128);      // warp-wide ex2 per sub-partition
            printf("mode %d (%s) warps/subpartition %d: %.2f cycles per warp-wide EX2 per sub-partition (%s)\n", mode,
                   mode == 0 ? "ex2 only" : mode == 1 ? "ffma+ex2+fadd+fadd" : mode == 2 ? "ffma+ex2+fadd+fadd+pack" : mode == 3 ? "tanh.approx only" : "rcp.approx only", threads / 128, (double)c / per_sub, cudaGetErrorString(cudaGetLastError()));
        }
    for (int pk = 0; pk < 2; pk++)
        for (int threads : {128, 256}) {
            long long c = 0;
            for (int rep = 0; rep < 2; rep++) { k2<<<148, threads>>>(out, cyc, -0.5f, -0.25f, iters, pk); cudaDeviceSynchronize(); }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("ex2.bf16x2 (%s) warps/subpartition %d: %.2f cycles per PAIR-wide warp instruction = per 64 exponentials (%s)\n", pk ? "inputs packed from fp32: ffma x2 + cvt + ex2" : "ex2 only",
                   threads / 128, (double)c / ((double)iters * 32 * (threads / 128)), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
