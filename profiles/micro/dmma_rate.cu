// DMMA issue-rate micro-benchmark: how many warps per SM (and accumulators per warp) does mma.sync.m8n8k4.f64 need to saturate?
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(int iters, double* out) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
void run(int warps, int sms) {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<NACC><<<sms, warps * 32>>>(iters, out);
    cudaEventRecord(e0);
    k<NACC><<<sms, warps * 32>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dmma = (double)sms * warps * iters * NACC;
    printf("warps/SM=%2d acc/warp=%2d : %.2f TFLOP/s, %.1f cycles per DMMA per SMSP (at 1.9 GHz)\n", warps, NACC, dmma * 512 / (ms * 1e-3) / 1e12,
           (ms * 1e-3) * 1.9e9 / (dmma / (sms * 4.0)));
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) { run<1>(w, sms); run<4>(w, sms); run<16>(w, sms); }
    return 0;
}
