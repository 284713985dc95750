// Does the scheduler issue other instructions in the shadow of an FP64 instruction?  (A DFMA occupies the FP64 pipe of its SM
// sub-partition for two cycles: measured peak 0.4965 warp-instr / clk / SMSP, scripts/micro/fp64_peak.cu.)
// Kernel: per thread 8 independent DFMA chains, and per DFMA `K` independent 32-bit integer FMAs (IMAD) / shared-memory loads on
// separate chains.  If time stays at the DFMA-only time while K grows to 1, the other pipes issue in the shadow; if it grows like
// (2 + K) / 2, every instruction takes its own issue slot(s).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/fp64_coissue scripts/micro/fp64_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool LDS>
__global__ void mix_kernel(int iters, double a, double b, int ia, int ib, double *out) {
    __shared__ int sm[1024];
    sm[threadIdx.x & 1023] = threadIdx.x;
    __syncthreads();
    double v[8];
    int q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { v[c] = threadIdx.x * 1e-3 + c; q[c] = threadIdx.x + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                v[c] = fma(v[c], a, b);
                if (K >= 1) { if (LDS) q[c] = sm[(q[c] + ia) & 1023]; else q[c] = q[c] * ia + ib; }
                if (K >= 2) { if (LDS) q[(c + 1) & 7] ^= sm[(q[c] + ib) & 1023]; else q[(c + 3) & 7] = q[(c + 3) & 7] * ib + ia; }
            }
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { s += v[c]; t += q[c]; }
    if (s == 123.456 || t == 0x7fffffff) out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int K, bool LDS> float run(int sms, int warps, int iters, double *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<K, LDS><<<sms, 32 * warps>>>(iters / 8, 1.0000001, 1e-9, 3, 7, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        mix_kernel<K, LDS><<<sms, 32 * warps>>>(iters, 1.0000001, 1e-9, 3, 7, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    const int iters = 2048;
    printf("{\"device\": \"%s\", \"rows\": [", p.name);
    bool first = true;
    for (int warps : {8, 16}) {
        float t0 = run<0, false>(sms, warps, iters, out);
        float t1 = run<1, false>(sms, warps, iters, out), t2 = run<2, false>(sms, warps, iters, out);
        float l1 = run<1, true>(sms, warps, iters, out), l2 = run<2, true>(sms, warps, iters, out);
        printf("%s{\"warps_per_sm\": %d, \"dfma_only_ms\": %.4f, \"plus_1_imad_per_dfma_ms\": %.4f, \"plus_2_imad_per_dfma_ms\": %.4f, \"plus_1_lds_per_dfma_ms\": %.4f, \"plus_2_lds_per_dfma_ms\": %.4f}",
               first ? "" : ", ", warps, t0, t1, t2, l1, l2);
        first = false;
    }
    printf("]}\n");
    return 0;
}
