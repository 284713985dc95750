// FP64 roof of one B200, measured ("must be measured by an FP64-FMA microbenchmark").
// A DFMA-only, register-resident kernel: every thread runs CHAINS independent dependent-FMA chains, so the FP64 pipe is the
// only resource in use.  Swept over resident warps per SM (4/8/12/16/32) and chains per thread (1/2/4/8); CUDA events,
// best of 5.  Prints one JSON object: TFLOP/s, warp-instructions per clock and SM at the sampled SM clock.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/fp64_peak scripts/micro/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(int iters, double a, double b, double *out) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

template <int CHAINS>
double run(int sms, int warps_per_sm, int iters, double *out, double *tflops) {
    int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm);
    int blocks_per_sm = warps_per_sm > 32 ? warps_per_sm / 32 : 1;
    dim3 grid(sms * blocks_per_sm), block(threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<CHAINS><<<grid, block>>>(iters / 8, 1.0000001, 1e-9, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<CHAINS><<<grid, block>>>(iters, 1.0000001, 1e-9, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double warp_instr = (double)sms * warps_per_sm * (double)iters * 16 * CHAINS;
    *tflops = warp_instr * 32 * 2 / (best * 1e-3) / 1e12;
    return warp_instr / (best * 1e-3);   // warp instructions per second
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024);
    const int iters = 4096;
    int wps[] = {4, 8, 12, 16, 32, 64};
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"kernel\": \"dependent DFMA chains, register resident\", \"sweep\": [", p.name, sms, clk_khz / 1e3);
    double best_tf = 0, best_ips = 0; int bw = 0, bc = 0; bool first = true;
    for (int w : wps)
        for (int chains : {1, 2, 4, 8}) {
            double tf, ips;
            switch (chains) {
                case 1: ips = run<1>(sms, w, iters, out, &tf); break;
                case 2: ips = run<2>(sms, w, iters, out, &tf); break;
                case 4: ips = run<4>(sms, w, iters, out, &tf); break;
                default: ips = run<8>(sms, w, iters, out, &tf); break;
            }
            printf("%s{\"warps_per_sm\": %d, \"chains\": %d, \"tflops\": %.3f, \"warp_instr_per_clk_sm_at_max_clock\": %.4f}", first ? "" : ", ", w, chains, tf,
                   ips / sms / (clk_khz * 1e3));
            first = false;
            if (tf > best_tf) { best_tf = tf; best_ips = ips; bw = w; bc = chains; }
        }
    printf("], \"peak\": {\"tflops\": %.3f, \"warp_instr_per_s\": %.6e, \"warp_instr_per_clk_sm_at_max_clock\": %.4f, \"warps_per_sm\": %d, \"chains\": %d}}\n", best_tf, best_ips,
           best_ips / sms / (clk_khz * 1e3), bw, bc);
    return 0;
}
