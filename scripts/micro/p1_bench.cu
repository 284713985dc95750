// Microbenchmark: how fast can one SM evaluate bending stencils / faces (elements.cuh tile forms) from shared memory into
// shared memory, as a function of resident warps and the register cap?  Not part of the product; informs the kernel design.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../eol_cloth_b200/csrc/tile_exec.cuh"
using namespace eolc;
using namespace eolc::tiles;

template <int NT, int MINB, int EDGE_WARPS>
__global__ void __launch_bounds__(NT, MINB) k(int iters, const double *__restrict__ xin, double *out) {
    extern __shared__ __align__(16) double sm[];
    double *xs = sm, *Xs = sm + 3 * 128, *scr = sm + 5 * 128;
    for (int i = threadIdx.x; i < 5 * 128; i += NT) sm[i] = xin[i];
    __syncthreads();
    FillParams P{24.75, 0.5, 0.05, 1e-5, 0, 0, -9.8, 2.5e-5};
    const int t = threadIdx.x, w = t >> 5;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t l0 = (t + it) & 63, l1 = l0 + 1, l2 = l0 + 17, l3 = l0 + 33;
        if (w < EDGE_WARPS) {
            double X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y;
            ld2(Xs + 2 * l0, X0x, X0y); ld2(Xs + 2 * l1, X1x, X1y); ld2(Xs + 2 * l2, X2x, X2y); ld2(Xs + 2 * l3, X3x, X3y);
            ParkEdge park{scr + EDGE_STRIDE * (t % 160)};
            edge_element_tile(ldx(xs, l0), ldx(xs, l1), ldx(xs, l2), ldx(xs, l3), X0x, X0y, X1x, X1y, X2x, X2y, X3x, X3y, P.beta, P.dhh, park);
        } else {
            double Xax, Xay, Xbx, Xby, Xcx, Xcy;
            ld2(Xs + 2 * l0, Xax, Xay); ld2(Xs + 2 * l1, Xbx, Xby); ld2(Xs + 2 * l2, Xcx, Xcy);
            ParkFace park{scr + EDGE_STRIDE * 160 + FACE_STRIDE * (t % 96)};
            face_element_tile(ldx(xs, l0), ldx(xs, l1), ldx(xs, l2), Xax, Xay, Xbx, Xby, Xcx, Xcy, P.mu, P.lam, P.rho, mk3(P.gx, P.gy, P.gz), P.dhh, park);
        }
        __syncthreads();
        acc += scr[(t * 7 + it) % 1000];
    }
    out[blockIdx.x * NT + t] = acc;
}

template <int NT, int MINB, int EW>
void run(const char *name, const double *xin, double *out, int sms) {
    const int iters = 2000;
    size_t smem = (5 * 128 + EDGE_STRIDE * 160 + FACE_STRIDE * 96) * 8;
    cudaFuncSetAttribute(k<NT, MINB, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<NT, MINB, EW>);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<NT, MINB, EW>, NT, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NT, MINB, EW><<<sms * occ, NT, smem>>>(10, xin, out);
    cudaEventRecord(e0);
    k<NT, MINB, EW><<<sms * occ, NT, smem>>>(iters, xin, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const int warps = NT / 32, fw = warps - EW;
    double edges = (double)sms * occ * iters * EW * 32, faces = (double)sms * occ * iters * fw * 32;
    double fp64 = edges * 546 + faces * 397;   // static FP64 instruction counts of the two element forms
    printf("%-34s regs %3d occ %d  %.3f ms  edges/s %.3e faces/s %.3e  FP64 warp-instr/clk/SM %.3f (peak 2)  err=%s\n", name, fa.numRegs, occ, ms,
           edges / ms * 1e3, faces / ms * 1e3, fp64 / 32 / (ms * 1e-3) / 1.965e9 / sms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double h[5 * 128];
    for (int i = 0; i < 128; ++i) { double X = (i % 16) * 0.1, Y = (i / 16) * 0.1; h[3 * i] = X + 0.001 * i; h[3 * i + 1] = Y; h[3 * i + 2] = 0.01 * ((i * 7) % 5); h[384 + 2 * i] = X; h[384 + 2 * i + 1] = Y; }
    double *xin, *out; cudaMalloc(&xin, sizeof h); cudaMalloc(&out, 8 * 2048 * 1024); cudaMemcpy(xin, h, sizeof h, cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    run<256, 1, 5>("256thr x1 (5 edge + 3 face warps)", xin, out, sms);
    run<256, 1, 8>("256thr x1 (8 edge warps)", xin, out, sms);
    run<128, 1, 4>("128thr x1 (4 edge warps)", xin, out, sms);
    run<384, 1, 8>("384thr x1 (8 edge + 4 face)", xin, out, sms);
    run<512, 1, 10>("512thr x1 (10 edge + 6 face)", xin, out, sms);
    run<512, 1, 16>("512thr x1 (16 edge)", xin, out, sms);
    run<768, 1, 15>("768thr x1 (15 edge + 9 face)", xin, out, sms);
    run<1024, 1, 20>("1024thr x1 (20 edge + 12 face)", xin, out, sms);
    return 0;
}
