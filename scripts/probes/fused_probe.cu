// fused_probe.cu -- development aid: times fused::hjb_fused_kernel<NE> alone on a 16384 x 2048 band with fixed step
// coefficients (no RK45 controller around it), so that kernel variants and ablations (-D flags) can be compared in one
// GPU call.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../optimal_crowds_b200/csrc
//                        [-D...] fused_probe.cu -o fused_probe_<tag>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "oc_hjb_fused.cuh"
#include "oc_rk45.h"

#ifndef NE_PROBE
#define NE_PROBE 3
#endif

int main(int argc, char **argv) {
    const int Nx = argc > 1 ? atoi(argv[1]) : 16384, Ny = argc > 2 ? atoi(argv[2]) : 2048;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    const size_t n = (size_t)Nx * Ny;
    std::vector<double> hy(n), hk(n), hc(n);
    for (size_t i = 0; i < n; i++) {
        const int x = (int)(i % Nx), yv = (int)(i / Nx);
        hy[i] = 1.0 + 1e-3 * std::sin(0.01 * x) * std::cos(0.013 * yv);
        hk[i] = 1e-2 * std::cos(0.02 * x + 0.01 * yv);
        const bool wall = (x % 160 < 20 && yv % 160 < 20) || x == 0 || yv == 0 || x == Nx - 1 || yv == Ny - 1;
        hc[i] = wall ? NAN : 0.0;
    }
    double *y, *k1, *coef, *ynew, *k7, *phi, *partial;
    cudaMalloc(&y, n * 8); cudaMalloc(&k1, n * 8); cudaMalloc(&coef, n * 8); cudaMalloc(&ynew, n * 8); cudaMalloc(&k7, n * 8);
    cudaMalloc(&phi, n * 8 * NE_PROBE + 8); cudaMalloc(&partial, 1 << 20);
    cudaMemcpy(y, hy.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(k1, hk.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(coef, hc.data(), n * 8, cudaMemcpyHostToDevice);
    unsigned *ticket; cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
    double *res_h, *res_d; cudaHostAlloc(&res_h, 16, cudaHostAllocMapped); cudaHostGetDevicePointer(&res_d, res_h, 0);
    fused::Args a{};
    const double h = -0.05;
    a.y = y; a.k1 = k1; a.coef = coef; a.ynew = ynew; a.k7 = k7; a.partial = partial;
    a.ha21 = h * rk45::A[1][0];
    for (int j = 0; j < 2; j++) a.ha3[j] = h * rk45::A[2][j];
    for (int j = 0; j < 3; j++) a.ha4[j] = h * rk45::A[3][j];
    for (int j = 0; j < 4; j++) a.ha5[j] = h * rk45::A[4][j];
    for (int j = 0; j < 5; j++) a.ha6[j] = h * rk45::A[5][j];
    for (int j = 0; j < 6; j++) a.hb[j] = h * rk45::B[j];
    for (int j = 0; j < 7; j++) a.he[j] = h * rk45::E[j];
    for (int e = 0; e < NE_PROBE; e++) {
        const double x = (e + 1.0) / (NE_PROBE + 1.0), pw[4] = {x, x * x, x * x * x, x * x * x * x};
        for (int j = 0; j < 7; j++) { double s = 0; for (int q = 0; q < 4; q++) s += rk45::P[j][q] * pw[q]; a.w[e][j] = h * s; }
        a.phi[e] = phi + (size_t)e * n;
    }
    a.A = -0.5 * 0.04 / (0.05 * 0.05); a.rtol = 1e-3; a.atol = 1e-6;
    a.Ny = Ny; a.Nx = Nx; a.row_base = 0; a.own0 = 0; a.own1 = Ny; a.phi_row_base = 0;
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    a.RC = getenv("OC_RC") ? atoi(getenv("OC_RC")) : fused::plan_chunk_rows(Nx, Ny, n_sm, 1, 0);
    a.ticket = ticket; a.result = res_d; a.result_seq = (unsigned long long *)(res_d + 1);
    fused::set_tensor_maps(a, Ny);
    printf("tma=%d ", a.tma);
    dim3 grid((Nx + fused::VX - 1) / fused::VX, (Ny + a.RC - 1) / a.RC);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, tot = 0;
    for (int it = 0; it < reps + 3; it++) {
        a.seq = it + 1;
        cudaEventRecord(e0);
        cudaError_t le = fused::launch<false>(NE_PROBE, a, grid, 0);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess || le != cudaSuccess) { printf("CUDA error: %s / %s\n", cudaGetErrorString(le), cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 3) { tot += ms; best = ms < best ? ms : best; }
    }
    const double bytes = 8.0 * n * (5 + NE_PROBE);
    printf("%-28s NE=%d grid %dx%d RC=%d smem=%zu : avg %.1f us  best %.1f us  %.0f GB/s (%.3f of 6551.7)  sum=%.6e\n",
           argc > 4 ? argv[4] : "probe", NE_PROBE, grid.x, grid.y, a.RC, sizeof(fused::Smem), tot / reps * 1e3, best * 1e3,
           bytes / (tot / reps) / 1e6, bytes / (tot / reps) / 1e6 / 6551.7, res_h[0]);
    return 0;
}
