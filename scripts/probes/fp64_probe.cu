// fp64_probe.cu -- development aid: DFMA dependent-issue latency and per-SM throughput as a function of resident warps
// and instruction-level parallelism (B200).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_probe.cu -o fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int ILP>
void run(int warps_per_sm, int n_sm, double *out) {
    const int iters = 2000;
    const int threads = 32 * warps_per_sm;  // one CTA per SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<n_sm, threads>>>(out, 10, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<ILP><<<n_sm, threads>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const double inst_per_warp = (double)iters * 16 * ILP;
    printf("ILP %d warps/SM %2d (%.1f/SMSP): %.2f cycles per DFMA per warp, %.3f warp-DFMA/clk/SM, %.2f TFLOP/s\n", ILP, warps_per_sm,
           warps_per_sm / 4.0, cyc / inst_per_warp, inst_per_warp * warps_per_sm / cyc,
           2.0 * inst_per_warp * 32 * warps_per_sm * n_sm / (ms * 1e-3) / 1e12);
}

int main() {
    int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, 8 * 1024 * 1024);
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, n_sm, out); run<2>(w, n_sm, out); run<4>(w, n_sm, out); run<8>(w, n_sm, out); }
    return 0;
}
