// cdp_probe.cu -- feasibility probe (development aid, not product code): a kernel whose last CTA tail-launches its own
// next iteration with by-value parameters (CUDA dynamic parallelism, cudaStreamTailLaunch), against the host-driven
// loop (launch, poll a mapped flag, launch).  Prints the per-iteration cost of both.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true cdp_probe.cu -o cdp_probe -lcudadevrt
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>

struct P {
    double c[80];
    int iter, n_iter, work;
    unsigned *ticket;
    double *buf;
    volatile unsigned long long *done;  // mapped host memory
    int self;  // 1: relaunch from the device
};

__global__ void __launch_bounds__(128) step(P p) {
    double v = p.buf[blockIdx.x * 128 + threadIdx.x];
    for (int i = 0; i < p.work; i++)
#pragma unroll
        for (int j = 0; j < 80; j++) v = fma(v, 0.999999, p.c[j]);
    p.buf[blockIdx.x * 128 + threadIdx.x] = v * 1e-30;
    __shared__ int last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *p.ticket = 0u;
        if (p.self && p.iter + 1 < p.n_iter) {
            P q = p;
            q.iter++;
            q.c[q.iter % 80] += 1e-9;
            step<<<gridDim.x, 128, 0, cudaStreamTailLaunch>>>(q);
        } else {
            __threadfence_system();
            *p.done = (unsigned long long)(p.iter + 1);
        }
    }
}

int main(int argc, char **argv) {
    int n_iter = argc > 1 ? atoi(argv[1]) : 2000;
    int grid = 296;
    unsigned *ticket; double *buf; unsigned long long *done_h, *done_d;
    cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
    cudaMalloc(&buf, grid * 128 * 8); cudaMemset(buf, 0, grid * 128 * 8);
    cudaHostAlloc(&done_h, 8, cudaHostAllocMapped); *done_h = 0;
    cudaHostGetDevicePointer(&done_d, done_h, 0);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int work : {0, 20, 400}) {
        P p{}; for (int j = 0; j < 80; j++) p.c[j] = 1e-3 * j;
        p.n_iter = n_iter; p.work = work; p.ticket = ticket; p.buf = buf; p.done = done_d;
        // (a) device-side tail launch chain
        for (int rep = 0; rep < 2; rep++) {
            *done_h = 0; p.iter = 0; p.self = 1;
            cudaEventRecord(e0, st);
            step<<<grid, 128, 0, st>>>(p);
            cudaEventRecord(e1, st);
            cudaError_t e = cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("work %3d tail-launch chain: %s done=%llu  %.3f us/iter (events span the chain: %s)\n", work,
                   cudaGetErrorString(e), *done_h, ms * 1e3 / n_iter, *done_h == (unsigned long long)n_iter ? "yes" : "NO");
        }
        // (b) host loop: launch, poll mapped flag, launch
        for (int rep = 0; rep < 2; rep++) {
            p.self = 0;
            cudaEventRecord(e0, st);
            auto t0 = std::chrono::steady_clock::now();
            for (int it = 0; it < n_iter; it++) {
                *done_h = 0; p.iter = it;
                step<<<grid, 128, 0, st>>>(p);
                while (*(volatile unsigned long long *)done_h != (unsigned long long)(it + 1)) {}
            }
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            auto t1 = std::chrono::steady_clock::now();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("work %3d host poll loop   : %.3f us/iter (events) %.3f us/iter (wall)\n", work, ms * 1e3 / n_iter,
                   std::chrono::duration<double, std::micro>(t1 - t0).count() / n_iter);
        }
        // (c) plain back-to-back launches (no dependence on results): the floor
        p.self = 0;
        cudaEventRecord(e0, st);
        for (int it = 0; it < n_iter; it++) { p.iter = it; step<<<grid, 128, 0, st>>>(p); }
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("work %3d back-to-back     : %.3f us/iter\n", work, ms * 1e3 / n_iter);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
