// oc_hjb_final.h -- host side of the fused step's in-launch final reduction (oc_hjb_fused.cuh): ticket counters on the
// device, result slots in mapped page-locked host memory, and the polling hand-over.  Shared by the single-GPU /
// batched solves (oc_hjb.cu) and the NVLink peer-memory row-band solve (oc_hjb_dist.cu).
#pragma once
#include <atomic>

#include "oc_common.h"
#include "oc_hjb_fused.cuh"

namespace ocfinal {

// device ticket counters + mapped host result slots (slot b: value at [2b], sequence number at [2b+1])
inline int ensure(oc_ctx *ctx, int slots) {
    if (ctx->fr_slots >= slots) return OC_OK;
    if (ctx->fr_ticket) cudaFree(ctx->fr_ticket);
    if (ctx->fr_result) cudaFreeHost(ctx->fr_result);
    ctx->fr_ticket = nullptr; ctx->fr_result = nullptr; ctx->fr_slots = 0;
    OC_CUDA(cudaMalloc(&ctx->fr_ticket, sizeof(unsigned) * slots));
    OC_CUDA(cudaMemset(ctx->fr_ticket, 0, sizeof(unsigned) * slots));
    OC_CUDA(cudaHostAlloc(&ctx->fr_result, sizeof(double) * 2 * slots, cudaHostAllocMapped));
    memset(ctx->fr_result, 0, sizeof(double) * 2 * slots);
    OC_CUDA(cudaDeviceSynchronize());
    ctx->fr_slots = slots;
    return OC_OK;
}
inline void set_args(oc_ctx *ctx, int slot, unsigned long long seq, fused::Args &fa) {
    double *dev = nullptr;
    cudaHostGetDevicePointer(&dev, ctx->fr_result, 0);
    fa.ticket = ctx->fr_ticket + slot;
    fa.result = dev + 2 * slot;
    fa.result_seq = reinterpret_cast<unsigned long long *>(dev + 2 * slot + 1);
    fa.seq = seq;
}
inline bool ready(oc_ctx *ctx, int slot, unsigned long long seq, double *out) {
    const volatile unsigned long long *f = reinterpret_cast<const volatile unsigned long long *>(ctx->fr_result + 2 * slot + 1);
    if (*f != seq) return false;
    std::atomic_thread_fence(std::memory_order_acquire);
    *out = *reinterpret_cast<const volatile double *>(ctx->fr_result + 2 * slot);
    return true;
}
// spin on the sequence word the last CTA writes (a few microseconds after the kernel's last store, instead of the copy +
// stream synchronisation round trip); the stream is queried from time to time so that a failed launch cannot hang the host
inline int wait(oc_ctx *ctx, int slot, unsigned long long seq, cudaStream_t st, double *out) {
    const volatile unsigned long long *f = reinterpret_cast<const volatile unsigned long long *>(ctx->fr_result + 2 * slot + 1);
    for (unsigned long long spin = 1;; spin++) {
        if (ready(ctx, slot, seq, out)) return OC_OK;
        if (*f == (seq | (1ull << 63))) {  // peer-memory mode: another rank never delivered its error sums
            oc::set_error("row-band solve: a peer rank did not arrive within the time-out");
            return OC_ERR_NCCL;
        }
        if ((spin & 0x3ffff) == 0) {
            cudaError_t e = cudaStreamQuery(st);
            if (e == cudaErrorNotReady) continue;
            if (e == cudaSuccess && ready(ctx, slot, seq, out)) return OC_OK;
            oc::set_error("fused RK45 step did not deliver its error sum: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return OC_ERR_CUDA;
        }
    }
}

}  // namespace ocfinal
