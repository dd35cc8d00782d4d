// PCIe ceiling for the host-buffer (e2e) step: pinned H2D, D2H and both directions at once, for the byte counts
// of one SWE 8192^2 fp32 step (3 fields each way), whole and in the slab sizes wsb_sim_step_host streams.
// build: nvcc -O2 -o profiles/tools/pcie_probe profiles/tools/pcie_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
// stand-in for one slab's step kernel: 588 one-warp CTAs that read the uploaded slab, write the output slab and
// stay resident for about `spin` clocks (a 64-row chunk sweep is latency bound at this occupancy)
__global__ void slab_kernel(const float4 *in, float4 *out, size_t n, long long spin) {
    const long long t0 = clock64();
    for (size_t i = blockIdx.x * 32 + threadIdx.x; i < n; i += (size_t)gridDim.x * 32) out[i] = in[i];
    while (clock64() - t0 < spin) {}
}
// persistent stand-in for the whole step: the CTAs of slab i spin until the upload stream has copied ready[i] = 1
// behind the slab's data, copy the slab, and count themselves done; the download stream waits on the counter
// with a stream memory operation (no events, no launch per slab)
__global__ void persistent_kernel(const float4 *in, float4 *out, size_t slab_f4, int ctas_per_slab,
                                  const volatile unsigned *ready, unsigned *done) {
    const int slab = blockIdx.x / ctas_per_slab, part = blockIdx.x % ctas_per_slab;
    if (threadIdx.x == 0)
        while (ready[slab] == 0) __nanosleep(200);
    __syncwarp();
    __threadfence();
    const float4 *src = in + slab * slab_f4;
    float4 *dst = out + slab * slab_f4;
    for (size_t i = (size_t)part * 32 + threadIdx.x; i < slab_f4; i += (size_t)ctas_per_slab * 32) dst[i] = src[i];
    __threadfence();
    __syncwarp();
    if (threadIdx.x == 0) atomicAdd(&done[slab], 1u);
}
typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
int main(int argc, char **argv) {
    const size_t bytes = 3ull * 8192 * 8192 * 4;
    const int nslab = argc > 1 ? atoi(argv[1]) : 32;
    char *hin, *hout, *din, *dout;
    CK(cudaHostAlloc(&hin, bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hout, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&din, bytes));
    CK(cudaMalloc(&dout, bytes));
    for (size_t i = 0; i < bytes; i += 4096) hin[i] = 1, hout[i] = 2;
    cudaStream_t a, b;
    CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, eb;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&eb));
    const size_t slab = bytes / nslab;
    cudaStream_t c;
    CK(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
    static cudaEvent_t up[1024], done[1024];
    for (int i = 0; i < nslab && i < 1024; ++i) {
        CK(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
        unsigned *ready, *done_cnt, *ones;
    CK(cudaMalloc(&ready, 4096)); CK(cudaMalloc(&done_cnt, 4096));
    CK(cudaHostAlloc(&ones, 4096, cudaHostAllocDefault));
    for (int i = 0; i < 1024; ++i) ones[i] = 1;
    WaitValueFn wait_value = nullptr;
    { void *p = nullptr; cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) wait_value = (WaitValueFn)p; }
    for (int mode = 0; mode < 9; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, a));
            CK(cudaStreamWaitEvent(b, e0, 0));
            if (mode == 0 || mode == 2) CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, a));
            if (mode == 1 || mode == 2) CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, b));
            if (mode == 3)
                for (int i = 0; i < nslab; ++i) {
                    CK(cudaMemcpyAsync(din + i * slab, hin + i * slab, slab, cudaMemcpyHostToDevice, a));
                    CK(cudaMemcpyAsync(hout + i * slab, dout + i * slab, slab, cudaMemcpyDeviceToHost, b));
                }
            if (mode == 8) {
                if (!wait_value) { printf("no cuStreamWaitValue32\n"); return 1; }
                const int cps = 588;
                CK(cudaMemsetAsync(ready, 0, 4096, c));
                CK(cudaMemsetAsync(done_cnt, 0, 4096, c));
                CK(cudaEventRecord(up[0], c));
                CK(cudaStreamWaitEvent(a, up[0], 0));
                CK(cudaStreamWaitEvent(b, up[0], 0));
                persistent_kernel<<<nslab * cps, 32, 0, c>>>((const float4 *)din, (float4 *)dout, slab / 16, cps, ready, done_cnt);
                for (int i = 0; i < nslab; ++i) {
                    for (int k = 0; k < 3; ++k)
                        CK(cudaMemcpyAsync(din + i * slab + k * (slab / 3), hin + i * slab + k * (slab / 3), slab / 3, cudaMemcpyHostToDevice, a));
                    CK(cudaMemcpyAsync(ready + i, ones + i, 4, cudaMemcpyHostToDevice, a));
                    if (wait_value((CUstream)b, (CUdeviceptr)(done_cnt + i), cps, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) { printf("wait_value failed\n"); return 1; }
                    for (int k = 0; k < 3; ++k)
                        CK(cudaMemcpyAsync(hout + i * slab + k * (slab / 3), dout + i * slab + k * (slab / 3), slab / 3, cudaMemcpyDeviceToHost, b));
                }
            }
            if (mode >= 4 && mode < 8)  // the dependency chain of wsb_sim_step_host: upload i -> (kernel i) -> download i
                for (int i = 0; i < nslab; ++i) {
                    for (int k = 0; k < 3; ++k)
                        CK(cudaMemcpyAsync(din + i * slab + k * (slab / 3), hin + i * slab + k * (slab / 3), slab / 3, cudaMemcpyHostToDevice, a));
                    CK(cudaEventRecord(up[i], a));
                    if (mode >= 5) {
                        CK(cudaStreamWaitEvent(c, up[i], 0));
                        if (mode == 5) CK(cudaMemsetAsync(dout + i * slab, 0, 4096, c));
                        else slab_kernel<<<588, 32, 0, c>>>((const float4 *)(din + i * slab), (float4 *)(dout + i * slab), slab / 16, mode == 6 ? 0 : 100000);
                        CK(cudaEventRecord(done[i], c));
                        CK(cudaStreamWaitEvent(b, done[i], 0));
                    } else {
                        CK(cudaStreamWaitEvent(b, up[i], 0));
                    }
                    for (int k = 0; k < 3; ++k)
                        CK(cudaMemcpyAsync(hout + i * slab + k * (slab / 3), dout + i * slab + k * (slab / 3), slab / 3, cudaMemcpyDeviceToHost, b));
                }
            CK(cudaEventRecord(eb, b));
            CK(cudaStreamWaitEvent(a, eb, 0));
            CK(cudaEventRecord(e1, a));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const char *names[9] = {"h2d", "d2h", "h2d+d2h concurrent", "h2d+d2h concurrent, slabs", "slabs x3 copies, up->down events",
                                "slabs x3, up->memset->down events", "slabs x3, up->copy kernel->down", "slabs x3, up->50us kernel->down", "slabs x3, flags + persistent kernel + wait-value"};
        const double moved = (mode >= 2 ? 2.0 : 1.0) * bytes;
        printf("%-48s %8.3f ms  %7.2f GB/s total (%zu MB per direction)\n", names[mode], best, moved / best * 1e-6, bytes >> 20);
    }
    return 0;
}
