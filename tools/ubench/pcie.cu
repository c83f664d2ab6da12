// Host <-> device copy rates of this box: pinned vs write-combined pinned source, one direction and both at once,
// 4 MiB pieces (one hop of cfg 2) and one 256 MiB piece. Build: nvcc -O2 -o tools/ubench/pcie tools/ubench/pcie.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    const size_t total = 256u << 20, piece = 4u << 20;
    void *h_def, *h_wc, *h_out, *d_a, *d_b;
    CK(cudaHostAlloc(&h_def, total, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_wc, total, cudaHostAllocWriteCombined));
    CK(cudaHostAlloc(&h_out, total, cudaHostAllocDefault));
    memset(h_def, 1, total); memset(h_wc, 1, total); memset(h_out, 0, total);
    CK(cudaMalloc(&d_a, total)); CK(cudaMalloc(&d_b, total));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto run = [&](const char *name, void *src, size_t pc, bool h2d, bool d2h) -> int {
        float best = 1e9f;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, s1));
            if (h2d) for (size_t o = 0; o < total; o += pc) CK(cudaMemcpyAsync((char *)d_a + o, (char *)src + o, pc, cudaMemcpyHostToDevice, s1));
            if (d2h) for (size_t o = 0; o < total; o += pc) CK(cudaMemcpyAsync((char *)h_out + o, (char *)d_b + o, pc, cudaMemcpyDeviceToHost, s2));
            CK(cudaStreamSynchronize(s2));
            CK(cudaEventRecord(b, s1));
            CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            if (ms < best) best = ms;
        }
        printf("%-46s %6.1f GB/s per direction\n", name, total / best / 1e6);
        return 0;
    };
    run("H2D pinned, 4 MiB pieces", h_def, piece, true, false);
    run("H2D pinned, one 256 MiB piece", h_def, total, true, false);
    run("H2D write-combined, 4 MiB pieces", h_wc, piece, true, false);
    run("D2H pinned, 4 MiB pieces", h_def, piece, false, true);
    run("H2D + D2H pinned, 4 MiB pieces", h_def, piece, true, true);
    run("H2D write-combined + D2H pinned, 4 MiB pieces", h_wc, piece, true, true);
    return 0;
}
