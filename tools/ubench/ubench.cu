// Micro-benchmarks that size the forward-group design on the box at hand (not product code; not linked into the library):
//   fp     : issue rate of FADD / FFMA / FADD2 / FMUL2 / FFMA2 with 16 warps per SM (thread-level flops per clock per SM)
//   l2     : bulk-copy (cp.async.bulk) streaming reads of a buffer of S MiB by all SMs, three 64 KiB stages per CTA
//   st     : 128-bit stores into a buffer of S MiB (L2-resident when small)
//   flag   : release/acquire hand-off latency between two CTAs through a global flag
//   launch : back-to-back launch cost of a 148-CTA kernel
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int OP> __global__ void __launch_bounds__(512, 1) fp_kernel(float *out, int iters, float a, float b, long long *cyc) {
    constexpr int NACC = 12;
    float x[NACC], y[NACC];
    unsigned long long p[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = i * 0.5f; p[i] = pk(x[i], y[i]); }
    const unsigned long long pa = pk(a, a), pb = pk(b, b);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (OP == 0) { x[i] = __fadd_rn(x[i], a); y[i] = __fadd_rn(y[i], b); }
            if (OP == 1) { x[i] = __fmaf_rn(x[i], a, b); y[i] = __fmaf_rn(y[i], b, a); }
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (OP == 5) { x[i] = __fmul_rn(x[i], a); y[i] = __fmul_rn(y[i], b); }
            if (OP == 6) { x[i] = __fmaf_rn(x[i], 0.999f, b); y[i] = __fmaf_rn(y[i], 1.001f, a); }   // immediate multiplier
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) { float lo, hi; unpk(p[i], lo, hi); s += x[i] + y[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// one thread per CTA streams chunks i = blockIdx.x, +gridDim.x, ... of 64 KiB into a ring of 3 stages (nothing consumes them;
// with STORE the stage is written back to `dst` with a bulk store: a TMA-only copy)
template <bool STORE> __global__ void __launch_bounds__(128, 1) l2_kernel(const char *src, char *dst, size_t nchunks, int reps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 3 * 65536);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 3; s++) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        size_t issued = 0, done = 0;
        const size_t mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x * (size_t)reps;
        auto chunk_of = [&](size_t j) { return (blockIdx.x + (j % (mine / reps)) * gridDim.x); };
        while (done < mine) {
            while (issued < mine && issued < done + 3) {
                const int s = issued % 3;
                if (STORE && issued >= 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                mbar_expect_tx(bars + s, 65536);
                const char *p = src + chunk_of(issued) * 65536;
                for (int k = 0; k < 4; k++) bulk_load(sm + s * 65536 + k * 16384, p + k * 16384, 16384, bars + s);
                issued++;
            }
            const int s = done % 3;
            while (!mbar_try_wait(bars + s, (done / 3) & 1)) {}
            if (STORE) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_store(dst + chunk_of(done) * 65536, sm + s * 65536, 65536);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            done++;
        }
        if (STORE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(512, 1) st_kernel(float4 *dst, size_t n4, int reps) {
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)blockIdx.x);
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void __launch_bounds__(512, 1) ld_kernel(const float4 *src, float *out, size_t n4, int reps) {
    float s = 0.f;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x * 4) {
            float4 a = __ldcg(src + i), b = {0, 0, 0, 0}, c = {0, 0, 0, 0}, d = {0, 0, 0, 0};
            const size_t st = (size_t)gridDim.x * blockDim.x;
            if (i + st < n4) b = __ldcg(src + i + st);
            if (i + 2 * st < n4) c = __ldcg(src + i + 2 * st);
            if (i + 3 * st < n4) d = __ldcg(src + i + 3 * st);
            s += a.x + b.y + c.z + d.w;
        }
    if (s == 123.456f) out[0] = s;
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// CTA 0 and CTA (gridDim.x - 1) ping-pong a counter: each bump is st.release after a 64 KiB "payload" store by 256 threads
__global__ void flag_kernel(unsigned *flag, float *payload, int rounds, unsigned long long *ns) {
    const bool a = blockIdx.x == 0, b = blockIdx.x == gridDim.x - 1;
    if (!a && !b) return;
    unsigned long long t0 = 0;
    for (int r = 0; r < rounds; r++) {
        const unsigned want = 2 * r + (a ? 0 : 1);
        if (threadIdx.x == 0) {
            if (r == 0 && a) t0 = gtime();
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v < want);
        }
        __syncthreads();
        payload[(a ? 0 : 16384) + threadIdx.x] = (float)r;  // some data the other side would read
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(flag, 1u);
    }
    if (a && threadIdx.x == 0) ns[0] = gtime() - t0;
}
__global__ void empty_kernel(int *p) { if (p && threadIdx.x == 1000) *p = 1; }

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char **argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, max clock %d MHz, L2 %d MiB\n", pr.name, sms, clk_khz / 1000, pr.l2CacheSize >> 20);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *out; CK(cudaMalloc(&out, sizeof(float) * 512 * sms * 2));
    long long *cyc; CK(cudaMalloc(&cyc, sizeof(long long) * sms * 2));
    {   // ---- fp ----
        const char *names[] = {"FADD (scalar, 2 per complex)", "FFMA (3 registers)", "FADD2 add.f32x2", "FMUL2 mul.f32x2", "FFMA2 fma.f32x2", "FMUL (scalar)", "FFMA (immediate multiplier)"};
        const int iters = 4096;
        auto run = [&](int op) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                switch (op) {
                case 0: fp_kernel<0><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 1: fp_kernel<1><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 2: fp_kernel<2><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 3: fp_kernel<3><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 4: fp_kernel<4><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 5: fp_kernel<5><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                case 6: fp_kernel<6><<<sms, 512>>>(out, iters, 1.0001f, 0.9999f, cyc); break;
                }
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            long long c0; CK(cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost));
            const double lane_ops = 512.0 * iters * 12 * 2;   // float results per CTA (2 per accumulator pair)
            printf("fp  %-30s %8.3f ms  %9lld cycles  %6.1f float results/clk/SM  (%.2f warp-instr/clk/SMSP scalar-equivalent)\n", names[op],
                   time_ms(e0, e1), c0, lane_ops / (double)c0, lane_ops / (double)c0 / 128.0);
        };
        for (int op = 0; op < 7; op++) run(op);
    }
    {   // ---- l2 / dram streaming reads with bulk copies ----
        CK(cudaFuncSetAttribute(l2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 64));
        CK(cudaFuncSetAttribute(l2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 64));
        const size_t big = (size_t)1 << 30;
        char *buf, *buf2; CK(cudaMalloc(&buf, big)); CK(cudaMalloc(&buf2, big)); CK(cudaMemset(buf, 1, big)); CK(cudaMemset(buf2, 0, big));
        for (size_t mib : {8, 16, 32, 48, 64, 96, 256, 1024}) {
            const size_t bytes = mib << 20, nchunks = bytes / 65536;
            const int reps = (int)std::max<size_t>(2, ((size_t)4 << 30) / bytes / 2);
            for (int w = 0; w < 2; w++) {
                CK(cudaEventRecord(e0));
                l2_kernel<false><<<sms, 128, 3 * 65536 + 64>>>(buf, nullptr, nchunks, reps);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            CK(cudaGetLastError());
            printf("l2  bulk-read  %5zu MiB x %4d : %8.1f GB/s\n", mib, reps, (double)bytes * reps / time_ms(e0, e1) / 1e6);
        }
        for (size_t mib : {8, 32, 64, 1024}) {
            const size_t bytes = mib << 20, nchunks = bytes / 65536;
            const int reps = (int)std::max<size_t>(2, ((size_t)2 << 30) / bytes / 2);
            for (int w = 0; w < 2; w++) {
                CK(cudaEventRecord(e0));
                l2_kernel<true><<<sms, 128, 3 * 65536 + 64>>>(buf, buf2, nchunks, reps);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            CK(cudaGetLastError());
            printf("l2  bulk-copy (load+store) %5zu MiB x %4d : %8.1f GB/s read + the same written\n", mib, reps, (double)bytes * reps / time_ms(e0, e1) / 1e6);
        }
        for (size_t mib : {8, 32, 64, 1024}) {
            const size_t bytes = mib << 20;
            const int reps = (int)std::max<size_t>(2, ((size_t)2 << 30) / bytes);
            for (int w = 0; w < 2; w++) {
                CK(cudaEventRecord(e0));
                st_kernel<<<sms, 512>>>(reinterpret_cast<float4 *>(buf2), bytes / 16, reps);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            printf("st  STG.128    %5zu MiB x %4d : %8.1f GB/s\n", mib, reps, (double)bytes * reps / time_ms(e0, e1) / 1e6);
            for (int w = 0; w < 2; w++) {
                CK(cudaEventRecord(e0));
                ld_kernel<<<sms, 512>>>(reinterpret_cast<const float4 *>(buf), out, bytes / 16, reps);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            }
            printf("ld  LDG.128 x4 %5zu MiB x %4d : %8.1f GB/s\n", mib, reps, (double)bytes * reps / time_ms(e0, e1) / 1e6);
        }
        CK(cudaFree(buf)); CK(cudaFree(buf2));
    }
    {   // ---- flag ----
        unsigned *flag; unsigned long long *ns; float *payload;
        CK(cudaMalloc(&flag, 4)); CK(cudaMalloc(&ns, 8)); CK(cudaMalloc(&payload, 4 * 32768));
        for (int grid : {2, sms}) {
            CK(cudaMemset(flag, 0, 4));
            const int rounds = 2000;
            flag_kernel<<<grid, 256>>>(flag, payload, rounds, ns);
            CK(cudaDeviceSynchronize());
            unsigned long long h; CK(cudaMemcpy(&h, ns, 8, cudaMemcpyDeviceToHost));
            printf("flag ping-pong between CTA 0 and CTA %d: %.0f ns per one-way hand-off (store payload, fence, atomicAdd, acquire-poll)\n", grid - 1, (double)h / (2.0 * rounds));
        }
    }
    {   // ---- launch ----
        for (int w = 0; w < 2; w++) {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 1000; i++) empty_kernel<<<sms, 512>>>(nullptr);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        printf("launch: %.2f us per back-to-back launch of an empty %d x 512 kernel\n", time_ms(e0, e1), sms);
    }
    return 0;
}
