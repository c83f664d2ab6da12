// TMA-fed variants of the two FFT passes for the 2^20-point complex transform (the headline configs:
// 2^20 c2c and 2^21 r2c). Same mathematics and data layout as fft_fwd.cuh; what changes is how the data
// reaches the butterflies and where the twiddles live:
//   * two CTAs of eight warps per SM; thread 0 of a CTA streams the CTA's NEXT tile into shared memory with TMA (pass 1: cp.async.bulk.tensor 2-D boxes of the strided column tile;
//     pass 2: cp.async.bulk 1-D copies of contiguous rows) as soon as the stage buffer is free again, i.e.
//     while the consumers are still in the second register DFT, the twiddle multiply and the global stores
//     of the current tile; a "full" mbarrier (expect_tx / complete_tx) tells the consumers when it has landed;
//   * the stage buffer doubles as the shared-memory exchange buffer between the two register DFTs; the |X|^2
//     tile of the fused waterfall epilogue has its own 36 KiB so the stage is released before the epilogue;
//   * pass 1 keeps ONE column tile per CTA across the frames of a batch: its slice of the Hann window is
//     fetched once, the intra-pass twiddles sit in shared memory, and the inter-pass twiddles
//     W_M^(n2*(q+32s)) = G_a * B^b (s = 4a+b) are held as 8+3 register values per thread.
#pragma once
#include <cuda.h>
#include <cstdint>
#include "fft_fwd.cuh"

namespace b200 {

constexpr int kTmaT = 8;            // columns (rows) per tile
constexpr int kTmaThreads = 256;    // 8 warps: T * 32 threads; thread 0 also drives TMA
constexpr int kS = 1024;            // sub-transform length of both passes

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {  // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A transfer that never lands would be a bug in this file (or a launch that lost the co-residency it relies on); rather
// than hang the GPU - or trap, which would poison the context of every engine in the process - a wait that has lasted
// more than five seconds of wall time (globaltimer, checked every 4096 failed attempts) latches an error word and the
// thread leaves the kernel; the threads that depended on it time out the same way. The word lives in page-locked host
// memory mapped into the device (no per-launch copy: a 4-byte device-to-host copy behind every launch group costs
// ~10 us of stream time, 0.17 us per frame); the next synchronising call of the host reports it.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ int *g_wait_err;  // mapped host word: 0, or which kind of bounded wait expired on this device (1 mbarrier, 2 frame counter)
__device__ __noinline__ void wait_timed_out(int code) {
    int *w = g_wait_err;
    if (w) {
        *reinterpret_cast<volatile int *>(w) = code;
        __threadfence_system();
    }
    asm volatile("exit;");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1; !mbar_try_wait(bar, parity); spins++) {
        if ((spins & 4095u) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 5000000000ull) wait_timed_out(1);
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// orders what this thread has observed of global memory (through an acquire) and its shared-memory accesses before its
// subsequent TMA (async proxy) operations
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { __syncthreads(); }

struct TmaSmem {
    static constexpr int kRow1 = 32 * kTmaT + 8;                        // pass-1 exchange row pitch (float2)
    static constexpr int kRow2 = 32 * kTmaT + 1;                        // pass-2 exchange row pitch (float2)
    static constexpr size_t kStage = sizeof(float2) * 32 * kRow1;       // >= 64 KiB tile, >= either exchange layout
    static constexpr size_t kTwA = sizeof(float2) * 32 * 32;            // intra-pass twiddles W_1024^(r q)
    static constexpr size_t kWindowC = sizeof(float) * kS * kTmaT;      // 32 KiB Hann slice, c2c (one weight per IQ pair)
    static constexpr size_t kWindowR = 2 * kWindowC;                    // 64 KiB, r2c (one weight per real sample)
    static constexpr int kPP = kTmaT + 1;                               // power tile pitch (floats): odd, conflict-free
    static constexpr size_t kPower = sizeof(float) * kS * kPP;          // 36 KiB |X|^2 tile of the fused waterfall epilogue
    static constexpr size_t kBars = 64;
    static constexpr size_t kPass1C = kStage + kTwA + kWindowC + kBars;
    static constexpr size_t kWinTab = sizeof(float2) * kS;             // r2c: table the Hann weights are built from on the fly
    static constexpr size_t kPass1R = kStage + kTwA + kWinTab + kBars;
    static constexpr size_t kPass2 = kStage + kTwA + kPower + kBars;
};
static_assert(TmaSmem::kStage >= sizeof(float2) * kS * kTmaT, "stage must hold a tile");
static_assert(TmaSmem::kStage >= sizeof(float2) * 32 * TmaSmem::kRow2, "stage must hold the pass-2 exchange");

// ------------------------------------------------------------------------------------------------
// pass 1: grid = (N2/T) * nsplit CTAs. CTA (tile, part) handles the frames f == part (mod nsplit) of column tile
// `tile`; block = 256 consumers + 1 producer warp.
//   ring_map  : hop ring as a 2-D tensor of N2 elements per row, nhops*N1/2 rows, box {T elements, 256 rows}; an element (an
//               IQ sample, or two consecutive real samples) is EB bytes: 8 = float pair (mapped as 2 uint32), 4 = 16-bit
//               pair (uint32), 2 = 8-bit pair (uint16) - raw ADC samples are converted here exactly as
//               SampleConverter<T>::read does (src/samplereader.cpp:29-40,59-66; load_sample in fft_fwd.cuh)
//   window_map: Hann window as {N2 (c2c) | 2*N2 (r2c), N1}, box {T | 2*T, 256}
// ------------------------------------------------------------------------------------------------
template <bool REAL, int EB = 8>
__global__ void __launch_bounds__(kTmaThreads, 2)
    fft_pass1_tma_kernel(const FwdParams p, const __grid_constant__ CUtensorMap ring_map,
                         const __grid_constant__ CUtensorMap window_map, int nframes, int order, int nsplit) {
    constexpr int T = kTmaT, RA = 32, RB = 32, N1 = kS, N2 = kS;
    constexpr int SHIFT = REAL ? 0 : 1;
    constexpr int ROW = TmaSmem::kRow1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    float2 *twA = reinterpret_cast<float2 *>(smem_raw + TmaSmem::kStage);
    float *win = reinterpret_cast<float *>(smem_raw + TmaSmem::kStage + TmaSmem::kTwA);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + TmaSmem::kStage + TmaSmem::kTwA +
                                                  (REAL ? TmaSmem::kWinTab : TmaSmem::kWindowC));
    uint64_t *full = bars, *wbar = bars + 1;

    // Work items (column tile, frame) of this CTA, item j = 0 .. nitems-1:
    //   order 0: CTA (tile, part) keeps ONE column tile (window slice and twiddles loaded once) and takes the frames
    //            f == part (mod nsplit); all CTAs sweep the batch frame by frame together, so the hop reads and the Y
    //            writes of neighbouring column tiles coalesce in L2 / DRAM pages
    //   order 1: tile-major list cut into gridDim.x equal contiguous chunks (balanced, but the CTAs are spread over
    //            all frames of the batch at any instant: measured slower, DRAM page locality is lost)
    //   order 2: frame-major list cut into equal contiguous chunks (balanced and frame-synchronous; the window slice
    //            and the twiddles change with every item)
    const int tid = threadIdx.x;
    constexpr int NT = N2 / T;
    int begin = 0, nitems = 0;
    if (order == 0) {
        begin = (int)blockIdx.x;  // tile * nsplit + part
        const int part = (int)blockIdx.x % nsplit;
        nitems = part < nframes ? (nframes - part + nsplit - 1) / nsplit : 0;
    } else {
        const int total = NT * nframes;
        const int per = total / (int)gridDim.x, rem = total % (int)gridDim.x;
        begin = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
        nitems = per + ((int)blockIdx.x < rem ? 1 : 0);
    }
    if (nitems <= 0) return;
    auto item_of = [&](int j, int &tile, int &f) {
        if (order == 0) {
            tile = begin / nsplit;
            f = begin % nsplit + j * nsplit;
        } else if (order == 1) {
            tile = (begin + j) / nframes;
            f = (begin + j) - tile * nframes;
        } else {
            f = (begin + j) / NT;
            tile = (begin + j) - f * NT;
        }
    };
    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(wbar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    for (int i = tid; i < 32 * 32; i += kTmaThreads) {
        twA[i] = p.twA1[i];
        if constexpr (REAL) reinterpret_cast<float2 *>(win)[i] = p.winT[i];  // (h cos, h sin)(2 pi row / 1024), h = 1/2
    }
    __syncthreads();

    // thread 0 drives TMA: the window slice of a column tile when it changes (c2c), and one item ahead of the consumers
    auto issue_item = [&](int tile, int f) {
        mbar_expect_tx(full, EB * N1 * T);
        const int hopA = (p.hop0 + f) % p.nhops, hopB = (p.hop0 + f + 1) % p.nhops;
        // rows 0..511 of the frame come from the older hop, 512..1023 from the newer one
        constexpr int kBox = EB * 256 * T;                  // bytes of one box of 256 rows
        const int x0 = tile * T * (EB == 8 ? 2 : 1);        // (a float pair is two elements of the uint32 map)
        tma_load_2d(smem_raw + 0 * kBox, &ring_map, x0, hopA * (N1 / 2), full);
        tma_load_2d(smem_raw + 1 * kBox, &ring_map, x0, hopA * (N1 / 2) + 256, full);
        tma_load_2d(smem_raw + 2 * kBox, &ring_map, x0, hopB * (N1 / 2), full);
        tma_load_2d(smem_raw + 3 * kBox, &ring_map, x0, hopB * (N1 / 2) + 256, full);
    };
    auto issue_window = [&](int tile) {
        if constexpr (!REAL) {
            constexpr uint32_t kWinBytes = sizeof(float) * N1 * T;
            mbar_expect_tx(wbar, kWinBytes);
            for (int b = 0; b < 4; b++)
                tma_load_2d(reinterpret_cast<unsigned char *>(win) + b * (kWinBytes / 4), &window_map, tile * T, b * 256, wbar);
        }
    };
    if (tid == 0) {
        int tile0, f0;
        item_of(0, tile0, f0);
        issue_window(tile0);
        issue_item(tile0, f0);
    }

    // ===== consumers =====
    const int c = tid % T;
    const int r = tid / T;
    const int q = r;
    const size_t M = (size_t)N1 * N2;
    // inter-pass twiddles of this thread, W_M^(n2*k1) for k1 = q + 32 s with s = 4a + b:
    //   G[a] = W_M^(n2*(q + 128 a)),  B[b-1] = W_M^(32*n2*b)   ->   tw(s) = G[a] * B[b-1]  (b = 0: G[a])
    // (the IQ k1 = 0 row, q = 0 and s = 0, carries the one-slot rotation W_M^(N1*n2) instead)
    float2 G[8], B[3], rot0;
    // r2c: a complex element holds the real samples 2 idx and 2 idx + 1, whose Hann weights 1/2 - 1/2 cos(A_row + B_b) are
    // built on the fly: A_row from the shared table, (cos, sin) B_b = W_size^(2 n2 + b) per thread (angle addition)
    float cB0 = 0.f, sB0 = 0.f, cB1 = 0.f, sB1 = 0.f;
    auto tw_lookup = [&](unsigned e) { return cmul(__ldg(p.TL + (e & 1023u)), __ldg(p.TH + ((e >> 10) & 1023u))); };
    int cur_tile = -1, wphase = 0;
    for (int it_local = 0; it_local < nitems; it_local++) {
        int tile, f;
        item_of(it_local, tile, f);
        const int n2 = tile * T + c;
        if (tile != cur_tile) {  // (CTA-uniform) new column tile: its twiddles and its slice of the Hann window
            cur_tile = tile;
#pragma unroll
            for (int a = 0; a < 8; a++) G[a] = tw_lookup((unsigned)n2 * (unsigned)(q + 128 * a));
#pragma unroll
            for (int b = 1; b < 4; b++) B[b - 1] = tw_lookup((unsigned)n2 * 32u * (unsigned)b);
            rot0 = tw_lookup((unsigned)N1 * (unsigned)n2);
            if constexpr (REAL) {
                const unsigned e0 = 2u * (unsigned)n2, e1 = e0 + 1u;
                const float2 w0 = cmul(__ldg(p.TLr + (e0 & 1023u)), __ldg(p.THr + (e0 >> 10)));
                const float2 w1 = cmul(__ldg(p.TLr + (e1 & 1023u)), __ldg(p.THr + (e1 >> 10)));
                cB0 = w0.x;
                sB0 = -w0.y;
                cB1 = w1.x;
                sB1 = -w1.y;
            } else {
                mbar_wait(wbar, wphase & 1);
                wphase++;
            }
        }
        mbar_wait(full, it_local & 1);
        float2 v[RA];
        // raw ADC samples: (x ^ topbit) as signed, / 2^(bits-1) (an exact multiply)
        const unsigned flip = (p.in_format == FMT_U16) ? 0x8000u : (p.in_format == FMT_U8) ? 0x80u : 0u;
#pragma unroll
        for (int j = 0; j < RA; j++) {
            float2 x;
            if constexpr (EB == 8) {
                x = sm[(r + RB * j) * T + c];
            } else if constexpr (EB == 4) {
                const unsigned w = reinterpret_cast<const unsigned *>(smem_raw)[(r + RB * j) * T + c];
                x.x = (float)(short)((w & 0xFFFFu) ^ flip) * (1.0f / 32768.0f);
                x.y = (float)(short)((w >> 16) ^ flip) * (1.0f / 32768.0f);
            } else {
                const unsigned w = reinterpret_cast<const unsigned short *>(smem_raw)[(r + RB * j) * T + c];
                x.x = (float)(signed char)((w & 0xFFu) ^ flip) * (1.0f / 128.0f);
                x.y = (float)(signed char)((w >> 8) ^ flip) * (1.0f / 128.0f);
            }
            if constexpr (REAL) {
                const float2 tab = reinterpret_cast<const float2 *>(win)[r + RB * j];
                x.x *= fmaf(tab.y, sB0, fmaf(-tab.x, cB0, 0.5f));
                x.y *= fmaf(tab.y, sB1, fmaf(-tab.x, cB1, 0.5f));
            } else {
                const float w = win[(r + RB * j) * T + c];
                x.x *= w;
                x.y *= w;
            }
            v[j] = x;
        }
        consumer_sync();  // the raw tile is in registers: the stage becomes the exchange buffer
        RegDft<RA>::run(v);
        sm[r * ROW + c] = v[0];
#pragma unroll
        for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], twA[qq * RB + r]);
        consumer_sync();
        float2 u[RB];
#pragma unroll
        for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
        consumer_sync();  // exchange consumed: the next item streams in while this one is finished
        if (tid == 0 && it_local + 1 < nitems) {
            fence_proxy_async();  // generic-proxy accesses to the stage / window are ordered before the TMA writes
            int ntile, nf;
            item_of(it_local + 1, ntile, nf);
            if (ntile != tile) issue_window(ntile);  // every thread finished reading the old slice before the first sync
            issue_item(ntile, nf);
        }
        RegDft<RB>::run(u);
        float2 *Y = p.Y + (size_t)f * M + n2;
#pragma unroll
        for (int s = 0; s < RB; s++) {
            const int k1 = q + RA * s;
            int u1 = k1 - SHIFT;
            if (u1 < 0) u1 += N1;
            float2 tw = (s & 3) ? cmul(G[s >> 2], B[(s & 3) - 1]) : G[s >> 2];
            if (SHIFT && s == 0 && q == 0) tw = rot0;  // the IQ k1 = 0 row: one-slot rotation
            Y[(size_t)u1 * N2] = cmul(u[s], tw);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: persistent grid (2 CTAs per SM); tile i = (frame, row tile) goes to CTA i % gridDim.x.
// ------------------------------------------------------------------------------------------------
template <int FUSE>
__global__ void __launch_bounds__(kTmaThreads, 2) fft_pass2_tma_kernel(const FwdParams p, int nframes) {
    constexpr int T = kTmaT, RA = 32, RB = 32, N1 = kS, N2 = kS;
    constexpr int ROW = TmaSmem::kRow2;
    constexpr int PP = TmaSmem::kPP;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    float2 *twA = reinterpret_cast<float2 *>(smem_raw + TmaSmem::kStage);
    float *pt = reinterpret_cast<float *>(smem_raw + TmaSmem::kStage + TmaSmem::kTwA);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + TmaSmem::kStage + TmaSmem::kTwA + TmaSmem::kPower);

    const int tid = threadIdx.x;
    const int tiles_per_frame = N1 / T;
    const int total = tiles_per_frame * nframes;
    const size_t M = (size_t)N1 * N2;
    if (tid == 0) {
        mbar_init(full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    for (int i = tid; i < 32 * 32; i += kTmaThreads) twA[i] = p.twA2[i];
    __syncthreads();
    auto issue_tile = [&](int i) {  // T consecutive rows of Y are one contiguous 64 KiB block
        const int frame = i / tiles_per_frame, tile = i - frame * tiles_per_frame;
        mbar_expect_tx(full, sizeof(float2) * N2 * T);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.Y + (size_t)frame * M + (size_t)tile * T * N2);
        for (int k = 0; k < 4; k++) bulk_load_1d(smem_raw + k * 16384, src + k * 16384, 16384, full);
    };
    if (tid == 0 && (int)blockIdx.x < total) issue_tile(blockIdx.x);

    int it = 0;
    for (int i = blockIdx.x; i < total; i += gridDim.x, it++) {
        const int frame = i / tiles_per_frame, tile = i - frame * tiles_per_frame;
        mbar_wait(full, it & 1);
        {   // stage A: lanes along n2
            const int r = tid % 32;
            const int c = tid / 32;
            float2 v[RA];
#pragma unroll
            for (int j = 0; j < RA; j++) v[j] = sm[c * N2 + r + RB * j];
            consumer_sync();
            RegDft<RA>::run(v);
            sm[r * ROW + c] = v[0];
#pragma unroll
            for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], twA[qq * RB + r]);
        }
        consumer_sync();
        const int c = tid % T;
        const int q = tid / T;
        float2 u[RB];
#pragma unroll
        for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
        consumer_sync();  // exchange consumed: the next tile streams in while this one is finished
        if (tid == 0 && i + (int)gridDim.x < total) {
            fence_proxy_async();
            issue_tile(i + gridDim.x);
        }
        RegDft<RB>::run(u);
        float2 *out = p.out + (size_t)frame * p.out_stride;
        const unsigned u1 = tile * T + c;
        const float scale = p.scale;
        float pw[RB];
#pragma unroll
        for (int s = 0; s < RB; s++) {
            const unsigned u2 = q + RA * s;
            const size_t k = ((size_t)u1 + (size_t)N1 * u2 + p.shift) & (M - 1);
            const float2 val = make_float2(u[s].x * scale, u[s].y * scale);
            out[k] = val;
            if (k < (size_t)p.additional) out[M + k] = val;  // IQ wrap tail, src/fft.cpp:96-97
            if constexpr (FUSE == 1) pw[s] = __fadd_rn(__fmul_rn(val.x, val.x), __fmul_rn(val.y, val.y));
            if constexpr (FUSE == 2)
                p.pscratch[(size_t)frame * M + (size_t)u2 * N1 + u1] = __fadd_rn(__fmul_rn(val.x, val.x), __fmul_rn(val.y, val.y));
        }
        if (p.npeers > 0) {  // NVLink peer copies of the frame (multi-GPU ingest rank only), off the common path
            for (int pe = 0; pe < p.npeers; pe++) {
                float2 *po = p.peers[pe] + (size_t)frame * p.out_stride;
#pragma unroll
                for (int s = 0; s < RB; s++) {
                    const unsigned k = (unsigned)(((size_t)u1 + (size_t)N1 * (q + RA * s) + p.shift) & (M - 1));
                    const float2 val = make_float2(u[s].x * scale, u[s].y * scale);
                    if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1]))
                        po[k] = val;
                    const unsigned kt = (unsigned)M + k;
                    if (k < (unsigned)p.additional &&
                        ((kt >= p.peer_lo[pe][0] && kt < p.peer_hi[pe][0]) || (kt >= p.peer_lo[pe][1] && kt < p.peer_hi[pe][1])))
                        po[kt] = val;
                }
            }
        }
        if constexpr (FUSE == 1) {
            // (the previous tile's epilogue readers are past the two barriers above)
#pragma unroll
            for (int s = 0; s < RB; s++) pt[(q + RA * s) * PP + c] = pw[s];
            consumer_sync();
            constexpr int LT = Log2<T>::v;
            const int L = p.levels;
            int8_t *quant = p.quant + (size_t)frame * p.pyr_stride;
            float *scr = p.pscratch + ((size_t)frame * (N1 / T) + tile) * N2;
#pragma unroll
            for (int u2 = tid; u2 < N2; u2 += kTmaThreads) {
                const unsigned d2 = (u2 + (N2 >> 1)) & (N2 - 1);
                const size_t dbase = (size_t)tile * T + (size_t)N1 * d2;
                float vv[T];
#pragma unroll
                for (int k = 0; k < T; k++) vv[k] = pt[u2 * PP + k];
                size_t lvl_off = 0;
                static_for<LT>([&](auto lvc) {
                    constexpr int lv = decltype(lvc)::value;
                    constexpr int CNT = T >> lv;
                    if (lv < L) {
                        unsigned w[(CNT + 3) / 4];
#pragma unroll
                        for (int k = 0; k < (CNT + 3) / 4; k++) w[k] = 0;
#pragma unroll
                        for (int k = 0; k < CNT; k++)
                            w[k / 4] |= (unsigned)quantize_dev(vv[k], p.size_log2 - lv) << (8 * (k % 4));
                        store_packed<CNT>(quant + lvl_off + (dbase >> lv), w);
                    }
                    lvl_off += M >> lv;
#pragma unroll
                    for (int k = 0; k < CNT / 2; k++) vv[k] = __fadd_rn(vv[2 * k], vv[2 * k + 1]);
                });
                scr[d2] = vv[0];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2, three-stage variant: ONE CTA per SM, two consumer groups of eight warps, three 64 KiB stages.
// Tile j of the CTA (j = 0, 1, 2, ...; global tile blockIdx.x + j*gridDim.x) is handled by group j % 2 in
// stage j % 3. A group hands its stage back as soon as its exchange is consumed and immediately streams tile
// j + 3 into it (which the OTHER group will consume), so every tile has a whole iteration of lead time and
// the SM always has 64-128 KiB of loads in flight - same 16 warps per SM as the two-CTA variant, twice the
// memory-level parallelism. The 1/N scale is folded into the (power-of-two scaled, hence bit-identical)
// intra-pass twiddles; spectrum / power addresses are one base pointer + compile-time offsets.
// ------------------------------------------------------------------------------------------------
constexpr int kP3Groups = 2;
constexpr int kP3Threads = kP3Groups * kTmaThreads;   // 512
constexpr int kP3Stages = 3;
struct P3Smem {
    static constexpr size_t kStage = sizeof(float2) * 32 * TmaSmem::kRow2;   // 65 792 B: raw tile (64 KiB) or exchange
    static constexpr size_t kTw = sizeof(float2) * 32 * 32;
    static constexpr size_t kTotal = kP3Stages * kStage + kTw + 64 + 64;  // + mbarriers + 2 x 8 warp sums (fused pyramid)
};
static_assert(P3Smem::kStage % 16 == 0, "stage must stay 16-byte aligned for bulk copies");
static_assert(P3Smem::kTotal <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kTmaThreads) : "memory");
}

struct GroupSync {
    int g;
    __device__ __forceinline__ void operator()() const { group_sync(g); }
};
// spin until *counter >= target (acquire); a counter that never gets there is a bug: give up after five seconds
__device__ __forceinline__ void wait_counter(const unsigned *counter, unsigned target) {
    unsigned long long t0 = 0;
    for (unsigned spins = 1;; spins++) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) return;
        __nanosleep(100);
        if ((spins & 1023u) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 5000000000ull) wait_timed_out(2);
        }
    }
}

// FUSE: 0 = spectrum only, 2 = + |X|^2 plane, 3 = spectrum + the whole waterfall pyramid inside this kernel: after a
// tile of frame f a group quantises two 4096-bin blocks of frame f - lag, whose bins every CTA has stored by then
// (per-frame completion counters `done`, release/acquire at GPU scope) and which are still in L2 - no second kernel,
// no re-read of the spectrum from DRAM. All CTAs are co-resident (grid <= SMs, one CTA per SM) and every wait targets
// work that precedes the waiter in the common tile order, so the waits cannot deadlock.
template <int FUSE, bool PEERS>
__global__ void __launch_bounds__(kP3Threads, 1)
    fft_pass2_tma3_kernel(const FwdParams p, int nframes, const PyrParams pyr, unsigned *done, int lag) {
    constexpr int T = kTmaT, RA = 32, RB = 32, N1 = kS, N2 = kS;
    constexpr int ROW = TmaSmem::kRow2;
    static_assert(FUSE == 0 || FUSE == 2 || FUSE == 3, "the three-stage kernel has no room for the smem-tiled waterfall epilogue");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2 *twA = reinterpret_cast<float2 *>(smem_raw + kP3Stages * P3Smem::kStage);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + kP3Stages * P3Smem::kStage + P3Smem::kTw);
    float *wsum = reinterpret_cast<float *>(smem_raw + kP3Stages * P3Smem::kStage + P3Smem::kTw + 64);

    const int tid = threadIdx.x;
    const int g = tid / kTmaThreads;       // consumer group
    const int gt = tid - g * kTmaThreads;  // thread within the group
    const int tiles_per_frame = N1 / T;
    const int total = tiles_per_frame * nframes;
    const size_t M = (size_t)N1 * N2;
    // One "full" barrier per (stage, consumer group): tile j completes on full[j % 3][j % 2], phase j / 6. A group only
    // ever waits on its own barriers, one phase after the other, so a parity wait can never be satisfied by an older
    // phase (with a barrier shared by both groups, a group running two phases ahead of a slow transfer would be).
    if (tid == 0) {
        for (int s = 0; s < kP3Stages * kP3Groups; s++) mbar_init(full + s, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    const float scale = p.scale;
    for (int i = tid; i < 32 * 32; i += kP3Threads) {
        const float2 w = p.twA2[i];
        twA[i] = make_float2(w.x * scale, w.y * scale);  // scale = 2^-k: exact, commutes with every later rounding
    }
    __syncthreads();
    auto issue_tile = [&](int j) {  // T consecutive rows of Y are one contiguous 64 KiB block
        const int i = blockIdx.x + j * gridDim.x;
        const int frame = i / tiles_per_frame, tile = i - frame * tiles_per_frame;
        const int st = j % kP3Stages;
        uint64_t *bar = full + st * kP3Groups + (j % kP3Groups);
        mbar_expect_tx(bar, sizeof(float2) * N2 * T);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.Y + (size_t)frame * M + (size_t)tile * T * N2);
        unsigned char *dst = smem_raw + (size_t)st * P3Smem::kStage;
        for (int k = 0; k < 4; k++) bulk_load_1d(dst + k * 16384, src + k * 16384, 16384, bar);
    };
    if (tid == 0)
        for (int j = 0; j < kP3Stages; j++)
            if ((int)blockIdx.x + j * (int)gridDim.x < total) issue_tile(j);

    auto pyramid_blocks = [&](int fp, int blk0, int count) {
        if (gt == 0) wait_counter(done + fp, (unsigned)tiles_per_frame);  // every bin of frame fp is stored and visible
        group_sync(g);
        for (int b = 0; b < count; b++) {
            pyramid_block<PYR_SPEC, 16, true, true, false>(pyr, fp, (unsigned)(blk0 + b), gt, wsum + 8 * g, GroupSync{g});
            group_sync(g);  // the block's warp sums are consumed before the next block overwrites them
        }
    };
    for (int j = g; (int)blockIdx.x + j * (int)gridDim.x < total; j += kP3Groups) {
        const int i = blockIdx.x + j * gridDim.x;
        const int frame = i / tiles_per_frame, tile = i - frame * tiles_per_frame;
        const int st = j % kP3Stages;
        float2 *sm = reinterpret_cast<float2 *>(smem_raw + (size_t)st * P3Smem::kStage);
        mbar_wait(full + st * kP3Groups + g, (j / (kP3Stages * kP3Groups)) & 1);
        {   // stage A: lanes along n2
            const int r = gt % 32;
            const int c = gt / 32;
            float2 v[RA];
#pragma unroll
            for (int jj = 0; jj < RA; jj++) v[jj] = sm[c * N2 + r + RB * jj];
            group_sync(g);
            RegDft<RA>::run(v);
            const float2 *tw = twA + r;
            sm[r * ROW + c] = make_float2(v[0].x * scale, v[0].y * scale);
#pragma unroll
            for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], tw[qq * RB]);
        }
        group_sync(g);
        const int c = gt % T;
        const int q = gt / T;
        float2 u[RB];
#pragma unroll
        for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
        group_sync(g);  // exchange consumed: the stage is free for tile j + 3
        if (gt == 0 && (int)blockIdx.x + (j + kP3Stages) * (int)gridDim.x < total) {
            fence_proxy_async();  // generic-proxy accesses to the stage are ordered before the TMA write
            issue_tile(j + kP3Stages);
        }
        RegDft<RB>::run(u);
        // bin k = (u1 + N1*u2 + shift) mod M with u2 = q + 32 s: one base pointer, compile-time offsets; the only
        // wrap is u = M-1 -> k = 0 (IQ), handled by redirecting that single store
        const unsigned u1 = tile * T + c;
        float2 *out = p.out + (size_t)frame * p.out_stride;
        float2 *o = out + u1 + p.shift + (size_t)N1 * q;
        const bool wraps = p.shift && u1 == N1 - 1 && q == RA - 1;
#pragma unroll
        for (int s = 0; s < RB - 1; s++) o[(size_t)N1 * RA * s] = u[s];
        *(wraps ? out : o + (size_t)N1 * RA * (RB - 1)) = u[RB - 1];
        if constexpr (FUSE == 2) {
            float *pp = p.pscratch + (size_t)frame * M + (size_t)q * N1 + u1;
#pragma unroll
            for (int s = 0; s < RB; s++)
                pp[(size_t)N1 * RA * s] = __fadd_rn(__fmul_rn(u[s].x, u[s].x), __fmul_rn(u[s].y, u[s].y));
        }
        // IQ wrap tail (src/fft.cpp:96-97): out[M + k] = out[k] for k < additional. k grows with s, so only
        // threads whose s = 0 bin is inside the tail ever enter.
        const unsigned k0 = u1 + p.shift + N1 * q;
        if (k0 < (unsigned)p.additional) {
#pragma unroll
            for (int s = 0; s < RB; s++)
                if (k0 + (unsigned)(N1 * RA * s) < (unsigned)p.additional && !(wraps && s == RB - 1))
                    o[M + (size_t)N1 * RA * s] = u[s];
        }
        if (wraps && p.additional > 0) out[M] = u[RB - 1];
        if constexpr (PEERS) {  // NVLink peer copies of the frame (multi-GPU ingest rank only)
            for (int pe = 0; pe < p.npeers; pe++) {
                float2 *po = p.peers[pe] + (size_t)frame * p.out_stride;
#pragma unroll
                for (int s = 0; s < RB; s++) {
                    const unsigned k = (unsigned)(((size_t)u1 + (size_t)N1 * (q + RA * s) + p.shift) & (M - 1));
                    if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1]))
                        po[k] = u[s];
                    const unsigned kt = (unsigned)M + k;
                    if (k < (unsigned)p.additional &&
                        ((kt >= p.peer_lo[pe][0] && kt < p.peer_hi[pe][0]) || (kt >= p.peer_lo[pe][1] && kt < p.peer_hi[pe][1])))
                        po[kt] = u[s];
                }
            }
        }
        if constexpr (FUSE == 3) {
            // publish this tile (every thread fences its own stores, then one arrival per tile) ...
            __threadfence();
            group_sync(g);
            if (gt == 0) atomicAdd(done + frame, 1u);
            // ... and quantise the two pyramid blocks that pair with it, `lag` frames back
            if (frame >= lag) pyramid_blocks(frame - lag, 2 * tile, 2);
        }
    }
    if constexpr (FUSE == 3) {
        // the last `lag` frames have no later tile to ride on: their blocks are spread over all groups
        constexpr int kBlocksPerFrame = N1 * N2 / (256 * 16);
        const int first = nframes > lag ? nframes - lag : 0;
        const int nblocks = (nframes - first) * kBlocksPerFrame;
        for (int b = (int)blockIdx.x * kP3Groups + g; b < nblocks; b += kP3Groups * (int)gridDim.x)
            pyramid_blocks(first + b / kBlocksPerFrame, b % kBlocksPerFrame, 1);
    }
}

}  // namespace b200
