// EXPERIMENTAL (B200_OPT_TMA 4) - written at the end of round 1, compiled, NOT YET RUN ON A GPU.
//
// Both FFT passes of the 2^20-point c2c transform in ONE persistent launch, so that the intermediate Y never makes its
// round trip through DRAM (16 of the 38 MB the forward group moves per frame, profiles/README.md): pass 1 runs up to
// K frames ahead of pass 2 and Y is a ring of K frame slots (K * 8 MiB, L2-resident) instead of a batch-sized buffer.
//
//   grid = number of SMs (>= 128), one CTA of two 8-warp consumer groups per SM, one 64 KiB stage per group
//   CTA b < 128 : group 0 = pass-1 worker of column tile b for every frame (its Hann slice and twiddles stay put, as
//                 in fft_pass1_tma_kernel order 0), group 1 = pass-2 worker
//   CTA b >= 128: both groups are pass-2 workers
//   pass-2 worker w of W takes the tiles i = w, w + W, ... of the frame-major list (frame = i / 128, tile = i % 128)
//
// Dependencies are per-frame counters in global memory (release: every thread fences its stores, one atomicAdd per
// tile; acquire: one thread spins, then the group barrier - the protocol of fft_pass2_tma3_kernel<3>):
//   doneP1[f] == 128  before a pass-2 worker lets TMA read Y of frame f
//   doneP2[f - K] == 128  before a pass-1 worker overwrites ring slot f % K
// Every wait targets a strictly older frame (K >= 2) and all CTAs are co-resident, so the waits cannot deadlock;
// they are time-bounded like the mbarrier waits.
#pragma once
#include "fft_tma.cuh"

namespace b200 {

struct FusedSmem {
    static constexpr size_t kStage0 = TmaSmem::kStage;   // pass-1 exchange pitch (264) - also fits the pass-2 layout
    static constexpr size_t kStage1 = P3Smem::kStage;    // pass-2 exchange pitch (257)
    static constexpr size_t kTw = sizeof(float2) * 32 * 32;
    static constexpr size_t kOffTwA = kStage0 + kStage1;
    static constexpr size_t kOffTwS = kOffTwA + kTw;
    static constexpr size_t kOffWin = kOffTwS + kTw;
    static constexpr size_t kOffBars = kOffWin + TmaSmem::kWindowC;
    static constexpr size_t kTotal = kOffBars + 64;
};
static_assert(FusedSmem::kStage0 >= FusedSmem::kStage1, "group 0 must be able to run the pass-2 layout in its stage");
static_assert(FusedSmem::kTotal <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
static_assert(FusedSmem::kStage0 % 16 == 0 && FusedSmem::kOffWin % 128 == 0, "TMA destinations must stay aligned");

// orders what this thread has observed of global memory (through the acquire on a counter) before its TMA reads of it
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(kP3Threads, 1)
    fft_fused12_kernel(const FwdParams p, const __grid_constant__ CUtensorMap ring_map,
                       const __grid_constant__ CUtensorMap window_map, int nframes, int K, unsigned *doneP1,
                       unsigned *doneP2) {
    constexpr int T = kTmaT, RA = 32, RB = 32, N1 = kS, N2 = kS, NT = N2 / T;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2 *twA = reinterpret_cast<float2 *>(smem_raw + FusedSmem::kOffTwA);   // W_1024^(r q)
    float2 *twS = reinterpret_cast<float2 *>(smem_raw + FusedSmem::kOffTwS);   // the same, times 1/N (pass 2)
    float *win = reinterpret_cast<float *>(smem_raw + FusedSmem::kOffWin);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + FusedSmem::kOffBars);  // full[0], full[1], window

    const int tid = threadIdx.x;
    const int g = tid / kTmaThreads;
    const int gt = tid - g * kTmaThreads;
    const size_t M = (size_t)N1 * N2;
    const bool p1cta = (int)blockIdx.x < NT;
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    const float scale = p.scale;
    for (int i = tid; i < 32 * 32; i += kP3Threads) {
        const float2 w = p.twA1[i];  // N1 == N2: both passes share W_1024^(r q)
        twA[i] = w;
        twS[i] = make_float2(w.x * scale, w.y * scale);  // scale = 2^-k: exact
    }
    __syncthreads();
    unsigned char *stage_raw = smem_raw + (g == 0 ? 0 : FusedSmem::kStage0);
    float2 *sm = reinterpret_cast<float2 *>(stage_raw);
    uint64_t *full = bars + g;

    if (p1cta && g == 0) {
        // ===================== pass-1 worker: column tile `tile`, frames 0 .. nframes-1 =====================
        constexpr int ROW = TmaSmem::kRow1;
        uint64_t *wbar = bars + 2;
        const int tile = blockIdx.x;
        auto issue_frame = [&](int f) {
            mbar_expect_tx(full, sizeof(float2) * N1 * T);
            const int hopA = (p.hop0 + f) % p.nhops, hopB = (p.hop0 + f + 1) % p.nhops;
            tma_load_2d(stage_raw + 0 * 16384, &ring_map, tile * T * 2, hopA * (N1 / 2), full);
            tma_load_2d(stage_raw + 1 * 16384, &ring_map, tile * T * 2, hopA * (N1 / 2) + 256, full);
            tma_load_2d(stage_raw + 2 * 16384, &ring_map, tile * T * 2, hopB * (N1 / 2), full);
            tma_load_2d(stage_raw + 3 * 16384, &ring_map, tile * T * 2, hopB * (N1 / 2) + 256, full);
        };
        if (gt == 0) {
            constexpr uint32_t kWinBytes = sizeof(float) * N1 * T;
            mbar_expect_tx(wbar, kWinBytes);
            for (int b = 0; b < 4; b++)
                tma_load_2d(reinterpret_cast<unsigned char *>(win) + b * (kWinBytes / 4), &window_map, tile * T, b * 256, wbar);
            if (nframes > 0) issue_frame(0);
        }
        const int c = gt % T;
        const int r = gt / T;
        const int q = r;
        const int n2 = tile * T + c;
        float2 G[8], B[3];
        auto tw_lookup = [&](unsigned e) { return cmul(__ldg(p.TL + (e & 1023u)), __ldg(p.TH + ((e >> 10) & 1023u))); };
#pragma unroll
        for (int a = 0; a < 8; a++) G[a] = tw_lookup((unsigned)n2 * (unsigned)(q + 128 * a));
#pragma unroll
        for (int b = 1; b < 4; b++) B[b - 1] = tw_lookup((unsigned)n2 * 32u * (unsigned)b);
        const float2 rot0 = tw_lookup((unsigned)N1 * (unsigned)n2);
        mbar_wait(wbar, 0);
        for (int f = 0; f < nframes; f++) {
            mbar_wait(full, f & 1);
            float2 v[RA];
#pragma unroll
            for (int j = 0; j < RA; j++) {
                float2 x = sm[(r + RB * j) * T + c];
                const float w = win[(r + RB * j) * T + c];
                x.x *= w;
                x.y *= w;
                v[j] = x;
            }
            group_sync(g);  // the raw tile is in registers: the stage becomes the exchange buffer
            RegDft<RA>::run(v);
            sm[r * ROW + c] = v[0];
#pragma unroll
            for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], twA[qq * RB + r]);
            group_sync(g);
            float2 u[RB];
#pragma unroll
            for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
            // back-pressure: ring slot f % K was last read by pass 2 of frame f - K
            if (gt == 0 && f >= K) wait_counter(doneP2 + (f - K), NT);
            group_sync(g);  // exchange consumed, slot free: the next frame streams in while this one is finished
            if (gt == 0 && f + 1 < nframes) {
                fence_proxy_async();
                issue_frame(f + 1);
            }
            RegDft<RB>::run(u);
            float2 *Y = p.Y + (size_t)(f % K) * M + n2;
#pragma unroll
            for (int s = 0; s < RB; s++) {
                const int k1 = q + RA * s;
                int u1 = k1 - 1;  // IQ display shift (fft_fwd.cuh): rows stored at (k1 - 1) mod N1
                if (u1 < 0) u1 += N1;
                float2 tw = (s & 3) ? cmul(G[s >> 2], B[(s & 3) - 1]) : G[s >> 2];
                if (s == 0 && q == 0) tw = rot0;  // the k1 = 0 row carries the one-slot rotation
                Y[(size_t)u1 * N2] = cmul(u[s], tw);
            }
            __threadfence();  // publish this tile of Y ...
            group_sync(g);
            if (gt == 0) atomicAdd(doneP1 + f, 1u);  // ... one arrival per tile
        }
        return;
    }

    // ===================== pass-2 worker =====================
    constexpr int ROW = TmaSmem::kRow2;
    const int W = NT + 2 * ((int)gridDim.x - NT);  // one worker on the pass-1 CTAs, two on the others
    const int w0 = p1cta ? (int)blockIdx.x : NT + 2 * ((int)blockIdx.x - NT) + g;
    const int total = NT * nframes;
    auto issue_tile = [&](int i) {  // T consecutive rows of Y (ring slot frame % K) are one contiguous 64 KiB block
        const int frame = i / NT, tile = i - frame * NT;
        wait_counter(doneP1 + frame, NT);  // every pass-1 tile of the frame is stored and visible ...
        fence_proxy_async_all();           // ... also to the async proxy that performs the copy
        mbar_expect_tx(full, sizeof(float2) * N2 * T);
        const unsigned char *src =
            reinterpret_cast<const unsigned char *>(p.Y + (size_t)(frame % K) * M + (size_t)tile * T * N2);
        for (int k = 0; k < 4; k++) bulk_load_1d(stage_raw + k * 16384, src + k * 16384, 16384, full);
    };
    if (gt == 0 && w0 < total) issue_tile(w0);
    int it = 0;
    for (int i = w0; i < total; i += W, it++) {
        const int frame = i / NT, tile = i - frame * NT;
        mbar_wait(full, it & 1);
        {   // stage A: lanes along n2
            const int r = gt % 32;
            const int c = gt / 32;
            float2 v[RA];
#pragma unroll
            for (int jj = 0; jj < RA; jj++) v[jj] = sm[c * N2 + r + RB * jj];
            group_sync(g);
            RegDft<RA>::run(v);
            const float2 *tw = twS + r;
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
        group_sync(g);  // exchange consumed: the stage is free for this worker's next tile
        if (gt == 0 && i + W < total) {
            fence_proxy_async();
            issue_tile(i + W);
        }
        RegDft<RB>::run(u);
        const unsigned u1 = tile * T + c;
        float2 *out = p.out + (size_t)frame * p.out_stride;
        float2 *o = out + u1 + 1 + (size_t)N1 * q;  // bin k = u + 1 (IQ display shift)
        const bool wraps = u1 == N1 - 1 && q == RA - 1;  // u = M - 1 -> k = 0
#pragma unroll
        for (int s = 0; s < RB - 1; s++) o[(size_t)N1 * RA * s] = u[s];
        *(wraps ? out : o + (size_t)N1 * RA * (RB - 1)) = u[RB - 1];
        const unsigned k0 = u1 + 1 + N1 * q;
        if (k0 < (unsigned)p.additional) {  // IQ wrap tail, src/fft.cpp:96-97
#pragma unroll
            for (int s = 0; s < RB; s++)
                if (k0 + (unsigned)(N1 * RA * s) < (unsigned)p.additional && !(wraps && s == RB - 1))
                    o[M + (size_t)N1 * RA * s] = u[s];
        }
        if (wraps && p.additional > 0) out[M] = u[RB - 1];
        __threadfence();  // the spectrum tile is visible before the slot is handed back to pass 1
        group_sync(g);
        if (gt == 0) atomicAdd(doneP2 + frame, 1u);
    }
}

}  // namespace b200
