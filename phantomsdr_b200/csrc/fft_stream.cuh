// The whole forward group of the 2^20-point c2c transform - FFT pass 1, FFT pass 2 and the waterfall pyramid - as ONE
// persistent, dataflow-scheduled launch per batch of frames (replaces fft_pass1_tma_kernel + fft_pass2_tma3_kernel +
// pyramid_kernel for the headline size; reference: FFTW::load_complex_input + fftwf_execute + power_and_quantize +
// half_and_quantize, src/fft_impl.cpp:136-174).
//
// Why: with one kernel per pass every intermediate makes a round trip through DRAM (the four-step intermediate Y is
// written and read once, the spectrum is read again by the quantiser: 38 MB of DRAM traffic per frame against 19 MB of
// algorithmic bytes). Here the three stages of consecutive frames run concurrently on every SM,
//
//      P1(f)  : 8 columns of frame f   -> window, 1024-point column DFTs, inter-pass twiddles -> Y ring slot f % K
//      P2(f)  : 8 rows of Y            -> 1024-point row DFTs -> normalised spectrum (+ wrap tail, + NVLink peers)
//      P3(f)  : 8192 display bins      -> |X|^2, approx-log2, int8, the pairwise-sum pyramid
//
// a few frames apart, so Y (a ring of K frame slots) and the spectrum a quantiser item reads are still in the 126 MB L2
// when they are consumed: DRAM sees the input once, the spectrum once and the pyramid once.
//
// Structure
//   * grid = G CTAs (<= SMs, all co-resident), one CTA of two 8-warp consumer groups per SM, three 66 KiB stages.
//     The host cuts the global item list (frame-slot major: P1 tiles of frame s, P2 tiles of frame s - lag1, P3 chunks of
//     frame s - lag2) round-robin over the CTAs; item j of a CTA is processed by group j % 2 in stage j % 3.
//   * every item is 64 KiB of input streamed into its stage by TMA (P1: 2-D boxes of the strided column tile of the hop
//     ring; P2: bulk copies of 8 contiguous rows of Y; P3: 2-D boxes with the 128-byte swizzle so that a thread's 16
//     consecutive bins are conflict-free 128-bit reads) and signalled by an mbarrier per (stage, group). The stage doubles
//     as the exchange buffer between the two register DFTs; it is handed back as soon as the exchange is consumed.
//   * dependencies between items of different CTAs are per-frame completion counters in global memory (release: every
//     thread fences its stores, one atomicAdd per item; acquire: ld.acquire.gpu + fence.proxy.async before the TMA read).
//     Loads are issued by a non-blocking "pump" (in item order, when the stage is free and the counter has arrived) that
//     the group leaders run at every hand-over and inside every wait, so no thread ever blocks on a dependency while
//     holding work that others wait for: the globally earliest unfinished item can always proceed. Waits are bounded
//     (two seconds) and end in an abort flag that every spin loop polls and the host reports - no trap.
//   * pass 1 builds the Hann weights on the fly (angle addition from a 1024-entry table in shared memory and one
//     cos/sin pair per thread; the 1/N normalisation is folded into that table), so an item needs no window slice and any
//     CTA can take any tile.
#pragma once
#include "fft_tma.cuh"

namespace b200 {

enum { IT_P1 = 0, IT_P2 = 1, IT_P3 = 2 };
__host__ __device__ inline unsigned stream_item(int type, int frame, int tile) {
    return ((unsigned)type << 28) | ((unsigned)frame << 16) | (unsigned)tile;
}

struct StreamSync {
    unsigned doneP1[64];   // pass-1 tiles of frame f stored
    unsigned doneP2[64];   // pass-2 tiles of frame f stored
    unsigned abort;        // a bounded wait expired somewhere: everybody leaves
    unsigned pad[31];
};

struct StreamParams {
    FwdParams fp;
    PyrParams pyr;
    const unsigned *items;   // [grid][max_items]
    const int *nitems;       // [grid]
    int max_items;
    int K;                   // Y ring slots (frames)
    int nframes;
    const float2 *winT;      // [1024] (h cos, h sin)(2 pi row / 1024), h = scale / 2
    float whalf;             // h
    unsigned spec_row0;      // tensor-map row (16 bins) that holds bin 1 of frame 0 of this launch
    unsigned spec_rows_per_frame;
    int nodeps;              // profiling aid: item types run without their dependencies (stage mask)
    StreamSync *sync;
};

struct StreamSmem {
    static constexpr size_t kStage = TmaSmem::kStage;                 // 67 584 B: raw tile, either exchange layout
    static constexpr size_t kOffTw = 3 * kStage;                      // W_1024^(r q)
    static constexpr size_t kOffWin = kOffTw + sizeof(float2) * 1024; // window table
    static constexpr size_t kOffBars = kOffWin + sizeof(float2) * 1024;
    static constexpr size_t kOffWsum = kOffBars + 64;                 // 2 x 8 warp sums of the pyramid tree
    static constexpr size_t kOffCtl = kOffWsum + 64;                  // pump state: cursor, lock, freed[3]
    static constexpr size_t kTotal = kOffCtl + 64;
};
static_assert(StreamSmem::kStage >= P3Smem::kStage && StreamSmem::kStage >= 65536, "stage must hold every layout");
static_assert(StreamSmem::kTotal <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
static_assert(StreamSmem::kStage % 1024 == 0, "swizzled TMA destinations need 1024-byte alignment");

__device__ __forceinline__ bool group_or(int g, bool pred) {  // barrier of the group's 256 threads + OR of pred
    unsigned r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "barrier.cta.red.or.pred q, %2, %3, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((unsigned)pred), "r"(g + 1), "n"(kTmaThreads)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    tma_load_2d(dst, map, x, y, bar);
}

template <bool PEERS>
__global__ void __launch_bounds__(kP3Threads, 1)
    fwd_stream_kernel(const StreamParams sp, const __grid_constant__ CUtensorMap ring_map,
                      const __grid_constant__ CUtensorMap spec_map) {
    constexpr int T = kTmaT, RA = 32, RB = 32, N1 = kS, N2 = kS, NT = N2 / T;
    constexpr unsigned long long kTimeoutNs = 2000000000ull;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2 *twA = reinterpret_cast<float2 *>(smem_raw + StreamSmem::kOffTw);
    float2 *winT = reinterpret_cast<float2 *>(smem_raw + StreamSmem::kOffWin);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + StreamSmem::kOffBars);  // [stage][group]
    float *wsum = reinterpret_cast<float *>(smem_raw + StreamSmem::kOffWsum);
    volatile int *ctl = reinterpret_cast<volatile int *>(smem_raw + StreamSmem::kOffCtl);  // [0] cursor [1] lock [2..4] freed

    const FwdParams &p = sp.fp;
    const int tid = threadIdx.x;
    const int g = tid / kTmaThreads;
    const int gt = tid - g * kTmaThreads;
    const size_t M = (size_t)N1 * N2;
    const unsigned *items = sp.items + (size_t)blockIdx.x * sp.max_items;
    const int nitems = sp.nitems[blockIdx.x];
    StreamSync *sy = sp.sync;
    if (tid == 0) {
        for (int s = 0; s < 6; s++) mbar_init(full + s, 1);
        for (int s = 0; s < 5; s++) ctl[s] = 0;
        fence_barrier_init();
        fence_proxy_async();
    }
    for (int i = tid; i < 32 * 32; i += kP3Threads) {
        twA[i] = p.twA1[i];  // N1 == N2: both passes share W_1024^(r q)
        winT[i] = sp.winT[i];
    }
    __syncthreads();

    // ---- the pump: issue the loads of this CTA's items, in order, whenever stage and dependency allow ----
    auto pump = [&]() {
        if (atomicCAS(const_cast<int *>(ctl + 1), 0, 1) != 0) return;  // the other leader is pumping
        for (;;) {
            const int jj = ctl[0];
            if (jj >= nitems) break;
            const int st = jj % 3;
            if (ctl[2 + st] < jj / 3) break;  // item jj - 3 still owns the stage
            const unsigned it = items[jj];
            const int type = it >> 28, frame = (it >> 16) & 0xFFF, tile = it & 0xFFFF;
            if (!sp.nodeps) {
                if (type == IT_P2 && ld_acquire(sy->doneP1 + frame) < (unsigned)NT) break;
                if (type == IT_P3 && ld_acquire(sy->doneP2 + frame) < (unsigned)NT) break;
            }
            uint64_t *bar = full + st * 2 + (jj & 1);
            unsigned char *dst = smem_raw + (size_t)st * StreamSmem::kStage;
            fence_proxy_async_all();  // stage reads (generic proxy) and the acquired global state before the async writes / reads
            mbar_expect_tx(bar, 65536);
            if (type == IT_P1) {
                const int hopA = (p.hop0 + frame) % p.nhops, hopB = (p.hop0 + frame + 1) % p.nhops;
                // rows 0..511 of the frame come from the older hop, 512..1023 from the newer one
                tma_load_2d(dst + 0 * 16384, &ring_map, tile * T * 2, hopA * (N1 / 2), bar);
                tma_load_2d(dst + 1 * 16384, &ring_map, tile * T * 2, hopA * (N1 / 2) + 256, bar);
                tma_load_2d(dst + 2 * 16384, &ring_map, tile * T * 2, hopB * (N1 / 2), bar);
                tma_load_2d(dst + 3 * 16384, &ring_map, tile * T * 2, hopB * (N1 / 2) + 256, bar);
            } else if (type == IT_P2) {  // T consecutive rows of Y (ring slot frame % K) are one contiguous 64 KiB block
                const unsigned char *src =
                    reinterpret_cast<const unsigned char *>(p.Y + (size_t)(frame % sp.K) * M + (size_t)tile * T * N2);
                for (int k = 0; k < 4; k++) bulk_load_1d(dst + k * 16384, src + k * 16384, 16384, bar);
            } else {  // 8192 display bins = 512 rows of 16 bins, starting at display bin 8192 * tile
                const unsigned row = sp.spec_row0 + (unsigned)frame * sp.spec_rows_per_frame +
                                     ((((unsigned)tile * 8192u + (unsigned)(M >> 1)) & (unsigned)(M - 1)) >> 4);
                tma_load_2d(dst, &spec_map, 0, (int)row, bar);
                tma_load_2d(dst + 32768, &spec_map, 0, (int)row + 256, bar);
            }
            ctl[0] = jj + 1;
        }
        __threadfence_block();
        atomicExch(const_cast<int *>(ctl + 1), 0);
    };
    // bounded waits; the leader of a group keeps the pump running while it waits. false = abort
    auto wait_full = [&](uint64_t *bar, unsigned parity) -> bool {
        if (mbar_test_wait(bar, parity)) return true;
        unsigned long long t0 = 0;
        if (gt < 32) {
            // the leader's warp polls (mbarrier.try_wait puts the whole warp to sleep until the phase completes, which it
            // cannot while the load it waits for has not been issued): lane 0 keeps the pump running
            for (unsigned spins = 1;; spins++) {
                if (mbar_test_wait(bar, parity)) return true;
                if (gt == 0) pump();
                __nanosleep(32);
                if ((spins & 63u) == 0) {
                    if (ld_relaxed(&sy->abort)) return false;
                    const unsigned long long now = global_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > kTimeoutNs) {
                        atomicExch(&sy->abort, 1u);
                        return false;
                    }
                }
            }
        }
        for (unsigned spins = 1;; spins++) {
            if (mbar_try_wait(bar, parity)) return true;
            if ((spins & 15u) == 0) {
                if (ld_relaxed(&sy->abort)) return false;
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kTimeoutNs) {
                    atomicExch(&sy->abort, 1u);
                    return false;
                }
            }
        }
    };
    auto wait_done = [&](const unsigned *counter, unsigned target) -> bool {  // leader only
        unsigned long long t0 = 0;
        for (unsigned spins = 1;; spins++) {
            if (ld_acquire(counter) >= target) return true;
            pump();
            __nanosleep(64);
            if ((spins & 63u) == 0) {
                if (ld_relaxed(&sy->abort)) return false;
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kTimeoutNs) {
                    atomicExch(&sy->abort, 1u);
                    return false;
                }
            }
        }
    };
    auto release_stage = [&](int st) {  // leader, after the group barrier that ends the stage's use
        atomicAdd(const_cast<int *>(ctl + 2 + st), 1);
        pump();
    };
    if (tid == 0) pump();

    auto tw_lookup = [&](unsigned e) { return cmul(__ldg(p.TL + (e & 1023u)), __ldg(p.TH + ((e >> 10) & 1023u))); };

    for (int jj = g; jj < nitems; jj += kP3Groups) {
        const unsigned it = items[jj];
        const int type = it >> 28, frame = (it >> 16) & 0xFFF, tile = it & 0xFFFF;
        const int st = jj % 3;
        float2 *sm = reinterpret_cast<float2 *>(smem_raw + (size_t)st * StreamSmem::kStage);
        uint64_t *bar = full + st * 2 + g;
        const unsigned parity = (jj / 6) & 1;
        if (type == IT_P1) {
            // ===================== pass 1: column tile `tile` of frame `frame` =====================
            constexpr int ROW = TmaSmem::kRow1;
            const int c = gt % T;
            const int r = gt / T;
            const int q = r;
            const int n2 = tile * T + c;
            // inter-pass twiddles W_M^(n2*k1), k1 = q + 32 s, s = 4a + b:  G[a] * B[b-1]  (the IQ k1 = 0 row carries the
            // one-slot rotation W_M^(N1*n2) instead); fetched before the wait so that the lookups overlap it
            float2 G[8], B[3];
#pragma unroll
            for (int a = 0; a < 8; a++) G[a] = tw_lookup((unsigned)n2 * (unsigned)(q + 128 * a));
#pragma unroll
            for (int b = 1; b < 4; b++) B[b - 1] = tw_lookup((unsigned)n2 * 32u * (unsigned)b);
            const float2 rot0 = tw_lookup((unsigned)N1 * (unsigned)n2);
            const float2 wn = __ldg(p.TL + n2);  // (cos, -sin)(2 pi n2 / M)
            const float cB = wn.x, sB = -wn.y, wh = sp.whalf;
            const bool ok = wait_full(bar, parity);
            float2 v[RA];
#pragma unroll
            for (int j = 0; j < RA; j++) {
                float2 x = sm[(r + RB * j) * T + c];
                // Hann weight of sample (r + 32 j) * 1024 + n2 times 1/N: h - h cos(A + B) = h - (h cA) cB + (h sA) sB
                const float2 tab = winT[r + RB * j];
                const float w = fmaf(tab.y, sB, fmaf(-tab.x, cB, wh));
                x.x *= w;
                x.y *= w;
                v[j] = x;
            }
            if (group_or(g, !ok)) return;  // the raw tile is in registers: the stage becomes the exchange buffer
            RegDft<RA>::run(v);
            sm[r * ROW + c] = v[0];
#pragma unroll
            for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], twA[qq * RB + r]);
            group_sync(g);
            float2 u[RB];
#pragma unroll
            for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
            // back-pressure: ring slot frame % K was last read by pass 2 of frame - K
            bool ok2 = true;
            if (gt == 0 && frame >= sp.K && !sp.nodeps) ok2 = wait_done(sy->doneP2 + (frame - sp.K), NT);
            if (group_or(g, !ok2)) return;  // exchange consumed, slot free
            if (gt == 0) release_stage(st);
            RegDft<RB>::run(u);
            float2 *Y = p.Y + (size_t)(frame % sp.K) * M + n2;
#pragma unroll
            for (int s = 0; s < RB; s++) {
                const int k1 = q + RA * s;
                int u1 = k1 - 1;  // IQ display shift (fft_fwd.cuh): rows stored at (k1 - 1) mod N1
                if (u1 < 0) u1 += N1;
                float2 tw = (s & 3) ? cmul(G[s >> 2], B[(s & 3) - 1]) : G[s >> 2];
                if (s == 0 && q == 0) tw = rot0;
                Y[(size_t)u1 * N2] = cmul(u[s], tw);
            }
            group_sync(g);  // every store of the tile is issued ...
            if (gt == 0) {
                __threadfence();  // ... the leader's fence is cumulative over what the barrier ordered before it (the grid-sync pattern) ...
                atomicAdd(sy->doneP1 + frame, 1u);  // ... one arrival per tile
            }
        } else if (type == IT_P2) {
            // ===================== pass 2: row tile `tile` of frame `frame` =====================
            constexpr int ROW = TmaSmem::kRow2;
            const bool ok = wait_full(bar, parity);
            {   // stage A: lanes along n2
                const int r = gt % 32;
                const int c = gt / 32;
                float2 v[RA];
#pragma unroll
                for (int j = 0; j < RA; j++) v[j] = sm[c * N2 + r + RB * j];
                if (group_or(g, !ok)) return;
                RegDft<RA>::run(v);
                const float2 *tw = twA + r;
                sm[r * ROW + c] = v[0];
#pragma unroll
                for (int qq = 1; qq < RA; qq++) sm[r * ROW + qq * T + c] = cmul(v[qq], tw[qq * RB]);
            }
            group_sync(g);
            const int c = gt % T;
            const int q = gt / T;
            float2 u[RB];
#pragma unroll
            for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
            group_sync(g);  // exchange consumed
            if (gt == 0) release_stage(st);
            RegDft<RB>::run(u);
            // bin k = (u1 + N1*u2 + 1) mod M with u2 = q + 32 s: one base pointer, compile-time offsets; the only wrap
            // is u = M-1 -> k = 0, handled by redirecting that single store. Bin 0 is always mirrored at index M: it is
            // the wrap tail's first bin (src/fft.cpp:96-97) and the last bin of the display row the quantiser reads.
            const unsigned u1 = tile * T + c;
            float2 *out = p.out + (size_t)frame * p.out_stride;
            float2 *o = out + u1 + 1 + (size_t)N1 * q;
            const bool wraps = u1 == N1 - 1 && q == RA - 1;
#pragma unroll
            for (int s = 0; s < RB - 1; s++) o[(size_t)N1 * RA * s] = u[s];
            *(wraps ? out : o + (size_t)N1 * RA * (RB - 1)) = u[RB - 1];
            const unsigned k0 = u1 + 1 + N1 * q;
            if (k0 < (unsigned)p.additional) {
#pragma unroll
                for (int s = 0; s < RB; s++)
                    if (k0 + (unsigned)(N1 * RA * s) < (unsigned)p.additional && !(wraps && s == RB - 1))
                        o[M + (size_t)N1 * RA * s] = u[s];
            }
            if (wraps) out[M] = u[RB - 1];
            if constexpr (PEERS) {  // NVLink peer copies of the frame (multi-GPU ingest rank only)
                for (int pe = 0; pe < p.npeers; pe++) {
                    float2 *po = p.peers[pe] + (size_t)frame * p.out_stride;
#pragma unroll
                    for (int s = 0; s < RB; s++) {
                        const unsigned k = (unsigned)(((size_t)u1 + (size_t)N1 * (q + RA * s) + 1) & (M - 1));
                        if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1]))
                            po[k] = u[s];
                        const unsigned kt = (unsigned)M + k;
                        if (k < (unsigned)p.additional &&
                            ((kt >= p.peer_lo[pe][0] && kt < p.peer_hi[pe][0]) || (kt >= p.peer_lo[pe][1] && kt < p.peer_hi[pe][1])))
                            po[kt] = u[s];
                    }
                }
            }
            group_sync(g);
            if (gt == 0) {
                __threadfence();  // the spectrum tile is visible before the frame is handed to the quantiser / the slot to pass 1
                atomicAdd(sy->doneP2 + frame, 1u);
            }
        } else {
            // ===================== waterfall: display bins [8192 tile, 8192 tile + 8192) of frame `frame` =====================
            const bool ok = wait_full(bar, parity);
            float pw[2][16];
            const unsigned char *stage = reinterpret_cast<const unsigned char *>(sm);
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int row = 256 * b + gt;  // 16 consecutive bins = one 128-byte row, 16-byte chunks XOR-swizzled by row % 8
                const unsigned char *rp = stage + (size_t)row * 128;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 x = *reinterpret_cast<const float4 *>(rp + ((i ^ (row & 7)) << 4));
                    float a, bb, cc, d;
                    const unsigned long long v0 = pk(x.x, x.y), v1 = pk(x.z, x.w);
                    unpk(pk_mul(v0, v0), a, bb);
                    unpk(pk_mul(v1, v1), cc, d);
                    pw[b][2 * i] = __fadd_rn(a, bb);
                    pw[b][2 * i + 1] = __fadd_rn(cc, d);
                }
            }
            if (group_or(g, !ok)) return;  // the chunk is in registers
            if (gt == 0) release_stage(st);
#pragma unroll
            for (int b = 0; b < 2; b++) {
                pyramid_tree<16, true, false>(sp.pyr, frame, (unsigned)(2 * tile + b), gt, pw[b], 0, wsum + 8 * g, GroupSync{g});
                group_sync(g);  // the block's warp sums are consumed before the next block overwrites them
            }
        }
    }
}

}  // namespace b200
