// Sequential tails of the per-client chain - DC blocker (src/utils.h:76-99,139-169), look-ahead AGC
// (src/utils/audioprocessing.cpp:17-74), float -> int16 (src/utils/dsp.cpp:152-165) - bit-exact float op order, as an
// elastic warp-specialised pipeline with ONE LANE PER CLIENT.
//
// The three float recurrences of the chain (two running sums, the AGC gain) are strictly serial per client; their
// critical path (the gain: sub, mul, max, sub = ~18 cycles per sample) is the floor of the whole client path. Everything
// here is arranged so that nothing else is slower than that chain:
//   * a CTA owns a group of 32 client slots; every warp is one pipeline stage and its 32 lanes are the 32 clients, so one
//     warp instruction advances 32 clients by one sample and index arithmetic is per warp, not per sample and client;
//   * stages hand chunks of 32 samples to one another through small rings in shared memory ([sample][client], pitch 33)
//     behind full / empty mbarriers - no CTA-wide barrier anywhere, a stage runs as far ahead as its ring allows;
//   * the AGC's sliding maximum over the 200 ms look-ahead is the maximum of three pieces that are all O(1) per sample:
//     the suffix maximum inside the block the window's old end walks through (kept beside the look-ahead ring in global
//     memory, built once when a block completes), the maxima of the whole blocks in between, and the running maximum of
//     the frame being pushed. Rings and state are stored [sample][client] per group, so a warp access is one 128-byte line.
//
//   stage (warp)   per sample and lane
//   0 load         audio (global, one client row at a time, lanes along samples) -> X ring, transposed
//   1 sum1         s1 = (s1 - X[t-D]) + X[t];  M[t] = s1 / D                               (serial: 2 adds)
//   2 sum2         s2 = (s2 - M[t-D]) + M[t];  Y[t] = X[t-D+1] - s2 / D                    (serial: 2 adds)
//   3 block        run = max(run, |Y[t]|) -> R ring; Y -> look-ahead ring; at the end of a frame the block's maximum and
//                  its suffix maxima
//   4,5 peak       (alternate chunks) delayed sample + window maximum -> desired gain 0.2 / (peak + 1e-10)
//   6 gain         g -= max(att (g - d), rel (g - d))                                      (serial: 4 ops)
//   7,8 out        (alternate chunks) delayed sample * g -> int16 (as int32) -> PCM rows, transposed back
//   9 suffix       suffix maxima of every finished block (needed kb - 1 frames later, when the window's old end gets there)
//
// A frame dropped by the NaN guard (src/signal.cpp:266-271: nothing is pushed) leaves a lane idle for that frame; the
// last D inputs / averages such a lane will need again are parked in two small save buffers and replayed.
#pragma once
#include <cstdint>
#include "clients.cuh"
#include "fft_tma.cuh"

namespace b200 {

// Ampere-style asynchronous 4-byte copies global -> shared: no registers, arbitrary (transposing) destinations
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kT2NPeak = 4, kT2NOut = 4, kT2NSuf = 2;   // warps of the stages that have no loop-carried state: chunks
                                                         // (suffix: frames) are dealt round-robin
constexpr int kT2Warps = 16;
constexpr int kT2Threads = 32 * kT2Warps;
constexpr int kT2CH = 32;      // samples per chunk
constexpr int kT2Depth = 4;    // chunks a producer may run ahead of its consumer
constexpr int kT2Pitch = 33;   // ring row pitch in floats: conflict-free for lanes along clients AND along samples

struct Tail2State {          // device memory, [group] major, 32 clients innermost
    float *dcx, *dcm;        // [groups][D][32]   last D inputs / first-stage averages, oldest first
    float *sum;              // [groups][2][32]
    float *gain;             // [groups][32]
    int *since;              // [groups][32]      samples pushed since the AGC was reset (saturates at L)
    int *blk;                // [groups][32]      look-ahead ring block that the next frame fills
    float *ring, *suf;       // [groups][NB * h][32]  look-ahead samples and their in-block suffix maxima of |.|
    float *cmax;             // [groups][NB][32]  block maxima of |.|
    int *err;                // a bounded wait expired
    int NB;                  // blocks (frames) in the look-ahead ring
    int kb;                  // the window's old end lies kb blocks behind the frame being pushed
    int col0;                // ... at this column when the frame starts
    int nsx;                 // X / M ring length in samples: D + (depth + 1) * CH
    int dpow2;               // D is a power of two: sum / D is an exact multiply
    int pcm16;               // 1: PCM rows are int16 (two per 32-bit word), 0: int32 as AudioEncoder::process takes them
};

// Shared memory of one group of 32 clients: 161 KB with the reference's D = 32 (108 KB at depth 2, measured 5 % slower). (The kernel can run G = 2 groups per CTA,
// each its own pipeline; the host does not use it - see launch_tail2 in engine.cu for the measurement.)
__host__ __device__ inline size_t tail2_smem(int D) {
    const size_t nsx = (size_t)D + (kT2Depth + 1) * kT2CH;
    const size_t chunk = sizeof(float) * kT2CH * kT2Pitch;
    //     X, M rings                                Y R D G rings          out (2 x 2), peak (2 x 2), suffix (2) staging   barriers
    return 2 * sizeof(float) * nsx * kT2Pitch + 4 * kT2Depth * chunk + 2 * (kT2NOut + kT2NPeak + kT2NSuf) * chunk + 1024;
}

// slow path of a bounded mbarrier wait, out of line: the stages' hot loops stay small (ten warps run ten different loops,
// and the instruction caches are what they compete for)
__device__ __noinline__ bool t2_wait_slow(uint64_t *bar, unsigned parity, int *s_abort, int *err, int code) {
    unsigned long long t0 = 0;
    for (unsigned spins = 1;; spins++) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((spins & 63u) == 0) {
            if (*reinterpret_cast<volatile int *>(s_abort)) return false;
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) {
                if (atomicCAS(s_abort, 0, 1) == 0) {  // first stage of this CTA to give up; err is a mapped host word
                    *reinterpret_cast<volatile int *>(err) = code;
                    __threadfence_system();
                }
                return false;
            }
        }
    }
}

template <int G>
__global__ void __launch_bounds__(kT2Threads *G) client_tail2_kernel(const ClientArrays ca, const ClientLaunch cl, const Tail2State st,
                                                                      const int groups) {
    extern __shared__ __align__(16) unsigned char t2_smem_all[];
    const int sub = G > 1 ? (int)threadIdx.x / kT2Threads : 0;   // group of this warp inside the CTA
    unsigned char *t2_smem = t2_smem_all + (size_t)sub * tail2_smem(ca.D);
    const int h = ca.h, D = ca.D, L = ca.L, F = cl.nframes, NB = st.NB;
    const int nsx = st.nsx;
    constexpr int CH = kT2CH, DEPTH = kT2Depth, P = kT2Pitch, NR = kT2Depth * kT2CH;
    float *rX = reinterpret_cast<float *>(t2_smem);
    float *rM = rX + (size_t)nsx * P;
    float *rY = rM + (size_t)nsx * P;     // DC blocker output
    float *rR = rY + NR * P;               // running maximum of |y| inside the frame
    float *rD = rR + NR * P;               // desired gain
    float *rG = rD + NR * P;               // gain
    float *stgO = rG + NR * P;             // out stage: [2 warps][2 buffers][CH][P] delayed samples (front of the look-ahead
                                           // buffer) copied straight from the look-ahead ring, turned into int16 results in place
    float *stgP = stgO + 2 * kT2NOut * CH * P;   // peak stage: [warps][2 buffers][CH][P] suffix-maximum rows of the look-ahead ring
    float *stgS = stgP + 2 * kT2NPeak * CH * P;  // suffix stage: [warps][2 buffers][CH][P]
    uint64_t *bars = reinterpret_cast<uint64_t *>(stgS + 2 * kT2NSuf * CH * P);
    // edges: 0 X (load -> sum1, sum2)  1 M (sum1 -> sum2)  2 Y (sum2 -> block)  3 R (block -> peak)  4 D (peak -> gain)
    //        5 (unused)  6 G (gain -> out)  7 B (block -> suffix, one slot per FRAME)
    constexpr int BS = 4;  // barrier slots reserved per edge and direction
    auto full = [&](int e) { return bars + (2 * e) * BS; };
    auto empty = [&](int e) { return bars + (2 * e + 1) * BS; };
    __shared__ int s_abort_all[G];
    __shared__ unsigned char s_valid_all[G][64][32];   // [frame][lane]: the NaN guard's verdict (src/signal.cpp:266-271)
    int &s_abort = s_abort_all[sub];
    unsigned char(*s_valid)[32] = s_valid_all[sub];

    // Stage of this warp. Warp w issues from scheduler w % 4: each scheduler gets ONE of the four stages with a loop-carried
    // recurrence (sum1, sum2, gain, block - the critical path) plus warps that run the same code as one another, so that
    // the 6 KB L0 instruction cache of a scheduler holds three loops, not four or five.
    enum { R_LOAD = 0, R_SUM1 = 1, R_SUM2 = 2, R_BLOCK = 3, R_GAIN = 4, R_PEAK = 5, R_OUT = R_PEAK + kT2NPeak, R_SUF = R_OUT + kT2NOut,
           R_IDLE = R_SUF + kT2NSuf };
    static_assert(R_IDLE <= kT2Warps && kT2NPeak == 4 && kT2NOut == 4 && kT2NSuf == 2, "role table below");
    const int tid = (int)threadIdx.x - sub * kT2Threads, hw_warp = tid >> 5, lane = tid & 31;
    //                                   scheduler 0..3 | 0..3                        | 0..3                            | 0..3
    constexpr int kRole[kT2Warps] = {R_SUM1, R_SUM2, R_GAIN, R_BLOCK, R_PEAK, R_PEAK + 2, R_OUT, R_OUT + 2,
                                     R_PEAK + 1, R_PEAK + 3, R_OUT + 1, R_OUT + 3, R_LOAD, R_SUF, R_SUF + 1, R_IDLE};
    const int warp = kRole[hw_warp];
    const int grp = (int)blockIdx.x * G + sub;
    const int slot = grp * 32 + lane;
    const bool in_range = slot < ca.max_clients;
    const int flags = in_range ? ca.slots[slot].flags : 0;
    const bool active = (flags & CF_ACTIVE) != 0;
    const bool reset_all = (flags & CF_RESET_ALL) != 0;
    const bool reset_agc = (flags & (CF_RESET_ALL | CF_RESET_AGC)) != 0;
    if (tid == 0) {
        const int ncons[8] = {2, 1, 1, 1, 1, 1, 1, 1};
        for (int e = 0; e < 8; e++)
            for (int s = 0; s < BS; s++) {
                mbar_init(full(e) + s, 1);
                mbar_init(empty(e) + s, ncons[e]);
            }
        s_abort = 0;
        fence_barrier_init();
    }
    for (int i = tid; i < F * 32; i += kT2Threads) {
        const int f = i >> 5, c = i & 31, sl = grp * 32 + c;
        s_valid[f][c] = (sl < ca.max_clients && (ca.slots[sl].flags & CF_ACTIVE)) ? ca.valid_a[(size_t)f * ca.max_clients + sl] : 0;
    }
    __syncthreads();
    if (grp >= groups) return;  // (odd number of groups: the second half of the last CTA)

    // bounded waits: a stage that never gets its chunk is a bug in this file - flag it and leave instead of hanging
    auto wait_bar = [&](uint64_t *bar, unsigned parity) -> bool {
        if (mbar_try_wait(bar, parity)) return true;
        return t2_wait_slow(bar, parity, &s_abort, st.err, 10000 + warp * 100 + (int)(bar - bars));  // (stage, edge / direction / slot)
    };
    // profiling aid (cl.prof != nullptr): cycles the warps of CTA 0 spend waiting for input / for ring space, and in total
    const bool prof = cl.prof != nullptr && grp == 0 && lane == 0;
    long long t_full = 0, t_empty = 0;
    const long long t_start = prof ? clock64() : 0;
#define T2_WAIT_FULL(e, q)                                                            \
    do {                                                                              \
        const long long t_a = prof ? clock64() : 0;                                   \
        if (!wait_bar(full(e) + (q) % DEPTH, ((q) / DEPTH) & 1)) return;              \
        if (prof) t_full += clock64() - t_a;                                          \
    } while (0)
#define T2_WAIT_EMPTY(e, q)                                                           \
    do {                                                                              \
        const long long t_a = prof ? clock64() : 0;                                   \
        if (!wait_bar(empty(e) + (q) % DEPTH, (((q) / DEPTH) & 1) ^ 1)) return;       \
        if (prof) t_empty += clock64() - t_a;                                         \
    } while (0)
#define T2_WAITC_FULL(e, c)                                                           \
    do {                                                                              \
        const long long t_a = prof ? clock64() : 0;                                   \
        if (!wait_bar(full(e) + (c).slot, (c).par)) return;                           \
        if (prof) t_full += clock64() - t_a;                                          \
    } while (0)
#define T2_WAITC_EMPTY(e, c)                                                          \
    do {                                                                              \
        const long long t_a = prof ? clock64() : 0;                                   \
        if (!wait_bar(empty(e) + (c).slot, (c).par ^ 1)) return;                      \
        if (prof) t_empty += clock64() - t_a;                                         \
    } while (0)
#define T2_SIGNALC(barp, c)                        \
    do {                                            \
        __syncwarp();                               \
        if (lane == 0) mbar_arrive((barp) + (c).slot); \
    } while (0)
#define T2_SIGNAL(barp, q)                         \
    do {                                            \
        __syncwarp();                               \
        if (lane == 0) mbar_arrive((barp) + (q) % DEPTH); \
    } while (0)

    const size_t g32 = (size_t)grp * 32 + lane;                      // [groups][32] scalars
    const float Df = (float)D, invD = 1.0f / Df;
    const bool dpow2 = st.dpow2 != 0;
    auto avg = [&](float s) { return dpow2 ? __fmul_rn(s, invD) : __fdiv_rn(s, Df); };
    auto wrapx = [&](int r) { return r >= nsx ? r - nsx : r; };
    // per-frame validity of this lane (frames dropped by the NaN guard leave every state untouched)
    auto valid_of = [&](int f) -> bool { return s_valid[f][lane] != 0; };
    const int cpf = (h + CH - 1) / CH;   // chunks per frame
    const int nq = F * cpf;
    constexpr int U = 8;                 // samples per register batch: loads of a batch are in flight together
    // chunk q = (frame f, first sample a, length len), advanced without divisions
    struct ChunkIt {
        int f, a, len, slot, par;  // ring slot q % DEPTH and its phase parity (q / DEPTH) & 1
    };
    auto chunk_first = [&]() { return ChunkIt{0, 0, min(CH, h), 0, 0}; };
    auto chunk_next = [&](ChunkIt &c) {
        c.a += CH;
        if (c.a >= h) {
            c.a = 0;
            c.f++;
        }
        c.len = min(CH, h - c.a);
        if (++c.slot == DEPTH) {
            c.slot = 0;
            c.par ^= 1;
        }
    };

    // the same walk for the stages whose warps take every N-th chunk: chunk number, frame, chunk in frame, and the two
    // per-frame counters that advance with every frame the NaN guard let through
    struct Walk {
        int q, f, ci, blk, since;  // blk: look-ahead block that frame f fills; since: samples pushed before frame f (saturates at L)
    };
    auto walk_first = [&](int blk0, int since0) { return Walk{0, 0, 0, blk0, since0}; };
    auto walk = [&](Walk &w, int n) {  // n chunks on (the caller keeps w.q + n < nq)
        for (int i = 0; i < n; i++) {
            w.q++;
            if (++w.ci == cpf) {
                w.ci = 0;
                if (valid_of(w.f)) {
                    if (++w.blk == NB) w.blk = 0;
                    w.since = min(w.since + h, L);
                }
                w.f++;
            }
        }
    };

    // Code size and instruction count matter here: every stage is ONE instruction stream, so its time is its instruction
    // count times the issue latency of dependent instructions. Loops stay rolled (register batches of eight samples,
    // addressed by one pointer plus compile-time offsets; the rare batch that straddles the end of a ring or the end of a
    // chunk takes a generic path), and global memory is read with asynchronous copies.
    // ring element (idx, lane) of a [samples][P] ring
#define T2_AT(ring, idx) ((ring) + (idx) * P + lane)
    if (warp == R_LOAD) {
        // ================= load: audio rows -> X ring (transposed), DC input history =================
        // history: ring positions nsx - D .. nsx - 1 precede position 0 of the common timeline
        for (int i = 0; i < D; i++) {
            const float v = (active && !reset_all) ? st.dcx[((size_t)grp * D + i) * 32 + lane] : 0.f;
            rX[(nsx - D + i) * P + lane] = v;
        }
        __syncwarp();
        // one client row at a time, lanes along the samples of the chunk (coalesced), copied straight to their transposed
        // place; up to DEPTH - 1 chunks are requested ahead of the one being handed over
        int pos_issue = 0, q_issue = 0;
        const int nrows = min(32, ca.max_clients - grp * 32);
        ChunkIt ci = chunk_first();
        auto issue = [&]() -> bool {
            if (!wait_bar(empty(0) + ci.slot, ci.par ^ 1)) return false;
            // (rows of dropped frames and of inactive slots are copied too: nothing reads them - a dropped frame's lane only
            // ever looks at the last D positions, which the replay below rewrites - and one unpredicated copy per row with
            // two pointer bumps is a fifth of the instructions of the selective loop)
            const float *src = ca.audio_pre + ((size_t)ci.f * ca.max_clients + grp * 32) * h + ci.a + lane;
            float *dst = rX + wrapx(pos_issue + lane) * P;
            if (lane < ci.len) {
#pragma unroll 8
                for (int c = 0; c < nrows; c++, src += h, dst++) cp_async4(dst, src);
            }
            cp_async_commit();
            pos_issue = wrapx(pos_issue + ci.len);
            q_issue++;
            chunk_next(ci);
            return true;
        };
        int pos = 0;
        ChunkIt c = chunk_first();
        for (int q = 0; q < nq; q++, chunk_next(c)) {
            const bool v = valid_of(c.f);
            while (q_issue < nq && q_issue < q + DEPTH - 1)
                if (!issue()) return;
            // groups complete in order: chunk q has landed when at most q_issue - q - 1 groups are still pending
            if (q_issue - q - 1 >= 2) cp_async_wait<2>();
            else if (q_issue - q - 1 == 1) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncwarp();
            // A dropped frame pushes nothing: its lane parks the last D inputs (still in the ring when the frame starts)
            // and replays them into the last D positions of the frame, where the next frame will look for them.
            if (!v && active) {
                float *park = st.dcx + (size_t)grp * D * 32 + lane;  // (read at the start, rewritten at the end: free in between)
                if (c.a == 0)
                    for (int i = 0; i < D; i++) park[i * 32] = rX[wrapx(pos + nsx - D + i) * P + lane];
                for (int j = max(c.a, h - D); j < c.a + c.len; j++) rX[wrapx(pos + (j - c.a)) * P + lane] = park[(j - (h - D)) * 32];
            }
            T2_SIGNALC(full(0), c);
            pos = wrapx(pos + c.len);
        }
        // the last D inputs of the batch are the history of the next launch
        __syncwarp();
        if (active)
            for (int i = 0; i < D; i++) st.dcx[((size_t)grp * D + i) * 32 + lane] = rX[wrapx(pos + nsx - D + i) * P + lane];
    } else if (warp == R_SUM1) {
        // ================= sum1: first running sum of the DC blocker, src/utils.h:80-85 =================
        float s1 = (active && !reset_all) ? st.sum[((size_t)grp * 2 + 0) * 32 + lane] : 0.f;
        for (int i = 0; i < D; i++) {
            const float v = (active && !reset_all) ? st.dcm[((size_t)grp * D + i) * 32 + lane] : 0.f;
            rM[(nsx - D + i) * P + lane] = v;
        }
        __syncwarp();
        int pos = 0;
        ChunkIt c = chunk_first();
        for (int q = 0; q < nq; q++, chunk_next(c)) {
            const int len = c.len;
            const bool v = valid_of(c.f);
            T2_WAITC_FULL(0, c);
            T2_WAITC_EMPTY(1, c);
            if (v) {
                int rn = pos, ro = wrapx(pos + nsx - D);
                for (int j0 = 0; j0 < len; j0 += U) {
                    if (j0 + U <= len && rn + U <= nsx && ro + U <= nsx) {  // the common case: plain pointers, no predicates
                        const float *po = T2_AT(rX, ro), *pn = T2_AT(rX, rn);
                        float *pm = T2_AT(rM, rn);
                        float xo[U], xn[U];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            xo[u] = po[u * P];
                            xn[u] = pn[u * P];
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            s1 = __fadd_rn(__fadd_rn(s1, -xo[u]), xn[u]);  // sum -= q.back(); sum += val
                            pm[u * P] = avg(s1);
                        }
                    } else {
                        for (int u = 0; u < U && j0 + u < len; u++) {
                            s1 = __fadd_rn(__fadd_rn(s1, -*T2_AT(rX, wrapx(ro + u))), *T2_AT(rX, wrapx(rn + u)));
                            *T2_AT(rM, wrapx(rn + u)) = avg(s1);
                        }
                    }
                    rn = wrapx(rn + U);
                    ro = wrapx(ro + U);
                }
            } else if (active) {  // dropped frame: park / replay the last D averages (see the load stage)
                float *park = st.dcm + (size_t)grp * D * 32 + lane;
                if (c.a == 0)
                    for (int i = 0; i < D; i++) park[i * 32] = rM[wrapx(pos + nsx - D + i) * P + lane];
                for (int j = max(c.a, h - D); j < c.a + len; j++) rM[wrapx(pos + (j - c.a)) * P + lane] = park[(j - (h - D)) * 32];
            }
            T2_SIGNALC(full(1), c);
            T2_SIGNALC(empty(0), c);
            pos = wrapx(pos + len);
        }
        __syncwarp();
        if (active) {
            st.sum[((size_t)grp * 2 + 0) * 32 + lane] = s1;
            for (int i = 0; i < D; i++) st.dcm[((size_t)grp * D + i) * 32 + lane] = rM[wrapx(pos + nsx - D + i) * P + lane];
        }
    } else if (warp == R_SUM2) {
        // ================= sum2: second running sum; y = x[delayed] - ma2, src/utils.h:145-149 =================
        float s2 = (active && !reset_all) ? st.sum[((size_t)grp * 2 + 1) * 32 + lane] : 0.f;
        int pos = 0;
        ChunkIt c = chunk_first();
        for (int q = 0; q < nq; q++, chunk_next(c)) {
            const int len = c.len;
            const bool v = valid_of(c.f);
            T2_WAITC_FULL(1, c);  // (the X chunk is complete as well: sum1 waited for it)
            T2_WAITC_EMPTY(2, c);
            if (v) {
                int rn = pos, ro = wrapx(pos + nsx - D), rd = wrapx(pos + nsx - D + 1);  // buffer[delay - 1]: the input of delay - 1 samples ago
                float *yp = T2_AT(rY, c.slot * CH);
                for (int j0 = 0; j0 < len; j0 += U, yp += U * P) {
                    if (j0 + U <= len && rn + U <= nsx && ro + U <= nsx && rd + U <= nsx) {
                        const float *po = T2_AT(rM, ro), *pn = T2_AT(rM, rn), *pd = T2_AT(rX, rd);
                        float mo[U], mn[U], xd[U];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            mo[u] = po[u * P];
                            mn[u] = pn[u * P];
                            xd[u] = pd[u * P];
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            s2 = __fadd_rn(__fadd_rn(s2, -mo[u]), mn[u]);
                            yp[u * P] = __fsub_rn(xd[u], avg(s2));
                        }
                    } else {
                        for (int u = 0; u < U && j0 + u < len; u++) {
                            s2 = __fadd_rn(__fadd_rn(s2, -*T2_AT(rM, wrapx(ro + u))), *T2_AT(rM, wrapx(rn + u)));
                            yp[u * P] = __fsub_rn(*T2_AT(rX, wrapx(rd + u)), avg(s2));
                        }
                    }
                    rn = wrapx(rn + U);
                    ro = wrapx(ro + U);
                    rd = wrapx(rd + U);
                }
            }
            T2_SIGNALC(full(2), c);
            T2_SIGNALC(empty(1), c);
            T2_SIGNALC(empty(0), c);
            pos = wrapx(pos + len);
        }
        if (active) st.sum[((size_t)grp * 2 + 1) * 32 + lane] = s2;
    } else if (warp == R_BLOCK) {
        // ================= block: running maximum, look-ahead ring, block maximum =================
        int blk = active ? st.blk[g32] : 0;
        float *ring = st.ring + (size_t)grp * NB * h * 32 + lane;
        float run = 0.f;
        ChunkIt c = chunk_first();
        for (int q = 0; q < nq; q++, chunk_next(c)) {
            const int f = c.f, a = c.a, len = c.len;
            const bool v = valid_of(f);
            if (a == 0) {
                run = 0.f;
                T2_WAIT_EMPTY(7, f);  // (the suffix stage is done with the frame that used this slot)
            }
            T2_WAITC_FULL(2, c);
            T2_WAITC_EMPTY(3, c);
            if (v) {
                float *rowp = ring + ((size_t)blk * h + a) * 32;
                const float *yp = T2_AT(rY, c.slot * CH);
                float *rp = T2_AT(rR, c.slot * CH);
                for (int j0 = 0; j0 < len; j0 += U, yp += U * P, rp += U * P, rowp += U * 32) {
                    if (j0 + U <= len) {
                        float y[U];
#pragma unroll
                        for (int u = 0; u < U; u++) y[u] = yp[u * P];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            rowp[u * 32] = y[u];
                            run = fmaxf(run, fabsf(y[u]));
                            rp[u * P] = run;
                        }
                    } else {
                        for (int u = 0; j0 + u < len; u++) {
                            const float y = yp[u * P];
                            rowp[u * 32] = y;
                            run = fmaxf(run, fabsf(y));
                            rp[u * P] = run;
                        }
                    }
                }
                // the block maximum is read by the peak stage as soon as the NEXT frame starts: publish it before the
                // frame's last chunk is handed over
                if (a + len == h) st.cmax[((size_t)grp * NB + blk) * 32 + lane] = run;
            }
            T2_SIGNALC(full(3), c);
            T2_SIGNALC(empty(2), c);
            if (a + len == h) {
                T2_SIGNAL(full(7), f);
                if (v && ++blk == NB) blk = 0;
            }
        }
        if (active) st.blk[g32] = blk;
    } else if (warp >= R_PEAK && warp < R_PEAK + kT2NPeak) {
        // ================= peak: window maximum -> desired gain (audioprocessing.cpp:41-52); chunks q = me (mod NP) ========
        constexpr int NP = kT2NPeak;
        const int me = warp - R_PEAK;
        const float *suf = st.suf + (size_t)grp * NB * h * 32 + lane;
        const float *cmx = st.cmax + (size_t)grp * NB * 32 + lane;
        const int kb = st.kb, col0 = st.col0;
        float *stg = stgP + me * 2 * CH * P + lane;   // [buffer][CH][P]
        // rows of the suffix maxima for chunk w: they were written at least kb - 1 frames ago, so they are requested one own
        // chunk ahead, independent of the frame being pushed
        auto request = [&](const Walk &w, int buf) {
            if (valid_of(w.f)) {
                const int a = w.ci * CH, len = min(CH, h - a);
                int c0 = w.blk - kb;
                if (c0 < 0) c0 += NB;
                const int c1 = (c0 + 1 == NB) ? 0 : c0 + 1;
                float *d = stg + buf * CH * P;
                // the chunk's window ends walk through block c0 from column col0 + a, then through block c1 from column 0
                const int i0 = col0 + a, n0 = max(0, min(len, h - i0));
                const float *g0 = suf + ((size_t)c0 * h + i0) * 32, *g1 = suf + ((size_t)c1 * h + (i0 + n0 - h)) * 32;
#pragma unroll 8
                for (int j = 0; j < n0; j++) cp_async4(d + j * P, g0 + (size_t)j * 32);
#pragma unroll 8
                for (int j = n0; j < len; j++) cp_async4(d + j * P, g1 + (size_t)(j - n0) * 32);
            }
            cp_async_commit();
        };
        if (me < nq) {
            Walk cur = walk_first(active ? st.blk[g32] : 0, 0), nxt = cur;
            walk(cur, me);
            request(cur, 0);
            float m0 = 0.f, m1 = 0.f;  // maxima of the whole blocks strictly between the walking block and this frame
            int f_m = -1;
            for (int k = 0;; k++) {
                const bool more = cur.q + NP < nq;
                if (more) {
                    nxt = cur;
                    walk(nxt, NP);
                    request(nxt, (k + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                const int q = cur.q, f = cur.f, a = cur.ci * CH, len = min(CH, h - a);
                const bool v = valid_of(f);
                T2_WAIT_FULL(3, q);  // (also orders this after the previous frame's block maximum)
                if (f != f_m) {      // first own chunk of the frame: the block maxima in between
                    f_m = f;
                    m0 = m1 = 0.f;
                    if (v) {
                        int c = cur.blk - kb + 1;
                        if (c < 0) c += NB;
                        for (int i = 1; i < kb; i++) {
                            const float x = cmx[(size_t)c * 32];
                            m0 = fmaxf(m0, x);
                            if (i >= 2) m1 = fmaxf(m1, x);
                            if (++c == NB) c = 0;
                        }
                    }
                }
                T2_WAIT_EMPTY(4, q);
                if (v) {
                    const float *d = stg + (k & 1) * CH * P;
                    const float *rp = T2_AT(rR, (q % DEPTH) * CH);
                    float *dp = T2_AT(rD, (q % DEPTH) * CH);
                    const int nfirst = max(0, min(len, h - (col0 + a)));  // samples whose window end is still in block c0
                    for (int j0 = 0; j0 < len; j0 += U, d += U * P, rp += U * P, dp += U * P) {
                        if (j0 + U <= len && (j0 + U <= nfirst || j0 >= nfirst)) {
                            const float mid = (j0 < nfirst) ? m0 : m1;
                            float pk[U];
#pragma unroll
                            for (int u = 0; u < U; u++) pk[u] = __fadd_rn(fmaxf(fmaxf(d[u * P], mid), rp[u * P]), 1e-10f);
#pragma unroll
                            for (int u = 0; u < U; u++) dp[u * P] = __fdiv_rn(ca.desired, pk[u]);
                        } else {
                            for (int u = 0; u < U && j0 + u < len; u++) {
                                const float mid = (j0 + u < nfirst) ? m0 : m1;
                                dp[u * P] = __fdiv_rn(ca.desired, __fadd_rn(fmaxf(fmaxf(d[u * P], mid), rp[u * P]), 1e-10f));
                            }
                        }
                    }
                }
                T2_SIGNAL(full(4), q);
                T2_SIGNAL(empty(3), q);
                if (!more) break;
                cur = nxt;
            }
        }
    } else if (warp == R_GAIN) {
        // ================= gain: attack / release recurrence, audioprocessing.cpp:54-63 =================
        float gain = (active && !reset_agc) ? st.gain[g32] : 0.f;
        int since = (active && !reset_agc) ? st.since[g32] : 0;
        const float att = ca.attack, rel = ca.release;
        ChunkIt c = chunk_first();
        for (int q = 0; q < nq; q++, chunk_next(c)) {
            const int a = c.a, len = c.len;
            const bool v = valid_of(c.f);
            T2_WAITC_FULL(4, c);
            T2_WAITC_EMPTY(6, c);
            const float *dp = T2_AT(rD, c.slot * CH);
            float *gp = T2_AT(rG, c.slot * CH);
            // outputs stay 0 (and the gain untouched) until the look-ahead buffer is full: sample index since + a + j
            // must reach L - 1 (audioprocessing.cpp:45,64-66)
            const int first = v ? max(0, min(len, L - 1 - since - a)) : len;
            const bool steady = __all_sync(0xffffffffu, first == 0);
            for (int j0 = 0; j0 < len; j0 += U, dp += U * P, gp += U * P) {
                if (steady && j0 + U <= len) {
                    float d[U];
#pragma unroll
                    for (int u = 0; u < U; u++) d[u] = dp[u * P];
                    // t = gain - desired; both arms are gain - c*t and attack >= release > 0 picks the arm by max()
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float t = __fsub_rn(gain, d[u]);
                        gain = __fsub_rn(gain, fmaxf(__fmul_rn(att, t), __fmul_rn(rel, t)));
                        gp[u * P] = gain;
                    }
                } else {
                    for (int u = 0; u < U && j0 + u < len; u++) {
                        if (j0 + u >= first) {
                            const float t = __fsub_rn(gain, dp[u * P]);
                            gain = __fsub_rn(gain, fmaxf(__fmul_rn(att, t), __fmul_rn(rel, t)));
                        }
                        gp[u * P] = gain;
                    }
                }
            }
            T2_SIGNALC(full(6), c);
            T2_SIGNALC(empty(4), c);
            if (v && a + len == h) since = min(since + h, L);
        }
        if (active) {
            st.gain[g32] = gain;
            st.since[g32] = since;
        }
    } else if (warp >= R_OUT && warp < R_OUT + kT2NOut) {
        // ================= out: delayed sample * gain -> int16 (dsp.cpp:152-165 with mult = 65536 / 4), transposed store ====
        // chunks q = me (mod NO)
        constexpr int NO = kT2NOut;
        const int me = warp - R_OUT;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        const float *ring = st.ring + (size_t)grp * NB * h * 32 + lane;
        const int kb = st.kb, col0 = st.col0;
        float *stg = stgO + me * 2 * CH * P;   // [buffer][CH][P]
        if (me == 0 && active)
            for (int f = 0; f < F; f++) ca.valid[(size_t)f * ca.max_clients + slot] = valid_of(f) ? 1 : 0;
        // The delayed samples of a chunk are rows of the look-ahead ring written at least kb - 1 frames ago: they are
        // requested one own chunk ahead, independent of every other stage.
        auto request = [&](const Walk &w, int buf) {
            if (valid_of(w.f)) {
                const int a = w.ci * CH, len = min(CH, h - a);
                int c0 = w.blk - kb;
                if (c0 < 0) c0 += NB;
                const int c1 = (c0 + 1 == NB) ? 0 : c0 + 1;
                float *d = stg + buf * CH * P + lane;
                const int i0 = col0 + a, n0 = max(0, min(len, h - i0));
                const float *g0 = ring + ((size_t)c0 * h + i0) * 32, *g1 = ring + ((size_t)c1 * h + (i0 + n0 - h)) * 32;
#pragma unroll 8
                for (int j = 0; j < n0; j++) cp_async4(d + j * P, g0 + (size_t)j * 32);
#pragma unroll 8
                for (int j = n0; j < len; j++) cp_async4(d + j * P, g1 + (size_t)(j - n0) * 32);
            }
            cp_async_commit();
        };
        if (me < nq) {
            Walk cur = walk_first(active ? st.blk[g32] : 0, (active && !reset_agc) ? st.since[g32] : 0), nxt = cur;
            walk(cur, me);
            request(cur, 0);
            for (int k = 0;; k++) {
                const bool more = cur.q + NO < nq;
                if (more) {
                    nxt = cur;
                    walk(nxt, NO);
                    request(nxt, (k + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                const int q = cur.q, f = cur.f, a = cur.ci * CH, len = min(CH, h - a);
                const bool v = valid_of(f);
                T2_WAIT_FULL(6, q);
                int *tile = reinterpret_cast<int *>(stg + (k & 1) * CH * P);  // results replace the delayed samples in place
                const float *op = reinterpret_cast<const float *>(tile) + lane;
                const float *gp = T2_AT(rG, (q % DEPTH) * CH);
                int *tp = tile + lane;
                // outputs stay 0 until the look-ahead buffer is full (audioprocessing.cpp:45,64-66)
                const int first = v ? max(0, min(len, L - 1 - cur.since - a)) : len;
                for (int j0 = 0; j0 < len; j0 += U, op += U * P, gp += U * P, tp += U * P) {
                    if (v && first == 0 && j0 + U <= len) {
                        float o[U], gg[U];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            o[u] = op[u * P];
                            gg[u] = gp[u * P];
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const float t = __fadd_rn(__fmul_rn(__fmul_rn(o[u], gg[u]), 16384.f), 32768.5f);
                            tp[u * P] = max(min(__float2int_rz(t) - 32768, 32767), -32768);
                        }
                    } else {
                        for (int u = 0; u < U && j0 + u < len; u++) {
                            int r = 0;
                            if (v && j0 + u >= first) {
                                const float t = __fadd_rn(__fmul_rn(__fmul_rn(op[u * P], gp[u * P]), 16384.f), 32768.5f);
                                r = max(min(__float2int_rz(t) - 32768, 32767), -32768);
                            }
                            tp[u * P] = r;
                        }
                    }
                }
                T2_SIGNAL(empty(6), q);
                __syncwarp();
                // one client row at a time, lanes along the samples of the chunk
                if (!st.pcm16) {
                    int *prow = ca.pcm + ((size_t)f * ca.max_clients + grp * 32) * h + a + lane;
                    const int *tr = tile + lane * P;
                    if (lane < len) {
                        if (amask == 0xffffffffu) {  // (the usual case: no predicate, two pointer bumps per row)
#pragma unroll 8
                            for (int c = 0; c < 32; c++, prow += h, tr++) *prow = *tr;
                        } else {
#pragma unroll 8
                            for (int c = 0; c < 32; c++)
                                if ((amask >> c) & 1u) prow[(size_t)c * h] = tr[c];
                        }
                    }
                } else {
                    // int16 rows: h int16 = h / 2 words (a is a multiple of 32 and h is even, so every row offset is even)
                    unsigned *p16 =
                        reinterpret_cast<unsigned *>(ca.pcm) + ((((size_t)f * ca.max_clients + grp * 32) * h + a) >> 1) + lane;
                    if (2 * lane < len) {
                        const int *t0 = tile + (2 * lane) * P, *t1 = t0 + ((2 * lane + 1 < len) ? P : 0);
                        const unsigned keep_hi = (2 * lane + 1 < len) ? 0xFFFFFFFFu : 0u;
                        const int hw = h >> 1;
                        if (amask == 0xffffffffu) {
#pragma unroll 8
                            for (int c = 0; c < 32; c++, p16 += hw, t0++, t1++)
                                *p16 = ((unsigned)*t0 & 0xFFFFu) | (((unsigned)*t1 << 16) & (keep_hi & 0xFFFF0000u));
                        } else {
#pragma unroll 8
                            for (int c = 0; c < 32; c++) {
                                if (!((amask >> c) & 1u)) continue;
                                p16[(size_t)c * hw] = ((unsigned)t0[c] & 0xFFFFu) | (((unsigned)t1[c] << 16) & (keep_hi & 0xFFFF0000u));
                            }
                        }
                    }
                }
                __syncwarp();
                if (!more) break;
                cur = nxt;
            }
        }
    } else if (warp >= R_SUF && warp < R_SUF + kT2NSuf) {
        // ================= suffix: in-block suffix maxima of |y| for every finished block; frames f = me (mod NS) =========
        constexpr int NS = kT2NSuf;
        const int me = warp - R_SUF;
        int blk = active ? st.blk[g32] : 0;
        const float *ring = st.ring + (size_t)grp * NB * h * 32 + lane;
        float *suf = st.suf + (size_t)grp * NB * h * 32 + lane;
        float *stg = stgS + me * 2 * CH * P + lane;   // [buffer][CH][P]
        for (int f = 0; f < F; f++) {
            const bool v = valid_of(f);
            if (f % NS == me) {
                T2_WAIT_FULL(7, f);
                if (v) {
                    const float *rowp = ring + (size_t)blk * h * 32;
                    float *sp = suf + (size_t)blk * h * 32;
                    // backwards in batches of CH rows; batch b + 1 is requested while batch b is scanned
                    auto issue = [&](int b) {
                        const int j1 = h - b * CH, j0 = max(0, j1 - CH);
                        float *d = stg + (b & 1) * CH * P;
                        const float *g = rowp + (size_t)j0 * 32;
#pragma unroll 4
                        for (int j = 0; j < j1 - j0; j++) cp_async4(d + j * P, g + (size_t)j * 32);
                        cp_async_commit();
                    };
                    issue(0);
                    float sm = 0.f;
                    for (int b = 0; b < cpf; b++) {
                        const int j1 = h - b * CH, j0 = max(0, j1 - CH);
                        if (b + 1 < cpf) {
                            issue(b + 1);
                            cp_async_wait<1>();
                        } else {
                            cp_async_wait<0>();
                        }
                        const float *d = stg + (b & 1) * CH * P;
                        float *so = sp + (size_t)j0 * 32;
#pragma unroll 4
                        for (int j = j1 - j0 - 1; j >= 0; j--) {
                            sm = fmaxf(sm, fabsf(d[j * P]));
                            so[(size_t)j * 32] = sm;
                        }
                    }
                }
                T2_SIGNAL(empty(7), f);
            }
            if (v && ++blk == NB) blk = 0;
        }
    }
#undef T2_AT
    if (prof) {
        cl.prof[warp * 3 + 0] += t_full;
        cl.prof[warp * 3 + 1] += t_empty;
        cl.prof[warp * 3 + 2] += clock64() - t_start;
    }
#undef T2_WAIT_FULL
#undef T2_WAIT_EMPTY
#undef T2_SIGNAL
#undef T2_WAITC_FULL
#undef T2_WAITC_EMPTY
#undef T2_SIGNALC
}

}  // namespace b200
