// Batched per-client demodulation: replaces N x AudioClient::send_audio (reference
// src/signal.cpp:102-298) plus the slice index math of signal_loop (src/websocket.cpp:156-185).
//
//   client_demod_kernel : one warp per client. Gather the client's bin slice from the spectrum,
//                         place it as the reference does per mode, inverse FFT of audio_fft_size
//                         points (mixed radix Stockham in shared memory), parity sign flip,
//                         overlap-add with device-resident state, USB/LSB/AM/FM demodulation, NaN guard.
//   client_tail_kernel  : one LANE per client for the strictly sequential tails - DC blocker
//                         (src/utils.h:139-169), look-ahead AGC (src/utils/audioprocessing.cpp:17-68),
//                         float -> int16 (src/utils/dsp.cpp:152-165) - bit-exact float op order.
#pragma once
#include <cstdint>
#include "regfft.cuh"

namespace b200 {

enum { MODE_USB = 0, MODE_LSB = 1, MODE_AM = 2, MODE_FM = 3 };
enum { CF_ACTIVE = 1, CF_RESET_AGC = 2, CF_RESET_ALL = 4 };
constexpr int kMaxStages = 24;

struct ClientSlot {  // host-written, one per slot
    int l, r;
    int m_floor;  // floor(audio_mid)
    int mode;
    int flags;
    int pad[3];
};

struct ClientArrays {
    // geometry
    int n, h;            // audio_fft_size, n/2
    int D;               // DC blocker delay
    int L;               // AGC look-ahead samples
    int NC;              // AGC ring rows (chunks of h samples)
    int max_clients;
    float attack, release, desired;
    int nstages;
    int radix[kMaxStages];
    unsigned mul_s[kMaxStages];      // floor(2^32 / s) + 1 for the stride s of each stage (t / s = umulhi(t, mul))
    unsigned mul_items[kMaxStages];  // same for items = n / radix
    unsigned mul_n;                  // same for n
    const float2 *Wn;    // exp(+2*pi*i*k/n), k < n
    // per-slot parameters and state (device)
    ClientSlot *slots;
    float *real_prev;    // [slot][h]
    float *real_hi;      // [slot][h]   only meaningful while diverged (NaN-dropped SSB frame)
    int *hi_diverged;    // [slot]
    float2 *bb_hi;       // [slot][h]
    float2 *bb_last;     // [slot]
    float *dc_x;         // [slot][D]  last D inputs, oldest first
    float *dc_m;         // [slot][D]  last D first-stage averages, oldest first
    float *dc_sum;       // [slot][2]
    float *agc_ring;     // [slot][NC][h]
    float *agc_cmax;     // [slot][NC]
    float *agc_gain;     // [slot]
    long long *agc_t0;   // [slot] samples pushed since reset
    // per-frame results (device)
    float *audio_pre;    // [frames][slot][h]
    unsigned char *valid_a;  // [frames][slot]
    float *pwr;          // [frames][slot]
    int *pcm;            // [frames][slot][h]
    unsigned char *valid;    // [frames][slot]
};

struct ClientLaunch {
    const float2 *spec;  // spectrum of frame 0
    size_t spec_stride;  // float2 between frames
    int nframes;
    unsigned long long frame_num0;
    size_t fft_size;     // reference fft_size
    int is_real;
    const int *order;    // active slots in (l, r) order
    int nactive;
    int fchunk;          // demod kernel: frames whose inverse FFTs are batched in shared memory at once
    int cpb;             // tail kernel: clients per block
    long long *prof;     // optional: per-phase SM clock totals of block 0 (profiling aid), 8 entries
    // frame-chunked demodulation (client_demod_warp_kernel): state is read from the `sin` copies and written to the arrays
    // of ClientArrays; clients whose batch must be replayed sequentially (a NaN-dropped frame) are flagged in `redo`
    const float *sin_real_prev;
    const float *sin_real_hi;
    const float2 *sin_bb_hi;
    const float2 *sin_bb_last;
    const int *sin_hi_diverged;
    unsigned char *redo;  // [max_clients]
    int chunk;            // frames per warp task
    int nchunks;
    int redo_only;        // client_demod_kernel: only the flagged clients, starting from the `sin` state
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// radix-R butterfly of an inverse (sign +1) DFT, in place; shared by the run-time and the compile-time stage code so that both
// round identically
template <int R> __device__ __forceinline__ void butterfly_r(float2 (&a)[R]) {
    if constexpr (R == 2) {
        float2 b0 = cadd(a[0], a[1]), b1 = csub(a[0], a[1]);
        a[0] = b0;
        a[1] = b1;
    } else if constexpr (R == 4) {
        float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        float2 t2 = cadd(a[1], a[3]), d = csub(a[1], a[3]);
        float2 t3 = make_float2(-d.y, d.x);  // +i * d
        a[0] = cadd(t0, t2);
        a[1] = cadd(t1, t3);
        a[2] = csub(t0, t2);
        a[3] = csub(t1, t3);
    } else if constexpr (R == 3) {
        const float hs = 0.86602540378443864676f;
        float2 sm = cadd(a[1], a[2]), d = csub(a[1], a[2]);
        float2 m = make_float2(a[0].x - 0.5f * sm.x, a[0].y - 0.5f * sm.y);
        float2 e = make_float2(-hs * d.y, hs * d.x);
        a[0] = cadd(a[0], sm);
        a[1] = cadd(m, e);
        a[2] = csub(m, e);
    } else if constexpr (R == 5) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
        float2 t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
        float2 m1 = make_float2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
        float2 m2 = make_float2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
        float2 v1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        float2 v2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        float2 n1 = make_float2(-v1.y, v1.x), n2 = make_float2(-v2.y, v2.x);  // i * v
        a[0] = make_float2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
        a[1] = cadd(m1, n1);
        a[4] = csub(m1, n1);
        a[2] = cadd(m2, n2);
        a[3] = csub(m2, n2);
    }
}

// One Stockham stage of radix R (sign +1) over `nf` independent transforms stored back to back:
//   x[t + j*items] -> y[q + s*(R*p + r)],  t = p*s + q,  items = n / R.
// The whole CTA shares the nf * items butterflies; divisions are multiply-high with host-made constants.
template <int R>
__device__ __forceinline__ void ifft_stage_fixed(const float2 *xb, float2 *yb, int n, int s, unsigned mul_s, unsigned mul_items,
                                                 int nf, const float2 *Wn, int tid, int nthr) {
    const int items = n / R;
    for (int T = tid; T < nf * items; T += nthr) {
        const int fr = __umulhi((unsigned)T, mul_items);
        const int t = T - fr * items;
        const int p = (s == 1) ? t : (int)__umulhi((unsigned)t, mul_s), q = t - p * s;
        const float2 *x = xb + fr * n;
        float2 *y = yb + fr * n;
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = x[t + j * items];
        butterfly_r<R>(a);
        const int ob = q + s * R * p;
        y[ob] = a[0];
        const int tw = p * s;  // W_ncur^(p r) = Wn[p*r*s]
#pragma unroll
        for (int r = 1; r < R; r++) y[ob + s * r] = cmul(a[r], __ldg(Wn + tw * r));
    }
}

// The same stage with everything but the lane known at compile time (one transform, one warp): loads and stores are one
// base address plus immediates, divisions by the stride become multiply-shifts, and the last stage (stride S = N / R, all
// twiddles 1) skips its multiplies.
template <int N, int R, int S>
__device__ __forceinline__ void ifft_stage_static(const float2 *x, float2 *y, const float2 *Wn, int lane) {
    constexpr int items = N / R;
    constexpr bool last = (S * R == N);
#pragma unroll
    for (int i = 0; i < (items + 31) / 32; i++) {
        const int t = lane + 32 * i;
        if ((i + 1) * 32 <= items || t < items) {
            const int p = (S == 1) ? t : t / S, q = t - p * S;
            float2 a[R];
#pragma unroll
            for (int j = 0; j < R; j++) a[j] = x[t + j * items];
            butterfly_r<R>(a);
            float2 *yo = y + (q + S * R * p);
            yo[0] = a[0];
            const int tw = p * S;  // W_ncur^(p r) = Wn[p*r*s]
#pragma unroll
            for (int r = 1; r < R; r++) yo[S * r] = last ? a[r] : cmul(a[r], __ldg(Wn + tw * r));
        }
    }
    __syncwarp();
}
template <int N> struct StaticRadix {  // the host's factorisation order (engine.cu: factorize): 4s, 2s, then odd primes
    static constexpr int first(int rem) { return rem % 4 == 0 ? 4 : rem % 2 == 0 ? 2 : rem % 3 == 0 ? 3 : rem % 5 == 0 ? 5 : 0; }
    static constexpr bool ok() {
        int rem = N;
        while (rem > 1) {
            const int r = first(rem);
            if (!r) return false;
            rem /= r;
        }
        return true;
    }
};
// runs the remaining stages (REM = N / S still to split) and returns the buffer that holds the result
template <int N, int S, int REM> __device__ __forceinline__ float2 *ifft_static(float2 *x, float2 *y, const float2 *Wn, int lane) {
    if constexpr (REM == 1) {
        return x;
    } else {
        constexpr int R = StaticRadix<N>::first(REM);
        static_assert(R != 0, "audio FFT size with a factor other than 2, 3, 5");
        ifft_stage_static<N, R, S>(x, y, Wn, lane);
        return ifft_static<N, S * R, REM / R>(y, x, Wn, lane);
    }
}

// generic (prime) radix: one thread per OUTPUT, R complex MACs each
__device__ __forceinline__ void ifft_stage_generic(const float2 *xb, float2 *yb, int n, int s, int R, unsigned mul_s,
                                                   unsigned mul_n, int nf, const float2 *Wn, int tid, int nthr) {
    const int items = n / R;
    const int step = n / R;  // W_R = Wn[n/R]
    for (int O = tid; O < nf * n; O += nthr) {
        const int fr = __umulhi((unsigned)O, mul_n);
        const int o = O - fr * n;
        const int t = o / R, r = o - t * R;
        const int p = (s == 1) ? t : (int)__umulhi((unsigned)t, mul_s), q = t - p * s;
        const float2 *x = xb + fr * n;
        float2 acc = x[t];
        int idx = 0;
        for (int j = 1; j < R; j++) {
            idx += r;
            if (idx >= R) idx -= R;
            const float2 w = __ldg(Wn + idx * step);
            const float2 v = x[t + j * items];
            acc.x += v.x * w.x - v.y * w.y;
            acc.y += v.x * w.y + v.y * w.x;
        }
        if (r) acc = cmul(acc, __ldg(Wn + p * s * r));
        yb[fr * n + q + s * (R * p + r)] = acc;
    }
}

// grid: one CTA per active client, TPB threads; dynamic smem 2 * fchunk * n float2. The inverse FFTs of up to
// `fchunk` consecutive frames are independent, so they are gathered and transformed together (full lanes, one
// barrier per stage for the whole chunk); only the overlap-add / demodulation walks the frames in order.
template <int TPB>
__global__ void __launch_bounds__(TPB) client_demod_kernel(const ClientArrays ca, const ClientLaunch cl) {
    extern __shared__ float2 smem_c[];
    __shared__ float s_pw[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ci = blockIdx.x;
    const int slot = cl.order[ci];
    const int n = ca.n, h = ca.h;
    const int FC = cl.fchunk;
    float2 *bufX = smem_c;
    float2 *bufY = bufX + (size_t)FC * n;
    const ClientSlot cs = ca.slots[slot];
    const size_t R = cl.is_real ? cl.fft_size / 2 : cl.fft_size;
    const size_t base_idx = cl.is_real ? 0 : cl.fft_size / 2 + 1;
    const size_t off = ((size_t)cs.l + base_idx) % R;  // src/websocket.cpp:182
    const int len = cs.r - cs.l;
    const int audio_m = cs.m_floor - cs.l;
    const int mode = cs.mode;

    float *real_prev = ca.real_prev + (size_t)slot * h;
    float *real_hi = ca.real_hi + (size_t)slot * h;
    float2 *bb_hi = ca.bb_hi + (size_t)slot * h;

    if (cl.redo_only) {
        // sequential replay of a client whose batch the frame-chunked kernel could not finish (a frame dropped by the NaN
        // guard changes what later frames overlap with): start again from the state the batch began with
        if (!cl.redo[slot]) return;
        for (int i = tid; i < h; i += TPB) {
            real_prev[i] = cl.sin_real_prev[(size_t)slot * h + i];
            real_hi[i] = cl.sin_real_hi[(size_t)slot * h + i];
            bb_hi[i] = cl.sin_bb_hi[(size_t)slot * h + i];
        }
        if (tid == 0) {
            ca.bb_last[slot] = cl.sin_bb_last[slot];
            ca.hi_diverged[slot] = cl.sin_hi_diverged[slot];
        }
        __syncthreads();
    }
    if (cs.flags & CF_RESET_ALL) {  // freshly opened slot: zeroed scratch as AudioClient's ctor (signal.cpp:38-52)
        for (int i = tid; i < h; i += TPB) {
            real_prev[i] = 0.f;
            real_hi[i] = 0.f;
            bb_hi[i] = make_float2(0.f, 0.f);
        }
        if (tid == 0) {
            ca.bb_last[slot] = make_float2(0.f, 0.f);
            ca.hi_diverged[slot] = 0;
        }
        __syncthreads();
    }

    for (int f0 = 0; f0 < cl.nframes; f0 += FC) {
        const int nf = min(FC, cl.nframes - f0);
        // ---- placement of the slice into the IFFT input (signal.cpp:125-198), all frames of the chunk at once:
        //      every input position looks up the bin that the reference copies there (or stays zero) ----
        for (int idx = tid; idx < nf * n; idx += TPB) {
            const int fr = __umulhi((unsigned)idx, ca.mul_n);
            const int kk = idx - fr * n;
            const float2 *buf = cl.spec + (size_t)(f0 + fr) * cl.spec_stride + off;
            float2 v = make_float2(0.f, 0.f);
            if (mode == MODE_USB || mode == MODE_LSB) {
                // c2r reads bins 0..n/2 of its input; the other half is the Hermitian mirror
                const int k2 = (kk <= n / 2) ? kk : n - kk;
                const int i = (mode == MODE_USB) ? (audio_m + k2) : (audio_m - k2);
                if (i >= 0 && i < len) {
                    v = buf[i];
                    if (k2 == 0 || k2 == n / 2) v.y = 0.f;  // c2r ignores Im of DC / Nyquist
                    else if (kk > n / 2) v.y = -v.y;
                }
            } else {
                const int d = (kk < n / 2) ? kk : kk - n;  // positive bins [0, n/2), negative [-n/2+1, -1]; kk = n/2 stays 0
                const int i = audio_m + d;
                if (kk != n / 2 && i >= 0 && i < len) v = buf[i];
            }
            bufX[idx] = v;
        }
        // slice power of every frame (signal.cpp:117-119): one warp per frame
        for (int fr = warp; fr < nf; fr += TPB / 32) {
            const float2 *buf = cl.spec + (size_t)(f0 + fr) * cl.spec_stride + off;
            float pw = 0.f;
            for (int i = lane; i < len; i += 32) {
                const float2 v = buf[i];
                pw += v.x * v.x + v.y * v.y;
            }
            pw = warp_sum(pw);
            if (lane == 0) s_pw[fr] = pw;
        }
        __syncthreads();
        // ---- inverse FFTs, unnormalised (signal.cpp:138,154,214) ----
        float2 *xb = bufX, *yb = bufY;
        int s = 1;
        for (int st = 0; st < ca.nstages; st++) {
            const int Rr = ca.radix[st];
            if (Rr == 4) ifft_stage_fixed<4>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], nf, ca.Wn, tid, TPB);
            else if (Rr == 2) ifft_stage_fixed<2>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], nf, ca.Wn, tid, TPB);
            else if (Rr == 3) ifft_stage_fixed<3>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], nf, ca.Wn, tid, TPB);
            else if (Rr == 5) ifft_stage_fixed<5>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], nf, ca.Wn, tid, TPB);
            else ifft_stage_generic(xb, yb, n, s, Rr, ca.mul_s[st], ca.mul_n, nf, ca.Wn, tid, TPB);
            s *= Rr;
            float2 *t = xb;
            xb = yb;
            yb = t;
            __syncthreads();
        }
        // ---- frames in order: parity sign flip, overlap-add with the carried state, demodulation ----
        for (int fr = 0; fr < nf; fr++) {
            const int f = f0 + fr;
            const float2 *x = xb + (size_t)fr * n;  // time-domain result of this frame
            float2 *y = yb + (size_t)fr * n;        // free scratch of this frame
            const unsigned long long frame_num = cl.frame_num0 + f;
            const int m_idx = cs.m_floor;
            const bool negate =
                (frame_num & 1ull) && (((m_idx % 2 == 0) && !cl.is_real) || ((m_idx % 2 == 1) && cl.is_real));
            const float sg = negate ? -1.f : 1.f;
            float *audio = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
            int nan_seen = 0;
            if (mode == MODE_USB || mode == MODE_LSB) {
                // signal.cpp:155-172: (LSB: time reverse), parity negate, overlap-add
                for (int t = tid; t < h; t += TPB) {
                    const float lo = (mode == MODE_USB) ? x[t].x : x[n - 1 - t].x;
                    const float o = __fadd_rn(sg * lo, real_prev[t]);
                    nan_seen |= (o != o);
                    audio[t] = o;
                }
                nan_seen = __syncthreads_or(nan_seen);
                float *dst = nan_seen ? real_hi : real_prev;  // signal.cpp:266-275: prev only advances on a sent frame
                for (int t = tid; t < h; t += TPB) {
                    const float hi = (mode == MODE_USB) ? x[h + t].x : x[n - 1 - (h + t)].x;
                    dst[t] = sg * hi;
                }
                if (tid == 0) ca.hi_diverged[slot] = nan_seen ? 1 : 0;
            } else {
                // signal.cpp:200-263
                const float2 prev_last = ca.bb_last[slot];
                float2 *bb = y;  // assemble the overlapped first half in the free buffer
                for (int t = tid; t < h; t += TPB) {
                    const float2 lo = x[t], old = bb_hi[t];
                    bb[t] = make_float2(__fadd_rn(sg * lo.x, old.x), __fadd_rn(sg * lo.y, old.y));
                }
                __syncthreads();
                for (int t = tid; t < h; t += TPB) {
                    const float2 hi = x[h + t];
                    bb_hi[t] = make_float2(sg * hi.x, sg * hi.y);
                    const float2 b = bb[t];
                    float o;
                    if (mode == MODE_AM) {
                        o = __fsqrt_rn(__fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y)));  // dsp.cpp:116-126
                    } else {
                        const float2 pv = (t == 0) ? prev_last : bb[t - 1];
                        const float c = pv.x, d = -pv.y;  // buf[i] * conj(prev), dsp.cpp:27-35
                        const float re = __fsub_rn(__fmul_rn(b.x, c), __fmul_rn(b.y, d));
                        const float im = __fadd_rn(__fmul_rn(b.x, d), __fmul_rn(b.y, c));
                        o = atan2f(im, re);
                    }
                    nan_seen |= (o != o);
                    audio[t] = o;
                }
                nan_seen = __syncthreads_or(nan_seen);
                if (tid == 0) ca.bb_last[slot] = bb[h - 1];
                if (!nan_seen && ca.hi_diverged[slot]) {
                    // audio_real_prev <- audio_real[n/2..n) left behind by a NaN-dropped SSB frame
                    for (int t = tid; t < h; t += TPB) real_prev[t] = real_hi[t];
                    __syncthreads();
                    if (tid == 0) ca.hi_diverged[slot] = 0;
                }
            }
            if (tid == 0) {
                ca.valid_a[(size_t)f * ca.max_clients + slot] = nan_seen ? 0 : 1;
                ca.pwr[(size_t)f * ca.max_clients + slot] = s_pw[fr];
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Frame-chunked demodulation: one WARP per (client, chunk of `cl.chunk` consecutive frames), nothing but warp-level
// synchronisation. The inverse FFTs of different frames are independent and the overlap-add only couples a frame to its
// predecessor, so a task that does not start at the first frame of the batch simply recomputes the predecessor's inverse
// FFT (and, for FM, the one before that: the discriminator's first phase step needs the last overlapped sample) with the
// very same stage code - its results are bit-identical to the sequential kernel's. Every stage of a 360-point transform is
// 72 .. 180 butterflies: a 32-lane warp is a far better fit than a 256-thread CTA, and eight independent warps per CTA
// overlap one another's shared-memory latency with no barrier in between.
// A frame dropped by the NaN guard (src/signal.cpp:266-275: the overlap state does not advance) breaks the independence;
// such clients are flagged and replayed by client_demod_kernel (redo_only) from the state the batch began with.
// Dynamic smem: kDemodWarps * (2 n + h) float2.
// ------------------------------------------------------------------------------------------------
constexpr int kDemodWarps = 8;

// NF: audio FFT size known at compile time (0 = run-time plan from ClientArrays)
template <int NF>
__global__ void __launch_bounds__(32 * kDemodWarps, 3) client_demod_warp_kernel(const ClientArrays ca, const ClientLaunch cl) {
    extern __shared__ float2 smem_c[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.x * (blockDim.x >> 5) + warp;  // (fewer than kDemodWarps warps per CTA for very long audio FFTs)
    if (task >= cl.nactive * cl.nchunks) return;
    const int ci = task / cl.nchunks, k = task - ci * cl.nchunks;
    const int slot = cl.order[ci];
    const int n = NF ? NF : ca.n, h = NF ? NF / 2 : ca.h;
    float2 *bufX = smem_c + (size_t)warp * (2 * n + h);
    float2 *bufY = bufX + n;
    float2 *prevC = bufY + n;  // overlap state: upper half of the previous frame (SSB: .x only)
    const ClientSlot cs = ca.slots[slot];
    const size_t R = cl.is_real ? cl.fft_size / 2 : cl.fft_size;
    const size_t base_idx = cl.is_real ? 0 : cl.fft_size / 2 + 1;
    const size_t off = ((size_t)cs.l + base_idx) % R;  // src/websocket.cpp:182
    const int len = cs.r - cs.l;
    const int audio_m = cs.m_floor - cs.l;
    const int mode = cs.mode;
    const bool ssb = mode == MODE_USB || mode == MODE_LSB;
    const bool fresh = (cs.flags & CF_RESET_ALL) != 0;
    const int f_begin = k * cl.chunk, f_end = min(cl.nframes, f_begin + cl.chunk);
    if (k == 0 && !fresh && cl.sin_hi_diverged[slot]) {  // state left behind by a dropped SSB frame: sequential replay
        if (lane == 0) cl.redo[slot] = 1;
    }

    // placement of the slice of frame f into the IFFT input (signal.cpp:125-198)
    auto gather = [&](int f) {
        const float2 *buf = cl.spec + (size_t)f * cl.spec_stride + off;
        for (int kk = lane; kk < n; kk += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (ssb) {
                // c2r reads bins 0..n/2 of its input; the other half is the Hermitian mirror
                const int k2 = (kk <= n / 2) ? kk : n - kk;
                const int i = (mode == MODE_USB) ? (audio_m + k2) : (audio_m - k2);
                if (i >= 0 && i < len) {
                    v = buf[i];
                    if (k2 == 0 || k2 == n / 2) v.y = 0.f;  // c2r ignores Im of DC / Nyquist
                    else if (kk > n / 2) v.y = -v.y;
                }
            } else {
                const int d = (kk < n / 2) ? kk : kk - n;  // positive bins [0, n/2), negative [-n/2+1, -1]; kk = n/2 stays 0
                const int i = audio_m + d;
                if (kk != n / 2 && i >= 0 && i < len) v = buf[i];
            }
            bufX[kk] = v;
        }
        __syncwarp();
    };
    // The same placement for a compile-time n, reading the slice ONCE: all loads of the lane are issued together (the slice
    // is at most n bins = (n + 31) / 32 per lane), the slice power is summed from the registers in the order of the loop
    // above (lane-strided, then the warp reduction), and every bin is then scattered to its place(s) in the zeroed input.
    auto load_place = [&](int f, bool want_pw) -> float {
        constexpr int NI = NF ? (NF + 31) / 32 : 1, NB = 6;  // NB loads in flight per lane (registers)
        const float2 *buf = cl.spec + (size_t)f * cl.spec_stride + off;
#pragma unroll
        for (int i = 0; i < NI; i++)
            if (lane + 32 * i < n) bufX[lane + 32 * i] = make_float2(0.f, 0.f);
        __syncwarp();
        float pw = 0.f;
#pragma unroll
        for (int i0 = 0; i0 < NI; i0 += NB) {
            float2 v[NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const int idx = lane + 32 * (i0 + i);
                v[i] = (i0 + i < NI && idx < len) ? buf[idx] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const int idx = lane + 32 * (i0 + i);
                if (i0 + i < NI && idx < len) {
                    const float2 x = v[i];
                    if (want_pw) pw += x.x * x.x + x.y * x.y;
                    if (ssb) {
                        const int k2 = (mode == MODE_USB) ? idx - audio_m : audio_m - idx;
                        if (k2 >= 0 && k2 <= n / 2) {
                            const bool edge = (k2 == 0 || k2 == n / 2);  // c2r ignores Im of DC / Nyquist
                            bufX[k2] = make_float2(x.x, edge ? 0.f : x.y);
                            if (!edge) bufX[n - k2] = make_float2(x.x, -x.y);  // Hermitian mirror
                        }
                    } else {
                        const int d = idx - audio_m;  // positive bins [0, n/2), negative [-n/2+1, -1]
                        if (d > -n / 2 && d < n / 2) bufX[d >= 0 ? d : d + n] = x;
                    }
                }
            }
        }
        __syncwarp();
        return want_pw ? warp_sum(pw) : 0.f;
    };
    // unnormalised inverse FFT of bufX (signal.cpp:138,154,214); returns the buffer that holds the result
    auto ifft = [&]() -> float2 * {
        if constexpr (NF != 0) return ifft_static<(NF ? NF : 4), 1, (NF ? NF : 4)>(bufX, bufY, ca.Wn, lane);
        float2 *xb = bufX, *yb = bufY;
        int s = 1;
        for (int st = 0; st < ca.nstages; st++) {
            const int Rr = ca.radix[st];
            if (Rr == 4) ifft_stage_fixed<4>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], 1, ca.Wn, lane, 32);
            else if (Rr == 2) ifft_stage_fixed<2>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], 1, ca.Wn, lane, 32);
            else if (Rr == 3) ifft_stage_fixed<3>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], 1, ca.Wn, lane, 32);
            else if (Rr == 5) ifft_stage_fixed<5>(xb, yb, n, s, ca.mul_s[st], ca.mul_items[st], 1, ca.Wn, lane, 32);
            else ifft_stage_generic(xb, yb, n, s, Rr, ca.mul_s[st], ca.mul_n, 1, ca.Wn, lane, 32);
            s *= Rr;
            float2 *t = xb;
            xb = yb;
            yb = t;
            __syncwarp();
        }
        return xb;
    };
    auto sign_of = [&](int f) -> float {  // signal.cpp:160-172,223-237: parity flip of every other frame
        const unsigned long long frame_num = cl.frame_num0 + f;
        const int m_idx = cs.m_floor;
        const bool negate = (frame_num & 1ull) && (((m_idx % 2 == 0) && !cl.is_real) || ((m_idx % 2 == 1) && cl.is_real));
        return negate ? -1.f : 1.f;
    };
    // the upper half of a transformed frame, as the next frame will overlap it
    auto take_upper = [&](const float2 *x, float sg) {
        for (int t = lane; t < h; t += 32) {
            if (ssb) prevC[t] = make_float2(sg * ((mode == MODE_USB) ? x[h + t].x : x[n - 1 - (h + t)].x), 0.f);
            else prevC[t] = make_float2(sg * x[h + t].x, sg * x[h + t].y);
        }
        __syncwarp();
    };

    // ---- the state this task starts from ----
    float2 last = make_float2(0.f, 0.f);  // FM: the last overlapped sample of the previous frame
    if (k == 0) {
        for (int t = lane; t < h; t += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (!fresh) v = ssb ? make_float2(cl.sin_real_prev[(size_t)slot * h + t], 0.f) : cl.sin_bb_hi[(size_t)slot * h + t];
            prevC[t] = v;
        }
        if (!fresh) last = cl.sin_bb_last[slot];
        __syncwarp();
    } else {
        float2 hi2 = make_float2(0.f, 0.f);  // upper-half sample h - 1 of frame f_begin - 2, as frame f_begin - 1 overlapped it
        if (mode == MODE_FM) {
            if (f_begin >= 2) {
                if constexpr (NF != 0) load_place(f_begin - 2, false);
                else gather(f_begin - 2);
                const float2 *x = ifft();
                const float sg = sign_of(f_begin - 2);
                hi2 = make_float2(sg * x[n - 1].x, sg * x[n - 1].y);
                __syncwarp();
            } else if (!fresh) {
                hi2 = cl.sin_bb_hi[(size_t)slot * h + h - 1];
            }
        }
        if constexpr (NF != 0) load_place(f_begin - 1, false);
        else gather(f_begin - 1);
        const float2 *x = ifft();
        const float sg = sign_of(f_begin - 1);
        if (!ssb) last = make_float2(__fadd_rn(sg * x[h - 1].x, hi2.x), __fadd_rn(sg * x[h - 1].y, hi2.y));
        take_upper(x, sg);
    }

    // ---- the task's own frames, in order ----
    for (int f = f_begin; f < f_end; f++) {
        float pw = 0.f;  // slice power (signal.cpp:117-119)
        if constexpr (NF != 0) {
            pw = load_place(f, true);
        } else {
            const float2 *buf = cl.spec + (size_t)f * cl.spec_stride + off;
            for (int i = lane; i < len; i += 32) {
                const float2 v = buf[i];
                pw += v.x * v.x + v.y * v.y;
            }
            pw = warp_sum(pw);
            gather(f);
        }
        float2 *x = ifft();
        float2 *y = (x == bufX) ? bufY : bufX;  // free scratch
        const float sg = sign_of(f);
        float *audio = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
        int nan_seen = 0;
        if (ssb) {
            // signal.cpp:155-172: (LSB: time reverse), parity negate, overlap-add
            for (int t = lane; t < h; t += 32) {
                const float lo = (mode == MODE_USB) ? x[t].x : x[n - 1 - t].x;
                const float o = __fadd_rn(sg * lo, prevC[t].x);
                nan_seen |= (o != o);
                audio[t] = o;
            }
        } else {
            // signal.cpp:200-263
            for (int t = lane; t < h; t += 32) {
                const float2 lo = x[t], old = prevC[t];
                y[t] = make_float2(__fadd_rn(sg * lo.x, old.x), __fadd_rn(sg * lo.y, old.y));
            }
            __syncwarp();
            for (int t = lane; t < h; t += 32) {
                const float2 b = y[t];
                float o;
                if (mode == MODE_AM) {
                    o = __fsqrt_rn(__fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y)));  // dsp.cpp:116-126
                } else {
                    const float2 pv = (t == 0) ? last : y[t - 1];
                    const float c = pv.x, d = -pv.y;  // buf[i] * conj(prev), dsp.cpp:27-35
                    const float re = __fsub_rn(__fmul_rn(b.x, c), __fmul_rn(b.y, d));
                    const float im = __fadd_rn(__fmul_rn(b.x, d), __fmul_rn(b.y, c));
                    o = atan2f(im, re);
                }
                nan_seen |= (o != o);
                audio[t] = o;
            }
            last = y[h - 1];
            __syncwarp();
        }
        nan_seen = __any_sync(0xffffffffu, nan_seen);
        take_upper(x, sg);
        if (lane == 0) {
            ca.valid_a[(size_t)f * ca.max_clients + slot] = nan_seen ? 0 : 1;
            ca.pwr[(size_t)f * ca.max_clients + slot] = pw;
            if (nan_seen) cl.redo[slot] = 1;
        }
    }

    // ---- the last task of a client leaves the state for the next batch ----
    if (k == cl.nchunks - 1) {
        for (int t = lane; t < h; t += 32) {
            const size_t o = (size_t)slot * h + t;
            ca.real_prev[o] = ssb ? prevC[t].x : (fresh ? 0.f : cl.sin_real_prev[o]);
            ca.real_hi[o] = fresh ? 0.f : cl.sin_real_hi[o];
            ca.bb_hi[o] = ssb ? (fresh ? make_float2(0.f, 0.f) : cl.sin_bb_hi[o]) : prevC[t];
        }
        if (lane == 0) {
            ca.bb_last[slot] = ssb ? (fresh ? make_float2(0.f, 0.f) : cl.sin_bb_last[slot]) : last;
            ca.hi_diverged[slot] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sequential tails. The float recurrences (two running sums of the DC blocker, the AGC gain) are
// strictly serial per client, so they run one LANE per client (cpb clients per block) and are
// stripped to the loop-carried adds; everything that is not loop-carried (divisions, window
// maxima, conversions) runs beside them with lanes along the sample index.
// Tiles are [j][client] with pitch cpb + 1 (odd): conflict-free for lanes along j and along client.
// ------------------------------------------------------------------------------------------------
constexpr int kTailThreads = 512;
constexpr int kTailWarps = kTailThreads / 32;
constexpr int kTailMaxCpb = 16;  // one warp per client in the parallel phases

__device__ __forceinline__ long long floordiv_ll(long long a, int b) {
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

// Shared-memory rows are [client][sample] with a pitch that is a multiple of 4 floats and == 4 (mod 32):
// 16-byte aligned for 128-bit accesses, and conflict-free both for lanes along the sample index and for
// lanes along the client index (quarter-warps of 128-bit accesses land on disjoint bank groups).
__host__ __device__ inline int tail_pitch(int len) {
    int p = (len + 3) & ~3;
    while ((p & 31) != 4) p += 4;
    return p;
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// running sum  s <- (s - ext[j]) + ext[j + D]  for j = 0..h-1 (src/utils.h:80-85: sum -= q.back(); sum += val),
// one lane per client. ext = [last D values of the previous frames | this frame]. Only the two dependent adds
// per sample sit on the loop-carried path: operands are fetched one group ahead with 128-bit loads.
__device__ __forceinline__ float running_sum_serial(float s, const float *ext, float *out, int h, int D) {
    int j = 0;
    if ((D & 3) == 0) {
#define RS4(OO, XX, dst)                                    \
    {                                                       \
        float4 r;                                           \
        r.x = s = __fadd_rn(__fadd_rn(s, -OO.x), XX.x);     \
        r.y = s = __fadd_rn(__fadd_rn(s, -OO.y), XX.y);     \
        r.z = s = __fadd_rn(__fadd_rn(s, -OO.z), XX.z);     \
        r.w = s = __fadd_rn(__fadd_rn(s, -OO.w), XX.w);     \
        sts4(dst, r);                                       \
    }
        if (h >= 8) {
            float4 o0 = lds4(ext), x0 = lds4(ext + D);
            for (; j + 8 <= h; j += 8) {
                const float4 o1 = lds4(ext + j + 4), x1 = lds4(ext + j + 4 + D);
                RS4(o0, x0, out + j)
                if (j + 12 <= h) {  // operands of the next group, fetched while this one's adds retire
                    o0 = lds4(ext + j + 8);
                    x0 = lds4(ext + j + 8 + D);
                }
                RS4(o1, x1, out + j + 4)
            }
        }
        for (; j + 4 <= h; j += 4) {
            const float4 o = lds4(ext + j), x = lds4(ext + j + D);
            RS4(o, x, out + j)
        }
#undef RS4
    }
    for (; j < h; j++) {
        s = __fadd_rn(__fadd_rn(s, -ext[j]), ext[j + D]);
        out[j] = s;
    }
    return s;
}

// KB = ceil(h / 32) when known at compile time (all row loads of a client are issued before any is
// consumed), 0 = runtime loops for unusually long audio frames.
template <int KB>
__global__ void __launch_bounds__(kTailThreads) client_tail_kernel(const ClientArrays ca, const ClientLaunch cl) {
    extern __shared__ __align__(16) float smem_t[];
    const int h = ca.h, D = ca.D, L = ca.L, NC = ca.NC;
    const int cpb = cl.cpb;
    const int pX = tail_pitch(D + h), pH = tail_pitch(h);
    float *tX = smem_t;            // [cpb][D + h]  DC input, extended by the last D inputs
    float *tM = tX + cpb * pX;     // [cpb][D + h]  first-stage averages, extended likewise (sum, then sum / D)
    float *tY = tM + cpb * pX;     // [cpb][h]      running sum 2 -> DC blocker output
    float *tC = tY + cpb * pH;     // [cpb][h]      look-ahead-delayed sample -> AGC output
    float *tP = tC + cpb * pH;     // [cpb][h]      window maximum over the old samples -> desired gain
    __shared__ int s_slot[32];
    __shared__ unsigned char s_valid[32];
    __shared__ long long s_t0[32];
    __shared__ float s_pmax[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g0 = blockIdx.x * cpb;
    if (tid < 32) s_slot[tid] = (tid < cpb && g0 + tid < cl.nactive) ? cl.order[g0 + tid] : -1;
    __syncthreads();

    // per-lane (= per-client) scalars live in warp 0's registers across frames
    float sum1 = 0.f, sum2 = 0.f, gain = 0.f;
    long long t0 = 0;
    const int my_slot = s_slot[lane];
    {
        // resets (AGC::reset, audioprocessing.cpp:70-74; fresh slot): lanes along the arrays
        for (int ci = warp; ci < cpb; ci += kTailWarps) {
            const int slot = s_slot[ci];
            if (slot < 0) continue;
            const int fl = ca.slots[slot].flags;
            if (fl & (CF_RESET_ALL | CF_RESET_AGC)) {
                float *ring = ca.agc_ring + (size_t)slot * NC * h;
                for (int i = lane; i < NC * h; i += 32) ring[i] = 0.f;
                for (int i = lane; i < NC; i += 32) ca.agc_cmax[(size_t)slot * NC + i] = 0.f;
            }
            if (fl & CF_RESET_ALL)
                for (int i = lane; i < D; i += 32) {
                    ca.dc_x[(size_t)slot * D + i] = 0.f;
                    ca.dc_m[(size_t)slot * D + i] = 0.f;
                }
        }
        if (warp == 0 && my_slot >= 0) {
            const int fl = ca.slots[my_slot].flags;
            if (!(fl & (CF_RESET_ALL | CF_RESET_AGC))) {
                gain = ca.agc_gain[my_slot];
                t0 = ca.agc_t0[my_slot];
            }
            if (!(fl & CF_RESET_ALL)) {
                sum1 = ca.dc_sum[2 * my_slot];
                sum2 = ca.dc_sum[2 * my_slot + 1];
            }
        }
    }
    __syncthreads();

    const float Df = (float)D;
    long long tprev = 0;
    const bool profiling = cl.prof != nullptr && blockIdx.x == 0 && tid == 0;
#define TAIL_PHASE(k)                              \
    if (profiling) {                               \
        const long long tn = clock64();            \
        cl.prof[k] += tn - tprev;                  \
        tprev = tn;                                \
    }
    for (int f = 0; f < cl.nframes; f++) {
        if (warp == 0) {
            s_valid[lane] = (my_slot >= 0) ? ca.valid_a[(size_t)f * ca.max_clients + my_slot] : 0;
            s_t0[lane] = t0;
        }
        __syncthreads();
        if (profiling) tprev = clock64();
        // ---- P0: loads, lanes along j. Old-sample window maxima and delayed samples. ----
        for (int ci = warp; ci < cpb; ci += kTailWarps) {
            const int slot = s_slot[ci];
            if (slot < 0 || !s_valid[ci]) continue;
            float *X = tX + ci * pX, *Mx = tM + ci * pX, *Cc = tC + ci * pH, *Pp = tP + ci * pH;
            const float *a = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
            // AGC window of output j is samples [t0+j-L+1, t0+j]. Its old part starts at chunk c0,
            // column col0 and, as j grows, walks through chunk c0 then chunk c0+1.
            const long long tt = s_t0[ci];
            const long long lo0 = tt - L + 1;
            const long long c0 = floordiv_ll(lo0, h);
            const int col0 = (int)(lo0 - c0 * h);
            const long long Fc = tt / h;  // chunk index of the samples produced by this frame
            const float *ring = ca.agc_ring + (size_t)slot * NC * h;
            const float *cmax = ca.agc_cmax + (size_t)slot * NC;
            const float *row0 = ring + (size_t)((((c0) % NC) + NC) % NC) * h;
            const float *row1 = ring + (size_t)((((c0 + 1) % NC) + NC) % NC) * h;
            // maxima of the whole chunks strictly between the walking chunk and this frame
            float m0 = 0.f, m1 = 0.f;
            for (long long c = c0 + 1 + lane; c <= Fc - 1; c += 32) {
                const float v = cmax[(int)(((c % NC) + NC) % NC)];
                m0 = fmaxf(m0, v);
                if (c >= c0 + 2) m1 = fmaxf(m1, v);
            }
            float ra[KB > 0 ? KB : 1], r0[KB > 0 ? KB : 1], r1[KB > 0 ? KB : 1];
            if constexpr (KB > 0) {
#pragma unroll
                for (int b = 0; b < KB; b++) {  // every global load of this client is in flight together
                    const int col = 32 * b + lane;
                    ra[b] = (col < h) ? a[col] : 0.f;
                    r0[b] = (col < h) ? row0[col] : 0.f;
                    r1[b] = (col < h) ? row1[col] : 0.f;
                }
            }
            for (int j = lane; j < D; j += 32) {
                X[j] = ca.dc_x[(size_t)slot * D + j];
                Mx[j] = ca.dc_m[(size_t)slot * D + j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
            }
            if constexpr (KB > 0) {
#pragma unroll
                for (int b = 0; b < KB; b++) {
                    const int col = 32 * b + lane;
                    if (col < h) X[D + col] = ra[b];
                }
            } else {
                for (int j = lane; j < h; j += 32) X[D + j] = a[j];
            }
            // The two old chunks go to shared memory (rows that are still free at this point), then each lane
            // owns a contiguous segment: local suffix max, one warp scan over the 32 segment maxima, fold back.
            float *R0 = tY + ci * pH, *R1 = tM + ci * pX + D;
            if constexpr (KB > 0) {
#pragma unroll
                for (int b = 0; b < KB; b++) {
                    const int col = 32 * b + lane;
                    if (col < h) {
                        R0[col] = r0[b];
                        R1[col] = r1[b];
                    }
                }
            } else {
                for (int col = lane; col < h; col += 32) {
                    R0[col] = row0[col];
                    R1[col] = row1[col];
                }
            }
            __syncwarp();
            constexpr bool kStatic = KB > 0;
            const int seg = kStatic ? KB : (h + 31) / 32;
            const int s0 = lane * seg;
            float run0 = 0.f, run1 = 0.f;
#pragma unroll
            for (int u = (kStatic ? KB : seg) - 1; u >= 0; u--) {
                const int col = s0 + u;
                if (col < h) {
                    run0 = fmaxf(run0, fabsf(R0[col]));
                    run1 = fmaxf(run1, fabsf(R1[col]));
                }
            }
            float in0 = run0, in1 = run1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {  // inclusive suffix max over lanes
                const float a0 = __shfl_down_sync(0xffffffffu, in0, o);
                const float a1 = __shfl_down_sync(0xffffffffu, in1, o);
                if (lane + o < 32) {
                    in0 = fmaxf(in0, a0);
                    in1 = fmaxf(in1, a1);
                }
            }
            run0 = __shfl_down_sync(0xffffffffu, in0, 1);
            run1 = __shfl_down_sync(0xffffffffu, in1, 1);
            if (lane == 31) run0 = run1 = 0.f;
#pragma unroll
            for (int u = (kStatic ? KB : seg) - 1; u >= 0; u--) {
                const int col = s0 + u;
                if (col < h) {
                    const float x0 = R0[col], x1 = R1[col];
                    run0 = fmaxf(run0, fabsf(x0));
                    run1 = fmaxf(run1, fabsf(x1));
                    const int ja = col - col0;      // output index served by (chunk c0, col)
                    const int jb = h + col - col0;  // ... by (chunk c0+1, col)
                    if (ja >= 0) {                  // ja < h always
                        Cc[ja] = x0;
                        Pp[ja] = fmaxf(run0, m0);
                    }
                    if (jb < h) {                   // jb >= 0 always
                        Cc[jb] = x1;
                        Pp[jb] = fmaxf(run1, m1);
                    }
                }
            }
        }
        __syncthreads();
        TAIL_PHASE(0)
        // ---- P1 (serial): first running sum of the DC blocker, src/utils.h:80-85 ----
        if (warp == 0 && my_slot >= 0 && s_valid[lane])
            sum1 = running_sum_serial(sum1, tX + lane * pX, tM + lane * pX + D, h, D);
        __syncthreads();
        TAIL_PHASE(1)
        // ---- P2 (parallel): getAverage() = sum / length ----
        for (int ci = warp; ci < cpb; ci += kTailWarps) {
            if (s_slot[ci] < 0 || !s_valid[ci]) continue;
            float *Mx = tM + ci * pX + D;
            if constexpr (KB > 0) {
#pragma unroll
                for (int b = 0; b < KB; b++) {
                    const int j = 32 * b + lane;
                    if (j < h) Mx[j] = __fdiv_rn(Mx[j], Df);
                }
            } else {
                for (int j = lane; j < h; j += 32) Mx[j] = __fdiv_rn(Mx[j], Df);
            }
        }
        __syncthreads();
        TAIL_PHASE(2)
        // ---- P3 (serial): second running sum ----
        if (warp == 0 && my_slot >= 0 && s_valid[lane])
            sum2 = running_sum_serial(sum2, tM + lane * pX, tY + lane * pH, h, D);
        __syncthreads();
        TAIL_PHASE(3)
        // ---- P4 (parallel): DC output y = x[delayed] - ma2 (src/utils.h:145-149: buffer[delay-1] is the input
        //      of delay-1 samples ago = ext[j+1]), running |y| maximum, window peak, desired gain
        //      (audioprocessing.cpp:48-52) ----
        for (int ci = warp; ci < cpb; ci += kTailWarps) {
            if (s_slot[ci] < 0 || !s_valid[ci]) continue;
            const float *X = tX + ci * pX;
            float *Y = tY + ci * pH, *Pp = tP + ci * pH;
            // lane owns the contiguous samples [lane*seg, lane*seg + seg): local running max, one warp scan of
            // the 32 segment maxima, then the exclusive prefix is folded back in
            constexpr bool kStatic = KB > 0;
            const int seg = kStatic ? KB : (h + 31) / 32;
            const int j0 = lane * seg;
            float run = 0.f;
#pragma unroll
            for (int u = 0; u < (kStatic ? KB : seg); u++) {
                const int j = j0 + u;
                if (j < h) {
                    const float y = __fsub_rn(X[j + 1], __fdiv_rn(Y[j], Df));
                    Y[j] = y;
                    run = fmaxf(run, fabsf(y));
                }
            }
            float incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float other = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = fmaxf(incl, other);
            }
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.f;
            run = excl;
#pragma unroll
            for (int u = 0; u < (kStatic ? KB : seg); u++) {
                const int j = j0 + u;
                if (j < h) {
                    run = fmaxf(run, fabsf(Y[j]));
                    const float peak = fmaxf(Pp[j], run);
                    Pp[j] = __fdiv_rn(ca.desired, __fadd_rn(peak, 1e-10f));
                }
            }
            const float carry = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) s_pmax[ci] = carry;
        }
        __syncthreads();
        TAIL_PHASE(4)
        // ---- P5 (serial): attack/release recurrence on the gain, audioprocessing.cpp:54-63 ----
        if (warp == 0 && my_slot >= 0 && s_valid[lane]) {
            const float att = ca.attack, rel = ca.release;
            float *Cc = tC + lane * pH;
            const float *Pp = tP + lane * pH;
            // outputs stay 0 until the look-ahead buffer is full (audioprocessing.cpp:45,64-66)
            long long first = (long long)L - 1 - t0;
            if (first < 0) first = 0;
            if (first > h) first = h;
            int j = 0;
            for (; j < (int)first; j++) Cc[j] = 0.f;  // AGC output is 0 until the buffer is full ...
            // gain -= attack*(gain-desired) if desired < gain, else gain += release*(desired-gain)
            // (audioprocessing.cpp:54-60). With t = gain - desired both arms are gain - c*t (negation and the
            // sign-symmetric product are exact), and because attack >= release > 0 the arm is picked by
            // max(attack*t, release*t): no predicate on the loop-carried path.
            auto step = [&](float d) {
                const float t = __fsub_rn(gain, d);
                gain = __fsub_rn(gain, fmaxf(__fmul_rn(att, t), __fmul_rn(rel, t)));
                return gain;
            };
#define G4(DD, dst)              \
    {                            \
        float4 r;                \
        r.x = step(DD.x);        \
        r.y = step(DD.y);        \
        r.z = step(DD.z);        \
        r.w = step(DD.w);        \
        sts4(dst, r);            \
    }
            // the gain sequence overwrites the desired-gain row; the product with the delayed sample is
            // formed in the parallel store phase
            float *Gg = tP + lane * pH;
            for (; (j & 3) && j < h; j++) Gg[j] = step(Pp[j]);
            if (j + 8 <= h) {
                float4 d0 = lds4(Pp + j);
                for (; j + 8 <= h; j += 8) {
                    const float4 d1 = lds4(Pp + j + 4);
                    G4(d0, Gg + j)
                    if (j + 12 <= h) d0 = lds4(Pp + j + 8);
                    G4(d1, Gg + j + 4)
                }
            }
            for (; j + 4 <= h; j += 4) {
                const float4 d = lds4(Pp + j);
                G4(d, Gg + j)
            }
#undef G4
            for (; j < h; j++) Gg[j] = step(Pp[j]);
            t0 += h;
        }
        __syncthreads();
        TAIL_PHASE(5)
        // ---- P6 (parallel): write back ----
        for (int ci = warp; ci < cpb; ci += kTailWarps) {
            const int slot = s_slot[ci];
            if (slot < 0) continue;
            int *pcm = ca.pcm + ((size_t)f * ca.max_clients + slot) * h;
            if (!s_valid[ci]) {
                for (int j = lane; j < h; j += 32) pcm[j] = 0;
                if (lane == 0) ca.valid[(size_t)f * ca.max_clients + slot] = 0;
                continue;
            }
            if (lane == 0) ca.valid[(size_t)f * ca.max_clients + slot] = 1;
            const float *X = tX + ci * pX, *Mx = tM + ci * pX, *Y = tY + ci * pH, *Cc = tC + ci * pH, *Gg = tP + ci * pH;
            const long long Fc = s_t0[ci] / h;
            const int row = (int)(Fc % NC);
            float *dst = ca.agc_ring + ((size_t)slot * NC + row) * h;
#pragma unroll
            for (int b = 0; b < (KB > 0 ? KB : 1); b++) {
                for (int j = 32 * b + lane; j < h; j += (KB > 0 ? h : 32)) {
                    // dsp.cpp:152-165 with mult = 65536/4
                    const float t = __fadd_rn(__fmul_rn(__fmul_rn(Cc[j], Gg[j]), 16384.f), 32768.5f);
                    int v = __float2int_rz(t) - 32768;
                    v = max(min(v, 32767), -32768);
                    pcm[j] = v;
                    dst[j] = Y[j];
                }
            }
            if (lane == 0) ca.agc_cmax[(size_t)slot * NC + row] = s_pmax[ci];
            // DC state: the last D entries of the extended rows
            for (int j = lane; j < D; j += 32) {
                ca.dc_x[(size_t)slot * D + j] = X[h + j];
                ca.dc_m[(size_t)slot * D + j] = Mx[h + j];
            }
        }
        __syncthreads();
        TAIL_PHASE(6)
    }
#undef TAIL_PHASE
    if (warp == 0 && my_slot >= 0) {
        ca.agc_gain[my_slot] = gain;
        ca.agc_t0[my_slot] = t0;
        ca.dc_sum[2 * my_slot] = sum1;
        ca.dc_sum[2 * my_slot + 1] = sum2;
    }
}

}  // namespace b200
