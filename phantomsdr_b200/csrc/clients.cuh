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
    int cpb;             // tail kernel: clients per block (<= 32), tiles are [j][cpb + 1]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One Stockham stage of radix R (sign +1):  x[t + j*items] -> y[q + s*(R*p + r)], t = p*s + q
template <int R>
__device__ __forceinline__ void ifft_stage_fixed(const float2 *x, float2 *y, int n, int s, const float2 *Wn, int lane) {
    const int items = n / R;
    for (int t = lane; t < items; t += 32) {
        const int p = t / s, q = t - p * s;
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = x[t + j * items];
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
        const int ob = q + s * R * p;
        y[ob] = a[0];
        const int tw = p * s;  // W_ncur^(p r) = Wn[p*r*s]
#pragma unroll
        for (int r = 1; r < R; r++) y[ob + s * r] = cmul(a[r], __ldg(Wn + tw * r));
    }
}

// generic (prime) radix: one lane per OUTPUT, R complex MACs each
__device__ __forceinline__ void ifft_stage_generic(const float2 *x, float2 *y, int n, int s, int R, const float2 *Wn,
                                                   int lane) {
    const int items = n / R;
    const int step = n / R;  // W_R = Wn[n/R]
    for (int o = lane; o < n; o += 32) {
        const int t = o / R, r = o - t * R;
        const int p = t / s, q = t - p * s;
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
        y[q + s * (R * p + r)] = acc;
    }
}

// grid: ceil(nactive / WPB) blocks of WPB warps; dynamic smem WPB * 2n float2
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) client_demod_kernel(const ClientArrays ca, const ClientLaunch cl) {
    extern __shared__ float2 smem_c[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * WPB + warp;
    if (ci >= cl.nactive) return;
    const int slot = cl.order[ci];
    const int n = ca.n, h = ca.h;
    float2 *bufX = smem_c + (size_t)warp * 2 * n;
    float2 *bufY = bufX + n;
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

    if (cs.flags & CF_RESET_ALL) {  // freshly opened slot: zeroed scratch as AudioClient's ctor (signal.cpp:38-52)
        for (int i = lane; i < h; i += 32) {
            real_prev[i] = 0.f;
            real_hi[i] = 0.f;
            bb_hi[i] = make_float2(0.f, 0.f);
        }
        if (lane == 0) {
            ca.bb_last[slot] = make_float2(0.f, 0.f);
            ca.hi_diverged[slot] = 0;
        }
        __syncwarp();
    }

    for (int f = 0; f < cl.nframes; f++) {
        const unsigned long long frame_num = cl.frame_num0 + f;
        const float2 *buf = cl.spec + (size_t)f * cl.spec_stride + off;
        for (int i = lane; i < n; i += 32) bufX[i] = make_float2(0.f, 0.f);
        __syncwarp();
        // gather + placement + slice power (signal.cpp:117-198)
        float pw = 0.f;
        for (int i = lane; i < len; i += 32) {
            const float2 v = buf[i];
            pw += v.x * v.x + v.y * v.y;
            if (mode == MODE_USB || mode == MODE_LSB) {
                const int kk = (mode == MODE_USB) ? (i - audio_m) : (audio_m - i);
                if (kk >= 0 && kk <= n / 2) {
                    if (kk == 0 || kk == n / 2) {
                        bufX[kk] = make_float2(v.x, 0.f);  // c2r ignores Im of DC / Nyquist
                    } else {
                        bufX[kk] = v;
                        bufX[n - kk] = make_float2(v.x, -v.y);
                    }
                }
            } else {
                const int d = i - audio_m;
                if (d >= 0 && d < n / 2) bufX[d] = v;
                else if (d < 0 && d >= -(n / 2) + 1) bufX[n + d] = v;
            }
        }
        pw = warp_sum(pw);
        __syncwarp();
        // inverse FFT, unnormalised (signal.cpp:138,154,214)
        float2 *x = bufX, *y = bufY;
        int s = 1;
        for (int st = 0; st < ca.nstages; st++) {
            const int Rr = ca.radix[st];
            if (Rr == 4) ifft_stage_fixed<4>(x, y, n, s, ca.Wn, lane);
            else if (Rr == 2) ifft_stage_fixed<2>(x, y, n, s, ca.Wn, lane);
            else if (Rr == 3) ifft_stage_fixed<3>(x, y, n, s, ca.Wn, lane);
            else if (Rr == 5) ifft_stage_fixed<5>(x, y, n, s, ca.Wn, lane);
            else ifft_stage_generic(x, y, n, s, Rr, ca.Wn, lane);
            s *= Rr;
            float2 *t = x;
            x = y;
            y = t;
            __syncwarp();
        }
        // x holds the time-domain result
        const int m_idx = cs.m_floor;
        const bool negate = (frame_num & 1ull) && (((m_idx % 2 == 0) && !cl.is_real) || ((m_idx % 2 == 1) && cl.is_real));
        const float sg = negate ? -1.f : 1.f;
        float *audio = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
        bool nan_seen = false;
        if (mode == MODE_USB || mode == MODE_LSB) {
            // signal.cpp:155-172: (LSB: time reverse), parity negate, overlap-add
            for (int t = lane; t < h; t += 32) {
                const float lo = (mode == MODE_USB) ? x[t].x : x[n - 1 - t].x;
                const float o = __fadd_rn(sg * lo, real_prev[t]);
                nan_seen |= (o != o);
                audio[t] = o;
            }
            nan_seen = __any_sync(0xffffffffu, nan_seen);
            float *dst = nan_seen ? real_hi : real_prev;  // signal.cpp:266-275: prev only advances on a sent frame
            __syncwarp();
            for (int t = lane; t < h; t += 32) {
                const float hi = (mode == MODE_USB) ? x[h + t].x : x[n - 1 - (h + t)].x;
                dst[t] = sg * hi;
            }
            if (lane == 0) ca.hi_diverged[slot] = nan_seen ? 1 : 0;
        } else {
            // signal.cpp:200-263
            const float2 prev_last = ca.bb_last[slot];
            float2 *bb = y;  // assemble the overlapped first half in the free buffer
            for (int t = lane; t < h; t += 32) {
                const float2 lo = x[t], old = bb_hi[t];
                bb[t] = make_float2(__fadd_rn(sg * lo.x, old.x), __fadd_rn(sg * lo.y, old.y));
            }
            __syncwarp();
            for (int t = lane; t < h; t += 32) {
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
            nan_seen = __any_sync(0xffffffffu, nan_seen);
            if (lane == 0) ca.bb_last[slot] = bb[h - 1];
            if (!nan_seen && ca.hi_diverged[slot]) {
                // audio_real_prev <- audio_real[n/2..n) left behind by a NaN-dropped SSB frame
                for (int t = lane; t < h; t += 32) real_prev[t] = real_hi[t];
                __syncwarp();
                if (lane == 0) ca.hi_diverged[slot] = 0;
            }
        }
        if (lane == 0) {
            ca.valid_a[(size_t)f * ca.max_clients + slot] = nan_seen ? 0 : 1;
            ca.pwr[(size_t)f * ca.max_clients + slot] = pw;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Sequential tails, one lane per client, 32 clients per block (8 warps help with tile movement).
// ------------------------------------------------------------------------------------------------
constexpr int kTailThreads = 256;
// tiles are [j][client] with pitch cpb + 1 (odd): conflict-free for lanes along j and lanes along client

__device__ __forceinline__ int floordiv(long long a, int b) {
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return (int)q;
}

__global__ void __launch_bounds__(kTailThreads) client_tail_kernel(const ClientArrays ca, const ClientLaunch cl) {
    extern __shared__ float smem_t[];
    const int h = ca.h, D = ca.D, L = ca.L, NC = ca.NC;
    const int cpb = cl.cpb, kPitch = cl.cpb + 1;
    float *tA = smem_t;                 // [h][33]  audio in, later AGC output
    float *tM = tA + h * kPitch;        // [h][33]  first-stage moving averages
    float *tY = tM + h * kPitch;        // [h][33]  DC blocker output
    float *tO = tY + h * kPitch;        // [2h][33] the two oldest AGC chunks touching the window
    float *tS = tO + 2 * h * kPitch;    // [2h][33] suffix maxima of |tO| within each chunk
    float *tDx = tS + 2 * h * kPitch;   // [D][33]
    float *tDm = tDx + D * kPitch;      // [D][33]
    __shared__ int s_slot[32];
    __shared__ int s_ca[32];
    __shared__ unsigned char s_valid[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g0 = blockIdx.x * cpb;
    if (tid < 32) s_slot[tid] = (tid < cpb && g0 + tid < cl.nactive) ? cl.order[g0 + tid] : -1;
    __syncthreads();

    // per-lane (= per-client) scalars live in warp 0's registers across frames
    float sum1 = 0.f, sum2 = 0.f, gain = 0.f;
    long long t0 = 0;
    const int my_slot = s_slot[lane];
    if (warp == 0 && my_slot >= 0) {
        const int fl = ca.slots[my_slot].flags;
        if (fl & (CF_RESET_ALL | CF_RESET_AGC)) {  // AGC::reset, audioprocessing.cpp:70-74
            float *ring = ca.agc_ring + (size_t)my_slot * NC * h;
            for (int i = 0; i < NC * h; i++) ring[i] = 0.f;
            for (int i = 0; i < NC; i++) ca.agc_cmax[(size_t)my_slot * NC + i] = 0.f;
            gain = 0.f;
            t0 = 0;
        } else {
            gain = ca.agc_gain[my_slot];
            t0 = ca.agc_t0[my_slot];
        }
        if (fl & CF_RESET_ALL) {
            for (int i = 0; i < D; i++) {
                ca.dc_x[(size_t)my_slot * D + i] = 0.f;
                ca.dc_m[(size_t)my_slot * D + i] = 0.f;
            }
            sum1 = sum2 = 0.f;
        } else {
            sum1 = ca.dc_sum[2 * my_slot];
            sum2 = ca.dc_sum[2 * my_slot + 1];
        }
    }
    __syncthreads();

    for (int f = 0; f < cl.nframes; f++) {
        if (warp == 0) {
            int v = 0, cidx = 0;
            if (my_slot >= 0) {
                v = ca.valid_a[(size_t)f * ca.max_clients + my_slot];
                cidx = floordiv(t0 - L + 1, h);
            }
            s_valid[lane] = (unsigned char)v;
            s_ca[lane] = cidx;
        }
        __syncthreads();
        // ---- tile loads (lanes along j: coalesced) ----
        for (int ci = warp; ci < cpb; ci += kTailThreads / 32) {
            const int slot = s_slot[ci];
            if (slot < 0 || !s_valid[ci]) continue;
            const float *a = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
            for (int j = lane; j < h; j += 32) tA[j * kPitch + ci] = a[j];
            for (int j = lane; j < D; j += 32) {
                tDx[j * kPitch + ci] = ca.dc_x[(size_t)slot * D + j];
                tDm[j * kPitch + ci] = ca.dc_m[(size_t)slot * D + j];
            }
            const int c0 = s_ca[ci];
            const float *ring = ca.agc_ring + (size_t)slot * NC * h;
            for (int rr = 0; rr < 2; rr++) {
                const int row = (((c0 + rr) % NC) + NC) % NC;
                const float *src = ring + (size_t)row * h;
                // suffix max of |x| inside the chunk, 32 at a time from the end
                float carry = 0.f;
                for (int base = ((h - 1) / 32) * 32; base >= 0; base -= 32) {
                    const int j = base + lane;
                    const float xv = (j < h) ? src[j] : 0.f;
                    float m = fabsf(xv);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float other = __shfl_down_sync(0xffffffffu, m, o);
                        if (lane + o < 32) m = fmaxf(m, other);
                    }
                    m = fmaxf(m, carry);
                    if (j < h) {
                        tO[(rr * h + j) * kPitch + ci] = xv;
                        tS[(rr * h + j) * kPitch + ci] = m;
                    }
                    carry = __shfl_sync(0xffffffffu, m, 0);
                }
            }
        }
        __syncthreads();
        // ---- sequential chains, lane = client ----
        if (warp == 0 && my_slot >= 0 && s_valid[lane]) {
            const int ci = lane;
            const float Df = (float)D;
            // DC blocker: src/utils.h:80-85,145-149.  xe = [dc_x | audio], me = [dc_m | ma1]
            for (int j = 0; j < h; j++) {
                const float xin = tA[j * kPitch + ci];
                const float old1 = (j < D) ? tDx[j * kPitch + ci] : tA[(j - D) * kPitch + ci];
                sum1 = __fadd_rn(__fadd_rn(sum1, -old1), xin);
                const float ma1 = __fdiv_rn(sum1, Df);
                tM[j * kPitch + ci] = ma1;
                const float old2 = (j < D) ? tDm[j * kPitch + ci] : tM[(j - D) * kPitch + ci];
                sum2 = __fadd_rn(__fadd_rn(sum2, -old2), ma1);
                const float ma2 = __fdiv_rn(sum2, Df);
                const float xd = (j + 1 < D) ? tDx[(j + 1) * kPitch + ci] : tA[(j + 1 - D) * kPitch + ci];
                tY[j * kPitch + ci] = __fsub_rn(xd, ma2);
            }
            // AGC: audioprocessing.cpp:40-68 with the sliding |x| maximum taken from chunk maxima
            const int c0 = s_ca[ci];
            const int F = (int)(t0 / h);
            const float *cmax = ca.agc_cmax + (size_t)my_slot * NC;
            float mfull[2];
            for (int rr = 0; rr < 2; rr++) {
                float m = 0.f;
                for (int c = c0 + rr + 1; c <= F - 1; c++) m = fmaxf(m, cmax[((c % NC) + NC) % NC]);
                mfull[rr] = m;
            }
            float pmax = 0.f;
            for (int j = 0; j < h; j++) {
                const float yv = tY[j * kPitch + ci];
                pmax = fmaxf(pmax, fabsf(yv));
                const long long lo = t0 + j - L + 1;  // oldest sample in the window
                float outv = 0.f;
                if (t0 + j + 1 >= L) {
                    const int c = floordiv(lo, h);
                    const int col = (int)(lo - (long long)c * h);
                    const int rr = c - c0;  // 0 or 1
                    const float cur = tO[(rr * h + col) * kPitch + ci];
                    const float peak = fmaxf(fmaxf(tS[(rr * h + col) * kPitch + ci], mfull[rr]), pmax);
                    const float desired = __fdiv_rn(ca.desired, __fadd_rn(peak, 1e-10f));
                    if (desired < gain) gain = __fsub_rn(gain, __fmul_rn(ca.attack, __fsub_rn(gain, desired)));
                    else gain = __fadd_rn(gain, __fmul_rn(ca.release, __fsub_rn(desired, gain)));
                    outv = __fmul_rn(cur, gain);
                }
                tA[j * kPitch + ci] = outv;
            }
            ca.agc_cmax[(size_t)my_slot * NC + (F % NC)] = pmax;
            t0 += h;
        }
        __syncthreads();
        // ---- write back (lanes along j) ----
        for (int ci = warp; ci < cpb; ci += kTailThreads / 32) {
            const int slot = s_slot[ci];
            if (slot < 0) continue;
            int *pcm = ca.pcm + ((size_t)f * ca.max_clients + slot) * h;
            if (!s_valid[ci]) {
                for (int j = lane; j < h; j += 32) pcm[j] = 0;
                if (lane == 0) ca.valid[(size_t)f * ca.max_clients + slot] = 0;
                continue;
            }
            if (lane == 0) ca.valid[(size_t)f * ca.max_clients + slot] = 1;
            for (int j = lane; j < h; j += 32) {
                // dsp.cpp:152-165 with mult = 65536/4
                const float x = tA[j * kPitch + ci];
                const float t = __fadd_rn(__fmul_rn(x, 16384.f), 32768.5f);
                int v = __float2int_rz(t) - 32768;
                v = max(min(v, 32767), -32768);
                pcm[j] = v;
            }
            // DC state: last D of [dc_x | audio_in] and [dc_m | ma1]; audio_in was overwritten in tA by
            // the AGC output, so re-read it from audio_pre (L2-resident)
            const float *a = ca.audio_pre + ((size_t)f * ca.max_clients + slot) * h;
            for (int j = lane; j < D; j += 32) {
                const int src = h + j;  // index into the (D + h)-long concatenation, minus D offset below
                float nx, nm;
                if (src >= D) {
                    nx = a[src - D];
                    nm = tM[(src - D) * kPitch + ci];
                } else {
                    nx = tDx[src * kPitch + ci];
                    nm = tDm[src * kPitch + ci];
                }
                ca.dc_x[(size_t)slot * D + j] = nx;
                ca.dc_m[(size_t)slot * D + j] = nm;
            }
        }
        // AGC ring row of this frame: needs each client's chunk index -> done by a second sweep
        __syncthreads();
        if (warp == 0) s_ca[lane] = (my_slot >= 0 && s_valid[lane]) ? (int)(((t0 / h) - 1) % NC) : -1;
        __syncthreads();
        for (int ci = warp; ci < cpb; ci += kTailThreads / 32) {
            const int slot = s_slot[ci];
            const int row = s_ca[ci];
            if (slot < 0 || row < 0) continue;
            float *dst = ca.agc_ring + ((size_t)slot * NC + row) * h;
            for (int j = lane; j < h; j += 32) dst[j] = tY[j * kPitch + ci];
        }
        __syncthreads();
    }
    if (warp == 0 && my_slot >= 0) {
        ca.agc_gain[my_slot] = gain;
        ca.agc_t0[my_slot] = t0;
        ca.dc_sum[2 * my_slot] = sum1;
        ca.dc_sum[2 * my_slot + 1] = sum2;
    }
}

}  // namespace b200
