// Forward spectrum path: fused {sample convert, Hann window, four-step FFT pass 1}, FFT pass 2 with
// 1/N normalisation + wrap tail + optional NVLink peer stores, and the waterfall quantiser/pyramid.
//
// Replaces FFTW::load_*_input + fftwf_execute + power_and_quantize + half_and_quantize
// (reference src/fft_impl.cpp:119-174) and cuFFT's window_* / power_and_quantize /
// half_and_quantize kernels (src/fft_cuda.cu:62-130).
//
// Decomposition of the M-point complex transform (M = size for c2c, size/2 for r2c packing):
//   M = N1*N2,  n = N2*n1 + n2,  k = k1 + N1*k2
//   pass 1: per column n2, N1-point DFT over n1, times W_M^(n2*k1)      -> Y[u1][n2]
//   pass 2: per row   k1, N2-point DFT over n2                          -> X[k1 + N1*k2]
// Each sub-DFT of S = RA*RB points runs on max(RA,RB) threads: an RA-point register DFT, one
// shared-memory exchange, an RB-point register DFT. A CTA carries T adjacent columns (rows) so that
// every global access is a run of T consecutive complex values (T*8 bytes).
//
// IQ display shift (src/fft_impl.cpp:148-160, base_idx = N/2+1): rows of Y are stored by
// u1 = (k1 - shift) mod N1 and, for IQ, the k1 = 0 row is pre-rotated by W_N2^(n2) so that slot u2 of
// that row holds k2 = u2+1. With u = u1 + N1*u2 the bin is k = (u + shift) mod M everywhere.
#pragma once
#include <cstdint>
#include "regfft.cuh"

namespace b200 {

enum { FMT_F32 = 0, FMT_U8 = 1, FMT_S8 = 2, FMT_U16 = 3, FMT_S16 = 4 };

constexpr int kMaxPeers = 8;

struct FwdParams {
    const void *ring;        // device hop ring, nhops hops
    size_t hop_bytes;        // bytes per hop in the ring
    int nhops;
    int hop0;                // frame f is made of hops (hop0+f) % nhops and (hop0+f+1) % nhops
    int in_format;           // FMT_*
    const float *window;     // size floats (Hann, host-built)
    float2 *Y;               // [frames][N1][N2] intermediate
    float2 *out;             // pass 2 output: spectrum (c2c) or Z scratch (r2c)
    size_t out_stride;       // float2 per frame in `out`
    int log2M, N1, N2;
    int shift;               // 1 = IQ display shift, 0 = none
    int is_real;
    float scale;             // applied by pass 2 (c2c: 1/size; r2c: 1)
    int additional;          // wrap tail bins (c2c)
    const float2 *twA1;      // [RA][RB] W_N1^(r*q)
    const float2 *twA2;      // [RA][RB] W_N2^(r*q)
    const float2 *TL;        // W_M^j, j < 1024
    const float2 *TH;        // W_M^(1024 j)
    int npeers;
    float2 *peers[kMaxPeers];  // extra spectrum destinations (NVLink peer memory)
    // bins each peer needs (its clients' sub-band): up to two half-open ranges of spectrum indices, tail included
    unsigned peer_lo[kMaxPeers][2], peer_hi[kMaxPeers][2];
    // fused waterfall epilogue of pass 2 (c2c): levels 0..log2(T)-1 straight from the FFT registers
    int8_t *quant;           // pyramid [frames][pyr_stride]
    size_t pyr_stride;
    float *pscratch;         // [frames][N1/T][N2] power sums of level log2(T)
    int levels;
    int size_log2;
    // transforms longer than 2^20 points: M = na * Mb. radix_split_kernel leaves na sub-sequences of Mb points in `pre`
    // (windowed, radix-na butterflies and W_M twiddles applied); passes 1 and 2 then run na sub-transforms per frame
    // and pass 2 interleaves their bins, k = ka + na * k_sub
    const float2 *pre;       // [frames][na][Mb]
    int na;                  // 1 = plain two-pass transform
    // r2c on the TMA path: Hann weights built on the fly from (h cos, h sin)(2 pi row / 1024), h = 1/2, and W_size^(2 n2 + b)
    const float2 *winT;
    const float2 *TLr, *THr; // W_size^j, j < 1024; W_size^(1024 j)
};

template <int A, int B> struct CMax { static constexpr int v = A > B ? A : B; };

__device__ __forceinline__ float2 load_sample(const void *hop, int fmt, size_t e) {
    // element e of a hop as a float pair (IQ sample, or two consecutive real samples)
    if (fmt == FMT_F32) return reinterpret_cast<const float2 *>(hop)[e];
    if (fmt == FMT_U8) {  // src/samplereader.cpp:29-40: (x ^ 0x80) as int8, / 128
        uchar2 b = reinterpret_cast<const uchar2 *>(hop)[e];
        return make_float2((float)(int8_t)(b.x ^ 0x80) / 128.0f, (float)(int8_t)(b.y ^ 0x80) / 128.0f);
    }
    if (fmt == FMT_S8) {
        char2 b = reinterpret_cast<const char2 *>(hop)[e];
        return make_float2((float)b.x / 128.0f, (float)b.y / 128.0f);
    }
    if (fmt == FMT_U16) {
        ushort2 b = reinterpret_cast<const ushort2 *>(hop)[e];
        return make_float2((float)(int16_t)(b.x ^ 0x8000) / 32768.0f, (float)(int16_t)(b.y ^ 0x8000) / 32768.0f);
    }
    short2 b = reinterpret_cast<const short2 *>(hop)[e];
    return make_float2((float)b.x / 32768.0f, (float)b.y / 32768.0f);
}

// ------------------------------------------------------------------------------------------------
// pass 1: grid (N2/T, frames), block T*max(RA,RB)
// ------------------------------------------------------------------------------------------------
template <int RA, int RB, int T, bool RAW, bool REAL, bool PRE = false>
__global__ void __launch_bounds__(T *CMax<RA, RB>::v) fft_pass1_kernel(const FwdParams p) {
    // the IQ display shift exists only for c2c (fft_impl.cpp:148-150) and only when pass 2 owns display-aligned runs
    constexpr int SHIFT = (REAL || PRE) ? 0 : 1;
    constexpr int PAD = (T < 16) ? (16 - T) : 0;   // keep the two r-rows of a half-warp on disjoint banks
    constexpr int ROW = RA * T + PAD;
    extern __shared__ float2 sm[];

    const int tid = threadIdx.x;
    const int c = tid % T;
    const int r = tid / T;
    const int frame = blockIdx.y;
    const int n2 = blockIdx.x * T + c;
    const int N2 = p.N2;
    const size_t M = (size_t)1 << p.log2M;
    const size_t half = M >> 1;

    const char *ring = reinterpret_cast<const char *>(p.ring);
    const void *hopA = ring + (size_t)((p.hop0 + frame) % p.nhops) * p.hop_bytes;
    const void *hopB = ring + (size_t)((p.hop0 + frame + 1) % p.nhops) * p.hop_bytes;

    if (r < RB) {
        float2 v[RA];
        const int fmt = RAW ? p.in_format : FMT_F32;  // float32 input: a compile-time constant, loads batch freely
#pragma unroll
        for (int j = 0; j < RA; j++) {
            const size_t idx = (size_t)(r + RB * j) * N2 + n2;  // complex element within the frame
            // first half of the frame (j < RA/2) comes from the older hop
            float2 x;
            if constexpr (PRE) x = p.pre[(size_t)frame * M + idx];  // sub-sequence `frame` of radix_split_kernel, already windowed
            else x = (j < RA / 2) ? load_sample(hopA, fmt, idx) : load_sample(hopB, fmt, idx - half);
            if constexpr (PRE) {
            } else if constexpr (REAL) {
                float2 w = __ldg(reinterpret_cast<const float2 *>(p.window) + idx);
                x.x *= w.x;
                x.y *= w.y;
            } else {
                float w = __ldg(p.window + idx);
                x.x *= w;
                x.y *= w;
            }
            v[j] = x;
        }
        RegDft<RA>::run(v);
#pragma unroll
        for (int q = 0; q < RA; q++) {
            float2 t = (q == 0) ? make_float2(1.f, 0.f) : __ldg(p.twA1 + q * RB + r);
            sm[r * ROW + q * T + c] = (q == 0) ? v[q] : cmul(v[q], t);
        }
    }
    __syncthreads();
    if (r < RA) {
        const int q = r;
        float2 u[RB];
#pragma unroll
        for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
        RegDft<RB>::run(u);
        float2 *Y = p.Y + (size_t)frame * M;
        const int N1 = p.N1;
#pragma unroll
        for (int s = 0; s < RB; s++) {
            const int k1 = q + RA * s;
            int u1 = k1 - SHIFT;
            if (u1 < 0) u1 += N1;
            // exponent of W_M: n2*k1, or the one-slot rotation N1*n2 for the IQ k1 = 0 row
            const unsigned e = (SHIFT && k1 == 0) ? (unsigned)N1 * (unsigned)n2 : (unsigned)n2 * (unsigned)k1;
            const float2 tw = cmul(__ldg(p.TL + (e & 1023u)), __ldg(p.TH + (e >> 10)));
            Y[(size_t)u1 * N2 + n2] = cmul(u[s], tw);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// transforms longer than 2^20 points, M = NA * Mb with n = Mb*na + nb and k = ka + NA*kb:
//   A_ka[nb] = W_M^(nb*ka) * sum_na w[n] x[n] W_NA^(na*ka)        (this kernel: fully coalesced, window + sample
//   X[ka + NA*kb] = DFT_Mb(A_ka)[kb]                               conversion fused; then NA two-pass sub-transforms)
// grid (Mb / 256, frames), block 256: one thread per nb.
// ------------------------------------------------------------------------------------------------
template <int NA, bool REAL>
__global__ void __launch_bounds__(256) radix_split_kernel(const FwdParams p, float2 *pre, const float2 *TLM, const float2 *THM,
                                                          int log2Mb) {
    const size_t Mb = (size_t)1 << log2Mb;
    const size_t nb = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int frame = blockIdx.y;
    const size_t half = Mb * NA / 2;
    const char *ring = reinterpret_cast<const char *>(p.ring);
    const void *hopA = ring + (size_t)((p.hop0 + frame) % p.nhops) * p.hop_bytes;
    const void *hopB = ring + (size_t)((p.hop0 + frame + 1) % p.nhops) * p.hop_bytes;
    float2 v[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) {
        const size_t n = Mb * a + nb;
        float2 x = (a < NA / 2) ? load_sample(hopA, p.in_format, n) : load_sample(hopB, p.in_format, n - half);
        if constexpr (REAL) {
            const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window) + n);
            x.x *= w.x;
            x.y *= w.y;
        } else {
            const float w = __ldg(p.window + n);
            x.x *= w;
            x.y *= w;
        }
        v[a] = x;
    }
    RegDft<NA>::run(v);
    float2 *dst = pre + (size_t)frame * NA * Mb + nb;
    dst[0] = v[0];
#pragma unroll
    for (int ka = 1; ka < NA; ka++) {
        const unsigned e = (unsigned)nb * (unsigned)ka;  // < 2^23
        dst[(size_t)ka * Mb] = cmul(v[ka], cmul(__ldg(TLM + (e & 1023u)), __ldg(THM + (e >> 10))));
    }
}

// ------------------------------------------------------------------------------------------------
// waterfall quantiser - bit-exact restatement of vec_log2 / power_and_quantize (src/fft_impl.cpp:14-44).
// Every float op is an explicit round-to-nearest intrinsic (no FMA contraction) in source order.
// ------------------------------------------------------------------------------------------------
// The exponent reaches the float domain without an I2F (the conversion pipe is quarter-rate): 0x4B000000 | e is the
// float 2^23 + e, and (2^23 + e) - (2^23 + 128 - offset) is exact, so log_val has the bits of
// (float)(e - 128) + (float)offset (fft_impl.cpp:16-17) for every e and every offset in range.
__device__ __forceinline__ float quant_bias(int power_offset) { return (float)(8388608 + 128 - power_offset); }
// (a & b) | c in one LOP3 with both masks in registers (as immediates ptxas needs two instructions)
__device__ __forceinline__ unsigned and_or(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ float exponent_as_float(unsigned bits) { return __uint_as_float(and_or(bits >> 23, 0xFFu, 0x4B000000u)); }
// exponent field replaced by 127; the sign bit stays, as in the reference (`*bit_exponent &= ~(255 << 23)`, fft_impl.cpp:19)
__device__ __forceinline__ float mantissa_1_2(unsigned bits) { return __uint_as_float(and_or(bits, 0x807FFFFFu, 0x3F800000u)); }
__device__ __forceinline__ float vec_log2_biased(float val, float bias) {
    const unsigned bits = __float_as_uint(val);
    const float log_val = __fsub_rn(exponent_as_float(bits), bias);
    val = mantissa_1_2(bits);  // mantissa forced to [1, 2)
    float poly = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(-0.34484843f, val), 2.02466578f), val), 0.67487759f);
    return __fadd_rn(log_val, poly);
}
__device__ __forceinline__ float vec_log2_dev(float val, int power_offset) { return vec_log2_biased(val, quant_bias(power_offset)); }
__device__ __forceinline__ int quantize_biased(float power, float bias) {
    float v = __fadd_rn(__fmul_rn(__fmul_rn(vec_log2_biased(power, bias), 0.3010299956639812f), 20.f), 127.f);
    v = fmaxf(v, -128.f);                 // std::max(-128.f, v); NaN -> -128
    return __float2int_rz(v);             // C truncation; the int8 store keeps the low byte (wraps above 127 like x86)
}
__device__ __forceinline__ int quantize_dev(float power, int power_offset) {
    return quantize_biased(power, quant_bias(power_offset)) & 0xFF;
}
// Two bins per instruction on the sm_100 packed-f32 pipe, keeping the reference's rounding: ptxas contracts every
// mul.f32x2 -> add.f32x2 chain into FFMA2 (one rounding instead of two, even with explicit .rn and -fmad=false), so an
// add that consumes a product is written as fma(product, 1, c) - round(round(a b) * 1 + c) is exactly the separately
// rounded add, and with its multiplier slot taken ptxas leaves the FMUL2 in front of it alone (checked in the SASS:
// FMUL2 followed by FFMA2 ..., R.F32 (= 1.0), c).
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long pk_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pk_add(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pk_add_scalar(unsigned long long a, float c) {  // never fused with a producer
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(pk(1.f, 1.f)), "l"(pk(c, c)));
    return d;
}
__device__ __forceinline__ void quantize2_biased(float p0, float p1, float bias, int &t0, int &t1) {
    const unsigned b0 = __float_as_uint(p0), b1 = __float_as_uint(p1);
    const unsigned long long lv = pk_add(pk(exponent_as_float(b0), exponent_as_float(b1)), pk(-bias, -bias));
    const unsigned long long m = pk(mantissa_1_2(b0), mantissa_1_2(b1));
    unsigned long long t = pk_mul(pk(-0.34484843f, -0.34484843f), m);
    t = pk_add_scalar(t, 2.02466578f);
    t = pk_mul(t, m);
    t = pk_add_scalar(t, -0.67487759f);
    unsigned long long v = pk_add(lv, t);  // both operands are sums: nothing to contract
    v = pk_mul(v, pk(0.3010299956639812f, 0.3010299956639812f));
    v = pk_mul(v, pk(20.f, 20.f));
    v = pk_add_scalar(v, 127.f);
    float vx, vy;
    unpk(v, vx, vy);
    t0 = __float2int_rz(fmaxf(vx, -128.f));
    t1 = __float2int_rz(fmaxf(vy, -128.f));
}
// Table-driven form (opt-in, B200_OPT_PACKED_MATH bit 1). Per power offset the quantiser is a step function of the 31
// non-sign bits of the power with at most one step per 1/8 octave, except for a band of a few ulps around each step
// where rounding makes it wiggle. tab[bits >> 20] = {lo, base | width << 8}: base below lo, base + 1 from lo + width on,
// the exact arithmetic inside the band. Built by build_quant_table (engine.cu) with the reference arithmetic;
// tools/quant_table.c checks the construction against all 2^31 inputs.
__device__ __forceinline__ int quantize_table(float power, const uint2 *tab, float bias) {
    const unsigned raw = __float_as_uint(power), b = raw & 0x7FFFFFFFu;
    const uint2 e = __ldg(tab + (b >> 20));
    const unsigned width = e.y >> 8;
    int q = (int)(e.y & 0xFFu) + (b >= e.x + width ? 1 : 0);
    // exact arithmetic inside the band (a handful of inputs per offset) and for a set sign bit (never for |X|^2; the
    // reference's polynomial would see a negative mantissa)
    if ((b - e.x < width) | (raw >> 31)) q = quantize_biased(power, bias);
    return q;
}
// CNT bins of one level -> little-endian packed bytes in w[(CNT+3)/4]
template <int CNT, bool PK>
__device__ __forceinline__ void quantize_pack(const float *pw, float bias, unsigned *w, const uint2 *tab = nullptr) {
    int t[CNT];
    if (tab) {
#pragma unroll
        for (int i = 0; i < CNT; i++) t[i] = quantize_table(pw[i], tab, bias);
    } else if constexpr (PK && CNT >= 2) {
#pragma unroll
        for (int i = 0; i < CNT; i += 2) quantize2_biased(pw[i], pw[i + 1], bias, t[i], t[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; i++) t[i] = quantize_biased(pw[i], bias);
    }
    if constexpr (CNT >= 4) {
#pragma unroll
        for (int i = 0; i < CNT / 4; i++)
            w[i] = __byte_perm(__byte_perm(t[4 * i], t[4 * i + 1], 0x0040), __byte_perm(t[4 * i + 2], t[4 * i + 3], 0x0040), 0x5410);
    } else if constexpr (CNT == 2) {
        w[0] = __byte_perm(t[0], t[1], 0x0040) & 0xFFFFu;
    } else {
        w[0] = (unsigned)t[0] & 0xFFu;
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: grid (N1/T, frames), block T*max(RA,RB)
// ------------------------------------------------------------------------------------------------
template <int NB> __device__ __forceinline__ void store_packed(int8_t *dst, const unsigned *w) {
    // NB quantised bytes packed little-endian in w[]; dst is NB-aligned
    if constexpr (NB >= 16) {
#pragma unroll
        for (int i = 0; i < NB / 16; i++)
            reinterpret_cast<uint4 *>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else if constexpr (NB == 8) {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    } else if constexpr (NB == 4) {
        *reinterpret_cast<unsigned *>(dst) = w[0];
    } else if constexpr (NB == 2) {
        *reinterpret_cast<unsigned short *>(dst) = (unsigned short)w[0];
    } else {
        *dst = (int8_t)w[0];
    }
}

template <int T> struct Log2 { static constexpr int v = 1 + Log2<T / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// FUSE: 0 = spectrum only; 1 = whole waterfall epilogue (levels 0..log2(T)-1 quantised here); 2 = |X|^2 stored
// straight from the FFT registers into the power scratch [u2][u1], quantiser + pyramid in pyramid_kernel<PYR_POWER>
template <int RA, int RB, int T, int FUSE>
__global__ void __launch_bounds__(T *CMax<RA, RB>::v) fft_pass2_kernel(const FwdParams p) {
    constexpr int TPC = CMax<RA, RB>::v;
    constexpr int ROW = RA * T + 1;  // odd stride: lanes along r hit distinct banks
    constexpr int PP = T + 4;        // pitch of the power tile (floats): float4 reads stay conflict-free
    extern __shared__ float2 sm[];

    const int tid = threadIdx.x;
    const int na = p.na;                         // sub-transforms per frame (1 = plain)
    const int frame = blockIdx.y / na;
    const int ka = blockIdx.y - frame * na;
    const int N1 = p.N1, N2 = p.N2;
    const size_t M = (size_t)1 << p.log2M;       // sub-transform length
    const float2 *Y = p.Y + (size_t)blockIdx.y * M;

    {   // stage A: lanes along n2 (contiguous in Y)
        const int r = tid % TPC;
        const int c = tid / TPC;
        if (r < RB) {
            const float2 *src = Y + (size_t)(blockIdx.x * T + c) * N2;
            float2 v[RA];
#pragma unroll
            for (int j = 0; j < RA; j++) v[j] = src[r + RB * j];
            RegDft<RA>::run(v);
#pragma unroll
            for (int q = 0; q < RA; q++) {
                float2 t = (q == 0) ? make_float2(1.f, 0.f) : __ldg(p.twA2 + q * RB + r);
                sm[r * ROW + q * T + c] = (q == 0) ? v[q] : cmul(v[q], t);
            }
        }
    }
    __syncthreads();
    float pw[RB];
    {   // stage B: lanes along u1 (contiguous in the spectrum)
        const int c = tid % T;
        const int q = tid / T;
        if (q < RA) {
            float2 u[RB];
#pragma unroll
            for (int rr = 0; rr < RB; rr++) u[rr] = sm[rr * ROW + q * T + c];
            RegDft<RB>::run(u);
            float2 *out = p.out + (size_t)frame * p.out_stride;
            const unsigned u1 = blockIdx.x * T + c;
            const float scale = p.scale;
#pragma unroll
            for (int s = 0; s < RB; s++) {
                const unsigned u2 = q + RA * s;
                const size_t ks = ((size_t)u1 + (size_t)N1 * u2 + p.shift) & (M - 1);
                const size_t k = (size_t)ka + (size_t)na * ks;  // bin of the full transform
                const float2 val = make_float2(u[s].x * scale, u[s].y * scale);
                out[k] = val;
                if (k < (size_t)p.additional) out[M * na + k] = val;  // IQ wrap tail, src/fft.cpp:96-97
                if constexpr (FUSE == 1) pw[s] = __fadd_rn(__fmul_rn(val.x, val.x), __fmul_rn(val.y, val.y));
                if constexpr (FUSE == 2)
                    p.pscratch[(size_t)frame * M + (size_t)u2 * N1 + u1] =
                        __fadd_rn(__fmul_rn(val.x, val.x), __fmul_rn(val.y, val.y));
            }
            if (p.npeers > 0) {  // NVLink peer copies of the frame (multi-GPU ingest rank only), off the common path
                for (int pe = 0; pe < p.npeers; pe++) {
                    float2 *po = p.peers[pe] + (size_t)frame * p.out_stride;
#pragma unroll
                    for (int s = 0; s < RB; s++) {
                        const unsigned ks = (unsigned)(((size_t)u1 + (size_t)N1 * (q + RA * s) + p.shift) & (M - 1));
                        const unsigned k = (unsigned)ka + (unsigned)na * ks;
                        const float2 val = make_float2(u[s].x * scale, u[s].y * scale);
                        if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1]))
                            po[k] = val;
                        const unsigned kt = (unsigned)(M * na) + k;
                        if (k < (unsigned)p.additional &&
                            ((kt >= p.peer_lo[pe][0] && kt < p.peer_hi[pe][0]) || (kt >= p.peer_lo[pe][1] && kt < p.peer_hi[pe][1])))
                            po[kt] = val;
                    }
                }
            }
        }
    }
    if constexpr (FUSE == 1) {
        // Waterfall epilogue (src/fft_impl.cpp:24-61,146-173): |X|^2 of this CTA's T x N2 bins goes through
        // shared memory so that one thread owns one display run of T bins and writes whole words.
        float *pt = reinterpret_cast<float *>(sm);
        __syncthreads();  // everyone is done reading the exchange buffer
        {
            const int c = tid % T;
            const int q = tid / T;
            if (q < RA) {
#pragma unroll
                for (int s = 0; s < RB; s++) pt[(q + RA * s) * PP + c] = pw[s];
            }
        }
        __syncthreads();
        constexpr int LT = Log2<T>::v;
        const int L = p.levels;
        int8_t *quant = p.quant + (size_t)frame * p.pyr_stride;
        float *scr = p.pscratch + ((size_t)frame * (N1 / T) + blockIdx.x) * N2;
        for (int u2 = tid; u2 < N2; u2 += T * TPC) {
            // display bin d = u1 + N1 * ((u2 + N2/2) mod N2): the base_idx = N/2+1 shift of fft_impl.cpp:148-160
            const unsigned d2 = (u2 + (N2 >> 1)) & (N2 - 1);
            const size_t dbase = (size_t)blockIdx.x * T + (size_t)N1 * d2;
            float v[T];
#pragma unroll
            for (int i = 0; i < T / 4; i++) {
                const float4 f = *reinterpret_cast<const float4 *>(pt + u2 * PP + 4 * i);
                v[4 * i] = f.x;
                v[4 * i + 1] = f.y;
                v[4 * i + 2] = f.z;
                v[4 * i + 3] = f.w;
            }
            size_t lvl_off = 0;
            const size_t R = M;
            static_for<LT>([&](auto lvc) {
                constexpr int lv = decltype(lvc)::value;
                constexpr int CNT = T >> lv;
                if (lv < L) {
                    unsigned w[(CNT + 3) / 4];
#pragma unroll
                    for (int i = 0; i < (CNT + 3) / 4; i++) w[i] = 0;
#pragma unroll
                    for (int i = 0; i < CNT; i++)
                        w[i / 4] |= (unsigned)quantize_dev(v[i], p.size_log2 - lv) << (8 * (i % 4));
                    store_packed<CNT>(quant + lvl_off + (dbase >> lv), w);
                }
                lvl_off += R >> lv;
#pragma unroll
                for (int i = 0; i < CNT / 2; i++) v[i] = __fadd_rn(v[2 * i], v[2 * i + 1]);
            });
            scr[d2] = v[0];  // level LT power, consumed by pyramid_kernel<PYR_SCRATCH>
        }
    }
}

enum { PYR_SPEC = 0, PYR_R2C = 1, PYR_SCRATCH = 2, PYR_POWER = 3 };

struct PyrParams {
    float2 *spec;            // spectrum (PYR_SPEC: input, already normalised; PYR_R2C: OUTPUT written here)
    size_t spec_stride;
    const float2 *Z;         // PYR_R2C: packed half-size transform (unnormalised), [frames][M]
    int8_t *quant;           // pyramid output [frames][pyr_stride]
    size_t pyr_stride;
    float *ptop;             // [frames][(R >> base_level) / 1024] sums ten levels above the base (deep pyramids only)
    int log2R;               // display bins R = 1 << log2R
    int levels;
    int size_log2;           // round(log2(size)) + brightness_offset
    float scale;             // PYR_R2C: 1/size
    const float2 *TLr;       // PYR_R2C: W_size^j, j < 1024
    const float2 *THr;       // PYR_R2C: W_size^(1024 j)
    // PYR_SCRATCH: level `base_level` power sums left by the fused pass-2 epilogue, [frames][ntiles][N2],
    // logical index i = tile + ntiles * d2
    const float *pscratch;
    int base_level;
    int ntiles, N2;
    int npeers;
    float2 *peers[kMaxPeers];
    unsigned peer_lo[kMaxPeers][2], peer_hi[kMaxPeers][2];  // PYR_R2C: bins each peer needs (two half-open ranges), as FwdParams
    int frame0, frame_step;  // the launch covers frames frame0, frame0 + frame_step, ... (blockIdx.y-th of them)
    int wf_first, wf_skip;   // PYR_R2C (every frame is split): only frames wf_first, wf_first + wf_skip, ... get a pyramid
    int natural;             // PYR_SPEC: 1 = display bin d is FFT bin d (r2c), 0 = the IQ shift below
    const uint2 *qtab;       // optional [3][2048] quantiser tables of levels 0..2 (see quantize_table), or nullptr
};

// grid ((R >> base_level) / (256*PER), frames), block 256: each thread owns PER (4 or 16) consecutive entries of the
// base level; the block then reduces log2(PER) levels in registers, five with warp shuffles and three through
// shared memory (pairwise sums, src/fft_impl.cpp:45-61,162-172) - log2(PER) + 8 levels above the base in all.
struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// The pairwise-sum tree over one block of 256 * PER base-level entries whose powers are already in registers
// (pw[i] = entry (blk * 256 + tid) * PER + i of level B): quantises and stores every level the block covers.
template <int PER, bool PK, bool TB = false, typename Sync = CtaSync>
__device__ __forceinline__ void pyramid_tree(const PyrParams &p, const int frame, const unsigned blk, const int tid, float (&pw)[PER],
                                             const int B, float *warp_sum_s, Sync sync) {
    constexpr int LP = (PER == 16) ? 4 : 2;  // levels reduced in registers
    const unsigned R = 1u << p.log2R;
    const unsigned d0 = (blk * 256u + tid) * PER;  // index at level B
    int8_t *quant = p.quant + (size_t)frame * p.pyr_stride;
    const int L = p.levels - B;  // levels still to produce, counted from the base
    const int off = p.size_log2 - B;
    unsigned lvl_off = 0;        // byte offset of level B
    for (int i = 0; i < B; i++) lvl_off += R >> i;
    const unsigned RB_ = R >> B;   // entries at the base level
    if (L <= 0) return;
    // relative levels 0 .. LP in registers: level lv has PER >> lv values per thread, stored as one word/vector
    static_for<LP + 1>([&](auto lvc) {
        constexpr int lv = decltype(lvc)::value;
        constexpr int CNT = PER >> lv;
        if (lv < L) {
            unsigned w[(CNT + 3) / 4];
            if constexpr (TB)  // table-driven levels 0..2 (opt-in instantiation; the default kernels do not carry this path)
                quantize_pack<CNT, PK>(pw, quant_bias(off - lv), w, (p.qtab && B + lv < 3) ? p.qtab + 2048 * (B + lv) : nullptr);
            else
                quantize_pack<CNT, PK>(pw, quant_bias(off - lv), w);
            store_packed<CNT>(quant + lvl_off + (d0 >> lv), w);
        }
        lvl_off += RB_ >> lv;
        if constexpr (CNT > 1) {
#pragma unroll
            for (int i = 0; i < CNT / 2; i++) pw[i] = __fadd_rn(pw[2 * i], pw[2 * i + 1]);
        }
    });
    // pw[0] now holds the thread's level-LP sum (already written above)
    if (L > LP + 1) {
        float s = pw[0];
        const int lane = tid & 31;
#pragma unroll
        for (int k = 1; k <= 5; k++) {  // relative levels LP+1 .. LP+5 inside the warp
            const int lv = LP + k;
            if (lv < L) {
                s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1 << (k - 1)));
                if ((lane & ((1 << k) - 1)) == 0) quant[lvl_off + (d0 >> lv)] = (int8_t)quantize_dev(s, off - lv);
                lvl_off += RB_ >> lv;
            }
        }
        if (L > LP + 6) {
            if (lane == 0) warp_sum_s[tid >> 5] = s;
            sync();
            if (tid < 8) {
                float w = warp_sum_s[tid];
                const unsigned b0 = blk * 256u * PER;
#pragma unroll
                for (int k = 1; k <= 3; k++) {  // relative levels LP+6 .. LP+8 across the eight warps
                    const int lv = LP + 5 + k;
                    if (lv < L) {
                        w = __fadd_rn(w, __shfl_xor_sync(0xffu, w, 1 << (k - 1)));
                        if ((tid & ((1 << k) - 1)) == 0)
                            quant[lvl_off + ((b0 + 32u * PER * tid) >> lv)] = (int8_t)quantize_dev(w, off - lv);
                        lvl_off += RB_ >> lv;
                    }
                }
                if (L > LP + 9 && tid == 0) p.ptop[(size_t)frame * (RB_ / (256 * PER)) + blk] = w;
            }
        }
    }
}

// One block of 256 threads (a CTA of pyramid_kernel, or one consumer group of the fused FFT kernel): `blk` = block
// index within the frame, `tid` = 0..255, `warp_sum_s` = 8 floats of shared memory, `sync` = barrier of those 256
// threads. CG: spectrum loads bypass L1 (the fused kernel reads bins that other SMs stored during the same launch).
template <int MODE, int PER, bool PK, bool CG, bool TB = false, typename Sync = CtaSync>
__device__ __forceinline__ void pyramid_block(const PyrParams &p, const int frame, const unsigned blk, const int tid,
                                              float *warp_sum_s, Sync sync) {
    static_assert(PER == 4 || PER == 16, "PER must be 4 or 16");
    const unsigned R = 1u << p.log2R;  // bins and pyramid offsets fit 32 bits (R <= 2^23)
    const int B = (MODE == PYR_SCRATCH) ? p.base_level : 0;
    const unsigned d0 = (blk * 256u + tid) * PER;  // index at level B

    float pw[PER];
    if constexpr (MODE == PYR_SPEC) {
        // display bin d <-> FFT bin (d + R/2 + 1) mod R   (src/fft_impl.cpp:148-160)
        const float2 *spec = p.spec + (size_t)frame * p.spec_stride;
        const unsigned k0 = p.natural ? d0 : ((d0 + (R >> 1) + 1) & (R - 1));
        // d0 is a multiple of PER, so the thread's PER bins start at k0 == 1 (mod PER): with the engine's buffer offset
        // (bin 1 on a 128-byte line) that is an aligned, contiguous run - 128-bit loads, except for the one thread
        // whose run wraps past bin R-1 and for externally bound, unaligned buffers
        if (k0 + PER <= R && (reinterpret_cast<uintptr_t>(spec + k0) & 15) == 0) {
            const float4 *src = reinterpret_cast<const float4 *>(spec + k0);
#pragma unroll
            for (int i = 0; i < PER / 2; i++) {
                const float4 x = CG ? __ldcg(src + i) : src[i];
                if constexpr (PK) {  // (re^2, im^2) in one FMUL2, then the separately rounded sum
                    float a, b, c, d;
                    const unsigned long long v0 = pk(x.x, x.y), v1 = pk(x.z, x.w);
                    unpk(pk_mul(v0, v0), a, b);
                    unpk(pk_mul(v1, v1), c, d);
                    pw[2 * i] = __fadd_rn(a, b);
                    pw[2 * i + 1] = __fadd_rn(c, d);
                } else {
                    pw[2 * i] = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
                    pw[2 * i + 1] = __fadd_rn(__fmul_rn(x.z, x.z), __fmul_rn(x.w, x.w));
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const float2 *src = spec + ((k0 + i) & (R - 1));
                const float2 x = CG ? __ldcg(src) : *src;
                pw[i] = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
            }
        }
    } else if constexpr (MODE == PYR_R2C) {
        // r2c split: X[k] = (Z[k] + conj(Z[M-k]))/2 - (i/2) W_size^k (Z[k] - conj(Z[M-k])), M = R
        // The thread's 16 bins k = d0 .. d0+15 need Z[d0 .. d0+15] (one aligned 128-byte run) and the mirrored run
        // Z[R-d0-15 .. R-d0] (contiguous as well, but it starts on an odd element: 64-bit loads at both ends, 128-bit
        // in between). Twiddles: W^(d0+i) = W^d0 * W^i - one table lookup per thread, sixteen small constants per block.
        float2 *spec = p.spec + (size_t)frame * p.spec_stride;
        const float2 *Z = p.Z + (size_t)frame * R;
        float2 a[PER], b[PER];
        {
            const float4 *src = reinterpret_cast<const float4 *>(Z + d0);
#pragma unroll
            for (int i = 0; i < PER / 2; i++) {
                const float4 x = src[i];
                a[2 * i] = make_float2(x.x, x.y);
                a[2 * i + 1] = make_float2(x.z, x.w);
            }
        }
        if (d0 == 0) {  // the run that pairs with bins 0 .. 15 wraps: Z[0], Z[R-1], ..., Z[R-15]
            b[0] = Z[0];
#pragma unroll
            for (int i = 1; i < PER; i++) b[i] = Z[R - i];
        } else {
            const float2 *m = Z + (R - d0 - (PER - 1));  // ascending addresses m[0 .. PER-1] = Z[R-d0-15 .. R-d0], b[i] = m[PER-1-i]
            b[PER - 1] = m[0];
            const float4 *mid = reinterpret_cast<const float4 *>(m + 1);
#pragma unroll
            for (int i = 0; i < (PER - 2) / 2; i++) {
                const float4 x = mid[i];
                b[PER - 2 - 2 * i] = make_float2(x.x, x.y);
                b[PER - 3 - 2 * i] = make_float2(x.z, x.w);
            }
            b[0] = m[PER - 1];
        }
        const float2 w0 = cmul(__ldg(p.TLr + (d0 & 1023)), __ldg(p.THr + (d0 >> 10)));
        float2 xs[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const float2 e = make_float2(a[i].x + b[i].x, a[i].y - b[i].y);
            const float2 o = make_float2(a[i].x - b[i].x, a[i].y + b[i].y);
            const float2 w = (i == 0) ? w0 : cmul(w0, __ldg(p.TLr + i));
            const float2 t = cmul(o, w);
            float2 x = make_float2(0.5f * (e.x + t.y), 0.5f * (e.y - t.x));
            x.x *= p.scale;
            x.y *= p.scale;
            xs[i] = x;
            pw[i] = __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
        }
        if (d0 == 0) {
            // Nyquist bin is left unnormalised by the reference (only outbuf_len = size/2 bins are divided,
            // src/fft_impl.cpp:152-154)
            const float2 ny = make_float2(a[0].x - a[0].y, 0.f);
            spec[R] = ny;
            for (int pe = 0; pe < p.npeers; pe++)
                if ((R >= p.peer_lo[pe][0] && R < p.peer_hi[pe][0]) || (R >= p.peer_lo[pe][1] && R < p.peer_hi[pe][1]))
                    (p.peers[pe] + (size_t)frame * p.spec_stride)[R] = ny;
        }
        if ((reinterpret_cast<uintptr_t>(spec + d0) & 15) == 0) {
            float4 *dst = reinterpret_cast<float4 *>(spec + d0);
#pragma unroll
            for (int i = 0; i < PER / 2; i++) dst[i] = make_float4(xs[2 * i].x, xs[2 * i].y, xs[2 * i + 1].x, xs[2 * i + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < PER; i++) spec[d0 + i] = xs[i];
        }
        for (int pe = 0; pe < p.npeers; pe++) {
            float2 *ps = p.peers[pe] + (size_t)frame * p.spec_stride;
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const unsigned k = d0 + i;
                if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1])) ps[k] = xs[i];
            }
        }
    } else if constexpr (MODE == PYR_POWER) {
        // |X|^2 left by FFT pass 2 in [u2][u1] order: display bin d = u1 + N1 * ((u2 + N2/2) mod N2)
        const int N1 = p.ntiles, N2 = p.N2;  // (ntiles carries N1 in this mode)
        const unsigned u1 = d0 & (unsigned)(N1 - 1), d2 = d0 / (unsigned)N1;
        const unsigned u2 = (d2 + (N2 >> 1)) & (unsigned)(N2 - 1);
        const float4 *src = reinterpret_cast<const float4 *>(p.pscratch + (size_t)frame * R + (size_t)(u2 * N1 + u1));
#pragma unroll
        for (int i = 0; i < PER / 4; i++) {
            const float4 f = src[i];
            pw[4 * i] = f.x;
            pw[4 * i + 1] = f.y;
            pw[4 * i + 2] = f.z;
            pw[4 * i + 3] = f.w;
        }
    } else {
        const float *scr = p.pscratch + (size_t)frame * p.ntiles * p.N2;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const unsigned idx = d0 + i;
            const unsigned tile = idx % (unsigned)p.ntiles, d2 = idx / (unsigned)p.ntiles;
            pw[i] = scr[tile * p.N2 + d2];
        }
    }
    if constexpr (MODE == PYR_R2C) {
        // waterfall cadence (src/fft.cpp:33,102-104): the split above is needed every frame, the pyramid only on send frames
        if (p.wf_skip > 1 && (frame < p.wf_first || (frame - p.wf_first) % p.wf_skip != 0)) return;
    }
    pyramid_tree<PER, PK, TB>(p, frame, blk, tid, pw, B, warp_sum_s, sync);
}

// r2c Hermitian split on its own: X[k] for four bins per thread, the arithmetic of the PYR_R2C branch above expression for
// expression (same twiddle decomposition W^(16 j) * W^i), so the spectrum is bit-identical; a streaming kernel at 40
// registers whose loads, unlike those of the 80-register split-and-quantise kernel, are hidden by occupancy. The pyramid
// then comes from pyramid_kernel<PYR_SPEC> with p.natural = 1 - on the send frames only.
__global__ void __launch_bounds__(256) r2c_split_kernel(const PyrParams p) {
    const unsigned R = 1u << p.log2R;
    const int frame = p.frame0 + (int)blockIdx.y * p.frame_step;
    const unsigned d0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    float2 *spec = p.spec + (size_t)frame * p.spec_stride;
    const float2 *Z = p.Z + (size_t)frame * R;
    float2 a[4], b[4];
    {
        const float4 *src = reinterpret_cast<const float4 *>(Z + d0);
        const float4 x0 = src[0], x1 = src[1];
        a[0] = make_float2(x0.x, x0.y);
        a[1] = make_float2(x0.z, x0.w);
        a[2] = make_float2(x1.x, x1.y);
        a[3] = make_float2(x1.z, x1.w);
    }
    if (d0 == 0) {  // the run that pairs with bins 0 .. 3 wraps: Z[0], Z[R-1], Z[R-2], Z[R-3]
        b[0] = Z[0];
        b[1] = Z[R - 1];
        b[2] = Z[R - 2];
        b[3] = Z[R - 3];
    } else {
        const float2 *m = Z + (R - d0 - 3);  // ascending m[0 .. 3] = Z[R-d0-3 .. R-d0], b[i] = m[3 - i]; m + 1 is 16-byte aligned
        b[3] = m[0];
        const float4 x = *reinterpret_cast<const float4 *>(m + 1);
        b[2] = make_float2(x.x, x.y);
        b[1] = make_float2(x.z, x.w);
        b[0] = m[3];
    }
    const unsigned base16 = d0 & ~15u, i0 = d0 & 15u;
    const float2 w0 = cmul(__ldg(p.TLr + (base16 & 1023)), __ldg(p.THr + (base16 >> 10)));
    float2 xs[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 e = make_float2(a[i].x + b[i].x, a[i].y - b[i].y);
        const float2 o = make_float2(a[i].x - b[i].x, a[i].y + b[i].y);
        const float2 w = (i0 + i == 0) ? w0 : cmul(w0, __ldg(p.TLr + i0 + i));
        const float2 t = cmul(o, w);
        float2 x = make_float2(0.5f * (e.x + t.y), 0.5f * (e.y - t.x));
        x.x *= p.scale;
        x.y *= p.scale;
        xs[i] = x;
    }
    if (d0 == 0) {  // Nyquist bin, left unnormalised by the reference (src/fft_impl.cpp:152-154)
        const float2 ny = make_float2(a[0].x - a[0].y, 0.f);
        spec[R] = ny;
        for (int pe = 0; pe < p.npeers; pe++)
            if ((R >= p.peer_lo[pe][0] && R < p.peer_hi[pe][0]) || (R >= p.peer_lo[pe][1] && R < p.peer_hi[pe][1]))
                (p.peers[pe] + (size_t)frame * p.spec_stride)[R] = ny;
    }
    if ((reinterpret_cast<uintptr_t>(spec + d0) & 15) == 0) {
        float4 *dst = reinterpret_cast<float4 *>(spec + d0);
        dst[0] = make_float4(xs[0].x, xs[0].y, xs[1].x, xs[1].y);
        dst[1] = make_float4(xs[2].x, xs[2].y, xs[3].x, xs[3].y);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) spec[d0 + i] = xs[i];
    }
    for (int pe = 0; pe < p.npeers; pe++) {
        float2 *ps = p.peers[pe] + (size_t)frame * p.spec_stride;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned k = d0 + i;
            if ((k >= p.peer_lo[pe][0] && k < p.peer_hi[pe][0]) || (k >= p.peer_lo[pe][1] && k < p.peer_hi[pe][1])) ps[k] = xs[i];
        }
    }
}

// (the Hermitian-split mode is bound by the latency of its mirrored loads: three CTAs per SM instead of the two its 82 registers allow)
template <int MODE, int PER, bool PK, bool TB = false>
__global__ void __launch_bounds__(256, MODE == PYR_R2C ? 3 : 1) pyramid_kernel(const PyrParams p) {
    __shared__ float warp_sum_s[8];
    pyramid_block<MODE, PER, PK, false, TB>(p, p.frame0 + (int)blockIdx.y * p.frame_step, blockIdx.x, threadIdx.x, warp_sum_s, CtaSync{});
}

// more than ten levels above the base (only for very deep pyramids): one block per frame, pairwise tree over
// the sums left in ptop. Tiny.
__global__ void pyramid_tail_kernel(const PyrParams p, int base_level, int levels_done) {
    // levels_done = relative levels already produced by pyramid_kernel (log2(PER) + 9); ptop holds the sums of the last one
    const int frame = p.frame0 + (int)blockIdx.x * p.frame_step;
    const size_t R = (size_t)1 << p.log2R;
    const int first = base_level + levels_done;
    size_t n = R >> (first - 1);
    float *buf = p.ptop + (size_t)frame * n;
    int8_t *quant = p.quant + (size_t)frame * p.pyr_stride;
    size_t lvl_off = 0;
    for (int lv = 0; lv < first; lv++) lvl_off += R >> lv;
    for (int lv = first; lv < p.levels; lv++) {
        n >>= 1;
        // in-place pairwise sum: element j <- buf[2j] + buf[2j+1]; read everything of a chunk before writing it
        for (size_t base = 0; base < n; base += blockDim.x) {
            size_t j = base + threadIdx.x;
            float v = 0.f;
            if (j < n) v = __fadd_rn(buf[2 * j], buf[2 * j + 1]);
            __syncthreads();
            if (j < n) {
                buf[j] = v;
                quant[lvl_off + j] = (int8_t)quantize_dev(v, p.size_log2 - lv);
            }
            __syncthreads();
        }
        lvl_off += R >> lv;
    }
}

// ------------------------------------------------------------------------------------------------
// Stream-ordered flags in (possibly peer) device memory: how the ingest rank tells the other ranks that a bank of
// spectrum frames has landed in their memory, and how they hand the bank back, without the host in the loop.
// ------------------------------------------------------------------------------------------------
struct FlagList {
    unsigned long long *ptr[kMaxPeers];
    int n;
};
__global__ void flag_signal_kernel(FlagList fl, unsigned long long value) {
    if (threadIdx.x < fl.n) {
        __threadfence_system();  // everything earlier kernels of this stream wrote is visible before the flag
        *reinterpret_cast<volatile unsigned long long *>(fl.ptr[threadIdx.x]) = value;
        __threadfence_system();
    }
}
// spins until every flag >= min_value; gives up after timeout_ns of wall time (%globaltimer; sets *err) so a lost peer
// cannot hang the GPU
__global__ void flag_wait_kernel(FlagList fl, unsigned long long min_value, unsigned long long timeout_ns, int *err) {
    if (threadIdx.x < fl.n) {
        auto now_ns = [] {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            return t;
        };
        const unsigned long long t0 = now_ns();
        volatile unsigned long long *f = reinterpret_cast<volatile unsigned long long *>(fl.ptr[threadIdx.x]);
        while (*f < min_value) {
            if (now_ns() - t0 > timeout_ns) {
                *err = 1;
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

}  // namespace b200
