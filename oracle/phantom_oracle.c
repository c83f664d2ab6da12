/*
 * phantom_oracle.c - CPU ORACLE for the PhantomSDR hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE. It is a plain-C restatement of the reference's
 * FFTW path (window -> forward FFT -> /N -> |X|^2 -> approx-log2 -> int8 pyramid,
 * and the per-client slice -> IFFT -> overlap-add -> demod -> DC -> AGC -> int16
 * chain). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker or the timed CPU
 * baseline. The product (phantomsdr_b200/csrc) never links or calls it.
 *
 * PARITY PINNING STATUS (see DESIGN.md "Oracle"):
 *   - The reference ships NO tests, fixtures or golden vectors for this path
 *      and its FFT provider (FFTW3f, system library,
 *     meson.build:32-38, wrap pin fftw-3.3.10) is not under /root/reference and not
 *     in this image: the forward FFT and the three small inverse FFTs are
 *     "PARITY UNPINNED" against FFTW itself; they restate FFTW's published
 *     definition (unnormalised DFT, forward sign -1, backward +1) and are checked
 *     against numpy/pocketfft in float64.
 *   - Everything that compiles from the reference's own sources IS pinned:
 *     oracle/Makefile builds oracle/_ref/libphantom_ref.so from
 *     /root/reference/src/utils/dsp.cpp, src/utils/audioprocessing.cpp and (through a
 *     stub boost::circular_buffer) src/utils.h; tests/test_oracle_vs_ref.py checks
 *     the restatements below bit-for-bit against that library (Hann window, FM
 *     discriminator, AM envelope, negate/add, float->int16, AGC, DC blocker).
 *
 * Floating-point convention: compiled with -ffp-contract=off (no FMA contraction),
 * every operation is a separately rounded IEEE binary32 op in the order the
 * reference source writes it. The reference itself is built with -march=native
 * (meson.build:14) so its own low-order bits depend on the build host; the
 * un-contracted order is the canonical definition used here and by the CUDA path.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define REAL float
#define SUFFIX _f32
#include "fft_generic.inc"
#undef REAL
#undef SUFFIX
#define REAL double
#define SUFFIX _f64
#include "fft_generic.inc"
#undef REAL
#undef SUFFIX

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Derived sizes                                                             */
/* ------------------------------------------------------------------------- */

/* src/spectrumserver.cpp:151  audio_max_fft_size = ceil((double)audio_max_sps * fft_size / sps / 4.) * 4 */
ORC_API int orc_audio_fft_size(int audio_max_sps, int fft_size, int sps) {
    return (int)(ceil((double)audio_max_sps * fft_size / sps / 4.) * 4);
}

/* src/spectrumserver.cpp:185-190 */
ORC_API int orc_downsample_levels(int fft_result_size, int min_waterfall_fft) {
    int levels = 0;
    for (int cur = fft_result_size; cur >= min_waterfall_fft; cur /= 2) levels++;
    return levels;
}

/* src/fft.cpp:33  skip_num = max(1, (int)floor(((float)sps / fft_size) / 10.) * 2) */
ORC_API int orc_skip_num(int sps, int fft_size) {
    int v = (int)floor(((float)sps / fft_size) / 10.) * 2;
    return v > 1 ? v : 1;
}

/* ------------------------------------------------------------------------- */
/* a1: Hann window - src/utils/dsp.cpp:6-11                                  */
/*   arr[i] = 0.5 * (1 - cosf(2 * M_PI * i / num));                          */
/* 2*M_PI*i/num is formed in double and narrowed to float for cosf; 1 - cosf */
/* is a float subtraction; 0.5 * (.) is a double product narrowed on store.  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_hann_window(float *arr, int num) {
    for (int i = 0; i < num; i++) {
        float c = cosf((float)(2 * M_PI * i / num));
        float one_minus = 1 - c;
        arr[i] = (float)(0.5 * (double)one_minus);
    }
}

/* ------------------------------------------------------------------------- */
/* L1 ingest - src/samplereader.cpp:29-40,59-66               */
/* unsigned types: XOR the top bit, reinterpret as signed; then / 2^(bits-1) */
/* ------------------------------------------------------------------------- */
ORC_API void orc_convert_u8(const uint8_t *in, float *out, size_t num) {
    for (size_t i = 0; i < num; i++) out[i] = ((float)(int8_t)(in[i] ^ 0x80)) / 128.0f;
}
ORC_API void orc_convert_s8(const int8_t *in, float *out, size_t num) {
    for (size_t i = 0; i < num; i++) out[i] = ((float)in[i]) / 128.0f;
}
ORC_API void orc_convert_u16(const uint16_t *in, float *out, size_t num) {
    for (size_t i = 0; i < num; i++) out[i] = ((float)(int16_t)(in[i] ^ 0x8000)) / 32768.0f;
}
ORC_API void orc_convert_s16(const int16_t *in, float *out, size_t num) {
    for (size_t i = 0; i < num; i++) out[i] = ((float)in[i]) / 32768.0f;
}

/* ------------------------------------------------------------------------- */
/* a6: vec_log2 - src/fft_impl.cpp:14-23                                     */
/* ------------------------------------------------------------------------- */
static inline float orc_vec_log2(float val, int power_offset) {
    uint32_t bits;
    memcpy(&bits, &val, 4);
    float log_val = (float)((int)((bits >> 23) & 0xFF) - 128) + power_offset;
    bits &= ~(255u << 23);
    bits += 127u << 23;
    memcpy(&val, &bits, 4);
    log_val += ((-0.34484843f) * val + 2.02466578f) * val - 0.67487759f;
    return log_val;
}

/* The reference stores std::max(-128.f, x) into an int8_t (fft_impl.cpp:40-42,57-59):
   C float->int8 conversion truncates toward zero; values above 127 are undefined
   behaviour there and wrap modulo 256 in practice on x86 (cvttss2si + byte store).
   The oracle DEFINES that wrap: truncate to int32, keep the low 8 bits. */
static inline int8_t orc_quantize(float power, int power_offset) {
    float v = orc_vec_log2(power, power_offset) * 0.3010299956639812f * 20.f + 127.f;
    v = v > -128.f ? v : -128.f; /* std::max(-128.f, v): returns -128 when v is NaN-free smaller */
    if (!(v == v)) v = -128.f;   /* std::max(a,b) returns a when b is NaN */
    if (v > 2147483520.f) v = 2147483520.f;
    int32_t t = (int32_t)v;
    return (int8_t)(uint8_t)(t & 0xFF);
}

ORC_API int8_t orc_quantize_one(float power, int power_offset) { return orc_quantize(power, power_offset); }

/* a5: power_and_quantize - src/fft_impl.cpp:24-44 */
static void orc_power_and_quantize(float *complexbuf, float *powerbuf, int8_t *quantizedbuf,
                                   float normalize, size_t len, int power_offset) {
#pragma omp parallel for
    for (size_t i = 0; i < len; i++) {
        complexbuf[i * 2] /= normalize;
        complexbuf[i * 2 + 1] /= normalize;
        float re = complexbuf[i * 2];
        float im = complexbuf[i * 2 + 1];
        float power = re * re + im * im;
        powerbuf[i] = power;
        quantizedbuf[i] = orc_quantize(power, power_offset);
    }
}

/* a7: half_and_quantize - src/fft_impl.cpp:45-61 */
static void orc_half_and_quantize(const float *powerbuf, float *halfbuf, int8_t *quantizedbuf,
                                  size_t len, int power_offset) {
#pragma omp parallel for
    for (size_t i = 0; i < len; i++) {
        float power = powerbuf[i * 2] + powerbuf[i * 2 + 1];
        halfbuf[i] = power;
        quantizedbuf[i] = orc_quantize(power, power_offset);
    }
}

/* ------------------------------------------------------------------------- */
/* The FFT backend object - class FFT / class FFTW, src/fft.h:33-105,        */
/* src/fft_impl.cpp:63-183                                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    size_t size;
    int size_log2;
    int downsample_levels;
    int additional_size;
    size_t outbuf_len; /* size (c2c) or size/2 (r2c) */
    int is_real;
    float *windowbuf;
    float *inbuf;
    float *outbuf;
    float *powerbuf;
    int8_t *quantizedbuf;
    orc_plan_f32 *plan;
    float *cscratch; /* r2c: complexified input / full complex output */
} orc_fft;

/* FFT::FFT - src/fft_impl.cpp:63-70 */
ORC_API orc_fft *orc_fft_create(size_t size, int downsample_levels, int brightness_offset) {
    orc_fft *f = (orc_fft *)calloc(1, sizeof(*f));
    f->size = size;
    f->downsample_levels = downsample_levels;
    f->windowbuf = (float *)aligned_alloc(64, sizeof(float) * size);
    f->size_log2 = (int)round(log2((double)size)) + brightness_offset;
    orc_hann_window(f->windowbuf, (int)size);
    return f;
}
/* FFT::set_output_additional_size - src/fft_impl.cpp:74 */
ORC_API void orc_fft_set_output_additional_size(orc_fft *f, size_t n) { f->additional_size = (int)n; }

/* FFTW::plan_c2c - src/fft_impl.cpp:89-103 (direction is always FORWARD at the call site fft.cpp:28) */
ORC_API int orc_fft_plan_c2c(orc_fft *f) {
    size_t size = f->size;
    f->inbuf = (float *)aligned_alloc(64, sizeof(float) * size * 2);
    f->outbuf = (float *)aligned_alloc(64, sizeof(float) * (size * 2 + (size_t)f->additional_size * 2));
    memset(f->outbuf, 0, sizeof(float) * (size * 2 + (size_t)f->additional_size * 2));
    f->outbuf_len = size;
    f->is_real = 0;
    f->powerbuf = (float *)aligned_alloc(64, sizeof(float) * size * 2);
    f->quantizedbuf = (int8_t *)aligned_alloc(64, size * 2);
    f->plan = orc_plan_create_f32((long)size, -1);
    return 0;
}
/* FFTW::plan_r2c - src/fft_impl.cpp:104-117 */
ORC_API int orc_fft_plan_r2c(orc_fft *f) {
    size_t size = f->size;
    f->inbuf = (float *)aligned_alloc(64, sizeof(float) * size);
    f->outbuf = (float *)aligned_alloc(64, sizeof(float) * (size + 2));
    memset(f->outbuf, 0, sizeof(float) * (size + 2));
    f->outbuf_len = size / 2;
    f->is_real = 1;
    f->powerbuf = (float *)aligned_alloc(64, sizeof(float) * size);
    f->quantizedbuf = (int8_t *)aligned_alloc(64, size);
    f->plan = orc_plan_create_f32((long)size, -1);
    f->cscratch = (float *)aligned_alloc(64, sizeof(float) * size * 4);
    return 0;
}
ORC_API void orc_fft_destroy(orc_fft *f) {
    if (!f) return;
    orc_plan_destroy_f32(f->plan);
    free(f->windowbuf);
    free(f->inbuf);
    free(f->outbuf);
    free(f->powerbuf);
    free(f->quantizedbuf);
    free(f->cscratch);
    free(f);
}
ORC_API float *orc_fft_window(orc_fft *f) { return f->windowbuf; }
ORC_API float *orc_fft_input(orc_fft *f) { return f->inbuf; }
ORC_API float *orc_fft_output(orc_fft *f) { return f->outbuf; }
ORC_API float *orc_fft_power(orc_fft *f) { return f->powerbuf; }
ORC_API int8_t *orc_fft_quantized(orc_fft *f) { return f->quantizedbuf; }
ORC_API int orc_fft_size_log2(orc_fft *f) { return f->size_log2; }

/* a3: FFTW::load_real_input - src/fft_impl.cpp:131-135 */
ORC_API int orc_fft_load_real_input(orc_fft *f, const float *a1, const float *a2) {
    size_t h = f->size / 2;
    for (size_t i = 0; i < h; i++) f->inbuf[i] = a1[i] * f->windowbuf[i];
    for (size_t i = 0; i < h; i++) f->inbuf[h + i] = a2[i] * f->windowbuf[h + i];
    return 0;
}
/* a3: FFTW::load_complex_input - src/fft_impl.cpp:136-143 (complex * real scalar: each part scaled) */
ORC_API int orc_fft_load_complex_input(orc_fft *f, const float *a1, const float *a2) {
    size_t h = f->size / 2;
    for (size_t i = 0; i < h; i++) {
        f->inbuf[2 * i] = a1[2 * i] * f->windowbuf[i];
        f->inbuf[2 * i + 1] = a1[2 * i + 1] * f->windowbuf[i];
    }
    for (size_t i = 0; i < h; i++) {
        f->inbuf[2 * (h + i)] = a2[2 * i] * f->windowbuf[h + i];
        f->inbuf[2 * (h + i) + 1] = a2[2 * i + 1] * f->windowbuf[h + i];
    }
    return 0;
}

/* a4: fftwf_execute(p) - src/fft_impl.cpp:145. Unnormalised forward DFT (sign -1).
   r2c returns bins 0..N/2 inclusive. */
ORC_API int orc_fft_transform(orc_fft *f) {
    if (!f->is_real) {
        orc_fft_exec_f32(f->plan, f->inbuf, f->outbuf);
    } else {
        size_t n = f->size;
        float *cin = f->cscratch, *cout = f->cscratch + 2 * n;
        for (size_t i = 0; i < n; i++) {
            cin[2 * i] = f->inbuf[i];
            cin[2 * i + 1] = 0.f;
        }
        orc_fft_exec_f32(f->plan, cin, cout);
        memcpy(f->outbuf, cout, sizeof(float) * (n + 2));
    }
    return 0;
}

/* a5+a7: the rest of FFTW::execute - src/fft_impl.cpp:146-173.
   outbuf must hold the raw (unnormalised) transform on entry. */
ORC_API int orc_fft_quantize(orc_fft *f) {
    int base_idx = 0;
    size_t size = f->size, outbuf_len = f->outbuf_len;
    if (!f->is_real) base_idx = (int)(size / 2 + 1);
    orc_power_and_quantize(&f->outbuf[(size_t)base_idx * 2], f->powerbuf, f->quantizedbuf, (float)size,
                           outbuf_len - base_idx, f->size_log2);
    orc_power_and_quantize(f->outbuf, &f->powerbuf[outbuf_len - base_idx],
                           &f->quantizedbuf[outbuf_len - base_idx], (float)size, base_idx, f->size_log2);
    size_t out_len = outbuf_len;
    int8_t *q = f->quantizedbuf;
    float *p = f->powerbuf;
    for (int i = 0; i < f->downsample_levels - 1; i++) {
        orc_half_and_quantize(p, p + out_len, q + out_len, out_len / 2, f->size_log2 - i - 1);
        p += out_len;
        q += out_len;
        out_len /= 2;
    }
    return 0;
}

/* FFTW::execute - src/fft_impl.cpp:144-174 */
ORC_API int orc_fft_execute(orc_fft *f) {
    orc_fft_transform(f);
    return orc_fft_quantize(f);
}

/* a8: IQ wrap copy done by the caller - src/fft.cpp:91-98 */
ORC_API void orc_fft_wrap_copy(orc_fft *f, size_t audio_max_fft_size) {
    if (!f->is_real) memcpy(&f->outbuf[2 * f->size], &f->outbuf[0], sizeof(float) * 2 * audio_max_fft_size);
}

/* float64 shadow of the forward transform of the current inbuf, normalised by 1/N. out: 2*size doubles
   (c2c) or size+2 doubles (r2c). Used only to arbitrate between two float32 implementations. */
ORC_API void orc_fft_shadow_f64(orc_fft *f, double *out) {
    size_t n = f->size;
    orc_plan_f64 *pl = orc_plan_create_f64((long)n, -1);
    double *cin = (double *)malloc(sizeof(double) * 2 * n);
    double *cout = (double *)malloc(sizeof(double) * 2 * n);
    if (!f->is_real) {
        for (size_t i = 0; i < 2 * n; i++) cin[i] = f->inbuf[i];
    } else {
        for (size_t i = 0; i < n; i++) {
            cin[2 * i] = f->inbuf[i];
            cin[2 * i + 1] = 0;
        }
    }
    orc_fft_exec_f64(pl, cin, cout);
    size_t nout = f->is_real ? n + 2 : 2 * n;
    for (size_t i = 0; i < nout; i++) out[i] = cout[i] / (double)n;
    free(cin);
    free(cout);
    orc_plan_destroy_f64(pl);
}

/* Stand-alone DFTs for tests (any length). */
ORC_API void orc_dft_f32(const float *in, float *out, long n, int sign) {
    orc_plan_f32 *pl = orc_plan_create_f32(n, sign);
    orc_fft_exec_f32(pl, in, out);
    orc_plan_destroy_f32(pl);
}
ORC_API void orc_dft_f64(const double *in, double *out, long n, int sign) {
    orc_plan_f64 *pl = orc_plan_create_f64(n, sign);
    orc_fft_exec_f64(pl, in, out);
    orc_plan_destroy_f64(pl);
}

/* ------------------------------------------------------------------------- */
/* a15: MovingAverage / DCBlocker - src/utils.h:76-99,139-169                */
/* boost::circular_buffer<T> q{length, 0}: FULL buffer of `length` zeros;    */
/* push_front on a full buffer overwrites the back. q[0] is the newest.      */
/* `sum` is a Neumaier<float> whose conversion returns `sum` only            */
/* (utils.h:24), so the correction term never reaches the output.            */
/* ------------------------------------------------------------------------- */
typedef struct {
    int length;
    int head; /* index of q[0] (newest) in buf */
    float *buf;
    float sum;
} orc_ma;

static void orc_ma_init(orc_ma *m, int length) {
    m->length = length;
    m->head = 0;
    m->buf = (float *)calloc((size_t)length, sizeof(float));
    m->sum = 0.f;
}
static inline float orc_ma_at(const orc_ma *m, int i) { return m->buf[(m->head + i) % m->length]; }
/* MovingAverage::insert - src/utils.h:80-85 */
static inline float orc_ma_insert(orc_ma *m, float val) {
    float back = orc_ma_at(m, m->length - 1);
    m->sum = m->sum + (-back);                        /* sum -= q.back()  (Neumaier += -value: sum = sum + value) */
    m->head = (m->head + m->length - 1) % m->length;  /* q.push_front(val) */
    m->buf[m->head] = val;
    m->sum = m->sum + val;                            /* sum += val */
    return m->sum / m->length;                        /* getAverage(): float / int */
}

typedef struct {
    int delay;
    orc_ma ma1, ma2;
} orc_dc;

static void orc_dc_init(orc_dc *d, int delay) {
    d->delay = delay;
    orc_ma_init(&d->ma1, delay);
    orc_ma_init(&d->ma2, delay);
}
/* DCBlocker::processSample - src/utils.h:145-149 */
static inline float orc_dc_sample(orc_dc *d, float s) {
    float ma1 = orc_ma_insert(&d->ma1, s);
    float ma2 = orc_ma_insert(&d->ma2, ma1);
    return orc_ma_at(&d->ma1, d->delay - 1) - ma2;
}

ORC_API orc_dc *orc_dc_create(int delay) {
    orc_dc *d = (orc_dc *)calloc(1, sizeof(*d));
    orc_dc_init(d, delay);
    return d;
}
ORC_API void orc_dc_destroy(orc_dc *d) {
    if (!d) return;
    free(d->ma1.buf);
    free(d->ma2.buf);
    free(d);
}
/* DCBlocker::removeDC - src/utils.h:150-157 */
ORC_API void orc_dc_remove(orc_dc *d, float *arr, int length) {
    for (int i = 0; i < length; i++) arr[i] = orc_dc_sample(d, arr[i]);
}

/* ------------------------------------------------------------------------- */
/* a16: AGC - src/utils/audioprocessing.cpp:5-74                             */
/* ------------------------------------------------------------------------- */
typedef struct {
    float desired_level, attack_coeff, release_coeff, gain, sample_rate;
    size_t look_ahead_samples;
    /* std::deque<float> lookahead_buffer / lookahead_max as rings of capacity look_ahead+2 */
    float *buf;
    size_t buf_head, buf_size;
    float *mx;
    size_t mx_head, mx_size;
    size_t cap;
} orc_agc;

/* AGC::AGC - audioprocessing.cpp:5-15. `exp` is the unqualified C double exp (only <cmath> is
   included), so the coefficient is formed in double and narrowed on store. */
ORC_API orc_agc *orc_agc_create(float desiredLevel, float attackTimeMs, float releaseTimeMs,
                                float lookAheadTimeMs, float sr) {
    orc_agc *a = (orc_agc *)calloc(1, sizeof(*a));
    a->desired_level = desiredLevel;
    a->gain = 0;
    a->sample_rate = sr;
    a->look_ahead_samples = (size_t)(lookAheadTimeMs * a->sample_rate / 1000.0f);
    a->attack_coeff = (float)(1 - exp((double)(-1.0f / (attackTimeMs * 0.001f * a->sample_rate))));
    a->release_coeff = (float)(1 - exp((double)(-1.0f / (releaseTimeMs * 0.001f * a->sample_rate))));
    a->cap = a->look_ahead_samples + 2;
    a->buf = (float *)calloc(a->cap, sizeof(float));
    a->mx = (float *)calloc(a->cap, sizeof(float));
    return a;
}
ORC_API void orc_agc_destroy(orc_agc *a) {
    if (!a) return;
    free(a->buf);
    free(a->mx);
    free(a);
}
ORC_API float orc_agc_attack(orc_agc *a) { return a->attack_coeff; }
ORC_API float orc_agc_release(orc_agc *a) { return a->release_coeff; }
ORC_API size_t orc_agc_lookahead(orc_agc *a) { return a->look_ahead_samples; }
ORC_API float orc_agc_gain(orc_agc *a) { return a->gain; }

/* AGC::pop - audioprocessing.cpp:30-36 */
static inline void orc_agc_pop(orc_agc *a) {
    float sample = a->buf[a->buf_head];
    a->buf_head = (a->buf_head + 1) % a->cap;
    a->buf_size--;
    if (sample == a->mx[a->mx_head]) {
        a->mx_head = (a->mx_head + 1) % a->cap;
        a->mx_size--;
    }
}
/* AGC::push - audioprocessing.cpp:17-28 */
static inline void orc_agc_push(orc_agc *a, float sample) {
    a->buf[(a->buf_head + a->buf_size) % a->cap] = sample;
    a->buf_size++;
    while (a->mx_size && fabsf(a->mx[(a->mx_head + a->mx_size - 1) % a->cap]) < fabsf(sample)) a->mx_size--;
    a->mx[(a->mx_head + a->mx_size) % a->cap] = sample;
    a->mx_size++;
    if (a->buf_size > a->look_ahead_samples) orc_agc_pop(a);
}
/* AGC::process - audioprocessing.cpp:40-68 */
ORC_API void orc_agc_process(orc_agc *a, float *arr, size_t len) {
    for (size_t i = 0; i < len; i++) {
        orc_agc_push(a, arr[i]);
        if (a->buf_size == a->look_ahead_samples) {
            float current_sample = a->buf[a->buf_head];
            float peak_sample = fabsf(a->mx[a->mx_head]); /* AGC::max() */
            float desired_gain = a->desired_level / (peak_sample + 1e-10f);
            if (desired_gain < a->gain) {
                a->gain = a->gain - a->attack_coeff * (a->gain - desired_gain);
            } else {
                a->gain = a->gain + a->release_coeff * (desired_gain - a->gain);
            }
            arr[i] = current_sample * a->gain;
        } else {
            arr[i] = 0.0f;
        }
    }
}
/* AGC::reset - audioprocessing.cpp:70-74 */
ORC_API void orc_agc_reset(orc_agc *a) {
    a->gain = 0;
    a->buf_head = a->buf_size = 0;
    a->mx_head = a->mx_size = 0;
}

/* ------------------------------------------------------------------------- */
/* a12/a13/a17 helpers - src/utils/dsp.cpp                                   */
/* ------------------------------------------------------------------------- */
/* polar_discriminator_fm - dsp.cpp:27-35: arg(buf[i] * conj(prev)).
   (a+bi)(c-di): libstdc++ complex<float> operator* without -ffast-math evaluates
   re = a*c - b*(-d), im = a*(-d) + b*c and only falls back to __mulsc3 on NaN. */
ORC_API void orc_polar_discriminator_fm(const float *buf, float prev_re, float prev_im, float *output,
                                        size_t len) {
    for (size_t i = 0; i < len; i++) {
        float a = buf[2 * i], b = buf[2 * i + 1];
        float c = prev_re, d = -prev_im;
        float re = a * c - b * d;
        float im = a * d + b * c;
        output[i] = atan2f(im, re);
        prev_re = a;
        prev_im = b;
    }
}
/* dsp_am_demod - dsp.cpp:116-126 */
ORC_API void orc_am_demod(const float *arr, float *output, size_t len) {
    for (size_t i = 0; i < len; i++) {
        float re = arr[2 * i], im = arr[2 * i + 1];
        output[i] = sqrtf(re * re + im * im);
    }
}
/* dsp_float_to_int16 - dsp.cpp:152-165 */
ORC_API void orc_float_to_int16(const float *arr, int32_t *output, float mult, size_t len) {
    const int32_t minimum = -32768, maximum = 32767;
    for (size_t i = 0; i < len; i++) {
        float t = arr[i] * mult + 32768.5f;
        /* (int32_t) of an out-of-range float is UB in the reference; define it as saturation */
        int32_t v;
        if (t >= 2147483648.f) v = 2147483647;
        else if (t <= -2147483648.f) v = (int32_t)0x80000000;
        else v = (int32_t)t;
        int64_t o = (int64_t)v - 32768;
        if (o > maximum) o = maximum;
        if (o < minimum) o = minimum;
        output[i] = (int32_t)o;
    }
}

/* ------------------------------------------------------------------------- */
/* a10-a17: AudioClient - src/signal.cpp:7-79 (ctor), :81-98, :102-298        */
/* ------------------------------------------------------------------------- */
enum { ORC_USB = 0, ORC_LSB = 1, ORC_AM = 2, ORC_FM = 3 }; /* src/client.h:43 */

typedef struct {
    int is_real, audio_fft_size, fft_result_size, audio_rate;
    int l, r, demodulation;
    double audio_mid;
    float *audio_fft_input;         /* n complex */
    float *audio_complex_baseband;  /* n complex */
    float *audio_complex_baseband_prev;
    float *audio_real, *audio_real_prev; /* n floats each */
    int32_t *audio_real_int16;
    orc_dc dc;
    orc_agc *agc;
    orc_plan_f32 *p_complex; /* backward, n */
    float *herm;             /* scratch for the c2r emulation */
} orc_client;

/* AudioClient::AudioClient - src/signal.cpp:7-79: zeroed scratch, DCBlocker(audio_max_sps/750*2),
   AGC(0.2f, 50.0f, 300.0f, 200.0f, audio_max_sps) */
ORC_API orc_client *orc_client_create(int is_real, int audio_fft_size, int audio_max_sps, int fft_result_size) {
    orc_client *c = (orc_client *)calloc(1, sizeof(*c));
    size_t n = (size_t)audio_fft_size;
    c->is_real = is_real;
    c->audio_fft_size = audio_fft_size;
    c->fft_result_size = fft_result_size;
    c->audio_rate = audio_max_sps;
    c->audio_fft_input = (float *)calloc(2 * n, sizeof(float));
    c->audio_complex_baseband = (float *)calloc(2 * n, sizeof(float));
    c->audio_complex_baseband_prev = (float *)calloc(2 * n, sizeof(float));
    c->audio_real = (float *)calloc(n, sizeof(float));
    c->audio_real_prev = (float *)calloc(n, sizeof(float));
    c->audio_real_int16 = (int32_t *)calloc(n, sizeof(int32_t));
    c->herm = (float *)calloc(4 * n, sizeof(float));
    orc_dc_init(&c->dc, audio_max_sps / 750 * 2);
    c->agc = orc_agc_create(0.2f, 50.0f, 300.0f, 200.0f, (float)audio_max_sps);
    c->p_complex = orc_plan_create_f32((long)n, +1);
    c->demodulation = ORC_USB;
    return c;
}
ORC_API void orc_client_destroy(orc_client *c) {
    if (!c) return;
    free(c->audio_fft_input);
    free(c->audio_complex_baseband);
    free(c->audio_complex_baseband_prev);
    free(c->audio_real);
    free(c->audio_real_prev);
    free(c->audio_real_int16);
    free(c->herm);
    free(c->dc.ma1.buf);
    free(c->dc.ma2.buf);
    orc_agc_destroy(c->agc);
    orc_plan_destroy_f32(c->p_complex);
    free(c);
}
/* AudioClient::set_audio_range - src/signal.cpp:81-94 (the multimap re-key lives in the caller) */
ORC_API void orc_client_set_audio_range(orc_client *c, int l, double m, int r) {
    c->audio_mid = m;
    c->l = l;
    c->r = r;
}
/* AudioClient::on_window_message validation - src/signal.cpp:300-314. Returns 1 if accepted. */
ORC_API int orc_client_on_window_message(orc_client *c, int new_l, double new_m, int new_r) {
    if (new_l < 0 || new_l >= c->fft_result_size || new_r < 0 || new_r >= c->fft_result_size || new_l > new_r)
        return 0;
    if (new_r - new_l > c->audio_fft_size) return 0;
    orc_client_set_audio_range(c, new_l, new_m, new_r);
    return 1;
}
/* AudioClient::set_audio_demodulation - src/signal.cpp:95-97 (no AGC reset) */
ORC_API void orc_client_set_audio_demodulation(orc_client *c, int mode) { c->demodulation = mode; }
/* AudioClient::on_demodulation_message - src/signal.cpp:316-328 (AGC reset) */
ORC_API void orc_client_on_demodulation_message(orc_client *c, int mode) {
    if (mode >= ORC_USB && mode <= ORC_FM) c->demodulation = mode;
    orc_agc_reset(c->agc);
}

/* fftwf_execute(p_real): c2r of length n from in[0..n/2]. FFTW defines the input as the
   non-redundant half of a Hermitian spectrum; the imaginary parts of in[0] and in[n/2]
   are not part of that definition and are taken as zero here (PARITY UNPINNED against
   FFTW's codelets for non-Hermitian-consistent input). */
static void orc_c2r(orc_client *c, const float *in, float *out) {
    int n = c->audio_fft_size;
    float *full = c->herm, *res = c->herm + 2 * n;
    full[0] = in[0];
    full[1] = 0.f;
    for (int k = 1; k < n / 2; k++) {
        full[2 * k] = in[2 * k];
        full[2 * k + 1] = in[2 * k + 1];
        full[2 * (n - k)] = in[2 * k];
        full[2 * (n - k) + 1] = -in[2 * k + 1];
    }
    full[2 * (n / 2)] = in[2 * (n / 2)];
    full[2 * (n / 2) + 1] = 0.f;
    orc_fft_exec_f32(c->p_complex, full, res);
    for (int t = 0; t < n; t++) out[t] = res[2 * t];
}

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* AudioClient::send_audio - src/signal.cpp:102-298.
 *   buf: &fft_buffer[(l + base_idx) % fft_result_size] (complex, interleaved) - formed by the
 *        caller exactly as websocket.cpp:182 does (orc_signal_slice_offset below).
 *   pcm_out[n/2], *pwr_out: what the reference hands to encoder->process / set_data (signal.cpp:287-291).
 *   audio_pre_dc (nullable, n/2 floats): tap of audio_real before DC removal (test aid).
 * Returns 1 if a packet would be sent, 0 if the frame is dropped by the NaN guard (signal.cpp:266-271).
 */
ORC_API int orc_client_send_audio(orc_client *c, const float *buf, size_t frame_num, int32_t *pcm_out,
                                  float *pwr_out, float *audio_pre_dc) {
    const int n = c->audio_fft_size;
    const int audio_l = 0;
    const int audio_r = c->r - c->l;
    const int audio_m = (int)floor(c->audio_mid) - c->l;
    const int audio_m_idx = (int)floor(c->audio_mid);
    const int is_real = c->is_real;
    int len = audio_r - audio_l;

    /* signal.cpp:117-119 - std::accumulate of std::norm. libstdc++'s norm() for floating types is
       re*re + im*im (bits/complex: _Norm_helper, the abs()^2 form is commented out there); pinned bit for
       bit against the reference's compiled send_audio by tests/test_oracle_vs_ref_fft.py */
    float average_power = 0.0f;
    for (int i = 0; i < len; i++) {
        float re = buf[2 * i], im = buf[2 * i + 1];
        average_power = average_power + (re * re + im * im);
    }

    float *in = c->audio_fft_input;
    float *bb = c->audio_complex_baseband;
    float *bbp = c->audio_complex_baseband_prev;
    float *ar = c->audio_real;
    int negate = (frame_num % 2 == 1) &&
                 ((audio_m_idx % 2 == 0 && !is_real) || (audio_m_idx % 2 == 1 && is_real));
    /* NB: audio_m_idx % 2 == 1 is false for negative odd values in C; floor(mid) >= 0 always here */

    if (c->demodulation == ORC_USB || c->demodulation == ORC_LSB) {
        memset(in, 0, sizeof(float) * 2 * n);
        if (c->demodulation == ORC_USB) {
            /* signal.cpp:125-138 */
            int copy_l = imax(audio_l, audio_m);
            int copy_r = imin(audio_r, audio_m + n);
            if (copy_r >= copy_l)
                memcpy(in + 2 * (copy_l - audio_m), buf + 2 * (copy_l - audio_l),
                       sizeof(float) * 2 * (size_t)(copy_r - copy_l));
            orc_c2r(c, in, ar);
        } else {
            /* signal.cpp:139-156 */
            int copy_l = imax(audio_l, audio_m - n + 1);
            int copy_r = imin(audio_r, audio_m + 1);
            if (copy_r >= copy_l) {
                /* reverse_copy(buf+copy_l, buf+copy_r, in + audio_m - copy_r + 1) */
                int cnt = copy_r - copy_l;
                float *dst = in + 2 * (audio_m - copy_r + 1);
                for (int j = 0; j < cnt; j++) {
                    dst[2 * j] = buf[2 * (copy_r - 1 - j - audio_l)];
                    dst[2 * j + 1] = buf[2 * (copy_r - 1 - j - audio_l) + 1];
                }
            }
            orc_c2r(c, in, ar);
            for (int i = 0; i < n / 2; i++) { /* std::reverse(audio_real) */
                float t = ar[i];
                ar[i] = ar[n - 1 - i];
                ar[n - 1 - i] = t;
            }
        }
        /* signal.cpp:160-168 */
        if (negate)
            for (int i = 0; i < n; i++) ar[i] = -ar[i];
        /* signal.cpp:170-172 */
        for (int i = 0; i < n / 2; i++) ar[i] += c->audio_real_prev[i];
    } else {
        /* AM / FM: signal.cpp:173-198 */
        memset(in, 0, sizeof(float) * 2 * n);
        int pos_copy_l = imax(audio_l, audio_m);
        int pos_copy_r = imin(audio_r, audio_m + n / 2);
        if (pos_copy_r >= pos_copy_l)
            memcpy(in + 2 * (pos_copy_l - audio_m), buf + 2 * (pos_copy_l - audio_l),
                   sizeof(float) * 2 * (size_t)(pos_copy_r - pos_copy_l));
        int neg_copy_l = imax(audio_l, audio_m - n / 2 + 1);
        int neg_copy_r = imin(audio_r, audio_m);
        if (neg_copy_r >= neg_copy_l)
            memcpy(in + 2 * (n - (audio_m - neg_copy_l)), buf + 2 * (neg_copy_l - audio_l),
                   sizeof(float) * 2 * (size_t)(neg_copy_r - neg_copy_l));
        /* signal.cpp:200-203 */
        float prev_re = bb[2 * (n / 2 - 1)], prev_im = bb[2 * (n / 2 - 1) + 1];
        memcpy(bbp, bb + 2 * (n / 2), sizeof(float) * 2 * (size_t)(n / 2));
        /* signal.cpp:214  (the carrier path :204-222 only feeds the HAS_LIQUID branch; this
           oracle restates the #else envelope branch) */
        orc_fft_exec_f32(c->p_complex, in, bb);
        /* signal.cpp:223-234 */
        if (negate)
            for (int i = 0; i < 2 * n; i++) bb[i] = -bb[i];
        /* signal.cpp:235-237 */
        for (int i = 0; i < 2 * (n / 2); i++) bb[i] += bbp[i];
        if (c->demodulation == ORC_AM) {
            orc_am_demod(bb, ar, (size_t)(n / 2)); /* signal.cpp:255-256 */
        } else {
            orc_polar_discriminator_fm(bb, prev_re, prev_im, ar, (size_t)(n / 2)); /* signal.cpp:259-263 */
        }
    }

    /* signal.cpp:266-271 */
    for (int i = 0; i < n / 2; i++) {
        if (isnan(ar[i])) return 0;
    }
    /* signal.cpp:274-275 */
    memcpy(c->audio_real_prev, ar + n / 2, sizeof(float) * (size_t)(n / 2));
    if (audio_pre_dc) memcpy(audio_pre_dc, ar, sizeof(float) * (size_t)(n / 2));
    /* signal.cpp:278-284 */
    orc_dc_remove(&c->dc, ar, n / 2);
    orc_agc_process(c->agc, ar, (size_t)(n / 2));
    orc_float_to_int16(ar, c->audio_real_int16, 65536 / 4, (size_t)(n / 2));
    /* signal.cpp:287-291 */
    if (pwr_out) *pwr_out = average_power;
    if (pcm_out) memcpy(pcm_out, c->audio_real_int16, sizeof(int32_t) * (size_t)(n / 2));
    return 1;
}

/* a9: signal_loop slice pointer - src/websocket.cpp:156-185: (l + base_idx) % fft_result_size,
   base_idx = fft_size/2 + 1 for IQ, 0 for real. */
ORC_API size_t orc_signal_slice_offset(int l, size_t fft_size, int is_real) {
    size_t base_idx = is_real ? 0 : fft_size / 2 + 1;
    size_t fft_result_size = is_real ? fft_size / 2 : fft_size;
    return ((size_t)l + base_idx) % fft_result_size;
}
/* waterfall_loop level base - src/websocket.cpp:207-236: level i starts at sum_{j<i} (R >> j) */
ORC_API size_t orc_waterfall_level_offset(int level, size_t fft_result_size) {
    size_t off = 0;
    for (int i = 0; i < level; i++) off += fft_result_size >> i;
    return off;
}

/* Batched convenience for the timed CPU baseline: run send_audio for `nclients` clients over the
   same spectrum, OpenMP over clients (the analogue of the asio pool, spectrumserver.cpp:253-258). */
ORC_API void orc_clients_send_audio(orc_client **clients, int nclients, const float *spectrum, size_t fft_size,
                                    int is_real, size_t frame_num, int32_t *pcm_out, float *pwr_out,
                                    uint8_t *valid_out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < nclients; i++) {
        orc_client *c = clients[i];
        size_t off = orc_signal_slice_offset(c->l, fft_size, is_real);
        int half = c->audio_fft_size / 2;
        int ok = orc_client_send_audio(c, spectrum + 2 * off, frame_num, pcm_out ? pcm_out + (size_t)i * half : NULL,
                                       pwr_out ? pwr_out + i : NULL, NULL);
        if (valid_out) valid_out[i] = (uint8_t)ok;
    }
}
