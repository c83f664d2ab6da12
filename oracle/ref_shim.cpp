// TEST INFRASTRUCTURE. extern "C" shim around the parts of the REFERENCE that compile from their
// own sources with no third-party dependency. Built by oracle/Makefile into
// oracle/_ref/libphantom_ref.so directly from /root/reference (nothing is copied into the repo):
//   src/utils/dsp.cpp              (Hann window, FM discriminator, AM envelope, negate/add, float->int16)
//   src/utils/audioprocessing.cpp  (AGC)
//   src/utils.h                    (MovingAverage / DCBlocker; needs boost::circular_buffer ->
//                                   oracle/ref_stub/boost/circular_buffer.hpp)
// tests/test_oracle_vs_ref.py uses it to pin oracle/phantom_oracle.c bit-for-bit.
#include <complex>
#include <cstdint>
#include <cstring>

#include "utils/dsp.h"
#include "utils/audioprocessing.h"
#include "utils.h"

extern "C" {
void ref_build_hann_window(float *arr, int num) { build_hann_window(arr, num); }
void ref_polar_discriminator_fm(float *buf, float prev_re, float prev_im, float *output, size_t len) {
    polar_discriminator_fm(reinterpret_cast<std::complex<float> *>(buf), std::complex<float>(prev_re, prev_im), output,
                           len);
}
void ref_dsp_negate_float(float *arr, size_t len) { dsp_negate_float(arr, len); }
void ref_dsp_negate_complex(float *arr, size_t len) {
    dsp_negate_complex(reinterpret_cast<std::complex<float> *>(arr), len);
}
void ref_dsp_add_float(float *a, float *b, size_t len) { dsp_add_float(a, b, len); }
void ref_dsp_add_complex(float *a, float *b, size_t len) {
    dsp_add_complex(reinterpret_cast<std::complex<float> *>(a), reinterpret_cast<std::complex<float> *>(b), len);
}
void ref_dsp_am_demod(float *arr, float *output, size_t len) {
    dsp_am_demod(reinterpret_cast<std::complex<float> *>(arr), output, len);
}
void ref_dsp_float_to_int16(float *arr, int32_t *output, float mult, size_t len) {
    dsp_float_to_int16(arr, output, mult, len);
}

void *ref_agc_create(float level, float attack_ms, float release_ms, float lookahead_ms, float sr) {
    return new AGC(level, attack_ms, release_ms, lookahead_ms, sr);
}
void ref_agc_destroy(void *a) { delete static_cast<AGC *>(a); }
void ref_agc_process(void *a, float *arr, size_t len) { static_cast<AGC *>(a)->process(arr, len); }
void ref_agc_reset(void *a) { static_cast<AGC *>(a)->reset(); }

void *ref_dc_create(int delay) { return new DCBlocker<float>(delay); }
void ref_dc_destroy(void *d) { delete static_cast<DCBlocker<float> *>(d); }
void ref_dc_remove(void *d, float *arr, int len) { static_cast<DCBlocker<float> *>(d)->removeDC(arr, len); }

// std::accumulate over std::norm exactly as src/signal.cpp:117-119 spells it
float ref_slice_power(float *buf, int len) {
    std::complex<float> *b = reinterpret_cast<std::complex<float> *>(buf);
    return std::accumulate(b, b + len, 0.0f, [](float a, std::complex<float> &x) { return a + std::norm(x); });
}
}
