// b200_fft.hpp - header-only C++ adapter from the C ABI (phantomsdr_b200.h) to the reference's
// FFT-backend interface, `class FFT` (reference src/fft.h:33-63).
//
// Inside the reference tree: `#include "fft.h"` first (its include guard FFT_H is then defined) and
// B200FFT derives from the reference's own abstract `FFT`, so broadcast_server::fft_task
// (src/fft.cpp:10-119) can drive it unmodified:
//     fft = std::make_unique<B200FFT>(fft_size, fft_threads, downsample_levels, brightness_offset);
// Stand-alone (this repo's host driver and tests): the same class without a base.
//
// Error behaviour follows the reference: methods return int (0 = OK) that the caller never checks
// (src/fft.cpp:25-29,61,68,90); construction failure throws std::runtime_error like cuFFT's ctor
// ("No CUDA devices found", src/fft_cuda.cu:10-13).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "phantomsdr_b200.h"

#ifdef FFT_H
#define B200_FFT_BASE : public FFT
#define B200_FFT_BASE_INIT FFT(size, nthreads, downsample_levels, brightness_offset),
#define B200_OVERRIDE override
#else
#define B200_FFT_BASE
#define B200_FFT_BASE_INIT
#define B200_OVERRIDE
#endif

class B200FFT B200_FFT_BASE {
  public:
#ifndef FFT_H
    enum direction { FORWARD, BACKWARD };  // FFT::direction, src/fft.h:35
#endif
    B200FFT(size_t size, int nthreads, int downsample_levels, int brightness_offset, int device = 0)
        : B200_FFT_BASE_INIT engine_(nullptr) {
        int rc = b200_engine_create(&engine_, size, nthreads, downsample_levels, brightness_offset, device);
        if (rc != 0) throw std::runtime_error(b200_last_error());
    }
    virtual ~B200FFT() { b200_engine_destroy(engine_); }
    B200FFT(const B200FFT &) = delete;
    B200FFT &operator=(const B200FFT &) = delete;

    virtual float *malloc(size_t nfloats) B200_OVERRIDE { return b200_malloc(engine_, nfloats); }
    virtual void free(float *buf) B200_OVERRIDE { b200_free(engine_, buf); }
    virtual int plan_c2c(direction d, int options) B200_OVERRIDE {
        return b200_plan_c2c(engine_, d == FORWARD ? B200_FORWARD : B200_BACKWARD, options);
    }
    virtual int plan_r2c(int options) B200_OVERRIDE { return b200_plan_r2c(engine_, options); }
    virtual void set_output_additional_size(size_t n) B200_OVERRIDE { b200_set_output_additional_size(engine_, n); }
    virtual float *get_output_buffer() B200_OVERRIDE { return b200_get_output_buffer(engine_); }
    virtual int8_t *get_quantized_buffer() B200_OVERRIDE { return b200_get_quantized_buffer(engine_); }
    virtual int load_real_input(float *a1, float *a2) B200_OVERRIDE { return b200_load_real_input(engine_, a1, a2); }
    virtual int load_complex_input(float *a1, float *a2) B200_OVERRIDE {
        return b200_load_complex_input(engine_, a1, a2);
    }
    virtual int execute() B200_OVERRIDE { return b200_execute(engine_); }

    // Beyond class FFT: the batched signal / waterfall slots and the device-resident API.
    b200_engine *engine() { return engine_; }

  private:
    b200_engine *engine_;
};
