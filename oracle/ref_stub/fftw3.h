// TEST INFRASTRUCTURE. Stand-in for <fftw3.h> so that the REFERENCE's own src/fft_impl.cpp and src/signal.cpp compile
// unmodified where they lie (oracle/Makefile target `ref`). FFTW3f itself is a system library that is neither under
// /root/reference nor in this image; only the declarations the reference uses exist here, and fftwf_execute forwards to a
// DFT supplied by the test (oracle/ref_shim_fft.cpp: the oracle's own mixed-radix DFT), so that everything AROUND the
// transform - loaders, 1/N, vec_log2, quantiser, pyramid, slice placement, overlap-add, demodulators - is the
// reference's compiled code and can be compared bit for bit.
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct stub_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_ESTIMATE (1U << 6)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
void fftwf_plan_with_nthreads(int nthreads);
int fftwf_init_threads(void);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
