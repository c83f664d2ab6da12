/*
 * phantomsdr_b200.h - C ABI of the B200-native spectrum / channeliser engine.
 *
 * This is the drop-in boundary for ONE hot path of PhantomSDR: the FFT backend behind
 * src/fft.h:33-63 (class FFT) and the two slot callbacks it feeds,
 * AudioClient::send_audio (src/signal.h:67, src/signal.cpp:102-298) and
 * WaterfallClient::send_waterfall (src/waterfall.h:13, src/waterfall.cpp:44-51).
 * Citations are relative to the reference tree. INTEGRATION.md shows the C++ adapter a
 * maintainer would add (class B200FFT : public FFT -> these calls).
 *
 * Conventions (mirroring the reference, src/fft.h:39-48): every call returns int, 0 = OK,
 * negative = -errno style failure (B200_E*). No C++ types, no torch types, no exceptions cross
 * this boundary. All entry points of one engine are single-caller (the reference calls them
 * from its one fft_thread, src/spectrumserver.cpp:245). There is NO CPU fallback: if no CUDA
 * device is usable, b200_engine_create fails with B200_ENODEV (the reference's cuFFT ctor
 * throws "No CUDA devices found", src/fft_cuda.cu:10-13).
 */
#ifndef PHANTOMSDR_B200_H
#define PHANTOMSDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* error codes (negative errno values) */
#define B200_OK 0
#define B200_EINVAL (-22)
#define B200_ENOMEM (-12)
#define B200_ENODEV (-19)
#define B200_ENOTSUP (-95)
#define B200_ESTATE (-1)  /* call out of order (e.g. execute before plan) */
#define B200_ECUDA (-5)   /* a CUDA runtime call failed; see b200_last_error() */

/* FFT::direction, src/fft.h:35 */
#define B200_FORWARD 0
#define B200_BACKWARD 1

/* enum demodulation_mode { USB, LSB, AM, FM }, src/client.h:43 */
#define B200_USB 0
#define B200_LSB 1
#define B200_AM 2
#define B200_FM 3

/* raw sample formats of SampleConverter<T>, src/samplereader.cpp:29-66 */
#define B200_FMT_F32 0
#define B200_FMT_U8 1
#define B200_FMT_S8 2
#define B200_FMT_U16 3
#define B200_FMT_S16 4

typedef struct b200_engine b200_engine;

int b200_abi_version(void);
/* Human-readable text of the last failure on this thread (never NULL). */
const char *b200_last_error(void);
/* Number of usable CUDA devices (0 if none / driver missing). */
int b200_device_count(void);

/* ------------------------------------------------------------------------------------------
 * FFT backend group  - replaces class FFTW / class cuFFT (src/fft.h:93-145)
 * ------------------------------------------------------------------------------------------ */

/* FFT::FFT(size, nthreads, downsample_levels, brightness_offset), src/fft.h:36, src/fft_impl.cpp:63-70.
 * Builds the periodic Hann window on the host exactly as build_hann_window (src/utils/dsp.cpp:6-11)
 * and uploads it; size_log2 = round(log2(size)) + brightness_offset. `size` must be a power of
 * two (2^16..2^23 for c2c, 2^17..2^23 for r2c: complex transform length 2^16..2^23 -> B200_ENOTSUP otherwise).
 * `nthreads` is accepted and ignored. `device` is the CUDA ordinal. */
int b200_engine_create(b200_engine **out, size_t size, int nthreads, int downsample_levels, int brightness_offset,
                       int device);
void b200_engine_destroy(b200_engine *e);

/* FFT::set_output_additional_size, src/fft.h:41, called before planning (src/spectrumserver.cpp:214). */
int b200_set_output_additional_size(b200_engine *e, size_t n);

/* FFT::malloc / FFT::free, src/fft.h:37-38: the caller's 3-deep input ring (src/fft.cpp:17-22).
 * Returns page-locked host memory (as cuFFT::malloc does, src/fft_cuda.cu:22-28); NULL on failure. */
float *b200_malloc(b200_engine *e, size_t nfloats);
void b200_free(b200_engine *e, float *buf);

/* FFT::plan_c2c / plan_r2c, src/fft.h:39-40; called once (src/fft.cpp:25-29). `options` (FFTW flags)
 * is ignored; direction must be B200_FORWARD (the only one the path uses, src/fft.cpp:28). */
int b200_plan_c2c(b200_engine *e, int direction, int options);
int b200_plan_r2c(b200_engine *e, int options);

/* FFT::get_output_buffer / get_quantized_buffer, src/fft.h:43-45. Stable, CPU-dereferenceable
 * (page-locked host mirrors refreshed by b200_execute). Output: complex float32 interleaved,
 * normalised by 1/size, natural FFT order; c2c: size + additional_size bins (the caller performs
 * the wrap memcpy itself, src/fft.cpp:96-97; the engine also fills the tail so device-side
 * clients see it); r2c: size/2 + 1 bins. Quantized: int8 pyramid, level i at offset
 * sum_{j<i}(R >> j), display order (src/fft_impl.cpp:146-173, src/websocket.cpp:207-236). */
float *b200_get_output_buffer(b200_engine *e);
int8_t *b200_get_quantized_buffer(b200_engine *e);

/* FFT::load_real_input / load_complex_input, src/fft.h:46-47: a1 = older half, a2 = newer half
 * (src/fft.cpp:51-71). Host pointers, preferably from b200_malloc. Asynchronous: enqueues the
 * host->device copy of the halves; the window multiply is fused into the FFT's first pass. If a1
 * is the pointer passed as a2 by the previous call, it is assumed unchanged and NOT re-uploaded
 * (the reference ring guarantees this); b200_set_option(B200_OPT_RELOAD_BOTH, 1) disables that. */
int b200_load_real_input(b200_engine *e, const float *a1, const float *a2);
int b200_load_complex_input(b200_engine *e, const float *a1, const float *a2);

/* FFT::execute, src/fft.h:48, src/fft_impl.cpp:144-174. Synchronous: on return the host mirrors
 * selected by B200_OPT_HOST_MIRROR hold the spectrum and the pyramid of the loaded frame. */
int b200_execute(b200_engine *e);

/* Waterfall cadence. The reference computes the pyramid in every FFT::execute but SENDS it only when
 * frame_num % skip_num == 0 (src/fft.cpp:33,102-104; skip_num = 6 at 35 MSPS IQ / 2^20). With skip_num > 1 the engine
 * computes the pyramid - and copies it to the host (b200_execute's mirror, b200_submit_block's pyramid_out rows) - only
 * for those frames; the spectrum and the clients are unaffected. The engine counts frames itself from 0 like fft_task;
 * b200_set_frame_number resynchronises the counter (b200_submit_block takes it from frame_num0). Default 1 = every frame. */
int b200_set_waterfall_cadence(b200_engine *e, int skip_num);
int b200_set_frame_number(b200_engine *e, uint64_t frame_num);

/* Engine options (behaviour a caller may need; the tuning and profiling knobs live in phantomsdr_b200_debug.h). */
#define B200_OPT_RELOAD_BOTH 1   /* 0 (default) / 1 */
#define B200_OPT_HOST_MIRROR 2   /* bitmask: 1 = spectrum, 2 = pyramid; default 3 */
#define B200_OPT_INPUT_FORMAT 3  /* B200_FMT_*: format of the halves given to b200_load_raw_input */
#define B200_OPT_PEER_STORES 8   /* 1 (default): FFT pass 2 stores the peers' sub-bands itself; 0: leave it to b200_push_peers */
#define B200_OPT_PCM16 18        /* N2: 1 = PCM rows are int16 [frame][slot][n/2] (half the D2H bytes; values identical), 0 (default) = int32
                                   as AudioEncoder::process takes them (src/audio.h:26-27). Pipelined tail kernel only. */
int b200_set_option(b200_engine *e, int option, int value);

/* SURVEY 8f N1 - SampleConverter<T>::read fused into the first FFT pass: halves are raw
 * u8/s8/u16/s16 samples (format from B200_OPT_INPUT_FORMAT), converted on the GPU exactly as
 * src/samplereader.cpp:29-40,59-66 ((x ^ topbit) as signed / 2^(bits-1)). */
int b200_load_raw_input(b200_engine *e, const void *a1, const void *a2);

/* ------------------------------------------------------------------------------------------
 * Device-resident streaming form of the same path (inputs already in HBM).
 * ------------------------------------------------------------------------------------------ */

/* Device buffers owned by the engine (valid after planning). */
void *b200_device_spectrum(b200_engine *e);  /* float2[R + additional] (c2c) / float2[size/2 + 1] (r2c) */
void *b200_device_quantized(b200_engine *e); /* int8 pyramid */
void *b200_device_hop_ring(b200_engine *e);  /* hop ring: nhops x hop_floats float32 */
size_t b200_hop_floats(b200_engine *e);      /* floats per hop: size (c2c) or size/2 (r2c) */
size_t b200_spectrum_bins(b200_engine *e);   /* bins incl. tail */
size_t b200_pyramid_bytes(b200_engine *e);
/* (Re)allocate the device hop ring with `nhops` hops (>= 2; default 3). */
int b200_set_hop_ring(b200_engine *e, size_t nhops);
/* Forward FFT + pyramid of the frame made of hops (hop_index, hop_index+1) mod nhops of the device
 * ring. Asynchronous on the engine's stream; no host copies. Same kernels as b200_execute. */
int b200_execute_device(b200_engine *e, size_t hop_index);
/* Streaming batch: `nframes` consecutive 50%-overlapped frames (frame f = hops hop_index+f,
 * hop_index+f+1) in one launch per kernel. b200_set_batch_frames(F) sizes the device outputs for
 * F frames (default 1): frame f's spectrum is at b200_device_spectrum() + f*b200_spectrum_stride()
 * float2, its pyramid at b200_device_quantized() + f*b200_pyramid_stride() bytes. */
int b200_set_batch_frames(b200_engine *e, int max_frames);
int b200_execute_device_batch(b200_engine *e, size_t hop_index, int nframes);
size_t b200_spectrum_stride(b200_engine *e);
size_t b200_pyramid_stride(b200_engine *e);
/* Software pipeline: `banks` (1..4, default 1) copies of the batch outputs. With banks > 1 the client
 * kernels run on a second internal stream, so the clients of batch k overlap the forward FFT of batch
 * k+1; the engine orders the two streams per bank with events. b200_select_bank picks the bank that the
 * following execute_device / clients_execute_device / device_spectrum / device_quantized calls use.
 * b200_bank_acquire makes the forward stream wait until the selected bank's previous clients are done
 * (implicit in execute_device; call it before writing the bank yourself, e.g. a broadcast on a non-ingest
 * rank). b200_join_streams makes the forward stream wait for all outstanding client work (e.g. before
 * recording a timing event on it). */
int b200_set_pipeline(b200_engine *e, int banks);
int b200_select_bank(b200_engine *e, int bank);
int b200_bank_acquire(b200_engine *e);
int b200_join_streams(b200_engine *e);
/* Make the client stream wait for a caller-owned cudaEvent_t (e.g. the end of a collective that fills the selected
 * bank on a communication stream) before the next b200_clients_execute_device. */
int b200_client_stream_wait_event(b200_engine *e, void *cuda_event);
/* Wait for everything enqueued on the engine's stream. */
int b200_sync(b200_engine *e);
/* The engine's cudaStream_t (as void*), so callers can order their own work / events on it. */
void *b200_stream(b200_engine *e);
/* Run on the caller's stream instead (e.g. the framework's current stream); NULL restores the
 * engine's own. The engine synchronises its previous stream first. */
int b200_set_stream(b200_engine *e, void *cuda_stream);
/* Adopt `dev_ptr` (device memory of b200_spectrum_bins() float2) as the spectrum buffer the client
 * kernels read - used on non-ingest ranks, where the frame arrives by NCCL broadcast / peer store.
 * Pass NULL to go back to the engine-owned buffer. */
int b200_bind_spectrum(b200_engine *e, void *dev_ptr);
/* Ingest rank only: additionally store every spectrum frame into these peer buffers (device
 * pointers valid on this device, e.g. from b200_ipc_open) from inside the last FFT pass, so the
 * NVLink transfer overlaps the butterflies. npeers = 0 turns it off. */
int b200_set_peer_spectra(b200_engine *e, int npeers, void *const *dev_ptrs);
/* Restrict what peer `peer` receives to the bins its clients read: two half-open ranges of spectrum indices
 * (the IQ wrap tail is indices R .. R+additional). Default after b200_set_peer_spectra: everything. */
int b200_set_peer_ranges(b200_engine *e, int peer, uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1);
/* Base of the spectrum allocation (what CUDA IPC exports) and the byte offset of bank 0 / frame 0 / bin 0 in it. */
void *b200_device_spectrum_base(b200_engine *e);
size_t b200_device_spectrum_offset(b200_engine *e);
/* 64 uint64 flags in device memory (zeroed; exportable with b200_ipc_export) and stream-ordered operations on
 * flags that may live in this GPU's memory or in an IPC-mapped peer's: b200_enqueue_signal stores `value` after
 * everything already enqueued on the chosen stream (0 = forward, 1 = client, 2 = copy stream); b200_enqueue_wait holds the
 * stream until every flag >= min_value (gives up after timeout_ms and latches b200_flag_error). */
/* Copy-engine form of the scatter: after the forward work already enqueued, DMA every peer's sub-band of the
 * selected bank (`nframes` frames) into that peer's bank on the engine's copy stream (selector 2 of the flag
 * calls). Use with b200_set_option(B200_OPT_PEER_STORES, 0). */
int b200_push_peers(b200_engine *e, int nframes);
/* Pull form of the scatter, called on a CLIENT rank: copy this rank's sub-band (two half-open bin ranges, as
 * b200_set_peer_ranges) of the selected bank from the ingest rank's spectrum (an IPC mapping of its
 * b200_device_spectrum_base() + b200_device_spectrum_offset()) into the same bank here, on the client stream - behind
 * the b200_enqueue_wait for the ingest rank's "bank ready" flag and in front of b200_clients_execute_device. The copy
 * is done by THIS GPU's copy engine, so the pulls of seven client ranks run on seven engines instead of the ingest
 * rank's own. */
int b200_pull_spectrum(b200_engine *e, const void *remote_spectrum, int nframes, uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1);
void *b200_flag_buffer(b200_engine *e);
int b200_enqueue_signal(b200_engine *e, int client_stream, void *const *flag_ptrs, int n, uint64_t value);
int b200_enqueue_wait(b200_engine *e, int client_stream, void *const *flag_ptrs, int n, uint64_t min_value, int timeout_ms);
int b200_flag_error(b200_engine *e);
/* CUDA IPC helpers for the above (64-byte handles). */
int b200_ipc_export(b200_engine *e, const void *dev_ptr, uint8_t handle[64]);
int b200_ipc_open(b200_engine *e, const uint8_t handle[64], void **dev_ptr);
int b200_ipc_close(b200_engine *e, void *dev_ptr);

/* ------------------------------------------------------------------------------------------
 * Signal slot group - replaces N x AudioClient::send_audio (src/signal.cpp:102-298) and the
 * slice index math of broadcast_server::signal_loop (src/websocket.cpp:156-185).
 * ------------------------------------------------------------------------------------------ */

/* AudioClient::AudioClient scratch/state for up to max_clients slots (src/signal.cpp:7-79):
 * audio_fft_size (multiple of 4, src/spectrumserver.cpp:151), DCBlocker(audio_max_sps/750*2),
 * AGC(0.2, 50 ms, 300 ms, 200 ms, audio_max_sps). */
int b200_clients_create(b200_engine *e, int max_clients, int audio_fft_size, int audio_max_sps);
/* New connection in `slot` (src/websocket.cpp:129-150): zeroed state, range and mode set. */
int b200_client_open(b200_engine *e, int slot, int l, double audio_mid, int r, int demodulation);
/* AudioClient::on_window_message, src/signal.cpp:300-314: validates 0 <= l <= r < R, r - l <= n;
 * returns B200_EINVAL (state untouched) when the reference would silently ignore the message. */
int b200_client_set_window(b200_engine *e, int slot, int l, double audio_mid, int r);
/* AudioClient::on_demodulation_message, src/signal.cpp:316-328: sets the mode and resets the AGC. */
int b200_client_set_demodulation(b200_engine *e, int slot, int demodulation);
/* AudioClient::on_close, src/signal.cpp:330-336. */
int b200_client_close(b200_engine *e, int slot);

/* One signal_loop() pass: every open client, visited in the reference's (l, r)-sorted multimap
 * order, demodulates frame `frame_num` (the server-global counter, src/websocket.cpp:182) from the
 * current spectrum. Host outputs (page-locked preferred), indexed by slot:
 *   pcm_out  [max_clients][audio_fft_size/2] int32  - what encoder->process receives (signal.cpp:291)
 *   pwr_out  [max_clients] float                   - average_power of set_data (signal.cpp:287)
 *   valid_out[max_clients] uint8                   - 0 where the reference drops the frame (NaN guard,
 *                                                    signal.cpp:266-271) or the slot is closed
 * Any of them may be NULL. Synchronous. */
int b200_clients_execute(b200_engine *e, uint64_t frame_num, int32_t *pcm_out, float *pwr_out, uint8_t *valid_out);
/* Asynchronous device-only form (results stay in device memory; see b200_device_pcm). With
 * nframes > 1 the clients run frames frame_num .. frame_num+nframes-1 of the last
 * b200_execute_device_batch in order, carrying their state from frame to frame. */
int b200_clients_execute_device(b200_engine *e, uint64_t frame_num, int nframes);
void *b200_device_pcm(b200_engine *e);   /* int32 [frames][max_clients][n/2] */
void *b200_device_pwr(b200_engine *e);   /* float [frames][max_clients] */
void *b200_device_valid(b200_engine *e); /* uint8 [frames][max_clients] */
/* Copy the results of the last b200_clients_execute_device (frame index `frame` of the batch) to
 * host buffers laid out as in b200_clients_execute. Synchronous. */
int b200_clients_fetch(b200_engine *e, int frame, int32_t *pcm_out, float *pwr_out, uint8_t *valid_out);
/* Asynchronous batch form of b200_clients_fetch: enqueue the device->host copies of the LAST client batch (nframes frames
 * of PCM [frame][slot][n/2], pwr [frame][slot], valid [frame][slot]; page-locked destinations) behind its kernels and
 * return; b200_clients_fetch_wait(slot) blocks until they have landed. Two slots (0 / 1), so the copies of batch k
 * overlap the kernels of batch k + 1. This is the hand-off a pool of AudioEncoder::process threads consumes
 * (src/signal.cpp:287-291, src/audio.cpp:65-82) without a device-wide wait. */
int b200_clients_fetch_async(b200_engine *e, int slot, int nframes, void *pcm_out, float *pwr_out, uint8_t *valid_out);
int b200_clients_fetch_wait(b200_engine *e, int slot);
/* Test tap: audio_real[0..n/2) before DC removal (signal.cpp:274) of the last executed frame,
 * copied to host float [max_clients][n/2]. */
int b200_clients_read_pre_dc(b200_engine *e, float *out);

/* ------------------------------------------------------------------------------------------
 * Pipelined block form of the same host-buffer path: load -> execute -> signal_loop for `nframes` frames per
 * call, with the host->device copy of block k+1, the kernels of block k and the device->host copy of block
 * k's results running on three streams. Up to min(4, (ring halves - 2) / F) blocks may be in flight: needs
 * b200_set_batch_frames(F), b200_set_pipeline(>= 2) and a hop ring of >= 2F+2 halves (4F+2 for four blocks, which keeps
 * the host->device link busy back to back). Halves that lie back to back in one b200_malloc buffer are uploaded in one
 * copy per contiguous run (measured at cfg 2: 39.5 -> 48 GB/s of host->device traffic with both changes).
 *   b200_stream_prime(older_half)     : the half that precedes the first frame (the reference reads two halves
 *                                       before its first transform, src/fft.cpp:50-67)
 *   b200_submit_block(new_halves[nframes], ...): frame f of the block = (previous newest half | new_halves[f]);
 *       outputs (page-locked host memory; any may be NULL): pcm [nframes][max_clients][n/2], pwr / valid
 *       [nframes][max_clients], pyramid [nframes][b200_pyramid_bytes()]. Returns at once.
 *   b200_wait_block()                 : blocks until the OLDEST submitted block's outputs are complete.
 * ------------------------------------------------------------------------------------------ */
int b200_stream_prime(b200_engine *e, const void *older_half);
int b200_submit_block(b200_engine *e, const void *const *new_halves, int nframes, uint64_t frame_num0, int32_t *pcm_out,
                      float *pwr_out, uint8_t *valid_out, int8_t *pyramid_out);
int b200_wait_block(b200_engine *e);

/* ------------------------------------------------------------------------------------------
 * Waterfall slot group - replaces N x WaterfallClient::send_waterfall (src/waterfall.cpp:44-51)
 * and the level-offset math of waterfall_loop (src/websocket.cpp:207-236).
 * ------------------------------------------------------------------------------------------ */
/* Gathers, for client i, quantized[level_offset(level[i]) + l[i] .. + r[i]) into out + out_offsets[i]
 * (host). Synchronous. The labels the reference sends are l << level, r << level (waterfall.cpp:47). */
int b200_waterfall_gather(b200_engine *e, int nclients, const int *level, const int *l, const int *r,
                          const size_t *out_offsets, int8_t *out);

/* Kernel-launch counter since creation (for bench.py's gpu_launches). */
uint64_t b200_launch_count(b200_engine *e);
/* Host-only (no GPU needed): the 2048-cell table of the table-driven waterfall quantiser for one power offset
 * (vec_log2 + power_and_quantize, src/fft_impl.cpp:14-44, as a step function of float_bits(power) >> 20): the int8
 * value is base[c] below lo[c], base[c] + 1 (mod 256) from hi[c] on, and needs the exact arithmetic inside [lo[c], hi[c]).
 * Arrays of 2048 entries each. Used by B200_OPT_PACKED_MATH bit 1 and by the CPU tests. */
int b200_quant_table(int power_offset, uint32_t *lo, uint32_t *hi, uint8_t *base);

#ifdef __cplusplus
}
#endif
#endif /* PHANTOMSDR_B200_H */
