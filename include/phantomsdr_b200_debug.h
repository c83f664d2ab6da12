/* Tuning knobs and profiling hooks of libphantomsdr_b200 - NOT part of the drop-in boundary (include/phantomsdr_b200.h).
 * They select between measured kernel variants (every variant is covered bit-for-bit by tests/test_gpu_variants.py) and
 * isolate stages for timing; values and numbering may change between builds. The environment variable
 * B200_OPTS="knob=value,knob=value" applies b200_debug_option after planning (used by the probes under tools/). */
#ifndef PHANTOMSDR_B200_DEBUG_H
#define PHANTOMSDR_B200_DEBUG_H
#include "phantomsdr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OPT_FUSED_PYRAMID 5 /* waterfall source (c2c; r2c always splits + quantises in the pyramid kernel):
                                   0 = the pyramid kernel re-reads the spectrum with aligned 128-bit loads; 1 = |X|^2, log and
                                   levels 0..log2(T)-1 inside pass 2, upper levels in a small kernel; 2 = FFT pass 2 stores
                                   |X|^2 straight from its registers into a compact power plane that the pyramid kernel
                                   reads. Default -1 = auto: 0 with the TMA passes, 2 with the generic ones. Measured at 2^20,
                                   64 frames/launch, three-stage pass 2: 8.5 (mode 0) / 9.0 (mode 2) us per frame. */
#define B200_OPT_TMA 6           /* 2^20-point transforms: 2 (default) = TMA-fed pass 1 + three-stage pass 2 (one CTA of two
                                   consumer groups per SM), 3 = same with the waterfall pyramid fused into pass 2 (per-frame
                                   completion counters; measured slower), 4 = c2c: pass 1, pass 2 and the pyramid of consecutive frames
                                   in ONE persistent dataflow-scheduled launch (fft_stream.cuh: the four-step intermediate and the
                                   spectrum the quantiser reads stay in L2), 1 = two-CTA pass 2, 0 = generic */
#define B200_OPT_TAIL_PIPELINE 7 /* 1 (default): frame-skewed software pipeline for the DC/AGC tails when >= 4 frames per call */
#define B200_OPT_PACKED_MATH 9   /* bit0 (default 1): waterfall quantiser on the packed-f32 pipe (FMUL2/FADD2), same IEEE rounding per lane;
                                   bit1 (default 0, not yet validated on a GPU): table-driven quantiser for levels 0..2 (b200_quant_table) */
#define B200_OPT_FWD_LANES 10    /* 1..4 streams that the sub-batches of one device batch alternate over (default 1) */
#define B200_OPT_FWD_SUB_FRAMES 11 /* frames per forward launch group inside a device batch (default: the whole batch) */
#define B200_OPT_PASS1_ORDER 12  /* tuning: work-item order of the TMA pass 1 (0 default: column tile sticky, frames swept together) */
#define B200_OPT_PYRAMID_LAG 13  /* B200_OPT_TMA 3: frames between a pass-2 tile and the pyramid blocks that ride on it (default 2) */
#define B200_OPT_DEMOD_CHUNK 19  /* frames per warp task of the frame-chunked demodulation kernel (default -1: chosen per launch from the
                                   client count; 0 = the sequential one-CTA-per-client kernel only). Results are bit-identical either way. */
#define B200_OPT_CLIENT_STAGE_MASK 20 /* profiling aid: bit0 = demodulation kernels, bit1 = tail kernel; default 3 */
#define B200_OPT_FWD_SMS 21      /* SMs the persistent pass-2 kernel sizes its grid for (0 = all): with the tail kernel resident on
                                  * some SMs a one-CTA-per-SM grid would run in two waves */
#define B200_OPT_PASS1_SPLIT 22  /* CTAs per column tile of the TMA pass 1 (default 16 = work units of four frames at 64 frames per launch) */
#define B200_OPT_DEMOD_GENERIC 23 /* comparison aid: 1 = run-time-plan demodulation kernel even for audio_fft_size 360 */
#define B200_OPT_TAIL_SMEM_KB 24 /* shared memory a tail-pipeline CTA asks for (default 224 KB = the whole SM: keeps every other CTA off its schedulers) */
#define B200_OPT_R2C_SPLIT_KERNEL 25 /* r2c: 1 (default) = Hermitian split in a streaming kernel of its own, then the c2c pyramid kernel on
                                      * the send frames; 0 = split and quantiser in one kernel (bit-identical, measured slower) */
#define B200_OPT_STREAM_GRID 14  /* B200_OPT_TMA 4: CTAs of the stream kernel (0 = one per SM) */
#define B200_OPT_STREAM_LAG1 15  /* ... frame slots between pass 1 and pass 2 of a frame in the item order (default 2) */
#define B200_OPT_STREAM_LAG2 16  /* ... between pass 1 and the quantiser (default 4) */
#define B200_OPT_STREAM_RING 17  /* ... frame slots of the L2-resident ring that holds the four-step intermediate (default 5) */
#define B200_OPT_STAGE_MASK 4    /* profiling aid: bit0 = FFT pass 1, bit1 = pass 2, bit2 = pyramid; default 7 */
/* Sets a tuning knob (the B200_OPT_* values above) or an engine option. 0 or a negative B200_E* code. */
int b200_debug_option(b200_engine *e, int knob, int value);

/* Profiling aid: accumulate, per stage warp of group 0 of the client tail pipeline (stage ids: 0 load, 1 sum1, 2 sum2,
 * 3 block, 4 gain, 5..8 peak, 9..12 out, 13..14 suffix), the SM-clock cycles spent waiting for input, waiting for ring
 * space, and in total: out[3 * stage + {0, 1, 2}]. out (nullable) receives the totals so far. */
int b200_debug_tail_profile(b200_engine *e, int enable, long long out[64]);

#ifdef __cplusplus
}
#endif
#endif /* PHANTOMSDR_B200_DEBUG_H */
