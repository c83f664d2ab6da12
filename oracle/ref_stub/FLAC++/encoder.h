// TEST INFRASTRUCTURE: just enough of FLAC++ for src/audio.h's class FlacEncoder and the setter calls in
// AudioClient's constructor (src/signal.cpp:16-27) to compile. Nothing is encoded.
#pragma once
#include <cstddef>
#include <cstdint>
typedef uint8_t FLAC__byte;
typedef int32_t FLAC__int32;
enum FLAC__StreamEncoderWriteStatus { FLAC__STREAM_ENCODER_WRITE_STATUS_OK = 0, FLAC__STREAM_ENCODER_WRITE_STATUS_FATAL_ERROR };
namespace FLAC { namespace Encoder {
class Stream {
  public:
    virtual ~Stream() {}
    bool set_channels(unsigned) { return true; }
    bool set_verify(bool) { return true; }
    bool set_compression_level(unsigned) { return true; }
    bool set_sample_rate(unsigned) { return true; }
    bool set_bits_per_sample(unsigned) { return true; }
    bool set_streamable_subset(bool) { return true; }
    int init() { return 0; }
    bool finish() { return true; }
    bool process_interleaved(const FLAC__int32 *, unsigned) { return true; }
  protected:
    virtual FLAC__StreamEncoderWriteStatus write_callback(const FLAC__byte buffer[], size_t bytes, unsigned samples,
                                                          unsigned current_frame) = 0;
};
} }
