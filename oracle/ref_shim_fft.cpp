// TEST INFRASTRUCTURE. Second reference shim: the REFERENCE's own src/fft_impl.cpp (class FFT / class FFTW: window,
// load_*_input, 1/N, vec_log2, power_and_quantize, half_and_quantize pyramid) and src/signal.cpp
// (AudioClient::send_audio: slice placement, parity flip, overlap-add, demodulators, NaN guard, DC, AGC, int16)
// compiled unmodified, where they lie under /root/reference, against the stand-in headers of oracle/ref_stub
// (fftw3.h, websocketpp, FLAC++, zstd, nlohmann, boost). Built by oracle/Makefile into
// oracle/_ref/libphantom_ref_fft.so. Nothing of the reference is copied into the repo.
//
// What is NOT the reference here: the DFT behind fftwf_execute (FFTW3f is absent - the test plugs the oracle's own
// DFT in through ref_set_dft, so both sides transform with identical arithmetic and everything around the transform
// is compared bit for bit), the websocket/encoder plumbing (a PacketSender and a FlacEncoder::process that record
// what send_audio hands to encoder->set_data / encoder->process, src/signal.cpp:287-291), src/client.cpp (glaze) and
// generate_unique_id (src/utils.cpp).
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "fft.h"
#include "signal.h"

// ---- fftw3 stand-in --------------------------------------------------------------------------------
typedef void (*dft_fn)(const float *in, float *out, long n, int sign);
static dft_fn g_dft = nullptr;
struct stub_fftwf_plan_s {
    int kind;  // 0 c2c, 1 r2c, 2 c2r
    int n, sign;
    float *in, *out;
    std::vector<float> a, b;
};
extern "C" {
void ref_set_dft(dft_fn f) { g_dft = f; }
void *fftwf_malloc(size_t n) { return std::aligned_alloc(64, (n + 63) / 64 * 64); }
void fftwf_free(void *p) { std::free(p); }
void fftwf_plan_with_nthreads(int) {}
int fftwf_init_threads(void) { return 1; }
static fftwf_plan make_plan(int kind, int n, int sign, void *in, void *out) {
    stub_fftwf_plan_s *p = new stub_fftwf_plan_s();
    p->kind = kind;
    p->n = n;
    p->sign = sign;
    p->in = static_cast<float *>(in);
    p->out = static_cast<float *>(out);
    p->a.resize(2 * (size_t)n);
    p->b.resize(2 * (size_t)n);
    return p;
}
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned) { return make_plan(0, n, sign, in, out); }
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned) { return make_plan(1, n, -1, in, out); }
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned) { return make_plan(2, n, +1, in, out); }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
void fftwf_execute(const fftwf_plan p) {
    if (!g_dft) std::abort();
    const int n = p->n;
    if (p->kind == 0) {
        g_dft(p->in, p->out, n, p->sign);
    } else if (p->kind == 1) {  // r2c: bins 0 .. n/2 of the transform of the real sequence
        for (int i = 0; i < n; i++) {
            p->a[2 * i] = p->in[i];
            p->a[2 * i + 1] = 0.f;
        }
        g_dft(p->a.data(), p->b.data(), n, -1);
        std::memcpy(p->out, p->b.data(), sizeof(float) * (n + 2));
    } else {  // c2r: Hermitian extension of in[0 .. n/2] (Im of DC and Nyquist are not part of FFTW's definition), sign +1
        float *full = p->a.data();
        const float *in = p->in;
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
        g_dft(full, p->b.data(), n, +1);
        for (int t = 0; t < n; t++) p->out[t] = p->b[2 * t];
    }
}
}

// ---- pieces of the reference that live in files this shim does not compile ---------------------------
std::string generate_unique_id() { return "oracle-pin"; }  // src/utils.cpp:12 (random id, irrelevant to the DSP)
Client::Client(connection_hdl hdl, PacketSender &sender, conn_type type)  // src/client.cpp:5-6
    : type{type}, hdl{hdl}, sender{sender}, frame_num{0}, mute{false} {}
void Client::on_window_message(int, std::optional<double> &, int, std::optional<int> &) {}
void Client::on_demodulation_message(std::string &) {}
void Client::on_userid_message(std::string &) {}
void Client::on_mute(bool) {}
void PacketSender::send_binary_packet(connection_hdl hdl, const void *data, size_t size) { send_binary_packet(hdl, {{data, size}}); }
void PacketSender::send_text_packet(connection_hdl hdl, const std::string &data) { send_text_packet(hdl, {data}); }

namespace {
struct Capture {
    bool sent = false;
    uint64_t frame_num = 0;
    int l = 0, r = 0;
    double m = 0, pwr = 0;
    std::vector<int32_t> pcm;
};
Capture g_cap;
struct Sender : PacketSender {
    waterfall_slices_t wf;
    waterfall_mutexes_t wfm;
    signal_slices_t slices;
    std::mutex mtx;
    void send_binary_packet(connection_hdl, const std::initializer_list<std::pair<const void *, size_t>> &) override {}
    void send_text_packet(connection_hdl, const std::initializer_list<std::string> &) override {}
    std::string ip_from_hdl(connection_hdl) override { return ""; }
    void log(connection_hdl, const std::string &) override {}
    waterfall_slices_t &get_waterfall_slices() override { return wf; }
    waterfall_mutexes_t &get_waterfall_slice_mtx() override { return wfm; }
    signal_slices_t &get_signal_slices() override { return slices; }
    std::mutex &get_signal_slice_mtx() override { return mtx; }
    void broadcast_signal_changes(const std::string &, int, double, int) override {}
};
}  // namespace

// src/audio.cpp stand-ins: record what send_audio hands over (src/signal.cpp:287-291)
AudioEncoder::AudioEncoder(websocketpp::connection_hdl hdl, PacketSender &sender) : hdl{hdl}, sender{sender}, stream{nullptr} {}
AudioEncoder::~AudioEncoder() {}
void AudioEncoder::set_data(uint64_t frame_num, int l, double m, int r, double pwr) {
    g_cap.frame_num = frame_num;
    g_cap.l = l;
    g_cap.m = m;
    g_cap.r = r;
    g_cap.pwr = pwr;
}
int AudioEncoder::send(const void *, size_t, unsigned) { return 0; }
FLAC__StreamEncoderWriteStatus FlacEncoder::write_callback(const FLAC__byte[], size_t, unsigned, unsigned) {
    return FLAC__STREAM_ENCODER_WRITE_STATUS_OK;
}
int FlacEncoder::process(int32_t *data, size_t size) {
    g_cap.sent = true;
    g_cap.pcm.assign(data, data + size);
    return 0;
}
int FlacEncoder::finish_encoder() { return 0; }
FlacEncoder::~FlacEncoder() {}

struct RefAudio {
    Sender sender;
    std::shared_ptr<AudioClient> client;
};

extern "C" {
// ---- class FFTW (src/fft.h:88-104, src/fft_impl.cpp:76-183) ----
void *ref_fftw_create(size_t size, int downsample_levels, int brightness_offset, size_t additional, int is_real) {
    FFTW *f = new FFTW(size, 1, downsample_levels, brightness_offset);
    f->set_output_additional_size(additional);  // src/spectrumserver.cpp:214
    if (is_real) f->plan_r2c(FFTW_MEASURE | FFTW_DESTROY_INPUT);  // src/fft.cpp:25-29
    else f->plan_c2c(FFT::FORWARD, FFTW_MEASURE | FFTW_DESTROY_INPUT);
    return f;
}
void ref_fftw_destroy(void *f) { delete static_cast<FFTW *>(f); }
void ref_fftw_load_real(void *f, float *a1, float *a2) { static_cast<FFTW *>(f)->load_real_input(a1, a2); }
void ref_fftw_load_complex(void *f, float *a1, float *a2) { static_cast<FFTW *>(f)->load_complex_input(a1, a2); }
void ref_fftw_execute(void *f) { static_cast<FFTW *>(f)->execute(); }
float *ref_fftw_input(void *f) { return static_cast<FFTW *>(f)->get_input_buffer(); }
float *ref_fftw_output(void *f) { return static_cast<FFTW *>(f)->get_output_buffer(); }
int8_t *ref_fftw_quantized(void *f) { return static_cast<FFTW *>(f)->get_quantized_buffer(); }

// ---- class AudioClient (src/signal.h:51-128, src/signal.cpp) ----
void *ref_audio_create(int is_real, int audio_fft_size, int audio_max_sps, int fft_result_size) {
    RefAudio *a = new RefAudio();
    a->client = std::make_shared<AudioClient>(connection_hdl(), a->sender, AUDIO_FLAC, is_real != 0, audio_fft_size,
                                              audio_max_sps, fft_result_size);
    a->client->it = a->sender.slices.insert({{0, 0}, a->client});  // src/websocket.cpp:141-142
    return a;
}
void ref_audio_destroy(void *a) {
    RefAudio *r = static_cast<RefAudio *>(a);
    r->sender.slices.clear();
    delete r;
}
void ref_audio_set_range(void *a, int l, double m, int r) { static_cast<RefAudio *>(a)->client->set_audio_range(l, m, r); }
void ref_audio_set_demodulation(void *a, int mode) {
    static_cast<RefAudio *>(a)->client->set_audio_demodulation(static_cast<demodulation_mode>(mode));
}
void ref_audio_on_demodulation_message(void *a, const char *name) {
    std::string s(name);
    static_cast<RefAudio *>(a)->client->on_demodulation_message(s);
}
int ref_audio_on_window_message(void *a, int l, double m, int r) {
    std::optional<double> mm = m;
    std::optional<int> lvl;
    g_cap.sent = false;
    static_cast<RefAudio *>(a)->client->on_window_message(l, mm, r, lvl);
    return 0;
}
// buf = &fft_buffer[(l + base_idx) % fft_result_size], as src/websocket.cpp:182 forms it. Returns 1 when
// encoder->process was reached (0: dropped by the NaN guard), with the int32 PCM and the power it was given.
int ref_audio_send(void *a, float *buf, size_t frame_num, int32_t *pcm_out, float *pwr_out) {
    g_cap.sent = false;
    static_cast<RefAudio *>(a)->client->send_audio(reinterpret_cast<std::complex<float> *>(buf), frame_num);
    if (!g_cap.sent) return 0;
    std::memcpy(pcm_out, g_cap.pcm.data(), sizeof(int32_t) * g_cap.pcm.size());
    *pwr_out = (float)g_cap.pwr;
    return 1;
}
}
