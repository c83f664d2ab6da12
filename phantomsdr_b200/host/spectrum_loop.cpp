// spectrum_loop - C++ host driver above the C ABI that mirrors broadcast_server::fft_task
// (reference src/fft.cpp:10-119) and signal_loop / waterfall_loop (src/websocket.cpp:156-236):
// raw samples on stdin -> 3-buffer ring with the read of the next half overlapped with the transform ->
// load -> execute -> (IQ) wrap copy -> batched audio clients -> waterfall rows every skip_num-th frame.
// What the reference sends to its encoders (int32 PCM per client, int8 waterfall rows) is written to
// stdout as a simple framed binary stream so tests can compare it with the oracle.
//
//   spectrum_loop --sps S --fft N [--real] --format u8|s8|u16|s16|f32 --frames K \
//                 --client l,mid,r,mode ... [--audio-sps 12000] [--waterfall-size 1024] [--brightness 0]
//
// The SampleConverter (src/samplereader.cpp) runs on the GPU: raw halves go through b200_load_raw_input.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <string>
#include <vector>

#include "../../include/b200_fft.hpp"

struct ClientArg {
    int l, r, mode;
    double mid;
};

static void die(const char *msg) {
    fprintf(stderr, "spectrum_loop: %s\n", msg);
    exit(2);
}

int main(int argc, char **argv) {
    long sps = 0, fft_size = 0, frames = 1;
    int is_real = 0, audio_sps = 12000, waterfall_size = 1024, brightness = 0, fmt = B200_FMT_U8;
    std::vector<ClientArg> clients;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * {
            if (i + 1 >= argc) die("missing value");
            return argv[++i];
        };
        if (a == "--sps") sps = atol(next());
        else if (a == "--fft") fft_size = atol(next());
        else if (a == "--real") is_real = 1;
        else if (a == "--frames") frames = atol(next());
        else if (a == "--audio-sps") audio_sps = atoi(next());
        else if (a == "--waterfall-size") waterfall_size = atoi(next());
        else if (a == "--brightness") brightness = atoi(next());
        else if (a == "--format") {
            std::string f = next();
            fmt = f == "u8" ? B200_FMT_U8 : f == "s8" ? B200_FMT_S8 : f == "u16" ? B200_FMT_U16 : f == "s16" ? B200_FMT_S16 : B200_FMT_F32;
        } else if (a == "--client") {
            ClientArg c{};
            if (sscanf(next(), "%d,%lf,%d,%d", &c.l, &c.mid, &c.r, &c.mode) != 4) die("bad --client l,mid,r,mode");
            clients.push_back(c);
        } else die("unknown argument");
    }
    if (sps <= 0 || fft_size <= 0) die("--sps and --fft are required");

    // derived sizes, src/spectrumserver.cpp:99-105,151,185-190 and src/fft.cpp:18,33
    const long fft_result_size = is_real ? fft_size / 2 : fft_size;
    const int audio_max_fft_size = (int)(ceil((double)audio_sps * fft_size / sps / 4.) * 4);
    int downsample_levels = 0;
    for (long cur = fft_result_size; cur >= waterfall_size; cur /= 2) downsample_levels++;
    const int skip_num = std::max(1, (int)floor(((float)sps / fft_size) / 10.) * 2);
    const size_t bytes_per = (fmt == B200_FMT_F32) ? 4 : (fmt == B200_FMT_U16 || fmt == B200_FMT_S16) ? 2 : 1;
    const size_t input_buffer_size = (size_t)fft_size / 2 * (2 - is_real);  // scalars per half

    B200FFT fft(fft_size, 1, downsample_levels, brightness);  // throws without a CUDA device
    fft.set_output_additional_size(audio_max_fft_size);
    b200_engine *e = fft.engine();
    float *input_buffers[3];
    for (auto &b : input_buffers) b = fft.malloc(input_buffer_size);  // pinned; raw formats use a prefix of it
    if (is_real) fft.plan_r2c(0);
    else fft.plan_c2c(B200FFT::FORWARD, 0);
    b200_set_option(e, B200_OPT_INPUT_FORMAT, fmt);
    float *fft_buffer = fft.get_output_buffer();
    int8_t *quantized = fft.get_quantized_buffer();
    const int h = audio_max_fft_size / 2;
    if (!clients.empty()) {
        if (b200_clients_create(e, (int)clients.size(), audio_max_fft_size, audio_sps)) die(b200_last_error());
        for (size_t i = 0; i < clients.size(); i++)
            if (b200_client_open(e, (int)i, clients[i].l, clients[i].mid, clients[i].r, clients[i].mode)) die(b200_last_error());
    }
    std::vector<int32_t> pcm(clients.size() * h);
    std::vector<float> pwr(clients.size());
    std::vector<uint8_t> valid(clients.size());

    auto read_half = [&](float *dst) { return fread(dst, bytes_per, input_buffer_size, stdin) == input_buffer_size; };
    int idx = 0;
    if (!read_half(input_buffers[0]) || !read_half(input_buffers[1])) die("short input");
    size_t frame_num = 0;
    std::future<bool> buffer_read = std::async(std::launch::deferred, [] { return true; });
    for (long it = 0; it < frames; it++) {
        if (!buffer_read.get()) break;
        float *buf0 = input_buffers[idx], *buf1 = input_buffers[(idx + 1) % 3], *buf2 = input_buffers[(idx + 2) % 3];
        buffer_read = std::async(std::launch::async, [&, buf2] { return read_half(buf2); });  // src/fft.cpp:56-67
        b200_load_raw_input(e, buf0, buf1);
        idx = (idx + 1) % 3;
        fft.execute();
        if (!is_real)  // src/fft.cpp:96-97 (the engine already filled the tail; kept for call-order parity)
            memcpy(fft_buffer + 2 * fft_result_size, fft_buffer, sizeof(float) * 2 * audio_max_fft_size);
        // signal_loop
        if (!clients.empty()) {
            if (b200_clients_execute(e, frame_num, pcm.data(), pwr.data(), valid.data())) die(b200_last_error());
            for (size_t i = 0; i < clients.size(); i++) {
                if (!valid[i]) continue;
                uint32_t hdr[4] = {0x41554449u /* 'AUDI' */, (uint32_t)frame_num, (uint32_t)i, (uint32_t)h};
                fwrite(hdr, sizeof(hdr), 1, stdout);
                fwrite(&pwr[i], sizeof(float), 1, stdout);
                fwrite(&pcm[i * h], sizeof(int32_t), h, stdout);
            }
        }
        // waterfall_loop: default client = last level, [0, waterfall_size) (src/websocket.cpp:195-198)
        if (frame_num % skip_num == 0) {
            size_t off = 0;
            for (int lv = 0; lv < downsample_levels - 1; lv++) off += fft_result_size >> lv;
            const uint32_t len = (uint32_t)std::min<long>(waterfall_size, fft_result_size >> (downsample_levels - 1));
            uint32_t hdr[4] = {0x57465241u /* 'WFRA' */, (uint32_t)frame_num, (uint32_t)(downsample_levels - 1), len};
            fwrite(hdr, sizeof(hdr), 1, stdout);
            fwrite(quantized + off, 1, len, stdout);
        }
        frame_num++;
    }
    buffer_read.wait();
    for (auto &b : input_buffers) fft.free(b);
    fflush(stdout);
    return 0;
}
