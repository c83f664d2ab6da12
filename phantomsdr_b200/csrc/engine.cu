// Host side of the B200 spectrum engine and its C ABI (include/phantomsdr_b200.h).
// Mirrors the reference's FFT backend life cycle (src/fft.h:33-63, src/fft_impl.cpp:63-183,
// src/fft_cuda.cu) and the AudioClient bookkeeping (src/signal.cpp:7-98,300-336) around the
// kernels in fft_fwd.cuh and clients.cuh. No CPU fallback exists: every compute entry point
// launches CUDA kernels or fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/phantomsdr_b200.h"
#include "../../include/phantomsdr_b200_debug.h"
#include "clients.cuh"
#include "clients_tail.cuh"
#include "fft_fwd.cuh"
#include "fft_tma.cuh"
#include "fft_stream.cuh"

using namespace b200;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

struct SubPlan {  // S = RA * RB points per sub-transform, T columns per CTA
    int S, RA, RB, T;
};

SubPlan sub_plan(int S, const char *tile_env) {
    int t1024 = 16;
    if (const char *v = getenv(tile_env)) {  // tuning aid: B200_TILE1 / B200_TILE2 = 8 | 16 for the 1024-point passes
        if (atoi(v) == 8) t1024 = 8;
        if (atoi(v) == 4) t1024 = 4;
    }
    switch (S) {
    case 256: return {256, 16, 16, 32};
    case 512: return {512, 16, 32, 16};
    case 1024: return {1024, 32, 32, t1024};
    }
    return {0, 0, 0, 0};
}

float2 *upload_f2(const std::vector<float2> &v) {
    float2 *d = nullptr;
    if (cudaMalloc(&d, sizeof(float2) * v.size()) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), sizeof(float2) * v.size(), cudaMemcpyHostToDevice);
    return d;
}

std::vector<float2> tw_sub(int RA, int RB) {  // [q][r] = exp(-2*pi*i*r*q/S)
    const int S = RA * RB;
    std::vector<float2> t((size_t)RA * RB);
    for (int q = 0; q < RA; q++)
        for (int r = 0; r < RB; r++) {
            const double a = -2.0 * M_PI * (double)((r * q) % S) / S;
            t[(size_t)q * RB + r] = make_float2((float)cos(a), (float)sin(a));
        }
    return t;
}

std::vector<float2> tw_pow(size_t count, size_t mult, size_t N) {  // [j] = exp(-2*pi*i*(j*mult mod N)/N)
    std::vector<float2> t(count);
    for (size_t j = 0; j < count; j++) {
        const double a = -2.0 * M_PI * (double)((j * mult) % N) / (double)N;
        t[j] = make_float2((float)cos(a), (float)sin(a));
    }
    return t;
}

}  // namespace

extern "C" int b200_set_option(struct b200_engine *e, int option, int value);

// ---- table-driven waterfall quantiser (see quantize_table in fft_fwd.cuh) ----------------------------------------
// op-for-op the reference arithmetic (src/fft_impl.cpp:14-44); the attribute keeps the host compiler from contracting
__attribute__((optimize("-ffp-contract=off"))) static int quant_exact_host(uint32_t bits, int off) {
    volatile float log_val = (float)((int)((bits >> 23) & 0xFF) - 128) + (float)off;
    uint32_t mb = (bits & ~(255u << 23)) + (127u << 23);
    float m;
    memcpy(&m, &mb, 4);
    volatile float t = -0.34484843f * m;
    t = t + 2.02466578f;
    t = t * m;
    t = t - 0.67487759f;
    volatile float v = log_val + t;
    v = v * 0.3010299956639812f;
    v = v * 20.f;
    v = v + 127.f;
    float vv = v;
    vv = vv > -128.f ? vv : -128.f;
    return (int)vv & 0xFF;
}
// 2048 cells indexed by bits >> 20: x = lo, y = base | (hi - lo) << 8
static void build_quant_table(int off, uint32_t *lo_out, uint32_t *hi_out, uint8_t *base_out) {
    const uint32_t scan = 64;  // the bands are <= 3 ulps wide (tools/quant_table.c); the bisection lands inside or next to them
    for (uint32_t c = 0; c < 2048; c++) {
        const uint32_t b0 = c << 20, b1 = b0 + (1u << 20);
        const int q0 = quant_exact_host(b0, off);
        uint32_t a = b0, b = b1;  // invariant: Q(a) == q0; b == b1 or Q(b) != q0
        if (quant_exact_host(b1 - 1, off) == q0) a = b1 - 1;
        while (b - a > 1) {
            const uint32_t mid = a + (b - a) / 2;
            if (quant_exact_host(mid, off) == q0) a = mid;
            else b = mid;
        }
        uint32_t lo = b, hi = b;
        const uint32_t s0 = b > b0 + scan ? b - scan : b0, s1 = b + scan < b1 ? b + scan : b1;
        for (uint32_t x = s0; x < s1; x++) {
            const bool same = quant_exact_host(x, off) == q0;
            if (!same && x < lo) lo = x;
            if (same && x + 1 > hi) hi = x + 1;
        }
        if (lo > hi) lo = hi;
        lo_out[c] = lo;
        hi_out[c] = hi;
        base_out[c] = (uint8_t)q0;
    }
}

struct b200_engine {
    int device = 0;
    size_t size = 0;
    int levels = 1;
    int size_log2 = 0;
    size_t additional = 0;
    bool planned = false;
    bool is_real = false;
    size_t M = 0;  // complex transform length
    int log2M = 0;
    int na = 1;        // M = na * Mb: radix-na split in front of na two-pass sub-transforms (M > 2^20)
    int log2Mb = 0;    // sub-transform length (== log2M when na == 1)
    float2 *d_pre = nullptr, *d_TLM = nullptr, *d_THM = nullptr;
    uint2 *d_qtab = nullptr;            // [3][2048] quantiser tables of levels 0..2 (opt_packed bit 1)
    unsigned *d_done = nullptr;         // [lanes][64] tiles of each frame stored by the fused pass-2 + pyramid kernel
    int opt_pyr_lag = 2;                // fused kernel: frames between a tile and the pyramid blocks that ride on it
    size_t R = 0;  // fft_result_size
    SubPlan sp1{}, sp2{};
    cudaStream_t stream = nullptr;      // the stream forward work is enqueued on
    cudaStream_t own_stream = nullptr;  // created by the engine
    cudaStream_t cstream = nullptr;     // client kernels when banks > 1 (overlaps the next forward batch)
    int banks = 1;                      // spectrum / pyramid banks (software pipeline depth)
    int cur_bank = 0;
    cudaEvent_t ev_fwd[4] = {};         // forward stream state at the time the bank's clients were enqueued
    cudaEvent_t ev_cli[4] = {};         // the bank's clients have finished
    bool cli_pending[4] = {};

    float *d_window = nullptr;
    float2 *d_Y = nullptr, *d_Z = nullptr, *d_spec = nullptr, *spec_bound = nullptr;
    float2 *d_spec_raw = nullptr;  // d_spec = d_spec_raw + 15: bin 1 (first bin of every IQ tile run) sits on a 128-byte line
    int8_t *d_quant = nullptr;
    float *d_ptop = nullptr;
    float *d_pscratch = nullptr;
    float2 *d_twA1 = nullptr, *d_twA2 = nullptr, *d_TL = nullptr, *d_TH = nullptr, *d_TLr = nullptr, *d_THr = nullptr;
    size_t spec_stride = 0, pyr_stride = 0, pyr_bytes = 0;
    int batch = 1;

    void *d_ring = nullptr;
    size_t nhops = 3;
    int in_format = B200_FMT_F32;
    size_t hop_samples = 0;  // scalar samples per hop (size for c2c IQ floats, size/2 for r2c)

    float *h_out = nullptr;
    int8_t *h_quant = nullptr;

    // load bookkeeping
    const void *last_a2 = nullptr;
    long head = -1;      // ring index of the newest hop
    long frame_hop0 = -1;
    int opt_reload_both = 0;
    int opt_mirror = 3;
    int opt_stage_mask = 7;
    int opt_fused_pyramid = -1;         // -1 auto: 0 with the TMA passes (2^20), 2 with the generic passes
    int opt_tma = 2;
    int opt_tail_pipe = 1;
    int opt_packed = 1;                 // bit0: packed-f32 (FMUL2/FADD2) waterfall quantiser
    int opt_pass1_order = 0;            // item order of the TMA pass 1 (see fft_pass1_tma_kernel)
    int opt_fwd_sms = 0;                // SMs the persistent forward kernels size their grids for (0 = all)
    int opt_p1_split = 16;              // CTAs per column tile of the TMA pass 1 (frames f == part mod split): 2048 work units of
                                        // four frames for the block scheduler (measured 3.16 vs 3.28 us/frame at two)
    int opt_lanes = 1;                  // forward lanes: sub-batches of a batch run on this many streams
    int opt_sub_frames = 64;            // frames per sub-batch (>= batch: one launch group per batch)
    cudaStream_t lane_stream[4] = {};
    cudaEvent_t ev_fork = nullptr, ev_lane[4] = {};
    bool tma_ok = false;
    int num_sms = 148;
    CUtensorMap ring_map{}, window_map{};
    // dataflow-scheduled forward group (fft_stream.cuh, B200_OPT_TMA 4)
    CUtensorMap spec_map{};             // the spectrum as rows of 16 bins from bin 1, 128-byte swizzle (quantiser items)
    bool spec_map_ok = false;
    StreamSync *d_ssync = nullptr;
    int *h_wait_err = nullptr;          // the device's mapped bounded-wait error word (fft_tma.cuh: g_wait_err); not owned
    unsigned *h_abort = nullptr;        // pinned mirror of StreamSync::abort, refreshed behind every launch
    float2 *d_winT = nullptr;
    unsigned *d_items = nullptr;
    int *d_nitems = nullptr;
    int items_frames = -1, items_grid = 0, items_lag1 = 0, items_lag2 = 0, items_max = 0, items_mask = 7;
    int opt_stream_grid = 0;            // CTAs of the stream kernel (0 = one per SM)
    int opt_stream_lag1 = 2;            // frame slots between pass 1 and pass 2 of a frame in the item order
    int opt_stream_lag2 = 4;            // ... between pass 1 and the quantiser
    int opt_stream_ring = 5;            // Y ring slots
    bool stream_used = false;

    int npeers = 0;
    float2 *peers[kMaxPeers] = {};
    unsigned peer_lo[kMaxPeers][2] = {}, peer_hi[kMaxPeers][2] = {};
    unsigned long long *d_flags = nullptr;  // 64 stream-ordered flags (exportable over CUDA IPC)
    int opt_peer_stores = 1;                // 1: FFT pass 2 stores into the peers itself; 0: b200_push_peers (copy engines)
    cudaEvent_t ev_push[4] = {}, ev_fwd_done = nullptr;
    bool push_pending[4] = {};
    cudaStream_t peer_stream[kMaxPeers] = {};  // b200_push_peers: one copy stream per peer
    cudaEvent_t ev_peer[kMaxPeers] = {};
    int *d_flag_err = nullptr;

    // clients
    bool have_clients = false;
    ClientArrays ca{};
    std::vector<ClientSlot> slots;
    std::vector<double> mids;
    bool slots_dirty = true;
    std::vector<int> order;
    int *d_order = nullptr;
    int tail_cpb = 32;
    size_t tail_smem = 0;
    // With banks > 1 the tails of batch k run on their own stream, concurrently with the forward group and the demodulation
    // of batch k + 1: they are latency-bound (serial recurrences) and occupy few SMs. Audio and PCM are double-buffered.
    cudaStream_t tstream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_demod[2] = {}, ev_tail[2] = {}, ev_join = nullptr, ev_fetch[2] = {};
    bool tail_pending[2] = {};
    int tail_buf = 0;                   // buffer the NEXT client batch uses
    int last_buf = 0;                   // buffer that holds the results of the last client batch
    bool use_tail2 = false;             // lane-per-client pipeline (clients_tail.cuh) with its own state layout
    Tail2State t2{};
    int *h_tail_err = nullptr;          // mapped host word behind t2.err
    int last_client_frames = 0;
    int demod_fchunk = 1;
    int opt_client_mask = 3;            // profiling aid: bit0 = demodulation kernels, bit1 = tail kernel
    int opt_demod_chunk = -1;           // frames per warp task of the frame-chunked demodulation (0 = sequential kernel only,
                                        // -1 = chosen per launch: see demod_chunk_for)
    int flag_waits = 0;                 // b200_enqueue_wait calls since stream_check last read the flag error word
    size_t tail2_dyn = 0;               // dynamic shared memory of a tail CTA (tail2_dynamic_smem)
    int opt_tail_smem_kb = 224;         // shared memory a tail CTA asks for: with its 3 KB of static memory exactly the SM's 227 KB,
                                        // so that no other CTA (not even a pyramid CTA with 1 KB) shares its schedulers
    int opt_r2c_split = 1;              // r2c: Hermitian split as its own streaming kernel + the c2c pyramid kernel (0: one kernel)
    int opt_demod_generic = 0;          // 1: never use the compile-time-size demodulation kernel (comparison aid)
    int demod_wpc = 0;                  // warps per CTA of client_demod_warp_kernel (0 = audio FFT too long: sequential kernel)
    int cstate = 0;                     // which copy of the overlap state the next client batch reads
    unsigned char *d_redo = nullptr;
    long long *d_prof = nullptr;

    // pipelined host-block streaming (b200_stream_prime / b200_submit_block / b200_wait_block)
    cudaStream_t copy_stream = nullptr;
    std::map<const char *, size_t> host_allocs;  // b200_malloc buffers (base -> bytes): what one host->device copy may span
    static constexpr int kMaxBlocks = 4;  // host blocks in flight (b200_submit_block); the hop ring decides how many of them
    cudaEvent_t ev_in[kMaxBlocks] = {}, ev_out[kMaxBlocks] = {}, ev_ring_free[kMaxBlocks] = {};
    bool blk_pending[kMaxBlocks] = {};
    int blk_depth = 2;                  // = min(kMaxBlocks, (nhops - 2) / batch), fixed by b200_stream_prime
    uint64_t blk_submitted = 0, blk_waited = 0;
    long stream_head = -1;  // ring index of the newest resident half

    uint64_t launches = 0;
    // waterfall cadence: pyramids are computed (and copied to the host) only for frames whose number is a
    // multiple of wf_skip, as the reference SENDS them (src/fft.cpp:33,102-104); 1 = every frame (FFT::execute's behaviour)
    int wf_skip = 1;
    int opt_pcm16 = 0;                  // PCM rows as int16 (N2: half the D2H bytes) instead of the encoder API's int32
    uint64_t fwd_frame = 0;             // number of the first frame of the next forward batch
    // waterfall slot gather (b200_waterfall_gather): grow-only device / pinned staging
    void *d_wf_desc = nullptr, *d_wf_out = nullptr, *h_wf_out = nullptr;
    size_t wf_desc_cap = 0, wf_out_cap = 0;

    size_t format_bytes() const {
        switch (in_format) {
        case B200_FMT_U8:
        case B200_FMT_S8: return 1;
        case B200_FMT_U16:
        case B200_FMT_S16: return 2;
        default: return 4;
        }
    }
    size_t hop_bytes_max() const { return hop_samples * 4; }
    float2 *spec_ptr() const {
        return spec_bound ? spec_bound : d_spec + (size_t)cur_bank * batch * spec_stride;
    }
    int8_t *quant_ptr() const { return d_quant + (size_t)cur_bank * batch * pyr_stride; }
    cudaStream_t client_stream() const { return banks > 1 ? cstream : stream; }
    bool tail_async() const { return use_tail2 && banks > 1 && tstream != nullptr; }
    cudaStream_t tail_stream() const { return tail_async() ? tstream : client_stream(); }
    // per-batch client buffers (audio before the tails, validity, PCM) of buffer b
    ClientArrays client_arrays(int b) const {
        ClientArrays c = ca;
        const size_t mc = ca.max_clients, h = ca.h, F = batch;
        c.audio_pre += (size_t)b * F * mc * h;
        c.valid_a += (size_t)b * F * mc;
        c.pcm += (size_t)b * F * mc * h;
        c.valid += (size_t)b * F * mc;
        return c;
    }
    // the two copies of the overlap state (frame-chunked demodulation reads one and writes the other)
    void state_ptrs(int which, float *&real_prev, float *&real_hi, float2 *&bb_hi, float2 *&bb_last, int *&hi_div) const {
        const size_t mc = ca.max_clients, h = ca.h;
        real_prev = ca.real_prev + (size_t)which * mc * h;
        real_hi = ca.real_hi + (size_t)which * mc * h;
        bb_hi = ca.bb_hi + (size_t)which * mc * h;
        bb_last = ca.bb_last + (size_t)which * mc;
        hi_div = ca.hi_diverged + (size_t)which * mc;
    }
};

namespace {

template <int RA, int RB, int T, bool RAW, bool REAL, bool PRE = false> int launch_pass1r(b200_engine *e, const FwdParams &p, int frames) {
    constexpr int threads = T * CMax<RA, RB>::v;
    constexpr int PAD = (T < 16) ? (16 - T) : 0;
    constexpr size_t smem = sizeof(float2) * RB * (RA * T + PAD);
    if (frames == 0) {  // preparation call from plan time
        CU(cudaFuncSetAttribute(fft_pass1_kernel<RA, RB, T, RAW, REAL, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
        return 0;
    }
    dim3 grid(p.N2 / T, frames);
    fft_pass1_kernel<RA, RB, T, RAW, REAL, PRE><<<grid, threads, smem, e->stream>>>(p);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
template <int RA, int RB, int T, bool RAW> int launch_pass1(b200_engine *e, const FwdParams &p, int frames) {
    if (frames == 0) {
        int rc = launch_pass1r<RA, RB, T, RAW, false>(e, p, 0);
        return rc ? rc : launch_pass1r<RA, RB, T, RAW, true>(e, p, 0);
    }
    return p.is_real ? launch_pass1r<RA, RB, T, RAW, true>(e, p, frames) : launch_pass1r<RA, RB, T, RAW, false>(e, p, frames);
}

template <int RA, int RB, int T, int FUSE> int launch_pass2(b200_engine *e, const FwdParams &p, int frames) {
    constexpr int threads = T * CMax<RA, RB>::v;
    constexpr size_t smem = sizeof(float2) * RB * (RA * T + 1);
    static_assert(FUSE != 1 || sizeof(float) * RA * RB * (T + 4) <= smem, "power tile must fit in the exchange buffer");
    if (frames == 0) {  // preparation call from plan time
        CU(cudaFuncSetAttribute(fft_pass2_kernel<RA, RB, T, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }
    dim3 grid(p.N1 / T, frames);
    fft_pass2_kernel<RA, RB, T, FUSE><<<grid, threads, smem, e->stream>>>(p);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <bool RAW> int dispatch_pass1_fmt(b200_engine *e, const FwdParams &p, int frames) {
    const SubPlan &sp = e->sp1;
    if (sp.S == 256) return launch_pass1<16, 16, 32, RAW>(e, p, frames);
    if (sp.S == 512) return launch_pass1<16, 32, 16, RAW>(e, p, frames);
    if (sp.S == 1024 && sp.T == 16) return launch_pass1<32, 32, 16, RAW>(e, p, frames);
    if (sp.S == 1024 && sp.T == 8) return launch_pass1<32, 32, 8, RAW>(e, p, frames);
    if (sp.S == 1024 && sp.T == 4) return launch_pass1<32, 32, 4, RAW>(e, p, frames);
    return fail(B200_ENOTSUP, "no pass-1 kernel for sub-transform %d", sp.S);
}
int dispatch_pass1(b200_engine *e, const FwdParams &p, int frames) {
    if (frames == 0) {
        int rc = dispatch_pass1_fmt<false>(e, p, 0);
        return rc ? rc : dispatch_pass1_fmt<true>(e, p, 0);
    }
    return p.in_format == FMT_F32 ? dispatch_pass1_fmt<false>(e, p, frames) : dispatch_pass1_fmt<true>(e, p, frames);
}
template <int FUSE> int dispatch_pass2_f(b200_engine *e, const FwdParams &p, int frames) {
    const SubPlan &sp = e->sp2;
    if (sp.S == 256) return launch_pass2<16, 16, 32, FUSE>(e, p, frames);
    if (sp.S == 512) return launch_pass2<16, 32, 16, FUSE>(e, p, frames);
    if (sp.S == 1024 && sp.T == 16) return launch_pass2<32, 32, 16, FUSE>(e, p, frames);
    if (sp.S == 1024 && sp.T == 8) return launch_pass2<32, 32, 8, FUSE>(e, p, frames);
    if (sp.S == 1024 && sp.T == 4) return launch_pass2<32, 32, 4, FUSE>(e, p, frames);
    return fail(B200_ENOTSUP, "no pass-2 kernel for sub-transform %d", sp.S);
}
int dispatch_pass2(b200_engine *e, const FwdParams &p, int frames, int fuse) {
    if (frames == 0) {
        int rc = dispatch_pass2_f<0>(e, p, 0);
        if (!rc) rc = dispatch_pass2_f<1>(e, p, 0);
        return rc ? rc : dispatch_pass2_f<2>(e, p, 0);
    }
    if (fuse == 1) return dispatch_pass2_f<1>(e, p, frames);
    if (fuse == 2) return dispatch_pass2_f<2>(e, p, frames);
    return dispatch_pass2_f<0>(e, p, frames);
}

// forward FFT + pyramid for `frames` consecutive frames starting at ring hop `hop0`
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D uint32 tensor {inner, rows} with row pitch = inner * 4 bytes, box {box_inner, 256}
int make_map_2d(CUtensorMap *map, void *base, uint64_t inner, uint64_t rows, uint32_t box_inner, bool swizzle128 = false,
                int elem_bytes = 4) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(B200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {inner * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {box_inner, 256};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

bool tma_path(const b200_engine *e) {
    return e->opt_tma && e->tma_ok && e->log2M == 20;  // (raw ADC formats included: pass 1 converts them)
}

int tma_prepare(b200_engine *e) {  // plan time: window map + kernel attributes
    e->tma_ok = false;
    if (e->log2M != 20) return 0;
    if (getenv("B200_NO_TMA")) return 0;
    if (!encode_tiled_fn()) return 0;  // no cuTensorMapEncodeTiled from this driver: the generic passes take over
    {
        // the word a bounded wait of the TMA kernels writes when it expires: one per device and process (mapped host memory
        // behind a __device__ pointer), shared by every engine on the device and never freed
        static int *host_word[64] = {};
        if (e->device >= 0 && e->device < 64) {
            if (!host_word[e->device]) {
                int *h = nullptr, *d = nullptr;
                CU(cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped));
                *h = 0;
                CU(cudaHostGetDevicePointer(&d, h, 0));
                CU(cudaMemcpyToSymbol(g_wait_err, &d, sizeof(d)));
                host_word[e->device] = h;
            }
            e->h_wait_err = host_word[e->device];
        }
    }
    CU(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device));
    if (!e->is_real) {
        int rc = make_map_2d(&e->window_map, e->d_window, kS, kS, kTmaT);
        if (rc) return rc;
    } else if (!e->d_winT) {
        // r2c pass 1 builds its Hann weights from (h cos, h sin)(2 pi row / 1024), h = 1/2 (the 1/N normalisation of r2c is
        // applied with the Hermitian split)
        std::vector<float2> wt(kS);
        for (int r = 0; r < kS; r++) {
            const double a = 2.0 * M_PI * (double)r / kS;
            wt[r] = make_float2((float)(0.5 * cos(a)), (float)(0.5 * sin(a)));
        }
        e->d_winT = upload_f2(wt);
        if (!e->d_winT) return fail(B200_ENOMEM, "window table allocation failed");
    }
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1C));
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1C));
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1C));
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1R));
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1R));
    CU(cudaFuncSetAttribute(fft_pass1_tma_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass1R));
    CU(cudaFuncSetAttribute(fft_pass2_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass2));
    CU(cudaFuncSetAttribute(fft_pass2_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass2));
    CU(cudaFuncSetAttribute(fft_pass2_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaSmem::kPass2));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    CU(cudaFuncSetAttribute(fft_pass2_tma3_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3Smem::kTotal));
    if (!e->is_real) {
        CU(cudaFuncSetAttribute(fwd_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamSmem::kTotal));
        CU(cudaFuncSetAttribute(fwd_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamSmem::kTotal));
        if (!e->d_ssync) {
            CU(cudaMalloc(&e->d_ssync, sizeof(StreamSync)));
            CU(cudaMemset(e->d_ssync, 0, sizeof(StreamSync)));
            CU(cudaHostAlloc(&e->h_abort, sizeof(unsigned), cudaHostAllocDefault));
            *e->h_abort = 0;
            // window table of pass 1: (h cos, h sin)(2 pi row / 1024), h = 1 / (2 size) (the 1/N normalisation rides on the window)
            std::vector<float2> wt(kS);
            const double h = 0.5 / (double)e->size;
            for (int r = 0; r < kS; r++) {
                const double a = 2.0 * M_PI * (double)r / kS;
                wt[r] = make_float2((float)(h * cos(a)), (float)(h * sin(a)));
            }
            e->d_winT = upload_f2(wt);
            if (!e->d_winT) return fail(B200_ENOMEM, "window table allocation failed");
        }
    }
    if (!e->d_done) {
        CU(cudaMalloc(&e->d_done, sizeof(unsigned) * 4 * 128));  // per-frame tile counters of the fused kernels: per lane 64 + 64
        CU(cudaMemset(e->d_done, 0, sizeof(unsigned) * 4 * 128));
    }
    e->tma_ok = true;
    return 0;
}

int tma_ring_map(b200_engine *e) {  // whenever the hop ring is (re)allocated
    if (!e->tma_ok || !e->d_ring) return 0;
    const uint64_t rows = (uint64_t)e->nhops * (kS / 2);
    switch (e->format_bytes()) {  // bytes per scalar sample; an element of the map row is a pair of them
    case 4: return make_map_2d(&e->ring_map, e->d_ring, 2 * kS, rows, 2 * kTmaT);
    case 2: return make_map_2d(&e->ring_map, e->d_ring, kS, rows, kTmaT);
    default: return make_map_2d(&e->ring_map, e->d_ring, kS, rows, kTmaT, false, 2);
    }
}

int launch_tma_pass1(b200_engine *e, const FwdParams &p, int frames) {
    // order 0: two CTAs share a column tile (even / odd frames); 1 / 2: equal chunks of the tile- / frame-major item list
    const int order = e->opt_pass1_order;
    const int nsplit = std::max(1, std::min(e->opt_p1_split, frames));
    const int grid = order == 0 ? (kS / kTmaT) * nsplit : std::min((kS / kTmaT) * frames, 2 * e->num_sms);
    const int eb = 2 * (int)e->format_bytes();  // bytes per element (IQ sample / pair of real samples) in the hop ring
#define B200_P1(REAL_, EB_, SMEM_) \
    fft_pass1_tma_kernel<REAL_, EB_><<<grid, kTmaThreads, SMEM_, e->stream>>>(p, e->ring_map, e->window_map, frames, order, nsplit)
    if (e->is_real) {
        if (eb == 8) B200_P1(true, 8, TmaSmem::kPass1R);
        else if (eb == 4) B200_P1(true, 4, TmaSmem::kPass1R);
        else B200_P1(true, 2, TmaSmem::kPass1R);
    } else {
        if (eb == 8) B200_P1(false, 8, TmaSmem::kPass1C);
        else if (eb == 4) B200_P1(false, 4, TmaSmem::kPass1C);
        else B200_P1(false, 2, TmaSmem::kPass1C);
    }
#undef B200_P1
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
int launch_tma_pass2(b200_engine *e, const FwdParams &p, int frames, int fuse, const PyrParams *pyr = nullptr, int lane = 0) {
    const int total = (kS / kTmaT) * frames;
    if (e->opt_tma >= 2 && fuse != 1) {  // three-stage variant: one CTA of two consumer groups per SM
        const int grid = std::min(total, e->opt_fwd_sms > 0 ? std::min(e->opt_fwd_sms, e->num_sms) : e->num_sms);
        const bool peers = p.npeers > 0;
        PyrParams none{};
        if (pyr) {  // spectrum + whole pyramid in one launch (FUSE 3): zero the per-frame tile counters first
            unsigned *done = e->d_done + 128 * lane;
            CU(cudaMemsetAsync(done, 0, sizeof(unsigned) * 64, e->stream));
            const int lag = std::max(1, e->opt_pyr_lag);
            if (!peers) fft_pass2_tma3_kernel<3, false><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, *pyr, done, lag);
            else fft_pass2_tma3_kernel<3, true><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, *pyr, done, lag);
        } else if (fuse == 2 && !peers)
            fft_pass2_tma3_kernel<2, false><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, none, nullptr, 0);
        else if (fuse == 2) fft_pass2_tma3_kernel<2, true><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, none, nullptr, 0);
        else if (!peers) fft_pass2_tma3_kernel<0, false><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, none, nullptr, 0);
        else fft_pass2_tma3_kernel<0, true><<<grid, kP3Threads, P3Smem::kTotal, e->stream>>>(p, frames, none, nullptr, 0);
        e->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    const int grid = std::min(total, 2 * e->num_sms);
    if (fuse == 1) fft_pass2_tma_kernel<1><<<grid, kTmaThreads, TmaSmem::kPass2, e->stream>>>(p, frames);
    else if (fuse == 2) fft_pass2_tma_kernel<2><<<grid, kTmaThreads, TmaSmem::kPass2, e->stream>>>(p, frames);
    else fft_pass2_tma_kernel<0><<<grid, kTmaThreads, TmaSmem::kPass2, e->stream>>>(p, frames);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <int NA> int launch_split_na(b200_engine *e, const FwdParams &p, int frames) {
    dim3 grid((unsigned)(((size_t)1 << e->log2Mb) / 256), frames);
    float2 *pre = const_cast<float2 *>(p.pre);
    if (e->is_real) radix_split_kernel<NA, true><<<grid, 256, 0, e->stream>>>(p, pre, e->d_TLM, e->d_THM, e->log2Mb);
    else radix_split_kernel<NA, false><<<grid, 256, 0, e->stream>>>(p, pre, e->d_TLM, e->d_THM, e->log2Mb);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
int launch_split(b200_engine *e, const FwdParams &p, int frames) {
    switch (e->na) {
    case 2: return launch_split_na<2>(e, p, frames);
    case 4: return launch_split_na<4>(e, p, frames);
    case 8: return launch_split_na<8>(e, p, frames);
    }
    return fail(B200_ENOTSUP, "no radix-%d split kernel", e->na);
}

int bank_acquire(b200_engine *e) {
    // the forward stream may only overwrite a bank once the clients that read it are done
    if (e->banks > 1 && e->cli_pending[e->cur_bank]) {
        CU(cudaStreamWaitEvent(e->stream, e->ev_cli[e->cur_bank], 0));
        e->cli_pending[e->cur_bank] = false;
    }
    if (e->push_pending[e->cur_bank]) {  // a copy-engine push to the peers may still be reading the bank
        CU(cudaStreamWaitEvent(e->stream, e->ev_push[e->cur_bank], 0));
        e->push_pending[e->cur_bank] = false;
    }
    return 0;
}

// ---- dataflow-scheduled forward group (fft_stream.cuh) ----------------------------------------------------------
bool stream_path(const b200_engine *e) {
    return e->opt_tma == 4 && e->tma_ok && e->spec_map_ok && e->d_ssync && e->log2M == 20 && e->na == 1 && !e->is_real &&
           e->in_format == B200_FMT_F32 && !e->spec_bound && e->levels <= 13 && (e->opt_packed & 1) &&
           e->opt_lanes <= 1 && e->num_sms >= 8;
}
// The global item list - frame slot s: the 128 pass-1 tiles of frame s, the 128 pass-2 tiles of frame s - lag1, the 128
// quantiser chunks of frame s - lag2 - dealt round-robin to the CTAs (items of frames outside the batch are left out
// BEFORE dealing, so the CTAs stay balanced through the prologue and the epilogue).
int stream_items(b200_engine *e, int frames, int grid) {
    const int lag1 = e->opt_stream_lag1, lag2 = std::max(e->opt_stream_lag2, lag1);
    const int mask = e->opt_stage_mask & 7;
    if (e->items_frames == frames && e->items_grid == grid && e->items_lag1 == lag1 && e->items_lag2 == lag2 && e->items_mask == mask)
        return 0;
    constexpr int NT = kS / kTmaT;
    std::vector<std::vector<unsigned>> lists(grid);
    size_t i = 0;
    for (int s = 0; s < frames + lag2; s++) {
        const int lag[3] = {0, lag1, lag2};
        for (int type = 0; type < 3; type++) {
            const int f = s - lag[type];
            if (f < 0 || f >= frames || !(mask & (1 << type))) continue;
            for (int t = 0; t < NT; t++) lists[i++ % grid].push_back(stream_item(type, f, t));
        }
    }
    size_t mx = 1;
    for (auto &l : lists) mx = std::max(mx, l.size());
    std::vector<unsigned> flat((size_t)grid * mx, 0u);
    std::vector<int> counts(grid);
    for (int c = 0; c < grid; c++) {
        counts[c] = (int)lists[c].size();
        std::copy(lists[c].begin(), lists[c].end(), flat.begin() + (size_t)c * mx);
    }
    // the previous table may still be read by a launch in flight
    CU(cudaStreamSynchronize(e->stream));
    if (e->d_items) cudaFree(e->d_items);
    if (e->d_nitems) cudaFree(e->d_nitems);
    e->d_items = nullptr;
    e->d_nitems = nullptr;
    CU(cudaMalloc(&e->d_items, sizeof(unsigned) * flat.size()));
    CU(cudaMalloc(&e->d_nitems, sizeof(int) * grid));
    CU(cudaMemcpy(e->d_items, flat.data(), sizeof(unsigned) * flat.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_nitems, counts.data(), sizeof(int) * grid, cudaMemcpyHostToDevice));
    e->items_frames = frames;
    e->items_grid = grid;
    e->items_lag1 = lag1;
    e->items_lag2 = lag2;
    e->items_mask = mask;
    e->items_max = (int)mx;
    return 0;
}
int launch_stream(b200_engine *e, const FwdParams &p, const PyrParams &q, int f0, int frames) {
    const int grid = std::max(1, std::min(e->opt_stream_grid > 0 ? e->opt_stream_grid : e->num_sms, e->num_sms));
    int rc = stream_items(e, frames, grid);
    if (rc) return rc;
    StreamParams sp{};
    sp.fp = p;
    sp.pyr = q;
    sp.items = e->d_items;
    sp.nitems = e->d_nitems;
    sp.max_items = e->items_max;
    sp.K = std::max(1, std::min(e->opt_stream_ring, e->batch));
    sp.nframes = frames;
    sp.winT = e->d_winT;
    sp.whalf = 0.5f / (float)e->size;
    sp.spec_rows_per_frame = (unsigned)(e->spec_stride / 16);
    sp.spec_row0 = (unsigned)(((size_t)e->cur_bank * e->batch + f0) * e->spec_stride / 16);
    sp.nodeps = (e->opt_stage_mask & 7) != 7;
    sp.sync = e->d_ssync;
    CU(cudaMemsetAsync(e->d_ssync, 0, sizeof(unsigned) * 128, e->stream));  // the counters; the abort flag is sticky
    if (p.npeers > 0)
        fwd_stream_kernel<true><<<grid, kP3Threads, StreamSmem::kTotal, e->stream>>>(sp, e->ring_map, e->spec_map);
    else
        fwd_stream_kernel<false><<<grid, kP3Threads, StreamSmem::kTotal, e->stream>>>(sp, e->ring_map, e->spec_map);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(e->h_abort, &e->d_ssync->abort, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
    e->stream_used = true;
    return 0;
}
// after a synchronisation point: did a bounded wait inside the stream kernel expire?
int stream_check(b200_engine *e) {
    if (e->h_wait_err && *e->h_wait_err)
        return fail(B200_ECUDA, "forward FFT kernels: a bounded wait expired (kind %d: 1 = TMA transfer, 2 = frame counter); results of "
                                "the batch are incomplete", *e->h_wait_err);
    if (e->stream_used && e->h_abort && *e->h_abort)
        return fail(B200_ECUDA, "forward stream kernel: a bounded wait expired (protocol timeout); results of the batch are incomplete");
    if (e->h_tail_err && *e->h_tail_err)
        return fail(B200_ECUDA, "client tail pipeline: a bounded wait expired (protocol timeout, code %d = 10000 + 100 * stage + barrier); "
                                "results of the batch are incomplete", *e->h_tail_err);
    if (e->flag_waits && e->d_flag_err) {  // stream-ordered peer flags were waited on since the last check
        int v = 0;
        if (cudaMemcpy(&v, e->d_flag_err, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && v)
            return fail(B200_ECUDA, "a peer flag wait timed out (b200_enqueue_wait): the peer's data did not arrive");
        e->flag_waits = 0;
    }
    return 0;
}

// frames [f0, f0 + frames) of the batch that starts at ring hop `hop0`, scratch slots of `lane`, on e->stream
int forward_range(b200_engine *e, long hop0, int f0, int frames, int lane) {
    const size_t slot0 = (size_t)lane * e->opt_sub_frames;  // first scratch frame slot of this lane
    FwdParams p{};
    p.ring = e->d_ring;
    p.hop_bytes = e->hop_samples * e->format_bytes();
    p.nhops = (int)e->nhops;
    p.hop0 = (int)((hop0 + f0) % (long)e->nhops);
    p.in_format = e->in_format;
    p.window = e->d_window;
    p.Y = e->d_Y + slot0 * e->M;
    p.na = e->na;
    p.pre = e->na > 1 ? e->d_pre + slot0 * e->M : nullptr;
    p.log2M = e->log2Mb;  // what passes 1 and 2 transform (the sub-transform when na > 1)
    p.N1 = e->sp1.S;
    p.N2 = e->sp2.S;
    p.is_real = e->is_real ? 1 : 0;
    p.twA1 = e->d_twA1;
    p.twA2 = e->d_twA2;
    p.TL = e->d_TL;
    p.TH = e->d_TH;
    p.TLr = e->d_TLr;
    p.THr = e->d_THr;
    p.winT = e->d_winT;
    float2 *spec = e->spec_ptr() + (size_t)f0 * e->spec_stride;
    if (!e->is_real) {
        p.shift = e->na > 1 ? 0 : 1;  // the display-aligned row order only exists in the plain two-pass transform
        p.out = spec;
        p.out_stride = e->spec_stride;
        p.scale = 1.0f / (float)e->size;
        p.additional = (int)e->additional;
        p.npeers = e->opt_peer_stores ? e->npeers : 0;
        for (int i = 0; i < p.npeers; i++) {
            // peer buffers are addressed like the local bank: same frame stride, same bank offset
            p.peers[i] = e->peers[i] + ((size_t)e->cur_bank * e->batch + f0) * e->spec_stride;
            for (int j = 0; j < 2; j++) {
                p.peer_lo[i][j] = e->peer_lo[i][j];
                p.peer_hi[i][j] = e->peer_hi[i][j];
            }
        }
    } else {
        p.shift = 0;
        p.out = e->d_Z + slot0 * e->M;
        p.out_stride = e->M;
        p.scale = 1.0f;
        p.additional = 0;
        p.npeers = 0;
    }
    // 0 none (pyramid kernel re-reads the spectrum), 1 full epilogue in pass 2, 2 |X|^2 plane from pass 2
    // waterfall cadence: frames (fwd_frame + f0 + i) % wf_skip == 0 of this range get a pyramid
    const int wf_skip = std::max(1, e->wf_skip);
    const int wf_first = (int)((wf_skip - (e->fwd_frame + (uint64_t)f0) % wf_skip) % wf_skip);
    const int wf_count = wf_first < frames ? (frames - wf_first + wf_skip - 1) / wf_skip : 0;
    const int fuse = (e->is_real || e->na > 1 || wf_skip > 1) ? 0 : (e->opt_fused_pyramid >= 0 ? e->opt_fused_pyramid : (tma_path(e) ? 0 : 2));
    int base_level = 0;
    while ((1 << base_level) < e->sp2.T) base_level++;
    int8_t *quant = e->quant_ptr() + (size_t)f0 * e->pyr_stride;
    float *pscratch = e->d_pscratch + slot0 * e->M;
    p.quant = quant;
    p.pyr_stride = e->pyr_stride;
    p.pscratch = pscratch;
    p.levels = e->levels;
    p.size_log2 = e->size_log2;
    int rc = 0;
    const bool tma = tma_path(e);
    if (tma) {  // pass-2 TMA kernels use T = kTmaT rows per tile
        base_level = 0;
        while ((1 << base_level) < kTmaT) base_level++;
    }
    PyrParams q{};
    q.spec = spec;
    q.spec_stride = e->spec_stride;
    q.Z = e->d_Z + slot0 * e->M;
    q.quant = quant;
    q.pyr_stride = e->pyr_stride;
    q.ptop = e->d_ptop + slot0 * std::max<size_t>(1, e->R / 1024);
    int log2R = 0;
    while (((size_t)1 << log2R) < e->R) log2R++;
    q.log2R = log2R;
    q.levels = e->levels;
    q.size_log2 = e->size_log2;
    q.scale = 1.0f / (float)e->size;
    q.TLr = e->d_TLr;
    q.THr = e->d_THr;
    q.pscratch = pscratch;
    q.base_level = fuse == 1 ? base_level : 0;
    q.ntiles = fuse == 2 ? e->sp1.S : (tma ? kS / kTmaT : e->sp1.S / e->sp2.T);
    q.N2 = e->sp2.S;
    q.npeers = (e->is_real && e->opt_peer_stores) ? e->npeers : 0;  // (copy-engine mode: b200_push_peers moves the bank)
    q.qtab = (e->opt_packed & 2) ? e->d_qtab : nullptr;
    q.frame0 = 0;
    q.frame_step = 1;
    q.wf_first = wf_first;
    q.wf_skip = wf_skip;
    for (int i = 0; i < q.npeers; i++) {
        // peer buffers are addressed like the local bank: same frame stride, same bank offset (as the c2c path above)
        q.peers[i] = e->peers[i] + ((size_t)e->cur_bank * e->batch + f0) * e->spec_stride;
        for (int j = 0; j < 2; j++) {
            q.peer_lo[i][j] = e->peer_lo[i][j];
            q.peer_hi[i][j] = e->peer_hi[i][j];
        }
    }
    // opt_tma 3: pass 2 of the TMA path also produces the whole pyramid (FUSE 3), unless a stage is masked out for profiling
    const bool fused = tma && e->opt_tma == 3 && fuse == 0 && wf_skip == 1 && !e->is_real && e->na == 1 && (e->opt_packed & 1) &&
                       (e->opt_stage_mask & 6) == 6 && frames <= 64 && e->levels > 0 &&
                       e->opt_lanes <= 1;  // its CTAs wait on one another: never two such grids competing for the SMs
    if (e->na > 1) {
        if (e->opt_stage_mask & 1) {
            rc = launch_split(e, p, frames);
            if (!rc) rc = launch_pass1r<32, 32, 16, false, true, true>(e, p, frames * e->na);
        }
        if (rc) return rc;
        if (e->opt_stage_mask & 2) rc = dispatch_pass2(e, p, frames * e->na, 0);
        if (rc) return rc;
    } else if (stream_path(e) && frames <= 64 && wf_skip == 1) {
        // all three stages in one persistent, dataflow-scheduled launch (fft_stream.cuh): Y and the spectrum a quantiser
        // item reads never leave L2. The 1/N normalisation rides on pass 1's window table.
        return launch_stream(e, p, q, f0, frames);
    } else {
        if (e->opt_stage_mask & 1) rc = tma ? launch_tma_pass1(e, p, frames) : dispatch_pass1(e, p, frames);
        if (rc) return rc;
        if (e->opt_stage_mask & 2) rc = tma ? launch_tma_pass2(e, p, frames, fuse, fused ? &q : nullptr, lane) : dispatch_pass2(e, p, frames, fuse);
        if (rc) return rc;
    }
    if (!(e->opt_stage_mask & 4)) return 0;
    if (fused) {  // only very deep pyramids have levels left (sums in ptop)
        if (e->levels > 13) {
            pyramid_tail_kernel<<<frames, 512, 0, e->stream>>>(q, 0, 13);
            e->launches++;
            CU(cudaGetLastError());
        }
        return 0;
    }
    if (e->levels > q.base_level) {
        // 16 entries per thread for the full-resolution inputs, 4 for the small per-tile sums of mode 1
        const int per = (fuse == 1) ? 4 : 16;
        // r2c: the kernel also does the Hermitian split, so it covers every frame and drops the pyramid of the frames
        // between two sends itself; c2c: only the send frames are launched
        int pyr_frames = frames;
        const bool split_first = e->is_real && e->opt_r2c_split && (e->R % 1024) == 0;
        if (split_first) {  // r2c: the Hermitian split of every frame as a streaming kernel of its own
            dim3 sgrid((unsigned)(e->R / 1024), frames);
            r2c_split_kernel<<<sgrid, 256, 0, e->stream>>>(q);
            e->launches++;
            CU(cudaGetLastError());
            q.natural = 1;
        }
        if ((!e->is_real || split_first) && wf_skip > 1) {
            q.frame0 = wf_first;
            q.frame_step = wf_skip;
            pyr_frames = wf_count;
        }
        if (pyr_frames > 0) {
            dim3 grid((unsigned)((e->R >> q.base_level) / (256 * per)), pyr_frames);
            const bool pk = e->opt_packed & 1;
            if (e->is_real && !split_first) {
                if (pk) pyramid_kernel<PYR_R2C, 16, true><<<grid, 256, 0, e->stream>>>(q);
                else pyramid_kernel<PYR_R2C, 16, false><<<grid, 256, 0, e->stream>>>(q);
            } else if (fuse == 1) {
                if (pk) pyramid_kernel<PYR_SCRATCH, 4, true><<<grid, 256, 0, e->stream>>>(q);
                else pyramid_kernel<PYR_SCRATCH, 4, false><<<grid, 256, 0, e->stream>>>(q);
            } else if (fuse == 2) {
                if (pk && q.qtab) pyramid_kernel<PYR_POWER, 16, true, true><<<grid, 256, 0, e->stream>>>(q);
                else if (pk) pyramid_kernel<PYR_POWER, 16, true><<<grid, 256, 0, e->stream>>>(q);
                else pyramid_kernel<PYR_POWER, 16, false><<<grid, 256, 0, e->stream>>>(q);
            } else {
                if (pk && q.qtab) pyramid_kernel<PYR_SPEC, 16, true, true><<<grid, 256, 0, e->stream>>>(q);
                else if (pk) pyramid_kernel<PYR_SPEC, 16, true><<<grid, 256, 0, e->stream>>>(q);
                else pyramid_kernel<PYR_SPEC, 16, false><<<grid, 256, 0, e->stream>>>(q);
            }
            e->launches++;
            CU(cudaGetLastError());
        }
        const int levels_done = (per == 16 ? 4 : 2) + 9;
        if (e->levels - q.base_level > levels_done && wf_count > 0) {
            if (wf_skip > 1) {  // (r2c included: only the send frames left sums behind)
                q.frame0 = wf_first;
                q.frame_step = wf_skip;
            }
            pyramid_tail_kernel<<<wf_skip > 1 ? wf_count : frames, 512, 0, e->stream>>>(q, q.base_level, levels_done);
            e->launches++;
            CU(cudaGetLastError());
        }
    }
    return 0;
}

// forward FFT + pyramid for `frames` consecutive frames starting at ring hop `hop0`. With lanes > 1 the batch is cut
// into sub-batches of opt_sub_frames frames that alternate over the lane streams: every sub-batch's intermediates
// (Y, |X|^2) live in its lane's small scratch slots, which are rewritten while still resident in L2 (no DRAM round
// trip), and the lanes fill each other's launch gaps and partial waves.
int run_forward_batch(b200_engine *e, long hop0, int frames);
// the frame counter that drives the waterfall cadence advances with every forward batch
int run_forward(b200_engine *e, long hop0, int frames) {
    int rc = run_forward_batch(e, hop0, frames);
    if (!rc) e->fwd_frame += (uint64_t)frames;
    return rc;
}
int run_forward_batch(b200_engine *e, long hop0, int frames) {
    {
        int rc0 = bank_acquire(e);
        if (rc0) return rc0;
    }
    int sub = std::max(1, std::min(e->opt_sub_frames, e->batch));
    int lanes = std::max(1, std::min({e->opt_lanes, 4, e->batch / sub}));
    if (sub >= frames) lanes = 1;
    if (lanes == 1 && sub >= frames) return forward_range(e, hop0, 0, frames, 0);
    if (lanes == 1) {  // sub-batches back to back on the one stream (scratch slot 0 reused)
        for (int f0 = 0; f0 < frames; f0 += sub) {
            int rc = forward_range(e, hop0, f0, std::min(sub, frames - f0), 0);
            if (rc) return rc;
        }
        return 0;
    }
    for (int l = 0; l < lanes; l++)
        if (!e->lane_stream[l]) {
            CU(cudaStreamCreateWithFlags(&e->lane_stream[l], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&e->ev_lane[l], cudaEventDisableTiming));
        }
    if (!e->ev_fork) CU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CU(cudaEventRecord(e->ev_fork, e->stream));
    cudaStream_t main_stream = e->stream;
    int rc = 0, used = 0;
    for (int f0 = 0, i = 0; f0 < frames && !rc; f0 += sub, i++) {
        const int l = i % lanes;
        if (i < lanes) {
            CU(cudaStreamWaitEvent(e->lane_stream[l], e->ev_fork, 0));
            used = i + 1;
        }
        e->stream = e->lane_stream[l];
        rc = forward_range(e, hop0, f0, std::min(sub, frames - f0), l);
        e->stream = main_stream;
    }
    for (int l = 0; l < used; l++) {
        CU(cudaEventRecord(e->ev_lane[l], e->lane_stream[l]));
        CU(cudaStreamWaitEvent(main_stream, e->ev_lane[l], 0));
    }
    return rc;
}

int plan_common(b200_engine *e, bool is_real) {
    if (e->planned) return fail(B200_ESTATE, "engine already planned");
    CU(cudaSetDevice(e->device));
    e->is_real = is_real;
    e->M = is_real ? e->size / 2 : e->size;
    e->R = is_real ? e->size / 2 : e->size;
    e->log2M = 0;
    while (((size_t)1 << e->log2M) < e->M) e->log2M++;
    int s1 = 0, s2 = 0;
    switch (e->log2M) {
    case 16: s1 = 256; s2 = 256; break;
    case 17: s1 = 512; s2 = 256; break;
    case 18: s1 = 512; s2 = 512; break;
    case 19: s1 = 1024; s2 = 512; break;
    case 20: s1 = 1024; s2 = 1024; break;
    case 21: case 22: case 23: s1 = 1024; s2 = 1024; break;  // radix-2/4/8 split + 2^20-point sub-transforms
    default:
        return fail(B200_ENOTSUP, "fft size %zu (%s) not supported: complex transform length must be 2^16..2^23", e->size,
                    is_real ? "r2c" : "c2c");
    }
    e->log2Mb = std::min(e->log2M, 20);
    e->na = 1 << (e->log2M - e->log2Mb);
    const size_t Mb = (size_t)1 << e->log2Mb;
    e->sp1 = sub_plan(s1, "B200_TILE1");
    e->sp2 = sub_plan(s2, "B200_TILE2");
    if (e->levels < 1) return fail(B200_EINVAL, "downsample_levels must be >= 1");
    if ((e->R >> (e->levels - 1)) < 1) return fail(B200_EINVAL, "too many downsample levels for %zu bins", e->R);
    e->hop_samples = is_real ? e->size / 2 : e->size;  // scalar samples per hop
    e->spec_stride = is_real ? (e->size / 2 + 16) : (e->size + ((e->additional + 15) & ~(size_t)15) + 16);
    e->pyr_bytes = 0;
    for (int i = 0; i < e->levels; i++) e->pyr_bytes += e->R >> i;
    e->pyr_stride = (e->pyr_bytes + 255) & ~(size_t)255;

    // twiddles (double precision on the host, rounded once)
    e->d_twA1 = upload_f2(tw_sub(e->sp1.RA, e->sp1.RB));
    e->d_twA2 = upload_f2(tw_sub(e->sp2.RA, e->sp2.RB));
    e->d_TL = upload_f2(tw_pow(1024, 1, Mb));
    e->d_TH = upload_f2(tw_pow(std::max<size_t>(1, Mb / 1024), 1024, Mb));
    if (e->na > 1) {
        e->d_TLM = upload_f2(tw_pow(1024, 1, e->M));
        e->d_THM = upload_f2(tw_pow(e->M / 1024, 1024, e->M));
        if (!e->d_TLM || !e->d_THM) return fail(B200_ENOMEM, "twiddle allocation failed");
    }
    if (is_real) {
        e->d_TLr = upload_f2(tw_pow(1024, 1, e->size));
        e->d_THr = upload_f2(tw_pow(std::max<size_t>(1, e->M / 1024), 1024, e->size));
    }
    if (!e->d_twA1 || !e->d_twA2 || !e->d_TL || !e->d_TH) return fail(B200_ENOMEM, "twiddle allocation failed");

    // pinned host mirrors (FFT::get_output_buffer / get_quantized_buffer must be CPU-readable)
    const size_t out_floats = is_real ? e->size + 2 : 2 * (e->size + e->additional);
    CU(cudaHostAlloc(&e->h_out, sizeof(float) * out_floats, cudaHostAllocDefault));
    memset(e->h_out, 0, sizeof(float) * out_floats);
    CU(cudaHostAlloc(&e->h_quant, e->pyr_stride, cudaHostAllocDefault));
    memset(e->h_quant, 0, e->pyr_stride);
    {   // opt the chosen kernels into their shared-memory footprint on this device
        FwdParams none{};
        int rc = dispatch_pass1(e, none, 0);
        if (rc) return rc;
        rc = dispatch_pass2(e, none, 0, false);
        if (rc) return rc;
        if (e->na > 1) rc = launch_pass1r<32, 32, 16, false, true, true>(e, none, 0);
        if (rc) return rc;
    }
    {
        int rc = tma_prepare(e);
        if (rc) return rc;
    }
    e->planned = true;
    if (const char *opts = getenv("B200_OPTS")) {  // tuning aid: "option=value,option=value" applied after planning
        int k = 0, v = 0, used = 0;
        while (sscanf(opts, "%d=%d%n", &k, &v, &used) == 2) {
            b200_debug_option(e, k, v);
            opts += used;
            if (*opts == ',') opts++;
        }
    }
    return 0;
}

int alloc_batch(b200_engine *e, int frames) {
    if (e->d_Y) cudaFree(e->d_Y);
    if (e->d_Z) cudaFree(e->d_Z);
    if (e->d_spec_raw) cudaFree(e->d_spec_raw);
    e->d_spec_raw = nullptr;
    if (e->d_quant) cudaFree(e->d_quant);
    if (e->d_ptop) cudaFree(e->d_ptop);
    if (e->d_pscratch) cudaFree(e->d_pscratch);
    e->d_pscratch = nullptr;
    e->d_Y = e->d_Z = e->d_spec = nullptr;
    e->d_quant = nullptr;
    e->d_ptop = nullptr;
    CU(cudaMalloc(&e->d_Y, sizeof(float2) * e->M * frames));
    if (e->d_pre) cudaFree(e->d_pre);
    e->d_pre = nullptr;
    if (e->na > 1) CU(cudaMalloc(&e->d_pre, sizeof(float2) * e->M * frames));
    if (e->is_real) CU(cudaMalloc(&e->d_Z, sizeof(float2) * e->M * frames));
    // IQ bins are stored in runs k = 8m+1 .. 8m+8 (display shift, fft_impl.cpp:148-160): offsetting the buffer by 15
    // elements makes every run start on a 64-byte boundary, so pass 2 writes whole sectors
    CU(cudaMalloc(&e->d_spec_raw, sizeof(float2) * (e->spec_stride * frames * e->banks + 16)));
    CU(cudaMemset(e->d_spec_raw, 0, sizeof(float2) * (e->spec_stride * frames * e->banks + 16)));
    e->d_spec = e->d_spec_raw + (e->is_real ? 16 : 15);
    e->spec_map_ok = false;
    if (e->tma_ok && !e->is_real && e->spec_stride % 16 == 0) {
        // quantiser items of the stream kernel read the spectrum as rows of 16 bins starting at bin 1 (a 128-byte line)
        const uint64_t rows = (uint64_t)e->spec_stride * frames * e->banks / 16;
        if (make_map_2d(&e->spec_map, e->d_spec + 1, 32, rows, 32, true) == 0) e->spec_map_ok = true;
    }
    CU(cudaMalloc(&e->d_quant, e->pyr_stride * frames * e->banks));
    CU(cudaMemset(e->d_quant, 0, e->pyr_stride * frames * e->banks));
    e->cur_bank = 0;
    for (int b = 0; b < 4; b++) e->cli_pending[b] = e->push_pending[b] = false;
    CU(cudaMalloc(&e->d_ptop, sizeof(float) * std::max<size_t>(1, e->R / 1024) * frames));
    CU(cudaMalloc(&e->d_pscratch, sizeof(float) * e->M * frames));  // |X|^2 of every bin (mode 2) or per-tile sums (mode 1)
    e->batch = frames;
    return 0;
}

int alloc_ring(b200_engine *e, size_t nhops) {
    if (e->d_ring) cudaFree(e->d_ring);
    e->d_ring = nullptr;
    CU(cudaMalloc(&e->d_ring, e->hop_bytes_max() * nhops));
    CU(cudaMemset(e->d_ring, 0, e->hop_bytes_max() * nhops));
    e->nhops = nhops;
    e->head = -1;
    e->last_a2 = nullptr;
    e->frame_hop0 = -1;
    return tma_ring_map(e);
}

int load_common(b200_engine *e, const void *a1, const void *a2) {
    if (!e->planned) return fail(B200_ESTATE, "load before plan");
    if (!a1 || !a2) return fail(B200_EINVAL, "null input half");
    CU(cudaSetDevice(e->device));
    const size_t hb = e->hop_samples * e->format_bytes();
    char *ring = reinterpret_cast<char *>(e->d_ring);
    long hopA, hopB;
    if (e->head >= 0 && a1 == e->last_a2 && !e->opt_reload_both) {
        hopA = e->head;
        hopB = (e->head + 1) % (long)e->nhops;
    } else {
        hopA = (e->head + 1) % (long)e->nhops;
        hopB = (hopA + 1) % (long)e->nhops;
        CU(cudaMemcpyAsync(ring + (size_t)hopA * hb, a1, hb, cudaMemcpyHostToDevice, e->stream));
    }
    CU(cudaMemcpyAsync(ring + (size_t)hopB * hb, a2, hb, cudaMemcpyHostToDevice, e->stream));
    e->head = hopB;
    e->last_a2 = a2;
    e->frame_hop0 = hopA;
    return 0;
}

std::vector<int> factorize(int n) {
    std::vector<int> r;
    while (n % 4 == 0) { r.push_back(4); n /= 4; }
    while (n % 2 == 0) { r.push_back(2); n /= 2; }
    for (int p = 3; n > 1; p += 2)
        while (n % p == 0) { r.push_back(p); n /= p; }
    return r;
}

constexpr int kDemodThreads = 256;

// Frames per warp task. A task of c frames transforms c + 1 (the predecessor is recomputed), and the GPU runs the tasks in
// waves of (SMs x resident warps): cost ~ ceil(clients * ceil(F / c) / slots) * (c + 1). At 1024 clients x 64 frames that
// picks 7 (10 240 tasks = 2.9 waves; measured 2.9 us/frame against 3.2 for 8, whose 8 192 tasks leave the third wave 31 % full).
static int demod_chunk_for(const b200_engine *e, int nactive, int nframes) {
    if (e->opt_demod_chunk >= 0) return e->opt_demod_chunk;
    const long slots = (long)e->num_sms * 3 * std::max(1, e->demod_wpc);  // three CTAs per SM (registers, shared memory)
    int best = 8;
    long best_cost = -1;
    for (int c = 4; c <= 16; c++) {
        const long tasks = (long)nactive * ((nframes + c - 1) / c);
        const long cost = ((tasks + slots - 1) / slots) * (c + 1);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

int launch_demod(b200_engine *e, const ClientArrays &ca_in, const ClientLaunch &cl_in) {
    const size_t smem = sizeof(float2) * 2 * e->ca.n * e->demod_fchunk;
    const size_t per_warp = sizeof(float2) * (2 * (size_t)e->ca.n + e->ca.h);
    if (cl_in.nactive == 0) {  // preparation call from clients_create
        CU(cudaFuncSetAttribute(client_demod_kernel<kDemodThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        e->demod_wpc = (int)std::min<size_t>(kDemodWarps, (200 * 1024) / per_warp);
        if (e->demod_wpc >= 1)
        {
            CU(cudaFuncSetAttribute(client_demod_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * e->demod_wpc)));
            CU(cudaFuncSetAttribute(client_demod_warp_kernel<360>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * e->demod_wpc)));
        }
        return 0;
    }
    cudaStream_t cs = e->client_stream();
    ClientArrays ca = ca_in;
    ClientLaunch cl = cl_in;
    const int chunk_frames = demod_chunk_for(e, cl_in.nactive, cl_in.nframes);
    if (chunk_frames > 0 && e->demod_wpc >= 1) {
        // frame-chunked: read state copy `cstate`, write the other one; then replay the (rare) flagged clients sequentially
        float *rp, *rh;
        float2 *bh, *bl;
        int *hd;
        e->state_ptrs(e->cstate, rp, rh, bh, bl, hd);
        cl.sin_real_prev = rp;
        cl.sin_real_hi = rh;
        cl.sin_bb_hi = bh;
        cl.sin_bb_last = bl;
        cl.sin_hi_diverged = hd;
        e->state_ptrs(e->cstate ^ 1, ca.real_prev, ca.real_hi, ca.bb_hi, ca.bb_last, ca.hi_diverged);
        cl.redo = e->d_redo;
        cl.chunk = std::max(1, std::min(chunk_frames, cl.nframes));
        cl.nchunks = (cl.nframes + cl.chunk - 1) / cl.chunk;
        cl.redo_only = 0;
        CU(cudaMemsetAsync(e->d_redo, 0, (size_t)e->ca.max_clients, cs));
        const int tasks = cl.nactive * cl.nchunks;
        const dim3 grid((tasks + e->demod_wpc - 1) / e->demod_wpc), block(32 * e->demod_wpc);
        // 360 = the reference's audio FFT size at 12 kHz audio and the headline frame rate: stages unrolled at compile time
        if (e->ca.n == 360 && !e->opt_demod_generic) client_demod_warp_kernel<360><<<grid, block, per_warp * e->demod_wpc, cs>>>(ca, cl);
        else client_demod_warp_kernel<0><<<grid, block, per_warp * e->demod_wpc, cs>>>(ca, cl);
        e->launches++;
        CU(cudaGetLastError());
        cl.redo_only = 1;
        client_demod_kernel<kDemodThreads><<<cl.nactive, kDemodThreads, smem, cs>>>(ca, cl);
        e->launches++;
        CU(cudaGetLastError());
        e->cstate ^= 1;
        return 0;
    }
    // sequential kernel only: one copy of the state, in place
    e->state_ptrs(e->cstate, ca.real_prev, ca.real_hi, ca.bb_hi, ca.bb_last, ca.hi_diverged);
    cl.redo_only = 0;
    client_demod_kernel<kDemodThreads><<<cl.nactive, kDemodThreads, smem, cs>>>(ca, cl);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <int KB> int launch_tail_kb(b200_engine *e, const ClientArrays &ca, const ClientLaunch &cl) {
    if (cl.nactive == 0) {  // preparation call from clients_create
        CU(cudaFuncSetAttribute(client_tail_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tail_smem));
        return 0;
    }
    const int blocks = (cl.nactive + cl.cpb - 1) / cl.cpb;
    client_tail_kernel<KB><<<blocks, kTailThreads, e->tail_smem, e->tail_stream()>>>(ca, cl);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
// Dynamic shared memory of a tail CTA: what the pipeline needs, raised to the requested size (default: the whole SM) but
// never beyond what the device grants a block once the kernel's static shared memory is taken off.
static size_t tail2_dynamic_smem(const b200_engine *e) {
    size_t want = std::max(tail2_smem(e->ca.D), (size_t)e->opt_tail_smem_kb * 1024);
    cudaFuncAttributes fa{};
    int optin = 0;
    if (cudaFuncGetAttributes(&fa, client_tail2_kernel<1>) == cudaSuccess &&
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device) == cudaSuccess && optin > 0) {
        const size_t room = (size_t)optin > fa.sharedSizeBytes ? (size_t)optin - fa.sharedSizeBytes : 0;
        want = std::min(want, room);
    }
    return want;
}

int launch_tail2(b200_engine *e, const ClientArrays &ca, const ClientLaunch &cl) {
    // One group of 32 clients per CTA, and the CTA asks for ALL the shared memory of its SM although it needs 204 KB:
    // every stage of the pipeline is a single warp on the critical path, and warps of other CTAs competing for the SM's
    // issue slots slow the whole chain. Measured at 1024 clients beside the forward kernels (cycles per frame of a
    // stage-profiled CTA): alone 16.4 K; beside the forward kernels 22.3 K at 204 KB (pyramid CTAs move in), 19.3 K at
    // 224 KB. Two groups per SM (two CTAs, or one CTA of twice the warps) cost every stage 40 %.
    if (cl.nactive == 0) {  // preparation call from clients_create
        e->tail2_dyn = tail2_dynamic_smem(e);
        if (e->tail2_dyn < tail2_smem(e->ca.D)) return fail(B200_ENOTSUP, "tail pipeline needs %zu bytes of shared memory", tail2_smem(e->ca.D));
        CU(cudaFuncSetAttribute(client_tail2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tail2_dyn));
        return 0;
    }
    const size_t smem = e->tail2_dyn;
    const int groups = (e->ca.max_clients + 31) / 32;
    client_tail2_kernel<1><<<groups, kT2Threads, smem, e->tail_stream()>>>(ca, cl, e->t2, groups);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
int launch_tail(b200_engine *e, const ClientArrays &ca, const ClientLaunch &cl) {
    if (e->use_tail2) return launch_tail2(e, ca, cl);
    const int kb = (e->ca.h + 31) / 32;
    if (kb <= 2) return launch_tail_kb<2>(e, ca, cl);
    if (kb <= 4) return launch_tail_kb<4>(e, ca, cl);
    if (kb <= 6) return launch_tail_kb<6>(e, ca, cl);
    if (kb <= 9) return launch_tail_kb<9>(e, ca, cl);
    if (kb <= 12) return launch_tail_kb<12>(e, ca, cl);
    return launch_tail_kb<0>(e, ca, cl);
}

int run_clients(b200_engine *e, uint64_t frame_num, int nframes) {
    if (!e->have_clients) return fail(B200_ESTATE, "clients not created");
    if (!e->planned) return fail(B200_ESTATE, "engine not planned");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    CU(cudaSetDevice(e->device));
    if (e->slots_dirty) {
        e->order.clear();
        for (int i = 0; i < (int)e->slots.size(); i++)
            if (e->slots[i].flags & CF_ACTIVE) e->order.push_back(i);
        // std::multimap<std::pair<int,int>, ...> iteration order (src/spectrumserver.h:164-165)
        std::stable_sort(e->order.begin(), e->order.end(), [&](int a, int b) {
            if (e->slots[a].l != e->slots[b].l) return e->slots[a].l < e->slots[b].l;
            return e->slots[a].r < e->slots[b].r;
        });
        if (e->tail_async())  // the previous batch's tails read the reset flags of ITS slot table when they start
            for (int b = 0; b < 2; b++)
                if (e->tail_pending[b]) CU(cudaStreamWaitEvent(e->client_stream(), e->ev_tail[b], 0));
        CU(cudaMemcpyAsync(e->ca.slots, e->slots.data(), sizeof(ClientSlot) * e->slots.size(), cudaMemcpyHostToDevice,
                           e->client_stream()));
        if (!e->order.empty())
            CU(cudaMemcpyAsync(e->d_order, e->order.data(), sizeof(int) * e->order.size(), cudaMemcpyHostToDevice,
                               e->client_stream()));
        // pageable source: the runtime stages it before returning, so host vectors may change afterwards
        e->slots_dirty = false;
    }
    e->last_client_frames = nframes;
    cudaStream_t cs = e->client_stream(), ts = e->tail_stream();
    const bool async_tail = e->tail_async();
    const int buf = async_tail ? e->tail_buf : 0;
    const ClientArrays cab = e->client_arrays(buf);
    if (e->banks > 1) {
        // everything enqueued on the forward stream so far (this bank's FFT, and any broadcast the caller
        // put behind it) must land before the clients read the bank
        CU(cudaEventRecord(e->ev_fwd[e->cur_bank], e->stream));
        CU(cudaStreamWaitEvent(cs, e->ev_fwd[e->cur_bank], 0));
    }
    if (async_tail && e->tail_pending[buf]) {  // the tails of two batches ago still own this audio / PCM buffer
        CU(cudaStreamWaitEvent(cs, e->ev_tail[buf], 0));
        e->tail_pending[buf] = false;
    }
    e->last_buf = buf;
    if (e->order.empty()) {
        CU(cudaMemsetAsync(cab.valid, 0, (size_t)e->ca.max_clients * nframes, cs));  // closed slots read back as invalid
        return 0;
    }
    ClientLaunch cl{};
    cl.spec = e->spec_ptr();
    cl.spec_stride = e->spec_stride;
    cl.nframes = nframes;
    cl.frame_num0 = frame_num;
    cl.fft_size = e->size;
    cl.is_real = e->is_real ? 1 : 0;
    cl.order = e->d_order;
    cl.nactive = (int)e->order.size();
    cl.cpb = e->tail_cpb;
    cl.fchunk = e->demod_fchunk;
    cl.prof = e->d_prof;
    int rc = (e->opt_client_mask & 1) ? launch_demod(e, cab, cl) : 0;
    if (rc) return rc;
    if (e->banks > 1) {  // the demodulation is the only reader of the spectrum bank
        CU(cudaEventRecord(e->ev_cli[e->cur_bank], cs));
        e->cli_pending[e->cur_bank] = true;
    }
    if (async_tail) {
        CU(cudaEventRecord(e->ev_demod[buf], cs));
        CU(cudaStreamWaitEvent(ts, e->ev_demod[buf], 0));
    }
    CU(cudaMemsetAsync(cab.valid, 0, (size_t)e->ca.max_clients * nframes, ts));  // closed slots read back as invalid
    rc = (e->opt_client_mask & 2) ? launch_tail(e, cab, cl) : 0;
    if (rc) return rc;
    if (async_tail) {
        CU(cudaEventRecord(e->ev_tail[buf], ts));
        e->tail_pending[buf] = true;
        e->tail_buf ^= 1;
    }
    // one-shot reset flags have been consumed by this launch
    bool any = false;
    for (auto &s : e->slots)
        if (s.flags & (CF_RESET_AGC | CF_RESET_ALL)) {
            s.flags &= ~(CF_RESET_AGC | CF_RESET_ALL);
            any = true;
        }
    if (any) e->slots_dirty = true;
    return 0;
}

struct WfDesc { size_t src, dst; long long len; };
// N x WaterfallClient::send_waterfall (src/waterfall.cpp:44-51): client c's row q_level[l .. r) -> its place in one buffer
__global__ void waterfall_gather_kernel(const int8_t *quant, const WfDesc *desc, int8_t *out) {
    const WfDesc d = desc[blockIdx.x];
    const int8_t *s = quant + d.src;
    int8_t *o = out + d.dst;
    for (long long i = threadIdx.x; i < d.len; i += blockDim.x) o[i] = s[i];
}

}  // namespace

// ================================================================================================
// C ABI (the only symbols the library exports)
// ================================================================================================
#pragma GCC visibility push(default)
extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }
const char *b200_last_error(void) { return g_err; }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200_engine_create(b200_engine **out, size_t size, int nthreads, int downsample_levels, int brightness_offset,
                       int device) {
    (void)nthreads;
    if (!out) return fail(B200_EINVAL, "null out pointer");
    *out = nullptr;
    if (size < 2 || (size & (size - 1))) return fail(B200_ENOTSUP, "fft size %zu is not a power of two", size);
    int count = b200_device_count();
    if (count == 0) return fail(B200_ENODEV, "No CUDA devices found");  // src/fft_cuda.cu:10-13
    if (device < 0 || device >= count) return fail(B200_ENODEV, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    b200_engine *e = new b200_engine();
    e->device = device;
    e->size = size;
    e->levels = downsample_levels;
    e->size_log2 = (int)round(log2((double)size)) + brightness_offset;  // src/fft_impl.cpp:68
    cudaError_t err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    e->stream = e->own_stream;
    if (err != cudaSuccess) {
        delete e;
        return fail(B200_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(err));
    }
    // Hann window built on the host exactly like build_hann_window (src/utils/dsp.cpp:6-11) and uploaded,
    // as the reference's own CUDA backend does (src/fft_cuda.cu:15-17)
    std::vector<float> w(size);
    const int num = (int)size;
    for (int i = 0; i < num; i++) w[i] = 0.5 * (1 - cosf(2 * M_PI * i / num));
    err = cudaMalloc(&e->d_window, sizeof(float) * size);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_window, w.data(), sizeof(float) * size, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        b200_engine_destroy(e);
        return fail(B200_ECUDA, "window upload failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return 0;
}

static void free_client_buffers(b200_engine *e);
void b200_engine_destroy(b200_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    free_client_buffers(e);
    void *dev[] = {e->d_ssync, e->d_winT, e->d_items, e->d_nitems, e->d_prof, e->d_window, e->d_Y, e->d_Z, e->d_spec_raw, e->d_quant, e->d_ptop,
                   e->d_pscratch, e->d_flags, e->d_flag_err, e->d_twA1, e->d_twA2, e->d_TL, e->d_TH, e->d_TLr, e->d_THr, e->d_pre, e->d_TLM, e->d_THM,
                   e->d_done, e->d_qtab, e->d_ring};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->h_quant) cudaFreeHost(e->h_quant);
    for (int i = 0; i < 2; i++)
        if (e->ev_fetch[i]) cudaEventDestroy(e->ev_fetch[i]);
    if (e->h_abort) cudaFreeHost(e->h_abort);
    if (e->d_wf_desc) cudaFree(e->d_wf_desc);
    if (e->d_wf_out) cudaFree(e->d_wf_out);
    if (e->h_wf_out) cudaFreeHost(e->h_wf_out);
    if (e->h_tail_err) cudaFreeHost(e->h_tail_err);
    if (e->ev_fwd_done) cudaEventDestroy(e->ev_fwd_done);
    for (int b = 0; b < 4; b++)
        if (e->ev_push[b]) cudaEventDestroy(e->ev_push[b]);
    if (e->cstream) {
        cudaStreamSynchronize(e->cstream);
        for (int b = 0; b < 4; b++) {
            if (e->ev_fwd[b]) cudaEventDestroy(e->ev_fwd[b]);
            if (e->ev_cli[b]) cudaEventDestroy(e->ev_cli[b]);
        }
        cudaStreamDestroy(e->cstream);
    }
    if (e->tstream) {
        cudaStreamSynchronize(e->tstream);
        cudaStreamSynchronize(e->d2h_stream);
        for (int b = 0; b < 2; b++) {
            cudaEventDestroy(e->ev_demod[b]);
            cudaEventDestroy(e->ev_tail[b]);
        }
        cudaEventDestroy(e->ev_join);
        cudaStreamDestroy(e->tstream);
        cudaStreamDestroy(e->d2h_stream);
    }
    for (int i = 0; i < kMaxPeers; i++)
        if (e->peer_stream[i]) {
            cudaStreamSynchronize(e->peer_stream[i]);
            cudaEventDestroy(e->ev_peer[i]);
            cudaStreamDestroy(e->peer_stream[i]);
        }
    if (e->copy_stream) {
        cudaStreamSynchronize(e->copy_stream);
        for (int i = 0; i < b200_engine::kMaxBlocks; i++) {
            cudaEventDestroy(e->ev_in[i]);
            cudaEventDestroy(e->ev_out[i]);
            cudaEventDestroy(e->ev_ring_free[i]);
        }
        cudaStreamDestroy(e->copy_stream);
    }
    for (int l = 0; l < 4; l++) {
        if (e->lane_stream[l]) {
            cudaStreamSynchronize(e->lane_stream[l]);
            cudaStreamDestroy(e->lane_stream[l]);
        }
        if (e->ev_lane[l]) cudaEventDestroy(e->ev_lane[l]);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

int b200_set_output_additional_size(b200_engine *e, size_t n) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (e->planned) return fail(B200_ESTATE, "set_output_additional_size after plan");
    e->additional = n;
    return 0;
}

float *b200_malloc(b200_engine *e, size_t nfloats) {
    if (!e) return nullptr;
    cudaSetDevice(e->device);
    float *p = nullptr;
    if (cudaHostAlloc(&p, sizeof(float) * nfloats, cudaHostAllocDefault) != cudaSuccess) {
        fail(B200_ENOMEM, "cudaHostAlloc of %zu floats failed", nfloats);
        return nullptr;
    }
    e->host_allocs[reinterpret_cast<const char *>(p)] = sizeof(float) * nfloats;
    return p;
}
void b200_free(b200_engine *e, float *buf) {
    if (!e || !buf) return;
    e->host_allocs.erase(reinterpret_cast<const char *>(buf));
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->last_a2 == buf) e->last_a2 = nullptr;
    cudaFreeHost(buf);
}

int b200_plan_c2c(b200_engine *e, int direction, int options) {
    (void)options;
    if (!e) return fail(B200_EINVAL, "null engine");
    if (direction != B200_FORWARD) return fail(B200_ENOTSUP, "only the forward transform is on this path (src/fft.cpp:28)");
    int rc = plan_common(e, false);
    if (rc) return rc;
    rc = alloc_batch(e, 1);
    if (rc) return rc;
    return alloc_ring(e, 3);
}
int b200_plan_r2c(b200_engine *e, int options) {
    (void)options;
    if (!e) return fail(B200_EINVAL, "null engine");
    int rc = plan_common(e, true);
    if (rc) return rc;
    rc = alloc_batch(e, 1);
    if (rc) return rc;
    return alloc_ring(e, 3);
}

float *b200_get_output_buffer(b200_engine *e) { return e ? e->h_out : nullptr; }
int8_t *b200_get_quantized_buffer(b200_engine *e) { return e ? e->h_quant : nullptr; }

int b200_load_real_input(b200_engine *e, const float *a1, const float *a2) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->is_real) return fail(B200_ESTATE, "load_real_input on a c2c plan");
    if (e->in_format != B200_FMT_F32) {
        e->in_format = B200_FMT_F32;
        e->last_a2 = nullptr;
    }
    return load_common(e, a1, a2);
}
int b200_load_complex_input(b200_engine *e, const float *a1, const float *a2) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (e->is_real) return fail(B200_ESTATE, "load_complex_input on an r2c plan");
    if (e->in_format != B200_FMT_F32) {
        e->in_format = B200_FMT_F32;
        e->last_a2 = nullptr;
    }
    return load_common(e, a1, a2);
}
int b200_load_raw_input(b200_engine *e, const void *a1, const void *a2) {
    if (!e) return fail(B200_EINVAL, "null engine");
    return load_common(e, a1, a2);
}

int b200_set_option(b200_engine *e, int option, int value) {
    switch (option) {
    case B200_OPT_RELOAD_BOTH:
    case B200_OPT_HOST_MIRROR:
    case B200_OPT_INPUT_FORMAT:
    case B200_OPT_PEER_STORES:
    case B200_OPT_PCM16: return b200_debug_option(e, option, value);
    default: return fail(B200_EINVAL, "unknown option %d (tuning knobs: b200_debug_option, phantomsdr_b200_debug.h)", option);
    }
}

int b200_debug_option(b200_engine *e, int option, int value) {
    if (!e) return fail(B200_EINVAL, "null engine");
    switch (option) {
    case B200_OPT_RELOAD_BOTH: e->opt_reload_both = value ? 1 : 0; return 0;
    case B200_OPT_HOST_MIRROR: e->opt_mirror = value & 3; return 0;
    case B200_OPT_STAGE_MASK: e->opt_stage_mask = value & 7; return 0;
    case B200_OPT_FUSED_PYRAMID:
        if (value < -1 || value > 2) return fail(B200_EINVAL, "fused pyramid mode must be -1 (auto), 0, 1 or 2");
        e->opt_fused_pyramid = value;
        return 0;
    case B200_OPT_TMA: e->opt_tma = value < 0 ? 0 : (value > 4 ? 4 : value); return 0;
    case B200_OPT_PEER_STORES: e->opt_peer_stores = value ? 1 : 0; return 0;
    case B200_OPT_TAIL_PIPELINE: e->opt_tail_pipe = value ? 1 : 0; return 0;
    case B200_OPT_PACKED_MATH:
        e->opt_packed = value & 3;
        if ((value & 2) && !e->d_qtab) {  // build and upload the tables of levels 0..2
            if (!e->planned) return fail(B200_ESTATE, "the quantiser tables need a planned engine");
            CU(cudaSetDevice(e->device));
            std::vector<uint2> tab(3 * 2048);
            std::vector<uint32_t> lo(2048), hi(2048);
            std::vector<uint8_t> base(2048);
            for (int lv = 0; lv < 3; lv++) {
                build_quant_table(e->size_log2 - lv, lo.data(), hi.data(), base.data());
                for (int c = 0; c < 2048; c++) tab[lv * 2048 + c] = make_uint2(lo[c], base[c] | ((hi[c] - lo[c]) << 8));
            }
            CU(cudaMalloc(&e->d_qtab, sizeof(uint2) * tab.size()));
            CU(cudaMemcpy(e->d_qtab, tab.data(), sizeof(uint2) * tab.size(), cudaMemcpyHostToDevice));
        }
        return 0;
    case B200_OPT_PYRAMID_LAG:
        if (value < 1 || value > 8) return fail(B200_EINVAL, "pyramid lag must be 1..8 frames");
        e->opt_pyr_lag = value;
        return 0;
    case B200_OPT_PASS1_ORDER:
        if (value < 0 || value > 2) return fail(B200_EINVAL, "pass-1 order must be 0, 1 or 2");
        e->opt_pass1_order = value;
        return 0;
    case B200_OPT_FWD_LANES:
        if (value < 1 || value > 4) return fail(B200_EINVAL, "forward lanes must be 1..4");
        e->opt_lanes = value;
        return 0;
    case B200_OPT_FWD_SUB_FRAMES:
        if (value < 1 || value > 64) return fail(B200_EINVAL, "sub-batch frames must be 1..64");
        e->opt_sub_frames = value;
        return 0;
    case B200_OPT_DEMOD_GENERIC: e->opt_demod_generic = value ? 1 : 0; return 0;
    case B200_OPT_R2C_SPLIT_KERNEL: e->opt_r2c_split = value ? 1 : 0; return 0;
    case B200_OPT_TAIL_SMEM_KB:
        if (value < 0 || value > 224) return fail(B200_EINVAL, "tail shared memory must be 0..224 KB");
        e->opt_tail_smem_kb = value;
        if (e->have_clients && e->use_tail2) {
            e->tail2_dyn = tail2_dynamic_smem(e);
            CU(cudaFuncSetAttribute(client_tail2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->tail2_dyn));
        }
        return 0;
    case B200_OPT_FWD_SMS:
        if (value < 0 || value > 1024) return fail(B200_EINVAL, "forward SM count must be 0 (all) .. 1024");
        e->opt_fwd_sms = value;
        return 0;
    case B200_OPT_PASS1_SPLIT:
        if (value < 1 || value > 64) return fail(B200_EINVAL, "pass-1 split must be 1..64");
        e->opt_p1_split = value;
        return 0;
    case B200_OPT_CLIENT_STAGE_MASK: e->opt_client_mask = value & 3; return 0;
    case B200_OPT_DEMOD_CHUNK:
        if (value < -1 || value > 64) return fail(B200_EINVAL, "demodulation chunk must be -1 (automatic), 0 (sequential kernel) .. 64 frames");
        e->opt_demod_chunk = value;
        return 0;
    case B200_OPT_PCM16:
        if (e->have_clients && !e->use_tail2) return fail(B200_ENOTSUP, "int16 PCM needs the pipelined tail kernel");
        e->opt_pcm16 = value ? 1 : 0;
        e->t2.pcm16 = e->opt_pcm16;
        return 0;
    case B200_OPT_STREAM_GRID:
        if (value < 0 || value > 1024) return fail(B200_EINVAL, "stream grid must be 0 (one CTA per SM) .. 1024");
        e->opt_stream_grid = value;
        return 0;
    case B200_OPT_STREAM_LAG1:
        if (value < 0 || value > 16) return fail(B200_EINVAL, "stream lag must be 0..16 frames");
        e->opt_stream_lag1 = value;
        return 0;
    case B200_OPT_STREAM_LAG2:
        if (value < 0 || value > 32) return fail(B200_EINVAL, "stream lag must be 0..32 frames");
        e->opt_stream_lag2 = value;
        return 0;
    case B200_OPT_STREAM_RING:
        if (value < 1 || value > 64) return fail(B200_EINVAL, "Y ring must be 1..64 frame slots");
        e->opt_stream_ring = value;
        return 0;
    case B200_OPT_INPUT_FORMAT:
        if (value < B200_FMT_F32 || value > B200_FMT_S16) return fail(B200_EINVAL, "unknown input format %d", value);
        if (value != e->in_format) {
            e->in_format = value;
            e->last_a2 = nullptr;  // ring contents are in the old format
            e->head = -1;
            int rc = tma_ring_map(e);  // the tensor map of the hop ring depends on the element size
            if (rc) return rc;
        }
        return 0;
    }
    return fail(B200_EINVAL, "unknown option %d", option);
}

int b200_execute(b200_engine *e) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "execute before plan");
    if (e->frame_hop0 < 0) return fail(B200_ESTATE, "execute before load_*_input");
    CU(cudaSetDevice(e->device));
    const bool send_frame = e->fwd_frame % (uint64_t)std::max(1, e->wf_skip) == 0;
    int rc = run_forward(e, e->frame_hop0, 1);
    if (rc) return rc;
    if (e->opt_mirror & 1) {
        const size_t bins = e->is_real ? e->size / 2 + 1 : e->size + e->additional;
        CU(cudaMemcpyAsync(e->h_out, e->spec_ptr(), sizeof(float2) * bins, cudaMemcpyDeviceToHost, e->stream));
    }
    if ((e->opt_mirror & 2) && send_frame)
        CU(cudaMemcpyAsync(e->h_quant, e->quant_ptr(), e->pyr_bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return stream_check(e);
}

int b200_set_waterfall_cadence(b200_engine *e, int skip_num) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (skip_num < 1) return fail(B200_EINVAL, "skip_num must be >= 1 (src/fft.cpp:33)");
    e->wf_skip = skip_num;
    return 0;
}
int b200_set_frame_number(b200_engine *e, uint64_t frame_num) {
    if (!e) return fail(B200_EINVAL, "null engine");
    e->fwd_frame = frame_num;
    return 0;
}

void *b200_device_spectrum(b200_engine *e) { return e ? e->spec_ptr() : nullptr; }
void *b200_device_quantized(b200_engine *e) { return e ? e->quant_ptr() : nullptr; }
void *b200_device_hop_ring(b200_engine *e) { return e ? e->d_ring : nullptr; }
size_t b200_hop_floats(b200_engine *e) { return e ? e->hop_samples : 0; }
size_t b200_spectrum_bins(b200_engine *e) {
    if (!e) return 0;
    return e->is_real ? e->size / 2 + 1 : e->size + e->additional;
}
size_t b200_pyramid_bytes(b200_engine *e) { return e ? e->pyr_bytes : 0; }
size_t b200_spectrum_stride(b200_engine *e) { return e ? e->spec_stride : 0; }
size_t b200_pyramid_stride(b200_engine *e) { return e ? e->pyr_stride : 0; }

int b200_set_hop_ring(b200_engine *e, size_t nhops) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "set_hop_ring before plan");
    if (nhops < 2) return fail(B200_EINVAL, "hop ring needs at least 2 hops");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return alloc_ring(e, nhops);
}
int b200_set_batch_frames(b200_engine *e, int max_frames) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "set_batch_frames before plan");
    if (e->have_clients) return fail(B200_ESTATE, "set_batch_frames must precede clients_create");
    if (max_frames < 1 || max_frames > 64) return fail(B200_EINVAL, "batch frames must be 1..64");
    if (e->spec_bound) return fail(B200_ESTATE, "spectrum is bound to an external buffer");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return alloc_batch(e, max_frames);
}
int b200_execute_device_batch(b200_engine *e, size_t hop_index, int nframes) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "execute before plan");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    if (hop_index >= e->nhops) return fail(B200_EINVAL, "hop index %zu outside ring of %zu", hop_index, e->nhops);
    CU(cudaSetDevice(e->device));
    return run_forward(e, (long)hop_index, nframes);
}
int b200_execute_device(b200_engine *e, size_t hop_index) { return b200_execute_device_batch(e, hop_index, 1); }
int b200_sync(b200_engine *e) {
    if (!e) return fail(B200_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->cstream) CU(cudaStreamSynchronize(e->cstream));
    if (e->tstream) CU(cudaStreamSynchronize(e->tstream));
    if (e->d2h_stream) CU(cudaStreamSynchronize(e->d2h_stream));
    return stream_check(e);
}
int b200_set_pipeline(b200_engine *e, int banks) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "set_pipeline before plan");
    if (banks < 1 || banks > 4) return fail(B200_EINVAL, "banks must be 1..4");
    if (e->spec_bound) return fail(B200_ESTATE, "spectrum is bound to an external buffer");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->cstream) CU(cudaStreamSynchronize(e->cstream));
    if (banks > 1 && !e->cstream) {
        CU(cudaStreamCreateWithFlags(&e->cstream, cudaStreamNonBlocking));
        for (int b = 0; b < 4; b++) {
            CU(cudaEventCreateWithFlags(&e->ev_fwd[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e->ev_cli[b], cudaEventDisableTiming));
        }
    }
    e->banks = banks;
    return alloc_batch(e, e->batch);
}
int b200_select_bank(b200_engine *e, int bank) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (bank < 0 || bank >= e->banks) return fail(B200_EINVAL, "bank %d outside 0..%d", bank, e->banks - 1);
    e->cur_bank = bank;
    return 0;
}
int b200_bank_acquire(b200_engine *e) {
    if (!e) return fail(B200_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    return bank_acquire(e);
}
int b200_client_stream_wait_event(b200_engine *e, void *cuda_event) {
    if (!e || !cuda_event) return fail(B200_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamWaitEvent(e->client_stream(), reinterpret_cast<cudaEvent_t>(cuda_event), 0));
    return 0;
}
int b200_join_streams(b200_engine *e) {
    if (!e) return fail(B200_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    if (e->banks > 1)
        for (int b = 0; b < e->banks; b++)
            if (e->cli_pending[b]) CU(cudaStreamWaitEvent(e->stream, e->ev_cli[b], 0));
    if (e->tail_async())
        for (int b = 0; b < 2; b++)
            if (e->tail_pending[b]) CU(cudaStreamWaitEvent(e->stream, e->ev_tail[b], 0));
    return 0;
}
void *b200_stream(b200_engine *e) { return e ? (void *)e->stream : nullptr; }
int b200_set_stream(b200_engine *e, void *cuda_stream) {
    if (!e) return fail(B200_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : e->own_stream;
    return 0;
}

int b200_bind_spectrum(b200_engine *e, void *dev_ptr) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "bind_spectrum before plan");
    e->spec_bound = reinterpret_cast<float2 *>(dev_ptr);
    return 0;
}
int b200_set_peer_spectra(b200_engine *e, int npeers, void *const *dev_ptrs) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (npeers < 0 || npeers > kMaxPeers) return fail(B200_EINVAL, "npeers must be 0..%d", kMaxPeers);
    e->npeers = npeers;
    for (int i = 0; i < npeers; i++) {
        e->peers[i] = reinterpret_cast<float2 *>(dev_ptrs[i]);
        e->peer_lo[i][0] = 0;  // default: the whole frame
        e->peer_hi[i][0] = 0xffffffffu;
        e->peer_lo[i][1] = e->peer_hi[i][1] = 0;
    }
    return 0;
}
int b200_set_peer_ranges(b200_engine *e, int peer, uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (peer < 0 || peer >= e->npeers) return fail(B200_EINVAL, "peer %d out of range", peer);
    e->peer_lo[peer][0] = lo0;
    e->peer_hi[peer][0] = hi0;
    e->peer_lo[peer][1] = lo1;
    e->peer_hi[peer][1] = hi1;
    return 0;
}
void *b200_device_spectrum_base(b200_engine *e) { return e ? e->d_spec_raw : nullptr; }
size_t b200_device_spectrum_offset(b200_engine *e) { return e ? (e->is_real ? 16 : 15) * sizeof(float2) : 0; }
void *b200_flag_buffer(b200_engine *e) {
    if (!e) return nullptr;
    if (!e->d_flags) {
        cudaSetDevice(e->device);
        if (cudaMalloc(&e->d_flags, 64 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
        cudaMemset(e->d_flags, 0, 64 * sizeof(unsigned long long));
        if (cudaMalloc(&e->d_flag_err, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(e->d_flag_err, 0, sizeof(int));
    }
    return e->d_flags;
}
static int stream_setup(b200_engine *e);
static int pick_stream(b200_engine *e, int which, cudaStream_t *out) {
    if (which == 2) {
        int rc = stream_setup(e);
        if (rc) return rc;
        *out = e->copy_stream;
    } else {
        *out = which ? e->client_stream() : e->stream;
    }
    return 0;
}
int b200_push_peers(b200_engine *e, int nframes) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    CU(cudaSetDevice(e->device));
    int rc = stream_setup(e);
    if (rc) return rc;
    if (!e->ev_fwd_done) {
        CU(cudaEventCreateWithFlags(&e->ev_fwd_done, cudaEventDisableTiming));
        for (int b = 0; b < 4; b++) CU(cudaEventCreateWithFlags(&e->ev_push[b], cudaEventDisableTiming));
    }
    // the selected bank as the forward stream leaves it -> every peer's sub-band, by the copy engines. One stream per peer:
    // on a single stream the seven pushes of an 8-GPU box ran one after the other (470 MB per 64-frame batch at one copy
    // engine's rate, 1.2 ms - longer than the forward group) and bounded the ingest rate; the copy stream joins them, so
    // whatever the caller enqueues there next (the "bank has landed" signal) still follows every push.
    CU(cudaEventRecord(e->ev_fwd_done, e->stream));
    const size_t bank_off = (size_t)e->cur_bank * e->batch * e->spec_stride;
    const float2 *src = e->d_spec + bank_off;
    const size_t pitch = e->spec_stride * sizeof(float2);
    for (int i = 0; i < e->npeers; i++) {
        if (!e->peer_stream[i]) {
            CU(cudaStreamCreateWithFlags(&e->peer_stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&e->ev_peer[i], cudaEventDisableTiming));
        }
        cudaStream_t ps = e->peer_stream[i];
        CU(cudaStreamWaitEvent(ps, e->ev_fwd_done, 0));
        float2 *dst = e->peers[i] + bank_off;
        for (int j = 0; j < 2; j++) {
            size_t lo = e->peer_lo[i][j], hi = std::min<size_t>(e->peer_hi[i][j], e->spec_stride);
            if (hi <= lo) continue;
            CU(cudaMemcpy2DAsync(dst + lo, pitch, src + lo, pitch, (hi - lo) * sizeof(float2), nframes, cudaMemcpyDeviceToDevice, ps));
        }
        CU(cudaEventRecord(e->ev_peer[i], ps));
        CU(cudaStreamWaitEvent(e->copy_stream, e->ev_peer[i], 0));
    }
    if (e->npeers == 0) CU(cudaStreamWaitEvent(e->copy_stream, e->ev_fwd_done, 0));
    CU(cudaEventRecord(e->ev_push[e->cur_bank], e->copy_stream));
    e->push_pending[e->cur_bank] = true;
    return 0;
}
int b200_pull_spectrum(b200_engine *e, const void *remote_spectrum, int nframes, uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1) {
    if (!e || !remote_spectrum) return fail(B200_EINVAL, "null argument");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    if (e->spec_bound) return fail(B200_ESTATE, "spectrum is bound to an external buffer");
    CU(cudaSetDevice(e->device));
    // the ingest rank's bank -> the same bank here, by THIS GPU's copy engine, behind whatever the client stream already
    // holds (the wait for the ingest rank's "bank ready" flag) and in front of the demodulation that follows
    const size_t bank_off = (size_t)e->cur_bank * e->batch * e->spec_stride;
    const float2 *src = static_cast<const float2 *>(remote_spectrum) + bank_off;
    float2 *dst = e->d_spec + bank_off;
    const size_t pitch = e->spec_stride * sizeof(float2);
    const uint32_t lo[2] = {lo0, lo1}, hi[2] = {hi0, hi1};
    for (int j = 0; j < 2; j++) {
        const size_t a = lo[j], b = std::min<size_t>(hi[j], e->spec_stride);
        if (b <= a) continue;
        CU(cudaMemcpy2DAsync(dst + a, pitch, src + a, pitch, (b - a) * sizeof(float2), nframes, cudaMemcpyDeviceToDevice, e->client_stream()));
    }
    return 0;
}
static int flag_list(b200_engine *e, void *const *flag_ptrs, int n, FlagList *fl) {
    if (!e || !flag_ptrs) return fail(B200_EINVAL, "null argument");
    if (n < 1 || n > kMaxPeers) return fail(B200_EINVAL, "1..%d flags per call", kMaxPeers);
    if (!b200_flag_buffer(e)) return fail(B200_ENOMEM, "flag buffer allocation failed");
    fl->n = n;
    for (int i = 0; i < n; i++) fl->ptr[i] = reinterpret_cast<unsigned long long *>(flag_ptrs[i]);
    return 0;
}
int b200_enqueue_signal(b200_engine *e, int client_stream, void *const *flag_ptrs, int n, uint64_t value) {
    FlagList fl;
    int rc = flag_list(e, flag_ptrs, n, &fl);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    cudaStream_t st;
    rc = pick_stream(e, client_stream, &st);
    if (rc) return rc;
    flag_signal_kernel<<<1, 32, 0, st>>>(fl, value);
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
int b200_enqueue_wait(b200_engine *e, int client_stream, void *const *flag_ptrs, int n, uint64_t min_value, int timeout_ms) {
    FlagList fl;
    int rc = flag_list(e, flag_ptrs, n, &fl);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    const unsigned long long timeout_ns = (unsigned long long)std::max(1, timeout_ms) * 1000000ull;
    cudaStream_t st;
    rc = pick_stream(e, client_stream, &st);
    if (rc) return rc;
    flag_wait_kernel<<<1, 32, 0, st>>>(fl, min_value, timeout_ns, e->d_flag_err);
    e->flag_waits++;
    e->launches++;
    CU(cudaGetLastError());
    return 0;
}
int b200_flag_error(b200_engine *e) {
    if (!e || !e->d_flag_err) return 0;
    int v = 0;
    cudaSetDevice(e->device);
    cudaMemcpy(&v, e->d_flag_err, sizeof(int), cudaMemcpyDeviceToHost);
    return v;
}
int b200_ipc_export(b200_engine *e, const void *dev_ptr, uint8_t handle[64]) {
    if (!e || !dev_ptr || !handle) return fail(B200_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, 64);
    return 0;
}
int b200_ipc_open(b200_engine *e, const uint8_t handle[64], void **dev_ptr) {
    if (!e || !dev_ptr || !handle) return fail(B200_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int b200_ipc_close(b200_engine *e, void *dev_ptr) {
    if (!e || !dev_ptr) return fail(B200_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// signal slot group
// ------------------------------------------------------------------------------------------------
// device / host buffers of the client table (b200_clients_create); safe on a partially created table
static void free_client_buffers(b200_engine *e) {
    void *dev[] = {e->d_redo, e->t2.dcx, e->t2.dcm, e->t2.sum, e->t2.gain, e->t2.since, e->t2.blk, e->t2.ring, e->t2.suf, e->t2.cmax,
                   e->d_order, (void *)e->ca.Wn, e->ca.slots, e->ca.real_prev, e->ca.real_hi, e->ca.hi_diverged, e->ca.bb_hi, e->ca.bb_last,
                   e->ca.dc_x, e->ca.dc_m, e->ca.dc_sum, e->ca.agc_ring, e->ca.agc_cmax, e->ca.agc_gain, e->ca.agc_t0, e->ca.audio_pre,
                   e->ca.valid_a, e->ca.pwr, e->ca.pcm, e->ca.valid};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (e->h_tail_err) cudaFreeHost(e->h_tail_err);
    e->h_tail_err = nullptr;
    e->d_redo = nullptr;
    e->d_order = nullptr;
    e->t2 = Tail2State{};
    e->ca = ClientArrays{};
    e->have_clients = false;
    e->use_tail2 = false;
}

static int clients_create_impl(b200_engine *e, int max_clients, int audio_fft_size, int audio_max_sps);
int b200_clients_create(b200_engine *e, int max_clients, int audio_fft_size, int audio_max_sps) {
    const int rc = clients_create_impl(e, max_clients, audio_fft_size, audio_max_sps);
    // a failure half way (out of memory, unsupported size) leaves nothing behind: the call can be repeated with other sizes
    if (rc && e && !e->have_clients) free_client_buffers(e);
    return rc;
}
static int clients_create_impl(b200_engine *e, int max_clients, int audio_fft_size, int audio_max_sps) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "clients_create before plan");
    if (e->have_clients) return fail(B200_ESTATE, "clients already created");
    if (max_clients < 1) return fail(B200_EINVAL, "max_clients must be >= 1");
    if (audio_fft_size < 4 || audio_fft_size % 4) return fail(B200_EINVAL, "audio_fft_size must be a positive multiple of 4");
    if (!e->is_real && (size_t)audio_fft_size > e->additional)
        return fail(B200_EINVAL, "output additional size %zu < audio_fft_size %d (src/spectrumserver.cpp:214)",
                    e->additional, audio_fft_size);
    CU(cudaSetDevice(e->device));
    ClientArrays &ca = e->ca;
    ca.n = audio_fft_size;
    ca.h = audio_fft_size / 2;
    ca.D = audio_max_sps / 750 * 2;  // DCBlocker<float>(audio_max_sps / 750 * 2), src/signal.cpp:54
    if (ca.D < 1) return fail(B200_EINVAL, "audio_max_sps %d gives an empty DC blocker", audio_max_sps);
    // AGC(0.2f, 50.0f, 300.0f, 200.0f, audio_max_sps), src/signal.cpp:55, src/utils/audioprocessing.cpp:5-15
    const float sr = (float)audio_max_sps;
    ca.L = (int)static_cast<size_t>(200.0f * sr / 1000.0f);
    ca.attack = (float)(1 - exp((double)(-1.0f / (50.0f * 0.001f * sr))));
    ca.release = (float)(1 - exp((double)(-1.0f / (300.0f * 0.001f * sr))));
    ca.desired = 0.2f;
    if (ca.L < ca.h + 1)
        return fail(B200_ENOTSUP, "AGC look-ahead (%d samples) shorter than one frame of audio (%d): not supported", ca.L,
                    ca.h);
    ca.NC = (ca.L - 1 + ca.h - 1) / ca.h + 2;
    ca.max_clients = max_clients;
    std::vector<int> rad = factorize(audio_fft_size);
    if ((int)rad.size() > kMaxStages) return fail(B200_ENOTSUP, "audio_fft_size has too many factors");
    ca.nstages = (int)rad.size();
    {
        auto magic = [](unsigned d) { return (unsigned)(((1ull << 32) / d) + 1); };  // x / d == umulhi(x, magic) for x*d < 2^32
        unsigned sdiv = 1;
        for (int i = 0; i < ca.nstages; i++) {
            ca.radix[i] = rad[i];
            ca.mul_s[i] = sdiv == 1 ? 0u : magic(sdiv);  // s = 1 handled below (umulhi(x, 2^32) does not fit)
            ca.mul_items[i] = magic((unsigned)(audio_fft_size / rad[i]));
            sdiv *= (unsigned)rad[i];
        }
        ca.mul_n = magic((unsigned)audio_fft_size);
    }
    // frames whose inverse FFTs are batched in shared memory by the demod kernel (2 buffers of fchunk * n complex)
    e->demod_fchunk = std::max(1, std::min(std::min(e->batch, 4), (int)((96 * 1024) / (sizeof(float2) * 2 * audio_fft_size))));  // 4: measured optimum (occupancy vs lane use)
    if (const char *v = getenv("B200_DEMOD_FCHUNK")) e->demod_fchunk = std::max(1, std::min(e->demod_fchunk, atoi(v)));
    {
        std::vector<float2> wn(audio_fft_size);
        for (int k = 0; k < audio_fft_size; k++) {
            const double a = 2.0 * M_PI * (double)k / audio_fft_size;
            wn[k] = make_float2((float)cos(a), (float)sin(a));
        }
        ca.Wn = upload_f2(wn);
        if (!ca.Wn) return fail(B200_ENOMEM, "twiddle allocation failed");
    }
    const size_t mc = max_clients, h = ca.h, F = e->batch;
#define ALLOC0(ptr, bytes)                        \
    CU(cudaMalloc((void **)&(ptr), (bytes)));     \
    CU(cudaMemset((ptr), 0, (bytes)))
    ALLOC0(ca.slots, sizeof(ClientSlot) * mc);
    ALLOC0(ca.real_prev, sizeof(float) * 2 * mc * h);   // two copies: see client_demod_warp_kernel
    ALLOC0(ca.real_hi, sizeof(float) * 2 * mc * h);
    ALLOC0(ca.hi_diverged, sizeof(int) * 2 * mc);
    ALLOC0(ca.bb_hi, sizeof(float2) * 2 * mc * h);
    ALLOC0(ca.bb_last, sizeof(float2) * 2 * mc);
    ALLOC0(e->d_redo, mc);
    ALLOC0(ca.dc_x, sizeof(float) * mc * ca.D);
    ALLOC0(ca.dc_m, sizeof(float) * mc * ca.D);
    ALLOC0(ca.dc_sum, sizeof(float) * mc * 2);
    ALLOC0(ca.agc_ring, sizeof(float) * mc * ca.NC * h);
    ALLOC0(ca.agc_cmax, sizeof(float) * mc * ca.NC);
    ALLOC0(ca.agc_gain, sizeof(float) * mc);
    ALLOC0(ca.agc_t0, sizeof(long long) * mc);
    ALLOC0(ca.audio_pre, sizeof(float) * 2 * F * mc * h);  // two batches: see tail_async()
    ALLOC0(ca.valid_a, 2 * F * mc);
    ALLOC0(ca.pwr, sizeof(float) * F * mc);
    ALLOC0(ca.pcm, sizeof(int) * 2 * F * mc * h);
    ALLOC0(ca.valid, 2 * F * mc);
    ALLOC0(e->d_order, sizeof(int) * mc);
    // lane-per-client tail pipeline: state stored [group of 32 slots][...][32]
    e->use_tail2 = e->opt_tail_pipe && ca.D <= (int)h && ca.L - 1 >= (int)h && tail2_smem(ca.D) <= 220 * 1024 &&
                   (ca.L - 1 + (int)h - 1) / (int)h >= kT2Depth + 3;  // (suffix maxima are always complete long before they are read)
    if (e->use_tail2) {
        Tail2State &t = e->t2;
        const size_t groups = (mc + 31) / 32;
        t.NB = (ca.L - 1 + (int)h - 1) / (int)h + 2;
        t.kb = (ca.L - 1 + (int)h - 1) / (int)h;                    // ceil((L - 1) / h)
        t.col0 = ((int)h - (ca.L - 1) % (int)h) % (int)h;           // (pos - (L - 1)) mod h for pos a multiple of h
        t.nsx = ca.D + (kT2Depth + 1) * kT2CH;
        t.dpow2 = (ca.D & (ca.D - 1)) == 0;
        t.pcm16 = e->opt_pcm16;
        ALLOC0(t.dcx, sizeof(float) * groups * ca.D * 32);
        ALLOC0(t.dcm, sizeof(float) * groups * ca.D * 32);
        ALLOC0(t.sum, sizeof(float) * groups * 2 * 32);
        ALLOC0(t.gain, sizeof(float) * groups * 32);
        ALLOC0(t.since, sizeof(int) * groups * 32);
        ALLOC0(t.blk, sizeof(int) * groups * 32);
        ALLOC0(t.ring, sizeof(float) * groups * t.NB * h * 32);
        ALLOC0(t.suf, sizeof(float) * groups * t.NB * h * 32);
        ALLOC0(t.cmax, sizeof(float) * groups * t.NB * 32);
        CU(cudaHostAlloc(&e->h_tail_err, sizeof(int), cudaHostAllocMapped));  // written by the kernel only when a wait expires
        *e->h_tail_err = 0;
        CU(cudaHostGetDevicePointer(&t.err, e->h_tail_err, 0));
    }
#undef ALLOC0
    e->slots.assign(mc, ClientSlot{});
    e->mids.assign(mc, 0.0);
    // launch geometry
    if (sizeof(float2) * 2 * ca.n > 200 * 1024) return fail(B200_ENOTSUP, "audio_fft_size %d too large", ca.n);
    e->tail_cpb = kTailMaxCpb;
    auto tail_bytes = [&](int cpb) {
        return sizeof(float) * (size_t)cpb * (2 * (size_t)tail_pitch(ca.D + ca.h) + 3 * (size_t)tail_pitch(ca.h));
    };
    while (e->tail_cpb > 1 && tail_bytes(e->tail_cpb) > 200 * 1024) e->tail_cpb /= 2;
    if (tail_bytes(e->tail_cpb) > 200 * 1024) return fail(B200_ENOTSUP, "audio_fft_size %d too large for the tail kernel", ca.n);
    e->tail_smem = tail_bytes(e->tail_cpb);
    {
        ClientLaunch none{};
        int rc = launch_demod(e, e->ca, none);
        if (rc) return rc;
        rc = launch_tail(e, e->ca, none);
        if (rc) return rc;
        if (e->use_tail2 && !e->tstream) {
            int lo = 0, hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU(cudaStreamCreateWithPriority(&e->tstream, cudaStreamNonBlocking, hi));  // few, long-running CTAs: schedule them first
            CU(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                CU(cudaEventCreateWithFlags(&e->ev_demod[b], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&e->ev_tail[b], cudaEventDisableTiming));
            }
            CU(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
        }
    }
    e->have_clients = true;
    e->slots_dirty = true;
    return 0;
}

static int check_slot(b200_engine *e, int slot) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->have_clients) return fail(B200_ESTATE, "clients not created");
    if (slot < 0 || slot >= e->ca.max_clients) return fail(B200_EINVAL, "slot %d out of range", slot);
    return 0;
}
static int window_ok(b200_engine *e, int l, int r) {
    // AudioClient::on_window_message, src/signal.cpp:300-314
    const int R = (int)e->R;
    if (l < 0 || l >= R || r < 0 || r >= R || l > r) return 0;
    if (r - l > e->ca.n) return 0;
    return 1;
}

int b200_client_open(b200_engine *e, int slot, int l, double audio_mid, int r, int demodulation) {
    int rc = check_slot(e, slot);
    if (rc) return rc;
    if (demodulation < B200_USB || demodulation > B200_FM) return fail(B200_EINVAL, "unknown demodulation %d", demodulation);
    if (!window_ok(e, l, r)) return fail(B200_EINVAL, "window [%d, %d) rejected (src/signal.cpp:304-311)", l, r);
    ClientSlot &s = e->slots[slot];
    s.l = l;
    s.r = r;
    s.m_floor = (int)floor(audio_mid);
    s.mode = demodulation;
    s.flags = CF_ACTIVE | CF_RESET_ALL;
    e->mids[slot] = audio_mid;
    e->slots_dirty = true;
    return 0;
}
int b200_client_set_window(b200_engine *e, int slot, int l, double audio_mid, int r) {
    int rc = check_slot(e, slot);
    if (rc) return rc;
    if (!(e->slots[slot].flags & CF_ACTIVE)) return fail(B200_ESTATE, "slot %d is closed", slot);
    if (!window_ok(e, l, r)) return fail(B200_EINVAL, "window [%d, %d) rejected (src/signal.cpp:304-311)", l, r);
    ClientSlot &s = e->slots[slot];
    s.l = l;
    s.r = r;
    s.m_floor = (int)floor(audio_mid);
    e->mids[slot] = audio_mid;
    e->slots_dirty = true;
    return 0;
}
int b200_client_set_demodulation(b200_engine *e, int slot, int demodulation) {
    int rc = check_slot(e, slot);
    if (rc) return rc;
    if (!(e->slots[slot].flags & CF_ACTIVE)) return fail(B200_ESTATE, "slot %d is closed", slot);
    ClientSlot &s = e->slots[slot];
    // unknown strings leave the mode unchanged but still reset the AGC (src/signal.cpp:316-328)
    if (demodulation >= B200_USB && demodulation <= B200_FM) s.mode = demodulation;
    s.flags |= CF_RESET_AGC;
    e->slots_dirty = true;
    return 0;
}
int b200_client_close(b200_engine *e, int slot) {
    int rc = check_slot(e, slot);
    if (rc) return rc;
    e->slots[slot].flags = 0;
    e->slots_dirty = true;
    return 0;
}

int b200_clients_execute_device(b200_engine *e, uint64_t frame_num, int nframes) {
    if (!e) return fail(B200_EINVAL, "null engine");
    return run_clients(e, frame_num, nframes);
}

int b200_clients_fetch(b200_engine *e, int frame, int32_t *pcm_out, float *pwr_out, uint8_t *valid_out) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->have_clients) return fail(B200_ESTATE, "clients not created");
    if (frame < 0 || frame >= e->batch) return fail(B200_EINVAL, "frame %d outside batch", frame);
    CU(cudaSetDevice(e->device));
    const size_t mc = e->ca.max_clients, h = e->ca.h;
    cudaStream_t cs = e->client_stream(), ts = e->tail_stream();
    const ClientArrays cab = e->client_arrays(e->last_buf);
    const size_t pb = e->opt_pcm16 ? sizeof(int16_t) : sizeof(int32_t);  // bytes per PCM sample
    if (pcm_out)
        CU(cudaMemcpyAsync(pcm_out, reinterpret_cast<const char *>(cab.pcm) + (size_t)frame * mc * h * pb, pb * mc * h, cudaMemcpyDeviceToHost, ts));
    if (pwr_out) CU(cudaMemcpyAsync(pwr_out, e->ca.pwr + (size_t)frame * mc, sizeof(float) * mc, cudaMemcpyDeviceToHost, cs));
    if (valid_out) CU(cudaMemcpyAsync(valid_out, cab.valid + (size_t)frame * mc, mc, cudaMemcpyDeviceToHost, ts));
    CU(cudaStreamSynchronize(cs));
    if (ts != cs) CU(cudaStreamSynchronize(ts));
    return stream_check(e);
}

// Asynchronous form of b200_clients_fetch for whole batches: enqueues the device->host copies of the LAST client batch
// (nframes frames of PCM / pwr / valid) behind its kernels, on the engine's own copy path, and returns; completion is
// b200_clients_fetch_wait(slot). Two slots, so that the copies of batch k overlap the kernels of batch k + 1.
int b200_clients_fetch_async(b200_engine *e, int slot, int nframes, void *pcm_out, float *pwr_out, uint8_t *valid_out) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->have_clients) return fail(B200_ESTATE, "clients not created");
    if (slot < 0 || slot > 1) return fail(B200_EINVAL, "slot must be 0 or 1");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    CU(cudaSetDevice(e->device));
    if (!e->ev_fetch[0])
        for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&e->ev_fetch[i], cudaEventDisableTiming));
    const size_t mc = e->ca.max_clients, h = e->ca.h;
    const ClientArrays cab = e->client_arrays(e->last_buf);
    cudaStream_t cs = e->client_stream();
    if (pwr_out) CU(cudaMemcpyAsync(pwr_out, e->ca.pwr, sizeof(float) * mc * nframes, cudaMemcpyDeviceToHost, cs));
    if (e->tail_async()) {
        CU(cudaEventRecord(e->ev_join, cs));
        cs = e->d2h_stream;
        CU(cudaStreamWaitEvent(cs, e->ev_tail[e->last_buf], 0));
        CU(cudaStreamWaitEvent(cs, e->ev_join, 0));
    }
    if (pcm_out)
        CU(cudaMemcpyAsync(pcm_out, cab.pcm, (e->opt_pcm16 ? sizeof(int16_t) : sizeof(int32_t)) * mc * h * nframes, cudaMemcpyDeviceToHost, cs));
    if (valid_out) CU(cudaMemcpyAsync(valid_out, cab.valid, mc * nframes, cudaMemcpyDeviceToHost, cs));
    if (e->tail_async()) CU(cudaEventRecord(e->ev_tail[e->last_buf], cs));  // the buffer is free again only when the copy has read it
    CU(cudaEventRecord(e->ev_fetch[slot], cs));
    return 0;
}
int b200_clients_fetch_wait(b200_engine *e, int slot) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (slot < 0 || slot > 1 || !e->ev_fetch[slot]) return fail(B200_ESTATE, "no asynchronous fetch in slot %d", slot);
    CU(cudaSetDevice(e->device));
    CU(cudaEventSynchronize(e->ev_fetch[slot]));
    return stream_check(e);
}

int b200_clients_execute(b200_engine *e, uint64_t frame_num, int32_t *pcm_out, float *pwr_out, uint8_t *valid_out) {
    if (!e) return fail(B200_EINVAL, "null engine");
    int rc = run_clients(e, frame_num, 1);
    if (rc) return rc;
    return b200_clients_fetch(e, 0, pcm_out, pwr_out, valid_out);
}

void *b200_device_pcm(b200_engine *e) { return e ? e->client_arrays(e->last_buf).pcm : nullptr; }
void *b200_device_pwr(b200_engine *e) { return e ? e->ca.pwr : nullptr; }
void *b200_device_valid(b200_engine *e) { return e ? e->client_arrays(e->last_buf).valid : nullptr; }

int b200_clients_read_pre_dc(b200_engine *e, float *out) {
    if (!e || !out) return fail(B200_EINVAL, "null argument");
    if (!e->have_clients) return fail(B200_ESTATE, "clients not created");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(out, e->client_arrays(e->last_buf).audio_pre, sizeof(float) * (size_t)e->ca.max_clients * e->ca.h, cudaMemcpyDeviceToHost,
                       e->client_stream()));
    CU(cudaStreamSynchronize(e->client_stream()));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// waterfall slot group
// ------------------------------------------------------------------------------------------------
int b200_waterfall_gather(b200_engine *e, int nclients, const int *level, const int *l, const int *r,
                          const size_t *out_offsets, int8_t *out) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (!e->planned) return fail(B200_ESTATE, "waterfall_gather before plan");
    if (nclients <= 0) return 0;
    if (!level || !l || !r || !out_offsets || !out) return fail(B200_EINVAL, "null argument");
    CU(cudaSetDevice(e->device));
    std::vector<size_t> src(nclients), dst(nclients);
    std::vector<int> len(nclients);
    size_t total = 0;
    for (int i = 0; i < nclients; i++) {
        if (level[i] < 0 || level[i] >= e->levels) return fail(B200_EINVAL, "client %d: level %d out of range", i, level[i]);
        const size_t width = e->R >> level[i];
        if (l[i] < 0 || r[i] < l[i] || (size_t)r[i] > width)
            return fail(B200_EINVAL, "client %d: slice [%d, %d) outside level %d", i, l[i], r[i], level[i]);
        size_t off = 0;  // src/websocket.cpp:233: fft_power_quantized += (fft_result_size >> i)
        for (int j = 0; j < level[i]; j++) off += e->R >> j;
        src[i] = off + l[i];
        len[i] = r[i] - l[i];
        dst[i] = out_offsets[i];
        total = std::max(total, dst[i] + (size_t)len[i]);
    }
    // one descriptor table up, one kernel, one copy down; device and pinned staging buffers only ever grow
    typedef WfDesc Desc;
    std::vector<Desc> desc(nclients);
    for (int i = 0; i < nclients; i++) desc[i] = {src[i], dst[i], (long long)len[i]};
    const size_t out_bytes = std::max<size_t>(total, 1);
    if (e->wf_desc_cap < (size_t)nclients) {
        if (e->d_wf_desc) cudaFree(e->d_wf_desc);
        e->d_wf_desc = nullptr;
        e->wf_desc_cap = 0;
        const size_t cap = std::max<size_t>(256, (size_t)nclients * 2);
        CU(cudaMalloc(&e->d_wf_desc, sizeof(Desc) * cap));
        e->wf_desc_cap = cap;
    }
    if (e->wf_out_cap < out_bytes) {
        if (e->d_wf_out) cudaFree(e->d_wf_out);
        if (e->h_wf_out) cudaFreeHost(e->h_wf_out);
        e->d_wf_out = e->h_wf_out = nullptr;
        e->wf_out_cap = 0;
        const size_t cap = std::max<size_t>(1 << 20, out_bytes * 2);
        CU(cudaMalloc(&e->d_wf_out, cap));
        CU(cudaHostAlloc(&e->h_wf_out, cap, cudaHostAllocDefault));
        e->wf_out_cap = cap;
    }
    CU(cudaMemcpyAsync(e->d_wf_desc, desc.data(), sizeof(Desc) * nclients, cudaMemcpyHostToDevice, e->stream));
    waterfall_gather_kernel<<<nclients, 256, 0, e->stream>>>(e->quant_ptr(), reinterpret_cast<const WfDesc *>(e->d_wf_desc),
                                                             reinterpret_cast<int8_t *>(e->d_wf_out));
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(e->h_wf_out, e->d_wf_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const int8_t *staged_p = reinterpret_cast<const int8_t *>(e->h_wf_out);
    // only the clients' own byte ranges of `out` are touched
    for (int i = 0; i < nclients; i++) memcpy(out + dst[i], staged_p + dst[i], (size_t)len[i]);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Pipelined host-block streaming: the same load -> execute -> clients path, nframes frames per call, with
// the H2D of block k+1, the kernels of block k and the D2H of block k's results on three streams.
// ------------------------------------------------------------------------------------------------
static int stream_setup(b200_engine *e) {
    if (e->copy_stream) return 0;
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < b200_engine::kMaxBlocks; i++) {
        CU(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&e->ev_ring_free[i], cudaEventDisableTiming));
    }
    return 0;
}

int b200_stream_prime(b200_engine *e, const void *older_half) {
    if (!e || !older_half) return fail(B200_EINVAL, "null argument");
    if (!e->planned) return fail(B200_ESTATE, "stream_prime before plan");
    CU(cudaSetDevice(e->device));
    int rc = stream_setup(e);
    if (rc) return rc;
    if (e->banks < 2) return fail(B200_ESTATE, "block streaming needs b200_set_pipeline(>= 2)");
    if ((long)e->nhops < 2 * (long)e->batch + 2) return fail(B200_ESTATE, "hop ring must hold at least 2*batch+2 halves");
    CU(cudaStreamSynchronize(e->stream));
    if (e->cstream) CU(cudaStreamSynchronize(e->cstream));
    CU(cudaStreamSynchronize(e->copy_stream));
    const size_t hb = e->hop_samples * e->format_bytes();
    CU(cudaMemcpyAsync(e->d_ring, older_half, hb, cudaMemcpyHostToDevice, e->copy_stream));
    CU(cudaStreamSynchronize(e->copy_stream));
    e->stream_head = 0;
    e->blk_submitted = e->blk_waited = 0;
    for (bool &p : e->blk_pending) p = false;
    // a block's new halves overwrite ring slots last read by the forward pass of the block blk_depth submissions ago
    e->blk_depth = (int)std::min<long>(b200_engine::kMaxBlocks, ((long)e->nhops - 2) / (long)e->batch);
    return 0;
}

int b200_submit_block(b200_engine *e, const void *const *new_halves, int nframes, uint64_t frame_num0, int32_t *pcm_out,
                      float *pwr_out, uint8_t *valid_out, int8_t *pyramid_out) {
    if (!e || !new_halves) return fail(B200_EINVAL, "null argument");
    if (e->stream_head < 0) return fail(B200_ESTATE, "submit_block before stream_prime");
    if (nframes < 1 || nframes > e->batch) return fail(B200_EINVAL, "nframes %d outside 1..%d", nframes, e->batch);
    if (e->blk_submitted - e->blk_waited >= (uint64_t)e->blk_depth)
        return fail(B200_ESTATE, "%d blocks already in flight (the hop ring holds no more): call b200_wait_block", e->blk_depth);
    CU(cudaSetDevice(e->device));
    const int slot = (int)(e->blk_submitted % (uint64_t)e->blk_depth);
    const size_t hb = e->hop_samples * e->format_bytes();
    char *ring = reinterpret_cast<char *>(e->d_ring);
    // the ring slots this block overwrites were last read by the forward pass of the block blk_depth submissions ago
    if (e->blk_submitted >= (uint64_t)e->blk_depth) CU(cudaStreamWaitEvent(e->copy_stream, e->ev_ring_free[slot], 0));
    const long hop0 = e->stream_head;
    // one copy per run of halves that are contiguous in ONE b200_malloc buffer and in the ring (a caller that keeps a
    // block's halves in one buffer pays one copy set-up instead of nframes: measured 2.30 -> 1.69 ms per block of 64 u8
    // halves of 1 MiB). Separate allocations that merely happen to be adjacent are never merged.
    for (int f = 0; f < nframes;) {
        const long dst = (hop0 + 1 + f) % (long)e->nhops;
        const char *src = static_cast<const char *>(new_halves[f]);
        const char *alloc_end = src + hb;
        {
            auto it = e->host_allocs.upper_bound(src);
            if (it != e->host_allocs.begin()) {
                --it;
                if (src >= it->first && src + hb <= it->first + it->second) alloc_end = it->first + it->second;
            }
        }
        int run = 1;
        while (f + run < nframes && dst + run < (long)e->nhops && static_cast<const char *>(new_halves[f + run]) == src + (size_t)run * hb &&
               src + (size_t)(run + 1) * hb <= alloc_end)
            run++;
        CU(cudaMemcpyAsync(ring + (size_t)dst * hb, new_halves[f], hb * run, cudaMemcpyHostToDevice, e->copy_stream));
        f += run;
    }
    CU(cudaEventRecord(e->ev_in[slot], e->copy_stream));
    e->stream_head = (hop0 + nframes) % (long)e->nhops;
    // forward on the engine stream
    e->cur_bank = (int)(e->blk_submitted % (uint64_t)e->banks);
    CU(cudaStreamWaitEvent(e->stream, e->ev_in[slot], 0));
    e->fwd_frame = frame_num0;
    int rc = run_forward(e, hop0, nframes);
    if (rc) return rc;
    CU(cudaEventRecord(e->ev_ring_free[slot], e->stream));
    if (pyramid_out) {
        // rows of the batch are pyr_stride apart on the device, pyr_bytes apart for the caller; with a waterfall cadence only
        // the rows of the send frames (frame number a multiple of wf_skip) are produced and copied, the others stay untouched
        const int skip = std::max(1, e->wf_skip);
        const int first = (int)((skip - frame_num0 % skip) % skip);
        const int count = first < nframes ? (nframes - first + skip - 1) / skip : 0;
        if (count > 0)
            CU(cudaMemcpy2DAsync(pyramid_out + (size_t)first * e->pyr_bytes, (size_t)skip * e->pyr_bytes,
                                 e->quant_ptr() + (size_t)first * e->pyr_stride, (size_t)skip * e->pyr_stride, e->pyr_bytes, count,
                                 cudaMemcpyDeviceToHost, e->stream));
    }
    cudaStream_t cs = e->client_stream();
    if (e->have_clients) {
        rc = run_clients(e, frame_num0, nframes);
        if (rc) return rc;
        const size_t mc = e->ca.max_clients, h = e->ca.h;
        const ClientArrays cab = e->client_arrays(e->last_buf);
        if (pwr_out) CU(cudaMemcpyAsync(pwr_out, e->ca.pwr, sizeof(float) * mc * nframes, cudaMemcpyDeviceToHost, cs));
        if (e->tail_async()) {
            // PCM and validity leave on their own stream once the tails of THIS block are done, so that the tails of the
            // next block (other buffer) do not wait for the copy
            cs = e->d2h_stream;
            CU(cudaStreamWaitEvent(cs, e->ev_tail[e->last_buf], 0));
            CU(cudaEventRecord(e->ev_join, e->client_stream()));
            CU(cudaStreamWaitEvent(cs, e->ev_join, 0));
        }
        if (pcm_out)
            CU(cudaMemcpyAsync(pcm_out, cab.pcm, (e->opt_pcm16 ? sizeof(int16_t) : sizeof(int32_t)) * mc * h * nframes, cudaMemcpyDeviceToHost, cs));
        if (valid_out) CU(cudaMemcpyAsync(valid_out, cab.valid, mc * nframes, cudaMemcpyDeviceToHost, cs));
        if (e->tail_async()) {  // the buffer is free again only when the copy has read it
            CU(cudaEventRecord(e->ev_tail[e->last_buf], cs));
        }
    }
    // block completion = client results on the host and (engine stream) the pyramid on the host
    CU(cudaEventRecord(e->ev_out[slot], e->stream));
    CU(cudaStreamWaitEvent(cs, e->ev_out[slot], 0));
    CU(cudaEventRecord(e->ev_out[slot], cs));
    e->blk_pending[slot] = true;
    e->blk_submitted++;
    return 0;
}

int b200_wait_block(b200_engine *e) {
    if (!e) return fail(B200_EINVAL, "null engine");
    if (e->blk_waited >= e->blk_submitted) return fail(B200_ESTATE, "no block in flight");
    CU(cudaSetDevice(e->device));
    const int slot = (int)(e->blk_waited % (uint64_t)e->blk_depth);
    CU(cudaEventSynchronize(e->ev_out[slot]));
    e->blk_pending[slot] = false;
    e->blk_waited++;
    return stream_check(e);
}

uint64_t b200_launch_count(b200_engine *e) { return e ? e->launches : 0; }

int b200_quant_table(int power_offset, uint32_t *lo, uint32_t *hi, uint8_t *base) {
    if (!lo || !hi || !base) return fail(B200_EINVAL, "null argument");
    build_quant_table(power_offset, lo, hi, base);
    return 0;
}

int b200_debug_tail_profile(b200_engine *e, int enable, long long out[64]) {
    if (!e) return fail(B200_EINVAL, "null engine");
    CU(cudaSetDevice(e->device));
    if (e->d_prof && out) {
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(out, e->d_prof, sizeof(long long) * 64, cudaMemcpyDeviceToHost));
    }
    if (enable && !e->d_prof) {
        CU(cudaMalloc(&e->d_prof, sizeof(long long) * 64));
    }
    if (e->d_prof) CU(cudaMemset(e->d_prof, 0, sizeof(long long) * 64));
    if (!enable && e->d_prof) {
        cudaFree(e->d_prof);
        e->d_prof = nullptr;
    }
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
