// Shared helpers for libosb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/osb200.h"

#define OSB_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace osb {

// thread-local error text (osb_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int ensure_init();          // lazy osb_init(current device)
int num_sms();              // queried once; 148 on B200
// One-shot hint of the calling thread for the NEXT persistent launches that ask: "only this many SMs are free for your first wave" (a
// concurrently running kernel holds the others).  A persistent kernel whose CTAs take a static share of the tiles must not launch more
// CTAs than can be resident at once: the ones that wait start when the first ones END, and the kernel takes up to twice as long.
void set_sm_budget(int sms, int launches = 1);  // the hint holds for the next `launches` persistent launches that ask; 0 clears
int take_sm_budget();                           // the hint (one use taken), or num_sms() when none is set
void count_launch();        // osb_launch_count bookkeeping
bool prof_on();             // osb_profile_enable: per-kernel CUDA-event timing on the launching stream
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);

#define OSB_LAUNCH(kern, grid, block, smem, stream, ...)              \
    do {                                                              \
        const bool _p = osb::prof_on();                               \
        if (_p) osb::prof_begin(#kern, (stream));                     \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);     \
        if (_p) osb::prof_end((stream));                              \
        osb::count_launch();                                          \
    } while (0)

#define OSB_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return osb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define OSB_CHECK_LAUNCH() OSB_CUDA(cudaGetLastError())

#define OSB_REQUIRE(cond, msg)                         \
    do {                                               \
        if (!(cond)) {                                 \
            osb::set_error("invalid argument: %s", msg); \
            return OSB_ERR_INVALID_ARG;                \
        }                                              \
    } while (0)

// Stream-ordered scratch allocation (cudaMallocAsync on the caller's stream; the
// default mempool keeps freed blocks cached, so steady-state calls do not hit the driver).
struct Scratch {
    cudaStream_t s;
    void* ptrs[24];
    int n = 0;
    explicit Scratch(cudaStream_t st) : s(st) {}
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], s);
    }
    template <typename T>
    cudaError_t alloc(T** p, size_t count) {
        if (n >= (int)(sizeof(ptrs) / sizeof(ptrs[0]))) return cudaErrorMemoryAllocation;
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, count * sizeof(T) + 256, s);
        if (e == cudaSuccess) {
            ptrs[n++] = q;
            *p = reinterpret_cast<T*>(q);
        }
        return e;
    }
};

// Function attributes (dynamic shared memory opt-in) belong to the CONTEXT: a process that drives several devices has to set them once
// per device, and a failure must not be remembered for ever.  Usage: static PerDeviceOnce once; OSB_CUDA(once.run([&] { return ...; }));
struct PerDeviceOnce {
    std::mutex mu;
    unsigned long long done = 0;  // bit d: device d is set up (up to 64 devices per process)
    template <typename F>
    cudaError_t run(F&& f) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
        e = f();
        if (e == cudaSuccess && dev < 64) done |= 1ull << dev;
        return e;
    }
};

// Host-call workspace: per-thread stream + pinned staging + device buffers.  The
// *_host entry points copy in, launch the *_dev path, copy out and synchronise.
struct HostWs {
    cudaStream_t stream = nullptr;
    void* pin[2] = {nullptr, nullptr};
    size_t pin_cap[2] = {0, 0};
    void* dev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t dev_cap[4] = {0, 0, 0, 0};
    int device = -1;
    int prepare();
    int dev_buf(int slot, size_t bytes, void** out);
    int pin_buf(int slot, size_t bytes, void** out);
    int h2d(void* d, const void* h, size_t bytes);   // through pinned staging when it fits
    int d2h(void* h, const void* d, size_t bytes);   // synchronises the stream
    int sync();
};
HostWs& host_ws();

// cross-file launchers (logmel.cu, spectral_gate.cu, gain.cu)
int launch_logmel(const void* d_audio, int fmt, long long n, long long batch, long long stride, int n_mels, float* d_out,
                  const unsigned long long* d_sumsq, float target_dbfs, cudaStream_t st, const double* d_sumsq_f = nullptr,
                  int requant = 0);
// d_sumsq (optional, zeroed by the caller): per-clip sum of squares of the float32 output, for a following normalize_gain
int launch_spectral_gate(const void* d_audio, int fmt, long long n, long long batch, long long stride, int sr, float* d_out,
                         cudaStream_t st, double* d_sumsq = nullptr);
int launch_pitch_shift(const void* d_in, bool in_f64, const long long* d_offsets, const long long* d_lens, long long batch, long long max_len,
                       int sample_rate, double semitones, float* d_out, cudaStream_t st);
int launch_sumsq_pcm16(const int16_t* d_in, long long n, long long batch, long long stride, unsigned long long* d_sumsq, cudaStream_t st);
int launch_normalize_f32(const float* d_in, void* d_out, int out_pcm16, long long n, long long batch, long long stride, int normalize,
                         float target_dbfs, cudaStream_t st);

// vad.cu: batched scoring (state [batch][2][128] in/out) and the integer segmenter, for the composed paths
// front_done (optional): recorded on st after the first chunk's front kernel(s); shared_gpu: another branch runs beside the recurrence
int vad_recurrence_sms(void* handle, long long batch);
int launch_vad_score(void* handle, const void* d_audio, int fmt, long long n, long long batch, long long stride, float* d_state,
                     float* d_probs, long long probs_stride, cudaStream_t st, cudaEvent_t front_done = nullptr, bool shared_gpu = false);
int launch_vad_segments(const float* d_probs, long long probs_stride, long long n_win, long long batch, long long n_samples, float thr,
                        int min_speech_ms, int silence_ms, int32_t* d_segs, int32_t* d_counts, int max_seg, cudaStream_t st);

static inline int grid_for(size_t work_items, int per_block, int max_waves = 8) {
    size_t b = (work_items + per_block - 1) / per_block;
    size_t cap = (size_t)OSB_NUM_SMS * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace osb

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
namespace osb {

__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    const uint4 r = ld_stream_u4(p);
    return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS UBLKCP): global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// clip to [-1,1], *32767, truncate toward zero  (src/audio/preprocessing.py:24-25,
// src/tts/pipeline.py:32-37).  NaN -> 0 like a saturating cvt; the reference never feeds NaN.
__device__ __forceinline__ int quant_pcm16(float x) {
    x = fminf(fmaxf(x, -1.0f), 1.0f);
    return __float2int_rz(__fmul_rn(x, 32767.0f));
}

// gain exactly as normalize_gain forms it, in f32:  rms -> 20*log10 -> target - cur -> /20 -> 10**x
// returns 1.0 and *silent=true when rms <= 1e-8 (the reference returns its input unchanged).
__device__ __forceinline__ float gain_from_meansq(double mean_sq, float target_dbfs, bool* silent) {
    float rms = __fsqrt_rn((float)mean_sq);
    if (rms <= 1e-8f) { *silent = true; return 1.0f; }
    *silent = false;
    float cur = __fmul_rn(20.0f, log10f(rms));
    float gdb = __fsub_rn(target_dbfs, cur);
    return powf(10.0f, __fdiv_rn(gdb, 20.0f));
}

__device__ __forceinline__ float apply_gain(float x, float gain, bool silent) {
    if (silent) return x;  // unchanged, NOT clipped (preprocessing.py:37-38)
    return fminf(fmaxf(__fmul_rn(x, gain), -1.0f), 1.0f);
}


}  // namespace osb
#endif
