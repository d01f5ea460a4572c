// Silero-VAD front end as ONE persistent tcgen05 kernel: a tile of 128 windows is carried from the pcm16 samples to the
// LSTM gate pre-activations on chip -- DFT conv -> |.| -> 4 x (Conv1d k=3 + ReLU) -> W_ih -- with the activations in shared
// memory (as the next layer's MMA operand), the accumulators in tensor memory and the weights streaming through a TMA ring.
//
// Replaces the stateless part of (reference file:line) SileroVAD.__call__ / get_speech_segments, src/vad/silero.py:63-91,
// :109-177: the per-window onnxruntime session.run at :86 / :149 up to the LSTM cell.
//
// Arithmetic: every GEMM is split-bf16, D += A_hi.B_hi + A_hi.B_lo + A_lo.B_hi with FP32 accumulation in TMEM (operand
// precision ~2^-16; tools/vad_precision_sim.py: 2.6e-5 on the probabilities, budget 1e-3; one- and two-term schemes do not
// fit the budget).  pcm16 samples are EXACT in two bf16 planes.
//
// Convolutions over the 3 / 2 / 1 positions of a window are sums of per-tap GEMMs accumulated in TMEM, issued as soon as the
// input position exists ("accumulate as available"): no im2col, no zero-padded taps (7 of 9 enc1 taps, 4 of 6 enc2, 2 of 3
// enc3, 1 of 3 enc4 are real), and a position's activations live in shared memory only until their last tap is issued.
// The 129th |STFT| channel (Nyquist) does not fit the 128-wide K tiling: it takes the imaginary-DC column of the DFT GEMM
// (identically zero) and enters enc1 as a rank-1 FP32 update in the epilogue.
//
// Warp roles (576 threads, one CTA per SM, persistent over tiles):
//   warps 0-7   epilogue: TMEM -> registers -> bias / ReLU / |re,im| -> bf16 hi+lo -> swizzled smem operand (or global store)
//   warps 8-15  staging: pcm16 / f32 samples -> bf16 hi+lo -> swizzled smem operand (two 64-sample chunk buffers in flight)
//   warp 16     one lane issues every tcgen05.mma and the commits that release buffers / publish accumulators
//   warp 17     one lane keeps the weight ring full (cp.async.bulk, 32 KB slots)
// Per tile the schedule is static: 57 MMA groups of K = 64 (12 tcgen05.mma each), one weight slot per group; every shared resource
// (4 operand buffers, 4 TMEM slots of 128 columns) alternates strictly full -> free, each transition on its own mbarrier.
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"
#include "vad_front.cuh"

namespace osb {

namespace vf {

constexpr int kRows = 128;                      // windows per tile = MMA M
constexpr int kKc = 64;                         // K per operand chunk: one 128-byte swizzle row of bf16
constexpr int kPlane = kRows * kKc * 2;         // 16 KB
constexpr int kBuf = 2 * kPlane;                // hi + lo: 32 KB
constexpr int kRing = 3;                        // weight ring depth
constexpr int kSlotsPerTile = 57;
constexpr int kEpiThreads = 256, kStageThreads = 256;
constexpr int kThreads = kEpiThreads + kStageThreads + 64;
// dynamic shared memory: [4 operand buffers][kRing weight slots][side |re128| 3 x 128 f32][barriers][tmem ptr]
constexpr int kOffRing = 4 * kBuf;
constexpr int kOffSide = kOffRing + kRing * kBuf;
constexpr int kOffBar = kOffSide + 3 * kRows * 4;
constexpr int kNumBar = 4 + 4 + 4 + 4 + 2 * kRing;  // pfull, pfree, tfull, tfree, bfull, bfree
constexpr int kOffTmem = kOffBar + kNumBar * 8;
constexpr int kSmem = kOffTmem + 16 + 1024;          // + slack to align the base to 1024 B

// the four operand buffers.  Writes per tile in schedule order -- A0: 12 audio chunks + h1_0[0:64] + h1_2[0:64] + h3 = 15;
// A1: 12 audio chunks + h1_0[64:128] + h1_2[64:128] = 14; M0 / M1: |STFT| chunks of three units + h1_1 half + h2_q + h4 half = 6 each.
enum { A0 = 0, A1 = 1, M0 = 2, M1 = 3 };

struct Consts {
    float e1b[128], e2b[64], e3b[64], e4b[128];
    float bsum[512];            // b_ih + b_hh
    float w1side[3][128];       // enc1 weights of input channel 128 (the Nyquist bin), per tap
};

struct Params {
    const void* audio;          // pcm16 or f32
    int fmt;                    // OSB_FMT_PCM16 | OSB_FMT_F32
    long long audio_stride;     // samples between streams
    int wins_per_stream;        // windows of this chunk per stream (T)
    long long win0;             // first window of the chunk
    long long total;            // windows in this launch = streams * T
    int n_tiles;                // end of the tile range of this launch
    int tile0;                  // first tile of this launch (the chunk's tiles may be split over two launches, vad.cu)
    const uint8_t* wimg;        // weight image
    const uint2* slots;         // [57] (byte offset, bytes) in consumption order
    float* pre;                 // [windows / 8][64 column groups][8][8]: W_ih.x + b_ih + b_hh, interleaved (vad.cu pre_at)
    // per-channel vectors the epilogue needs, in the kernel's constant bank: every lane of a warp reads the same element (a warp holds
    // 32 rows of the same columns), which the constant cache serves as a broadcast; with 227 KB of shared memory carved out there is
    // no L1 left for them and the L2 round trip (~700 cycles per dependent load) was the epilogue's whole cost
    Consts c;
    unsigned long long* trace;  // debug: clock stamps of CTA 0 (OSB_VF_TRACE), or null
};

// debug trace: role r (0 MMA, 1 epilogue, 2 staging, 3 producer), tile iteration it (< 4), event e (< 64) of CTA 0
__device__ __forceinline__ void stamp(const Params& p, int r, int it, int e) {
    if (p.trace && blockIdx.x == 0 && it < 4 && e < 64) p.trace[(r * 4 + it) * 64 + e] = clock64();
}

__device__ __forceinline__ void stamp_val(const Params& p, int r, int it, int e, unsigned long long v) {
    if (p.trace && blockIdx.x == 0 && it < 4 && e < 64) p.trace[(r * 4 + it) * 64 + e] = v;
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
// sqrt.approx (2 ulp): the magnitude is split into bf16 hi + lo (2^-16) right after, an IEEE square root would cost ~25 instructions more
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
// tcgen05.ld 32x32b: lane i of the warp receives N consecutive 32-bit columns of TMEM lane (warp % 4) * 32 + i
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 8 consecutive K elements (one 16-byte unit) of row r, unit index u (0..7) inside a [128][64] bf16 SW128 K-major plane
__device__ __forceinline__ uint32_t unit_off(int r, int u) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)); }

// Shared bookkeeping of one role: global write counts per operand buffer (all roles advance them identically) and the flip bits of
// the barriers this role alone waits on.
struct Book {
    uint32_t wpar;        // bit b: parity of the number of writes of operand buffer b so far
    uint32_t wany;        // bit b: buffer b has been written at least once
    uint32_t ph;          // flip bits: MMA: pfull[0..3] bits 0-3, tfree[0..3] bits 4-7 ; epilogue: tfull[0..3] bits 0-3
    uint32_t tany;        // bit t: an accumulation into TMEM slot t has been started before (MMA role)
    __device__ __forceinline__ void other_writes(int b, int k) {  // k writes of buffer b by another role
        wpar ^= (uint32_t)(k & 1) << b;
        wany |= 1u << b;
    }
};

struct Sm {
    uint8_t* base;
    __device__ __forceinline__ uint8_t* buf(int i) const { return base + (size_t)i * kBuf; }
    uint8_t* ring;
    float* side;
    uint64_t *pfull, *pfree, *tfull, *tfree, *bfull, *bfree;
    uint32_t tmem;
};

// ------------------------------------------------------------------ role: MMA issuer (one thread)
struct MmaRole {
    const Sm sm;
    Book bk;
    uint32_t slot = 0;    // weight slots consumed so far (ring position = slot % kRing)
    const Params& p;
    int it = 0;
    __device__ __forceinline__ MmaRole(const Sm& s, const Book& b, const Params& pp) : sm(s), bk(b), p(pp) {}
    long long cyc_p = 0, cyc_b = 0, cyc_t = 0;  // debug: cycles spent waiting for operands / weights / TMEM
    __device__ __forceinline__ void wait_full(int buf) {
        const long long c0 = p.trace ? clock64() : 0;
        mbar_wait(&sm.pfull[buf], (bk.ph >> buf) & 1u);
        bk.ph ^= 1u << buf;
        fence_after();
        if (p.trace) cyc_p += clock64() - c0;
    }
    __device__ __forceinline__ void release(int buf) { commit(&sm.pfree[buf]); }
    __device__ __forceinline__ void acc_begin(int t) {  // before the first MMA of a new accumulation into TMEM slot t
        if ((bk.tany >> t) & 1u) {
            const long long c0 = p.trace ? clock64() : 0;
            mbar_wait(&sm.tfree[t], (bk.ph >> (4 + t)) & 1u);
            bk.ph ^= 1u << (4 + t);
            fence_after();
            if (p.trace) cyc_t += clock64() - c0;
        }
        bk.tany |= 1u << t;
    }
    __device__ __forceinline__ void acc_done(int t) { commit(&sm.tfull[t]); }
    // one group: A = operand buffer `buf` (K = 64), B = next weight slot (rows = N), D = TMEM column `col`, N columns
    __device__ __forceinline__ void group(int buf, uint32_t col, int N, bool accumulate) {
        const uint32_t rs = slot % kRing;
        const long long c0 = p.trace ? clock64() : 0;
        mbar_wait(&sm.bfull[rs], (slot / kRing) & 1u);
        fence_after();
        if (p.trace) cyc_b += clock64() - c0;
        const uint32_t a_hi = smem_u32(sm.buf(buf)), a_lo = a_hi + kPlane;
        const uint32_t b_hi = smem_u32(sm.ring + (size_t)rs * kBuf), b_lo = b_hi + (uint32_t)N * 128u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
        const uint32_t d = sm.tmem + col;
#pragma unroll
        for (int k = 0; k < kKc / 16; ++k) {
            const uint32_t ko = k * 32;
            mma_bf16(d, desc_sw128(a_hi + ko), desc_sw128(b_hi + ko), idesc, (accumulate || k) ? 1u : 0u);
            mma_bf16(d, desc_sw128(a_hi + ko), desc_sw128(b_lo + ko), idesc, 1u);
            mma_bf16(d, desc_sw128(a_lo + ko), desc_sw128(b_hi + ko), idesc, 1u);
        }
        commit(&sm.bfree[rs]);
        ++slot;
    }
    __device__ __forceinline__ void run_tile() {
        stamp(p, 0, it, 0);
        // ---- DFT conv + enc1, interleaved: unit u = (frame f, bin half h)
#pragma unroll 1
        for (int u = 0; u < 7; ++u) {
            stamp(p, 0, it, 1 + 2 * u);
            if (u < 6) {
                acc_begin(0);
#pragma unroll 1
                for (int kc = 0; kc < 4; ++kc) {
                    const int buf = kc & 1;  // audio chunks alternate A0 / A1
                    wait_full(buf);
                    group(buf, 0, 128, kc > 0);
                    release(buf);
                }
                acc_done(0);
            }
            stamp(p, 0, it, 2 + 2 * u);
            if (u >= 1) {  // enc1 contributions of |STFT| chunk (f, h) of unit u-1: positions p = f-1, f, f+1
                const int v = u - 1, f = v >> 1, h = v & 1, buf = M0 + (v & 1);
                wait_full(buf);
#pragma unroll 1
                for (int p = (f > 0 ? f - 1 : 0); p <= (f < 2 ? f + 1 : 2); ++p) {
                    const bool first = (h == 0) && (f == (p > 0 ? p - 1 : 0));
                    if (first) acc_begin(1 + p);
                    group(buf, 128u * (1 + p), 128, !first);
                }
                release(buf);
                if (v == 3) acc_done(1);
                if (v == 5) { acc_done(2); acc_done(3); }
            }
        }
        // ---- enc2 (stride 2): q0 <- h1_0 tap1, h1_1 tap2 ; q1 <- h1_1 tap0, h1_2 tap1.  ACC2 = TMEM slot 0, 64 + 64 columns
        stamp(p, 0, it, 15);
        acc_begin(0);
        wait_full(A0); wait_full(A1);
        group(A0, 0, 64, false); group(A1, 0, 64, true);
        release(A0); release(A1);
        wait_full(M0); wait_full(M1);
        group(M0, 0, 64, true); group(M1, 0, 64, true);
        group(M0, 64, 64, false); group(M1, 64, 64, true);
        release(M0); release(M1);
        wait_full(A0); wait_full(A1);
        group(A0, 64, 64, true); group(A1, 64, 64, true);
        release(A0); release(A1);
        acc_done(0);
        // ---- enc3 (stride 2): out <- h2_0 tap1, h2_1 tap2.  ACC3 = slot 1, 64 columns
        stamp(p, 0, it, 16);
        acc_begin(1);
        wait_full(M0); wait_full(M1);
        group(M0, 128, 64, false); group(M1, 128, 64, true);
        release(M0); release(M1);
        acc_done(1);
        // ---- enc4: out <- h3 tap1.  ACC4 = slot 2
        stamp(p, 0, it, 17);
        acc_begin(2);
        wait_full(A0);
        group(A0, 256, 128, false);
        release(A0);
        acc_done(2);
        // ---- W_ih: four 128-wide gate blocks, K = 128
        stamp(p, 0, it, 18);
        wait_full(M0); wait_full(M1);
        stamp(p, 0, it, 19);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            acc_begin(j);
            group(M0, 128u * j, 128, false);
            group(M1, 128u * j, 128, true);
            acc_done(j);
        }
        release(M0); release(M1);
        stamp(p, 0, it, 20);
        stamp_val(p, 0, it, 40, (unsigned long long)cyc_p); stamp_val(p, 0, it, 41, (unsigned long long)cyc_b); stamp_val(p, 0, it, 42, (unsigned long long)cyc_t);
        cyc_p = cyc_b = cyc_t = 0;
        ++it;
    }
};

// ------------------------------------------------------------------ role: weight ring producer (one thread)
__device__ __forceinline__ void producer_loop(const Sm& sm, const Params& p, int n_my_tiles) {
    uint32_t slot = 0;
    for (int t = 0; t < n_my_tiles; ++t)
        for (int i = 0; i < kSlotsPerTile; ++i, ++slot) {
            const uint32_t rs = slot % kRing;
            if (slot >= kRing) mbar_wait(&sm.bfree[rs], ((slot / kRing) - 1) & 1u);
            const uint2 e = __ldg(p.slots + i);
            mbar_expect_tx(&sm.bfull[rs], e.y);
            bulk_g2s(sm.ring + (size_t)rs * kBuf, p.wimg + e.x, e.y, &sm.bfull[rs]);
        }
}

// ------------------------------------------------------------------ writer-side bookkeeping shared by the staging and epilogue roles
__device__ __forceinline__ void writer_acquire(const Sm& sm, Book& bk, int buf, bool mine) {
    // before write number J (0-based, counted over ALL writers of this buffer): the readers of write J-1 must have finished
    const uint32_t jpar = (bk.wpar >> buf) & 1u, prev = (bk.wany >> buf) & 1u;
    bk.wpar ^= 1u << buf;
    bk.wany |= 1u << buf;
    if (mine && prev) mbar_wait(&sm.pfree[buf], jpar ^ 1u);
}
template <int COUNT = 1>
__device__ __forceinline__ void writer_publish(const Sm& sm, int buf) {
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&sm.pfull[buf])), "n"(COUNT) : "memory");
}

// ------------------------------------------------------------------ role: audio staging (256 threads)
struct StageRole {
    const Sm sm;
    Book bk;
    const Params& p;
    int t;  // 0..255
    int it = 0;
    __device__ __forceinline__ StageRole(const Sm& s, const Book& b, const Params& pp, int tid) : sm(s), bk(b), p(pp), t(tid) {}
    // One 64-sample chunk = 128 rows x 8 operand units (8 samples = 16 bytes of pcm16 each).  Thread t takes the four units
    // q = t + 256 k (row q / 8, unit q % 8): the eight lanes of a quarter-warp read one row's 128 contiguous bytes, a warp-wide load
    // touches 4 lines instead of 32 (one line per lane cost 32 L1 wavefronts per load and made this role the kernel's bottleneck),
    // and the matching 16-byte shared-memory stores of a quarter-warp fall into eight different bank groups (XOR swizzle).
    // The pcm16 loads of chunk c+1 are issued BEFORE chunk c is converted (two register sets); float32 input loads inside the conversion.
    // int16 -> float without the quarter-rate I2F: as_float(0x4B400000 + v) = 12582912 + v exactly, and one FFMA rescales it to v / 32768.
    long long cyc_acq = 0, cyc_cvt = 0, cyc_u0 = 0, cyc_pub = 0;
    long long rbase[4];  // per tile: sample offset of window row (t / 8 + 32 k) in the audio buffer, -1 past the end
    __device__ __forceinline__ void tile_rows(long long tile) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long wi = tile * kRows + (t >> 3) + 32 * k;
            if (wi < p.total) {
                const long long sidx = wi / p.wins_per_stream;
                rbase[k] = sidx * p.audio_stride + (p.win0 + (wi - sidx * p.wins_per_stream)) * 512 + 8 * (t & 7);
            } else rbase[k] = -1;
        }
    }
    __device__ __forceinline__ static int chunk_off(int c) { return 128 * (c >> 3) + 64 * (c & 3); }  // c = 4 u + kc, u = 2 f + h
    __device__ __forceinline__ void load_pcm(int c, uint4 (&w)[4]) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (rbase[k] < 0) { w[k] = make_uint4(0u, 0u, 0u, 0u); continue; }
            const int16_t* s16 = reinterpret_cast<const int16_t*>(p.audio) + rbase[k] + chunk_off(c);
            if ((((uintptr_t)s16) & 15) == 0) w[k] = ld_stream_u4(s16);
            else {
                uint32_t q[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = (uint32_t)(uint16_t)__ldg(s16 + 2 * i) | ((uint32_t)(uint16_t)__ldg(s16 + 2 * i + 1) << 16);
                w[k] = make_uint4(q[0], q[1], q[2], q[3]);
            }
        }
    }
    __device__ __forceinline__ void put_unit(int buf, int k, const float (&v)[8]) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
        uint8_t* hi = sm.buf(buf) + unit_off((t >> 3) + 32 * k, t & 7);
        *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(hi + kPlane) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    __device__ __forceinline__ void acquire(int buf, int c) {
        const long long c0 = p.trace ? clock64() : 0;
        writer_acquire(sm, bk, buf, true);
        if (p.trace) cyc_acq += clock64() - c0;
        if (t == 0) stamp(p, 2, it, 2 * c);
    }
    __device__ __forceinline__ void publish(int buf, int c) {
        writer_publish(sm, buf);
        if (t == 0) stamp(p, 2, it, 2 * c + 1);
    }
    __device__ __forceinline__ void convert_pcm(int buf, int c, const uint4 (&w)[4]) {
        acquire(buf, c);
        const long long c1 = p.trace ? clock64() : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t ws[4] = {w[k].x, w[k].y, w[k].z, w[k].w};
            float v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s0 = (int)(short)(ws[i] & 0xFFFFu), s1 = (int)ws[i] >> 16;
                v[2 * i] = fmaf(__int_as_float(0x4B400000 + s0), 3.0517578125e-05f, -384.0f);      // (12582912 + s) / 32768 - 384
                v[2 * i + 1] = fmaf(__int_as_float(0x4B400000 + s1), 3.0517578125e-05f, -384.0f);
            }
            put_unit(buf, k, v);
            if (p.trace && k == 0) cyc_u0 += clock64() - c1;
        }
        if (p.trace) cyc_cvt += clock64() - c1;
        const long long c2 = p.trace ? clock64() : 0;
        publish(buf, c);
        if (p.trace) cyc_pub += clock64() - c2;
    }
    __device__ __forceinline__ void convert_f32(int buf, int c) {
        float4 g[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (rbase[k] < 0) { g[2 * k] = g[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
            const float* sf = reinterpret_cast<const float*>(p.audio) + rbase[k] + chunk_off(c);
            if ((((uintptr_t)sf) & 15) == 0) {
                g[2 * k] = ld_stream_f4(sf);
                g[2 * k + 1] = ld_stream_f4(sf + 4);
            } else {
                g[2 * k] = make_float4(__ldg(sf), __ldg(sf + 1), __ldg(sf + 2), __ldg(sf + 3));
                g[2 * k + 1] = make_float4(__ldg(sf + 4), __ldg(sf + 5), __ldg(sf + 6), __ldg(sf + 7));
            }
        }
        acquire(buf, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v[8] = {g[2 * k].x, g[2 * k].y, g[2 * k].z, g[2 * k].w, g[2 * k + 1].x, g[2 * k + 1].y, g[2 * k + 1].z, g[2 * k + 1].w};
            put_unit(buf, k, v);
        }
        publish(buf, c);
    }
    // all tiles of this CTA: 24 chunks per tile, chunk c -> buffer A(c & 1)
    __device__ __forceinline__ void run(int n_my) {
        const bool pcm = p.fmt == OSB_FMT_PCM16;
        uint4 wa[4], wb[4];
        if (n_my > 0) tile_rows((long long)p.tile0 + blockIdx.x);
        if (pcm && n_my > 0) load_pcm(0, wa);
        for (int i = 0; i < n_my; ++i) {
            const long long tile = (long long)p.tile0 + blockIdx.x + (long long)i * gridDim.x;
            if (pcm) {
#pragma unroll 1
                for (int c = 0; c < 24; c += 2) {
                    load_pcm(c + 1, wb);
                    convert_pcm(A0, c, wa);
                    if (c + 2 < 24) load_pcm(c + 2, wa);
                    else if (i + 1 < n_my) { tile_rows(tile + gridDim.x); load_pcm(0, wa); }
                    convert_pcm(A1, c + 1, wb);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < 24; ++c) convert_f32(c & 1, c);
                if (i + 1 < n_my) tile_rows(tile + gridDim.x);
            }
            // The epilogue's writes of these two buffers (h1_0, h1_2, h3) come next.  A parity wait can only tell "the phase I mean" from
            // "the one before it", so this role must not run more than one phase ahead: it waits through the release of every one of those
            // writes, in schedule order, before it may ask for the buffers again (next tile).  Pure waits: nothing is written or published.
            writer_acquire(sm, bk, A0, true); writer_acquire(sm, bk, A1, true);   // h1_0
            writer_acquire(sm, bk, A0, true); writer_acquire(sm, bk, A1, true);   // h1_2
            writer_acquire(sm, bk, A0, true);                                     // h3
            if (t == 0) { stamp(p, 2, it, 48); stamp_val(p, 2, it, 50, (unsigned long long)cyc_acq); stamp_val(p, 2, it, 51, (unsigned long long)cyc_cvt); stamp_val(p, 2, it, 52, (unsigned long long)cyc_u0); stamp_val(p, 2, it, 53, (unsigned long long)cyc_pub); }
            cyc_acq = cyc_cvt = cyc_u0 = cyc_pub = 0;
            ++it;
        }
    }
};

// ------------------------------------------------------------------ role: epilogue (256 threads = warps 0-7)
struct EpiRole {
    const Sm sm;
    Book bk;
    const Params& p;
    int warp, lane, row, half;  // row = TMEM lane of this thread, half = which 64 columns of a 128-column slot
    uint32_t lane_base;
    __device__ __forceinline__ EpiRole(const Sm& s, const Book& b, const Params& pp, int tid) : sm(s), bk(b), p(pp) {
        warp = tid >> 5; lane = tid & 31;
        row = (warp & 3) * 32 + lane;
        half = warp >> 2;
        lane_base = (uint32_t)((warp & 3) * 32) << 16;
    }
    int it = 0, ev = 0;
    __device__ __forceinline__ void wait_acc(int t) {
        if (row == 0 && half == 0) stamp(p, 1, it, ev++);
        mbar_wait(&sm.tfull[t], (bk.ph >> t) & 1u);
        bk.ph ^= 1u << t;
        fence_after();
        if (row == 0 && half == 0) stamp(p, 1, it, ev++);
    }
    __device__ __forceinline__ void free_acc(int t) {
        fence_before();
        arrive(&sm.tfree[t]);
    }
    // 8 activations = one 16-byte unit (channels 8 unit .. 8 unit + 7 of a 64-channel chunk) of this thread's row -> hi / lo planes of `buf`
    __device__ __forceinline__ void put8(int buf, int unit, const float (&x)[8]) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(x[2 * j], x[2 * j + 1], h[j], l[j]);
        uint8_t* hi = sm.buf(buf) + unit_off(row, unit);
        *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(hi + kPlane) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    // |STFT| of unit (f, h): TMEM slot 0 holds (re, im) pairs of bins 64h .. 64h+63 -> 64 magnitudes = chunk h of position f.
    // This thread: columns 64 half .. +63 = 32 bins = channels 32 half .. +31 of the chunk, 16 columns (one operand unit) at a time.
    __device__ __forceinline__ void mags(int f, int h, int buf) {
        wait_acc(0);
        writer_acquire(sm, bk, buf, true);
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            uint32_t v[16];
            tmem_ld16(sm.tmem + lane_base + (uint32_t)(64 * half + 16 * u), v);
            if (u == 3) free_acc(0);
            float m[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                m[i] = sqrt_approx(re * re + im * im);
            }
            if (u == 0 && h == 0 && half == 0) {  // columns 0, 1 of the first half are (re 0, re 128): two real bins, not a pair
                sm.side[f * kRows + row] = fabsf(__uint_as_float(v[1]));
                m[0] = fabsf(__uint_as_float(v[0]));
            }
            put8(buf, 4 * half + u, m);
        }
        writer_publish(sm, buf);
    }
    // a 128-column accumulator -> bias (+ Nyquist-channel rank-1 term) + ReLU -> chunk buffers (b_lo: channels 0-63, b_hi: 64-127)
    template <bool SIDE>
    __device__ __forceinline__ void act128(int t, const float* __restrict__ bias, int b_lo, int b_hi, int side_pos) {
        wait_acc(t);
        const uint32_t ta = sm.tmem + lane_base + (uint32_t)(128 * t + 64 * half);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (SIDE) {  // |re 128| of positions side_pos-1, side_pos, side_pos+1 (zero outside the window)
            if (side_pos >= 1) s0 = sm.side[(side_pos - 1) * kRows + row];
            s1 = sm.side[side_pos * kRows + row];
            if (side_pos <= 1) s2 = sm.side[(side_pos + 1) * kRows + row];
        }
        const int buf = half ? b_hi : b_lo;
        // every thread waits for (and later arrives on) both buffers although it writes one: an arrival for write J of a buffer is
        // then always preceded by that thread's own wait for the readers of write J-1
        writer_acquire(sm, bk, b_lo, true);
        writer_acquire(sm, bk, b_hi, true);
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            uint32_t v[8];
            tmem_ld8(ta + 8 * u, v);
            if (u == 7) free_acc(t);
            const int oc = 64 * half + 8 * u;
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = __uint_as_float(v[i]) + bias[oc + i];  // constant bank, warp-uniform index
                if (SIDE) a = fmaf(s0, p.c.w1side[0][oc + i], fmaf(s1, p.c.w1side[1][oc + i], fmaf(s2, p.c.w1side[2][oc + i], a)));
                x[i] = fmaxf(a, 0.f);
            }
            put8(buf, u, x);
        }
        writer_publish(sm, b_lo);  // every epilogue thread arrives on both barriers (count = 256 each); each wrote one of the buffers
        writer_publish(sm, b_hi);
    }
    // 64 columns at `col` of slot t -> bias + ReLU -> one 64-channel chunk buffer (each thread: 32 of the 64 channels)
    __device__ __forceinline__ void act64(int t, int col, const float* __restrict__ bias, int buf, bool last_reader) {
        writer_acquire(sm, bk, buf, true);
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            uint32_t v[8];
            tmem_ld8(sm.tmem + lane_base + (uint32_t)(128 * t + col + 32 * half + 8 * u), v);
            if (u == 3 && last_reader) free_acc(t);
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[i]) + bias[32 * half + 8 * u + i], 0.f);
            put8(buf, 4 * half + u, x);
        }
        writer_publish(sm, buf);
    }
    __device__ __forceinline__ void store_pre(int j, long long tile) {
        wait_acc(j);
        const uint32_t ta = sm.tmem + lane_base + (uint32_t)(128 * j + 64 * half);
        const long long wi = tile * kRows + row;
        // interleaved layout [window / 8][column / 8][window % 8][column % 8] (vad.cu pre_at): lane = row, so eight lanes fill 256
        // contiguous bytes with two store instructions -- with the linear [window][512] layout an instruction wrote 32 half-sectors
        // 2 KB apart and this loop cost 9,000 cycles per block
        float* out = p.pre + (wi >> 3) * 4096 + (long long)(16 * j + 8 * half) * 64 + (wi & 7) * 8;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t v[16];
            tmem_ld16(ta + 16 * u, v);
            if (u == 3) free_acc(j);
            if (wi < p.total) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 128 * j + 64 * half + 16 * u + 4 * q;  // constant bank, warp-uniform
                    *reinterpret_cast<float4*>(out + (2 * u + (q >> 1)) * 64 + 4 * (q & 1)) =
                        make_float4(__uint_as_float(v[4 * q]) + p.c.bsum[col], __uint_as_float(v[4 * q + 1]) + p.c.bsum[col + 1],
                                    __uint_as_float(v[4 * q + 2]) + p.c.bsum[col + 2], __uint_as_float(v[4 * q + 3]) + p.c.bsum[col + 3]);
                }
            }
        }
    }
    __device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
    __device__ __forceinline__ void run_tile(long long tile) {
        // the staging role's 12 + 12 writes of A0 / A1 come first in the schedule; by the time this role asks for those buffers (after the
        // last |STFT| unit, i.e. after every DFT MMA) their releases have all happened, so skipping the phases cannot alias
        bk.other_writes(A0, 12); bk.other_writes(A1, 12);
#pragma unroll 1
        for (int u = 0; u < 6; ++u) mags(u >> 1, u & 1, M0 + (u & 1));
        epi_sync();                          // side[] of all three positions is written
        act128<true>(1, p.c.e1b, A0, A1, 0); // h1_0
        act128<true>(2, p.c.e1b, M0, M1, 1); // h1_1
        act128<true>(3, p.c.e1b, A0, A1, 2); // h1_2
        wait_acc(0);                         // enc2: q0 = columns 0-63 -> M0, q1 = columns 64-127 -> M1
        act64(0, 0, p.c.e2b, M0, false);
        act64(0, 64, p.c.e2b, M1, true);
        wait_acc(1);                         // enc3 -> A0
        act64(1, 0, p.c.e3b, A0, true);
        act128<false>(2, p.c.e4b, M0, M1, 0);  // enc4 -> h4
#pragma unroll 1
        for (int j = 0; j < 4; ++j) store_pre(j, tile);
        epi_sync();                          // side[] is rewritten by the next tile's mags
        if (row == 0 && half == 0) stamp(p, 1, it, ev++);
        ++it;
        ev = 0;
    }
};

__global__ void __launch_bounds__(kThreads, 1) k_vad_front_fused(Params p) {
    extern __shared__ uint8_t smraw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    Sm sm;
    sm.base = base;
    sm.ring = base + kOffRing;
    sm.side = reinterpret_cast<float*>(base + kOffSide);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kOffBar);
    sm.pfull = bars; sm.pfree = bars + 4; sm.tfull = bars + 8; sm.tfree = bars + 12; sm.bfull = bars + 16; sm.bfree = bars + 16 + kRing;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(base + kOffTmem);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&sm.pfull[i], 256);
            mbar_init(&sm.pfree[i], 1);
            mbar_init(&sm.tfull[i], 1);
            mbar_init(&sm.tfree[i], 256);
        }
        for (int i = 0; i < kRing; ++i) { mbar_init(&sm.bfull[i], 1); mbar_init(&sm.bfree[i], 1); }
    }
    if (warp == 0) {  // 512 TMEM columns = four 128-column accumulator slots (one CTA per SM by construction: 227 KB of smem)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    fence_after();
    sm.tmem = *tmem_ptr;
    const int n_my = (p.n_tiles - p.tile0 > (int)blockIdx.x) ? (p.n_tiles - p.tile0 - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    Book bk;
    bk.wpar = bk.wany = bk.ph = bk.tany = 0;
    if (warp < 8) {
        EpiRole r(sm, bk, p, tid);
        for (int i = 0; i < n_my; ++i) r.run_tile((long long)p.tile0 + blockIdx.x + (long long)i * gridDim.x);
    } else if (warp < (kEpiThreads + kStageThreads) / 32) {
        StageRole r(sm, bk, p, tid - kEpiThreads);
        r.run(n_my);
    } else if (tid == kEpiThreads + kStageThreads) {
        MmaRole r(sm, bk, p);
        for (int i = 0; i < n_my; ++i) r.run_tile();
    } else if (tid == kEpiThreads + kStageThreads + 32) {
        producer_loop(sm, p, n_my);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(sm.tmem), "r"(512));
}

}  // namespace vf

// ------------------------------------------------------------------ host side: weight image in consumption order
struct VadFront {
    uint8_t* wimg = nullptr;
    uint2* slots = nullptr;
    vf::Consts consts;  // host copy: travels to the device as part of the kernel parameters
};

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// one block = `rows` x 64 weights -> [hi plane rows x 128 B][lo plane], 128B-swizzled K-major
template <typename F>
static uint32_t add_block(std::vector<uint8_t>& img, int rows, F&& at) {
    const uint32_t off = (uint32_t)img.size();
    img.resize(img.size() + (size_t)2 * rows * 128, 0);
    uint16_t* hi = reinterpret_cast<uint16_t*>(img.data() + off);
    uint16_t* lo = hi + (size_t)rows * 64;
    for (int r = 0; r < rows; ++r)
        for (int e = 0; e < 64; ++e) {
            const float v = at(r, e);
            const uint16_t h = f2bf(v), l = f2bf(v - bf2f(h));
            const size_t o = ((size_t)(r >> 3) * 1024 + (r & 7) * 128 + (((e >> 3) ^ (r & 7)) * 16) + (e & 7) * 2) / 2;
            hi[o] = h;
            lo[o] = l;
        }
    return off;
}

int vad_front_create(const float* w, const VadFrontLayout& L, VadFront** out) {
    std::vector<uint8_t> img;
    auto basis = [&](int row, int n) { return w[L.basis + (size_t)row * 256 + n]; };  // [258][256]: rows 0-128 real, 129-257 imaginary
    uint32_t o_basis[2][4], o_e1[3][2], o_e2[3][2], o_e3[3], o_e4, o_ih[4][2];
    for (int h = 0; h < 2; ++h)
        for (int kc = 0; kc < 4; ++kc)
            o_basis[h][kc] = add_block(img, 128, [&](int r, int e) {
                const int bin = 64 * h + (r >> 1), n = 64 * kc + e;
                if (h == 0 && r == 1) return basis(128, n);                 // Nyquist bin (real) in the imaginary-DC column
                return (r & 1) ? basis(129 + bin, n) : basis(bin, n);
            });
    for (int tap = 0; tap < 3; ++tap)
        for (int h = 0; h < 2; ++h)
            o_e1[tap][h] = add_block(img, 128, [&](int oc, int e) { return w[L.e1w + ((size_t)oc * 129 + 64 * h + e) * 3 + tap]; });
    for (int tap = 0; tap < 3; ++tap)
        for (int kc = 0; kc < 2; ++kc)
            o_e2[tap][kc] = add_block(img, 64, [&](int oc, int e) { return w[L.e2w + ((size_t)oc * 128 + 64 * kc + e) * 3 + tap]; });
    for (int tap = 0; tap < 3; ++tap) o_e3[tap] = add_block(img, 64, [&](int oc, int e) { return w[L.e3w + ((size_t)oc * 64 + e) * 3 + tap]; });
    o_e4 = add_block(img, 128, [&](int oc, int e) { return w[L.e4w + ((size_t)oc * 64 + e) * 3 + 1]; });
    for (int j = 0; j < 4; ++j)
        for (int kc = 0; kc < 2; ++kc)
            o_ih[j][kc] = add_block(img, 128, [&](int r, int e) { return w[L.wih + (size_t)(128 * j + r) * 128 + 64 * kc + e]; });
    // consumption order of one tile: must mirror MmaRole::run_tile group by group
    std::vector<uint2> slots;
    auto push = [&](uint32_t off, int rows) { slots.push_back(make_uint2(off, (uint32_t)(2 * rows * 128))); };
    for (int u = 0; u < 7; ++u) {
        if (u < 6)
            for (int kc = 0; kc < 4; ++kc) push(o_basis[u & 1][kc], 128);
        if (u >= 1) {
            const int v = u - 1, f = v >> 1, h = v & 1;
            for (int p = (f > 0 ? f - 1 : 0); p <= (f < 2 ? f + 1 : 2); ++p) push(o_e1[f - p + 1][h], 128);
        }
    }
    push(o_e2[1][0], 64); push(o_e2[1][1], 64);                               // h1_0 -> q0, tap 1
    push(o_e2[2][0], 64); push(o_e2[2][1], 64);                               // h1_1 -> q0, tap 2
    push(o_e2[0][0], 64); push(o_e2[0][1], 64);                               // h1_1 -> q1, tap 0
    push(o_e2[1][0], 64); push(o_e2[1][1], 64);                               // h1_2 -> q1, tap 1
    push(o_e3[1], 64); push(o_e3[2], 64);                                     // h2_0 tap 1, h2_1 tap 2
    push(o_e4, 128);
    for (int j = 0; j < 4; ++j) { push(o_ih[j][0], 128); push(o_ih[j][1], 128); }
    if ((int)slots.size() != vf::kSlotsPerTile) { set_error("internal: VAD front slot table has %zu entries", slots.size()); return OSB_ERR_CUDA; }
    VadFront* f = new VadFront();
    memcpy(f->consts.e1b, w + L.e1b, sizeof(f->consts.e1b));
    memcpy(f->consts.e2b, w + L.e2b, sizeof(f->consts.e2b));
    memcpy(f->consts.e3b, w + L.e3b, sizeof(f->consts.e3b));
    memcpy(f->consts.e4b, w + L.e4b, sizeof(f->consts.e4b));
    for (int i = 0; i < 512; ++i) f->consts.bsum[i] = w[L.bih + i] + w[L.bhh + i];
    for (int tap = 0; tap < 3; ++tap)
        for (int oc = 0; oc < 128; ++oc) f->consts.w1side[tap][oc] = w[L.e1w + ((size_t)oc * 129 + 128) * 3 + tap];
    cudaError_t e = cudaMalloc(&f->wimg, img.size());
    if (e == cudaSuccess) e = cudaMemcpy(f->wimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&f->slots, slots.size() * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMemcpy(f->slots, slots.data(), slots.size() * sizeof(uint2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        vad_front_destroy(f);
        return cuda_fail(e, "VAD front weight image", __FILE__, __LINE__);
    }
    *out = f;
    return OSB_OK;
}

void vad_front_destroy(VadFront* f) {
    if (!f) return;
    cudaFree(f->wimg);
    cudaFree(f->slots);
    delete f;
}

int vad_front_tiles(long long total_windows) { return (int)((total_windows + vf::kRows - 1) / vf::kRows); }

int launch_vad_front_fused(const VadFront* f, const void* d_audio, int fmt, long long audio_stride, int wins_per_stream, long long win0,
                           long long total_windows, float* d_pre, cudaStream_t st, int tile_begin, int tile_end, int max_ctas) {
    static PerDeviceOnce once;
    OSB_CUDA(once.run([&] { return cudaFuncSetAttribute(vf::k_vad_front_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, vf::kSmem); }));
    vf::Params p;
    p.audio = d_audio; p.fmt = fmt; p.audio_stride = audio_stride; p.wins_per_stream = wins_per_stream; p.win0 = win0; p.total = total_windows;
    p.n_tiles = vad_front_tiles(total_windows);
    if (tile_end >= 0 && tile_end < p.n_tiles) p.n_tiles = tile_end;
    p.tile0 = tile_begin > 0 ? tile_begin : 0;
    p.wimg = f->wimg; p.slots = f->slots; p.pre = d_pre; p.c = f->consts;
    const int ctas = (max_ctas > 0 && max_ctas < num_sms()) ? max_ctas : num_sms();
    const int grid = (p.n_tiles - p.tile0) < ctas ? (p.n_tiles - p.tile0) : ctas;
    if (grid <= 0) return OSB_OK;
    p.trace = nullptr;
    const char* tpath = getenv("OSB_VF_TRACE");  // debug: clock stamps of CTA 0's first four tiles, written as text after a device sync
    if (tpath) {
        OSB_CUDA(cudaMalloc(&p.trace, 16 * 64 * 8));
        OSB_CUDA(cudaMemset(p.trace, 0, 16 * 64 * 8));
    }
    OSB_LAUNCH(vf::k_vad_front_fused, grid, vf::kThreads, vf::kSmem, st, p);
    OSB_CHECK_LAUNCH();
    if (tpath) {
        std::vector<unsigned long long> h(16 * 64);
        OSB_CUDA(cudaDeviceSynchronize());
        OSB_CUDA(cudaMemcpy(h.data(), p.trace, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        if (FILE* f = fopen(tpath, "w")) {
            unsigned long long t0 = ~0ull;
            for (size_t i = 0; i < h.size(); ++i) if ((i % 64) < 40 && h[i] && h[i] < t0) t0 = h[i];
            static const char* names[4] = {"mma", "epi", "stage", "prod"};
            for (int r = 0; r < 4; ++r)
                for (int it = 0; it < 4; ++it) {
                    fprintf(f, "%s tile%d:", names[r], it);
                    for (int e = 0; e < 64; ++e) {
                        const unsigned long long v = h[(r * 4 + it) * 64 + e];
                        if (v) fprintf(f, " %d=%llu", e, e < 40 ? v - t0 : v);
                    }
                    fprintf(f, "\n");
                }
            fclose(f);
        }
    }
    return OSB_OK;
}

}  // namespace osb
