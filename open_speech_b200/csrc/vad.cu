// Silero-VAD-shaped frame scoring and speech segmenting.
//
// Replaces (reference file:line):
//   SileroVAD.__call__ / is_speech / get_speech_segments   src/vad/silero.py:63-177
//   (the per-window onnxruntime session.run at :86 and :149, and the integer segmenter :133-177)
//
// Structure (the reference feeds raw 512-sample windows, no 64-sample context):
//   front  (stateless, batched over every window of every stream; GEMM-shaped):
//          right-reflect-pad 64 -> 3 frames x 256 -> Hann-DFT conv (258x256) -> |.| ->
//          4x Conv1d(k=3)+ReLU -> W_ih.x + b_ih + b_hh          => gate pre-activations [W][512]
//   recur  (sequential over windows, one CTA per stream): gates += W_hh.h ; LSTM cell ; head
//   segment(integer state machine, one warp per stream): bit-exact with silero.py:133-177
#include <cstdlib>
#include <mutex>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "vad_front.cuh"

namespace osb {

constexpr int kWin = 512, kHid = 128, kGates = 512;
constexpr int kMagC = 132;            // 129 magnitude channels padded to a multiple of 4 floats
constexpr int kK1 = 400;              // enc1 GEMM K: 3*132 = 396 padded to a multiple of 16

// flat host weight blob layout (must match vad/silero.py WEIGHT_LAYOUT)
constexpr size_t oBasis = 0, nBasis = 258 * 256;
constexpr size_t oE1w = oBasis + nBasis, nE1w = 128 * 129 * 3, oE1b = oE1w + nE1w;
constexpr size_t oE2w = oE1b + 128, nE2w = 64 * 128 * 3, oE2b = oE2w + nE2w;
constexpr size_t oE3w = oE2b + 64, nE3w = 64 * 64 * 3, oE3b = oE3w + nE3w;
constexpr size_t oE4w = oE3b + 64, nE4w = 128 * 64 * 3, oE4b = oE4w + nE4w;
constexpr size_t oWih = oE4b + 128, nW = 512 * 128, oWhh = oWih + nW;
constexpr size_t oBih = oWhh + nW, oBhh = oBih + 512, oDw = oBhh + 512, oDb = oDw + 128;
constexpr size_t kBlobFloats = oDb + 1;

struct VadModel {
    int device;
    // GEMM "B" matrices, row-major [N][K] with K padded as the kernels expect
    float *basis;   // [258][256], rows interleaved re0,im0,re1,im1,... so |.| pairs are adjacent columns
    float *e1w, *e1b;  // [128][400]  (k-major: col = k*132 + ic)
    float *e2w, *e2b;  // [64][384]   (col = k*128 + ic)
    float *e3w, *e3b;  // [64][192]   (col = k*64 + ic)
    float *e4w, *e4b;  // [128][192]
    float *wih, *bsum; // [512][128], b_ih + b_hh
    float *whh;        // [512][128]
    float *whh_perm;   // the same weights in the per-thread block order of k_vad_recur
    uint32_t *whh_tc;  // the same weights as fp16 mma.sync A fragments in the register order of k_vad_recur_tc
    int recur_tc;      // 1: tensor-pipe recurrence, eight streams per CTA (default) | 0: FP32 FFMA lock-step kernels (cross-check)
    float *dw;         // [128]
    float db;
    // tcgen05 path: per layer, B as split-bf16 (hi, lo) tiles pre-arranged in the 128B-swizzled K-major
    // shared-memory image, [n_tile][k_chunk][plane][NT*64]
    struct Tc { uint16_t* img; int NT, n_tiles, k_chunks; } tc[6];
    VadFront* fused;   // weight image of the fused persistent front kernel (vad_front.cu)
    int use_tc;        // 2: fused tcgen05 kernel (default) | 1: one tcgen05 GEMM per layer | 0: FP32 FFMA GEMMs (cross-check)
};

// ------------------------------------------------------------------ batched front GEMM
// C[r][n] = act( sum_k A(r,k) * B[n][k] + bias[n] ),   A rows addressed through a descriptor
struct GemmDesc {
    const void* A;
    long long a_outer;  // elements between consecutive "outer" groups of rows
    int a_icount;       // rows per outer group
    int a_istride;      // elements between rows inside a group
    const float* B;
    const float* bias;
    float* C;
    long long c_outer;
    int c_icount, c_istride, c_offset;
    int M, N, K, relu;
    // audio A-mode only: rows are (stream, window, frame); frames read the stream's samples
    long long audio_stride;  // samples between streams
    int wins_per_stream;     // windows of this chunk per stream
    long long win0;          // first window of the chunk
};

constexpr int BM = 128, BN = 64, BK = 16;

template <int AMODE>
__device__ __forceinline__ void load_a8(const GemmDesc& d, int r, int k0, float v[8]) {
    if (r >= d.M) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        return;
    }
    if (AMODE == 0) {
        const float* p = reinterpret_cast<const float*>(d.A) + (long long)(r / d.a_icount) * d.a_outer + (long long)(r % d.a_icount) * d.a_istride + k0;
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        // r = (stream*wins + win)*3 + frame ; sample s = 128*frame + k in the 576-sample padded window
        const int f = r % 3, wq = r / 3;
        const int sidx = wq / d.wins_per_stream;
        const long long win = d.win0 + (wq - sidx * d.wins_per_stream);
        const long long base = (long long)sidx * d.audio_stride + win * kWin;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int s = 128 * f + k0 + i;
            s = s < kWin ? s : (2 * kWin - 2 - s);  // right reflect pad (no edge repeat)
            if (AMODE == 1) v[i] = ((float)reinterpret_cast<const int16_t*>(d.A)[base + s] * 3.0517578125e-05f);
            else v[i] = reinterpret_cast<const float*>(d.A)[base + s];
        }
    }
}

template <int AMODE, int EPI>  // EPI 0: bias(+relu) ; 1: magnitude of adjacent (re,im) column pairs
__global__ void __launch_bounds__(256) k_vad_gemm(GemmDesc d) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int a_row = tid & 127, a_k = (tid >> 7) * 8;
    const int b_n = tid & 63, b_k = (tid >> 6) * 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float av[8];
    float4 bv;
    auto gload = [&](int kt) {
        load_a8<AMODE>(d, m0 + a_row, kt * BK + a_k, av);
        const int n = n0 + b_n;
        bv = (n < d.N) ? *reinterpret_cast<const float4*>(d.B + (long long)n * d.K + kt * BK + b_k) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k + i][a_row] = av[i];
        Bs[buf][b_k + 0][b_n] = bv.x; Bs[buf][b_k + 1][b_n] = bv.y; Bs[buf][b_k + 2][b_n] = bv.z; Bs[buf][b_k + 3][b_n] = bv.w;
    };
    const int KT = d.K / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty * 8 + i;
        if (r >= d.M) continue;
        float* crow = d.C + (long long)(r / d.c_icount) * d.c_outer + (long long)(r % d.c_icount) * d.c_istride + d.c_offset;
        const int n = n0 + tx * 4;
        if (EPI == 1) {
            // columns (n, n+1) = (re, im) of bin n/2
            if (n < d.N) crow[n >> 1] = sqrtf(acc[i][0] * acc[i][0] + acc[i][1] * acc[i][1]);
            if (n + 2 < d.N) crow[(n >> 1) + 1] = sqrtf(acc[i][2] * acc[i][2] + acc[i][3] * acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j < d.N) {
                    float v = acc[i][j] + (d.bias ? d.bias[n + j] : 0.f);
                    crow[n + j] = d.relu ? fmaxf(v, 0.f) : v;
                }
            }
        }
    }
}


// ------------------------------------------------------------------ tcgen05 front GEMM (split-bf16, FP32 accumulate in TMEM)
// C[128 rows x N] per CTA.  A rows (f32 activations or pcm16 audio frames) are split by the CTA's threads into
// bf16 hi + lo and written to shared memory in the canonical 128B-swizzled K-major layout; B (weights) arrives
// pre-split and pre-swizzled through the TMA unit (cp.async.bulk).  Three MMAs per k-step (hi*hi, hi*lo, lo*hi)
// give ~2^-16 relative operand precision, FP32 accumulation in tensor memory; the epilogue reads TMEM with
// tcgen05.ld and fuses bias/ReLU (or the |re,im| magnitude of the DFT conv).
constexpr int kTcM = 128, kTcKc = 64;                 // rows per CTA, K elements per chunk (one 128 B swizzle row of bf16)
constexpr int kTcNTmax = 144;
constexpr int kTcAPlane = kTcM * kTcKc * 2;           // 16 KB
constexpr int kTcABuf = 2 * kTcAPlane;                // hi + lo
constexpr int kTcBPlaneMax = kTcNTmax * kTcKc * 2;    // 18 KB
constexpr int kTcBBuf = 2 * kTcBPlaneMax;
constexpr int kTcNB = 4;                              // B ring depth
constexpr int kTcSmem = 2 * kTcABuf + kTcNB * kTcBBuf + 1024;

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // start address (>>4) | LBO=1 (unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) | version 1 | SWIZZLE_128B
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    // packed conversions (cvt.rn.bf16x2.f32): a in the low half, b in the high half; the residual of each element
    // against its own bf16 head is exact in float32
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Warp roles: warps 0-15 (512 threads) stage A and run the epilogue; warp 16 (one elected lane) owns the tensor core:
// it waits for a staged A chunk and the matching B tiles, issues the MMAs and refills the B ring.  Staging of chunk
// k+1 therefore overlaps the MMAs of chunk k; the hand-over in both directions is by mbarrier (fullA / doneA).
constexpr int kTcThreads = 544;
template <int AMODE, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) k_vad_gemm_tc(GemmDesc d, const uint16_t* __restrict__ Bimg, int NT, int n_tiles, int k_chunks) {
    extern __shared__ uint8_t smraw[];
    uint8_t* smb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Abuf = smb;                    // [2][hi|lo][128 x 128 B]
    uint8_t* Bbuf = smb + 2 * kTcABuf;      // [4][hi|lo][NT x 128 B]
    __shared__ __align__(8) uint64_t fullB[kTcNB], doneB[kTcNB], fullA[2], doneA[2], doneAll;
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTcM;
    const uint32_t b_plane = (uint32_t)NT * kTcKc * 2, b_tile = 2 * b_plane;

    if (tid == 0) {
        for (int i = 0; i < kTcNB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&doneB[i], 1); }
        mbar_init(&doneA[0], 1); mbar_init(&doneA[1], 1); mbar_init(&doneAll, 1);
        mbar_init(&fullA[0], 512); mbar_init(&fullA[1], 512);
    }
    if (warp == 0) {  // one full warp allocates all 512 TMEM columns (1 CTA per SM by construction: 209 KB of smem)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_sm)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sm;
    const int steps = k_chunks * n_tiles;
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    auto issue_b_load = [&](int s) {  // step s = kc * n_tiles + nt ; image order is [nt][kc]
        const int kc = s / n_tiles, nt = s - kc * n_tiles, b = s % kTcNB;
        const uint16_t* src = Bimg + ((size_t)nt * k_chunks + kc) * (b_tile / 2);
        mbar_expect_tx(&fullB[b], b_tile);
        bulk_g2s(Bbuf + (size_t)b * kTcBBuf, src, b_plane, &fullB[b]);
        bulk_g2s(Bbuf + (size_t)b * kTcBBuf + kTcBPlaneMax, src + b_plane / 2, b_plane, &fullB[b]);
    };
    if (tid == 512)
        for (int s = 0; s < kTcNB - 1 && s < steps; ++s) issue_b_load(s);

    const int arow = tid & 127, aq = (tid >> 7) & 3;  // thread -> (row, 16-element quarter of the 64-wide chunk)
    const float* a_row = nullptr;  // AMODE 0: start of this thread's activation row (the same for every chunk)
    if (AMODE == 0 && warp < 16 && m0 + arow < d.M) {
        const int r = m0 + arow;
        a_row = reinterpret_cast<const float*>(d.A) + (long long)(r / d.a_icount) * d.a_outer + (long long)(r % d.a_icount) * d.a_istride;
    }
    // A(kc) -> registers: 16 elements per thread
    auto load_A = [&](int kc, float (&v)[16]) {
            const int r = m0 + arow, k0 = kc * kTcKc + 16 * aq;
            if (r >= d.M) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            } else if (AMODE == 0) {
                const float* p = a_row + k0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
            } else {
                const int f = r % 3, wq = r / 3;
                const int sidx = wq / d.wins_per_stream;
                const long long win = d.win0 + (wq - sidx * d.wins_per_stream);
                const long long base = (long long)sidx * d.audio_stride + win * kWin;
                const int s0 = 128 * f + k0;
                const int16_t* p16 = reinterpret_cast<const int16_t*>(d.A) + base + s0;
                if (AMODE == 1 && s0 + 16 <= kWin && (((uintptr_t)p16) & 15) == 0) {
                    const uint4 w0 = *reinterpret_cast<const uint4*>(p16), w1 = *reinterpret_cast<const uint4*>(p16 + 8);
                    const uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[2 * i] = (float)(int16_t)(ws[i] & 0xFFFF) * 3.0517578125e-05f;
                        v[2 * i + 1] = (float)(int16_t)(ws[i] >> 16) * 3.0517578125e-05f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int si = s0 + i;
                        si = si < kWin ? si : (2 * kWin - 2 - si);
                        if (AMODE == 1) v[i] = (float)reinterpret_cast<const int16_t*>(d.A)[base + si] * 3.0517578125e-05f;
                        else v[i] = reinterpret_cast<const float*>(d.A)[base + si];
                    }
                }
            }
    };
    // registers -> split bf16 hi/lo -> 2 x 16 B per plane at swizzled positions of buffer kc & 1
    auto store_A = [&](int kc, const float (&v)[16]) {
        const int ab = kc & 1;
            uint8_t* hi_row = Abuf + (size_t)ab * kTcABuf + (arow >> 3) * 1024 + (arow & 7) * 128;
            uint8_t* lo_row = hi_row + kTcAPlane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_bf16x2(v[8 * c + 2 * j], v[8 * c + 2 * j + 1], h[j], l[j]);
                const int chunk = (2 * aq + c) ^ (arow & 7);
                *reinterpret_cast<uint4*>(hi_row + chunk * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(lo_row + chunk * 16) = make_uint4(l[0], l[1], l[2], l[3]);
            }
    };
    // the loads of chunk kc+1 are in flight while chunk kc waits for its buffer, is split and stored
    auto stage = [&](int kc, float (&cur)[16], float (&nxt)[16]) {
        const int ab = kc & 1;
        if (kc + 1 < k_chunks) load_A(kc + 1, nxt);
        if (kc >= 2) mbar_wait(&doneA[ab], (uint32_t)(((kc >> 1) - 1) & 1));  // MMAs of chunk kc-2 have drained this buffer
        store_A(kc, cur);
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&fullA[ab]);
    };
    if (warp < 16) {
        float va[16], vb[16];
        load_A(0, va);
        for (int kc = 0; kc < k_chunks; kc += 2) {
            stage(kc, va, vb);
            if (kc + 1 < k_chunks) stage(kc + 1, vb, va);
        }
    }
    for (int kc = 0; kc < k_chunks; ++kc) {
        const int ab = kc & 1;
        if (tid == 512) {
            mbar_wait(&fullA[ab], (uint32_t)((kc >> 1) & 1));
            tc_fence_after();
            const uint32_t a_hi = smem_u32(Abuf + (size_t)ab * kTcABuf), a_lo = a_hi + kTcAPlane;
            for (int nt = 0; nt < n_tiles; ++nt) {
                const int s = kc * n_tiles + nt, b = s % kTcNB;
                mbar_wait(&fullB[b], (uint32_t)((s / kTcNB) & 1));
                tc_fence_after();
                const uint32_t b_hi = smem_u32(Bbuf + (size_t)b * kTcBBuf), b_lo = b_hi + kTcBPlaneMax;
                const uint32_t dcol = tmem_base + (uint32_t)(nt * NT);
#pragma unroll
                for (int k = 0; k < kTcKc / 16; ++k) {
                    const uint32_t ko = k * 32;  // 16 bf16 = 32 B along the swizzled row
                    umma_bf16(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), idesc, (kc | k) ? 1u : 0u);
                    umma_bf16(dcol, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + ko), idesc, 1u);
                    umma_bf16(dcol, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + ko), idesc, 1u);
                }
                umma_commit(&doneB[b]);
                if (nt == n_tiles - 1) umma_commit(&doneA[ab]);
                // refill the ring: the buffer of step s-1 is free once its MMAs completed
                const int sn = s + kTcNB - 1;
                if (sn < steps) {
                    if (s >= 1) mbar_wait(&doneB[(s - 1) % kTcNB], (uint32_t)(((s - 1) / kTcNB) & 1));
                    issue_b_load(sn);
                }
            }
            if (kc == k_chunks - 1) umma_commit(&doneAll);
        }
    }
    // ---- epilogue: TMEM -> registers -> shared -> global.  tcgen05.ld hands lane i of a warp 32 consecutive columns of
    // row (warp & 3) * 32 + i.  Bias / ReLU (or the |re,im| magnitude) happen in registers, the 32 x 32 block is
    // transposed through shared memory (the operand buffers are free now) and leaves as float4s with eight lanes on
    // one output row: every store instruction writes four complete 128-byte row segments.
    if (warp < 16) {
        mbar_wait(&doneAll, 0);
        tc_fence_after();
        const int ncb = (n_tiles * NT + 31) / 32;
        constexpr int kTs = 36;  // padded row of the transpose tile (floats): 16-byte aligned, conflict-free float4 rows
        float* tile = reinterpret_cast<float*>(smb) + warp * (32 * kTs);
        const int rbase = m0 + (warp & 3) * 32;
        // this lane's four output rows in the store phase: row = rbase + 4 * j + (lane >> 3), columns 4 * (lane & 7) .. +3
        float* crow[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = rbase + 4 * j + (lane >> 3);
            crow[j] = r < d.M ? d.C + (long long)(r / d.c_icount) * d.c_outer + (long long)(r % d.c_icount) * d.c_istride + d.c_offset : nullptr;
        }
        const int ncols = (EPI == 1) ? 16 : 32;           // output columns per block
        const int nvalid = (EPI == 1) ? d.N / 2 : d.N;    // output columns that exist
        for (int cb = warp >> 2; cb < ncb; cb += 4) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                         "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int n0 = cb * ncols;
            if (EPI == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float mg[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float re = __uint_as_float(v[8 * j + 2 * e]), im = __uint_as_float(v[8 * j + 2 * e + 1]);
                        mg[e] = sqrtf(re * re + im * im);
                    }
                    *reinterpret_cast<float4*>(tile + lane * kTs + 4 * j) = make_float4(mg[0], mg[1], mg[2], mg[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float x[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int n = n0 + 4 * j + e;
                        x[e] = __uint_as_float(v[4 * j + e]) + ((d.bias && n < d.N) ? __ldg(d.bias + n) : 0.f);
                        if (d.relu) x[e] = fmaxf(x[e], 0.f);
                    }
                    *reinterpret_cast<float4*>(tile + lane * kTs + 4 * j) = make_float4(x[0], x[1], x[2], x[3]);
                }
            }
            __syncwarp();
            const int c4 = 4 * (lane & 7);
            if (c4 < ncols) {
                const int n = n0 + c4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!crow[j] || n >= nvalid) continue;
                    const float4 t = *reinterpret_cast<const float4*>(tile + (4 * j + (lane >> 3)) * kTs + c4);
                    if (n + 3 < nvalid) *reinterpret_cast<float4*>(crow[j] + n) = t;
                    else {
                        crow[j][n] = t.x;
                        if (n + 1 < nvalid) crow[j][n + 1] = t.y;
                        if (n + 2 < nvalid) crow[j][n + 2] = t.z;
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
}

// ------------------------------------------------------------------ recurrence + head
// One CTA of 512 threads walks S streams in lock step; W_hh (64 K floats) stays in the CTA's registers (+ a slice in
// shared memory: the register file is exactly 64 K words), so a step costs one pass over the weights for S streams.
//
// The hidden vector has to reach every FMA from shared memory, and the shared-memory pipe - not the FMA pipe - bounds
// the step.  So a thread does not own one 128-long row (32 h loads per stream) but a 4-row x 32-column block (8 h
// loads per stream): lane l of warp w holds W[32w + 4(l&7) + i][32(l>>3) + k], i < 4, k < 32 (the 8 lanes of a
// quarter-warp share one column quarter, so each 16-byte h load is a single address per quarter-warp), the four lanes
// l, l^8, l^16, l^24 combine their partial sums with two shuffle rounds, and each lane ends up with one finished gate row.
// The matvec issues as packed FFMA2.  S is chosen so that the grid is at most one wave (148 CTAs).
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// cell non-linearities on the SFU (ex2 + rcp): absolute error ~1e-7, three decimal orders inside the 1e-3 probability
// budget, and they sit on the serial chain of every step
// 1 / x for x in [1, 2]: the bare MUFU.RCP (1 ulp).  __frcp_rn carries a slow-path branch for operands it cannot meet here, and a branch on
// the serial chain of every step is a scheduling fence and two extra Newton steps (256 x 1 h: 196 -> 170 ms)
__device__ __forceinline__ float rcp_1to2(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// sigmoid with explicit roundings (both recurrence kernels must round alike): t = exp(-|x|), r = 1/(1+t); sigmoid(|x|) = r, sigmoid(-|x|) = t r
__device__ __forceinline__ float sigmoidf_det(float x) {
    const float t = __expf(-fabsf(x));
    const float r = rcp_1to2(__fadd_rn(1.0f, t));
    return x >= 0.f ? r : __fmul_rn(t, r);
}
__device__ __forceinline__ float tanhf_fast(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf(__fmul_rn(__fsub_rn(1.0f, t), rcp_1to2(__fadd_rn(1.0f, t))), x);
}

// Gate pre-activation of window w (flattened (stream, t) index of the launch), gate row `own`.  Linear: [w][512].  Interleaved (what the
// fused front writes, vad_front.cu): [w / 8][own / 8][w % 8][own % 8] -- eight lanes (rows) of the front's epilogue fill 256 contiguous
// bytes with two store instructions, the eight gate rows a quarter-warp of this kernel reads are one 32-byte sector (as in the linear
// layout), and the eight windows a stream consumes over eight steps sit in one contiguous 16 KB block.
__device__ __forceinline__ long long pre_at(long long w, int own, bool tiled) {
    return tiled ? (w >> 3) * 4096 + (long long)((own >> 3) * 64 + (own & 7)) + (w & 7) * 8 : w * kGates + own;
}

constexpr int kHq = 36;  // floats between the four 32-float quarters of h in shared memory: the quarters sit in different banks
template <int S>
struct RecurCfg {
    static constexpr int RKG = S == 1 ? 6 : (S == 2 ? 5 : (S == 3 ? 4 : 3));  // of the 8 four-column groups of a thread's block, those kept in registers
    static constexpr int SKG = 8 - RKG;
    // W slice | h, double-buffered | per-warp partials of the head: a ring of 64 steps (row of 16 padded to 17)
    static constexpr int smem = (SKG * 4 * 4 * kGates + 2 * S * 4 * kHq + 64 * S * 17) * (int)sizeof(float);
};
// Rows are dealt so that ONE WARP owns all four gates of its eight units: warp-local row r = 4 j + gate is row
// 128 gate + 8 warp + j of W_hh (PyTorch gate order i, f, g, o).  After the shuffle reduction the four lanes (j, q = 0..3)
// hold the four gates of unit 8 warp + j, so the cell update is three more shuffles inside the warp: no gate exchange
// through shared memory and ONE block barrier per step (h and the head partials are double-buffered) instead of two.
__host__ __device__ inline int recur_row(int warp, int r) { return 128 * (r & 3) + 8 * warp + (r >> 2); }

// whh_perm: float4 chunk c = i*8 + kg of thread tid at [(c*512 + tid)*4], i = row of the block, kg = column group
template <int S, bool ILV>  // ILV: the pre-activations come in the fused front's interleaved layout (pre_at)
__global__ void __launch_bounds__(512, 1) k_vad_recur(const float* __restrict__ pre, long long pre_stream_stride, int n_steps,
                                                      const float4* __restrict__ whh_perm, const float* __restrict__ dw, float db,
                                                      float* __restrict__ state, float* __restrict__ probs, long long probs_stride,
                                                      long long win0, int batch) {
    constexpr int RKG = RecurCfg<S>::RKG, SKG = RecurCfg<S>::SKG;
    extern __shared__ __align__(16) float sm[];
    float4* w_sm = reinterpret_cast<float4*>(sm);                 // [4][SKG][512] float4
    float* h_sm = sm + SKG * 4 * 4 * kGates;                      // [2][S][4 quarters x 36]
    float* part_sm = h_sm + 2 * S * 4 * kHq;                      // [64 steps][S][17]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane >> 3, j = lane & 7, b0 = blockIdx.x * S;
    const int ns = min(S, batch - b0);                            // live streams of this CTA
    const int gate = ((q & 1) << 1) | (q >> 1);                   // the gate row this lane holds after the shuffle reduction
    const int unit = 8 * warp + j;
    const int own = 128 * gate + unit;                            // its row in W_hh / in the gate pre-activations
    const int upos = (unit >> 5) * kHq + (unit & 31);
    float4 w[4][RKG];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int kg = 0; kg < RKG; ++kg) w[i][kg] = whh_perm[(i * 8 + kg) * kGates + tid];
#pragma unroll
        for (int kg = RKG; kg < 8; ++kg) w_sm[(i * SKG + kg - RKG) * kGates + tid] = whh_perm[(i * 8 + kg) * kGates + tid];
    }
    // the q == 0 lane of a unit carries its cell state and writes its h
    float c[S];
    const float dw_u = dw[unit];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float* st = state + (long long)(b0 + (s < ns ? s : 0)) * 2 * kHid;
        c[s] = s < ns ? st[kHid + unit] : 0.f;
        if (q == 0) h_sm[s * 4 * kHq + upos] = s < ns ? st[unit] : 0.f;
    }
    const float* p[S];  // this thread's gate row of the NEXT window to load, per stream
    int ph[S];          // ILV: that window's index mod 8 inside its 16 KB block
    float pre_v[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const long long w0 = (long long)(b0 + (s < ns ? s : 0)) * (pre_stream_stride / kGates);  // first window of the stream in the launch
        p[s] = pre + pre_at(w0, own, ILV);
        ph[s] = (int)(w0 & 7);
        pre_v[s] = n_steps > 0 ? *p[s] : 0.f;
    }
    auto advance = [&](int s) {  // p[s] -> the same gate row of the next window
        if (ILV) {
            ph[s] = (ph[s] + 1) & 7;
            p[s] += ph[s] ? 8 : 4096 - 56;
        } else p[s] += kGates;
    };
    __syncthreads();
    for (int t = 0; t < n_steps; ++t) {
        const float* hr = h_sm + (t & 1) * (S * 4 * kHq);          // h of the previous step
        float* hw = h_sm + ((t + 1) & 1) * (S * 4 * kHq);          // h of this step
        float* pw = part_sm + (t & 63) * (S * 17);
        float pre_next[S];
        float2 a[4][S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            advance(s);
            pre_next[s] = (t + 1 < n_steps) ? *p[s] : 0.f;  // prefetch (consumed one step later)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i][s] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) {
            float4 h4[S];
#pragma unroll
            for (int s = 0; s < S; ++s) h4[s] = *reinterpret_cast<const float4*>(hr + s * 4 * kHq + q * kHq + 4 * kg);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w4 = kg < RKG ? w[i][kg < RKG ? kg : 0] : w_sm[(i * SKG + (kg < RKG ? 0 : kg - RKG)) * kGates + tid];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    a[i][s] = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(h4[s].x, h4[s].y), a[i][s]);
                    a[i][s] = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(h4[s].z, h4[s].w), a[i][s]);
                }
            }
        }
        // Phase by phase over the S streams (independent chains interleave), no divergent region: every lane runs the cell
        // arithmetic, only the unit's q == 0 lane holds meaningful state and stores.
        float x[S], act[S], tt[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            // combine the four column quarters of a row group: two shuffle rounds, one finished gate per lane
            const float p0 = a[0][s].x + a[0][s].y, p1 = a[1][s].x + a[1][s].y, p2 = a[2][s].x + a[2][s].y, p3 = a[3][s].x + a[3][s].y;
            const bool hi = q & 1;
            const float v0 = (hi ? p2 : p0) + __shfl_xor_sync(0xffffffffu, hi ? p0 : p2, 8);
            const float v1 = (hi ? p3 : p1) + __shfl_xor_sync(0xffffffffu, hi ? p1 : p3, 8);
            const bool hi2 = q & 2;
            x[s] = ((hi2 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, hi2 ? v0 : v1, 16)) + pre_v[s];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            // the lane's own gate: g -> tanh, i/f/o -> sigmoid, from one exponential and one reciprocal:
            // t = exp(-k|x|), r = 1/(1+t):  sigmoid(|x|) = r, sigmoid(-|x|) = t r (k = 1),  tanh(|x|) = (1-t) r (k = 2)
            const float ax = fabsf(x[s]);
            const float t = __expf(gate == 2 ? -2.0f * ax : -ax);
            const float r = rcp_1to2(__fadd_rn(1.0f, t));
            // (explicit roundings: every S instantiation must round alike, the packing of streams into CTAs is invisible)
            act[s] = gate == 2 ? copysignf(__fmul_rn(__fsub_rn(1.0f, t), r), x[s]) : (x[s] >= 0.f ? r : __fmul_rn(t, r));  // sigmoid(-|x|) = t r
        }
        float gg[S], gf[S], go[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {  // the unit's q == 0 lane collects g, f, o
            gg[s] = __shfl_sync(0xffffffffu, act[s], j + 8);
            gf[s] = __shfl_sync(0xffffffffu, act[s], j + 16);
            go[s] = __shfl_sync(0xffffffffu, act[s], j + 24);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            c[s] = __fmaf_rn(gf[s], c[s], __fmul_rn(act[s], gg[s]));
            const float h = __fmul_rn(go[s], tanhf_fast(c[s]));
            if (q == 0) hw[s * 4 * kHq + upos] = h;
            tt[s] = q == 0 ? __fmul_rn(fmaxf(h, 0.f), dw_u) : 0.f;  // head: relu(h) . w_dec
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int s = 0; s < S; ++s) tt[s] += __shfl_xor_sync(0xffffffffu, tt[s], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) pw[s * 17 + warp] = tt[s];
        }
        __syncthreads();
        // The head's sigmoid is off the serial chain: every 32 steps warp s turns the last 32 rows of partials of stream s into
        // 32 probabilities, one per lane, while the other half of the ring takes the next steps (one exp per lane per 32 steps
        // instead of a dependent exp + divide in front of every step's barrier).
        if ((t & 31) == 31 || t == n_steps - 1) {
            const int tb = t & ~31;  // first step of the block
            if (warp < ns && tb + lane <= t) {
                const float* ps = part_sm + ((tb + lane) & 63) * (S * 17) + warp * 17;
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 4) sum += (ps[k] + ps[k + 1]) + (ps[k + 2] + ps[k + 3]);
                probs[(long long)(b0 + warp) * probs_stride + win0 + tb + lane] = sigmoidf_acc(sum + db);
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) pre_v[s] = pre_next[s];
    }
    if (q == 0) {
        const float* hf = h_sm + (n_steps & 1) * (S * 4 * kHq);
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (s < ns) {
                float* st = state + (long long)(b0 + s) * 2 * kHid;
                st[unit] = hf[s * 4 * kHq + upos];
                st[kHid + unit] = c[s];
            }
    }
}

template <int S>
struct RecurMbCfg {
    static constexpr int RKG = S == 1 ? 6 : (S == 2 ? 5 : (S == 3 ? 4 : 3));  // of the 8 four-column groups of a thread's block, those kept in registers
    static constexpr int SKG = 8 - RKG;
    static constexpr int smem = (SKG * 4 * 4 * kGates + S * (4 * kHq + kGates) + kHid + 64 * S * 17) * (int)sizeof(float);
};
// The same recurrence for S >= 2 streams per CTA: there the in-warp cell update of k_vad_recur costs more than it saves (every
// lane repeats the cell arithmetic of its unit: measured 33.7 ms against 30.5 ms on 256 streams), so the gates go through shared
// memory to 128 S cell threads and the step takes two barriers.  Same row dealing, same whh_perm, same arithmetic.
template <int S, bool ILV>
__global__ void __launch_bounds__(512, 1) k_vad_recur_mb(const float* __restrict__ pre, long long pre_stream_stride, int n_steps,
                                                      const float4* __restrict__ whh_perm, const float* __restrict__ dw, float db,
                                                      float* __restrict__ state, float* __restrict__ probs, long long probs_stride,
                                                      long long win0, int batch) {
    constexpr int RKG = RecurMbCfg<S>::RKG, SKG = RecurMbCfg<S>::SKG;
    extern __shared__ __align__(16) float sm[];
    float4* w_sm = reinterpret_cast<float4*>(sm);                 // [4][SKG][512] float4
    float* h_sm = sm + SKG * 4 * 4 * kGates;                      // [S][4 quarters x 36]
    float* g_sm = h_sm + S * 4 * kHq;                             // [S][512]
    float* dw_sm = g_sm + S * kGates;                             // [128]
    float* part_sm = dw_sm + kHid;                                // [64 steps][S][17]: ring of head partials
    const int tid = threadIdx.x, q = (tid & 31) >> 3, b0 = blockIdx.x * S;
    const int ns = min(S, batch - b0);                            // live streams of this CTA
    const int own = 128 * (((q & 1) << 1) | (q >> 1)) + 8 * (tid >> 5) + (tid & 7);  // row dealing of recur_row()
    float4 w[4][RKG];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int kg = 0; kg < RKG; ++kg) w[i][kg] = whh_perm[(i * 8 + kg) * kGates + tid];
#pragma unroll
        for (int kg = RKG; kg < 8; ++kg) w_sm[(i * SKG + kg - RKG) * kGates + tid] = whh_perm[(i * 8 + kg) * kGates + tid];
    }
    // cell threads: thread (cs, cu) owns unit cu of stream cs
    const int cs = tid >> 7, cu = tid & (kHid - 1);
    const bool cell = cs < ns;
    const int hpos = cs * 4 * kHq + (cu >> 5) * kHq + (cu & 31);
    float c = 0.f;
    float* st = state + (long long)(b0 + (cell ? cs : 0)) * 2 * kHid;
    if (tid < S * kHid) h_sm[hpos] = cell ? st[cu] : 0.f;
    if (cell) c = st[kHid + cu];
    if (tid < kHid) dw_sm[tid] = dw[tid];
    const float* p[S];  // this thread's gate row of the NEXT window to load, per stream
    int ph[S];          // ILV: that window's index mod 8 inside its 16 KB block
    float pre_v[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const long long w0 = (long long)(b0 + (s < ns ? s : 0)) * (pre_stream_stride / kGates);  // first window of the stream in the launch
        p[s] = pre + pre_at(w0, own, ILV);
        ph[s] = (int)(w0 & 7);
        pre_v[s] = n_steps > 0 ? *p[s] : 0.f;
    }
    auto advance = [&](int s) {  // p[s] -> the same gate row of the next window
        if (ILV) {
            ph[s] = (ph[s] + 1) & 7;
            p[s] += ph[s] ? 8 : 4096 - 56;
        } else p[s] += kGates;
    };
    float* pr = probs + (long long)(b0 + (cell ? cs : 0)) * probs_stride + win0;
    __syncthreads();
    for (int t = 0; t < n_steps; ++t) {
        float pre_next[S];
        float2 a[4][S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            advance(s);
            pre_next[s] = (t + 1 < n_steps) ? *p[s] : 0.f;  // prefetch (consumed one step later)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i][s] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) {
            float4 h4[S];
#pragma unroll
            for (int s = 0; s < S; ++s) h4[s] = *reinterpret_cast<const float4*>(h_sm + s * 4 * kHq + q * kHq + 4 * kg);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w4 = kg < RKG ? w[i][kg < RKG ? kg : 0] : w_sm[(i * SKG + (kg < RKG ? 0 : kg - RKG)) * kGates + tid];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    a[i][s] = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(h4[s].x, h4[s].y), a[i][s]);
                    a[i][s] = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(h4[s].z, h4[s].w), a[i][s]);
                }
            }
        }
        // combine the four column quarters of a row group: two shuffle rounds, one finished row per lane
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float p0 = a[0][s].x + a[0][s].y, p1 = a[1][s].x + a[1][s].y, p2 = a[2][s].x + a[2][s].y, p3 = a[3][s].x + a[3][s].y;
            const bool hi = q & 1;
            const float v0 = (hi ? p2 : p0) + __shfl_xor_sync(0xffffffffu, hi ? p0 : p2, 8);
            const float v1 = (hi ? p3 : p1) + __shfl_xor_sync(0xffffffffu, hi ? p1 : p3, 8);
            const bool hi2 = q & 2;
            const float r = (hi2 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, hi2 ? v0 : v1, 16);
            g_sm[s * kGates + own] = r + pre_v[s];
        }
        __syncthreads();
        if (tid < S * kHid) {
            // same arithmetic, same roundings and the same reduction tree as k_vad_recur: which kernel scored a stream is invisible
            const float* g = g_sm + cs * kGates;
            const float ai = sigmoidf_det(g[cu]), af = sigmoidf_det(g[kHid + cu]), ag = tanhf_fast(g[2 * kHid + cu]), ao = sigmoidf_det(g[3 * kHid + cu]);
            c = __fmaf_rn(af, c, __fmul_rn(ai, ag));
            const float h = __fmul_rn(ao, tanhf_fast(c));
            h_sm[hpos] = h;
            // head: relu(h) . w_dec -> sigmoid ; one partial per eight units, sixteen per stream
            float part = __fmul_rn(fmaxf(h, 0.f), dw_sm[cu]);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            if ((tid & 7) == 0) part_sm[(t & 63) * (S * 17) + cs * 17 + (cu >> 3)] = part;
        }
        __syncthreads();
        if (((t & 31) == 31 || t == n_steps - 1) && cell && cu < 32) {  // as in k_vad_recur: 32 probabilities per 32 steps, one per lane
            const int tb = t & ~31;
            if (tb + cu <= t) {
                const float* ps = part_sm + ((tb + cu) & 63) * (S * 17) + cs * 17;
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 4) sum += (ps[k] + ps[k + 1]) + (ps[k + 2] + ps[k + 3]);
                pr[tb + cu] = sigmoidf_acc(sum + db);
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) pre_v[s] = pre_next[s];
    }
    if (cell) {
        st[cu] = h_sm[hpos];
        st[kHid + cu] = c;
    }
}

// ------------------------------------------------------------------ recurrence on the tensor pipe, eight streams per CTA
// The FFMA kernels above spend ~1,400 SM-cycles per stream and step on a 512 x 128 matrix-vector product whose FP32 floor alone is 512.
// Batched over EIGHT streams the product is a [512 x 128] x [128 x 8] matrix product per step, and the warp-level mma.sync is the tensor
// instruction whose shape (m16n8k16) and latency (tens of cycles, operands and result in registers, no descriptor, no TMEM round trip,
// no commit/wait) fit a chain that is a few hundred cycles long and strictly serial: tcgen05.mma needs N >= 16 at M = 128 and an
// asynchronous commit -> tcgen05.ld -> fence hand-over per step, which is most of such a step.
//   * W_hh lives in REGISTERS as fp16 A fragments for the whole launch (64 registers per thread: warp w owns the four gates of units
//     8w .. 8w+7 = two 16-row tiles x eight k tiles), h as two fp16 planes (hi + lo) in shared memory, double-buffered: two MMAs per
//     tile and k step (W.h_hi + W.h_lo), FP32 accumulate.  Probability error of this operand scheme through the recurrence, simulated
//     on 120 s clips (tools/vad_precision_sim.py recur): 1.1e-4 (budget 1e-3; W rounded to fp16 is the whole of it).
//   * Row r of tile mt: gate r / 4, unit 8w + 2 (r % 4) + mt, so a lane's accumulators hold two gates of one unit for two streams; the
//     partner lane (lane ^ 16) holds the other two gates: two shuffles, then every lane owns ONE stream's cell for two adjacent units
//     (c in registers, h written as one packed fp16 pair per plane).
//   * k slot (kt, s) of the fragments = unit (kt / 2) * 32 + ((s % 8) / 2) * 8 + (kt % 2) * 4 + (s / 8) * 2 + s % 2: with that order a
//     lane's B fragments of all eight k tiles are four 16-byte chunks of the stream's row, 64 bytes apart, bank-conflict free at a row
//     stride of 320 bytes.
//   * one barrier per step; the head's partial sums go through a 64-step ring as in the FFMA kernels.
// Which CTA column a stream sits in is invisible (every output element of an MMA is the same dot product whatever its neighbours hold).
constexpr int kTcStreams = 8;
constexpr int kTcRow = 320;  // bytes between the streams' rows of an h plane
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__host__ __device__ inline int recur_tc_kslot_unit(int kt, int s) { return (kt >> 1) * 32 + ((s & 7) >> 1) * 8 + (kt & 1) * 4 + (s >> 3) * 2 + (s & 1); }

template <bool ILV, int TERMS>  // TERMS 2: W.h_hi + W.h_lo | 1: W.h_hi only (h rounded to fp16)
__global__ void __launch_bounds__(512, 1) k_vad_recur_tc(const float* __restrict__ pre, long long pre_stream_stride, int n_steps,
                                                      const uint32_t* __restrict__ wimg, const float* __restrict__ dw, float db,
                                                      float* __restrict__ state, float* __restrict__ probs, long long probs_stride,
                                                      long long win0, int batch) {
    __shared__ __align__(16) unsigned char h_sm[2][2][kTcStreams * kTcRow];  // [buffer][plane hi | lo][stream][128 fp16 (+ pad)]
    __shared__ float part_sm[64][kTcStreams][17];                            // [step ring][stream][warp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3, b0 = blockIdx.x * kTcStreams;
    const int ns = min(kTcStreams, batch - b0);  // live streams of this CTA
    uint32_t wa[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < 8; ++kt)
#pragma unroll
            for (int r = 0; r < 4; ++r) wa[mt][kt][r] = wimg[((((warp * 2 + mt) * 8 + kt) * 4 + r) << 5) + lane];
    // this lane's cell: stream cs, units u0 and u0 + 1
    const bool is_a = g < 4;
    const int cs = 2 * tig + (g >> 2), u0 = 8 * warp + 2 * (g & 3);
    const bool live = cs < ns;
    const int srd = live ? cs : 0;  // dead columns read stream 0's inputs and store nothing
    float* st = state + (long long)(b0 + srd) * 2 * kHid;
    float2 h = live ? *reinterpret_cast<const float2*>(st + u0) : make_float2(0.f, 0.f);
    float2 c = live ? *reinterpret_cast<const float2*>(st + kHid + u0) : make_float2(0.f, 0.f);
    const float2 dwv = *reinterpret_cast<const float2*>(dw + u0);
    auto store_h = [&](unsigned char* buf) {  // h as fp16 hi + lo, one packed pair per plane
        const __half2 hi = __floats2half2_rn(h.x, h.y);
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(buf + cs * kTcRow + u0 * 2) = hi;
        if (TERMS == 2) {
            const __half2 lo = __floats2half2_rn(__fsub_rn(h.x, hf.x), __fsub_rn(h.y, hf.y));
            *reinterpret_cast<__half2*>(buf + kTcStreams * kTcRow + cs * kTcRow + u0 * 2) = lo;
        }
    };
    store_h(&h_sm[0][0][0]);
    // gate pre-activations of the lane's cell: four float2 (gates i, f, g, o of units u0, u0 + 1) per window
    const long long w_first = (long long)(b0 + srd) * (pre_stream_stride / kGates);
    const float* p = pre + pre_at(w_first, u0, ILV);  // gate G at p + G * (ILV ? 1024 : 128)
    constexpr int kGateStep = ILV ? 16 * 64 : kHid;
    int ph = (int)(w_first & 7);
    float2 pv[4];
#pragma unroll
    for (int G = 0; G < 4; ++G) pv[G] = n_steps > 0 ? *reinterpret_cast<const float2*>(p + G * kGateStep) : make_float2(0.f, 0.f);
    const unsigned char* my_b = &h_sm[0][0][0] + g * kTcRow + tig * 16;
    __syncthreads();
    for (int t = 0; t < n_steps; ++t) {
        const unsigned char* hr = my_b + (t & 1) * (2 * kTcStreams * kTcRow);
        float ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int k = 0; k < 4; ++k) ahi[mt][k] = alo[mt][k] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 bh = *reinterpret_cast<const uint4*>(hr + j * 64);
            if (TERMS == 2) {
                const uint4 bl = *reinterpret_cast<const uint4*>(hr + kTcStreams * kTcRow + j * 64);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_f16(ahi[mt], wa[mt][2 * j], bh.x, bh.y);
                    mma_f16(alo[mt], wa[mt][2 * j], bl.x, bl.y);
                    mma_f16(ahi[mt], wa[mt][2 * j + 1], bh.z, bh.w);
                    mma_f16(alo[mt], wa[mt][2 * j + 1], bl.z, bl.w);
                }
            } else {  // even and odd k tiles on separate accumulators: four chains of four instead of two of eight
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_f16(ahi[mt], wa[mt][2 * j], bh.x, bh.y);
                    mma_f16(alo[mt], wa[mt][2 * j + 1], bh.z, bh.w);
                }
            }
        }
        float hn[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float d0 = __fadd_rn(ahi[mt][0], alo[mt][0]), d1 = __fadd_rn(ahi[mt][1], alo[mt][1]);
            const float d2 = __fadd_rn(ahi[mt][2], alo[mt][2]), d3 = __fadd_rn(ahi[mt][3], alo[mt][3]);
            const float x = __shfl_xor_sync(0xffffffffu, is_a ? d1 : d0, 16), y = __shfl_xor_sync(0xffffffffu, is_a ? d3 : d2, 16);
            const float gi = __fadd_rn(is_a ? d0 : x, mt ? pv[0].y : pv[0].x), gf = __fadd_rn(is_a ? x : d1, mt ? pv[1].y : pv[1].x);
            const float gg = __fadd_rn(is_a ? d2 : y, mt ? pv[2].y : pv[2].x), go = __fadd_rn(is_a ? y : d3, mt ? pv[3].y : pv[3].x);
            float& cc = mt ? c.y : c.x;
            cc = __fmaf_rn(sigmoidf_det(gf), cc, __fmul_rn(sigmoidf_det(gi), tanhf_fast(gg)));
            hn[mt] = __fmul_rn(sigmoidf_det(go), tanhf_fast(cc));
        }
        h = make_float2(hn[0], hn[1]);
        // the next window's pre-activations: issued here, consumed a step later
        if (ILV) {
            ph = (ph + 1) & 7;
            p += ph ? 8 : 4096 - 56;
        } else p += kGates;
        if (t + 1 < n_steps) {
#pragma unroll
            for (int G = 0; G < 4; ++G) pv[G] = *reinterpret_cast<const float2*>(p + G * kGateStep);
        }
        store_h(&h_sm[(t + 1) & 1][0][0]);
        // head: relu(h) . w_dec over the warp's eight units of this stream
        float part = __fmaf_rn(fmaxf(h.y, 0.f), dwv.y, __fmul_rn(fmaxf(h.x, 0.f), dwv.x));
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        if ((g & 3) == 0) part_sm[t & 63][cs][warp] = part;
        __syncthreads();
        if (((t & 31) == 31 || t == n_steps - 1) && tid < 32 * kTcStreams) {  // 32 probabilities per stream every 32 steps
            const int tb = t & ~31;
            if (tb + lane <= t && warp < ns) {
                const float* ps = part_sm[(tb + lane) & 63][warp];
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 4) sum += (ps[k] + ps[k + 1]) + (ps[k + 2] + ps[k + 3]);
                probs[(long long)(b0 + warp) * probs_stride + win0 + tb + lane] = sigmoidf_acc(sum + db);
            }
        }
    }
    if (live) {
        *reinterpret_cast<float2*>(st + u0) = h;
        *reinterpret_cast<float2*>(st + kHid + u0) = c;
    }
}

// ------------------------------------------------------------------ segmenter (bit-exact integer machine)
// one warp per stream: 32 probabilities per coalesced load -> ballot -> lane 0 walks the bits.
__global__ void __launch_bounds__(32) k_vad_segment(const float* __restrict__ probs, long long probs_stride, long long n_win,
                                                    long long n_samples, float thr, int min_speech_windows, int silence_windows,
                                                    int* __restrict__ segs, int* __restrict__ counts, int max_seg) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* p = probs + (long long)b * probs_stride;
    int* out = segs + (long long)b * max_seg * 2;
    bool in_speech = false;
    int speech_start = 0, silence_count = 0, speech_windows = 0, nseg = 0;
    float v = lane < n_win ? p[lane] : 0.f;
    for (long long base = 0; base < n_win; base += 32) {
        const long long k = base + lane;
        const bool sp = (k < n_win) && (v >= thr);
        v = (k + 32 < n_win) ? p[k + 32] : 0.f;  // the next word's probabilities are in flight while lane 0 walks this one
        const unsigned mask = __ballot_sync(0xffffffffu, sp);
        if (lane == 0) {
            const int lim = (int)((n_win - base) < 32 ? (n_win - base) : 32);
            const unsigned full = lim == 32 ? 0xffffffffu : ((1u << lim) - 1u);
            // whole-word fast paths (the machine's state after the word is the same as after walking its bits): silence outside
            // speech changes nothing, speech inside speech only counts
            if (mask == 0u && !in_speech) continue;
            if (mask == full && in_speech) { silence_count = 0; speech_windows += lim; continue; }
            for (int i = 0; i < lim; ++i) {
                const int cur_ms = (int)(((base + i) * kWin * 1000) / 16000);
                if ((mask >> i) & 1u) {
                    silence_count = 0;
                    if (!in_speech) { in_speech = true; speech_start = cur_ms; speech_windows = 0; }
                    ++speech_windows;
                } else if (in_speech) {
                    if (++silence_count >= silence_windows) {
                        if (speech_windows >= min_speech_windows) {
                            if (nseg < max_seg) { out[2 * nseg] = speech_start; out[2 * nseg + 1] = cur_ms; }
                            ++nseg;
                        }
                        in_speech = false; silence_count = 0; speech_windows = 0;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        if (in_speech && speech_windows >= min_speech_windows) {
            if (nseg < max_seg) { out[2 * nseg] = speech_start; out[2 * nseg + 1] = (int)((n_samples * 1000) / 16000); }
            ++nseg;
        }
        counts[b] = nseg;
    }
}

// ------------------------------------------------------------------ VAD-gated assembly
// _extract_speech_segments (src/wyoming/stt_handler.py:43-115): keep [start_ms, end_ms) * (rate // 1000) of every segment,
// at the ORIGINAL rate, concatenated.  plan: one thread turns the segment list into (src start, length, dst offset).
__global__ void k_vad_gather_plan(const int* __restrict__ segs, const int* __restrict__ count, int max_seg, long long n_samples,
                                  int samples_per_ms, long long* __restrict__ plan /* [max_seg][3] */, long long* __restrict__ total) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = min(*count, max_seg);
    long long off = 0;
    for (int i = 0; i < n; ++i) {
        const long long s = (long long)segs[2 * i] * samples_per_ms;
        long long e = (long long)segs[2 * i + 1] * samples_per_ms;
        if (e > n_samples) e = n_samples;
        const long long len = (s < n_samples && e > s) ? e - s : 0;
        plan[3 * i] = s; plan[3 * i + 1] = len; plan[3 * i + 2] = off;
        off += len;
    }
    total[0] = off;
    total[1] = n;
}

__global__ void __launch_bounds__(256) k_vad_gather(const int16_t* __restrict__ pcm, const long long* __restrict__ plan,
                                                    const long long* __restrict__ total, int16_t* __restrict__ out) {
    const int seg = blockIdx.y;
    if (seg >= (int)total[1]) return;
    const long long s = plan[3 * seg], len = plan[3 * seg + 1], off = plan[3 * seg + 2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) out[off + i] = pcm[s + i];
}

// ------------------------------------------------------------------ host orchestration
static int upload(float** dst, const std::vector<float>& v) {
    OSB_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
    OSB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return OSB_OK;
}

static int upload_words(uint32_t** dst, const std::vector<uint32_t>& v) {
    OSB_CUDA(cudaMalloc(dst, v.size() * sizeof(uint32_t)));
    OSB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return OSB_OK;
}

// conv weight [oc][ic][3] -> GEMM B [oc][Kpad], col = k*icp + ic
static std::vector<float> relay_conv(const float* w, int oc, int ic, int icp, int kpad) {
    std::vector<float> o((size_t)oc * kpad, 0.f);
    for (int a = 0; a < oc; ++a)
        for (int c = 0; c < ic; ++c)
            for (int k = 0; k < 3; ++k) o[(size_t)a * kpad + k * icp + c] = w[((size_t)a * ic + c) * 3 + k];
    return o;
}

static uint16_t f2bf(float f) {  // round-to-nearest-even float -> bf16 bits
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

// B [N][Kp] f32 -> [n_tile][k_chunk][hi|lo][NT x 64] bf16 in the 128B-swizzled K-major shared-memory image
static int build_tc_image(const std::vector<float>& B, int N, int Kp, int NT, VadModel::Tc* out) {
    const int n_tiles = (N + NT - 1) / NT, k_chunks = (Kp + kTcKc - 1) / kTcKc;
    const size_t plane = (size_t)NT * kTcKc, tile = 2 * plane;
    std::vector<uint16_t> img((size_t)n_tiles * k_chunks * tile, 0);
    for (int nt = 0; nt < n_tiles; ++nt)
        for (int kc = 0; kc < k_chunks; ++kc) {
            uint16_t* t = img.data() + ((size_t)nt * k_chunks + kc) * tile;
            for (int r = 0; r < NT; ++r)
                for (int e = 0; e < kTcKc; ++e) {
                    const int n = nt * NT + r, k = kc * kTcKc + e;
                    const float v = (n < N && k < Kp) ? B[(size_t)n * Kp + k] : 0.f;
                    const uint16_t hi = f2bf(v), lo = f2bf(v - bf2f(hi));
                    const size_t off = ((size_t)(r >> 3) * 1024 + (r & 7) * 128 + (((e >> 3) ^ (r & 7)) * 16) + (e & 7) * 2) / 2;
                    t[off] = hi;
                    t[plane + off] = lo;
                }
        }
    OSB_CUDA(cudaMalloc(&out->img, img.size() * 2));
    OSB_CUDA(cudaMemcpy(out->img, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    out->NT = NT; out->n_tiles = n_tiles; out->k_chunks = k_chunks;
    return OSB_OK;
}

template <int AMODE, int EPI>
static int launch_gemm_tc(const GemmDesc& d, const VadModel::Tc& L, cudaStream_t st) {
    static PerDeviceOnce once;  // per template instance; VAD sessions are scored from several threads
    OSB_CUDA(once.run([&] { return cudaFuncSetAttribute(k_vad_gemm_tc<AMODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem); }));
    OSB_LAUNCH((k_vad_gemm_tc<AMODE, EPI>), (d.M + kTcM - 1) / kTcM, kTcThreads, kTcSmem, st, d, (const uint16_t*)L.img, L.NT, L.n_tiles, L.k_chunks);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

template <int AMODE, int EPI>
static int launch_gemm(const GemmDesc& d, cudaStream_t st) {
    dim3 grid((d.M + BM - 1) / BM, (d.N + BN - 1) / BN);
    OSB_LAUNCH((k_vad_gemm<AMODE, EPI>), grid, 256, 0, st, d);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

static int share_width() {
    static const int w = [] { const char* e = getenv("OSB_VAD_SHARE_WIDTH"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 4) ? v : 4; }();
    return w;
}

static bool pipeline_enabled() {  // read per call: a test (and the bench's A/B leg) switches it inside one process
    const char* e = getenv("OSB_VAD_PIPELINE");
    return !(e && atoi(e) == 0);
}

// per-thread side stream + events of the pipelined chunks (created once per device)
struct FrontSide {
    cudaStream_t front = nullptr;
    cudaEvent_t start = nullptr, joined = nullptr, front_done[2] = {nullptr, nullptr}, recur_done[2] = {nullptr, nullptr};
    int device = -1;
};
// Joins the side stream into the caller's stream when vad_score leaves, on the error paths too: declared behind the Scratch, so it runs
// before the scratch buffers are handed back on the caller's stream and nothing on the side stream can outlive them.
struct FrontJoin {
    FrontSide* fs = nullptr;
    cudaStream_t st = nullptr;
    ~FrontJoin() {
        if (fs && cudaEventRecord(fs->joined, fs->front) == cudaSuccess) cudaStreamWaitEvent(st, fs->joined, 0);
    }
};
static int front_side(FrontSide** out) {
    static thread_local FrontSide s;
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    if (s.device != dev) {
        if (s.front) {
            cudaStreamDestroy(s.front);
            for (cudaEvent_t e : {s.start, s.joined, s.front_done[0], s.front_done[1], s.recur_done[0], s.recur_done[1]}) cudaEventDestroy(e);
            s = FrontSide{};
        }
        OSB_CUDA(cudaStreamCreateWithFlags(&s.front, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&s.start, &s.joined, &s.front_done[0], &s.front_done[1], &s.recur_done[0], &s.recur_done[1]})
            OSB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        s.device = dev;
    }
    *out = &s;
    return OSB_OK;
}

static int vad_score(VadModel* m, const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, float* d_state,
                     float* d_probs, int64_t probs_stride, cudaStream_t st, cudaEvent_t front_done = nullptr, bool shared_gpu = false) {
    const long long n_win = n / kWin;
    if (n_win == 0 || batch == 0) return OSB_OK;
    // chunk the window axis: windows per chunk ~ 4 x 148 SMs x 128 rows, so the three GEMM shapes (3W, 2W, W rows) run
    // 12 / 8 / 4 whole waves per launch.  Larger chunks amortise the launches and the recurrence prologue (measured: x4 is
    // 9 % faster on the front than one wave) at ~0.7 GB of activations per chunk, part of which leaves L2.
    // (the fused front keeps nothing but the pre-activations in HBM: 32 waves of 128-window tiles per launch, 1.2 GB: 256 x 60 s is one chunk)
    int waves = m->use_tc == 2 ? 32 : 4;
    if (const char* e = getenv("OSB_VAD_CHUNK_WAVES")) { const int v = atoi(e); if (v >= 1 && v <= 256) waves = v; }
    long long T = ((long long)OSB_NUM_SMS * 128 * waves + batch - 1) / batch;
    if (T < 1) T = 1;
    if (T > n_win) T = n_win;
    const long long W = batch * T;  // windows per chunk
    Scratch scr(st);
    const bool fused = m->use_tc == 2;
    float *mag = nullptr, *h1 = nullptr, *h2 = nullptr, *h3 = nullptr, *h4 = nullptr, *pre;
    if (!fused) {
        OSB_CUDA(scr.alloc(&mag, (size_t)W * 5 * kMagC + 128));
        OSB_CUDA(scr.alloc(&h1, (size_t)W * 5 * 128 + 64));
        OSB_CUDA(scr.alloc(&h2, (size_t)W * 4 * 64 + 64));
        OSB_CUDA(scr.alloc(&h3, (size_t)W * 3 * 64 + 64));
        OSB_CUDA(scr.alloc(&h4, (size_t)W * 128 + 64));
        // zero once: the padding rows/channels are never written by the GEMMs
        OSB_CUDA(cudaMemsetAsync(mag, 0, ((size_t)W * 5 * kMagC + 128) * 4, st));
        OSB_CUDA(cudaMemsetAsync(h1, 0, ((size_t)W * 5 * 128 + 64) * 4, st));
        OSB_CUDA(cudaMemsetAsync(h2, 0, ((size_t)W * 4 * 64 + 64) * 4, st));
        OSB_CUDA(cudaMemsetAsync(h3, 0, ((size_t)W * 3 * 64 + 64) * 4, st));
    }
    static PerDeviceOnce once;
    OSB_CUDA(once.run([&] {
        cudaError_t e = cudaFuncSetAttribute(k_vad_recur<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurCfg<1>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurCfg<1>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<2>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<2>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<3>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<3>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<4>::smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vad_recur_mb<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RecurMbCfg<4>::smem);
        return e;
    }));
    // FP32 kernels (cross-check mode), streams per recurrence CTA: as few as keep the grid within one wave of SMs
    // shared_gpu (the composed chain runs its feature branch beside this one, stt_pipeline.cu): four streams per CTA from 75 streams on.
    // A recurrence CTA owns its SM (the whole register file) and is bound by the latency of its serial chain, so a wider CTA costs the
    // VAD branch time but hands the SMs it vacates to kernels that can fill them.
    int rs = batch <= OSB_NUM_SMS ? 1 : (batch <= 2 * OSB_NUM_SMS ? 2 : 4);
    if (shared_gpu && batch > OSB_NUM_SMS / 2) rs = share_width();
    const bool rtc = m->recur_tc != 0;
    if (rtc) rs = kTcStreams;  // the tensor-pipe kernel: always eight streams per CTA
    // Pipelined chunks (fused front, more than one chunk, the GPU to ourselves): the front of chunk i + 1 runs on a side stream BESIDE the
    // recurrence of chunk i, on the SMs the recurrence leaves free (neither kernel shares an SM: 227 KB of shared memory each), into the
    // other half of a double-buffered `pre`.  The tensor-pipe recurrence takes one SM per eight streams (256 streams: 32 SMs, the
    // front on the other 116: 139 ms per 256 x 1 h instead of 223 in series); the FP32 kernels get as few SMs as keep them in one
    // wave with at least kFrontMinSms left over (256 streams: three per CTA on 86 SMs, front on 62).
    constexpr int kFrontMinSms = 48;
    bool pipelined = false;
    if (fused && !shared_gpu && n_win > T && pipeline_enabled()) {
        int prs = rs;
        if (!rtc) {
            if (const char* e = getenv("OSB_VAD_PIPE_WIDTH")) { const int v = atoi(e); if (v >= 1 && v <= 4) prs = v > rs ? v : rs; }  // experiments
            while (prs < 4 && OSB_NUM_SMS - (int)((batch + prs - 1) / prs) < kFrontMinSms) ++prs;
        }
        if (OSB_NUM_SMS - (int)((batch + prs - 1) / prs) >= kFrontMinSms) { pipelined = true; rs = prs; }
    }
    const size_t pre_floats = (size_t)((W + 127) / 128) * 128 * kGates + 64;  // whole 128-window tiles (the fused front's tiled layout)
    OSB_CUDA(scr.alloc(&pre, pre_floats * (pipelined ? 2 : 1)));
    FrontSide* fs = nullptr;
    FrontJoin join;
    int rc;
    if (pipelined) {
        if ((rc = front_side(&fs))) return rc;
        join.fs = fs;
        join.st = st;
        OSB_CUDA(cudaEventRecord(fs->start, st));  // the audio and `pre` exist on st from here on
        OSB_CUDA(cudaStreamWaitEvent(fs->front, fs->start, 0));
    }
    const int front_ctas = OSB_NUM_SMS - (int)((batch + rs - 1) / rs);
    long long chunk = 0;
    float* const pre0 = pre;
    for (long long w0 = 0; w0 < n_win; w0 += T, ++chunk) {
        const int t = (int)((n_win - w0) < T ? (n_win - w0) : T);
        const int Wc = (int)(batch * t);
        GemmDesc d{};
        if (pipelined) {
            const int b = (int)(chunk & 1);
            pre = pre0 + (size_t)b * pre_floats;
            if (chunk >= 2) OSB_CUDA(cudaStreamWaitEvent(fs->front, fs->recur_done[b], 0));  // the recurrence of chunk - 2 has read this half
            // chunk 0 has the whole GPU; later fronts start while a recurrence holds its SMs
            if ((rc = launch_vad_front_fused(m->fused, d_audio, fmt, stride, t, w0, (long long)batch * t, pre, fs->front, 0, -1, chunk ? front_ctas : 0)))
                return rc;
            OSB_CUDA(cudaEventRecord(fs->front_done[b], fs->front));
            OSB_CUDA(cudaStreamWaitEvent(st, fs->front_done[b], 0));
        } else if (fused) {
            if ((rc = launch_vad_front_fused(m->fused, d_audio, fmt, stride, t, w0, (long long)batch * t, pre, st))) return rc;
        } else {
        // L0: DFT conv + magnitude -> mag[w][1+f][0..128]
        d.A = d_audio; d.audio_stride = stride; d.wins_per_stream = t; d.win0 = w0;
        d.B = m->basis; d.bias = nullptr; d.C = mag; d.c_outer = 5 * kMagC; d.c_icount = 3; d.c_istride = kMagC; d.c_offset = kMagC;
        d.M = Wc * 3; d.N = 258; d.K = 256; d.relu = 0;
        if (m->use_tc) rc = (fmt == OSB_FMT_PCM16) ? launch_gemm_tc<1, 1>(d, m->tc[0], st) : launch_gemm_tc<2, 1>(d, m->tc[0], st);
        else rc = (fmt == OSB_FMT_PCM16) ? launch_gemm<1, 1>(d, st) : launch_gemm<2, 1>(d, st);
        if (rc) return rc;
        // L1: enc1 129->128, k3 s1 p1, 3 positions
        d = GemmDesc{};
        d.A = mag; d.a_outer = 5 * kMagC; d.a_icount = 3; d.a_istride = kMagC;
        d.B = m->e1w; d.bias = m->e1b; d.C = h1; d.c_outer = 5 * 128; d.c_icount = 3; d.c_istride = 128; d.c_offset = 128;
        d.M = Wc * 3; d.N = 128; d.K = kK1; d.relu = 1;
        if ((rc = m->use_tc ? launch_gemm_tc<0, 0>(d, m->tc[1], st) : launch_gemm<0, 0>(d, st))) return rc;
        // L2: enc2 128->64, k3 s2 p1, 2 positions
        d = GemmDesc{};
        d.A = h1; d.a_outer = 5 * 128; d.a_icount = 2; d.a_istride = 2 * 128;
        d.B = m->e2w; d.bias = m->e2b; d.C = h2; d.c_outer = 4 * 64; d.c_icount = 2; d.c_istride = 64; d.c_offset = 64;
        d.M = Wc * 2; d.N = 64; d.K = 384; d.relu = 1;
        if ((rc = m->use_tc ? launch_gemm_tc<0, 0>(d, m->tc[2], st) : launch_gemm<0, 0>(d, st))) return rc;
        // L3: enc3 64->64, k3 s2 p1, 1 position
        d = GemmDesc{};
        d.A = h2; d.a_outer = 4 * 64; d.a_icount = 1; d.a_istride = 0;
        d.B = m->e3w; d.bias = m->e3b; d.C = h3; d.c_outer = 3 * 64; d.c_icount = 1; d.c_istride = 0; d.c_offset = 64;
        d.M = Wc; d.N = 64; d.K = 192; d.relu = 1;
        if ((rc = m->use_tc ? launch_gemm_tc<0, 0>(d, m->tc[3], st) : launch_gemm<0, 0>(d, st))) return rc;
        // L4: enc4 64->128, k3 s1 p1, 1 position
        d = GemmDesc{};
        d.A = h3; d.a_outer = 3 * 64; d.a_icount = 1; d.a_istride = 0;
        d.B = m->e4w; d.bias = m->e4b; d.C = h4; d.c_outer = 128; d.c_icount = 1; d.c_istride = 0; d.c_offset = 0;
        d.M = Wc; d.N = 128; d.K = 192; d.relu = 1;
        if ((rc = m->use_tc ? launch_gemm_tc<0, 0>(d, m->tc[4], st) : launch_gemm<0, 0>(d, st))) return rc;
        // L5: W_ih.x + (b_ih + b_hh) -> gate pre-activations
        d = GemmDesc{};
        d.A = h4; d.a_outer = 128; d.a_icount = 1; d.a_istride = 0;
        d.B = m->wih; d.bias = m->bsum; d.C = pre; d.c_outer = kGates; d.c_icount = 1; d.c_istride = 0; d.c_offset = 0;
        d.M = Wc; d.N = kGates; d.K = 128; d.relu = 0;
        if ((rc = m->use_tc ? launch_gemm_tc<0, 0>(d, m->tc[5], st) : launch_gemm<0, 0>(d, st))) return rc;
        }
        if (front_done && w0 == 0) OSB_CUDA(cudaEventRecord(front_done, st));  // from here on this branch leaves SMs free
        // recurrence over the chunk's t windows, one CTA per stream
        const unsigned rg = (unsigned)((batch + rs - 1) / rs);
#define OSB_RECUR(KERN, SMEM) OSB_LAUNCH(KERN, rg, 512, SMEM, st, pre, (long long)t * kGates, t, (const float4*)m->whh_perm, m->dw, m->db, d_state, d_probs, \
                                         (long long)probs_stride, w0, (int)batch)
#define OSB_RECUR_TC(ILV, TERMS) OSB_LAUNCH((k_vad_recur_tc<ILV, TERMS>), rg, 512, 0, st, pre, (long long)t * kGates, t, (const uint32_t*)m->whh_tc, m->dw, m->db, \
                                            d_state, d_probs, (long long)probs_stride, w0, (int)batch)
        if (rtc) {
            if (m->recur_tc == 2) { if (fused) OSB_RECUR_TC(true, 1); else OSB_RECUR_TC(false, 1); }
            else { if (fused) OSB_RECUR_TC(true, 2); else OSB_RECUR_TC(false, 2); }
        } else if (rs == 1) { if (fused) OSB_RECUR((k_vad_recur<1, true>), RecurCfg<1>::smem); else OSB_RECUR((k_vad_recur<1, false>), RecurCfg<1>::smem); }
        else if (rs == 2) { if (fused) OSB_RECUR((k_vad_recur_mb<2, true>), RecurMbCfg<2>::smem); else OSB_RECUR((k_vad_recur_mb<2, false>), RecurMbCfg<2>::smem); }
        else if (rs == 3) { if (fused) OSB_RECUR((k_vad_recur_mb<3, true>), RecurMbCfg<3>::smem); else OSB_RECUR((k_vad_recur_mb<3, false>), RecurMbCfg<3>::smem); }
        else { if (fused) OSB_RECUR((k_vad_recur_mb<4, true>), RecurMbCfg<4>::smem); else OSB_RECUR((k_vad_recur_mb<4, false>), RecurMbCfg<4>::smem); }
#undef OSB_RECUR
#undef OSB_RECUR_TC
        OSB_CHECK_LAUNCH();
        if (pipelined) OSB_CUDA(cudaEventRecord(fs->recur_done[chunk & 1], st));
    }
    return OSB_OK;
}

static int vad_segments(const float* d_probs, int64_t probs_stride, int64_t n_win, int64_t batch, int64_t n_samples, float thr,
                        int min_speech_ms, int silence_ms, int32_t* d_segs, int32_t* d_counts, int max_seg, cudaStream_t st) {
    const int window_ms = kWin * 1000 / 16000;
    int silence_windows = silence_ms / window_ms;
    if (silence_windows < 1) silence_windows = 1;
    int min_speech_windows = min_speech_ms / window_ms;
    if (min_speech_windows < 1) min_speech_windows = 1;
    OSB_LAUNCH(k_vad_segment, (unsigned)batch, 32, 0, st, d_probs, (long long)probs_stride, (long long)n_win, (long long)n_samples, thr,
               min_speech_windows, silence_windows, d_segs, d_counts, max_seg);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int launch_vad_score(void* handle, const void* d_audio, int fmt, long long n, long long batch, long long stride, float* d_state,
                     float* d_probs, long long probs_stride, cudaStream_t st, cudaEvent_t front_done, bool shared_gpu) {
    if (!handle) { set_error("invalid argument: null VAD handle"); return OSB_ERR_INVALID_ARG; }
    return vad_score(reinterpret_cast<VadModel*>(handle), d_audio, fmt, n, batch, stride, d_state, d_probs, probs_stride, st, front_done, shared_gpu);
}
// SMs the recurrence of `batch` streams holds while it runs beside other kernels (stt_pipeline.cu sizes the first feature kernel by it)
int vad_recurrence_sms(void* handle, long long batch) {
    if (!handle || batch <= 0) return 0;
    const VadModel* m = reinterpret_cast<const VadModel*>(handle);
    int rs = kTcStreams;
    if (!m->recur_tc) {
        rs = batch <= OSB_NUM_SMS ? 1 : (batch <= 2 * OSB_NUM_SMS ? 2 : 4);
        if (batch > OSB_NUM_SMS / 2) rs = share_width();
    }
    const long long ctas = (batch + rs - 1) / rs;
    return (int)(ctas < OSB_NUM_SMS ? ctas : OSB_NUM_SMS);
}
int launch_vad_segments(const float* d_probs, long long probs_stride, long long n_win, long long batch, long long n_samples, float thr,
                        int min_speech_ms, int silence_ms, int32_t* d_segs, int32_t* d_counts, int max_seg, cudaStream_t st) {
    return vad_segments(d_probs, probs_stride, n_win, batch, n_samples, thr, min_speech_ms, silence_ms, d_segs, d_counts, max_seg, st);
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_vad_create(const float* weights_host, size_t n_floats, void** handle) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(weights_host && handle, "null argument");
    if (n_floats != kBlobFloats) {
        set_error("invalid argument: VAD weight blob has %zu floats, expected %zu", n_floats, (size_t)kBlobFloats);
        return OSB_ERR_INVALID_ARG;
    }
    const float* w = weights_host;
    VadModel* m = new VadModel();
    OSB_CUDA(cudaGetDevice(&m->device));
    std::vector<float> basis((size_t)258 * 256);
    for (int k = 0; k < 129; ++k)
        for (int n = 0; n < 256; ++n) {
            basis[(size_t)(2 * k) * 256 + n] = w[oBasis + (size_t)k * 256 + n];
            basis[(size_t)(2 * k + 1) * 256 + n] = w[oBasis + (size_t)(129 + k) * 256 + n];
        }
    std::vector<float> bsum(512);
    for (int i = 0; i < 512; ++i) bsum[i] = w[oBih + i] + w[oBhh + i];
    auto vec = [&](size_t off, size_t n) { return std::vector<float>(w + off, w + off + n); };
    if ((rc = upload(&m->basis, basis)) || (rc = upload(&m->e1w, relay_conv(w + oE1w, 128, 129, kMagC, kK1))) ||
        (rc = upload(&m->e1b, vec(oE1b, 128))) || (rc = upload(&m->e2w, relay_conv(w + oE2w, 64, 128, 128, 384))) ||
        (rc = upload(&m->e2b, vec(oE2b, 64))) || (rc = upload(&m->e3w, relay_conv(w + oE3w, 64, 64, 64, 192))) ||
        (rc = upload(&m->e3b, vec(oE3b, 64))) || (rc = upload(&m->e4w, relay_conv(w + oE4w, 128, 64, 64, 192))) ||
        (rc = upload(&m->e4b, vec(oE4b, 128))) || (rc = upload(&m->wih, vec(oWih, nW))) || (rc = upload(&m->bsum, bsum)) ||
        (rc = upload(&m->whh, vec(oWhh, nW))) || (rc = upload(&m->dw, vec(oDw, 128)))) {
        delete m;
        return rc;
    }
    {   // W_hh in the block order of the recurrence kernel: chunk (i, kg) of thread tid = W[row_i][32 q + 4 kg .. +3]
        std::vector<float> perm(nW);
        for (int tid = 0; tid < 512; ++tid)
            for (int i = 0; i < 4; ++i)
                for (int kg = 0; kg < 8; ++kg) {
                    const int row = recur_row(tid >> 5, (tid & 7) * 4 + i), col = 32 * ((tid & 31) >> 3) + 4 * kg;
                    for (int e = 0; e < 4; ++e) perm[((size_t)(i * 8 + kg) * 512 + tid) * 4 + e] = w[oWhh + (size_t)row * 128 + col + e];
                }
        if ((rc = upload(&m->whh_perm, perm))) {
            delete m;
            return rc;
        }
    }
    {   // W_hh as fp16 A fragments of mma.m16n8k16 (k_vad_recur_tc): word [((warp * 2 + mt) * 8 + kt) * 4 + r][lane]
        std::vector<uint32_t> img((size_t)16 * 2 * 8 * 4 * 32);
        for (int warp = 0; warp < 16; ++warp)
            for (int mt = 0; mt < 2; ++mt)
                for (int kt = 0; kt < 8; ++kt)
                    for (int r = 0; r < 4; ++r)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int g = lane >> 2, tig = lane & 3, rr = g + 8 * (r & 1);                 // row of the 16-row tile
                            const int row = (rr >> 2) * kHid + 8 * warp + 2 * (rr & 3) + mt;               // gate rr / 4 of that unit
                            uint32_t word = 0;
                            for (int e = 0; e < 2; ++e) {
                                const int col = recur_tc_kslot_unit(kt, 2 * tig + 8 * (r >> 1) + e);
                                const __half hv = __float2half_rn(w[oWhh + (size_t)row * kHid + col]);
                                unsigned short bits;
                                memcpy(&bits, &hv, 2);
                                word |= (uint32_t)bits << (16 * e);
                            }
                            img[((size_t)(((warp * 2 + mt) * 8 + kt) * 4 + r) << 5) + lane] = word;
                        }
        if ((rc = upload_words(&m->whh_tc, img))) {
            delete m;
            return rc;
        }
    }
    m->db = w[oDb];
    {   // tcgen05 operand images
        std::vector<float> e1 = relay_conv(w + oE1w, 128, 129, kMagC, 448);
        if ((rc = build_tc_image(basis, 258, 256, 144, &m->tc[0])) || (rc = build_tc_image(e1, 128, 448, 128, &m->tc[1])) ||
            (rc = build_tc_image(relay_conv(w + oE2w, 64, 128, 128, 384), 64, 384, 64, &m->tc[2])) ||
            (rc = build_tc_image(relay_conv(w + oE3w, 64, 64, 64, 192), 64, 192, 64, &m->tc[3])) ||
            (rc = build_tc_image(relay_conv(w + oE4w, 128, 64, 64, 192), 128, 192, 128, &m->tc[4])) ||
            (rc = build_tc_image(vec(oWih, nW), 512, 128, 128, &m->tc[5]))) {
            delete m;
            return rc;
        }
    }
    {
        VadFrontLayout L{oBasis, oE1w, oE1b, oE2w, oE2b, oE3w, oE3b, oE4w, oE4b, oWih, oBih, oBhh};
        if ((rc = vad_front_create(w, L, &m->fused))) {
            delete m;
            return rc;
        }
    }
    const char* env = getenv("OSB_VAD_GEMM");  // "ffma" | "layers" | (default) fused
    m->use_tc = (env && strcmp(env, "ffma") == 0) ? 0 : ((env && strcmp(env, "layers") == 0) ? 1 : 2);
    const char* renv = getenv("OSB_VAD_RECUR");  // "ffma" | (default) tensor pipe
    m->recur_tc = (renv && strcmp(renv, "ffma") == 0) ? 0 : ((renv && strcmp(renv, "fp16x1") == 0) ? 2 : 1);
    *handle = m;
    return OSB_OK;
}

int osb_vad_destroy(void* handle) {
    if (!handle) return OSB_OK;
    VadModel* m = reinterpret_cast<VadModel*>(handle);
    float* ptrs[] = {m->basis, m->e1w, m->e1b, m->e2w, m->e2b, m->e3w, m->e3b, m->e4w, m->e4b, m->wih, m->bsum, m->whh, m->whh_perm, m->dw};
    for (float* p : ptrs) cudaFree(p);
    for (auto& t : m->tc) cudaFree(t.img);
    cudaFree(m->whh_tc);
    vad_front_destroy(m->fused);
    delete m;
    return OSB_OK;
}

int osb_vad_set_gemm(void* handle, int use_tcgen05) {
    OSB_REQUIRE(handle, "null VAD handle");
    OSB_REQUIRE(use_tcgen05 >= 0 && use_tcgen05 <= 2, "mode must be 0 (FFMA), 1 (tcgen05 per layer) or 2 (fused tcgen05)");
    reinterpret_cast<VadModel*>(handle)->use_tc = use_tcgen05;
    return OSB_OK;
}

int osb_vad_set_recurrence(void* handle, int tensor_pipe) {
    OSB_REQUIRE(handle, "null VAD handle");
    OSB_REQUIRE(tensor_pipe >= 0 && tensor_pipe <= 2, "mode must be 0 (FP32 FFMA kernels), 1 (tensor pipe, h as fp16 hi + lo) or 2 (tensor pipe, h as one fp16 plane)");
    reinterpret_cast<VadModel*>(handle)->recur_tc = tensor_pipe;
    return OSB_OK;
}

int osb_vad_score_dev(void* handle, const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, float* d_state,
                      float* d_probs, int64_t probs_stride, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(handle, "null VAD handle");
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n, "bad sizes");
    if (n / kWin == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(d_audio && d_state && d_probs && probs_stride >= n / kWin, "bad buffers");
    return vad_score(reinterpret_cast<VadModel*>(handle), d_audio, fmt, n, batch, stride, d_state, d_probs, probs_stride, (cudaStream_t)stream);
}

int osb_vad_segments_dev(const float* d_probs, int64_t probs_stride, int64_t n_win, int64_t batch, int64_t n_samples,
                         float threshold, int min_speech_ms, int silence_ms, int32_t* d_segments, int32_t* d_counts, int max_seg,
                         void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(batch >= 0 && n_win >= 0 && max_seg >= 0, "bad sizes");
    if (batch == 0) return OSB_OK;
    OSB_REQUIRE(d_counts && (d_probs || n_win == 0) && (d_segments || max_seg == 0), "null buffer");
    return vad_segments(d_probs, probs_stride, n_win, batch, n_samples, threshold, min_speech_ms, silence_ms, d_segments, d_counts, max_seg,
                        (cudaStream_t)stream);
}

int osb_vad_score_host(void* handle, const void* audio, int fmt, int64_t n, float* state, float* probs, float* max_prob) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(handle && state, "null argument");
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    if (max_prob) *max_prob = 0.f;
    const long long n_win = n / kWin;
    if (n_win <= 0) return OSB_OK;
    const size_t es = fmt == OSB_FMT_PCM16 ? 2 : 4;
    void *da, *dst, *dp;
    if ((rc = ws.dev_buf(0, (size_t)n * es, &da)) || (rc = ws.dev_buf(1, 2 * kHid * 4, &dst)) || (rc = ws.dev_buf(2, (size_t)n_win * 4, &dp))) return rc;
    if ((rc = ws.h2d(da, audio, (size_t)n_win * kWin * es))) return rc;
    OSB_CUDA(cudaMemcpyAsync(dst, state, 2 * kHid * 4, cudaMemcpyHostToDevice, ws.stream));
    if ((rc = vad_score(reinterpret_cast<VadModel*>(handle), da, fmt, n, 1, n, (float*)dst, (float*)dp, n_win, ws.stream))) return rc;
    OSB_CUDA(cudaMemcpyAsync(state, dst, 2 * kHid * 4, cudaMemcpyDeviceToHost, ws.stream));
    std::vector<float> tmp;
    float* out = probs;
    if (!out) { tmp.resize((size_t)n_win); out = tmp.data(); }
    if ((rc = ws.d2h(out, dp, (size_t)n_win * 4))) return rc;
    if (max_prob) {
        float mx = 0.f;
        for (long long i = 0; i < n_win; ++i) if (out[i] > mx) mx = out[i];
        *max_prob = mx;
    }
    return OSB_OK;
}

int osb_vad_extract_speech_host(void* handle, const int16_t* pcm, int64_t n, int rate, float threshold, int min_speech_ms, int silence_ms,
                                int16_t* out, int64_t* out_n, int* n_segments) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(handle && pcm && out && out_n && rate >= 1000, "bad arguments");
    *out_n = 0;
    if (n_segments) *n_segments = 0;
    if (n <= 0) return OSB_OK;
    int up = 1, down = 1;
    long long n16 = n;
    if (rate != 16000) {
        int a = 16000, b = rate;
        while (b) { const int t = a % b; a = b; b = t; }
        up = 16000 / a; down = rate / a;
        n16 = (n * up + down - 1) / down;
        OSB_REQUIRE(n >= 2, "need at least two samples to resample");
    }
    const long long n_win = n16 / kWin;
    const int max_seg = (int)(n_win / 2 + 2);  // a segment needs a speech window and a silence window: never more than this
    void *d_in, *d_16k, *d_misc, *d_out;
    const size_t misc = 2 * kHid * 4 + (size_t)(n_win + 1) * 4 + (size_t)max_seg * 8 + 64 + (size_t)max_seg * 24 + 64;
    if ((rc = ws.dev_buf(0, (size_t)n * 2 + 16, &d_in)) || (rc = ws.dev_buf(1, (size_t)n16 * 2 + 16, &d_16k)) ||
        (rc = ws.dev_buf(2, misc, &d_misc)) || (rc = ws.dev_buf(3, (size_t)n * 2 + 16, &d_out))) return rc;
    if ((rc = ws.h2d(d_in, pcm, (size_t)n * 2))) return rc;
    const int16_t* a16 = (const int16_t*)d_in;
    if (rate != 16000) {
        if ((rc = osb_resample_poly_dev((const int16_t*)d_in, (int16_t*)d_16k, n, 1, n, n16, up, down, ws.stream))) return rc;
        a16 = (const int16_t*)d_16k;
    }
    float* d_state = (float*)d_misc;
    float* d_probs = d_state + 2 * kHid;
    int32_t* d_segs = (int32_t*)(d_probs + n_win + 1);
    int32_t* d_cnt = d_segs + (size_t)max_seg * 2;
    long long* d_plan = (long long*)(((uintptr_t)(d_cnt + 4) + 15) & ~(uintptr_t)15);
    long long* d_total = d_plan + (size_t)max_seg * 3;
    OSB_CUDA(cudaMemsetAsync(d_state, 0, 2 * kHid * 4, ws.stream));  // a fresh SileroVAD per call (stt_handler.py:68-71)
    if ((rc = vad_score(reinterpret_cast<VadModel*>(handle), a16, OSB_FMT_PCM16, n16, 1, n16, d_state, d_probs, n_win > 0 ? n_win : 1, ws.stream))) return rc;
    if ((rc = vad_segments(d_probs, n_win > 0 ? n_win : 1, n_win, 1, n16, threshold, min_speech_ms, silence_ms, d_segs, d_cnt, max_seg, ws.stream))) return rc;
    OSB_LAUNCH(k_vad_gather_plan, 1, 32, 0, ws.stream, d_segs, d_cnt, max_seg, (long long)n, rate / 1000, d_plan, d_total);
    OSB_CHECK_LAUNCH();
    long long tot[2] = {0, 0};
    OSB_CUDA(cudaMemcpyAsync(tot, d_total, sizeof(tot), cudaMemcpyDeviceToHost, ws.stream));
    if ((rc = ws.sync())) return rc;
    if (n_segments) *n_segments = (int)tot[1];
    if (tot[1] == 0 || tot[0] == 0) return OSB_OK;  // caller keeps the original audio (stt_handler.py:88-90, :112-115)
    long long per = (n / (tot[1] > 0 ? tot[1] : 1) / 8 + 255) / 256;
    if (per < 1) per = 1;
    if (per > 64) per = 64;
    OSB_LAUNCH(k_vad_gather, dim3((unsigned)per, (unsigned)tot[1]), 256, 0, ws.stream, (const int16_t*)d_in, d_plan, d_total, (int16_t*)d_out);
    OSB_CHECK_LAUNCH();
    if ((rc = ws.d2h(out, d_out, (size_t)tot[0] * 2))) return rc;
    *out_n = tot[0];
    return OSB_OK;
}

int osb_vad_segments_host(void* handle, const void* audio, int fmt, int64_t n, float* state, float threshold, int min_speech_ms,
                          int silence_ms, int32_t* segments, int max_seg, int* n_seg) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(handle && state && n_seg && (segments || max_seg == 0), "null argument");
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    *n_seg = 0;
    if (n <= 0) return OSB_OK;
    const long long n_win = n / kWin;
    const size_t es = fmt == OSB_FMT_PCM16 ? 2 : 4;
    void *da, *dst, *dp, *dsg;
    if ((rc = ws.dev_buf(0, (size_t)n * es + 16, &da)) || (rc = ws.dev_buf(1, 2 * kHid * 4, &dst)) ||
        (rc = ws.dev_buf(2, (size_t)(n_win + 1) * 4, &dp)) || (rc = ws.dev_buf(3, ((size_t)max_seg * 2 + 4) * 4, &dsg))) return rc;
    if ((rc = ws.h2d(da, audio, (size_t)n_win * kWin * es))) return rc;
    OSB_CUDA(cudaMemcpyAsync(dst, state, 2 * kHid * 4, cudaMemcpyHostToDevice, ws.stream));
    if ((rc = vad_score(reinterpret_cast<VadModel*>(handle), da, fmt, n, 1, n, (float*)dst, (float*)dp, n_win > 0 ? n_win : 1, ws.stream))) return rc;
    int32_t* d_counts = reinterpret_cast<int32_t*>(dsg) + (size_t)max_seg * 2;
    if ((rc = vad_segments((const float*)dp, n_win > 0 ? n_win : 1, n_win, 1, n, threshold, min_speech_ms, silence_ms, (int32_t*)dsg, d_counts, max_seg, ws.stream))) return rc;
    OSB_CUDA(cudaMemcpyAsync(state, dst, 2 * kHid * 4, cudaMemcpyDeviceToHost, ws.stream));
    std::vector<int32_t> buf((size_t)max_seg * 2 + 1);
    if ((rc = ws.d2h(buf.data(), dsg, ((size_t)max_seg * 2 + 1) * 4))) return rc;
    int cnt = buf[(size_t)max_seg * 2];
    if (cnt > max_seg) {
        set_error("segment buffer too small: %d segments, capacity %d", cnt, max_seg);
        return OSB_ERR_BUFFER;
    }
    memcpy(segments, buf.data(), (size_t)cnt * 2 * 4);
    *n_seg = cnt;
    return OSB_OK;
}

}  // extern "C"
