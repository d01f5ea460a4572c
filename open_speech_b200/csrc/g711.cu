// G.711 mu-law / A-law <-> PCM16 and the realtime path's linear resampler.
//
// Replaces (reference file:line):
//   audioop.ulaw2lin / alaw2lin / lin2ulaw / lin2alaw   src/realtime/audio_buffer.py:52,55,76,79
//   _resample_linear (np.interp over linspace grids)    src/realtime/audio_buffer.py:20-34
//
// HBM-bound byte movers: 128-bit loads/stores, closed-form codec arithmetic (no table, so no
// divergent constant-cache or shared-memory bank traffic), grid = multiple of 148 CTAs.
// The interpolation is IEEE f64 with explicit _rn intrinsics so that nvcc cannot contract
// mul+add into an FMA: numpy's arr_interp computes slope*(x-xp[i]) + fp[i] as two roundings.
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace osb {

__device__ __forceinline__ int ulaw2lin(uint32_t b) {
    uint32_t u = ~b & 0xFFu;
    int t = (int)(((u & 0x0Fu) << 3) + 0x84u);
    t <<= (u & 0x70u) >> 4;
    return (u & 0x80u) ? (0x84 - t) : (t - 0x84);
}

__device__ __forceinline__ int alaw2lin(uint32_t b) {
    uint32_t a = (b ^ 0x55u) & 0xFFu;
    int t = (int)((a & 0x0Fu) << 4);
    int seg = (int)((a & 0x70u) >> 4);
    t = (seg == 0) ? (t + 8) : ((seg == 1) ? (t + 0x108) : ((t + 0x108) << (seg - 1)));
    return (a & 0x80u) ? t : -t;
}

// audioop.lin2ulaw(width=2): st_14linear2ulaw(sample >> 2)
__device__ __forceinline__ uint32_t lin2ulaw(int s) {
    int v = s >> 2;
    uint32_t mask = 0xFFu;
    if (v < 0) { v = -v; mask = 0x7Fu; }
    v = min(v, 8159) + 0x21;
    int seg = max(0, 26 - __clz(v));  // first i with v <= (0x40<<i)-1
    if (seg >= 8) return 0x7Fu ^ mask;
    uint32_t uval = (uint32_t)(seg << 4) | ((uint32_t)(v >> (seg + 1)) & 0xFu);
    return uval ^ mask;
}

// audioop.lin2alaw(width=2): st_linear2alaw(sample >> 3)
__device__ __forceinline__ uint32_t lin2alaw(int s) {
    int v = s >> 3;
    uint32_t mask = 0xD5u;
    if (v < 0) { mask = 0x55u; v = -v - 1; }
    int seg = (v == 0) ? 0 : max(0, 27 - __clz(v));  // first i with v <= (0x20<<i)-1
    uint32_t aval = (uint32_t)(seg << 4) | ((uint32_t)((seg < 2) ? (v >> 1) : (v >> seg)) & 0xFu);
    return aval ^ mask;
}

template <int LAW>
__device__ __forceinline__ int expand(uint32_t b) { return LAW == OSB_FMT_ULAW ? ulaw2lin(b) : alaw2lin(b); }
template <int LAW>
__device__ __forceinline__ uint32_t compress(int s) { return LAW == OSB_FMT_ULAW ? lin2ulaw(s) : lin2alaw(s); }

// ---------------------------------------------------------------- flat decode / encode
template <int LAW>
__global__ void __launch_bounds__(256) k_g711_decode(const uint8_t* __restrict__ in, int16_t* __restrict__ out, size_t n) {
    size_t nvec = n / 16;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < nvec; v += nthr) {
        uint4 w = ld_stream_u4(in + v * 16);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int s0 = expand<LAW>(ws[k] & 0xFF), s1 = expand<LAW>((ws[k] >> 8) & 0xFF);
            int s2 = expand<LAW>((ws[k] >> 16) & 0xFF), s3 = expand<LAW>(ws[k] >> 24);
            o[2 * k] = (uint32_t)(s0 & 0xFFFF) | ((uint32_t)s1 << 16);
            o[2 * k + 1] = (uint32_t)(s2 & 0xFFFF) | ((uint32_t)s3 << 16);
        }
        st_stream_u4(out + v * 16, make_uint4(o[0], o[1], o[2], o[3]));
        st_stream_u4(out + v * 16 + 8, make_uint4(o[4], o[5], o[6], o[7]));
    }
    for (size_t i = nvec * 16 + tid; i < n; i += nthr) out[i] = (int16_t)expand<LAW>(in[i]);
}

template <int LAW>
__global__ void __launch_bounds__(256) k_g711_encode(const int16_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    size_t nvec = n / 16;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < nvec; v += nthr) {
        uint4 a = ld_stream_u4(in + v * 16), b = ld_stream_u4(in + v * 16 + 8);
        uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w0 = ws[2 * k], w1 = ws[2 * k + 1];
            o[k] = compress<LAW>((int)(int16_t)(w0 & 0xFFFF)) | (compress<LAW>((int)(int16_t)(w0 >> 16)) << 8) |
                   (compress<LAW>((int)(int16_t)(w1 & 0xFFFF)) << 16) | (compress<LAW>((int)(int16_t)(w1 >> 16)) << 24);
        }
        st_stream_u4(out + v * 16, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (size_t i = nvec * 16 + tid; i < n; i += nthr) out[i] = (uint8_t)compress<LAW>((int)in[i]);
}

// ---------------------------------------------------------------- linear resample (np.interp)
struct LinArgs {
    const void* in;
    void* out;
    long long n_in, n_out, batch, in_stride, out_stride;
    double step_o, step_n;  // 1/(n_in-1), 1/(n_out-1) formed in f64 on the host like np.linspace
    int exact_int;          // 1: the integer fast path below is provably bit-identical for these sizes
    const unsigned* tab;    // exact_int: per output j, (floor(p) << 16) | (1 << 15 if interior coincidence) | remainder r
    unsigned magic;            // floor(2^(31+l) / (m-1)) + 1, l = ceil(log2(m-1)): exact floor division of any dividend < 2^31 by m-1 < 2^15
    int shift;                 // l - 1:  U / (m-1) == __umulhi(U, magic) >> shift   (Granlund-Montgomery, N = 31)
};

template <int IN_FMT>
__device__ __forceinline__ int load_in(const void* base, long long idx) {
    if (IN_FMT == OSB_FMT_PCM16) return (int)__ldg(reinterpret_cast<const int16_t*>(base) + idx);
    return expand<IN_FMT>((uint32_t)__ldg(reinterpret_cast<const uint8_t*>(base) + idx));
}

__device__ __forceinline__ double grid_pt(long long i, long long last, double step) {
    return (i == last) ? 1.0 : __dmul_rn((double)i, step);  // np.linspace: arange*step, end forced
}

template <int IN_FMT>
__device__ __forceinline__ int interp_one(const void* in, long long j, const LinArgs& a) {
    const long long n = a.n_in, m = a.n_out;
    if (n == 1) return load_in<IN_FMT>(in, 0);
    const double xn = (m == 1) ? 0.0 : grid_pt(j, m - 1, a.step_n);
    long long i = (long long)(xn * (double)(n - 1));
    i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    double xo = grid_pt(i, n - 1, a.step_o);
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {  // candidate bucket is off by at most one
        if (xo > xn) {
            --i;
            xo = grid_pt(i, n - 1, a.step_o);
        } else if (i < n - 1) {
            double xo1 = grid_pt(i + 1, n - 1, a.step_o);
            if (xo1 <= xn) { ++i; xo = xo1; }
        }
    }
    const int f0 = load_in<IN_FMT>(in, i);
    if (i == n - 1 || xo == xn) return f0;
    const int f1 = load_in<IN_FMT>(in, i + 1);
    const double xo1 = grid_pt(i + 1, n - 1, a.step_o);
    const double slope = __ddiv_rn((double)(f1 - f0), __dsub_rn(xo1, xo));
    const double y = __dadd_rn(__dmul_rn(slope, __dsub_rn(xn, xo)), (double)f0);
    return __double2int_rz(y);  // astype(int16): truncation toward zero (|y| <= 32768)
}

// Integer fast path.  In exact arithmetic output j sits at p = j (n-1) / (m-1) = i + r / (m-1) and is y = f0 + (f1 - f0) r / (m-1), a rational
// with denominator m-1.  numpy's float64 evaluation differs from it by < 3e-11 (n-1) (two roundings on each grid point, a division, a
// multiply and an add on |values| <= 65535), so whenever y is NOT an integer -- at least 1 / (m-1) away from one -- truncating the
// exact value gives numpy's int16 bit for bit; the host enables this path only for m <= 32768 and (n-1)(m-1) <= 1e8, a margin of > 100x.
// The same margin makes the float64 bucket search land on i = floor(p) whenever r != 0.  Left to the float64 path (returns false):
// interior grid coincidences (r == 0, flagged in the table) and exactly integral interpolants with f1 != f0 (~1 % of the outputs for
// 160 -> 320), where the sign of a 1e-12 rounding error decides the truncation.  The two ends are exact in numpy as well (xn == xo).
// Per output: one table word (i, r: the same for every chunk), two samples, one multiply-shift division by the invariant m-1.
template <int IN_FMT>
__device__ __forceinline__ bool interp_int(const void* in, long long j, const LinArgs& a, int& y) {
    const unsigned e = __ldg(a.tab + j);
    const unsigned i = e >> 16, r = e & 0x7FFFu;
    if (e & 0x8000u) return false;                 // interior coincidence of the two grids
    const int f0 = load_in<IN_FMT>(in, i);
    if (r == 0) { y = f0; return true; }           // j = 0 or j = m-1
    const int f1 = load_in<IN_FMT>(in, i + 1);
    const unsigned den = (unsigned)(a.n_out - 1);
    // U = (y + 32768) (m-1) >= 0 and < 2^31; q = floor(U / den) by one high multiply and a shift
    const unsigned U = (unsigned)(f0 + 32768) * den + (unsigned)((f1 - f0) * (int)r);
    const unsigned q = __umulhi(U, a.magic) >> a.shift;
    const unsigned rem = U - q * den;
    if (rem == 0 && f1 != f0) return false;        // exactly integral: float64's rounding error decides
    const int yf = (int)q - 32768;                 // floor of the exact value
    y = (yf >= 0 || rem == 0) ? yf : yf + 1;       // astype(int16) truncates toward zero
    return true;
}

template <int IN_FMT, int OUT_FMT>
__global__ void __launch_bounds__(256) k_resample_linear(LinArgs a, int vec_ok) {
    const long long groups = (a.n_out + 7) / 8;
    const long long total = groups * a.batch;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const int in_es = (IN_FMT == OSB_FMT_PCM16) ? 2 : 1;
    for (long long g = tid; g < total; g += nthr) {
        const long long c = g / groups, j0 = (g - c * groups) * 8;
        const void* in = reinterpret_cast<const char*>(a.in) + c * a.in_stride * in_es;
        int v[8];
        if (a.exact_int) {
            unsigned slow = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = 0;
                if (j0 + k < a.n_out && !interp_int<IN_FMT>(in, j0 + k, a, v[k])) slow |= 1u << k;
            }
            // the rare float64 evaluations: every pass, each lane that still has one takes its lowest; ~1 pass per warp iteration
            while (__any_sync(0xffffffffu, slow != 0)) {
                if (slow) {
                    const int k = __ffs(slow) - 1;
                    slow &= slow - 1;
                    const int y = interp_one<IN_FMT>(in, j0 + k, a);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        if (kk == k) v[kk] = y;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (j0 + k < a.n_out) ? interp_one<IN_FMT>(in, j0 + k, a) : 0;
        }
        if (OUT_FMT == OSB_FMT_PCM16) {
            int16_t* o = reinterpret_cast<int16_t*>(a.out) + c * a.out_stride + j0;
            if (vec_ok && j0 + 8 <= a.n_out) {
                uint4 w;
                w.x = (uint32_t)(v[0] & 0xFFFF) | ((uint32_t)v[1] << 16);
                w.y = (uint32_t)(v[2] & 0xFFFF) | ((uint32_t)v[3] << 16);
                w.z = (uint32_t)(v[4] & 0xFFFF) | ((uint32_t)v[5] << 16);
                w.w = (uint32_t)(v[6] & 0xFFFF) | ((uint32_t)v[7] << 16);
                st_stream_u4(o, w);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (j0 + k < a.n_out) o[k] = (int16_t)v[k];
            }
        } else {
            uint8_t* o = reinterpret_cast<uint8_t*>(a.out) + c * a.out_stride + j0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (j0 + k < a.n_out) o[k] = (uint8_t)compress<OUT_FMT>((int)(int16_t)v[k]);
        }
    }
}

// Tiled variant for many short rows (the realtime door replayed over whole recordings: 768,000 chunks of 160 -> 320 per 256 x 60 s).
// A CTA takes R dense rows at a time: the wire bytes are expanded ONCE into shared memory (the per-output kernel above expands two
// samples per output: 4x the work for 2x upsampling), the (i, r) table is read from shared memory, the results are assembled in shared
// memory and leave as whole 16-byte stores, and the ~1 % of outputs that need the float64 evaluation are queued and then worked off by
// consecutive threads instead of stalling a warp per straggler.  Same arithmetic (interp_int / interp_one): bit-identical.
constexpr int kLinTileRows = 32, kLinSlowCap = 1024;

// the queue of float64 evaluations is full (adversarial input: it holds 10 % of a 160 -> 320 tile, real audio needs ~1 %): evaluate in
// place, out of line so that the eight unrolled outputs of the main loop do not each carry a copy of the float64 path
// (scalars by value: a reference to the kernel's argument block would force the whole block into local memory)
template <int IN_FMT>
__device__ __noinline__ int interp_one_overflow(const void* in, long long j, long long n_in, long long n_out, double step_o, double step_n) {
    LinArgs a;
    a.n_in = n_in; a.n_out = n_out; a.step_o = step_o; a.step_n = step_n;
    return interp_one<IN_FMT>(in, j, a);
}

template <int IN_FMT>
__global__ void __launch_bounds__(256) k_resample_linear_tiled(LinArgs a, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char lsm[];
    const int n_in = (int)a.n_in, n_out = (int)a.n_out;
    unsigned* tab = reinterpret_cast<unsigned*>(lsm);                                     // [n_out]
    int16_t* outb = reinterpret_cast<int16_t*>(tab + n_out);                              // [R][n_out]   (n_out % 8 == 0: 16-byte rows)
    int16_t* samp = outb + kLinTileRows * n_out;                                          // [R][n_in]
    unsigned* slow = reinterpret_cast<unsigned*>(samp + ((kLinTileRows * n_in + 2) & ~1));  // [kLinSlowCap]  row << 16 | j  (one spare sample in between)
    __shared__ int slow_n;
    const int tid = threadIdx.x;
    constexpr int in_es = (IN_FMT == OSB_FMT_PCM16) ? 2 : 1;
    for (int j = tid; j < n_out; j += 256) tab[j] = a.tab[j];
    const int gpr = n_out / 8;  // 8-output groups per row
    const unsigned den = (unsigned)(n_out - 1), magic = a.magic;
    const int shift = a.shift;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * kLinTileRows;
        const int rows = (int)((a.batch - row0) < kLinTileRows ? (a.batch - row0) : kLinTileRows);
        const unsigned char* inb = reinterpret_cast<const unsigned char*>(a.in) + row0 * n_in * in_es;
        if (tid == 0) slow_n = 0;
        // expand / copy the tile's input once: 4 samples per thread and pass (rows are dense and the tile starts 4-byte aligned)
        const int words = rows * n_in * in_es / 4;
        for (int w = tid; w < words; w += 256) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(inb) + w);
            if (IN_FMT == OSB_FMT_PCM16) {
                reinterpret_cast<uint32_t*>(samp)[w] = v;
            } else {
                const uint32_t lo = (uint32_t)(expand<IN_FMT>(v & 0xFFu) & 0xFFFF) | ((uint32_t)expand<IN_FMT>((v >> 8) & 0xFFu) << 16);
                const uint32_t hi = (uint32_t)(expand<IN_FMT>((v >> 16) & 0xFFu) & 0xFFFF) | ((uint32_t)expand<IN_FMT>(v >> 24) << 16);
                reinterpret_cast<uint2*>(samp)[w] = make_uint2(lo, hi);
            }
        }
        __syncthreads();
        const int groups = rows * gpr;
        for (int g = tid; g < groups; g += 256) {
            const int r = g / gpr, j0 = (g - r * gpr) * 8;
            const int16_t* sr = samp + r * n_in;
            const uint4 e0 = *reinterpret_cast<const uint4*>(tab + j0), e1 = *reinterpret_cast<const uint4*>(tab + j0 + 4);
            const unsigned e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            int v[8];
            unsigned slowmask = 0;
            // straight-line: every output loads both neighbours (samp has one spare element behind the last row) and runs the division;
            // what the table or the remainder hands over to float64 is collected in a mask and queued after the loop
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned i = e[k] >> 16, rr = e[k] & 0x7FFFu;
                const int f0 = sr[i], f1 = sr[i + 1];
                const unsigned U = (unsigned)(f0 + 32768) * den + (unsigned)((f1 - f0) * (int)rr);
                const unsigned q = __umulhi(U, magic) >> shift;
                const unsigned rem = U - q * den;
                const int yf = (int)q - 32768;
                v[k] = (yf >= 0 || rem == 0) ? yf : yf + 1;  // rr == 0: U = (f0 + 32768) den, so q - 32768 = f0 exactly
                // interior coincidence of the two grids (flag), or an exactly integral interpolant between different samples
                const bool hand_over = (e[k] & 0x8000u) || (rr != 0 && rem == 0 && f1 != f0);
                slowmask |= hand_over ? (1u << k) : 0u;
            }
            while (slowmask) {
                const int k = __ffs(slowmask) - 1;
                slowmask &= slowmask - 1;
                const int pos = atomicAdd(&slow_n, 1);
                if (pos < kLinSlowCap) {
                    slow[pos] = ((unsigned)r << 16) | (unsigned)(j0 + k);
                } else {
                    const int y = interp_one_overflow<IN_FMT>(inb + (long long)r * n_in * in_es, j0 + k, a.n_in, a.n_out, a.step_o, a.step_n);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        if (kk == k) v[kk] = y;
                }
            }
            uint4 w;
            w.x = (uint32_t)(v[0] & 0xFFFF) | ((uint32_t)v[1] << 16);
            w.y = (uint32_t)(v[2] & 0xFFFF) | ((uint32_t)v[3] << 16);
            w.z = (uint32_t)(v[4] & 0xFFFF) | ((uint32_t)v[5] << 16);
            w.w = (uint32_t)(v[6] & 0xFFFF) | ((uint32_t)v[7] << 16);
            *reinterpret_cast<uint4*>(outb + r * n_out + j0) = w;
        }
        __syncthreads();
        {   // the float64 evaluations, one per thread
            const int ns = slow_n < kLinSlowCap ? slow_n : kLinSlowCap;
            for (int sidx = tid; sidx < ns; sidx += 256) {
                const unsigned it = slow[sidx];
                const int r = (int)(it >> 16), j = (int)(it & 0xFFFFu);
                outb[r * n_out + j] = (int16_t)interp_one<IN_FMT>(inb + (long long)r * n_in * in_es, j, a);
            }
        }
        __syncthreads();
        {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<int16_t*>(a.out) + row0 * n_out);
            const int vecs = rows * n_out / 8;
            for (int w = tid; w < vecs; w += 256) st_stream_u4(o + w, reinterpret_cast<const uint4*>(outb)[w]);
        }
        // (the next tile's expansion writes samp and resets the queue counter; outb is rewritten only after its next barrier)
    }
}

template <int IN_FMT>
static int launch_linear_out(const LinArgs& a, int out_fmt, int vec_ok, int grid, cudaStream_t st) {
    switch (out_fmt) {
        case OSB_FMT_PCM16: OSB_LAUNCH((k_resample_linear<IN_FMT, OSB_FMT_PCM16>), grid, 256, 0, st, a, vec_ok); break;
        case OSB_FMT_ULAW: OSB_LAUNCH((k_resample_linear<IN_FMT, OSB_FMT_ULAW>), grid, 256, 0, st, a, vec_ok); break;
        case OSB_FMT_ALAW: OSB_LAUNCH((k_resample_linear<IN_FMT, OSB_FMT_ALAW>), grid, 256, 0, st, a, vec_ok); break;
        default: set_error("invalid argument: out_fmt %d", out_fmt); return OSB_ERR_INVALID_ARG;
    }
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

// (i, r) table of the integer fast path, per (device, n_in, n_out); built once, a few hundred words for the realtime sizes
struct LinTab { unsigned* d = nullptr; };
static std::mutex g_lin_mu;
static std::map<unsigned long long, LinTab> g_lin_tabs;

static int linear_table(long long n, long long m, const unsigned** out) {
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    const unsigned long long key = ((unsigned long long)dev << 56) ^ ((unsigned long long)n << 24) ^ (unsigned long long)m;
    std::lock_guard<std::mutex> lk(g_lin_mu);
    auto it = g_lin_tabs.find(key);
    if (it == g_lin_tabs.end()) {
        std::vector<unsigned> h((size_t)m);
        for (long long j = 0; j < m; ++j) {
            const long long num = j * (n - 1), i = num / (m - 1), r = num - i * (m - 1);
            const bool interior = (r == 0) && j != 0 && j != m - 1;
            h[(size_t)j] = ((unsigned)i << 16) | (interior ? 0x8000u : 0u) | (unsigned)r;
        }
        LinTab t;
        OSB_CUDA(cudaMalloc(&t.d, h.size() * sizeof(unsigned)));
        OSB_CUDA(cudaMemcpy(t.d, h.data(), h.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        it = g_lin_tabs.emplace(key, t).first;
    }
    *out = it->second.d;
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_g711_decode_dev(const uint8_t* d_in, int16_t* d_out, size_t n, int law, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(law == OSB_FMT_ULAW || law == OSB_FMT_ALAW, "law must be OSB_FMT_ULAW or OSB_FMT_ALAW");
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
    int grid = grid_for(n / 16 + 1, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (law == OSB_FMT_ULAW) OSB_LAUNCH(k_g711_decode<OSB_FMT_ULAW>, grid, 256, 0, st, d_in, d_out, n);
    else OSB_LAUNCH(k_g711_decode<OSB_FMT_ALAW>, grid, 256, 0, st, d_in, d_out, n);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_g711_encode_dev(const int16_t* d_in, uint8_t* d_out, size_t n, int law, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(law == OSB_FMT_ULAW || law == OSB_FMT_ALAW, "law must be OSB_FMT_ULAW or OSB_FMT_ALAW");
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
    int grid = grid_for(n / 16 + 1, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (law == OSB_FMT_ULAW) OSB_LAUNCH(k_g711_encode<OSB_FMT_ULAW>, grid, 256, 0, st, d_in, d_out, n);
    else OSB_LAUNCH(k_g711_encode<OSB_FMT_ALAW>, grid, 256, 0, st, d_in, d_out, n);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_resample_linear_dev(const void* d_in, int in_fmt, void* d_out, int out_fmt, int64_t n_in, int64_t n_out,
                            int64_t batch, int64_t in_stride, int64_t out_stride, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(n_in >= 0 && n_out >= 0 && batch >= 0, "negative size");
    if (n_in == 0 || n_out == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(in_stride >= n_in && out_stride >= n_out, "stride smaller than row");
    LinArgs a;
    a.in = d_in; a.out = d_out; a.n_in = n_in; a.n_out = n_out; a.batch = batch;
    a.in_stride = in_stride; a.out_stride = out_stride;
    a.step_o = n_in > 1 ? 1.0 / (double)(n_in - 1) : 0.0;
    a.step_n = n_out > 1 ? 1.0 / (double)(n_out - 1) : 0.0;
    a.exact_int = (n_in > 1 && n_in <= 65536 && n_out > 1 && n_out <= 32768 && (double)(n_in - 1) * (double)(n_out - 1) <= 1.0e8) ? 1 : 0;
    if (const char* e = getenv("OSB_LINEAR_F64")) a.exact_int = (e[0] == '1') ? 0 : a.exact_int;  // cross-check switch
    a.tab = nullptr; a.magic = 0; a.shift = 0;
    if (a.exact_int) {
        if ((rc = linear_table(n_in, n_out, &a.tab))) return rc;
        int l = 0;
        while ((1ll << l) < n_out - 1) ++l;
        if (l >= 1) {  // n_out == 2 has no interior output: the division is never reached
            a.shift = l - 1;
            a.magic = (unsigned)((1ull << (31 + l)) / (unsigned long long)(n_out - 1) + 1ull);
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    // many short dense rows: the tiled kernel (a single tick of the realtime path keeps the per-output kernel: more CTAs, lower latency)
    if (a.exact_int && out_fmt == OSB_FMT_PCM16 && in_stride == n_in && out_stride == n_out && n_out % 8 == 0 && n_out <= 4096 && n_in <= 4096 &&
        (n_in * (in_fmt == OSB_FMT_PCM16 ? 2 : 1)) % 4 == 0 && ((uintptr_t)d_in & 3) == 0 && ((uintptr_t)d_out & 15) == 0 &&
        batch >= 16384 && !getenv("OSB_LINEAR_NO_TILES")) {
        const size_t smem = (size_t)n_out * 4 + (size_t)kLinTileRows * n_out * 2 + (size_t)((kLinTileRows * n_in + 2) & ~1ll) * 2 + kLinSlowCap * 4;
        if (smem <= 200 * 1024) {
            const long long n_tiles = (batch + kLinTileRows - 1) / kLinTileRows;
            const int per_sm = (int)std::min<size_t>(8, (220 * 1024) / (smem + 1024));
            const long long want = (long long)num_sms() * (per_sm < 1 ? 1 : per_sm);
            const unsigned grid_t = (unsigned)(n_tiles < want ? n_tiles : want);
            static PerDeviceOnce once;
            OSB_CUDA(once.run([&] {
                cudaError_t e = cudaFuncSetAttribute(k_resample_linear_tiled<OSB_FMT_PCM16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resample_linear_tiled<OSB_FMT_ULAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resample_linear_tiled<OSB_FMT_ALAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                return e;
            }));
            switch (in_fmt) {
                case OSB_FMT_PCM16: OSB_LAUNCH(k_resample_linear_tiled<OSB_FMT_PCM16>, grid_t, 256, smem, st, a, n_tiles); break;
                case OSB_FMT_ULAW: OSB_LAUNCH(k_resample_linear_tiled<OSB_FMT_ULAW>, grid_t, 256, smem, st, a, n_tiles); break;
                case OSB_FMT_ALAW: OSB_LAUNCH(k_resample_linear_tiled<OSB_FMT_ALAW>, grid_t, 256, smem, st, a, n_tiles); break;
                default: set_error("invalid argument: in_fmt %d", in_fmt); return OSB_ERR_INVALID_ARG;
            }
            OSB_CHECK_LAUNCH();
            return OSB_OK;
        }
    }
    int vec_ok = (out_fmt == OSB_FMT_PCM16) && (((uintptr_t)d_out & 15) == 0) && (out_stride % 8 == 0);
    long long groups = (n_out + 7) / 8 * batch;
    int grid = grid_for((size_t)groups, 256);
    switch (in_fmt) {
        case OSB_FMT_PCM16: return launch_linear_out<OSB_FMT_PCM16>(a, out_fmt, vec_ok, grid, st);
        case OSB_FMT_ULAW: return launch_linear_out<OSB_FMT_ULAW>(a, out_fmt, vec_ok, grid, st);
        case OSB_FMT_ALAW: return launch_linear_out<OSB_FMT_ALAW>(a, out_fmt, vec_ok, grid, st);
    }
    set_error("invalid argument: in_fmt %d", in_fmt);
    return OSB_ERR_INVALID_ARG;
}

// ---------------------------------------------------------------- host-pointer wrappers
int osb_g711_decode_host(const uint8_t* in, int16_t* out, size_t n, int law) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, n, &di)) || (rc = ws.dev_buf(1, n * 2, &dout))) return rc;
    if ((rc = ws.h2d(di, in, n))) return rc;
    if ((rc = osb_g711_decode_dev((const uint8_t*)di, (int16_t*)dout, n, law, ws.stream))) return rc;
    return ws.d2h(out, dout, n * 2);
}

int osb_g711_encode_host(const int16_t* in, uint8_t* out, size_t n, int law) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, n * 2, &di)) || (rc = ws.dev_buf(1, n, &dout))) return rc;
    if ((rc = ws.h2d(di, in, n * 2))) return rc;
    if ((rc = osb_g711_encode_dev((const int16_t*)di, (uint8_t*)dout, n, law, ws.stream))) return rc;
    return ws.d2h(out, dout, n);
}

int osb_resample_linear_host(const void* in, int in_fmt, void* out, int out_fmt, int64_t n_in, int64_t n_out,
                             int64_t batch, int64_t in_stride, int64_t out_stride) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n_in <= 0 || n_out <= 0 || batch <= 0) return OSB_OK;
    size_t ies = in_fmt == OSB_FMT_PCM16 ? 2 : 1, oes = out_fmt == OSB_FMT_PCM16 ? 2 : 1;
    size_t ib = (size_t)((batch - 1) * in_stride + n_in) * ies, ob = (size_t)((batch - 1) * out_stride + n_out) * oes;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, ib, &di)) || (rc = ws.dev_buf(1, ob, &dout))) return rc;
    if ((rc = ws.h2d(di, in, ib))) return rc;
    if ((rc = osb_resample_linear_dev(di, in_fmt, dout, out_fmt, n_in, n_out, batch, in_stride, out_stride, ws.stream))) return rc;
    return ws.d2h(out, dout, ob);
}

}  // extern "C"
