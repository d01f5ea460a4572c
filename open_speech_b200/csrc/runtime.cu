// libosb200 runtime: device selection, error text, per-thread host workspace.
#include <atomic>
#include <cstdarg>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace osb {

static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_sms{0};
static thread_local int t_device = -1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear sticky-less errors so later calls report their own
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return OSB_ERR_NO_DEVICE;
    return OSB_ERR_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-kernel timing (bench.py roofline): event pairs recorded on the launching stream
struct ProfRec { std::string name; cudaEvent_t a, b; };
static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::map<std::string, std::pair<double, long long>> g_prof_acc;  // name -> (ms, launches)
static thread_local cudaEvent_t t_prof_a = nullptr;
static thread_local const char* t_prof_name = nullptr;

bool prof_on() { return g_prof.load(std::memory_order_relaxed); }
void prof_begin(const char* name, cudaStream_t st) {
    cudaEventCreate(&t_prof_a);
    cudaEventRecord(t_prof_a, st);
    t_prof_name = name;
}
void prof_end(cudaStream_t st) {
    cudaEvent_t b;
    cudaEventCreate(&b);
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(ProfRec{t_prof_name ? t_prof_name : "?", t_prof_a, b});
}
static void prof_drain() {  // caller holds g_prof_mu and has synchronised the device
    for (auto& r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& e = g_prof_acc[r.name];
            e.first += ms;
            e.second += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
}

int num_sms() {
    int v = g_sms.load();
    return v > 0 ? v : OSB_NUM_SMS;
}

static thread_local int t_sm_budget = 0, t_sm_budget_uses = 0;
void set_sm_budget(int sms, int launches) {
    t_sm_budget = (sms > 0 && launches > 0) ? sms : 0;
    t_sm_budget_uses = t_sm_budget ? launches : 0;
}
int take_sm_budget() {
    const int n = num_sms(), b = t_sm_budget;
    if (t_sm_budget_uses > 0 && --t_sm_budget_uses == 0) t_sm_budget = 0;
    return (b > 0 && b < n) ? b : n;
}

int ensure_init() {
    if (t_device >= 0) return OSB_OK;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    return osb_init(dev);
}

int HostWs::prepare() {
    int rc = ensure_init();
    if (rc) return rc;
    if (device != t_device) {  // first use on this thread, or the thread switched device
        if (stream) {
            cudaStreamDestroy(stream);
            for (auto& p : pin) { if (p) cudaFreeHost(p); p = nullptr; }
            for (auto& d : dev) { if (d) cudaFree(d); d = nullptr; }
            pin_cap[0] = pin_cap[1] = 0;
            dev_cap[0] = dev_cap[1] = dev_cap[2] = dev_cap[3] = 0;
        }
        OSB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        device = t_device;
    }
    return OSB_OK;
}

static size_t round_up_cap(size_t bytes) {
    size_t c = 1 << 16;
    while (c < bytes) c <<= 1;
    return c;
}

int HostWs::dev_buf(int slot, size_t bytes, void** out) {
    if (bytes > dev_cap[slot]) {
        if (dev[slot]) { OSB_CUDA(cudaStreamSynchronize(stream)); OSB_CUDA(cudaFree(dev[slot])); dev[slot] = nullptr; dev_cap[slot] = 0; }
        size_t c = round_up_cap(bytes);
        OSB_CUDA(cudaMalloc(&dev[slot], c));
        dev_cap[slot] = c;
    }
    *out = dev[slot];
    return OSB_OK;
}

int HostWs::pin_buf(int slot, size_t bytes, void** out) {
    if (bytes > pin_cap[slot]) {
        if (pin[slot]) { OSB_CUDA(cudaStreamSynchronize(stream)); OSB_CUDA(cudaFreeHost(pin[slot])); pin[slot] = nullptr; pin_cap[slot] = 0; }
        size_t c = round_up_cap(bytes);
        OSB_CUDA(cudaMallocHost(&pin[slot], c));
        pin_cap[slot] = c;
    }
    *out = pin[slot];
    return OSB_OK;
}

static const size_t kPinLimit = (size_t)256 << 20;  // above this, copy straight from the caller's pages

static bool is_pinned(const void* h) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int HostWs::h2d(void* d, const void* h, size_t bytes) {
    if (bytes == 0) return OSB_OK;
    if (bytes <= kPinLimit && !is_pinned(h)) {
        void* p;
        int rc = pin_buf(0, bytes, &p);
        if (rc) return rc;
        memcpy(p, h, bytes);
        OSB_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, stream));
    } else {
        OSB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream));
    }
    return OSB_OK;
}

int HostWs::d2h(void* h, const void* d, size_t bytes) {
    if (bytes == 0) return sync();
    if (bytes <= kPinLimit && !is_pinned(h)) {
        void* p;
        int rc = pin_buf(1, bytes, &p);
        if (rc) return rc;
        OSB_CUDA(cudaMemcpyAsync(p, d, bytes, cudaMemcpyDeviceToHost, stream));
        OSB_CUDA(cudaStreamSynchronize(stream));
        memcpy(h, p, bytes);
    } else {
        OSB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream));
        OSB_CUDA(cudaStreamSynchronize(stream));
    }
    return OSB_OK;
}

int HostWs::sync() {
    OSB_CUDA(cudaStreamSynchronize(stream));
    return OSB_OK;
}

HostWs& host_ws() {
    static thread_local HostWs ws;
    return ws;
}

}  // namespace osb

extern "C" {

int osb_version(void) { return OSB_VERSION; }

int osb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int osb_init(int device) {
    int n = osb_device_count();
    if (n <= 0) {
        osb::set_error("no CUDA device visible: libosb200 has no CPU path");
        return OSB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        osb::set_error("invalid argument: device %d out of range (0..%d)", device, n - 1);
        return OSB_ERR_INVALID_ARG;
    }
    OSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OSB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        osb::set_error("device %d is sm_%d%d; libosb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return OSB_ERR_UNSUPPORTED;
    }
    osb::g_sms.store(prop.multiProcessorCount);
    // keep stream-ordered scratch cached in the pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    osb::t_device = device;
    return OSB_OK;
}

const char* osb_last_error(void) { return osb::t_err; }

uint64_t osb_launch_count(void) { return osb::g_launches.load(); }

namespace osb {
__global__ void __launch_bounds__(1024, 1) k_poison_smem(int words, unsigned int* sink) {
    extern __shared__ unsigned int psm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) psm[i] = 0x7FC00000u | (unsigned)(i & 0xFFFF);  // quiet NaNs
    __syncthreads();
    if (sink && psm[(threadIdx.x * 7919) % words] == 0u) *sink = 1;  // keeps the stores alive
}
}  // namespace osb

int osb_debug_poison_smem(void) {
    int rc = osb::ensure_init();
    if (rc) return rc;
    int dev = 0, max_smem = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    OSB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    OSB_CUDA(cudaFuncSetAttribute(osb::k_poison_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    OSB_CUDA(cudaDeviceSynchronize());
    osb::k_poison_smem<<<4 * osb::num_sms(), 1024, max_smem>>>(max_smem / 4, nullptr);  // one CTA per SM at a time, four waves
    OSB_CUDA(cudaGetLastError());
    OSB_CUDA(cudaDeviceSynchronize());
    return OSB_OK;
}

int osb_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(osb::g_prof_mu);
    if (on) osb::g_prof_acc.clear();
    osb::g_prof.store(on != 0);
    return OSB_OK;
}

int osb_profile_report(char* buf, size_t capacity) {
    OSB_REQUIRE(buf && capacity > 2, "report buffer too small");
    OSB_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(osb::g_prof_mu);
    osb::prof_drain();
    std::string out = "{";
    bool first = true;
    for (auto& kv : osb::g_prof_acc) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 kv.second.second);
        out += tmp;
        first = false;
    }
    out += "}";
    OSB_REQUIRE(out.size() + 1 <= capacity, "report buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return OSB_OK;
}

}  // extern "C"
