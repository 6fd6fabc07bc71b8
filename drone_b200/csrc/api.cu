// api.cu -- host side of libb200drone.so: the C ABI declared in include/b200drone.h.
//
// Thin by design: owns device memory, fills the kernel argument structs and
// launches on the caller's stream.  There is NO CPU implementation behind this
// file -- without a CUDA device every entry point that would compute returns
// B2D_ECUDA.
#include "../../include/b200drone.h"
#include "race_kernels.cuh"
#include "swarm_kernels.cuh"
#include "advantage_kernels.cuh"
#include "policy_kernels.cuh"
#include "rollout_kernels.cuh"

#include <cstdarg>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace b2d;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) return fail(B2D_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

enum { KIND_RACE = 0, KIND_SWARM = 1 };

// dst = clamp(src, -1, 1), a NaN stays a NaN (DR/dronelib.h:73-79,437): host_copy.c (SIMD, memcpy speed)
extern "C" void b2d_clamp_copy(float *dst, const float *src, size_t n);
static void clamp_copy(float *dst, const float *src, size_t count) { b2d_clamp_copy(dst, src, count); }

// Host-side copy of the action batch into the caller-visible (pinned) action buffer, optionally clamping.
// The step hands over all its chunks at once; chunks are claimed in order from an atomic counter and published
// with a release store; the issuing thread waits for chunk j (claiming chunks itself meanwhile) before it
// enqueues chunk j's transfers.  Helper threads are OFF by default: measured on the bench box (16 vCPUs, one
// B200, 1 M envs) the step takes 2.62 / 2.77 / 3.10 / 3.48 ms with 1 / 2 / 4 / 6 copying threads -- the copy
// competes with the step's own DMA traffic for host memory bandwidth, so one core copying at SIMD speed is the
// optimum there; B2D_COPY_THREADS=n enables n - 1 helpers for hosts where it is not.
class CopyPool {
public:
    static constexpr int MAX_CHUNKS = 64;
    explicit CopyPool(int workers) {
        for (int w = 0; w < workers; w++) threads_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    // chunk j covers floats [bounds[j], bounds[j + 1]) of dst / src
    void begin(float *dst, const float *src, const size_t *bounds, int nchunks, bool clamp) {
        {
            std::unique_lock<std::mutex> g(m_);
            idle_cv_.wait(g, [&] { return busy_ == 0; }); // stragglers of the previous step have left their claim loops
            dst_ = dst; src_ = src; clamp_ = clamp; nchunks_ = nchunks;
            for (int j = 0; j <= nchunks; j++) bounds_[j] = bounds[j];
            for (int j = 0; j < nchunks; j++) done_[j].store(0, std::memory_order_relaxed);
            next_.store(0, std::memory_order_relaxed);
            gen_++;
        }
        if (!threads_.empty()) cv_.notify_all();
    }
    void wait(int j) { // chunk j is in place when this returns
        while (done_[j].load(std::memory_order_acquire) == 0) {
            if (!claim_one()) std::this_thread::yield();
        }
    }

private:
    bool claim_one() {
        const int c = next_.fetch_add(1, std::memory_order_relaxed);
        if (c >= nchunks_) return false;
        float *d = dst_ + bounds_[c];
        const float *s = src_ + bounds_[c];
        const size_t n = bounds_[c + 1] - bounds_[c];
        if (clamp_) clamp_copy(d, s, n);
        else if (d != s) memcpy(d, s, n * sizeof(float));
        done_[c].store(1, std::memory_order_release);
        return true;
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                busy_++;
            }
            while (claim_one()) {
            }
            {
                std::lock_guard<std::mutex> g(m_);
                busy_--;
            }
            idle_cv_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, idle_cv_;
    unsigned long long gen_ = 0;
    int busy_ = 0;
    bool stop_ = false;
    float *dst_ = nullptr;
    const float *src_ = nullptr;
    bool clamp_ = false;
    int nchunks_ = 0;
    size_t bounds_[MAX_CHUNKS + 1];
    std::atomic<int> done_[MAX_CHUNKS];
    std::atomic<int> next_{0};
};

// Every entry point that touches the device runs with the handle's device current and restores the
// caller's device on the way out (a handle on cuda:1 may be stepped while cuda:0 is current).
struct DeviceScope {
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
        if (cur != dev) {
            ok = cudaSetDevice(dev) == cudaSuccess;
            prev = cur;
        }
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define DEVICE_SCOPE(v)                                                                      \
    DeviceScope dev_scope_((v)->device);                                                     \
    if (!dev_scope_.ok) return fail(B2D_ECUDA, "cudaSetDevice(%d) failed", (v)->device)

struct b2d_vec {
    int kind;
    int device;
    int num_envs, num_agents /* rows */, obs_dim, blob_floats, payload_floats;
    int math, write_clamped;
    bool host_clamp; // host buffers: leave clamp(action, -1, 1) in the caller's action array like DR/dronelib.h:437
    CopyPool *pool;  // host buffers: threads that share the action copy
    int step_ctas;   // race: CTAs of an overlapped (tape) launch = the largest grid; swarm: unused
    int single_ctas; // race: CTAs of a launch that runs alone
    RaceDev race;
    SwarmDev swarm;
    // device contract buffers (owned unless external)
    b2d_buffers dev;
    bool own_obs, own_act, own_rew, own_term, own_trunc;
    // host mirrors for the *_host entry points
    b2d_buffers host;
    bool has_host;
    bool registered[5];
    std::vector<void *> allocs; // state arrays etc.
    float *d_payload;
    float *d_blob_tmp;
    int *d_ids_tmp;
    size_t blob_tmp_cap;
    long long *d_log_out;
    long long h_log_out[16];
    double h_flog_out[8];
    long long launches;
    int swarm_grid[2]; // swarm: persistent step grid per math mode (resident CTA slots of the handle's device)
    uint32_t seq; // step-kernel launches so far (RaceDev::seq)
    uint32_t last_full_seq;  // race: seq of the last launch that covered all tiles with the full grid (0: none)
    cudaStream_t last_stream; // ... and the stream it went to
    cudaStream_t copy_streams[2];
    cudaEvent_t ev_step, ev_copy[2];
    // optional per-kernel timing (b2d_profile_kernels): event triples of the profiled steps
    bool profile;
    std::vector<cudaEvent_t> prof_events;
};

template <class T> static int dev_alloc(b2d_vec *v, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return fail(B2D_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    e = cudaMemset(q, 0, count * sizeof(T));
    if (e != cudaSuccess) return fail(B2D_ECUDA, "cudaMemset: %s", cudaGetErrorString(e));
    v->allocs.push_back(q);
    *p = (T *)q;
    return B2D_OK;
}

// page-locked already (cudaHostAlloc / cudaHostRegister by the caller)?  Plain host memory reports cudaMemoryTypeUnregistered.
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

static int check_ext(const b2d_buffers *ext) {
    if (!ext) return B2D_OK;
    if (ext->location != B2D_MEM_DEVICE && ext->location != B2D_MEM_HOST)
        return fail(B2D_EINVAL, "buffers.location must be B2D_MEM_DEVICE or B2D_MEM_HOST");
    if (ext->location == B2D_MEM_DEVICE) {
        if (((uintptr_t)ext->observations & 15u) || ((uintptr_t)ext->actions & 15u))
            return fail(B2D_EINVAL, "observations and actions must be 16-byte aligned");
        if (((uintptr_t)ext->rewards & 3u)) return fail(B2D_EINVAL, "rewards must be 4-byte aligned");
    }
    return B2D_OK;
}

// allocate / adopt the contract buffers; rows = num_agents
static int setup_buffers(b2d_vec *v, const b2d_buffers *ext) {
    const size_t rows = (size_t)v->num_agents;
    const bool ext_dev = ext && ext->location == B2D_MEM_DEVICE;
    memset(&v->dev, 0, sizeof(v->dev));
    v->dev.location = B2D_MEM_DEVICE;
    int rc;
    if (ext_dev && ext->observations) v->dev.observations = ext->observations;
    else if ((rc = dev_alloc(v, &v->dev.observations, rows * v->obs_dim))) return rc;
    if (ext_dev && ext->actions) v->dev.actions = ext->actions;
    else if ((rc = dev_alloc(v, &v->dev.actions, rows * 4))) return rc;
    if (ext_dev && ext->rewards) v->dev.rewards = ext->rewards;
    else if ((rc = dev_alloc(v, &v->dev.rewards, rows))) return rc;
    if (ext_dev && ext->terminals) v->dev.terminals = ext->terminals;
    else if ((rc = dev_alloc(v, &v->dev.terminals, rows))) return rc;
    if (ext_dev && ext->truncations) v->dev.truncations = ext->truncations;
    else if ((rc = dev_alloc(v, &v->dev.truncations, rows))) return rc;
    v->has_host = ext && ext->location == B2D_MEM_HOST;
    if (v->has_host) {
        v->host = *ext;
        int workers = 0; // see CopyPool: helpers lose on the measured host
        if (const char *env = getenv("B2D_COPY_THREADS")) workers = atoi(env) > 0 ? atoi(env) - 1 : 0;
        v->pool = new (std::nothrow) CopyPool(workers);
        if (!v->pool) return fail(B2D_ENOMEM, "out of host memory");
        // page-lock the caller's buffers for the life of the handle (they must outlive it anyway:
        // ownership as in the reference).  Already-pinned memory reports an error that is ignored.
        void *ptrs[5] = {ext->observations, ext->actions, ext->rewards, ext->terminals, ext->truncations};
        const size_t bytes[5] = {rows * v->obs_dim * 4, rows * 16, rows * 4, rows, rows};
        for (int k = 0; k < 5; k++) {
            v->registered[k] = false;
            if (!ptrs[k]) continue;
            if (host_is_pinned(ptrs[k])) continue; // the caller's own pinned allocation (the Python wrappers do that)
            if (cudaHostRegister(ptrs[k], bytes[k], cudaHostRegisterDefault) == cudaSuccess) v->registered[k] = true;
            else cudaGetLastError();
        }
    }
    return B2D_OK;
}

static int finish_create(b2d_vec *v) {
    int rc;
    if ((rc = dev_alloc(v, &v->d_log_out, 16))) return rc;
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&v->copy_streams[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&v->ev_copy[k], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&v->ev_step, cudaEventDisableTiming));
    CUDA_TRY(cudaDeviceSynchronize());
    return B2D_OK;
}

extern "C" int b2d_race_create(b2d_vec **out, const b2d_race_cfg *cfg, const b2d_buffers *ext) {
    if (!out || !cfg) return fail(B2D_EINVAL, "b2d_race_create: null argument");
    *out = nullptr;
    if (cfg->num_envs <= 0) return fail(B2D_EINVAL, "num_envs must be greater than 0");
    if (cfg->max_rings <= 0 || cfg->max_rings > 4096) return fail(B2D_EINVAL, "max_rings must be in [1, 4096]");
    if (cfg->max_moves <= 0) return fail(B2D_EINVAL, "max_moves must be positive");
    if (cfg->math != B2D_MATH_FAST && cfg->math != B2D_MATH_STRICT) return fail(B2D_EINVAL, "unknown math mode");
    int rc = check_ext(ext);
    if (rc) return rc;
    DeviceScope scope(cfg->device);
    if (!scope.ok) return fail(B2D_ECUDA, "cudaSetDevice(%d) failed", cfg->device);
    b2d_vec *v = new (std::nothrow) b2d_vec();
    if (!v) return fail(B2D_ENOMEM, "out of host memory");
    v->kind = KIND_RACE;
    v->device = cfg->device;
    v->num_envs = v->num_agents = cfg->num_envs;
    v->obs_dim = B2D_RACE_OBS;
    v->blob_floats = B2D_RACE_BLOB + 6 * cfg->max_rings;
    v->payload_floats = v->blob_floats;
    v->math = cfg->math;
    v->write_clamped = cfg->write_clamped_actions == 1;
    v->host_clamp = cfg->write_clamped_actions < 0;
    RaceDev &d = v->race;
    memset(&d, 0, sizeof(d));
    d.n = cfg->num_envs;
    d.ld = (cfg->num_envs + RACE_LD_ALIGN - 1) / RACE_LD_ALIGN * RACE_LD_ALIGN;
    d.max_rings = cfg->max_rings;
    d.max_moves = cfg->max_moves;
    d.key0 = (uint32_t)cfg->seed;
    d.key1 = (uint32_t)(cfg->seed >> 32);
    d.env_id_base = cfg->env_id_base;
    d.reset_mode = B2D_RESET_PHILOX;
    {   // persistent grid: one resident set of step CTAs per SM (or fewer for small N)
        int sms = 0, per_sm = 0, per_sm_strict = 0;
        cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(race_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RACE_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(race_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RACE_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(race_step_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(race_step_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, race_step_kernel<false>, RACE_BLOCK, RACE_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_strict, race_step_kernel<true>, RACE_BLOCK, RACE_SMEM_BYTES);
        if (e != cudaSuccess || sms < 1 || per_sm < 1 || per_sm_strict < 1) {
            delete v;
            return fail(B2D_ECUDA, "race_step_kernel setup on device %d: %s", cfg->device,
                        e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on an SM");
        }
        if (per_sm > RACE_MIN_CTAS) per_sm = RACE_MIN_CTAS; // what the kernel is tuned for (__launch_bounds__)
        const int warps_per_cta = RACE_BLOCK / 32;
        const int ntiles = (cfg->num_envs + 31) / 32;
        int step_ctas = (ntiles + warps_per_cta - 1) / warps_per_cta;
        // overlapped launches (b2d_vec_step_tape) run best with every CTA slot taken (finer per-CTA chains:
        // 68 vs 74 us per 1M envs), a launch that runs alone with one CTA per SM fewer (78 vs 84 us)
        const int alone = sms * (per_sm > 1 ? per_sm - 1 : 1);
        v->single_ctas = step_ctas < alone ? step_ctas : alone;
        if (step_ctas > sms * per_sm) step_ctas = sms * per_sm;
        v->step_ctas = step_ctas;
        d.max_grid = step_ctas;
    }
    const size_t ld = d.ld;
    if ((rc = setup_buffers(v, ext)) || (rc = dev_alloc(v, &d.S, RACE_HOT_SLOTS * ld)) ||
        (rc = dev_alloc(v, &d.chain, (size_t)v->step_ctas)) || (rc = dev_alloc(v, &d.cta_score, (size_t)v->step_ctas)) ||
        (rc = dev_alloc(v, &d.ctl, 1)) || (rc = finish_create(v))) {
        b2d_vec_close(v);
        return rc;
    }
    race_ctl_reset_kernel<<<1, 256>>>(d.ctl, d.cta_score, 0u, (unsigned int)v->step_ctas, 1);
    {
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            b2d_vec_close(v);
            return fail(B2D_ECUDA, "race_ctl_reset_kernel: %s", cudaGetErrorString(e));
        }
    }
#if B2D_EXPERIMENT_TIMING
    cudaMalloc(&d.trace, (size_t)2 * v->step_ctas * 4 * sizeof(unsigned long long));
    cudaMemset(d.trace, 0, (size_t)2 * v->step_ctas * 4 * sizeof(unsigned long long));
#endif
    d.obs = v->dev.observations;
    d.act_in = v->dev.actions;
    d.act_out = v->write_clamped ? v->dev.actions : nullptr;
    d.rew = v->dev.rewards;
    d.term = v->dev.terminals;
    *out = v;
    return B2D_OK;
}

// closed-form formation targets of the swarm tasks, evaluated on the host with the reference's own
// expressions (DS/drone_swarm.h:246-261 orbit, :273-281 cube, :299-307 flag): PI is raylib's float
// literal, sqrt / cos / sin are the double libm calls.  table[3][A][3]
static void swarm_formation_table(int A, std::vector<float> &tab) {
    tab.assign((size_t)3 * A * 3, 0.0f);
    for (int idx = 0; idx < A; idx++) {
        {
            float Rr = 8.0f;
            // the reference is C: sqrt / cos / sin below are the DOUBLE libm functions (a C++ overload
            // on float arguments would silently pick sqrtf / cosf / sinf)
            float phi = (float)((double)3.14159265358979323846f * (sqrt((double)5.0f) - (double)1.0f));
            float y = 1.0f - 2 * ((float)idx / (float)A);
            float radius = sqrtf(1.0f - y * y);
            float theta = phi * idx;
            float x = (float)(cos((double)theta) * (double)radius);
            float z = (float)(sin((double)theta) * (double)radius);
            float *o = &tab[((size_t)0 * A + idx) * 3];
            o[0] = Rr * x; o[1] = Rr * z; o[2] = Rr * y;
        }
        {
            int i = idx;
            float z = i / 16;
            i = i % 16;
            float x = (float)(i % 4);
            float y = (float)(i / 4);
            float *o = &tab[((size_t)1 * A + idx) * 3];
            o[0] = 4 * x - 6; o[1] = 4 * y - 6; o[2] = 4 * z - 6;
        }
        {
            float x = (float)(idx % 8);
            float y = (float)(idx / 8);
            x = 2.0f * x - 7;
            y = 5 - 1.5f * y;
            float *o = &tab[((size_t)2 * A + idx) * 3];
            o[0] = 0.0f; o[1] = x; o[2] = y;
        }
    }
}

extern "C" int b2d_swarm_create(b2d_vec **out, const b2d_swarm_cfg *cfg, const b2d_buffers *ext) {
    if (!out || !cfg) return fail(B2D_EINVAL, "b2d_swarm_create: null argument");
    *out = nullptr;
    if (cfg->num_envs <= 0) return fail(B2D_EINVAL, "num_envs must be greater than 0");
    if (cfg->num_agents <= 0 || cfg->num_agents > SWARM_BLOCK) return fail(B2D_EINVAL, "num_agents must be in [1, 128]");
    if (cfg->max_rings <= 0 || cfg->max_rings > 4096) return fail(B2D_EINVAL, "max_rings must be in [1, 4096]");
    if (cfg->math != B2D_MATH_FAST && cfg->math != B2D_MATH_STRICT) return fail(B2D_EINVAL, "unknown math mode");
    if ((long long)cfg->num_envs * cfg->num_agents > 0x7fffffffLL / 64) return fail(B2D_EINVAL, "too many agents");
    int rc = check_ext(ext);
    if (rc) return rc;
    DeviceScope scope(cfg->device);
    if (!scope.ok) return fail(B2D_ECUDA, "cudaSetDevice(%d) failed", cfg->device);
    b2d_vec *v = new (std::nothrow) b2d_vec();
    if (!v) return fail(B2D_ENOMEM, "out of host memory");
    v->kind = KIND_SWARM;
    v->device = cfg->device;
    v->num_envs = cfg->num_envs;
    v->num_agents = cfg->num_envs * cfg->num_agents;
    v->obs_dim = B2D_SWARM_OBS;
    v->blob_floats = cfg->num_agents * B2D_SWARM_AGENT_BLOB + 2 + 6 * cfg->max_rings;
    v->payload_floats = cfg->num_agents * B2D_SWARM_AGENT_PAYLOAD + 2 + 6 * cfg->max_rings;
    v->math = cfg->math;
    v->write_clamped = cfg->write_clamped_actions == 1;
    v->host_clamp = cfg->write_clamped_actions < 0;
    v->step_ctas = 1;
    memset(&v->race, 0, sizeof(v->race));
    SwarmDev &d = v->swarm;
    memset(&d, 0, sizeof(d));
    d.n = cfg->num_envs;
    d.A = cfg->num_agents;
    d.R = d.max_rings = cfg->max_rings;
    d.rows = v->num_agents;
    d.ld = (d.rows + 255) / 256 * 256;
    d.epc = SWARM_BLOCK / d.A;
    d.key0 = (uint32_t)cfg->seed;
    d.key1 = (uint32_t)(cfg->seed >> 32);
    d.env_id_base = cfg->env_id_base;
    d.reset_mode = B2D_RESET_PHILOX;
    const size_t ld = d.ld;
    float *form = nullptr;
    if ((rc = setup_buffers(v, ext)) || (rc = dev_alloc(v, &d.S, 5 * ld)) || (rc = dev_alloc(v, &d.P, 3 * ld)) ||
        (rc = dev_alloc(v, &d.T, ld)) || (rc = dev_alloc(v, &d.U, ld)) || (rc = dev_alloc(v, &d.V, ld)) ||
        (rc = dev_alloc(v, &d.W, ld)) || (rc = dev_alloc(v, &d.RS, 4 * ld)) || (rc = dev_alloc(v, &d.E, (size_t)d.n)) ||
        (rc = dev_alloc(v, &d.G0, (size_t)d.R * d.n)) || (rc = dev_alloc(v, &d.G1, (size_t)d.R * d.n)) ||
        (rc = dev_alloc(v, &form, (size_t)9 * d.A)) || (rc = dev_alloc(v, &d.ctl, 1)) || (rc = finish_create(v))) {
        b2d_vec_close(v);
        return rc;
    }
    std::vector<float> tab;
    swarm_formation_table(d.A, tab);
    const unsigned int one = 1;
    if (cudaMemcpy(form, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(&d.ctl->grid, &one, sizeof(one), cudaMemcpyHostToDevice) != cudaSuccess) {
        b2d_vec_close(v);
        return fail(B2D_ECUDA, "swarm table upload failed");
    }
    d.form = form;
    if ((rc = swarm_step_setup(d, cfg->device, v->swarm_grid))) { // per handle: its device's SM count and kernel attributes
        b2d_vec_close(v);
        return fail(B2D_ECUDA, "swarm_kernel setup on device %d failed", cfg->device);
    }
    if ((rc = dev_alloc(v, &d.chain, (size_t)(v->swarm_grid[0] > v->swarm_grid[1] ? v->swarm_grid[0] : v->swarm_grid[1])))) {
        b2d_vec_close(v);
        return rc;
    }
    d.obs = v->dev.observations;
    d.act_in = v->dev.actions;
    d.act_out = v->write_clamped ? v->dev.actions : nullptr;
    d.rew = v->dev.rewards;
    d.term = v->dev.terminals;
    // every drone's first respawn is prepared from the start (vec_reset re-keys and regenerates)
    swarm_fill_slots_kernel<<<(d.rows + 127) / 128, 128>>>(d);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        b2d_vec_close(v);
        return fail(B2D_ECUDA, "swarm_fill_slots_kernel failed");
    }
    *out = v;
    return B2D_OK;
}

extern "C" int b2d_vec_close(b2d_vec *v) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    DeviceScope scope(v->device);
    cudaDeviceSynchronize();
#if B2D_EXPERIMENT_TIMING
    if (v->kind == KIND_RACE && v->race.ctl) {
        unsigned long long h[12];
        cudaMemcpy(h, v->race.ctl->dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const double it = (double)h[4], w = (double)h[7];
        fprintf(stderr, "[b2d timing] per tile iteration (cycles): wait_inputs %.0f  compute %.0f  store %.0f  (unused) %.0f  pass %.0f  tail_pass %.0f | per warp-launch: total %.0f  iterations %.2f\n",
                h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[8] / it, h[5] / it, h[6] / w, it / w);
        if (getenv("B2D_TRACE_FILE")) { // the last two launches, CTA by CTA: shows launch t+1 starting inside launch t
            std::vector<unsigned long long> tr((size_t)2 * v->step_ctas * 4);
            cudaMemcpy(tr.data(), v->race.trace, tr.size() * 8, cudaMemcpyDeviceToHost);
            FILE *f = fopen(getenv("B2D_TRACE_FILE"), "w");
            if (f) {
                fprintf(f, "launch_seq,cta,smid,entry_ns,go_ns,done_ns\n");
                for (int which = 0; which < 2; which++) {
                    const uint32_t seq = v->seq - (uint32_t)(1 - which); // older launch first
                    const size_t base = (size_t)(seq & 1u) * v->step_ctas * 4;
                    for (int c = 0; c < v->step_ctas; c++)
                        fprintf(f, "%u,%d,%llu,%llu,%llu,%llu\n", seq, c, tr[base + c * 4], tr[base + c * 4 + 1], tr[base + c * 4 + 2],
                                tr[base + c * 4 + 3]);
                }
                fclose(f);
            }
        }
        fprintf(stderr, "[b2d timing] slowest warp loop %.0f cycles; CTA busy: mean %.0f, slowest %.0f cycles (any launch)\n", (double)h[9],
                (double)h[10] / (w / RACE_WARPS), (double)h[11]);
    }
#endif
#if B2D_RO_TIMING
    if (v->kind == KIND_RACE && v->race.ctl) {
        unsigned long long h[12];
        cudaMemcpy(h, v->race.ctl->dbg, sizeof(h), cudaMemcpyDeviceToHost);
        if (h[8]) {
            const char *names[8] = {"obs->smem", "barrier1", "stores+noise", "wait GEMM1", "GELU", "barrier2", "GEMM2 trip", "sample+env"};
            double tot = 0;
            for (int m = 0; m < 8; m++) tot += (double)h[m];
            fprintf(stderr, "[b2d rollout timing] share of a warp's step loop:");
            for (int m = 0; m < 8; m++) fprintf(stderr, " %s %.1f%%", names[m], 100.0 * h[m] / tot);
            fprintf(stderr, " | cycles per warp launch %.0f\n", tot / (double)h[8]);
        }
    }
#endif
    if (v->has_host) {
        void *ptrs[5] = {v->host.observations, v->host.actions, v->host.rewards, v->host.terminals, v->host.truncations};
        for (int k = 0; k < 5; k++)
            if (v->registered[k] && ptrs[k]) cudaHostUnregister(ptrs[k]);
    }
    delete v->pool;
    v->pool = nullptr;

    for (void *p : v->allocs) cudaFree(p);
    if (v->d_payload) cudaFree(v->d_payload);
    if (v->d_blob_tmp) cudaFree(v->d_blob_tmp);
    if (v->d_ids_tmp) cudaFree(v->d_ids_tmp);
    for (int k = 0; k < 2; k++) {
        if (v->copy_streams[k]) cudaStreamDestroy(v->copy_streams[k]);
        if (v->ev_copy[k]) cudaEventDestroy(v->ev_copy[k]);
    }
    if (v->ev_step) cudaEventDestroy(v->ev_step);
    delete v;
    return B2D_OK;
}

static int launch_check(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2D_ECUDA, "%s launch: %s", what, cudaGetErrorString(e));
    return B2D_OK;
}

extern "C" int b2d_vec_reset(b2d_vec *v, uint64_t seed, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    if (v->kind == KIND_RACE) {
        RaceDev &d = v->race;
        d.key0 = (uint32_t)seed;
        d.key1 = (uint32_t)(seed >> 32);
        if (d.reset_mode == B2D_RESET_INJECT && !d.payload) return fail(B2D_ESTATE, "inject mode without a payload");
        race_ctl_reset_kernel<<<1, 256, 0, st>>>(d.ctl, d.cta_score, 0u, (unsigned int)v->step_ctas, 0);
        race_reset_kernel<<<(d.n + 127) / 128, 128, 0, st>>>(d);
        v->launches += 2;
        return launch_check("race_reset_kernel");
    }
    if (v->swarm.reset_mode == B2D_RESET_INJECT && !v->swarm.payload) return fail(B2D_ESTATE, "inject mode without a payload");
    CUDA_TRY(cudaMemsetAsync(&v->swarm.ctl->ctas_done, 0, sizeof(unsigned int), st));
    swarm_vec_reset(v->swarm, seed, st, &v->launches);
    return launch_check("swarm_reset_kernel");
}

// One launch of the step kernel.  Launch overlap: a launch that covers all tiles with the full grid, outside
// stream capture, and directly follows such a launch of the SAME handle on the SAME stream is made
// programmatically dependent on it (PDL): it may begin while its predecessor drains, and every CTA then waits
// for exactly its own predecessor's completion flag (race_step_kernel, chain_wait).  Whatever else sits between
// two steps in the stream (a policy kernel, a copy) is an ordinary full dependency, so the flags are already
// set when the next step arrives; `allow_overlap` = false forces a plain launch.
static int step_impl(b2d_vec *v, const float *actions, cudaStream_t st, bool allow_overlap = true, int tile_begin = 0,
                     int tile_end = -1, bool first_chunk = true, bool last_chunk = true) {
    if ((v->kind == KIND_RACE ? v->race.reset_mode : v->swarm.reset_mode) == B2D_RESET_INJECT && !v->d_payload)
        return fail(B2D_ESTATE, "inject mode without a payload (b2d_set_reset_payload)");
    if (v->kind == KIND_RACE) {
        RaceDev d = v->race;
        if (actions) d.act_in = actions;
        const bool full = tile_begin == 0 && tile_end < 0;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        const bool capturing = cap != cudaStreamCaptureStatusNone;
        // inside a capture the sequence numbers the flags carry would be frozen into the graph: plain launches there
        const bool full_grid = full && !capturing;
        const bool overlap = allow_overlap && full_grid && v->last_full_seq != 0 && v->last_full_seq == v->seq && v->last_stream == st;
        d.seq = ++v->seq;
        d.chain_wait = overlap ? 1 : 0;
        {   // tiles of this launch: the whole vector, or one chunk of it (host-buffer pipeline)
            const int ntiles = (d.n + 31) / 32;
            d.tile_begin = tile_begin;
            d.tile_end = tile_end < 0 || tile_end > ntiles ? ntiles : tile_end;
            d.count_step = last_chunk;
            d.score_add = !first_chunk;
        }
        cudaEvent_t pe[2] = {nullptr, nullptr};
        if (v->profile) {
            for (int k = 0; k < 2; k++) { cudaEventCreate(&pe[k]); v->prof_events.push_back(pe[k]); }
            cudaEventRecord(pe[0], st);
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        // every CTA slot for launches that can overlap their neighbours (finer per-CTA chains: 68 vs 74 us per 1M
        // envs), one CTA per SM fewer for launches that run alone (chunks of a host step, captured launches)
        cfg.gridDim = dim3((unsigned)(full_grid ? v->step_ctas : v->single_ctas));
        cfg.blockDim = dim3(RACE_BLOCK);
        cfg.dynamicSmemBytes = RACE_SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = overlap ? 1 : 0;
        cudaError_t e = v->math == B2D_MATH_STRICT ? cudaLaunchKernelEx(&cfg, race_step_kernel<true>, d)
                                                   : cudaLaunchKernelEx(&cfg, race_step_kernel<false>, d);
        if (e != cudaSuccess) return fail(B2D_ECUDA, "race_step_kernel launch: %s", cudaGetErrorString(e));
        if (v->profile) cudaEventRecord(pe[1], st);
        v->launches += 1;
        v->last_full_seq = full_grid ? d.seq : 0;
        v->last_stream = st;
        return launch_check("race_step_kernel");
    }
    {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        const bool capturing = cap != cudaStreamCaptureStatusNone; // sequence numbers would be frozen into the graph
        const bool full = tile_begin == 0 && tile_end < 0;
        const bool overlap = B2D_SW_OVERLAP && allow_overlap && full && !capturing && v->last_full_seq != 0 && v->last_full_seq == v->seq &&
                             v->last_stream == st;
        const unsigned int seq = ++v->seq;
        cudaError_t e = swarm_vec_step(v->swarm, actions, v->math, v->swarm_grid[v->math == B2D_MATH_STRICT ? 1 : 0], st, &v->launches,
                                       seq, overlap, tile_begin, tile_end);
        if (e != cudaSuccess) return fail(B2D_ECUDA, "swarm_kernel launch: %s", cudaGetErrorString(e));
        v->last_full_seq = (capturing || !full) ? 0 : seq;
        v->last_stream = st;
    }
    return launch_check("swarm_step_kernel");
}

extern "C" int b2d_vec_step(b2d_vec *v, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    DEVICE_SCOPE(v);
    return step_impl(v, nullptr, (cudaStream_t)stream);
}

extern "C" int b2d_vec_step_from(b2d_vec *v, const float *device_actions, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    if (!device_actions || ((uintptr_t)device_actions & 15u)) return fail(B2D_EINVAL, "actions must be a 16-byte aligned device pointer");
    DEVICE_SCOPE(v);
    return step_impl(v, device_actions, (cudaStream_t)stream);
}

// K steps over an action tape that is already on the device.  The launches after the first are
// allowed to overlap the tail of their predecessor (programmatic dependent launch + per-CTA
// completion flags): CTA c of step t+1 needs only CTA c of step t, so a straggling CTA no longer
// idles the whole GPU between steps.  Inside a stream capture the launches are plain (the
// sequence numbers the flags carry are host state a graph replay would not advance).
extern "C" int b2d_vec_step_tape(b2d_vec *v, const float *device_tape, int tape_len, int first, int steps, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    if (!device_tape || ((uintptr_t)device_tape & 15u) || tape_len <= 0 || first < 0 || steps < 0)
        return fail(B2D_EINVAL, "b2d_vec_step_tape: bad argument");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = (size_t)v->num_agents * 4;
    for (int k = 0; k < steps; k++) {
        const float *a = device_tape + (size_t)((first + k) % tape_len) * stride;
        int rc = step_impl(v, a, st);
        if (rc) return rc;
    }
    return B2D_OK;
}

// Host-buffer step.  The PCIe transfers dominate (16 B in, 121 B out per env against a kernel of
// 80 us per million envs), so the step is issued as a pipeline over chunks of the env range:
// actions of chunk j+1 go up and its tiles are stepped while the results of chunk j come down on
// the copy streams.  `host_actions` (optional) are copied into the caller-visible action buffer
// like the reference's wrapper does (`self.actions[:] = actions`), chunk by chunk, so that the
// 1 ms CPU copy of 16 MB hides behind the transfers as well.
static int step_host_impl(b2d_vec *v, const float *host_actions, cudaStream_t st) {
    const size_t rows = (size_t)v->num_agents;
    const bool copy_in = host_actions && host_actions != v->host.actions;
    int chunks = 1;
    if (rows >= (1u << 17)) chunks = rows >= (1u << 19) ? 8 : 4;
    // chunk boundaries in tiles (race: 32 envs; swarm: the envs one CTA steps together); the first chunk is split 1:3 so
    // that the first results start coming down after 1/32 of the vector instead of 1/8 (the pipeline's lead-in is pure latency)
    const size_t tile_rows = v->kind == KIND_RACE ? 32 : (size_t)v->swarm.epc * v->swarm.A;
    const size_t tiles = (rows + tile_rows - 1) / tile_rows, per_tiles = (tiles + chunks - 1) / chunks;
    std::vector<size_t> bnd;
    bnd.push_back(0);
    if (chunks > 1 && per_tiles >= 4) bnd.push_back(per_tiles / 4);
    for (int j = 1; j < chunks && (size_t)j * per_tiles < tiles; j++) bnd.push_back((size_t)j * per_tiles);
    bnd.push_back(tiles);
    const int nchunks = (int)bnd.size() - 1;
    static const bool late_small = !(getenv("B2D_HOST_SMALL_COPIES") && atoi(getenv("B2D_HOST_SMALL_COPIES")) == 1);
    // Actions straight from the caller's array when it is PAGE-LOCKED memory (cudaHostAlloc / cudaHostRegister by the
    // caller: `DroneRace.pinned_actions()` hands out such an array).  The reference's wrapper copies the caller's actions
    // into the env's action buffer and the step leaves them clamped there (`self.actions[:] = actions`, DR/dronelib.h:437).
    // That 16 B/env CPU copy otherwise sits in front of every chunk's upload -- one core at memcpy speed needs 1.5-2.5 ms
    // for 1 M envs, which is the whole PCIe budget on a slow or shared host.  From pinned memory the chunks go up by DMA
    // from where they are (the kernel clamps what it reads) and the clamped copy into the env's buffer is made by this
    // thread AFTER the last chunk is issued, while the results drain -- same buffer contents at return, nothing in front of
    // the transfers.  The library never page-locks caller memory itself: a registration would outlive a freed array whose
    // address the allocator hands out again, and DMA would then read the old pages.  (Checked per call: one driver query.)
    const float *h2d_src = v->host.actions;
    bool deferred_copy = false;
    static const bool inplace_ok = !(getenv("B2D_HOST_INPLACE_ACTIONS") && atoi(getenv("B2D_HOST_INPLACE_ACTIONS")) == 0);
    if (copy_in && inplace_ok && !v->write_clamped && host_is_pinned(host_actions) &&
        host_is_pinned(reinterpret_cast<const char *>(host_actions) + rows * 4 * sizeof(float) - 1)) {
        h2d_src = host_actions;
        deferred_copy = true;
    }
    const bool host_copy = (v->host_clamp || copy_in) && !deferred_copy;
    if (host_copy) { // all chunks of the action copy go to the pool at once; chunk j is awaited right before its H2D
        size_t fb[CopyPool::MAX_CHUNKS + 1];
        for (int j = 0; j <= nchunks; j++) fb[j] = (bnd[j] * tile_rows < rows ? bnd[j] * tile_rows : rows) * 4;
        v->pool->begin(v->host.actions, copy_in ? host_actions : v->host.actions, fb, nchunks, v->host_clamp);
    }
    for (int j = 0; j < nchunks; j++) {
        const size_t r0 = bnd[j] * tile_rows;
        const size_t r1 = bnd[j + 1] * tile_rows < rows ? bnd[j + 1] * tile_rows : rows;
        const size_t nr = r1 - r0;
        cudaStream_t cs = v->copy_streams[j & 1];
        // the copy of later chunks overlaps the transfers of the chunks already in flight
        // (with host_clamp the caller-visible buffer receives the clamped values, as the reference leaves them)
        if (host_copy) v->pool->wait(j);
        CUDA_TRY(cudaMemcpyAsync(v->dev.actions + r0 * 4, h2d_src + r0 * 4, nr * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
        int rc = step_impl(v, nullptr, st, false, (int)bnd[j], (int)bnd[j + 1], j == 0, j == nchunks - 1);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(v->ev_step, st));
        CUDA_TRY(cudaStreamWaitEvent(cs, v->ev_step, 0));
        CUDA_TRY(cudaMemcpyAsync(v->host.observations + r0 * v->obs_dim, v->dev.observations + r0 * v->obs_dim,
                                 nr * v->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, cs));
        // rewards and terminals (5 B per env) come down as two copies of the whole vector after the last chunk's
        // kernel instead of two small copies per chunk: every DMA operation costs ~10 us of link time whatever its size
        if (!late_small) {
            CUDA_TRY(cudaMemcpyAsync(v->host.rewards + r0, v->dev.rewards + r0, nr * sizeof(float), cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaMemcpyAsync(v->host.terminals + r0, v->dev.terminals + r0, nr, cudaMemcpyDeviceToHost, cs));
        } else if (j == nchunks - 1) {
            cudaStream_t os = v->copy_streams[(j + 1) & 1];
            CUDA_TRY(cudaStreamWaitEvent(os, v->ev_step, 0));
            CUDA_TRY(cudaMemcpyAsync(v->host.rewards, v->dev.rewards, rows * sizeof(float), cudaMemcpyDeviceToHost, os));
            CUDA_TRY(cudaMemcpyAsync(v->host.terminals, v->dev.terminals, rows, cudaMemcpyDeviceToHost, os));
        }
        if (v->write_clamped)
            CUDA_TRY(cudaMemcpyAsync(v->host.actions + r0 * 4, v->dev.actions + r0 * 4, nr * 4 * sizeof(float), cudaMemcpyDeviceToHost, cs));
    }
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaEventRecord(v->ev_copy[k], v->copy_streams[k]));
        CUDA_TRY(cudaStreamWaitEvent(st, v->ev_copy[k], 0));
    }
    if (deferred_copy) { // the env's own action buffer, as the reference leaves it; overlaps the results coming down
        if (v->host_clamp) clamp_copy(v->host.actions, host_actions, rows * 4);
        else memcpy(v->host.actions, host_actions, rows * 4 * sizeof(float));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2D_OK;
}

extern "C" int b2d_vec_step_host(b2d_vec *v, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    if (!v->has_host) return fail(B2D_ESTATE, "handle was not created with host buffers");
    DEVICE_SCOPE(v);
    return step_host_impl(v, nullptr, (cudaStream_t)stream);
}

extern "C" int b2d_vec_step_host_from(b2d_vec *v, const float *host_actions, void *stream) {
    if (!v || !host_actions) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    if (!v->has_host) return fail(B2D_ESTATE, "handle was not created with host buffers");
    DEVICE_SCOPE(v);
    return step_host_impl(v, host_actions, (cudaStream_t)stream);
}

extern "C" int b2d_vec_reset_host(b2d_vec *v, uint64_t seed, void *stream) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    if (!v->has_host) return fail(B2D_ESTATE, "handle was not created with host buffers");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = b2d_vec_reset(v, seed, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(v->host.observations, v->dev.observations, (size_t)v->num_agents * v->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2D_OK;
}

// ---------------------------------------------------------------- vec_log
extern "C" int b2d_vec_log_begin(b2d_vec *v, void *stream, long long **device_sums, int *count) {
    if (!v) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    Ctl *ctl = v->kind == KIND_RACE ? v->race.ctl : v->swarm.ctl;
    if (v->kind == KIND_RACE) race_log_snapshot_kernel<<<1, 32, 0, st>>>(ctl, v->race.cta_score, v->d_log_out);
    else swarm_log_snapshot_kernel<<<1, 32, 0, st>>>(ctl, v->d_log_out);
    v->launches += 1;
    if (device_sums) *device_sums = v->d_log_out;
    if (count) *count = 16;
    return launch_check("log_snapshot_kernel");
}

extern "C" int b2d_vec_log_end(b2d_vec *v, float out[B2D_LOG_FIELDS], void *stream) {
    if (!v || !out) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(v->h_log_out, v->d_log_out, 16 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return b2d_log_average(v->kind, v->kind == KIND_RACE ? v->race.max_rings : v->swarm.max_rings, v->h_log_out, 16, out);
}

// ncclAllReduce of the process's own NCCL, resolved lazily (no link-time dependency; nccl.h: ncclInt64 = 4, ncclSum = 0)
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
static nccl_allreduce_fn resolve_nccl_allreduce() {
    static nccl_allreduce_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllReduce");
        }
        fn = (nccl_allreduce_fn)sym;
    }
    return fn;
}

extern "C" int b2d_vec_log_reduce(b2d_vec *v, float out[B2D_LOG_FIELDS], void *nccl_comm, void *stream) {
    if (!v || !out) return fail(B2D_EINVAL, "Missing or invalid vec env handle");
    long long *sums = nullptr;
    int count = 0;
    int rc = b2d_vec_log_begin(v, stream, &sums, &count);
    if (rc) return rc;
    if (nccl_comm) {
        nccl_allreduce_fn all_reduce = resolve_nccl_allreduce();
        if (!all_reduce) return fail(B2D_ESTATE, "b2d_vec_log_reduce: no NCCL in this process (ncclAllReduce not found)");
        DEVICE_SCOPE(v);
        const int nrc = all_reduce(sums, sums, (size_t)count, 4 /* ncclInt64 */, 0 /* ncclSum */, nccl_comm, (cudaStream_t)stream);
        if (nrc != 0) return fail(B2D_ECUDA, "ncclAllReduce failed with ncclResult_t %d", nrc);
    }
    return b2d_vec_log_end(v, out, stream);
}

extern "C" int b2d_log_average(int kind, int max_rings, const long long *a, int count, float out[B2D_LOG_FIELDS]) {
    if (!a || !out || count < 16 || max_rings <= 0) return fail(B2D_EINVAL, "b2d_log_average: bad argument");
    for (int k = 0; k < B2D_LOG_FIELDS; k++) out[k] = 0.0f;
    const double n = (double)a[ACC_N];
    if (a[ACC_N] == 0) return B2D_OK;
    if (kind == KIND_RACE) {
        // Log field order: DR/dronelib.h:52-63; averaging: EB:588-591
        out[0] = (float)((double)a[ACC_RETURN] / n);
        out[1] = (float)((double)a[ACC_LENGTH] / n);
        out[2] = 0.0f;
        out[3] = (float)((double)a[ACC_COLLISION] / n);
        out[4] = (float)((double)a[ACC_OOB] / n);
        out[5] = (float)((double)a[ACC_TIMEOUT] / n);
        out[6] = (float)((double)a[ACC_COUNT] / n); // score: only episodes that ended in the last step
        out[7] = (float)((double)a[ACC_RINGS] / (double)max_rings / n);
        out[8] = (float)n;
    } else {
        swarm_log_finish(a, out);
    }
    return B2D_OK;
}

extern "C" int b2d_vec_log(b2d_vec *v, float out[B2D_LOG_FIELDS], void *stream) {
    int rc = b2d_vec_log_begin(v, stream, nullptr, nullptr);
    if (rc) return rc;
    return b2d_vec_log_end(v, out, stream);
}

// ---------------------------------------------------------------- introspection
extern "C" int b2d_get_buffers(const b2d_vec *v, b2d_buffers *out) {
    if (!v || !out) return fail(B2D_EINVAL, "null argument");
    *out = v->dev;
    return B2D_OK;
}
extern "C" int b2d_num_agents(const b2d_vec *v) { return v ? v->num_agents : fail(B2D_EINVAL, "null handle"); }
extern "C" int b2d_obs_dim(const b2d_vec *v) { return v ? v->obs_dim : fail(B2D_EINVAL, "null handle"); }
extern "C" int b2d_state_blob_floats(const b2d_vec *v) { return v ? v->blob_floats : fail(B2D_EINVAL, "null handle"); }
extern "C" long long b2d_kernel_launches(const b2d_vec *v) { return v ? v->launches : -1; }

extern "C" int b2d_step_count(b2d_vec *v, uint32_t *steps, void *stream) {
    if (!v || !steps) return fail(B2D_EINVAL, "null argument");
    DEVICE_SCOPE(v);
    Ctl *ctl = v->kind == KIND_RACE ? v->race.ctl : v->swarm.ctl;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    uint32_t w[2] = {0, 1};
    CUDA_TRY(cudaMemcpy(w, &ctl->ctas_done, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    *steps = v->kind == KIND_RACE ? w[0] : (w[1] ? w[0] / w[1] : 0);
    return B2D_OK;
}

extern "C" int b2d_guard_replays(b2d_vec *v, unsigned long long *count, void *stream) {
    if (!v || !count) return fail(B2D_EINVAL, "null argument");
    DEVICE_SCOPE(v);
    Ctl *ctl = v->kind == KIND_RACE ? v->race.ctl : v->swarm.ctl;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaMemcpy(count, &ctl->guard_replays, sizeof(*count), cudaMemcpyDeviceToHost));
    return B2D_OK;
}

extern "C" int b2d_set_step_count(b2d_vec *v, uint32_t steps) {
    if (!v) return fail(B2D_EINVAL, "null handle");
    DEVICE_SCOPE(v);
    Ctl *ctl = v->kind == KIND_RACE ? v->race.ctl : v->swarm.ctl;
    CUDA_TRY(cudaDeviceSynchronize());
    const uint32_t done = steps;
    CUDA_TRY(cudaMemcpy(&ctl->ctas_done, &done, sizeof(uint32_t), cudaMemcpyHostToDevice));
    return B2D_OK;
}

// ---------------------------------------------------------------- state hooks
// Rings that come from outside (put_state, the reset payload of the parity hook) need storage:
// [R][ld] x 24 B, allocated the first time such a hook is used (lazily generated rings need none).
static int ensure_external_rings(b2d_vec *v) {
    if (v->kind != KIND_RACE || v->race.X0) return B2D_OK;
    RaceDev &d = v->race;
    int rc;
    if ((rc = dev_alloc(v, &d.X0, (size_t)d.max_rings * d.ld)) || (rc = dev_alloc(v, &d.X1, (size_t)d.max_rings * d.ld))) {
        d.X0 = nullptr;
        d.X1 = nullptr;
        return rc;
    }
    return B2D_OK;
}

static int ensure_tmp(b2d_vec *v, int n) {
    const size_t need = (size_t)n * v->blob_floats;
    if (need > v->blob_tmp_cap) {
        if (v->d_blob_tmp) cudaFree(v->d_blob_tmp);
        if (v->d_ids_tmp) cudaFree(v->d_ids_tmp);
        v->d_blob_tmp = nullptr;
        v->d_ids_tmp = nullptr;
        v->blob_tmp_cap = 0;
        if (cudaMalloc(&v->d_blob_tmp, need * sizeof(float)) != cudaSuccess) return fail(B2D_ENOMEM, "cudaMalloc blob staging");
        if (cudaMalloc(&v->d_ids_tmp, (size_t)n * sizeof(int)) != cudaSuccess) return fail(B2D_ENOMEM, "cudaMalloc id staging");
        v->blob_tmp_cap = need;
    }
    return B2D_OK;
}

extern "C" int b2d_get_state(b2d_vec *v, const int *env_ids, int n, float *host_blobs) {
    if (!v || !host_blobs || n <= 0 || n > v->num_envs) return fail(B2D_EINVAL, "b2d_get_state: bad argument");
    DEVICE_SCOPE(v);
    int rc = ensure_tmp(v, n);
    if (rc) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if (env_ids) {
        for (int k = 0; k < n; k++)
            if (env_ids[k] < 0 || env_ids[k] >= v->num_envs) return fail(B2D_EINVAL, "env id out of range");
        CUDA_TRY(cudaMemcpy(v->d_ids_tmp, env_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (v->kind == KIND_RACE) race_pack_kernel<<<(n + 127) / 128, 128>>>(v->race, env_ids ? v->d_ids_tmp : nullptr, n, v->d_blob_tmp);
    else swarm_pack_kernel<<<(n * v->swarm.A + 127) / 128, 128>>>(v->swarm, env_ids ? v->d_ids_tmp : nullptr, n, v->d_blob_tmp);
    v->launches += 1;
    if ((rc = launch_check("pack_kernel"))) return rc;
    CUDA_TRY(cudaMemcpy(host_blobs, v->d_blob_tmp, (size_t)n * v->blob_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return B2D_OK;
}

extern "C" int b2d_put_state(b2d_vec *v, const int *env_ids, int n, const float *host_blobs) {
    if (!v || !host_blobs || n <= 0 || n > v->num_envs) return fail(B2D_EINVAL, "b2d_put_state: bad argument");
    DEVICE_SCOPE(v);
    int rc = ensure_tmp(v, n);
    if (rc) return rc;
    if ((rc = ensure_external_rings(v))) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if (env_ids) {
        for (int k = 0; k < n; k++)
            if (env_ids[k] < 0 || env_ids[k] >= v->num_envs) return fail(B2D_EINVAL, "env id out of range");
        CUDA_TRY(cudaMemcpy(v->d_ids_tmp, env_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMemcpy(v->d_blob_tmp, host_blobs, (size_t)n * v->blob_floats * sizeof(float), cudaMemcpyHostToDevice));
    if (v->kind == KIND_RACE) race_unpack_kernel<<<(n + 127) / 128, 128>>>(v->race, env_ids ? v->d_ids_tmp : nullptr, n, v->d_blob_tmp);
    else swarm_unpack_kernel<<<(n * v->swarm.A + 127) / 128, 128>>>(v->swarm, env_ids ? v->d_ids_tmp : nullptr, n, v->d_blob_tmp);
    v->launches += 1;
    if ((rc = launch_check("unpack_kernel"))) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    return B2D_OK;
}

// ---------------------------------------------------------------- render / checkpoint bridge
static_assert(sizeof(b2d_ref_drone) == 208, "b2d_ref_drone must match the reference's Drone (dronelib.h:191-247)");
static_assert(sizeof(b2d_ref_ring) == 44, "b2d_ref_ring must match the reference's Ring (dronelib.h:161-166)");

static void ref_ring_from(const float *g, b2d_ref_ring *r) {
    r->pos[0] = g[0]; r->pos[1] = g[1]; r->pos[2] = g[2];
    r->normal[0] = g[3]; r->normal[1] = g[4]; r->normal[2] = g[5];
    const bool none = g[3] == 0.0f && g[4] == 0.0f && g[5] == 0.0f; // a zeroed ring of a non-race swarm task
    r->radius = none ? 0.0f : 2.0f;
    // shortest rotation taking +z to the normal: q = (1 + n.z, z x n) normalised
    float w = 1.0f + g[5], x = -g[4], y = g[3], z = 0.0f;
    const float n2 = w * w + x * x + y * y;
    if (none || n2 < 1e-12f) { w = none ? 1.0f : 0.0f; x = none ? 0.0f : 1.0f; y = 0.0f; }
    else { const float inv = 1.0f / sqrtf(n2); w *= inv; x *= inv; y *= inv; }
    r->orientation[0] = w; r->orientation[1] = x; r->orientation[2] = y; r->orientation[3] = z;
}

// b[0:17] state, b[17:30] the 13 randomised params (blob order)
static void ref_drone_core(const float *b, b2d_ref_drone *d) {
    memset(d, 0, sizeof(*d));
    for (int k = 0; k < 3; k++) { d->pos[k] = b[k]; d->vel[k] = b[3 + k]; d->omega[k] = b[10 + k]; d->prev_pos[k] = b[k]; }
    for (int k = 0; k < 4; k++) { d->quat[k] = b[6 + k]; d->rpms[k] = b[13 + k]; }
    d->mass = b[17]; d->ixx = b[18]; d->iyy = b[19]; d->izz = b[20]; d->arm_len = b[21]; d->k_thrust = b[22];
    d->k_ang_damp = b[23]; d->k_drag = b[24]; d->b_drag = b[25]; d->gravity = b[26]; d->max_rpm = b[27];
    d->max_vel = 50.0f; d->max_omega = 50.0f; // BASE_MAX_VEL / BASE_MAX_OMEGA, DR/dronelib.h:286-287
    d->k_mot = b[28]; d->j_mot = b[29];
}

extern "C" int b2d_race_blob_to_ref(const float *blob, int max_rings, b2d_ref_drone *drone, b2d_ref_ring *rings, int *tick,
                                    int *ring_idx, float *episodic_return) {
    if (!blob || !drone || !rings || max_rings <= 0) return fail(B2D_EINVAL, "b2d_race_blob_to_ref: bad argument");
    ref_drone_core(blob, drone);
    drone->ring_idx = 0; // the race env keeps ring_idx / score on DroneRace, not on the Drone
    if (tick) *tick = (int)blob[30];
    if (ring_idx) *ring_idx = (int)blob[31];
    if (episodic_return) *episodic_return = blob[32];
    for (int r = 0; r < max_rings; r++) ref_ring_from(blob + B2D_RACE_BLOB + 6 * r, &rings[r]);
    return B2D_OK;
}

extern "C" int b2d_swarm_blob_to_ref(const float *blob, int num_agents, int max_rings, b2d_ref_drone *drones, b2d_ref_ring *rings,
                                     int *tick, int *task) {
    if (!blob || !drones || !rings || num_agents <= 0 || max_rings <= 0) return fail(B2D_EINVAL, "b2d_swarm_blob_to_ref: bad argument");
    for (int a = 0; a < num_agents; a++) {
        const float *b = blob + (size_t)a * B2D_SWARM_AGENT_BLOB;
        b2d_ref_drone *d = &drones[a];
        ref_drone_core(b, d);
        for (int k = 0; k < 3; k++) { d->spawn_pos[k] = b[30 + k]; d->target_pos[k] = b[33 + k]; d->target_vel[k] = b[36 + k]; }
        d->last_abs_reward = b[39]; d->last_target_reward = b[40]; d->last_collision_reward = b[41];
        d->episode_return = b[42]; d->collisions = b[43]; d->episode_length = (int)b[44]; d->score = b[45]; d->ring_idx = (int)b[46];
    }
    const float *eb = blob + (size_t)num_agents * B2D_SWARM_AGENT_BLOB;
    if (tick) *tick = (int)eb[0];
    if (task) *task = (int)eb[1];
    for (int r = 0; r < max_rings; r++) ref_ring_from(eb + 2 + 6 * r, &rings[r]);
    return B2D_OK;
}

extern "C" int b2d_export_ref(b2d_vec *v, int env_id, b2d_ref_drone *drones, b2d_ref_ring *rings, int *tick, int *aux,
                              float *episodic_return) {
    if (!v || !drones || !rings) return fail(B2D_EINVAL, "b2d_export_ref: null argument");
    if (env_id < 0 || env_id >= v->num_envs) return fail(B2D_EINVAL, "env id out of range");
    std::vector<float> blob((size_t)v->blob_floats);
    int rc = b2d_get_state(v, &env_id, 1, blob.data());
    if (rc) return rc;
    if (v->kind == KIND_RACE) return b2d_race_blob_to_ref(blob.data(), v->race.max_rings, drones, rings, tick, aux, episodic_return);
    if (episodic_return) *episodic_return = 0.0f;
    return b2d_swarm_blob_to_ref(blob.data(), v->swarm.A, v->swarm.max_rings, drones, rings, tick, aux);
}

extern "C" int b2d_observe(b2d_vec *v, void *stream) {
    if (!v) return fail(B2D_EINVAL, "null handle");
    DEVICE_SCOPE(v);
    cudaStream_t st = (cudaStream_t)stream;
    if (v->kind == KIND_RACE) race_observe_kernel<<<(v->race.n + 127) / 128, 128, 0, st>>>(v->race);
    else swarm_observe_launch(v->swarm, st);
    v->launches += 1;
    return launch_check("observe_kernel");
}

// ---------------------------------------------------------------- configuration hooks
extern "C" int b2d_set_math(b2d_vec *v, int math) {
    if (!v || (math != B2D_MATH_FAST && math != B2D_MATH_STRICT)) return fail(B2D_EINVAL, "unknown math mode");
    v->math = math;
    return B2D_OK;
}

extern "C" int b2d_set_reset_mode(b2d_vec *v, int mode) {
    if (!v || (mode != B2D_RESET_PHILOX && mode != B2D_RESET_INJECT)) return fail(B2D_EINVAL, "unknown reset mode");
    DEVICE_SCOPE(v);
    if (mode == B2D_RESET_INJECT) {
        CUDA_TRY(cudaDeviceSynchronize());
        int rc = ensure_external_rings(v);
        if (rc) return rc;
    }
    if (v->kind == KIND_RACE) v->race.reset_mode = mode;
    else v->swarm.reset_mode = mode;
    return B2D_OK;
}

extern "C" int b2d_set_reset_payload(b2d_vec *v, const float *host_payload) {
    if (!v || !host_payload) return fail(B2D_EINVAL, "null argument");
    DEVICE_SCOPE(v);
    const size_t bytes = (size_t)v->num_envs * v->payload_floats * sizeof(float);
    if (!v->d_payload) {
        if (cudaMalloc(&v->d_payload, bytes) != cudaSuccess) return fail(B2D_ENOMEM, "cudaMalloc payload");
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(v->d_payload, host_payload, bytes, cudaMemcpyHostToDevice));
    if (v->kind == KIND_RACE) v->race.payload = v->d_payload;
    else v->swarm.payload = v->d_payload;
    return B2D_OK;
}

// Per-kernel timing of the steps issued while profiling is on.  enable=1 starts recording
// CUDA events around the step kernel of every subsequent step; enable=0 synchronises, returns the
// mean duration in microseconds (out_us[0] = step kernel, out_us[1] = 0: the separate adoption
// kernel of early builds is gone, the slot is kept for ABI stability, out_us[2] = profiled steps)
// and stops.  Not capturable.
extern "C" int b2d_profile_kernels(b2d_vec *v, int enable, float out_us[3]) {
    if (!v) return fail(B2D_EINVAL, "null handle");
    if (enable) {
        v->profile = true;
        return B2D_OK;
    }
    DEVICE_SCOPE(v);
    v->profile = false;
    CUDA_TRY(cudaDeviceSynchronize());
    double a = 0, b = 0;
    const size_t n = v->prof_events.size() / 2;
    for (size_t k = 0; k < n; k++) {
        float t0 = 0;
        cudaEventElapsedTime(&t0, v->prof_events[2 * k], v->prof_events[2 * k + 1]);
        a += t0;
    }
    for (cudaEvent_t e : v->prof_events) cudaEventDestroy(e);
    v->prof_events.clear();
    if (out_us) {
        out_us[0] = n ? (float)(a / n * 1e3) : 0.0f;
        out_us[1] = n ? (float)(b / n * 1e3) : 0.0f;
        out_us[2] = (float)n;
    }
    return B2D_OK;
}

// ---------------------------------------------------------------- advantage (trainer side)
extern "C" int b2d_puff_advantage(const float *values, const float *rewards, const float *dones, const float *importance,
                                  float *advantages, float *abs_sum, int num_rows, int horizon, long long row_stride,
                                  long long t_stride, float gamma, float lambda, float rho_clip, float c_clip, int math,
                                  void *stream) {
    if (!values || !rewards || !dones || !importance || !advantages) return fail(B2D_EINVAL, "b2d_puff_advantage: null tensor");
    if (num_rows < 0 || horizon < 0 || row_stride <= 0 || t_stride <= 0) return fail(B2D_EINVAL, "b2d_puff_advantage: bad shape");
    if (math != B2D_MATH_FAST && math != B2D_MATH_STRICT) return fail(B2D_EINVAL, "unknown math mode");
    if (num_rows == 0) return B2D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (num_rows + 255) / 256;
    if (math == B2D_MATH_STRICT)
        puff_advantage_kernel<true><<<grid, 256, 0, st>>>(values, rewards, dones, importance, advantages, abs_sum, num_rows, horizon,
                                                          row_stride, t_stride, gamma, lambda, rho_clip, c_clip);
    else
        puff_advantage_kernel<false><<<grid, 256, 0, st>>>(values, rewards, dones, importance, advantages, abs_sum, num_rows, horizon,
                                                           row_stride, t_stride, gamma, lambda, rho_clip, c_clip);
    return launch_check("puff_advantage_kernel");
}

// ---------------------------------------------------------------- fused policy step (rollout side)
template <class K> static int policy_launch(K kernel, int slot, int threads, int warps, size_t smem, const PolicyArgs &a, cudaStream_t st) {
    static int grid_for[4][64] = {{0}};
    static size_t smem_for[4][64] = {{0}};
    static std::mutex cache_mu; // per-device launch geometry, shared by every caller thread of the process
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(B2D_EINVAL, "device ordinal out of range");
    std::unique_lock<std::mutex> lk(cache_mu);
    if (grid_for[slot][dev] == 0 || smem_for[slot][dev] != smem) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) return fail(B2D_ECUDA, "policy kernel does not fit on an SM (%zu bytes of shared memory)", smem);
        grid_for[slot][dev] = per_sm * sms;
        smem_for[slot][dev] = smem;
    }
    const int slots = grid_for[slot][dev];
    lk.unlock();
    const int tiles = (a.rows + POL_TILE - 1) / POL_TILE;
    int grid = (tiles + warps - 1) / warps;
    if (grid > slots) grid = slots;
    kernel<<<grid, threads, smem, st>>>(a);
    return launch_check("policy_act_kernel");
}

static size_t policy_smem_fp32(int D, int H) { return ((size_t)D * H + 7 * (size_t)H + (size_t)POL_WARPS * POL_TILE * D) * sizeof(float); }
static size_t policy_smem_tf32(int D, int H, int warps) {
    const size_t nt = H / 8, ks = (D + 7) / 8;
    return nt * ks * 256 + nt * 256 + (size_t)H * 4 + (size_t)warps * POL_TILE * (D + 8) * 4;
}

extern "C" int b2d_policy_act(const b2d_policy_weights *w, const b2d_policy_io *io, uint64_t noise_seed,
                              unsigned int *device_counter, int deterministic, void *stream) {
    if (!w || !io) return fail(B2D_EINVAL, "b2d_policy_act: null argument");
    if (!w->encoder_weight || !w->encoder_bias || !w->decoder_mean_weight || !w->decoder_mean_bias || !w->decoder_logstd ||
        !w->value_weight || !w->value_bias)
        return fail(B2D_EINVAL, "b2d_policy_act: null weight tensor");
    if (w->hidden < 8 || w->hidden > 256 || (w->hidden & 7)) return fail(B2D_EINVAL, "b2d_policy_act: hidden must be a multiple of 8 in [8, 256]");
    if (w->precision != B2D_POLICY_FP32 && w->precision != B2D_POLICY_TF32) return fail(B2D_EINVAL, "b2d_policy_act: unknown precision");
    if (!io->observations || !io->rewards || !io->terminals || !io->env_actions) return fail(B2D_EINVAL, "b2d_policy_act: null env buffer");
    if (!device_counter) return fail(B2D_EINVAL, "b2d_policy_act: null call counter");
    if (io->rows < 0) return fail(B2D_EINVAL, "b2d_policy_act: negative row count");
    if (io->obs_dim != B2D_RACE_OBS && io->obs_dim != B2D_SWARM_OBS) return fail(B2D_EINVAL, "b2d_policy_act: obs_dim must be 29 or 41");
    if (((uintptr_t)io->env_actions & 15) || ((uintptr_t)io->store_actions & 15))
        return fail(B2D_EINVAL, "b2d_policy_act: action buffers must be 16-byte aligned");
    if (io->rows == 0) return B2D_OK;
    PolicyArgs a;
    a.enc_w = w->encoder_weight;
    a.enc_b = w->encoder_bias;
    a.mean_w = w->decoder_mean_weight;
    a.mean_b = w->decoder_mean_bias;
    a.logstd = w->decoder_logstd;
    a.value_w = w->value_weight;
    a.value_b = w->value_bias;
    a.hidden = w->hidden;
    a.obs = io->observations;
    a.rew = io->rewards;
    a.term = io->terminals;
    a.env_act = io->env_actions;
    a.st_obs = io->store_observations;
    a.st_act = io->store_actions;
    a.st_logp = io->store_logprobs;
    a.st_rew = io->store_rewards;
    a.st_term = io->store_terminals;
    a.st_val = io->store_values;
    a.rows = io->rows;
    a.row_id_base = io->row_id_base;
    a.seed_lo = (uint32_t)noise_seed;
    a.seed_hi = (uint32_t)(noise_seed >> 32);
    a.counter = device_counter;
    a.deterministic = deterministic ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int H = a.hidden;
    if (w->precision == B2D_POLICY_TF32) {
        if (io->obs_dim == B2D_RACE_OBS)
            return policy_launch(policy_act_tf32_kernel<B2D_RACE_OBS, 8, 4>, 2, 256, 8, policy_smem_tf32(B2D_RACE_OBS, H, 8), a, st);
        return policy_launch(policy_act_tf32_kernel<B2D_SWARM_OBS, 4, 2>, 3, 128, 4, policy_smem_tf32(B2D_SWARM_OBS, H, 4), a, st);
    }
    if (io->obs_dim == B2D_RACE_OBS)
        return policy_launch(policy_act_kernel<B2D_RACE_OBS>, 0, POL_THREADS, POL_WARPS, policy_smem_fp32(B2D_RACE_OBS, H), a, st);
    return policy_launch(policy_act_kernel<B2D_SWARM_OBS>, 1, POL_THREADS, POL_WARPS, policy_smem_fp32(B2D_SWARM_OBS, H), a, st);
}

// ---------------------------------------------------------------- the rollout as one kernel
extern "C" int b2d_race_rollout(b2d_vec *v, const b2d_policy_weights *w, const b2d_rollout_store *store, int horizon,
                                uint64_t noise_seed, unsigned int *device_counter, int deterministic, void *stream) {
    if (!v || !w || !store) return fail(B2D_EINVAL, "b2d_race_rollout: null argument");
    if (v->kind != KIND_RACE) return fail(B2D_EINVAL, "b2d_race_rollout: race handles only");
    if (!w->encoder_weight || !w->encoder_bias || !w->decoder_mean_weight || !w->decoder_mean_bias || !w->decoder_logstd ||
        !w->value_weight || !w->value_bias)
        return fail(B2D_EINVAL, "b2d_race_rollout: null weight tensor");
    if (w->hidden != RO_HIDDEN) return fail(B2D_EINVAL, "b2d_race_rollout: hidden must be 128");
    if (horizon <= 0) return fail(B2D_EINVAL, "b2d_race_rollout: horizon must be positive");
    if (!device_counter) return fail(B2D_EINVAL, "b2d_race_rollout: null call counter");
    if (((uintptr_t)store->actions & 15)) return fail(B2D_EINVAL, "b2d_race_rollout: the action store must be 16-byte aligned");
    if (v->race.reset_mode != B2D_RESET_PHILOX) return fail(B2D_ESTATE, "b2d_race_rollout: not available in inject mode");
    DEVICE_SCOPE(v);
    static int grid_for[2][64] = {{0}};
    static std::mutex cache_mu;
    const int dev = v->device, m = v->math == B2D_MATH_STRICT ? 1 : 0;
    if (dev < 0 || dev >= 64) return fail(B2D_EINVAL, "device ordinal out of range");
    std::unique_lock<std::mutex> lk(cache_mu);
    if (grid_for[m][dev] == 0) {
        int per_sm = 0, sms = 0;
        if (m) {
            CUDA_TRY(cudaFuncSetAttribute(race_rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RO_SMEM_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(race_rollout_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, race_rollout_kernel<true>, RO_THREADS, RO_SMEM_BYTES));
        } else {
            CUDA_TRY(cudaFuncSetAttribute(race_rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RO_SMEM_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(race_rollout_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, race_rollout_kernel<false>, RO_THREADS, RO_SMEM_BYTES));
        }
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) return fail(B2D_ECUDA, "race_rollout_kernel does not fit on an SM");
        // The occupancy calculator answers 1 for a kernel that allocates tensor memory (it cannot know how many of
        // the SM's 512 TMEM columns a CTA will ask for); the kernel takes 160, registers (168 x 128) and shared
        // memory (72 KB) are sized for RO_CTAS_PER_SM = 3, and three CTAs per SM is what runs (measured: 346 / 206 /
        // 161 us per step with 1 / 2 / 3 CTAs per SM; a fourth waits in tcgen05.alloc).
        per_sm = RO_CTAS_PER_SM;
        if (getenv("B2D_RO_CTAS")) per_sm = atoi(getenv("B2D_RO_CTAS")); // measurement aid
        grid_for[m][dev] = per_sm * sms;
    }
    const int ro_slots = grid_for[m][dev];
    lk.unlock();
    // the episode bank: allocated on first use (not inside a stream capture), topped up before every launch
    cudaStream_t st = (cudaStream_t)stream;
    if (!v->race.bank) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap != cudaStreamCaptureStatusNone)
            return fail(B2D_ESTATE, "b2d_race_rollout: call once outside stream capture first (allocates the episode bank)");
        int rc = dev_alloc(v, &v->race.bank, (size_t)v->race.n * RACE_BANK_SLOTS * 6);
        if (rc) return rc;
    }
    race_bank_fill_kernel<<<(v->race.n * RACE_BANK_SLOTS + 127) / 128, 128, 0, st>>>(v->race);
    v->launches += 1;
    RolloutArgs a;
    a.d = v->race;
    a.enc_w = w->encoder_weight; a.enc_b = w->encoder_bias; a.mean_w = w->decoder_mean_weight; a.mean_b = w->decoder_mean_bias;
    a.logstd = w->decoder_logstd; a.value_w = w->value_weight; a.value_b = w->value_bias;
    a.st_obs = store->observations; a.st_act = store->actions; a.st_logp = store->logprobs; a.st_rew = store->rewards;
    a.st_term = store->terminals; a.st_val = store->values;
    a.env_act = v->dev.actions;
    a.horizon = horizon;
    a.row_id_base = v->race.env_id_base;
    a.seed_lo = (uint32_t)noise_seed;
    a.seed_hi = (uint32_t)(noise_seed >> 32);
    a.counter = device_counter;
    a.deterministic = deterministic ? 1 : 0;
    const int chunks = (v->race.n + RO_THREADS - 1) / RO_THREADS;
    const int grid = chunks < ro_slots ? chunks : ro_slots;
    if (m) race_rollout_kernel<true><<<grid, RO_THREADS, RO_SMEM_BYTES, st>>>(a);
    else race_rollout_kernel<false><<<grid, RO_THREADS, RO_SMEM_BYTES, st>>>(a);
    v->launches += 1;
    return launch_check("race_rollout_kernel");
}

extern "C" const char *b2d_last_error(void) { return g_err; }
extern "C" int b2d_version(void) { return B2D_VERSION; }
