// altb_api.cu -- C ABI (include/altair_b200.h) over the sm_100a kernels.  Host-side plumbing only:
// scene validation, first-event setup, device buffers, launches, multi-device fan-out inside
// one process.  No CPU fallback: every entry point needs a CUDA device.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "altb_kernels.cuh"

using namespace altb;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ALTB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static constexpr double PI_D = 3.14159265358979323846;
static constexpr uint64_t DEFAULT_BATCH = 1ull << 28;   // 8 GiB of records per launch: one tail per 2.7e8 rays (+1.2 % vs 2^26)

struct DevCtx {
    int dev = -1;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    altb_record* rec = nullptr; uint64_t rec_cap = 0;
    unsigned int* counter = nullptr;
    unsigned long long* counts = nullptr; uint64_t counts_cap = 0;
    unsigned long long* stats = nullptr;      // [8]
    float* tables = nullptr; uint64_t tables_cap = 0;   // floats
    float4* tiles = nullptr; uint64_t tiles_cap = 0;
    float4* lines = nullptr; uint64_t lines_cap = 0;    // escaping rays' test lines for the line maps (2 float4 each): rectangle list from the front, tile list from the back
    float4* raw = nullptr; uint64_t raw_cap = 0;        // LINES sink of k_trace: escaping rays' end points and directions (2 float4 each)
    float2* sincos = nullptr;                           // SC_N-entry azimuth table (altb_math.cuh: SinCosTab)
    // second record buffer + counter + two worker streams: consecutive launches of a DIRECTION-mode job alternate between
    // them, so that the tail of one persistent launch (its last long rays) overlaps the start of the next one
    altb_record* rec2 = nullptr; uint64_t rec2_cap = 0;
    unsigned int* counter2 = nullptr;
    QEntry* rq[2] = {nullptr, nullptr};                 // k_trace's resume queues (one set per launch in flight)
    unsigned long long* gstat[2] = {nullptr, nullptr};  // k_trace's per-block statistics (direction sink), same
    unsigned long long* stats_scratch = nullptr; uint64_t stats_scratch_cap = 0;
    // LINE-map tables of the last map spec (setup_map)
    bool map_cached = false; altb_map_spec map_key; MapParams map_M; int map_n_tiles = 0; size_t map_line_smem = 0, map_rect_smem = 0;
    std::vector<float> tab_host; std::vector<float4> tiles_host;
    double* dirtab = nullptr; uint64_t dirtab_cap = 0; int dir_nt = 0, dir_np = 0; std::vector<double> dirtab_host;    // direction_bin edges
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// NCCL through dlopen: a context with several devices merges the per-device maps with ONE all-reduce over NVLink (SURVEY 8b:
// "ctx owns streams / NCCL comms / device buffers").  Loaded on demand so that single-GPU hosts need no NCCL and a torch
// process keeps the copy it already loaded (same soname); only the six entry points below are used.
struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};
static const NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;      // 0 not tried, 1 ok, -1 unavailable
    if (state == 0) {
        state = -1;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.CommInitAll && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

struct altb_ctx {
    std::vector<DevCtx> devs;
    std::vector<ncclComm_t> comms;    // one per device when the context owns several (empty: host merge)
    uint64_t batch = DEFAULT_BATCH;
    int contract = ALTB_CONTRACT_EXACT;
    bool batch_user = false;          // altb_set_batch was called: the direction sink honours it too (default there: 2^31)
    uint64_t launches = 0;
    uint64_t trace_launches = 0;      // k_trace launches only (roofline: average launch duration)
};

// ---------------------------------------------------------------------------------- scene setup
static int make_geom(const altb_scene* sc, Geom& g, KConsts& k) {
    if (!(sc->r_inner > 0) || !(sc->r_outer >= sc->r_inner) || !(sc->world_half > sc->r_outer))
        return fail(ALTB_E_SCENE, "scene: need 0 < r_inner <= r_outer < world_half");
    if (!(sc->theta_max_deg > 90.0) || !(sc->theta_max_deg < 180.0))
        return fail(ALTB_E_SCENE, "scene: theta_max_deg must be in (90,180)");
    if (sc->max_bounces < 1) return fail(ALTB_E_SCENE, "scene: max_bounces < 1");
    if (!(sc->reflectance >= 0.0)) return fail(ALTB_E_SCENE, "scene: reflectance must be >= 0");
    const double th = sc->theta_max_deg * PI_D / 180.0;
    g.R1 = sc->r_inner; g.R2 = sc->r_outer;
    g.R1sq = g.R1 * g.R1; g.R2sq = g.R2 * g.R2;
    g.cth = cos(th); g.sth = sin(th);
    g.zc = g.R1 * g.cth;
    const double t = g.sth / g.cth;
    g.T2 = t * t;
    g.H = sc->world_half; g.exit_z = sc->exit_z;
    g.lambertian = sc->lambertian; g.brdf_kind = sc->brdf_kind;
    g.max_bounces = sc->max_bounces; g.count_all = sc->count_all_status;
    double ps = 0.0, bs = 0.0;
    if (sc->brdf_kind == 1 || sc->brdf_kind == 3) {
        const double sum = sc->brdf_param[1] + sc->brdf_param[2];
        if (!(sum > 0)) return fail(ALTB_E_SCENE, "scene: brdf specular+diffuse must be > 0");
        ps = sc->brdf_param[1] / sum;
        bs = sc->brdf_param[0] * PI_D / 6.0;
    } else if (sc->brdf_kind == 2) {
        const double ne = sc->brdf_param[0];
        if (!(ne >= 1.0 && ne <= 8.0) || ne != floor(ne)) return fail(ALTB_E_SCENE, "scene: cos^n lobe exponent must be an integer in 1..8");
        if (!(sc->brdf_param[1] > 0.0 && sc->brdf_param[1] <= 90.0)) return fail(ALTB_E_SCENE, "scene: cos^n lobe max angle must be in (0,90] deg");
    } else if (sc->brdf_kind != 0) return fail(ALTB_E_SCENE, "scene: brdf_kind %d not supported", sc->brdf_kind);
    k.lobe_n = sc->brdf_kind == 2 ? (int)sc->brdf_param[0] : 0;
    k.lobe_ang = (float)(sc->brdf_kind == 2 ? sc->brdf_param[1] * PI_D / 180.0 : 0.0);
    k.rho = (float)sc->reflectance; k.sigma = (float)sc->roughness_rad;
    k.two_r1 = (float)(2.0 * g.R1); k.neg_inv_r1 = (float)(-1.0 / g.R1); k.nr_c = (float)(-0.5 / g.R1sq);
    k.zc = (float)g.zc; k.p_spec = (float)ps; k.brdf_s = (float)bs; k.exit_zf = (float)g.exit_z; k.inv_r2 = (float)(1.0 / g.R2);
    // integer forms of the two comparisons (u_abs = k 2^-24, u_sel = k 2^-14 are exact in f32):
    //   rho < u_abs   <=>  k > floor(rho 2^24)  <=>  w0 > (floor(rho 2^24) << 8 | 0xff)      (w0 = k << 8 | low byte)
    //   u_sel < p     <=>  k < ceil(p 2^14)
    const double fa = floor((double)k.rho * 16777216.0);
    k.abs_thr = fa >= 16777216.0 ? 0xffffffffu : (((uint32_t)fa << 8) | 0xffu);
    k.spec_thr = (uint32_t)ceil((double)k.p_spec * 16384.0);
    // Box-Muller radius of a 20-bit u1: |g| <= sqrt(2 * 20 ln 2) = 5.2655
    k.tilt_small = fabs((double)k.sigma) * 5.2656 <= (double)SINCOS_TINY_MAX ? 2 : (fabs((double)k.sigma) * 5.2656 <= (double)SINCOS_DIRECT_MAX ? 1 : 0);
    k.spec_small = fabs((double)k.brdf_s) * 5.2656 <= (double)SINCOS_DIRECT_MAX;
    return 0;
}

// log table of DrawTabs::log_u20 (altb_math.cuh): double precision, rounded once
static std::vector<float4> make_log_table() {
    std::vector<float4> t(LG_N);
    for (int hi = 0; hi < LG_N; hi++) {
        const double mh = 1.0 + hi / 128.0;
        const bool half = hi >= 53;                                 // m_hi > ~sqrt(2): measure from the next power of two
        t[hi] = make_float4((float)log(half ? mh * 0.5 : mh), (float)(1.0 / mh), half ? -146.0f : -147.0f, 0.0f);
    }
    return t;
}

// ---------------------------------------------------------------------------------- context
extern "C" const char* altb_last_error(void) { return g_err.c_str(); }
extern "C" int altb_version(void) { return ALTB_VERSION; }
extern "C" int altb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void altb_destroy(altb_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->comms.empty())
        if (const NcclApi* N = nccl_api())
            for (size_t i = 0; i < ctx->comms.size(); i++)
                if (ctx->comms[i]) { cudaSetDevice(ctx->devs[i].dev); cudaStreamSynchronize(ctx->devs[i].stream); N->CommDestroy(ctx->comms[i]); }
    for (auto& d : ctx->devs) {
        if (d.dev < 0) continue;
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        cudaFree(d.rec); cudaFree(d.counter); cudaFree(d.counts); cudaFree(d.stats); cudaFree(d.tables); cudaFree(d.tiles); cudaFree(d.lines); cudaFree(d.raw); cudaFree(d.sincos);
        cudaFree(d.rec2); cudaFree(d.counter2); cudaFree(d.rq[0]); cudaFree(d.rq[1]); cudaFree(d.gstat[0]); cudaFree(d.gstat[1]); cudaFree(d.stats_scratch); cudaFree(d.dirtab);
        for (auto& a : d.aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
        if (d.fork_ev) cudaEventDestroy(d.fork_ev);
        for (auto& e : d.join_ev) if (e) cudaEventDestroy(e);
        for (auto& e : d.ev) if (e) cudaEventDestroy(e);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

// k_trace's statistics buffer: [SMs][MAX_SLOTS][STAT_WORDS] block words, then [SMs][TRACE_THREADS][2] per-thread words (TraceParams::lane_acc)
static inline size_t gstat_words(int sms) { return (size_t)sms * MAX_SLOTS * STAT_WORDS + (size_t)sms * TRACE_THREADS * 2; }

extern "C" int altb_create(altb_ctx** out, const int* devices, int n_devices) {
    if (!out) return fail(ALTB_E_ARG, "altb_create: out is NULL");
    *out = nullptr;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail < 1) {
        cudaGetLastError();
        return fail(ALTB_E_CUDA, "altb_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (n_devices <= 0) n_devices = devices ? 0 : avail;
    if (n_devices < 1 || n_devices > avail) return fail(ALTB_E_ARG, "altb_create: n_devices=%d, visible=%d", n_devices, avail);
    altb_ctx* ctx = new altb_ctx();
    ctx->devs.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        DevCtx& d = ctx->devs[i];
        const int dev = devices ? devices[i] : i;
        if (dev < 0 || dev >= avail) { altb_destroy(ctx); return fail(ALTB_E_ARG, "altb_create: bad device %d", dev); }
        cudaDeviceProp prop;
        if (cudaSetDevice(dev) != cudaSuccess) { altb_destroy(ctx); return fail(ALTB_E_CUDA, "altb_create: cudaSetDevice(%d) failed", dev); }
        d.dev = dev;        // from here on altb_destroy frees whatever this device already owns
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&d.counter, 4 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMalloc(&d.counter2, 4 * sizeof(unsigned int)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.aux[0], cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.aux[1], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.fork_ev, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.join_ev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d.join_ev[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaMalloc(&d.stats, 8 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&d.rq[0], (size_t)prop.multiProcessorCount * TRACE_WARPS * RQCAP * sizeof(QEntry)) != cudaSuccess ||
            cudaMalloc(&d.rq[1], (size_t)prop.multiProcessorCount * TRACE_WARPS * RQCAP * sizeof(QEntry)) != cudaSuccess ||
            cudaMalloc(&d.gstat[0], gstat_words(prop.multiProcessorCount) * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&d.gstat[1], gstat_words(prop.multiProcessorCount) * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&d.sincos, TABS_BYTES) != cudaSuccess) {
            const char* msg = cudaGetErrorString(cudaGetLastError());
            altb_destroy(ctx);
            return fail(ALTB_E_CUDA, "altb_create: device %d init failed: %s", dev, msg);
        }
        d.sm_count = prop.multiProcessorCount;
        k_make_sincos_table<<<SC_N / 256, 256, 0, d.stream>>>(d.sincos);
        ctx->launches++;
        const std::vector<float4> lg = make_log_table();
        if (cudaMemcpyAsync(d.sincos + SC_N, lg.data(), LG_N * sizeof(float4), cudaMemcpyHostToDevice, d.stream) != cudaSuccess ||
            cudaStreamSynchronize(d.stream) != cudaSuccess) {
            const char* msg = cudaGetErrorString(cudaGetLastError());
            altb_destroy(ctx);
            return fail(ALTB_E_CUDA, "altb_create: device %d: %s (is this an sm_100a GPU?)", dev, msg);
        }
        for (auto& ev : d.ev) cudaEventCreate(&ev);
    }
    if (n_devices > 1 && !getenv("ALTB_NO_NCCL")) {
        // the context owns the communicators of its devices; without NCCL on the host the maps are merged on the host instead
        if (const NcclApi* N = nccl_api()) {
            std::vector<int> ids(n_devices);
            for (int i = 0; i < n_devices; i++) ids[i] = ctx->devs[i].dev;
            ctx->comms.assign(n_devices, nullptr);
            const ncclResult_t r = N->CommInitAll(ctx->comms.data(), n_devices, ids.data());
            if (r != ncclSuccess) {
                const std::string msg = N->GetErrorString(r);
                ctx->comms.clear();
                altb_destroy(ctx);
                return fail(ALTB_E_CUDA, "altb_create: ncclCommInitAll over %d devices failed: %s", n_devices, msg.c_str());
            }
        }
    }
    *out = ctx;
    return 0;
}

// how a multi-device context merges its per-device maps: 0 = one device (nothing to merge), 1 = NCCL all-reduce, 2 = host sum
extern "C" int altb_collective(const altb_ctx* ctx) {
    if (!ctx || ctx->devs.size() < 2) return 0;
    return ctx->comms.empty() ? 2 : 1;
}

extern "C" int altb_set_batch(altb_ctx* ctx, uint64_t batch_rays) {
    if (!ctx) return fail(ALTB_E_ARG, "ctx is NULL");
    ctx->batch = batch_rays ? batch_rays : DEFAULT_BATCH;
    ctx->batch_user = batch_rays != 0;
    if (ctx->batch > (1ull << 31) - 1) ctx->batch = (1ull << 31) - 1;   // lane indices are 31 bits in the single-scene instances
    return 0;
}

extern "C" int altb_set_contract(altb_ctx* ctx, int contract) {
    if (!ctx) return fail(ALTB_E_ARG, "ctx is NULL");
    if (contract != ALTB_CONTRACT_EXACT && contract != ALTB_CONTRACT_FAST && contract != ALTB_CONTRACT_FAST7) return fail(ALTB_E_ARG, "altb_set_contract: unknown contract %d", contract);
    ctx->contract = contract;
    return 0;
}
extern "C" int altb_get_contract(const altb_ctx* ctx) { return ctx ? ctx->contract : -1; }

extern "C" uint64_t altb_launch_count(const altb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t altb_trace_launch_count(const altb_ctx* ctx) { return ctx ? ctx->trace_launches : 0; }

template <typename T>
static int ensure(T*& p, uint64_t& cap, uint64_t need) {
    if (need <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    CK(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return 0;
}

// ---------------------------------------------------------------------------------- trace launch
struct TraceSetup { TraceParams P; bool rough; int model; bool rescatter; };   // rescatter: brdf_kind 3 (k_rescatter after the Lambertian trace)

static int setup_trace(const altb_scene* sc, const altb_source* src, uint64_t seed, TraceSetup& ts) {
    memset(&ts.P, 0, sizeof ts.P);
    if (int rc = make_geom(sc, ts.P.g, ts.P.k)) return rc;
    const int kind0 = launch_ray(ts.P.g, src->pos, src->dir, ts.P.d0, ts.P.x0);
    if (kind0 < 0) return fail(ALTB_E_SOURCE, "source must lie strictly inside the inner sphere with a non-zero direction");
    ts.P.kind0 = kind0;
    for (int i = 0; i < 3; i++) { ts.P.x0f[i] = (float)ts.P.x0[i]; ts.P.d0f[i] = (float)ts.P.d0[i]; }
    ts.P.keys = philox_expand(seed);
    ts.rough = sc->roughness_rad != 0.0;
    ts.model = !sc->lambertian ? 2 : (sc->brdf_kind == 1 ? 1 : (sc->brdf_kind == 2 ? 3 : 0));
    ts.rescatter = sc->brdf_kind == 3;
    if (ts.rescatter && !sc->lambertian) return fail(ALTB_E_SCENE, "scene: brdf_kind 3 (post-hoc re-scatter) needs lambertian = 1");
    // one slot: this scene
    ts.P.n_slots = 1;
    ts.P.slots[0].zc = ts.P.g.zc; ts.P.slots[0].T2 = ts.P.g.T2; ts.P.slots[0].cth = ts.P.g.cth; ts.P.slots[0].sth = ts.P.g.sth;
    ts.P.slots[0].zcf = ts.P.k.zc; ts.P.slots[0].scene = 0;
    return 0;
}

// Can scene b ride in the same launch as scene a?  Everything but theta_max must agree: what the hot loop reads
// (KConsts except zc), the common Geom fields, the first event of the source rays and the kernel instance.
static bool same_launch(const TraceSetup& a, const TraceSetup& b) {
    KConsts ka = a.P.k, kb = b.P.k;
    ka.zc = kb.zc = 0.f;
    const Geom &ga = a.P.g, &gb = b.P.g;
    return a.rough == b.rough && a.model == b.model && a.P.kind0 == EV_WALL && b.P.kind0 == EV_WALL &&
           memcmp(&ka, &kb, sizeof ka) == 0 && ga.R1 == gb.R1 && ga.R2 == gb.R2 && ga.H == gb.H && ga.exit_z == gb.exit_z &&
           ga.lambertian == gb.lambertian && ga.brdf_kind == gb.brdf_kind && ga.max_bounces == gb.max_bounces &&
           ga.count_all == gb.count_all && memcmp(a.P.x0f, b.P.x0f, sizeof a.P.x0f) == 0 && memcmp(a.P.d0f, b.P.d0f, sizeof a.P.d0f) == 0 &&
           memcmp(&a.P.keys, &b.P.keys, sizeof a.P.keys) == 0;
}

template <bool R, int M, int S, int C>
static cudaError_t launch_trace_t(const TraceParams& P, altb_record* rec, unsigned int* counter, int blocks, cudaStream_t st) {
    static thread_local int attr_dev = -1;      // the opt-in to >48 kB of dynamic shared memory is per device and per kernel
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(k_trace<R, M, S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACE_SMEM);
        if (e != cudaSuccess) return e;
        attr_dev = dev;
    }
    k_trace<R, M, S, C><<<blocks, TRACE_THREADS, TRACE_SMEM, st>>>(P, rec, counter);
    return cudaGetLastError();
}

// kernel instance = (roughness on/off, reflection model, sink, arithmetic contract); the fast contract is built for the
// Lambert and CustomMirror models (0, 1)
template <int S, int C>
static cudaError_t launch_trace_sc(bool rough, int model, const TraceParams& P, altb_record* rec, unsigned int* counter, int blocks, cudaStream_t st) {
    if (rough) {
        if (model == 0) return launch_trace_t<true, 0, S, C>(P, rec, counter, blocks, st);
        if (model == 1) return launch_trace_t<true, 1, S, C>(P, rec, counter, blocks, st);
        if constexpr (C == CONTRACT_EXACT) {
            if (model == 2) return launch_trace_t<true, 2, S, C>(P, rec, counter, blocks, st);
            return launch_trace_t<true, 3, S, C>(P, rec, counter, blocks, st);
        }
    } else {
        if (model == 0) return launch_trace_t<false, 0, S, C>(P, rec, counter, blocks, st);
        if (model == 1) return launch_trace_t<false, 1, S, C>(P, rec, counter, blocks, st);
        if constexpr (C == CONTRACT_EXACT) {
            if (model == 2) return launch_trace_t<false, 2, S, C>(P, rec, counter, blocks, st);
            return launch_trace_t<false, 3, S, C>(P, rec, counter, blocks, st);
        }
    }
    return cudaErrorInvalidValue;
}
template <int C>
static cudaError_t launch_trace_c(int sink, bool rough, int model, const TraceParams& P, altb_record* rec, unsigned int* counter, int blocks, cudaStream_t st) {
    if (sink == SINK_LINES) return launch_trace_sc<SINK_LINES, C>(rough, model, P, rec, counter, blocks, st);
    if (sink != SINK_DIRECTION) return launch_trace_sc<SINK_RECORDS, C>(rough, model, P, rec, counter, blocks, st);
    if (P.n_slots > 1) return launch_trace_sc<SINK_DIRECTION_BATCHED, C>(rough, model, P, rec, counter, blocks, st);
    return launch_trace_sc<SINK_DIRECTION, C>(rough, model, P, rec, counter, blocks, st);
}

template <bool R, int M>
static cudaError_t launch_generic_t(const TraceParams& P, altb_record* rec, cudaStream_t st) {
    k_trace_generic<R, M><<<(P.n + 127) / 128, 128, 0, st>>>(P, rec);
    return cudaGetLastError();
}

// Trace rays [ray_id0, ray_id0+n) of every slot of ts.P.  sink == SINK_RECORDS: into rec[slot * n + i];
// SINK_DIRECTION: into P.counts_base / P.stats_base (set by the caller).  The caller keeps [ray_id0, ray_id0+n) inside one
// 2^32-aligned window of ray ids (pieces()), so that the high counter word is uniform.
static int run_trace(altb_ctx* ctx, DevCtx& d, TraceSetup& ts, int sink, uint64_t ray_id0, uint32_t n, altb_record* rec,
                     unsigned int* counter, QEntry* rq, unsigned long long* gstat, cudaStream_t st) {
    if (n == 0) return 0;
    TraceParams& P = ts.P;
    if (ts.rescatter) {         // brdf_kind 3: the second stage works on the records of the first
        if (sink != SINK_RECORDS) return fail(ALTB_E_ARG, "run_trace: brdf_kind 3 goes through the record path");
        P.ray_id0 = ray_id0; P.n = n; P.sincos = d.sincos;
    }
    auto rescatter = [&]() -> int {
        if (ts.rough) k_rescatter<true><<<(n + 127) / 128, 128, 0, st>>>(P, rec);
        else k_rescatter<false><<<(n + 127) / 128, 128, 0, st>>>(P, rec);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    };
    if (P.kind0 == EV_EXIT) {   // the source points straight out of the port: every ray is the same record
        altb_record proto;
        for (int i = 0; i < 3; i++) { proto.pos[i] = (float)P.x0[i]; proto.dir[i] = (float)P.d0[i]; }
        proto.n_hits = 0; proto.status = ALTB_EXITED;
        if (sink == SINK_DIRECTION)
            k_all_exit_direction<<<1, 32, 0, st>>>(proto, n, P.n_theta, P.n_phi, P.dir_tab, P.k.exit_zf, P.counts_base + (size_t)P.slots[0].scene * P.nb,
                                                   P.stats_base + (size_t)P.slots[0].scene * 8);
        else k_fill_records<<<d.sm_count * 4, 256, 0, st>>>(rec, n, proto);
        ctx->launches++;
        CK(cudaGetLastError());
        return ts.rescatter ? rescatter() : 0;
    }
    P.ray_id0 = ray_id0; P.ctr_lo0 = (uint32_t)ray_id0; P.ctr_hi = (uint32_t)(ray_id0 >> 32);
    P.n = n;
    P.sincos = d.sincos; P.rq = rq; P.gstat = gstat;
    P.lane_acc = gstat + (size_t)d.sm_count * MAX_SLOTS * STAT_WORDS;       // behind the block statistics, same allocation
    uint32_t sbits = 1;
    while ((1u << sbits) < P.n_slots) sbits++;
    P.shift = 32 - sbits; P.imask = (1u << P.shift) - 1u;
    if ((uint64_t)n >= (1ull << P.shift)) return fail(ALTB_E_ARG, "run_trace: %u rays x %u slots do not fit one launch", n, P.n_slots);
    if (P.n_slots > 1 && sink != SINK_DIRECTION) return fail(ALTB_E_ARG, "run_trace: batched scenes need the direction sink");
    if (sink == SINK_LINES && P.kind0 != EV_WALL) return fail(ALTB_E_ARG, "run_trace: the lines sink needs a source whose first event is the wall");
    const uint64_t total = (uint64_t)n * P.n_slots;
    int blocks = d.sm_count;                           // persistent: one 1024-thread block per SM
    const uint64_t warps_needed = (total + 31) / 32;
    if ((uint64_t)blocks * TRACE_WARPS > warps_needed) blocks = (int)((warps_needed + TRACE_WARPS - 1) / TRACE_WARPS);
    // ids are claimed in chunks; small enough that the tail (last chunk per warp) stays short
    uint32_t chunk = 256;
    while (chunk > 32 && (uint64_t)chunk * blocks * TRACE_WARPS * 4 > total) chunk >>= 1;
    if (const char* e = getenv("ALTB_CHUNK")) { const long v = atol(e); if (v >= 32 && v <= 65536) chunk = (uint32_t)v; }   // tuning knob
    P.chunk = chunk;
    P.cps = (n + chunk - 1) / chunk;
    P.n_chunks = P.cps * P.n_slots;
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    if (sink != SINK_RECORDS) CK(cudaMemsetAsync(gstat, 0, (size_t)blocks * P.n_slots * STAT_WORDS * sizeof(unsigned long long), st));
    const bool lane_acc = ALTB_LANE_ACC && sink != SINK_RECORDS && P.n_slots == 1 && P.kind0 == EV_WALL;
    if (lane_acc) CK(cudaMemsetAsync(P.lane_acc, 0, (size_t)blocks * TRACE_THREADS * 2 * sizeof(unsigned long long), st));
    cudaError_t le = cudaSuccess;
    if (P.kind0 != EV_WALL) {   // source aimed at the port rim: generic tracer (k_trace's fresh rays start on the sphere)
        if (sink != SINK_RECORDS || P.n_slots != 1) return fail(ALTB_E_ARG, "run_trace: rim-aimed sources go through the record path");
#define GEN(RR, MM) le = launch_generic_t<RR, MM>(P, rec, st)
        if (ts.rough) { if (ts.model == 0) GEN(true, 0); else if (ts.model == 1) GEN(true, 1); else if (ts.model == 2) GEN(true, 2); else GEN(true, 3); }
        else          { if (ts.model == 0) GEN(false, 0); else if (ts.model == 1) GEN(false, 1); else if (ts.model == 2) GEN(false, 2); else GEN(false, 3); }
#undef GEN
    } else {
        // fast instances: Lambert / CustomMirror with the small-angle flags their hot loop has compiled in (bounce_step)
        const bool fast = ctx->contract != ALTB_CONTRACT_EXACT && ts.model <= 1 && (!ts.rough || P.k.tilt_small == 2) &&
                          (ts.model != 1 || P.k.spec_small == 1);
        le = !fast ? launch_trace_c<CONTRACT_EXACT>(sink, ts.rough, ts.model, P, rec, counter, blocks, st)
           : ctx->contract == ALTB_CONTRACT_FAST7 ? launch_trace_c<CONTRACT_FAST7>(sink, ts.rough, ts.model, P, rec, counter, blocks, st)
                                                  : launch_trace_c<CONTRACT_FAST>(sink, ts.rough, ts.model, P, rec, counter, blocks, st);
    }
    ctx->launches++;
    ctx->trace_launches++;
    CK(le);
    if (sink != SINK_RECORDS) {
        if (lane_acc) { k_reduce_lane_acc<<<blocks, TRACE_THREADS, 0, st>>>(P); ctx->launches++; }
        k_reduce_trace_stats<<<(P.n_slots + 63) / 64, 64, 0, st>>>(P, blocks);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return ts.rescatter ? rescatter() : 0;
}

// [off, off+len) pieces of a ray-id range: at most `cap` rays each, none straddling a multiple of 2^32 of the GLOBAL id
static inline uint64_t piece_len(uint64_t ray_id, uint64_t left, uint64_t cap) {
    const uint64_t to_wrap = (1ull << 32) - (ray_id & 0xffffffffull);
    return std::min(std::min(left, cap), to_wrap);
}

// ---------------------------------------------------------------------------------- map setup
struct MapSetup { MapParams M; int n_tiles; size_t line_smem; size_t dir_smem; size_t rect_smem; };

static int setup_map(altb_ctx* ctx, DevCtx& d, const altb_scene* sc, const Geom& g, const KConsts& k, const altb_map_spec* map,
                     MapSetup& ms, cudaStream_t st) {
    (void)sc; (void)ctx;
    if (map->n_theta < 1 || map->n_phi < 1 || (int64_t)map->n_theta * map->n_phi > (1 << 24))
        return fail(ALTB_E_ARG, "map: bad bin counts %d x %d", map->n_theta, map->n_phi);
    if (map->map_mode < ALTB_MAP_LINE || map->map_mode > ALTB_MAP_TWOFOLD) return fail(ALTB_E_ARG, "map: bad map_mode %d", map->map_mode);
    const bool grouped = map->map_mode == ALTB_MAP_PER_POSITION || map->map_mode == ALTB_MAP_TWOFOLD;
    if (grouped && map->rays_per_position < 1) return fail(ALTB_E_ARG, "map: rays_per_position must be >= 1");
    if (map->map_mode == ALTB_MAP_TWOFOLD && (map->n_phi & 1)) return fail(ALTB_E_ARG, "map: twofold needs an even n_phi");
    MapParams& M = ms.M;
    memset(&M, 0, sizeof M);
    const int nt = map->n_theta, np = map->n_phi;
    M.n_theta = nt; M.n_phi = np; M.mode = map->map_mode; M.rays_per_position = map->rays_per_position;
    M.count_all = g.count_all; M.exit_zf = k.exit_zf;
    const double hw = map->det_width / 2;
    M.w2 = (float)(hw * hw);
    ms.dir_smem = (size_t)nt * np * sizeof(unsigned int);
    M.use_smem_hist = ms.dir_smem <= 160 * 1024;
    if (!M.use_smem_hist) ms.dir_smem = 0;
    ms.n_tiles = 0; ms.line_smem = 0; ms.rect_smem = 0;
    if (map->map_mode == ALTB_MAP_DIRECTION) {
        // bin edges of direction_bin: cos(i w_theta), i = 0..n_theta; (cos, sin)(j w_phi), j = 0..n_phi (the last one = the first)
        if (d.dir_nt != nt || d.dir_np != np) {
            if (d.dir_nt) CK(cudaDeviceSynchronize());                    // kernels of an earlier call may still read the old table
            std::vector<double>& t = d.dirtab_host;
            t.assign((size_t)nt + 1 + 2 * ((size_t)np + 1), 0.0);
            for (int i = 0; i <= nt; i++) t[i] = i == 0 ? 1.0 : cos(i * (90.0 / nt) * PI_D / 180.0);
            for (int j = 0; j <= np; j++) {
                const double a = (j % np) * (360.0 / np) * PI_D / 180.0;
                t[(size_t)nt + 1 + j] = cos(a); t[(size_t)nt + 1 + np + 1 + j] = sin(a);
            }
            if (int rc = ensure(d.dirtab, d.dirtab_cap, (uint64_t)t.size())) return rc;
            CK(cudaMemcpyAsync(d.dirtab, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            d.dir_nt = nt; d.dir_np = np;
        }
        M.dir_tab = d.dirtab;
        return 0;
    }
    (void)grouped;
    if (!(map->det_radius > 0) || !(map->det_width > 0)) return fail(ALTB_E_ARG, "map: det_radius/det_width must be > 0");

    // The tables depend on the map spec only: a context that keeps mapping with the same spec (every step of a run, every
    // scene of a sweep) builds and uploads them once.
    if (d.map_cached && d.map_key.n_theta == nt && d.map_key.n_phi == np && d.map_key.det_radius == map->det_radius &&
        d.map_key.det_width == map->det_width) {
        M.t_theta = d.map_M.t_theta; M.t_phi = d.map_M.t_phi; M.nt_theta = d.map_M.nt_theta; M.nt_phi = d.map_M.nt_phi;
        M.rs = d.map_M.rs; M.pz = d.map_M.pz; M.st = d.map_M.st; M.ct = d.map_M.ct; M.cp = d.map_M.cp; M.sp = d.map_M.sp;
        M.tiles = d.map_M.tiles; M.supers = d.map_M.supers;
        M.row4 = d.map_M.row4; M.col2 = d.map_M.col2; M.det_R = d.map_M.det_R; M.det_Wr = d.map_M.det_Wr;
        M.rc2 = d.map_M.rc2; M.RR = d.map_M.RR; M.n2R = d.map_M.n2R; M.p2R = d.map_M.p2R;
        ms.n_tiles = d.map_n_tiles; ms.line_smem = d.map_line_smem; ms.rect_smem = d.map_rect_smem;
        return 0;
    }
    if (d.map_cached) { CK(cudaDeviceSynchronize()); d.map_cached = false; }     // kernels of an earlier call may still read the old tables
    // per-row / per-column tables (Detector::setPosition, fluxAtObserverFast.C:61-80), rounded once to f32
    std::vector<float>& tab = d.tab_host;
    tab.assign((size_t)5 * nt + 2 * np, 0.f);        // rs, pz, st, ct [nt each], cp, sp [np each], R ct^2 [nt]
    std::vector<double> px((size_t)nt * np), py((size_t)nt * np), pzv((size_t)nt * np);
    for (int i = 0; i < nt; i++) {
        const double th = (i + 0.5) * 90.0 / nt * PI_D / 180.0;
        tab[i] = (float)(map->det_radius * sin(th));            // rs
        tab[nt + i] = (float)(-100.0 - map->det_radius * cos(th));  // pz
        tab[2 * nt + i] = (float)sin(th);                       // st
        tab[3 * nt + i] = (float)cos(th);                       // ct
        tab[4 * nt + 2 * np + i] = (float)(map->det_radius * cos(th) * cos(th));   // R ct^2: the row part of (L - p).n (line_hit)
    }
    for (int j = 0; j < np; j++) {
        const double ph = (j + 0.5) * 360.0 / np * PI_D / 180.0;
        tab[4 * nt + j] = (float)cos(ph);
        tab[4 * nt + np + j] = (float)sin(ph);
    }
    for (int i = 0; i < nt; i++)
        for (int j = 0; j < np; j++) {
            px[(size_t)i * np + j] = (double)tab[i] * tab[4 * nt + j];
            py[(size_t)i * np + j] = (double)tab[i] * tab[4 * nt + np + j];
            pzv[(size_t)i * np + j] = tab[nt + i];
        }
    // tile shape: the (t_theta x t_phi) <= 32 bins with the smallest mean bounding radius
    const int shapes[6][2] = {{32, 1}, {16, 2}, {8, 4}, {4, 8}, {2, 16}, {1, 32}};
    double best = 1e300; int bt = 8, bp = 4;
    std::vector<float4>& best_tiles = d.tiles_host;     // staging lives in the context: nothing to wait for after the uploads
    best_tiles.clear();
    auto bound = [&](int i0, int i1, int j0, int j1) {     // bounding sphere of the detector centres of bins [i0,i1) x [j0,j1)
        double cx = 0, cy = 0, cz = 0; int cnt = 0;
        for (int i = i0; i < i1; i++)
            for (int j = j0; j < j1; j++) { cx += px[(size_t)i * np + j]; cy += py[(size_t)i * np + j]; cz += pzv[(size_t)i * np + j]; cnt++; }
        cx /= cnt; cy /= cnt; cz /= cnt;
        double r = 0;
        for (int i = i0; i < i1; i++)
            for (int j = j0; j < j1; j++) {
                const double dx = px[(size_t)i * np + j] - cx, dy = py[(size_t)i * np + j] - cy, dz = pzv[(size_t)i * np + j] - cz;
                r = std::max(r, sqrt(dx * dx + dy * dy + dz * dz));
            }
        const double rad = hw + r + 0.5;     // conservative: disk radius + tile radius + f32 slack [cm]
        return make_float4((float)cx, (float)cy, (float)cz, (float)(rad * rad * 1.0001));
    };
    for (auto& s : shapes) {
        const int tt = s[0], tp = s[1];
        const int ntt = (nt + tt - 1) / tt, ntp = (np + tp - 1) / tp;
        if ((size_t)ntt * ntp > 4096) continue;
        std::vector<float4> tl((size_t)ntt * ntp);
        double sum = 0;
        for (int a = 0; a < ntt; a++)
            for (int b = 0; b < ntp; b++) {
                double cx = 0, cy = 0, cz = 0; int cnt = 0;
                for (int i = a * tt; i < std::min(nt, (a + 1) * tt); i++)
                    for (int j = b * tp; j < std::min(np, (b + 1) * tp); j++) {
                        cx += px[(size_t)i * np + j]; cy += py[(size_t)i * np + j]; cz += pzv[(size_t)i * np + j]; cnt++;
                    }
                cx /= cnt; cy /= cnt; cz /= cnt;
                double r = 0;
                for (int i = a * tt; i < std::min(nt, (a + 1) * tt); i++)
                    for (int j = b * tp; j < std::min(np, (b + 1) * tp); j++) {
                        const double dx = px[(size_t)i * np + j] - cx, dy = py[(size_t)i * np + j] - cy, dz = pzv[(size_t)i * np + j] - cz;
                        r = std::max(r, sqrt(dx * dx + dy * dy + dz * dz));
                    }
                sum += r;
                const double rad = hw + r + 0.5;     // conservative: disk radius + tile radius + f32 slack [cm]
                tl[(size_t)a * ntp + b] = make_float4((float)cx, (float)cy, (float)cz, (float)(rad * rad * 1.0001));
            }
        const double mean = sum / ((double)ntt * ntp) + 1e-9 * ntt * ntp;   // tie-break: fewer tiles
        if (mean < best) { best = mean; bt = tt; bp = tp; best_tiles.swap(tl); }
    }
    M.t_theta = bt; M.t_phi = bp;
    M.nt_theta = (nt + bt - 1) / bt; M.nt_phi = (np + bp - 1) / bp;
    ms.n_tiles = M.nt_theta * M.nt_phi;
    const int nst = (M.nt_theta + SUPER - 1) / SUPER, nsp = (M.nt_phi + SUPER - 1) / SUPER;
    for (int a = 0; a < nst; a++)
        for (int b = 0; b < nsp; b++)
            best_tiles.push_back(bound(a * SUPER * bt, std::min(nt, (a + 1) * SUPER * bt), b * SUPER * bp, std::min(np, (b + 1) * SUPER * bp)));
    ms.line_smem = (size_t)LINE_BATCH * 2 * sizeof(float4) + (size_t)ms.n_tiles * LINE_WORDS * sizeof(uint32_t) +
                   (size_t)(ms.n_tiles + nst * nsp) * sizeof(float4) + (size_t)nst * nsp * sizeof(uint32_t) + ((size_t)3 * nt + 2 * np) * sizeof(float);
    if (ms.line_smem > 200 * 1024) return fail(ALTB_E_ARG, "map: %d x %d bins need %zu B of shared memory", nt, np, ms.line_smem);
    // packed copies for the ray-stationary kernel: (st, ct, R ct^2, 0) per row, (cp, sp) per column, 16-byte aligned behind the flat tables
    const size_t flat = ((size_t)5 * nt + 2 * np + 3) & ~(size_t)3;
    tab.resize(flat + (size_t)4 * nt + 2 * np);
    for (int i = 0; i < nt; i++) {
        tab[flat + 4 * i] = tab[2 * nt + i]; tab[flat + 4 * i + 1] = tab[3 * nt + i]; tab[flat + 4 * i + 2] = tab[4 * nt + 2 * np + i];
        tab[flat + 4 * i + 3] = 0.f;
    }
    for (int j = 0; j < np; j++) { tab[flat + 4 * nt + 2 * j] = tab[4 * nt + j]; tab[flat + 4 * nt + 2 * j + 1] = tab[4 * nt + np + j]; }
    if (int rc = ensure(d.tables, d.tables_cap, (uint64_t)tab.size())) return rc;
    if (int rc = ensure(d.tiles, d.tiles_cap, (uint64_t)best_tiles.size())) return rc;
    CK(cudaMemcpyAsync(d.tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d.tiles, best_tiles.data(), best_tiles.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    M.rs = d.tables; M.pz = d.tables + nt; M.st = d.tables + 2 * nt; M.ct = d.tables + 3 * nt;
    M.cp = d.tables + 4 * nt; M.sp = d.tables + 4 * nt + np; M.rc2 = d.tables + 4 * nt + 2 * np;
    M.RR = (float)(map->det_radius * map->det_radius); M.n2R = (float)(-2.0 * map->det_radius); M.p2R = (float)(2.0 * map->det_radius);
    M.tiles = d.tiles; M.supers = d.tiles + ms.n_tiles;
    M.row4 = reinterpret_cast<const float4*>(d.tables + flat); M.col2 = reinterpret_cast<const float2*>(d.tables + flat + 4 * nt);
    M.det_R = (float)map->det_radius; M.det_Wr = (float)(hw + 0.05);      // f32 evaluation of the test moves the rim by < 1e-3 cm
    {
        const size_t nrp = ((size_t)nt + 1) / 2;
        ms.rect_smem = nrp * (sizeof(float4) + sizeof(float2)) + (size_t)RECT_THREADS * sizeof(float4) + (size_t)2 * np * sizeof(float2) +
                       2 * nrp * (size_t)rect_hist_stride(np) * sizeof(unsigned int);
    }
    d.map_key = *map; d.map_M = M; d.map_n_tiles = ms.n_tiles; d.map_line_smem = ms.line_smem; d.map_rect_smem = ms.rect_smem; d.map_cached = true;
    return 0;
}

// the two LINE-map kernels on the line buffer (rectangle list from the front, tile list from the back; n_lines[0] / [1])
static int run_line_kernels(altb_ctx* ctx, DevCtx& d, const MapSetup& ms, const MapParams& Mp, const float4* lines, uint32_t lines_cap,
                            const unsigned int* n_lines, uint32_t n_hint, unsigned long long* d_counts, cudaStream_t st) {
    if (!Mp.force_tiles) {
        if (ms.rect_smem > 48 * 1024) CK(cudaFuncSetAttribute(k_map_line_rect, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_map_line_rect, RECT_THREADS, ms.rect_smem));
        if (per_sm < 1) per_sm = 1;
        int blocks = d.sm_count * per_sm;
        const int need = (int)((n_hint + 8 * (RECT_THREADS / 32) - 1) / (8 * (RECT_THREADS / 32)));     // >= 8 rays per warp, or fewer blocks
        if (blocks > need) blocks = std::max(need, 1);
        k_map_line_rect<<<blocks, RECT_THREADS, ms.rect_smem, st>>>(lines, n_lines, Mp, d_counts);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaFuncSetAttribute(k_map_line, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 1;    // persistent blocks: exactly the resident count (registers, shared memory, threads)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_map_line, LINE_THREADS, ms.line_smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = d.sm_count * per_sm;
    const int need = (int)((n_hint + LINE_BATCH - 1) / LINE_BATCH);
    if (blocks > need) blocks = need;
    k_map_line<<<blocks, LINE_THREADS, ms.line_smem, st>>>(lines + 2 * (size_t)lines_cap, n_lines + 1, Mp, d_counts);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// records d.rec[0..n) -> counts (+stats)
static int run_map(altb_ctx* ctx, DevCtx& d, const MapSetup& ms, const altb_record* rec, unsigned int* counter, uint32_t n, uint64_t ray_base,
                   unsigned long long* d_counts, unsigned long long* d_stats, int* d_bin, cudaStream_t st) {
    if (n == 0) return 0;
    const MapParams& M = ms.M;
    if (M.mode == ALTB_MAP_DIRECTION) {
        if (ms.dir_smem > 48 * 1024)      // per device: cheap enough to repeat
            CK(cudaFuncSetAttribute(k_map_direction, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        int per_sm = 1;    // the 64.8 kB histogram limits residency: more warps per block, grid = resident blocks
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_map_direction, DIR_THREADS, ms.dir_smem));
        int blocks = d.sm_count * (per_sm < 1 ? 1 : per_sm);
        const int need = (int)((n + DIR_THREADS - 1) / DIR_THREADS);
        if (blocks > need) blocks = need;
        k_map_direction<<<blocks, DIR_THREADS, ms.dir_smem, st>>>(rec, n, M, d_counts, d_stats, d_bin);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (d_stats) {
        int blocks = d.sm_count * 4;
        const int need = (int)((n + 255) / 256);
        if (blocks > need) blocks = need;
        k_stats<<<blocks, 256, 0, st>>>(rec, n, M.count_all, M.exit_zf, d_stats);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (M.mode == ALTB_MAP_PER_POSITION || M.mode == ALTB_MAP_TWOFOLD) {
        int blocks = d.sm_count * 8;
        const int need = (int)((n + 255) / 256);
        if (blocks > need) blocks = need;
        k_map_per_position<<<blocks, 256, 0, st>>>(rec, n, M, (unsigned long long)ray_base, d_counts);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    // ---- LINE / TRACEONCE_COMPAT on records: escaping rays -> test lines (+ candidate rectangles) -> the two map kernels
    if (int rc = ensure(d.lines, d.lines_cap, 2 * (uint64_t)std::max<uint64_t>(d.rec_cap, n))) return rc;      // 2 float4 per ray, both lists
    const uint32_t lines_cap = (uint32_t)std::min<uint64_t>(d.lines_cap / 2, 0xffffffffull);
    MapParams Mp = M;
    Mp.force_tiles = ms.rect_smem > 200 * 1024 || getenv("ALTB_LINE_TILES") != nullptr;
    CK(cudaMemsetAsync(counter + 2, 0, 2 * sizeof(unsigned int), st));
    {
        int cb = d.sm_count * 8;
        const int need = (int)((n + 255) / 256);
        if (cb > need) cb = need;
        k_prepare_lines<<<cb, 256, 0, st>>>(rec, n, Mp, d.lines, lines_cap, counter + 2);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return run_line_kernels(ctx, d, ms, Mp, d.lines, lines_cap, counter + 2, n, d_counts, st);
}

// ---------------------------------------------------------------------------------- hot path
// One launch slot of a job: stream, chunk counter, resume queues and (record sink) the record buffer.  Jobs with several
// launches and no per-kernel timing alternate between two slots on two worker streams forked from / joined to the caller's
// stream, so that the tail of one persistent launch (its last long rays) overlaps the start of the next one.
struct LaunchSlot { cudaStream_t st; unsigned int* counter; QEntry* rq; unsigned long long* gstat; altb_record* rec; };

static int fluxmap_on_device(altb_ctx* ctx, DevCtx& d, const altb_scene* scenes, int n_scenes, const altb_source* src,
                             uint64_t ray_id0, uint64_t n_rays, uint64_t seed, const altb_map_spec* map,
                             unsigned long long* d_counts, unsigned long long* d_stats, cudaStream_t st,
                             float* t_trace_ms, float* t_map_ms) {
    CK(cudaSetDevice(d.dev));
    const uint64_t nb = (uint64_t)map->n_theta * map->n_phi;
    if (n_rays == 0) return 0;
    std::vector<TraceSetup> tss((size_t)n_scenes);
    for (int s = 0; s < n_scenes; s++)
        if (int rc = setup_trace(&scenes[s], src, seed, tss[s])) return rc;
    MapSetup ms;                                         // tables depend on the map spec only; count_all / exit_z are patched per scene
    if (int rc = setup_map(ctx, d, &scenes[0], tss[0].P.g, tss[0].P.k, map, ms, st)) return rc;
    if (!d_stats) {                                      // the direction sink always keeps statistics
        if (int rc = ensure(d.stats_scratch, d.stats_scratch_cap, (uint64_t)n_scenes * 8)) return rc;
        d_stats = d.stats_scratch;
    }
    // ---- plan: which scenes share launches
    const bool dir_mode = map->map_mode == ALTB_MAP_DIRECTION;
    auto dir_sink_ok = [&](int s) { return dir_mode && !tss[s].rescatter && !tss[s].P.g.count_all && (tss[s].P.kind0 == EV_WALL || tss[s].P.kind0 == EV_EXIT); };
    const bool line_mode = map->map_mode == ALTB_MAP_LINE || map->map_mode == ALTB_MAP_TRACEONCE_COMPAT;
    auto lines_sink_ok = [&](int s) { return line_mode && !tss[s].rescatter && !tss[s].P.g.count_all && tss[s].P.kind0 == EV_WALL && !getenv("ALTB_LINE_RECORDS"); };
    std::vector<std::vector<int>> groups;
    std::vector<char> seen((size_t)n_scenes, 0);
    for (int s = 0; s < n_scenes; s++) {
        if (seen[s]) continue;
        seen[s] = 1;
        groups.push_back({s});
        if (!dir_sink_ok(s) || tss[s].P.kind0 != EV_WALL || getenv("ALTB_NO_BATCH")) continue;
        for (int t = s + 1; t < n_scenes && (int)groups.back().size() < MAX_SLOTS; t++)
            if (!seen[t] && dir_sink_ok(t) && same_launch(tss[s], tss[t])) { seen[t] = 1; groups.back().push_back(t); }
    }
    // ---- launches of the whole job, to decide about the overlap
    const uint64_t batch_rec = std::min<uint64_t>(std::min<uint64_t>(ctx->batch, (1ull << 31) - 1), n_rays);
    auto dir_cap = [&](size_t G) -> uint64_t {          // rays per slot of one direction-sink launch of G slots
        uint32_t sbits = 1;
        while ((1u << sbits) < G) sbits++;
        const uint64_t total = ctx->batch_user ? ctx->batch : 0xfff00000ull;      // dense work index < 2^32
        return std::max<uint64_t>(1, std::min<uint64_t>(total / G, (1ull << (32 - sbits)) - 1));
    };
    uint64_t n_launches = 0;
    bool any_rec = false;
    for (auto& g : groups) {
        const bool dsink = dir_sink_ok(g[0]);
        const uint64_t cap = dsink ? dir_cap(g.size()) : batch_rec;
        n_launches += (n_rays + cap - 1) / cap;
        any_rec |= !dsink && !lines_sink_ok(g[0]);
    }
    if (any_rec) if (int rc = ensure(d.rec, d.rec_cap, batch_rec)) return rc;
    const bool overlap = !t_trace_ms && dir_mode && n_launches >= 2 && !getenv("ALTB_NO_OVERLAP");
    LaunchSlot ls[2] = {{st, d.counter, d.rq[0], d.gstat[0], d.rec}, {st, d.counter2, d.rq[1], d.gstat[1], d.rec}};
    // the worker streams are joined to the caller's stream on EVERY way out of this function once they are forked (the CK /
    // ensure returns inside the launch loop included): work queued behind `st` must never overtake a launch still running
    struct Join {
        DevCtx& d; cudaStream_t st; bool armed;
        cudaError_t run() {
            cudaError_t first = cudaSuccess;
            if (armed) for (int i = 0; i < 2; i++) {
                cudaError_t e = cudaEventRecord(d.join_ev[i], d.aux[i]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(st, d.join_ev[i], 0);
                if (e != cudaSuccess && first == cudaSuccess) first = e;
            }
            armed = false;
            return first;
        }
        ~Join() { run(); }
    } join{d, st, false};
    if (overlap) {
        if (any_rec) { if (int rc = ensure(d.rec2, d.rec2_cap, batch_rec)) return rc; ls[1].rec = d.rec2; }
        CK(cudaEventRecord(d.fork_ev, st));
        join.armed = true;
        for (int i = 0; i < 2; i++) { CK(cudaStreamWaitEvent(d.aux[i], d.fork_ev, 0)); ls[i].st = d.aux[i]; }
    }
    uint64_t launch_no = 0;
    int rc_all = 0;
    for (size_t gi = 0; gi < groups.size() && !rc_all; gi++) {
        const std::vector<int>& g = groups[gi];
        TraceSetup& ts = tss[g[0]];
        const bool dsink = dir_sink_ok(g[0]), lsink = lines_sink_ok(g[0]);
        float tt = 0.f, tm = 0.f;
        MapParams Ml = ms.M;
        if (lsink) {
            ts.P.slots[0].scene = (uint32_t)g[0];
            ts.P.stats_base = d_stats; ts.P.counts_base = d_counts; ts.P.nb = (uint32_t)nb; ts.P.n_theta = map->n_theta; ts.P.n_phi = map->n_phi;
            Ml.force_tiles = ms.rect_smem > 200 * 1024 || getenv("ALTB_LINE_TILES") != nullptr;
            ts.P.rp = {map->n_theta, map->n_phi, Ml.force_tiles, map->map_mode == ALTB_MAP_TRACEONCE_COMPAT, Ml.det_R, Ml.det_Wr};
            if (int rc = ensure(d.lines, d.lines_cap, 2 * (uint64_t)std::max<uint64_t>(d.rec_cap, batch_rec))) return rc;
            if (int rc = ensure(d.raw, d.raw_cap, 2 * (uint64_t)batch_rec)) return rc;
            ts.P.lines = d.raw; ts.P.lines_cap = (uint32_t)std::min<uint64_t>(d.raw_cap / 2, 0xffffffffull);
        }
        if (dsink) {
            ts.P.n_slots = (uint32_t)g.size();
            for (size_t j = 0; j < g.size(); j++) { ts.P.slots[j] = tss[g[j]].P.slots[0]; ts.P.slots[j].scene = (uint32_t)g[j]; }
            ts.P.counts_base = d_counts; ts.P.stats_base = d_stats;
            ts.P.nb = (uint32_t)nb; ts.P.n_theta = map->n_theta; ts.P.n_phi = map->n_phi; ts.P.dir_tab = ms.M.dir_tab;
        }
        const uint64_t cap = dsink ? dir_cap(g.size()) : batch_rec;
        MapSetup msg = ms;
        msg.M.count_all = ts.P.g.count_all; msg.M.exit_zf = ts.P.k.exit_zf;
        for (uint64_t off = 0; off < n_rays && !rc_all; launch_no++) {
            const uint32_t n = (uint32_t)piece_len(ray_id0 + off, n_rays - off, cap);
            const LaunchSlot& L = ls[overlap ? (launch_no & 1) : 0];
            if (t_trace_ms) CK(cudaEventRecord(d.ev[0], L.st));
            if (lsink) { ts.P.n_lines = L.counter + 1; CK(cudaMemsetAsync(L.counter + 1, 0, 3 * sizeof(unsigned int), L.st)); }     // [1] raw, [2] rect, [3] tile
            rc_all = run_trace(ctx, d, ts, dsink ? SINK_DIRECTION : (lsink ? SINK_LINES : SINK_RECORDS), ray_id0 + off, n, L.rec, L.counter, L.rq, L.gstat, L.st);
            if (!rc_all && t_trace_ms) CK(cudaEventRecord(d.ev[1], L.st));
            if (!rc_all && lsink) {
                const uint32_t lcap = (uint32_t)std::min<uint64_t>(d.lines_cap / 2, 0xffffffffull);
                int cb = d.sm_count * 8;
                const int need = (int)((n + 255) / 256);
                if (cb > need) cb = need;
                k_prepare_raw<<<cb, 256, 0, L.st>>>(d.raw, L.counter + 1, ts.P.rp, d.lines, lcap, L.counter + 2);
                ctx->launches++;
                CK(cudaGetLastError());
                rc_all = run_line_kernels(ctx, d, ms, Ml, d.lines, lcap, L.counter + 2, n, d_counts + (size_t)g[0] * nb, L.st);
            }
            else if (!rc_all && !dsink)
                rc_all = run_map(ctx, d, msg, L.rec, L.counter, n, ray_id0 + off, d_counts + (size_t)g[0] * nb, d_stats + (size_t)g[0] * 8, nullptr, L.st);
            if (!rc_all && t_trace_ms) {
                CK(cudaEventRecord(d.ev[2], L.st));
                CK(cudaEventSynchronize(d.ev[2]));
                float a = 0.f, b = 0.f;
                CK(cudaEventElapsedTime(&a, d.ev[0], d.ev[1]));
                CK(cudaEventElapsedTime(&b, d.ev[1], d.ev[2]));
                tt += a; tm += b;
            }
            off += n;
        }
        // scenes that shared launches share their device time equally
        if (t_trace_ms) for (int sidx : g) { t_trace_ms[sidx] = tt / g.size(); t_map_ms[sidx] = tm / g.size(); }
    }
    if (overlap) {                                      // join (the guard above does the same on the early returns)
        const cudaError_t e = join.run();
        if (e != cudaSuccess && !rc_all) rc_all = fail(ALTB_E_CUDA, "fluxmap: stream join failed: %s", cudaGetErrorString(e));
    }
    return rc_all;
}

extern "C" int altb_trace_fluxmap_dev(altb_ctx* ctx, const altb_scene* scenes, int n_scenes, const altb_source* src,
                                      uint64_t ray_id0, uint64_t n_rays, uint64_t seed, const altb_map_spec* map,
                                      uint64_t* d_counts, uint64_t* d_stats, void* cuda_stream) {
    if (!ctx || !scenes || n_scenes < 1 || !src || !map || !d_counts) return fail(ALTB_E_ARG, "altb_trace_fluxmap_dev: NULL/empty argument");
    return fluxmap_on_device(ctx, ctx->devs[0], scenes, n_scenes, src, ray_id0, n_rays, seed, map,
                             reinterpret_cast<unsigned long long*>(d_counts), reinterpret_cast<unsigned long long*>(d_stats),
                             (cudaStream_t)cuda_stream, nullptr, nullptr);
}

extern "C" int altb_trace_fluxmap(altb_ctx* ctx, const altb_scene* scenes, int n_scenes, const altb_source* src,
                                  uint64_t ray_id0, uint64_t n_rays, uint64_t seed, const altb_map_spec* map,
                                  uint64_t* counts, altb_stats* stats) {
    if (!ctx || !scenes || n_scenes < 1 || !src || !map || !counts) return fail(ALTB_E_ARG, "altb_trace_fluxmap: NULL/empty argument");
    const uint64_t nb = (uint64_t)map->n_theta * map->n_phi;
    const int nd = (int)ctx->devs.size();
    const uint64_t words = (uint64_t)n_scenes * (nb + 8);
    // rays are split evenly over the context's devices; each accumulates its own map; the maps are merged by one NCCL
    // all-reduce (contexts with several devices own communicators) or, without NCCL on the host, summed on the host
    std::vector<std::vector<float>> tms(nd, std::vector<float>(2 * n_scenes, 0.f));
    for (int i = 0; i < nd; i++) {
        DevCtx& d = ctx->devs[i];
        CK(cudaSetDevice(d.dev));
        if (int rc = ensure(d.counts, d.counts_cap, words)) return rc;
        CK(cudaMemsetAsync(d.counts, 0, words * sizeof(unsigned long long), d.stream));
    }
    // launch everything first when more than one device is used (async), time per device when one
    for (int i = 0; i < nd; i++) {
        DevCtx& d = ctx->devs[i];
        const uint64_t lo = n_rays * i / nd, hi = n_rays * (i + 1) / nd;
        float* tt = nd == 1 ? tms[i].data() : nullptr;
        float* tm = nd == 1 ? tms[i].data() + n_scenes : nullptr;
        if (int rc = fluxmap_on_device(ctx, d, scenes, n_scenes, src, ray_id0 + lo, hi - lo, seed, map,
                                       d.counts, d.counts + (size_t)n_scenes * nb, d.stream, tt, tm)) return rc;
    }
    std::vector<unsigned long long> host(words);
    if (stats) memset(stats, 0, sizeof(altb_stats) * n_scenes);
    auto add_host = [&](int i) {
        for (uint64_t j = 0; j < (uint64_t)n_scenes * nb; j++) counts[j] += host[j];
        if (stats)
            for (int s = 0; s < n_scenes; s++) {
                const unsigned long long* p = host.data() + (size_t)n_scenes * nb + (size_t)s * 8;
                stats[s].n_rays += p[0]; stats[s].n_exited += p[1]; stats[s].n_exit_port += p[2];
                stats[s].n_absorbed += p[3]; stats[s].n_suspended += p[4]; stats[s].n_bounces += p[5];
                stats[s].t_trace_s += tms[i][s] * 1e-3; stats[s].t_map_s += tms[i][n_scenes + s] * 1e-3;
            }
    };
    if (nd > 1 && !ctx->comms.empty()) {
        // ONE all-reduce of [counts | stats] (uint64 sums: bit-identical whatever the device count), one copy to the host
        const NcclApi* N = nccl_api();
        ncclResult_t r = N->GroupStart();
        for (int i = 0; i < nd && r == ncclSuccess; i++) {
            DevCtx& d = ctx->devs[i];
            r = N->AllReduce(d.counts, d.counts, words, ncclUint64, ncclSum, ctx->comms[i], d.stream);
        }
        const ncclResult_t r2 = N->GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) return fail(ALTB_E_CUDA, "altb_trace_fluxmap: ncclAllReduce failed: %s", N->GetErrorString(r));
        DevCtx& d0 = ctx->devs[0];
        CK(cudaSetDevice(d0.dev));
        CK(cudaMemcpyAsync(host.data(), d0.counts, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d0.stream));
        for (int i = 0; i < nd; i++) { CK(cudaSetDevice(ctx->devs[i].dev)); CK(cudaStreamSynchronize(ctx->devs[i].stream)); }
        add_host(0);
        return 0;
    }
    for (int i = 0; i < nd; i++) {
        DevCtx& d = ctx->devs[i];
        CK(cudaSetDevice(d.dev));
        CK(cudaMemcpyAsync(host.data(), d.counts, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        add_host(i);
    }
    return 0;
}

// ---------------------------------------------------------------------------------- per-ray results
static void read_stats(const unsigned long long* p, altb_stats* s) {
    memset(s, 0, sizeof *s);
    s->n_rays = p[0]; s->n_exited = p[1]; s->n_exit_port = p[2]; s->n_absorbed = p[3]; s->n_suspended = p[4]; s->n_bounces = p[5];
}

extern "C" int altb_trace_records(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                                  uint64_t n_rays, uint64_t seed, altb_record* records, altb_stats* stats) {
    if (!ctx || !scene || !src || (!records && n_rays)) return fail(ALTB_E_ARG, "altb_trace_records: NULL argument");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    TraceSetup ts;
    if (int rc = setup_trace(scene, src, seed, ts)) return rc;
    const uint64_t batch = std::min<uint64_t>(ctx->batch, std::max<uint64_t>(n_rays, 1));
    if (int rc = ensure(d.rec, d.rec_cap, batch)) return rc;
    CK(cudaMemsetAsync(d.stats, 0, 8 * sizeof(unsigned long long), d.stream));
    CK(cudaEventRecord(d.ev[0], d.stream));
    for (uint64_t off = 0, n = 0; off < n_rays; off += n) {
        n = piece_len(ray_id0 + off, n_rays - off, batch);
        if (int rc = run_trace(ctx, d, ts, SINK_RECORDS, ray_id0 + off, (uint32_t)n, d.rec, d.counter, d.rq[0], d.gstat[0], d.stream)) return rc;
        if (stats) {
            k_stats<<<d.sm_count * 4, 256, 0, d.stream>>>(d.rec, n, ts.P.g.count_all, ts.P.k.exit_zf, d.stats);
            ctx->launches++;
        }
        CK(cudaMemcpyAsync(records + off, d.rec, (size_t)n * sizeof(altb_record), cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    CK(cudaEventRecord(d.ev[1], d.stream));
    CK(cudaEventSynchronize(d.ev[1]));
    if (stats) {
        unsigned long long h[8];
        CK(cudaMemcpy(h, d.stats, sizeof h, cudaMemcpyDeviceToHost));
        read_stats(h, stats);
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        stats->t_trace_s = ms * 1e-3;
    }
    return 0;
}

extern "C" int altb_trace_exit_rays(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                                    uint64_t n_rays, uint64_t seed, double* last_pos, double* last_dir,
                                    uint32_t* n_points, uint8_t* status, altb_stats* stats) {
    if (!ctx || !scene || !src) return fail(ALTB_E_ARG, "altb_trace_exit_rays: NULL argument");
    const uint64_t step = 1ull << 22;
    std::vector<altb_record> buf((size_t)std::min<uint64_t>(step, std::max<uint64_t>(n_rays, 1)));
    altb_stats tot; memset(&tot, 0, sizeof tot);
    for (uint64_t off = 0; off < n_rays; off += step) {
        const uint64_t n = std::min<uint64_t>(step, n_rays - off);
        altb_stats s;
        if (int rc = altb_trace_records(ctx, scene, src, ray_id0 + off, n, seed, buf.data(), &s)) return rc;
        for (uint64_t i = 0; i < n; i++) {
            const altb_record& r = buf[i];
            for (int c = 0; c < 3; c++) {
                if (last_pos) last_pos[3 * (off + i) + c] = r.pos[c];
                if (last_dir) last_dir[3 * (off + i) + c] = r.dir[c];
            }
            if (n_points) n_points[off + i] = 1 + r.n_hits + (r.status == ALTB_EXITED ? 1u : 0u);
            if (status) status[off + i] = (uint8_t)r.status;
        }
        tot.n_rays += s.n_rays; tot.n_exited += s.n_exited; tot.n_exit_port += s.n_exit_port; tot.n_absorbed += s.n_absorbed;
        tot.n_suspended += s.n_suspended; tot.n_bounces += s.n_bounces; tot.t_trace_s += s.t_trace_s;
    }
    if (stats) *stats = tot;
    return 0;
}

// ---------------------------------------------------------------------------------- map on host records
extern "C" int altb_map_records_at(altb_ctx* ctx, const altb_scene* scene, const altb_map_spec* map,
                                   const altb_record* records, uint64_t ray_id0, uint64_t n, uint64_t* counts);
extern "C" int altb_map_records(altb_ctx* ctx, const altb_scene* scene, const altb_map_spec* map,
                                const altb_record* records, uint64_t n, uint64_t* counts) {
    return altb_map_records_at(ctx, scene, map, records, 0, n, counts);
}
extern "C" int altb_map_records_at(altb_ctx* ctx, const altb_scene* scene, const altb_map_spec* map,
                                   const altb_record* records, uint64_t ray_id0, uint64_t n, uint64_t* counts) {
    if (!ctx || !scene || !map || (!records && n) || !counts) return fail(ALTB_E_ARG, "altb_map_records: NULL argument");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    Geom g; KConsts k;
    if (int rc = make_geom(scene, g, k)) return rc;
    MapSetup ms;
    if (int rc = setup_map(ctx, d, scene, g, k, map, ms, d.stream)) return rc;
    const uint64_t nb = (uint64_t)map->n_theta * map->n_phi;
    if (int rc = ensure(d.counts, d.counts_cap, nb + 8)) return rc;
    CK(cudaMemsetAsync(d.counts, 0, nb * sizeof(unsigned long long), d.stream));
    const uint64_t batch = std::min<uint64_t>(ctx->batch, std::max<uint64_t>(n, 1));
    if (int rc = ensure(d.rec, d.rec_cap, batch)) return rc;
    for (uint64_t off = 0; off < n; off += batch) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(batch, n - off);
        CK(cudaMemcpyAsync(d.rec, records + off, (size_t)m * sizeof(altb_record), cudaMemcpyHostToDevice, d.stream));
        if (int rc = run_map(ctx, d, ms, d.rec, d.counter, m, ray_id0 + off, d.counts, nullptr, nullptr, d.stream)) return rc;
    }
    std::vector<unsigned long long> host(nb);
    CK(cudaMemcpyAsync(host.data(), d.counts, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    for (uint64_t j = 0; j < nb; j++) counts[j] += host[j];
    return 0;
}

// ---------------------------------------------------------------------------------- physical disks
extern "C" int altb_detector_sweep(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                                   uint64_t n_rays, uint64_t seed, const double* det_center, const double* det_rot,
                                   uint32_t m, double det_r, double det_halfthick, uint64_t* hits, altb_stats* stats) {
    if (!ctx || !scene || !src || !det_center || !det_rot || !hits || m == 0) return fail(ALTB_E_ARG, "altb_detector_sweep: NULL/empty argument");
    if (m > 2048) return fail(ALTB_E_ARG, "altb_detector_sweep: at most 2048 poses per call (per-block shared-memory tables)");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    TraceSetup ts;
    if (int rc = setup_trace(scene, src, seed, ts)) return rc;
    const uint64_t batch = std::min<uint64_t>(ctx->batch, std::max<uint64_t>(n_rays, 1));
    if (int rc = ensure(d.rec, d.rec_cap, batch)) return rc;
    double* d_geo = nullptr;
    unsigned long long* d_hits = nullptr;
    CK(cudaMalloc(&d_geo, (size_t)m * 12 * sizeof(double)));
    if (cudaMalloc(&d_hits, (size_t)m * sizeof(unsigned long long)) != cudaSuccess) { cudaFree(d_geo); return fail(ALTB_E_NOMEM, "cudaMalloc hits"); }
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d_geo, det_center, (size_t)m * 3 * sizeof(double), cudaMemcpyHostToDevice, d.stream) != cudaSuccess ||
            cudaMemcpyAsync(d_geo + (size_t)m * 3, det_rot, (size_t)m * 9 * sizeof(double), cudaMemcpyHostToDevice, d.stream) != cudaSuccess ||
            cudaMemsetAsync(d_hits, 0, (size_t)m * sizeof(unsigned long long), d.stream) != cudaSuccess ||
            cudaMemsetAsync(d.stats, 0, 8 * sizeof(unsigned long long), d.stream) != cudaSuccess) { rc = fail(ALTB_E_CUDA, "detector_sweep: upload failed"); break; }
        for (uint64_t off = 0, n = 0; off < n_rays && !rc; off += n) {
            n = piece_len(ray_id0 + off, n_rays - off, batch);
            rc = run_trace(ctx, d, ts, SINK_RECORDS, ray_id0 + off, (uint32_t)n, d.rec, d.counter, d.rq[0], d.gstat[0], d.stream);
            if (rc) break;
            k_stats<<<d.sm_count * 4, 256, 0, d.stream>>>(d.rec, n, ts.P.g.count_all, ts.P.k.exit_zf, d.stats);
            int blocks = d.sm_count * 4;
            if (const char* e = getenv("ALTB_DISK_BLOCKS")) blocks = d.sm_count * std::max(1, atoi(e));      // tuning knob
            const int need = (int)((n + 7) / 8);
            if (blocks > need) blocks = need;
            k_disk_hits<<<blocks, DISK_THREADS, (size_t)m * sizeof(unsigned int), d.stream>>>(d.rec, n, ts.P.g, d_geo, d_geo + (size_t)m * 3, m, det_r,
                                                                                                   det_halfthick, d_hits);
            ctx->launches += 2;
            if (cudaGetLastError() != cudaSuccess) rc = fail(ALTB_E_CUDA, "detector_sweep: launch failed");
        }
        if (rc) break;
        std::vector<unsigned long long> h(m);
        unsigned long long hs[8];
        if (cudaMemcpyAsync(h.data(), d_hits, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream) != cudaSuccess ||
            cudaMemcpyAsync(hs, d.stats, sizeof hs, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess ||
            cudaStreamSynchronize(d.stream) != cudaSuccess) { rc = fail(ALTB_E_CUDA, "detector_sweep: %s", cudaGetErrorString(cudaGetLastError())); break; }
        for (uint32_t j = 0; j < m; j++) hits[j] += h[j];
        if (stats) read_stats(hs, stats);
    } while (0);
    cudaFree(d_geo); cudaFree(d_hits);
    return rc;
}

// ---------------------------------------------------------------------------------- replay
template <bool R, int M, int C, bool AZ>
static void launch_replay_t(const ReplayParams& P, const double* ray0, const float4* tape, const unsigned long long* off,
                            const uint32_t* order, altb_record* rec, cudaStream_t st) {
    k_replay<R, M, C, AZ><<<(P.n + 127) / 128, 128, 0, st>>>(P, ray0, tape, off, order, rec);
}
// replay instance = (roughness, model, contract, azimuth precision); the fast contract is built for models 0 and 1
template <int C, bool AZ>
static void launch_replay_c(bool rough, int model, const ReplayParams& P, const double* ray0, const float4* tape, const unsigned long long* off,
                            const uint32_t* order, altb_record* rec, cudaStream_t st) {
    if (rough) {
        if (model == 0) return launch_replay_t<true, 0, C, AZ>(P, ray0, tape, off, order, rec, st);
        if (model == 1) return launch_replay_t<true, 1, C, AZ>(P, ray0, tape, off, order, rec, st);
        if constexpr (C == CONTRACT_EXACT) {
            if (model == 2) return launch_replay_t<true, 2, C, AZ>(P, ray0, tape, off, order, rec, st);
            return launch_replay_t<true, 3, C, AZ>(P, ray0, tape, off, order, rec, st);
        }
    } else {
        if (model == 0) return launch_replay_t<false, 0, C, AZ>(P, ray0, tape, off, order, rec, st);
        if (model == 1) return launch_replay_t<false, 1, C, AZ>(P, ray0, tape, off, order, rec, st);
        if constexpr (C == CONTRACT_EXACT) {
            if (model == 2) return launch_replay_t<false, 2, C, AZ>(P, ray0, tape, off, order, rec, st);
            return launch_replay_t<false, 3, C, AZ>(P, ray0, tape, off, order, rec, st);
        }
    }
}

extern "C" int altb_replay_ex(altb_ctx* ctx, const altb_scene* scene, const double* ray0, const float* tape,
                              const uint64_t* tape_off, uint64_t n_rays, const altb_map_spec* map, uint32_t flags,
                              altb_record* records, int32_t* bin, uint8_t* port);
extern "C" int altb_replay(altb_ctx* ctx, const altb_scene* scene, const double* ray0, const float* tape,
                           const uint64_t* tape_off, uint64_t n_rays, const altb_map_spec* map,
                           altb_record* records, int32_t* bin, uint8_t* port) {
    return altb_replay_ex(ctx, scene, ray0, tape, tape_off, n_rays, map, 0u, records, bin, port);
}

extern "C" int altb_replay_ex(altb_ctx* ctx, const altb_scene* scene, const double* ray0, const float* tape,
                              const uint64_t* tape_off, uint64_t n_rays, const altb_map_spec* map, uint32_t flags,
                              altb_record* records, int32_t* bin, uint8_t* port) {
    if (flags & ~(uint32_t)ALTB_REPLAY_FULL_AZIMUTH) return fail(ALTB_E_ARG, "altb_replay_ex: unknown flags 0x%x", flags);
    if (!ctx || !scene || !ray0 || !tape_off || (!tape && n_rays && tape_off[n_rays])) return fail(ALTB_E_ARG, "altb_replay: NULL argument");
    if (n_rays == 0) return 0;
    if (n_rays > (1ull << 31)) return fail(ALTB_E_ARG, "altb_replay: at most 2^31 rays per call");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    ReplayParams P;
    if (int rc = make_geom(scene, P.g, P.k)) return rc;
    if (scene->brdf_kind == 3) return fail(ALTB_E_SCENE, "altb_replay: brdf_kind 3 (two rays per id) has no tape format");
    P.n = (uint32_t)n_rays; P.sincos = d.sincos;
    const bool rough = scene->roughness_rad != 0.0;
    const int model = !scene->lambertian ? 2 : (scene->brdf_kind == 1 ? 1 : (scene->brdf_kind == 2 ? 3 : 0));
    const uint64_t n_rec = tape_off[n_rays];
    if (int rc = ensure(d.rec, d.rec_cap, n_rays)) return rc;
    double* d_ray0 = nullptr; float* d_tape = nullptr; unsigned long long* d_off = nullptr; int* d_bin = nullptr;
    uint32_t* d_order = nullptr;
    // schedule: longest tape first (counting sort by record count), so that the rays of a warp finish together
    std::vector<uint32_t> order(n_rays);
    {
        uint64_t maxlen = 0;
        for (uint64_t i = 0; i < n_rays; i++) {
            if (tape_off[i + 1] < tape_off[i]) return fail(ALTB_E_ARG, "altb_replay: tape_off must be non-decreasing");
            maxlen = std::max<uint64_t>(maxlen, tape_off[i + 1] - tape_off[i]);
        }
        const uint64_t nbk = std::min<uint64_t>(maxlen, 1u << 20) + 1;          // lengths above 2^20 share the first bucket
        std::vector<uint64_t> start(nbk + 1, 0);
        auto bucket = [&](uint64_t i) { return nbk - 1 - std::min<uint64_t>(tape_off[i + 1] - tape_off[i], nbk - 1); };
        for (uint64_t i = 0; i < n_rays; i++) start[bucket(i) + 1]++;
        for (uint64_t b = 0; b < nbk; b++) start[b + 1] += start[b];
        for (uint64_t i = 0; i < n_rays; i++) order[start[bucket(i)]++] = (uint32_t)i;
    }
    int rc = 0;
    do {
        if (cudaMalloc(&d_order, n_rays * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemcpyAsync(d_order, order.data(), n_rays * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) {
            rc = fail(ALTB_E_NOMEM, "altb_replay: cudaMalloc failed"); break;
        }
        if (cudaMalloc(&d_ray0, n_rays * 6 * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&d_tape, std::max<uint64_t>(n_rec, 1) * 8 * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&d_off, (n_rays + 1) * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&d_bin, n_rays * sizeof(int)) != cudaSuccess) { rc = fail(ALTB_E_NOMEM, "altb_replay: cudaMalloc failed"); break; }
        if (cudaMemcpyAsync(d_ray0, ray0, n_rays * 6 * sizeof(double), cudaMemcpyHostToDevice, d.stream) != cudaSuccess ||
            (n_rec && cudaMemcpyAsync(d_tape, tape, n_rec * 8 * sizeof(float), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) ||
            cudaMemcpyAsync(d_off, tape_off, (n_rays + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, d.stream) != cudaSuccess) {
            rc = fail(ALTB_E_CUDA, "altb_replay: upload failed"); break;
        }
        const float4* t4 = reinterpret_cast<const float4*>(d_tape);
        cudaEventRecord(d.ev[0], d.stream);
        {
            const bool fast = ctx->contract != ALTB_CONTRACT_EXACT && model <= 1, az = (flags & ALTB_REPLAY_FULL_AZIMUTH) != 0;
            if (fast) { if (az) launch_replay_c<CONTRACT_FAST, true>(rough, model, P, d_ray0, t4, d_off, d_order, d.rec, d.stream);
                        else launch_replay_c<CONTRACT_FAST, false>(rough, model, P, d_ray0, t4, d_off, d_order, d.rec, d.stream); }
            else      { if (az) launch_replay_c<CONTRACT_EXACT, true>(rough, model, P, d_ray0, t4, d_off, d_order, d.rec, d.stream);
                        else launch_replay_c<CONTRACT_EXACT, false>(rough, model, P, d_ray0, t4, d_off, d_order, d.rec, d.stream); }
        }
        cudaEventRecord(d.ev[1], d.stream);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { rc = fail(ALTB_E_CUDA, "altb_replay: launch failed"); break; }
        if (getenv("ALTB_TIMING")) {       // replay is the one HBM-bound kernel (32 B of recorded draws per surface hit)
            float ms = 0.f;
            cudaEventSynchronize(d.ev[1]);
            cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
            fprintf(stderr, "[altb] k_replay: %llu rays, %llu tape records (%.3f GB), %.3f ms -> %.1f GB/s\n", (unsigned long long)n_rays,
                    (unsigned long long)n_rec, n_rec * 32e-9, ms, n_rec * 32e-9 / (ms * 1e-3));
        }
        std::vector<altb_record> tmp;
        altb_record* out = records;
        if (!out) { tmp.resize(n_rays); out = tmp.data(); }
        if (bin && map) {
            MapSetup ms;
            altb_map_spec m2 = *map; m2.map_mode = ALTB_MAP_DIRECTION;
            rc = setup_map(ctx, d, scene, P.g, P.k, &m2, ms, d.stream);
            if (rc) break;
            MapParams M = ms.M; M.use_smem_hist = 0;
            k_map_direction<<<d.sm_count * 2, DIR_THREADS, 0, d.stream>>>(d.rec, (uint32_t)n_rays, M, nullptr, nullptr, d_bin);
            ctx->launches++;
            if (cudaMemcpyAsync(bin, d_bin, n_rays * sizeof(int), cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) { rc = fail(ALTB_E_CUDA, "altb_replay: download failed"); break; }
        }
        if (cudaMemcpyAsync(out, d.rec, n_rays * sizeof(altb_record), cudaMemcpyDeviceToHost, d.stream) != cudaSuccess ||
            cudaStreamSynchronize(d.stream) != cudaSuccess) { rc = fail(ALTB_E_CUDA, "altb_replay: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (port)
            for (uint64_t i = 0; i < n_rays; i++)
                port[i] = (uint8_t)((P.g.count_all || out[i].status == ALTB_EXITED) && out[i].pos[2] < P.k.exit_zf);
    } while (0);
    cudaFree(d_ray0); cudaFree(d_tape); cudaFree(d_off); cudaFree(d_bin); cudaFree(d_order);
    return rc;
}

// ---------------------------------------------------------------------------------- RNG probe
extern "C" int altb_draws_lobe(altb_ctx* ctx, uint64_t seed, uint64_t ray_id0, uint64_t n, uint32_t k, int lobe_n, double lobe_deg, float* out);
extern "C" int altb_draws(altb_ctx* ctx, uint64_t seed, uint64_t ray_id0, uint64_t n, uint32_t k, float* out) {
    return altb_draws_lobe(ctx, seed, ray_id0, n, k, 0, 0.0, out);
}
extern "C" int altb_draws_lobe(altb_ctx* ctx, uint64_t seed, uint64_t ray_id0, uint64_t n, uint32_t k, int lobe_n, double lobe_deg, float* out) {
    if (!ctx || (!out && n)) return fail(ALTB_E_ARG, "altb_draws: NULL argument");
    if (n == 0) return 0;
    if (n > (1ull << 28)) return fail(ALTB_E_ARG, "altb_draws: n too large");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    float* buf = nullptr;
    CK(cudaMalloc(&buf, n * 8 * sizeof(float)));
    if (ctx->contract == ALTB_CONTRACT_FAST7)
        k_draws<CONTRACT_FAST7><<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(philox_expand(seed), d.sincos, ray_id0, (uint32_t)n, k, lobe_n, (float)(lobe_deg * PI_D / 180.0), buf);
    else if (ctx->contract == ALTB_CONTRACT_FAST)
        k_draws<CONTRACT_FAST><<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(philox_expand(seed), d.sincos, ray_id0, (uint32_t)n, k, lobe_n, (float)(lobe_deg * PI_D / 180.0), buf);
    else
        k_draws<CONTRACT_EXACT><<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(philox_expand(seed), d.sincos, ray_id0, (uint32_t)n, k, lobe_n, (float)(lobe_deg * PI_D / 180.0), buf);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, buf, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(ALTB_E_CUDA, "altb_draws: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------- math probe
extern "C" int altb_probe_f32(altb_ctx* ctx, int op, const float* x, uint64_t n, float* y) {
    if (!ctx || ((!x || !y) && n)) return fail(ALTB_E_ARG, "altb_probe_f32: NULL argument");
    if (op < 0 || op > 4) return fail(ALTB_E_ARG, "altb_probe_f32: op %d", op);
    if (n == 0) return 0;
    if (n > (1ull << 30)) return fail(ALTB_E_ARG, "altb_probe_f32: n too large");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    float *dx = nullptr, *dy = nullptr;
    if (cudaMalloc(&dx, n * sizeof(float)) != cudaSuccess || cudaMalloc(&dy, n * sizeof(float)) != cudaSuccess) {
        cudaFree(dx);
        return fail(ALTB_E_NOMEM, "altb_probe_f32: cudaMalloc failed");
    }
    cudaError_t e = cudaMemcpyAsync(dx, x, n * sizeof(float), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
        k_probe_f32<<<d.sm_count * 8, 256, 0, d.stream>>>(op, d.sincos, dx, (uint32_t)n, dy);
        ctx->launches++;
        e = cudaMemcpyAsync(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost, d.stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(dx); cudaFree(dy);
    if (e != cudaSuccess) return fail(ALTB_E_CUDA, "altb_probe_f32: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------- tilted normals past the horizon
extern "C" int altb_count_horizon(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0,
                                  uint64_t n_rays, uint64_t seed, uint64_t* n_events, uint64_t* n_rays_flagged, uint64_t* n_hits) {
    if (!ctx || !scene || !src) return fail(ALTB_E_ARG, "altb_count_horizon: NULL argument");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    TraceSetup ts;
    if (int rc = setup_trace(scene, src, seed, ts)) return rc;
    if (ts.rescatter) return fail(ALTB_E_SCENE, "altb_count_horizon: not defined for brdf_kind 3");
    unsigned long long h[3] = {0, 0, 0};
    if (ts.rough && n_rays) {                       // no roughness: no tilt, nothing to count but the hits (left at 0)
        CK(cudaMemsetAsync(d.stats, 0, 8 * sizeof(unsigned long long), d.stream));
        ts.P.sincos = d.sincos;
        for (uint64_t off = 0, n = 0; off < n_rays; off += n) {
            n = piece_len(ray_id0 + off, n_rays - off, 1ull << 30);
            ts.P.ray_id0 = ray_id0 + off; ts.P.n = (uint32_t)n;
            const unsigned blocks = (unsigned)((n + 127) / 128);
            if (ts.model == 0) k_horizon_count<0><<<blocks, 128, 0, d.stream>>>(ts.P, d.stats);
            else if (ts.model == 1) k_horizon_count<1><<<blocks, 128, 0, d.stream>>>(ts.P, d.stats);
            else if (ts.model == 2) k_horizon_count<2><<<blocks, 128, 0, d.stream>>>(ts.P, d.stats);
            else k_horizon_count<3><<<blocks, 128, 0, d.stream>>>(ts.P, d.stats);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(h, d.stats, sizeof h, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    if (n_events) *n_events = h[0];
    if (n_rays_flagged) *n_rays_flagged = h[1];
    if (n_hits) *n_hits = h[2];
    return 0;
}

// ---------------------------------------------------------------------------------- FP32 peak probe
extern "C" int altb_measure_fp32_peak(altb_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(ALTB_E_ARG, "altb_measure_fp32_peak: NULL argument");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    float* buf = nullptr;
    CK(cudaMalloc(&buf, 16));
    const int iters = 4096, blocks = d.sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(d.ev[0], d.stream));
        k_fma_peak<<<blocks, 256, 0, d.stream>>>(buf, iters);
        CK(cudaEventRecord(d.ev[1], d.stream));
        CK(cudaEventSynchronize(d.ev[1]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
        const double flop = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
        if (rep > 0 && ms > 0) best = std::max(best, flop / (ms * 1e-3) * 1e-12);
        ctx->launches++;
    }
    cudaFree(buf);
    *tflops = best;
    return 0;
}

// ---------------------------------------------------------------------------------- polylines (small N)
#include "altb_paths.cuh"

template <bool R, int M>
static void launch_paths_t(const TraceParams& P, f3 sp, uint32_t mp, float* pts, uint32_t* np_, uint8_t* st, cudaStream_t s) {
    k_trace_paths<R, M><<<(P.n + 127) / 128, 128, 0, s>>>(P, sp, mp, pts, np_, st);
}

extern "C" int altb_trace_paths(altb_ctx* ctx, const altb_scene* scene, const altb_source* src, uint64_t ray_id0, uint64_t n_rays,
                                uint64_t seed, uint32_t max_points, float* points, uint32_t* n_points, uint8_t* status) {
    if (!ctx || !scene || !src || !points || !n_points || max_points < 2) return fail(ALTB_E_ARG, "altb_trace_paths: NULL/empty argument");
    if (n_rays == 0) return 0;
    if (n_rays * (uint64_t)max_points > (1ull << 30)) return fail(ALTB_E_ARG, "altb_trace_paths: n_rays * max_points too large (polylines are for small N)");
    DevCtx& d = ctx->devs[0];
    CK(cudaSetDevice(d.dev));
    TraceSetup ts;
    if (int rc = setup_trace(scene, src, seed, ts)) return rc;
    if (ts.rescatter) return fail(ALTB_E_SCENE, "altb_trace_paths: a polyline is ONE ray; brdf_kind 3 traces two per id");
    ts.P.ray_id0 = ray_id0; ts.P.n = (uint32_t)n_rays; ts.P.chunk = 0; ts.P.sincos = d.sincos;
    float* d_pts = nullptr; uint32_t* d_np = nullptr; uint8_t* d_st = nullptr;
    const size_t nb = (size_t)n_rays * max_points * 3 * sizeof(float);
    int rc = 0;
    do {
        if (cudaMalloc(&d_pts, nb) != cudaSuccess || cudaMalloc(&d_np, n_rays * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc(&d_st, n_rays) != cudaSuccess) { rc = fail(ALTB_E_NOMEM, "altb_trace_paths: cudaMalloc failed"); break; }
        cudaMemsetAsync(d_pts, 0, nb, d.stream);
        const f3 sp = {(float)src->pos[0], (float)src->pos[1], (float)src->pos[2]};
        const bool R = ts.rough; const int M = ts.model;
#define GO(RR, MM) launch_paths_t<RR, MM>(ts.P, sp, max_points, d_pts, d_np, d_st, d.stream)
        if (R) { if (M == 0) GO(true, 0); else if (M == 1) GO(true, 1); else if (M == 2) GO(true, 2); else GO(true, 3); }
        else   { if (M == 0) GO(false, 0); else if (M == 1) GO(false, 1); else if (M == 2) GO(false, 2); else GO(false, 3); }
#undef GO
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { rc = fail(ALTB_E_CUDA, "altb_trace_paths: launch failed"); break; }
        if (cudaMemcpyAsync(points, d_pts, nb, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess ||
            cudaMemcpyAsync(n_points, d_np, n_rays * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.stream) != cudaSuccess ||
            (status && cudaMemcpyAsync(status, d_st, n_rays, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) ||
            cudaStreamSynchronize(d.stream) != cudaSuccess) { rc = fail(ALTB_E_CUDA, "altb_trace_paths: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    cudaFree(d_pts); cudaFree(d_np); cudaFree(d_st);
    return rc;
}
