// mppi_rtc.cu — user-supplied MPPI models: the reference's Mppi::new takes `dynamics` and `cost` as fn pointers
// (src/mppi.rs:9-10,16-22); a GPU library cannot call host functions from a kernel, so the same two functions are handed
// over as CUDA source and the SAME fused kernel (mppi_kernel.cuh, embedded in this library as text) is compiled around
// them with NVRTC for sm_100a.  libnvrtc is opened on first use (dlopen), so the library itself does not depend on it.
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mppi_rtc.h"
#include "rtc_headers_gen.inc"

namespace mpcb {
namespace {

thread_local std::string g_log;

// compiled modules of this process, keyed by everything that goes into the compilation: a second handle on the same
// model (another rank of a sharded controller, another device, a re-created filter) skips the ~2.5 s NVRTC run
struct CachedCubin {
    std::vector<char> cubin;
    std::string lowered[3];
};
std::mutex g_cache_mutex;
std::map<std::string, CachedCubin> g_cache;

struct NvrtcApi {
    void* lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
};

// cudaLibrary* entry points of the CUDA 12 runtime, looked up at run time so that an older libcudart.so.12 only
// disables user models instead of failing to load the whole library
struct LibraryApi {
    cudaError_t (*LoadData)(cudaLibrary_t*, const void*, enum cudaJitOption*, void**, unsigned int, enum cudaLibraryOption*, void**,
                            unsigned int) = nullptr;
    cudaError_t (*GetKernel)(cudaKernel_t*, cudaLibrary_t, const char*) = nullptr;
    cudaError_t (*Unload)(cudaLibrary_t) = nullptr;
};

template <typename F>
bool sym(void* lib, const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(lib, name));
    return *out != nullptr;
}

// both tables are filled once, on first use (function-local statics: thread-safe initialisation)
NvrtcApi load_nvrtc() {
    NvrtcApi api;
    const char* override_path = getenv("MPCB_NVRTC_PATH");
    const char* names[] = {override_path, "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (n && (lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    }
    if (!lib) return api;
    const bool ok = sym(lib, "nvrtcCreateProgram", &api.CreateProgram) && sym(lib, "nvrtcDestroyProgram", &api.DestroyProgram) &&
                    sym(lib, "nvrtcCompileProgram", &api.CompileProgram) && sym(lib, "nvrtcGetProgramLogSize", &api.GetProgramLogSize) &&
                    sym(lib, "nvrtcGetProgramLog", &api.GetProgramLog) && sym(lib, "nvrtcAddNameExpression", &api.AddNameExpression) &&
                    sym(lib, "nvrtcGetLoweredName", &api.GetLoweredName) && sym(lib, "nvrtcGetCUBINSize", &api.GetCUBINSize) &&
                    sym(lib, "nvrtcGetCUBIN", &api.GetCUBIN) && sym(lib, "nvrtcGetErrorString", &api.GetErrorString);
    if (!ok) {
        dlclose(lib);
        return api;
    }
    api.lib = lib;
    return api;
}
NvrtcApi* nvrtc_api() {
    static NvrtcApi api = load_nvrtc();
    return api.lib ? &api : nullptr;
}

LibraryApi load_library_api() {
    LibraryApi api;
    Dl_info info;
    void* rt = nullptr;
    if (dladdr(reinterpret_cast<void*>(static_cast<cudaError_t (*)(int*)>(&cudaGetDeviceCount)), &info) && info.dli_fname)
        rt = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
    if (!rt) rt = dlopen("libcudart.so.12", RTLD_NOW | RTLD_NOLOAD);
    if (!rt) rt = RTLD_DEFAULT;
    if (!(sym(rt, "cudaLibraryLoadData", &api.LoadData) && sym(rt, "cudaLibraryGetKernel", &api.GetKernel) &&
          sym(rt, "cudaLibraryUnload", &api.Unload)))
        api = LibraryApi();
    return api;
}
LibraryApi* library_api() {
    static LibraryApi api = load_library_api();
    return api.LoadData ? &api : nullptr;
}

// The adapter between the user's two functions and the model interface of mppi_kernel.cuh (load / step / cost).
// It follows the user source, so that `::dynamics` and `::cost` are declared when the templates are parsed.
const char* kAdapter = R"RTC(
namespace mpcb {
template <typename real> __device__ __forceinline__ const real* user_params(const ModelConsts& mc);
template <> __device__ __forceinline__ const float* user_params<float>(const ModelConsts& mc) { return mc.kf; }
template <> __device__ __forceinline__ const double* user_params<double>(const ModelConsts& mc) { return mc.k; }
template <typename real>
struct UserCost {
    const real* p;  // the caller's parameters, read from the kernel-parameter (constant) bank
    __device__ __forceinline__ real operator()(const real (&x)[MPCB_USER_STATE_DIM]) const { return ::cost(x, p); }
    __device__ __forceinline__ real acc(const real (&x)[MPCB_USER_STATE_DIM], real s) const { return s + ::cost(x, p); }
};
template <typename real>
struct ModelUser {
    static constexpr int kStateDim = MPCB_USER_STATE_DIM;  // S of Mppi<N,K,S>
    UserCost<real> cost;
    __device__ __forceinline__ void load(const ModelConsts& mc) { cost.p = user_params<real>(mc); }
    __device__ __forceinline__ void step(real (&x)[MPCB_USER_STATE_DIM], real u) const { ::dynamics(x, u, cost.p); }
};
}  // namespace mpcb
)RTC";

}  // namespace

const char* rtc_log() { return g_log.c_str(); }

void rtc_unload(RtcModule* m) {
    if (m && m->library) {
        LibraryApi* la = library_api();
        if (la) la->Unload(reinterpret_cast<cudaLibrary_t>(m->library));
        m->library = nullptr;
    }
}

namespace {

// compiles `src` for sm_100a, resolves the three kernels named by `names` and (load) loads the cubin on the current device
mpcb_status load_cubin(const CachedCubin& c, const std::string (&names)[3], RtcModule* out) {
    LibraryApi* la = library_api();
    if (!la) {
        set_error("user models need cudaLibraryLoadData (CUDA runtime >= 12.0 with library management)");
        return MPCB_RTC_ERROR;
    }
    cudaLibrary_t lib = nullptr;
    cudaError_t e = la->LoadData(&lib, c.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        set_error("cudaLibraryLoadData failed: %s", cudaGetErrorString(e));
        return MPCB_CUDA_ERROR;
    }
    out->library = lib;
    for (int i = 0; i < 3; ++i) {
        cudaKernel_t k = nullptr;
        if (la->GetKernel(&k, lib, c.lowered[i].c_str()) != cudaSuccess) {
            set_error("kernel %s not found in the compiled user module", names[i].c_str());
            rtc_unload(out);
            return MPCB_RTC_ERROR;
        }
        out->kernel[i] = k;
    }
    return MPCB_OK;
}

// compiles `src` for sm_100a, resolves the three kernels named by `names` and (load) loads the cubin on the current device
mpcb_status compile_and_load(const std::string& src, const std::string (&names)[3], std::vector<const char*> opts, bool load, RtcModule* out) {
    g_log.clear();
    std::string key = src;
    for (const char* o : opts) { key += '\n'; key += o; }
    for (const std::string& n : names) { key += '\n'; key += n; }
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) return load ? load_cubin(it->second, names, out) : MPCB_OK;
    }
    NvrtcApi* nv = nvrtc_api();
    if (!nv) {
        set_error("user models need NVRTC: libnvrtc.so.12 could not be opened (%s); set MPCB_NVRTC_PATH", dlerror());
        return MPCB_RTC_ERROR;
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = nv->CreateProgram(&prog, src.c_str(), "mpcb_user_model.cu", kRtcHeaderCount, kRtcHeaderSources, kRtcHeaderNames);
    if (r != NVRTC_SUCCESS) {
        set_error("nvrtcCreateProgram: %s", nv->GetErrorString(r));
        return MPCB_RTC_ERROR;
    }
    for (int i = 0; i < 3; ++i) nv->AddNameExpression(prog, names[i].c_str());
    opts.insert(opts.begin(), {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"});
    r = nv->CompileProgram(prog, (int)opts.size(), opts.data());
    size_t n = 0;
    if (nv->GetProgramLogSize(prog, &n) == NVRTC_SUCCESS && n > 1) {
        g_log.resize(n);
        nv->GetProgramLog(prog, &g_log[0]);
        while (!g_log.empty() && g_log.back() == '\0') g_log.pop_back();
    }
    if (r != NVRTC_SUCCESS) {
        set_error("user model does not compile (%s); see mpcb_rtc_log()", nv->GetErrorString(r));
        nv->DestroyProgram(&prog);
        return MPCB_RTC_ERROR;
    }
    CachedCubin c;
    size_t cn = 0;
    mpcb_status st = MPCB_OK;
    if (nv->GetCUBINSize(prog, &cn) != NVRTC_SUCCESS || cn == 0) {
        set_error("nvrtcGetCUBIN returned nothing");
        st = MPCB_RTC_ERROR;
    } else {
        c.cubin.resize(cn);
        nv->GetCUBIN(prog, c.cubin.data());
        for (int i = 0; i < 3 && st == MPCB_OK; ++i) {
            const char* low = nullptr;
            if (nv->GetLoweredName(prog, names[i].c_str(), &low) != NVRTC_SUCCESS || !low) {
                set_error("kernel %s not found in the compiled user module", names[i].c_str());
                st = MPCB_RTC_ERROR;
            } else {
                c.lowered[i] = low;
            }
        }
    }
    nv->DestroyProgram(&prog);
    if (st != MPCB_OK) return st;
    if (load) st = load_cubin(c, names, out);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        g_cache.emplace(std::move(key), std::move(c));
    }
    return st;
}

}  // namespace

mpcb_status rtc_compile_mppi_user(const char* user_src, int state_dim, bool f64, int block, bool vt, bool load, RtcModule* out) {
    std::string src = "#define MPCB_USER_STATE_DIM " + std::to_string(state_dim) + "\n#include \"mppi_kernel.cuh\"\n#line 1 \"user_model.cu\"\n";
    src += user_src;
    src += "\n#line 1 \"mpcb_user_adapter.cu\"\n";
    src += kAdapter;
    std::string names[3];
    for (int noise = 0; noise < 3; ++noise) {
        char buf[160];
        snprintf(buf, sizeof(buf), "mpcb::mppi_rollout_kernel<mpcb::ModelUser, %s, %d, %d, 1, %s>", f64 ? "double" : "float", block, noise,
                 vt ? "true" : "false");
        names[noise] = buf;
    }
    // the FP64 path is reference arithmetic: no FMA contraction (like the -fmad=false of the built-in FP64 kernels)
    std::vector<const char*> opts;
    if (f64) opts.push_back("-fmad=false");
    return compile_and_load(src, names, opts, load, out);
}

mpcb_status rtc_compile_ukf_user(const char* user_src, int n, int o, int sqrt_mode, int order, bool fast, bool load, RtcModule* out) {
    // helpers first, then the user's fx / hx, then the kernel that calls them
    std::string src = "#include \"models.cuh\"\n#line 1 \"user_model.cu\"\n";
    src += user_src;
    src += "\n#line 1 \"mpcb_user_adapter.cu\"\n#define MPCB_UKF_USER 1\n#include \"ukf_kernel.cuh\"\n";
    std::string names[3];
    for (int mode = 0; mode < 3; ++mode) {
        char buf[160];
        snprintf(buf, sizeof(buf), "mpcb::ukf_kernel<%d, %d, %d, %d, %d, %d, %s>", n, o, (int)MPCB_MODEL_USER_UKF, sqrt_mode, order, mode,
                 fast ? "true" : "false");
        names[mode] = buf;
    }
    std::vector<const char*> opts;
    if (!fast) opts.push_back("-fmad=false");  // cfg.exact: the reference's operation order without FMA
    return compile_and_load(src, names, opts, load, out);
}

}  // namespace mpcb
