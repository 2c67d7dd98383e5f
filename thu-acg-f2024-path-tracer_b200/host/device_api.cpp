#include "device_api.hpp"

#include <dlfcn.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace pt {
namespace {
void anchor() {}
template <class F> void bind(void* lib, const char* name, F& out) {
    void* p = dlsym(lib, name);
    if (!p) throw std::runtime_error(std::string("libptb200.so does not export ") + name);
    out = reinterpret_cast<F>(p);
}
}  // namespace

const DeviceApi& device_api() {
    static DeviceApi api{};
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        try {
            // next to this library (lib/libptb200_host.so -> lib/libptb200.so); a copy already mapped by the process is reused
            std::string path = "libptb200.so";
            Dl_info info{};
            if (dladdr(reinterpret_cast<void*>(&anchor), &info) && info.dli_fname) {
                const std::string self = info.dli_fname;
                const size_t slash = self.rfind('/');
                if (slash != std::string::npos) path = self.substr(0, slash + 1) + "libptb200.so";
            }
            void* lib = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
            if (!lib) throw std::runtime_error(std::string("cannot load the CUDA library: ") + dlerror());
            bind(lib, "pt_ctx_create", api.ctx_create); bind(lib, "pt_ctx_destroy", api.ctx_destroy); bind(lib, "pt_last_error", api.last_error);
            bind(lib, "pt_scene_create", api.scene_create); bind(lib, "pt_scene_destroy", api.scene_destroy);
            bind(lib, "pt_scene_build_env_sampler", api.scene_build_env_sampler); bind(lib, "pt_render", api.render);
            bind(lib, "pt_render_multi", api.render_multi); bind(lib, "pt_sah_sweep", api.sah_sweep);
        } catch (const std::exception& e) { error = e.what(); }
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}
}  // namespace pt
