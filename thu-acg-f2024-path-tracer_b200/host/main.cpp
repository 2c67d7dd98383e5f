// ptb200 — command line mirroring the reference binary (src/main.rs:620-645):
//   -s/--scene N   scene 1..7 (default 1; 70 = our mesh variant of scene 7)
//   -q/--quality   1920 px x 4000 spp instead of 600 px x 100 spp
// plus knobs the reference lacks: --spp, --width, --seed, --device, --gpus, --assets, --out, --drop-nonfinite
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pt_host.hpp"

int main(int argc, char** argv) {
    int scene = 1; bool quality = false; long spp = -1, width = -1; unsigned long long seed = 1; int device = 0, gpus = 1;
    std::string assets = "assets", outdir = "demo"; bool drop = false, env_is = false, nee = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-s" || a == "--scene") scene = atoi(next());
        else if (a == "-q" || a == "--quality") quality = true;
        else if (a == "--spp") spp = atol(next());
        else if (a == "--width") width = atol(next());
        else if (a == "--seed") seed = strtoull(next(), nullptr, 10);
        else if (a == "--device") device = atoi(next());
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--assets") assets = next();
        else if (a == "--out") outdir = next();
        else if (a == "--drop-nonfinite") drop = true;
        else if (a == "--env-importance") env_is = true;
        else if (a == "--nee") nee = true;
        else { fprintf(stderr, "usage: ptb200 [-s N] [-q] [--spp N] [--width N] [--seed N] [--device N] [--gpus N] [--assets DIR] [--out DIR] [--drop-nonfinite] [--env-importance] [--nee]\n"); return 2; }
    }
    uint32_t w = quality ? 1920 : 600, s = quality ? 4000 : 100;  // main.rs:633
    if (width > 0) w = (uint32_t)width;
    if (spp > 0) s = (uint32_t)spp;
    try {
        auto b = pt::build_scene(scene, w, s, seed, assets);
        pt::RenderOptions o; o.seed = seed; o.device = device; o.nan_policy = drop ? PT_NAN_DROP : PT_NAN_REFERENCE; o.env_importance = env_is; o.nee = nee; o.gpus = gpus;
        return b->camera.render(b->world, outdir + "/" + b->output_name, o) ? 1 : 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
