// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED except via demo/*.png.
//
// CPU restatement of the reference integrator: src/camera.rs:51-228.
#pragma once
#include "oracle_bsdf.hpp"

namespace orc {

// Environment importance sampler — NOT in the reference (PT_RENDER_ENV_IMPORTANCE, include/pt_b200.h).
// The oracle restates the product's definition so device and CPU can be compared sample for sample: a
// piecewise-constant density over rows x cols cells of the lat-long map, weight = sum over the cell's texels of
// luminance * sin(theta of the texel row) + a uniform floor of 5 % of the mean weight; sequential f64 sums.
struct EnvDist {
    std::vector<double> marginal, cond;  // CDFs: [rows + 1], [rows][cols + 1]
    uint32_t rows = 0, cols = 0;
    void build(const ImageData& img, uint32_t max_rows, uint32_t max_cols) {
        const uint32_t W = img.width, H = img.height;
        rows = std::min<uint32_t>(H, max_rows ? max_rows : 512u); cols = std::min<uint32_t>(W, max_cols ? max_cols : 1024u);
        std::vector<double> w((size_t)rows * cols, 0.0);
        for (uint32_t j = 0; j < H; j++) {
            double st = std::sin(((double)j + 0.5) / (double)H * PI);
            size_t r = (size_t)((uint64_t)j * rows / H);
            for (uint32_t i = 0; i < W; i++) {
                const uint8_t* p = img.rgb + ((size_t)j * W + i) * 3;
                double lum = luminance(Vec3((double)p[0] / 255.0, (double)p[1] / 255.0, (double)p[2] / 255.0));
                w[r * cols + (size_t)((uint64_t)i * cols / W)] += lum * st;
            }
        }
        double total = 0.0;
        for (double v : w) total += v;
        double floor_w = total > 0.0 ? 0.05 * total / ((double)rows * (double)cols) : 1.0;
        marginal.assign(rows + 1, 0.0); cond.assign((size_t)rows * (cols + 1), 0.0);
        std::vector<double> row_sum(rows);
        double all = 0.0;
        for (uint32_t r = 0; r < rows; r++) {
            double rs = 0.0;
            for (uint32_t c = 0; c < cols; c++) { w[(size_t)r * cols + c] += floor_w; rs += w[(size_t)r * cols + c]; }
            row_sum[r] = rs; all += rs;
            double* cr = cond.data() + (size_t)r * (cols + 1);
            for (uint32_t c = 0; c < cols; c++) cr[c + 1] = cr[c] + w[(size_t)r * cols + c] / rs;
            cr[cols] = 1.0;
        }
        for (uint32_t r = 0; r < rows; r++) marginal[r + 1] = marginal[r] + row_sum[r] / all;
        marginal[rows] = 1.0;
    }
    static uint32_t find(const double* cdf, uint32_t n, double u) {  // largest i in [0, n) with cdf[i] <= u
        uint32_t lo = 0, hi = n;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid; else hi = mid; }
        return lo;
    }
    Vec3 sample(double u1, double u2) const {
        uint32_t r = find(marginal.data(), rows, u1);
        double m0 = marginal[r], m1 = marginal[r + 1];
        const double* row = cond.data() + (size_t)r * (cols + 1);
        uint32_t c = find(row, cols, u2);
        double c0 = row[c], c1 = row[c + 1];
        double fr = m1 > m0 ? (u1 - m0) / (m1 - m0) : 0.5, fc = c1 > c0 ? (u2 - c0) / (c1 - c0) : 0.5;
        double theta = (((double)r + fr) / (double)rows) * PI;
        double phi = (((double)c + fc) / (double)cols) * (2.0 * PI) - PI;
        double st = std::sin(theta);
        return Vec3(st * std::cos(phi), std::cos(theta), st * std::sin(phi));
    }
    double pdf(Vec3 dir) const {
        double theta = std::acos(dir.y), phi = std::atan2(dir.z, dir.x);
        double st = std::sin(theta);
        if (!(st > 0.0)) return 0.0;
        double fu = (phi + PI) / (2.0 * PI) * (double)cols, fv = theta / PI * (double)rows;
        uint32_t c = fu > 0.0 ? (uint32_t)fu : 0u, r = fv > 0.0 ? (uint32_t)fv : 0u;
        if (c >= cols) c = cols - 1;
        if (r >= rows) r = rows - 1;
        const double* row = cond.data() + (size_t)r * (cols + 1);
        double cell = (marginal[r + 1] - marginal[r]) * (row[c + 1] - row[c]);
        return cell * (double)rows * (double)cols / (2.0 * PI * PI * st);
    }
};

struct Camera {
    const EnvDist* env_dist = nullptr;  // non-null: PT_RENDER_ENV_IMPORTANCE (ours, not the reference's)
    // public fields, camera.rs:23-36
    double aspect_ratio = 1.0; uint32_t image_width = 0, samples_per_pixel = 0, max_depth = 0;
    double vfov = 0; Vec3 look_from, look_at, vup;
    double blur_strength = 0, focal_length = 0, defocus_angle = 0;
    bool env_is_map = false; Vec3 env_color; const ImageTexture* env_map = nullptr;
    // derived, camera.rs:38-47
    Vec3 forward, right, up; uint32_t image_height = 0; double pixel_sample_scale = 0;
    Vec3 center, pixel00, pixel_du, pixel_dv;

    void init() {  // camera.rs:51-77
        image_height = (uint32_t)((double)image_width / aspect_ratio);
        pixel_sample_scale = 1.0 / (double)samples_per_pixel;
        center = look_from;
        double theta = to_radians(vfov);
        double h = std::tan(theta / 2.0);
        double viewport_height = 2.0 * h * focal_length;
        double viewport_width = viewport_height * ((double)image_width / (double)image_height);
        forward = normalize(look_from - look_at);
        right = normalize(cross(vup, forward));
        up = cross(forward, right);
        Vec3 viewport_u = right * viewport_width;
        Vec3 viewport_v = up * -viewport_height;
        pixel_du = viewport_u / (double)image_width;
        pixel_dv = viewport_v / (double)image_height;
        Vec3 upperleft = center - (forward * focal_length) - (viewport_u / 2.0) - (viewport_v / 2.0);
        pixel00 = upperleft + (pixel_du + pixel_dv) * 0.5;
    }
    static void random_offsets(Rng& rng, double& x, double& y) {  // camera.rs:133-138
        double radius = std::sqrt(rng.next());
        double angle = rng.next() * 2.0 * PI;
        x = radius * std::cos(angle); y = radius * std::sin(angle);
    }
    Vec3 sample_environment(const Ray& ray) const {  // camera.rs:140-151
        if (!env_is_map) return env_color;
        double theta = std::acos(ray.direction.y);
        double phi = std::atan2(ray.direction.z, ray.direction.x);
        double u = (phi + PI) / (2.0 * PI);
        double v = 1.0 - theta / PI;
        return env_map->value(u, v, Vec3(0, 0, 0));
    }
    Ray generate_ray(uint32_t r, uint32_t c, Rng& rng) const {  // camera.rs:153-168
        double bx, by; random_offsets(rng, bx, by);
        bx = bx * blur_strength; by = by * blur_strength;
        Vec3 sample_location = pixel00 + (pixel_dv * ((double)r + bx)) + (pixel_du * ((double)c + by));
        double radius = std::tan(to_radians(defocus_angle / 2.0)) * focal_length;
        Vec3 dof_right = right * radius, dof_up = up * radius;
        double px, py; random_offsets(rng, px, py);
        Vec3 origin = center + (dof_right * px) + (dof_up * py);
        Vec3 direction = sample_location - origin;
        double time = rng.next();
        return Ray::make(origin, direction, time);
    }
    // camera.rs:170-228.  `record`, when non-null, receives every ray handed to intersect_all.
    // `dropped`, when non-null, selects PT_NAN_DROP (include/pt_b200.h; NOT reference behaviour): a non-finite
    // contribution is skipped and a non-finite throughput ends the path, each counted once in *dropped; finite
    // contributions the sample made before that stay.  With dropped == nullptr this is the reference's loop verbatim.
    Vec3 trace(uint32_t r, uint32_t c, const World& world, Rng& rng, std::vector<Ray>* record = nullptr, uint64_t* dropped = nullptr) const {
        auto fin3 = [](const Vec3& v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); };
        const double eps = 1e-3;
        const uint32_t min_bounces = 5;
        Vec3 radiance(0, 0, 0), throughput(1, 1, 1);
        Ray ray = generate_ray(r, c, rng);
        g_cnt.paths++;
        for (uint32_t bounces = 0; bounces < max_depth; bounces++) {
            if (record) record->push_back(ray);
            g_path_key = PathKey{rng.seed, rng.pixel, rng.sample, bounces};  // only volumes (ours) read it
            auto hit = world.intersect_all(ray, Interval{eps, INF});
            if (!hit) {
                Vec3 env = throughput * sample_environment(ray);
                if (dropped && !fin3(env)) { ++*dropped; break; }
                radiance += env;
                break;
            }
            const HitInfo& info = hit->first;
            Vec3 emission = info.mat->emitted(info.u, info.v, info.point);
            if (dropped && !fin3(throughput * emission)) { ++*dropped; break; }
            radiance += throughput * emission;
            if (bounces > min_bounces) {  // Russian roulette, camera.rs:190-196
                double p = clamp_(luminance(throughput), 0.01, 1.0);
                if (rng.next() > p) break;
                throughput /= p;
            }
            double p_light = world.lights.is_empty() ? 0.0 : 0.5;  // camera.rs:199
            double p_bsdf = 1.0 - p_light;
            double p_env = 0.0;
            if (env_dist) {  // ours (PT_RENDER_ENV_IMPORTANCE): the environment map joins the mixture as a third sampler
                p_env = world.lights.is_empty() ? 0.5 : 0.25;
                p_light = world.lights.is_empty() ? 0.0 : 0.5 - p_env;
                p_bsdf = 0.5;
            }
            double rr = rng.next();
            std::optional<Vec3> dir;
            if (rr < p_light) dir = world.lights.sample(info.point, ray.time, rng);
            else if (env_dist && rr < p_light + p_env) { double u1 = rng.next(), u2 = rng.next(); dir = env_dist->sample(u1, u2); }
            else dir = info.mat->sample(ray, info, rng);
            if (!dir) break;
            double bsdf_pdf = info.mat->pdf(-ray.direction, *dir, info);
            double light_pdf = world.lights.pdf(info.point, *dir, ray.time);
            double pdf = p_bsdf * bsdf_pdf + p_light * light_pdf;
            if (env_dist) pdf = pdf + p_env * env_dist->pdf(*dir);
            Vec3 brdf = info.mat->eval(-ray.direction, *dir, info);
            Vec3 attenuation = brdf / pdf;
            double e = 1e-3 * signum_(dot(*dir, info.geometric_normal));  // bsdf::EPS, camera.rs:217
            Ray next = Ray::make(info.point + e * info.geometric_normal, *dir, ray.time);
            throughput *= attenuation;
            if (dropped && !fin3(throughput)) { ++*dropped; break; }
            ray = next;
        }
        return radiance;
    }

    // PT_RENDER_NEE (include/pt_b200.h) — NOT the reference's integrator: next-event estimation with the balance heuristic.
    // Same hit handling, Russian roulette, offsets and sampling routines as trace(); the difference is how directions are
    // chosen and weighted.  The device spawns the shadow ray as a child path and resolves it one iteration later; here it
    // is resolved on the spot, which adds the same terms.
    Vec3 trace_nee(uint32_t r, uint32_t c, const World& world, Rng& rng, uint64_t* dropped = nullptr) const {
        auto fin3 = [](const Vec3& v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); };
        auto add = [&](Vec3& radiance, const Vec3& v) {  // false: the contribution was dropped (PT_NAN_DROP)
            if (!fin3(v)) { if (dropped) { ++*dropped; return false; } }
            radiance += v;
            return true;
        };
        const double eps = 1e-3;
        Vec3 radiance(0, 0, 0), throughput(1, 1, 1);
        Ray ray = generate_ray(r, c, rng);
        g_cnt.paths++;
        double w_prev = 1.0;  // MIS weight of the BSDF strategy for the direction that produced `ray`
        for (uint32_t bounces = 0; bounces < max_depth; bounces++) {
            g_path_key = PathKey{rng.seed, rng.pixel, rng.sample, bounces};
            auto hit = world.intersect_all(ray, Interval{eps, INF});
            if (!hit) { add(radiance, throughput * sample_environment(ray)); break; }
            const HitInfo& info = hit->first;
            if (info.mat->is_emitter()) {  // emitters end the path (DiffuseLight::sample -> None)
                add(radiance, throughput * info.mat->emitted(info.u, info.v, info.point) * (bounces == 0 ? 1.0 : w_prev));
                break;
            }
            if (!fin3(throughput) && !add(radiance, throughput * 0.0)) break;
            if (bounces > 5) {
                double p = clamp_(luminance(throughput), 0.01, 1.0);
                if (rng.next() > p) break;
                throughput /= p;
            }
            const bool deeper = bounces + 1 < max_depth;
            if (!world.lights.is_empty()) {
                if (auto ld = world.lights.sample(info.point, ray.time, rng)) {
                    double pl = world.lights.pdf(info.point, *ld, ray.time);
                    double pb = info.mat->pdf(-ray.direction, *ld, info);
                    Vec3 fl = info.mat->eval(-ray.direction, *ld, info);
                    Vec3 W = throughput * (fl / (pl + pb));
                    if (deeper && pl > 0.0 && (W.x != 0.0 || W.y != 0.0 || W.z != 0.0)) {
                        double e = 1e-3 * signum_(dot(*ld, info.geometric_normal));
                        Ray shadow = Ray::make(info.point + e * info.geometric_normal, *ld, ray.time);
                        g_path_key.bounce = bounces + 1;
                        auto sh = world.intersect_all(shadow, Interval{eps, INF});
                        if (sh && sh->first.mat->is_emitter()) add(radiance, W * sh->first.mat->emitted(sh->first.u, sh->first.v, sh->first.point));
                    }
                }
            }
            auto dir = info.mat->sample(ray, info, rng);
            if (!dir) break;
            double pb = info.mat->pdf(-ray.direction, *dir, info);
            Vec3 f = info.mat->eval(-ray.direction, *dir, info);
            double pl = world.lights.is_empty() ? 0.0 : world.lights.pdf(info.point, *dir, ray.time);
            w_prev = pl > 0.0 ? (double)(float)(pb / (pb + pl)) : 1.0;  // carried as fp32 by the device
            double e = 1e-3 * signum_(dot(*dir, info.geometric_normal));
            Ray next = Ray::make(info.point + e * info.geometric_normal, *dir, ray.time);
            throughput *= f / pb;
            if (dropped && !fin3(throughput)) { ++*dropped; break; }
            if (!deeper) break;
            ray = next;
        }
        return radiance;
    }
};

}  // namespace orc
