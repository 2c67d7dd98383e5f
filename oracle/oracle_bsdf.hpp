// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).  PARITY UNPINNED except via demo/*.png.
//
// CPU restatement of the reference's shading layer:
//   src/bsdf/{mod,sampling,diffuse,metal,glass,principled,sheen,clearcoat,mix}.rs, src/material.rs:150-191
// Random variates come from orc::Rng in the reference's program order.
#pragma once
#include "oracle_scene.hpp"

namespace orc {

// ---------------------------------------------------------------- bsdf/mod.rs:61-97
inline Vec3 tint(Vec3 base_color) {  // mod.rs:61-68
    if (luminance(base_color) > 0.0) return base_color / luminance(base_color);
    return Vec3(1, 1, 1);
}
inline double r0(double eta) { return powi2((eta - 1.0) / (eta + 1.0)); }  // mod.rs:70-72
namespace fresnel {
inline double dielectric(Vec3 w, Vec3 h, double eta_i, double eta_o) {  // mod.rs:77-88 (== glass.rs:51-62)
    double c = std::fabs(dot(w, h));
    double g_squared = powi2(eta_o / eta_i) - 1.0 + c * c;
    if (g_squared < 0.0) return 1.0;
    double g = std::sqrt(g_squared);
    double gmc = g - c, gpc = g + c;
    double x = (c * gpc - 1.0) / (c * gmc + 1.0);
    return 0.5 * (gmc * gmc) / (gpc * gpc) * (1.0 + x * x);
}
inline Vec3 schlick(Vec3 r0v, double angle) {  // mod.rs:90-92 (== metal.rs:111-113, clearcoat.rs:64-66)
    return r0v + (1.0 - r0v) * powi5(1.0 - angle);
}
inline double schlick_weight(double x) { return powi5(clamp_(1.0 - x, 0.0, 1.0)); }  // mod.rs:94-96
}  // namespace fresnel

// ---------------------------------------------------------------- bsdf/sampling.rs
inline Vec3 to_local(Vec3 normal, Vec3 w) { return quat_mul_vec3(get_rotation_to_z(normal), w); }                // :8-11
inline Vec3 to_world(Vec3 normal, Vec3 w) { return quat_mul_vec3(quat_inverse(get_rotation_to_z(normal)), w); }  // :13-16
inline Vec3 cosine_sample_hemisphere(Rng& rng) {  // :18-24 (phi drawn first; RNG contract: phi = U*2pi)
    double phi = rng.next() * (2.0 * PI);
    double r2 = rng.next();
    double r2s = std::sqrt(r2);
    return Vec3(r2s * std::cos(phi), r2s * std::sin(phi), std::sqrt(1.0 - r2));
}
namespace ggx {
inline double D(Vec3 h, double roughness) {  // :38-43
    double cos_theta = fmax_(h.z, 0.001);
    double alpha2 = fmax_(roughness * roughness, 0.001);
    double denom = (alpha2 - 1.0) * (cos_theta * cos_theta) + 1.0;
    return alpha2 / (PI * denom * denom);
}
inline double G1(Vec3 w, double roughness) {  // :51-55
    double alpha2 = fmax_(roughness * roughness, 0.001);
    double c = std::fabs(w.z);
    return 2.0 * c / (c + std::sqrt(c * c * (1.0 - alpha2) + alpha2));
}
inline double G(Vec3 v, Vec3 l, double roughness) { return G1(v, roughness) * G1(l, roughness); }  // :45-49
inline Vec3 sample_ggx_vndf(Vec3 v, double a2, Rng& rng) {  // :66-94
    v = normalize(Vec3(v.x * a2, v.y * a2, v.z));
    Vec3 t1 = v.z < 0.9999 ? normalize(cross(v, Vec3(0, 0, 1))) : Vec3(1, 0, 0);
    Vec3 t2 = cross(t1, v);
    double e1 = rng.next(), e2 = rng.next();
    double a = 1.0 / (1.0 + v.z);
    double r = std::sqrt(e1);
    double phi = e2 < a ? e2 / a * PI : PI + (e2 - a) / (1.0 - a) * PI;
    double p1 = r * std::cos(phi);
    double p2 = r * std::sin(phi) * (e2 < a ? 1.0 : v.z);
    Vec3 n = p1 * t1 + p2 * t2 + std::sqrt(fmax_(1.0 - p1 * p1 - p2 * p2, 0.0)) * v;
    return normalize(Vec3(a2 * n.x, a2 * n.y, fmax_(n.z, 0.0)));
}
inline Vec3 sample_microfacet_normal(Vec3 v, double roughness, Rng& rng) {  // :57-64 (stretch = roughness^2, Q15)
    Vec3 h = sample_ggx_vndf(v, roughness * roughness, rng);
    return h.z < 0.0 ? -h : h;
}
}  // namespace ggx
namespace gtr1 {
inline double D(double abs_cos_theta, double alpha_g) {  // :121-125 (log2, Q16)
    double alpha2 = alpha_g * alpha_g;
    double t = 1.0 + (alpha2 - 1.0) * abs_cos_theta * abs_cos_theta;
    return (alpha2 - 1.0) / (PI * t * std::log2(alpha2));
}
inline Vec3 sample_microfacet_normal(double alpha, Rng& rng) {  // :127-142
    double e1 = rng.next(), e2 = rng.next();
    double alpha2 = alpha * alpha;
    double cos_theta = (1.0 - std::pow(alpha2, 1.0 - e1)) / (1.0 - alpha2);
    double sin_theta = std::sqrt(fmax_(1.0 - cos_theta * cos_theta, 0.0));
    double phi = 2.0 * PI * e2;
    Vec3 h(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta);
    return h.z < 0.0 ? -h : h;
}
}  // namespace gtr1

// ---------------------------------------------------------------- bsdf/diffuse.rs
struct DiffuseBRDF : BxDF {
    const Texture* base_color; const ImageTexture* nmap;
    DiffuseBRDF(const Texture* c, const ImageTexture* n) : base_color(c), nmap(n) {}
    std::optional<Vec3> sample(const Ray&, const HitInfo& info, Rng& rng) const override {  // :51-54
        Vec3 d = cosine_sample_hemisphere(rng);
        return to_world(info.shading_normal, d);
    }
    double pdf(Vec3, Vec3 light_dir, const HitInfo& info) const override {  // :56-59
        Vec3 l = to_local(info.shading_normal, light_dir);
        return std::fabs(l.z) / PI;
    }
    Vec3 eval(Vec3, Vec3 light_dir, const HitInfo& info) const override {  // :61-65
        Vec3 color = base_color->value(info.u, info.v, info.point);
        Vec3 l = to_local(info.shading_normal, light_dir);
        return std::fabs(l.z) * (color / PI);
    }
    const ImageTexture* normal_map() const override { return nmap; }  // :81-83
};

// ---------------------------------------------------------------- bsdf/metal.rs
struct MetalBRDF : BxDF {
    const Texture *base_color, *roughness;
    MetalBRDF(const Texture* c, const Texture* r) : base_color(c), roughness(r) {}
    std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const override {  // :39-54
        Vec3 v = to_local(info.shading_normal, -ray.direction);
        double rough = roughness->value(info.u, info.v, info.point).x;
        Vec3 h = ggx::sample_microfacet_normal(v, rough, rng);
        Vec3 dir = to_world(info.shading_normal, reflect(-v, h));
        if (dot(dir, info.shading_normal) <= 0.0) return std::nullopt;
        return dir;
    }
    double pdf(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :56-67
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        Vec3 h = normalize(v + l);
        double rough = roughness->value(info.u, info.v, info.point).x;
        double pdf_h = ggx::G1(v, rough) * std::fabs(dot(v, h)) * ggx::D(h, rough) / std::fabs(v.z);
        double jacobian = 1.0 / (4.0 * std::fabs(dot(l, h)));
        return pdf_h * jacobian;
    }
    Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :69-80
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        Vec3 h = normalize(v + l);
        double rough = roughness->value(info.u, info.v, info.point).x;
        Vec3 color = base_color->value(info.u, info.v, info.point);
        double d = ggx::D(h, rough);
        double g = ggx::G(v, l, rough);
        Vec3 f = fresnel::schlick(color, dot(l, h));
        return std::fabs(l.z) * (f * g * d / (4.0 * std::fabs(l.z) * std::fabs(v.z)));
    }
};

// shared by glass.rs:92-163 and principled.rs:170-246
inline Vec3 generalized_half(Vec3 v, Vec3 l, double eta_i, double eta_o, bool reflect_) {
    if (reflect_) return normalize(l + v) * signum_(v.z);
    return -normalize(l * eta_o + v * eta_i);
}
inline double glass_pdf_impl(Vec3 v, Vec3 l, Vec3 h, double rough, double eta_i, double eta_o, bool reflect_) {
    double pdf_h = ggx::G1(v, rough) * std::fabs(dot(v, h)) * ggx::D(h, rough) / std::fabs(v.z);
    double f = fresnel::dielectric(v, h, eta_i, eta_o);
    double jacobian;
    if (reflect_) {
        jacobian = f * 1.0 / (4.0 * std::fabs(dot(l, h)));
    } else {
        double v_dot_h = dot(v, h), l_dot_h = dot(l, h);
        jacobian = (1.0 - f) * (eta_o * eta_o * std::fabs(l_dot_h)) / powi2(eta_i * v_dot_h + eta_o * l_dot_h);
    }
    return pdf_h * jacobian;
}
inline double glass_eval_factor(Vec3 v, Vec3 l, Vec3 h, double rough, double eta_i, double eta_o, bool reflect_) {
    double d = ggx::D(h, rough);
    double g = ggx::G(v, l, rough);
    double f = fresnel::dielectric(v, h, eta_i, eta_o);
    if (reflect_) return f * g * d / (4.0 * std::fabs(l.z) * std::fabs(v.z));
    double l_dot_h = dot(l, h), v_dot_h = dot(v, h);
    double term1 = std::fabs((l_dot_h * v_dot_h) / (l.z * v.z));
    double term2 = (eta_o * eta_o) / powi2(eta_i * v_dot_h + eta_o * l_dot_h);
    return term1 * term2 * (1.0 - f) * g * d;
}
inline Vec3 sample_glass_local(Vec3 v, double rough, double eta_i, double eta_o, Rng& rng) {
    Vec3 h = ggx::sample_microfacet_normal(v, rough, rng);
    double f = fresnel::dielectric(v, h, eta_i, eta_o);
    if (rng.next() < f) return reflect(-v, h);
    Vec3 t = refract(-v, h, eta_i / eta_o);
    if (t == Vec3(0, 0, 0)) t = reflect(-v, h);
    return t;
}

// ---------------------------------------------------------------- bsdf/glass.rs
struct GlassBSDF : BxDF {
    const Texture *base_color, *roughness; double ior;
    GlassBSDF(const Texture* c, const Texture* r, double ior_) : base_color(c), roughness(r), ior(ior_) {}
    std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const override {  // :66-90
        Vec3 v = to_local(info.shading_normal, -ray.direction);
        double rough = roughness->value(info.u, info.v, info.point).x;
        double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
        return to_world(info.shading_normal, sample_glass_local(v, rough, eta_i, eta_o, rng));
    }
    double pdf(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :92-123
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        bool refl = l.z * v.z > 0.0;
        double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
        Vec3 h = generalized_half(v, l, eta_i, eta_o, refl);
        double rough = roughness->value(info.u, info.v, info.point).x;
        return glass_pdf_impl(v, l, h, rough, eta_i, eta_o, refl);
    }
    Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :125-163 (colourless, Q19)
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        bool refl = l.z * v.z > 0.0;
        double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
        Vec3 h = generalized_half(v, l, eta_i, eta_o, refl);
        double rough = roughness->value(info.u, info.v, info.point).x;
        return Vec3::splat(glass_eval_factor(v, l, h, rough, eta_i, eta_o, refl)) * std::fabs(l.z);
    }
};

// ---------------------------------------------------------------- bsdf/principled.rs
struct PrincipledBSDF : BxDF {
    const Texture* base_color;
    double metallic, roughness, subsurface, specular, specular_tint, ior, spec_trans, sheen, sheen_tint, clearcoat,
        clearcoat_gloss;
    double get_alpha_g() const { return (1.0 - clearcoat_gloss) * 0.1 + clearcoat_gloss * 0.001; }  // :75-77
    void lobe_weights(double& d, double& s, double& g, double& c) const {  // :79-85
        d = (1.0 - metallic) * (1.0 - spec_trans);
        s = 1.0 - spec_trans * (1.0 - metallic);
        g = spec_trans * (1.0 - metallic);
        c = 0.25 * clearcoat;
    }
    static void lobe_probabilities(double d, double s, double g, double c, double& dp, double& sp, double& gp, double& cp) {
        double inv_total = 1.0 / (d + s + g + c);  // :87-100
        dp = d * inv_total; sp = s * inv_total; gp = g * inv_total; cp = c * inv_total;
    }
    std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const override {  // :262-277
        double dw, sw, gw, cw, dp, sp, gp, cp;
        lobe_weights(dw, sw, gw, cw);
        lobe_probabilities(dw, sw, gw, cw, dp, sp, gp, cp);
        double r = rng.next();
        Vec3 n = info.geometric_normal;  // Q18
        if (r < dp) return to_world(n, cosine_sample_hemisphere(rng));  // :102-104
        Vec3 v = to_local(n, -ray.direction);
        if (r < dp + sp) {  // :106-118
            Vec3 h = ggx::sample_microfacet_normal(v, roughness, rng);
            Vec3 dir = to_world(n, reflect(-v, h));
            if (dot(dir, n) <= 0.0) return std::nullopt;
            return dir;
        }
        if (r < dp + sp + gp) {  // :120-142
            double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
            return to_world(n, sample_glass_local(v, roughness, eta_i, eta_o, rng));
        }
        Vec3 h = gtr1::sample_microfacet_normal(0.25, rng);  // :144-155
        Vec3 dir = to_world(n, reflect(-v, h));
        if (dot(dir, n) <= 0.0) return std::nullopt;
        return dir;
    }
    double clearcoat_pdf(Vec3 v, Vec3 l, Vec3 h) const {  // :187-192
        double pdf_h = ggx::G1(v, 0.25) * std::fabs(dot(v, h)) * gtr1::D(std::fabs(dot(l, h)), get_alpha_g()) / std::fabs(v.z);
        return pdf_h * (1.0 / (4.0 * std::fabs(dot(l, h))));
    }
    double pdf(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :279-315
        double dw, sw, gw, cw, dp, sp, gp, cp;
        lobe_weights(dw, sw, gw, cw);
        lobe_probabilities(dw, sw, gw, cw, dp, sp, gp, cp);
        Vec3 v = to_local(info.geometric_normal, view_dir), l = to_local(info.geometric_normal, light_dir);
        bool refl = l.z * v.z > 0.0;
        double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
        Vec3 h = generalized_half(v, l, eta_i, eta_o, refl);
        double pdf = 0.0;
        if (dp > 0.0 && refl) pdf += dp * (std::fabs(l.z) / PI);  // :157-159
        if (sp > 0.0 && refl) {                                   // :161-168
            double pdf_h = ggx::G1(v, roughness) * std::fabs(dot(v, h)) * ggx::D(h, roughness) / std::fabs(v.z);
            pdf += sp * (pdf_h * (1.0 / (4.0 * std::fabs(dot(l, h)))));
        }
        if (gp > 0.0) pdf += gp * glass_pdf_impl(v, l, h, roughness, eta_i, eta_o, refl);  // :170-185
        if (cp > 0.0 && refl) pdf += cp * clearcoat_pdf(v, l, h);
        return pdf;
    }
    Vec3 eval_diffuse(Vec3 color, Vec3 v, Vec3 l, Vec3 h) const {  // :196-213
        double l_dot_h = dot(l, h);
        double rr = 2.0 * roughness * l_dot_h * l_dot_h;
        double fl = fresnel::schlick_weight(l.z), fv = fresnel::schlick_weight(v.z);
        double f_retro = rr * (fl + fv + fl * fv * (rr - 1.0));
        double f_d = (1.0 - 0.5 * fl) * (1.0 - 0.5 * fv);
        double fss90 = 0.5 * rr;
        double f_ss = lerp(1.0, fss90, fl) * lerp(1.0, fss90, fv);
        double ss = 1.25 * (f_ss * (1.0 / (l.z + v.z) - 0.5) + 0.5);
        return color / PI * lerp(f_d + f_retro, ss, subsurface);
    }
    Vec3 eval_clearcoat(Vec3 v, Vec3 l, Vec3 h) const {  // :248-258 (extra |l.z|, Q17)
        double d = gtr1::D(std::fabs(dot(l, h)), get_alpha_g());
        double g = ggx::G(v, l, 0.25);
        Vec3 f = fresnel::schlick(Vec3::splat(r0(1.5)), dot(l, h));
        return std::fabs(l.z) * (f * d * g / (4.0 * std::fabs(l.z) * std::fabs(v.z)));
    }
    Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :317-366
        Vec3 color = base_color->value(info.u, info.v, info.point);
        double dw, sw, gw, cw, dp, sp, gp, cp;
        lobe_weights(dw, sw, gw, cw);
        lobe_probabilities(dw, sw, gw, cw, dp, sp, gp, cp);
        Vec3 v = to_local(info.geometric_normal, view_dir), l = to_local(info.geometric_normal, light_dir);
        bool refl = l.z * v.z > 0.0;
        double eta_i = info.front_face ? 1.0 : ior, eta_o = info.front_face ? ior : 1.0;
        Vec3 h = generalized_half(v, l, eta_i, eta_o, refl);
        Vec3 brdf(0, 0, 0);
        if (dp > 0.0 && refl) {
            Vec3 c_tint = tint(color);
            Vec3 c_sheen = lerp(Vec3(1, 1, 1), c_tint, sheen_tint);
            Vec3 sheen_term = sheen * c_sheen * fresnel::schlick_weight(std::fabs(dot(l, h)));
            Vec3 diffuse_term = eval_diffuse(color, v, l, h);
            brdf += dw * (diffuse_term + sheen_term);
        }
        if (sp > 0.0 && refl) {
            Vec3 c_tint = tint(color);
            Vec3 ks = lerp(Vec3(1, 1, 1), c_tint, specular_tint);
            Vec3 c0 = lerp(specular * r0(eta_i / eta_o) * ks, color, metallic);
            Vec3 metallic_fresnel = fresnel::schlick(c0, dot(l, h));
            Vec3 dielectric_fresnel = Vec3::splat(fresnel::dielectric(v, h, eta_i, eta_o));
            Vec3 fr = lerp(dielectric_fresnel, metallic_fresnel, metallic);
            double d = ggx::D(h, roughness), g = ggx::G(v, l, roughness);  // :215-224
            brdf += sw * (fr * g * d / (4.0 * std::fabs(l.z) * std::fabs(v.z)));
        }
        if (gp > 0.0) brdf += gw * Vec3::splat(glass_eval_factor(v, l, h, roughness, eta_i, eta_o, refl));
        if (cp > 0.0 && refl) brdf += cw * eval_clearcoat(v, l, h);
        return brdf * std::fabs(l.z);
    }
};

// ---------------------------------------------------------------- material.rs:150-191
struct DiffuseLight : BxDF {
    const Texture* emission;
    explicit DiffuseLight(const Texture* e) : emission(e) {}
    std::optional<Vec3> sample(const Ray&, const HitInfo&, Rng&) const override { return std::nullopt; }
    double pdf(Vec3, Vec3, const HitInfo&) const override { return 1.0; }
    Vec3 eval(Vec3, Vec3, const HitInfo&) const override { return Vec3(1, 1, 1); }
    Vec3 emitted(double u, double v, Vec3 p) const override { return emission->value(u, v, p); }
    bool is_emitter() const override { return true; }
};

// ---------------------------------------------------------------- volume.rs:18 phase_function (stub in the reference)
// NOT reference behaviour (include/pt_b200.h, PT_MAT_ISOTROPIC): direction uniform on the sphere, albedo from a texture.
struct IsotropicMaterial : BxDF {
    const Texture* albedo;
    explicit IsotropicMaterial(const Texture* a) : albedo(a) {}
    std::optional<Vec3> sample(const Ray&, const HitInfo&, Rng& rng) const override {
        double u1 = rng.next(), u2 = rng.next();
        double z = 1.0 - 2.0 * u1;
        double r = std::sqrt(fmax_(0.0, 1.0 - z * z));
        double phi = 2.0 * PI * u2;
        return Vec3(r * std::cos(phi), r * std::sin(phi), z);
    }
    double pdf(Vec3, Vec3, const HitInfo&) const override { return 1.0 / (4.0 * PI); }
    Vec3 eval(Vec3, Vec3, const HitInfo& info) const override { return albedo->value(info.u, info.v, info.point) * (1.0 / (4.0 * PI)); }
};

// ---------------------------------------------------------------- bsdf/sheen.rs
struct SheenBRDF : BxDF {
    Vec3 base_color; double sheen_tint;
    SheenBRDF(Vec3 c, double t) : base_color(c), sheen_tint(t) {}
    std::optional<Vec3> sample(const Ray&, const HitInfo& info, Rng& rng) const override {
        Vec3 d = cosine_sample_hemisphere(rng);
        return to_world(info.geometric_normal, d);
    }
    double pdf(Vec3, Vec3 light_dir, const HitInfo& info) const override {
        return std::fabs(to_local(info.geometric_normal, light_dir).z) / PI;
    }
    Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :36-43
        Vec3 v = to_local(info.geometric_normal, view_dir), l = to_local(info.geometric_normal, light_dir);
        Vec3 h = normalize(v + l);
        Vec3 c_sheen = lerp(Vec3(1, 1, 1), tint(base_color), sheen_tint);
        return c_sheen * powi5(1.0 - std::fabs(dot(l, h))) * std::fabs(l.z);
    }
};

// ---------------------------------------------------------------- bsdf/clearcoat.rs
struct ClearcoatBRDF : BxDF {
    double alpha_g;
    explicit ClearcoatBRDF(double a) : alpha_g(a) {}
    std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const override {  // :23-35
        Vec3 v = to_local(info.shading_normal, -ray.direction);
        Vec3 h = gtr1::sample_microfacet_normal(0.25, rng);
        Vec3 dir = to_world(info.shading_normal, reflect(-v, h));
        if (dot(dir, info.shading_normal) <= 0.0) return std::nullopt;
        return dir;
    }
    double pdf(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :37-45
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        Vec3 h = normalize(v + l);
        double pdf_h = ggx::G1(v, 0.25) * std::fabs(dot(v, h)) * gtr1::D(std::fabs(dot(l, h)), alpha_g) / std::fabs(v.z);
        return pdf_h * (1.0 / (4.0 * std::fabs(dot(l, h))));
    }
    Vec3 eval(Vec3 view_dir, Vec3 light_dir, const HitInfo& info) const override {  // :47-61
        Vec3 v = to_local(info.shading_normal, view_dir), l = to_local(info.shading_normal, light_dir);
        Vec3 h = normalize(v + l);
        double d = gtr1::D(std::fabs(dot(l, h)), alpha_g);
        double g = ggx::G(v, l, 0.25);
        Vec3 f = fresnel::schlick(Vec3::splat(r0(1.5)), dot(l, h));
        return std::fabs(l.z) * (f * d * g / (4.0 * std::fabs(l.z) * std::fabs(v.z)));
    }
};

// ---------------------------------------------------------------- bsdf/mix.rs
struct MixBxDf : BxDF {
    double t; const BxDF *b1, *b2;
    MixBxDf(double t_, const BxDF* a, const BxDF* b) : t(clamp_(t_, 0.0, 1.0)), b1(a), b2(b) {}
    std::optional<Vec3> sample(const Ray& ray, const HitInfo& info, Rng& rng) const override {  // :25-32
        double p = rng.next();
        if (t < p) return b1->sample(ray, info, rng);
        return b2->sample(ray, info, rng);
    }
    double pdf(Vec3 v, Vec3 l, const HitInfo& info) const override {  // :34-38
        double p1 = (1.0 - t) * b1->pdf(v, l, info);
        double p2 = t * b2->pdf(v, l, info);
        return p1 + p2;
    }
    Vec3 eval(Vec3 v, Vec3 l, const HitInfo& info) const override {  // :40-44
        Vec3 w1 = (1.0 - t) * b1->eval(v, l, info);
        Vec3 w2 = t * b2->eval(v, l, info);
        return w1 + w2;
    }
};

}  // namespace orc
