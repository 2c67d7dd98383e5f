// Device shading library: textures (src/texture.rs), BSDF lobes and samplers (src/bsdf/*.rs) and the
// emitter (src/material.rs:150-191).  f64, reference operation order, quirks Q15-Q23 kept.
#pragma once
#include "geom.cuh"

namespace ptd {

// ---------------------------------------------------------------- textures
PT_D d3 texture_value(const DScene& S, uint32_t tex, double u, double v, d3 p) {
    // CheckerTexture children are arbitrary textures (texture.rs:27-31): descend iteratively
    for (int depth = 0; depth < 16; depth++) {
        const DTexture& t = S.textures[tex];
        if (t.kind == PT_TEX_SOLID) return mk(t.value[0], t.value[1], t.value[2]);
        if (t.kind == PT_TEX_IMAGE) return image_value(S, t.image, u, v);
        // texture.rs:44-53: floor(p * inv_scale) as i32 (saturating), wrapping add, sign-keeping %
        double fx = floor(p.x * t.inv_scale), fy = floor(p.y * t.inv_scale), fz = floor(p.z * t.inv_scale);
        int32_t x = __double2int_rz(fx), y = __double2int_rz(fy), z = __double2int_rz(fz);  // cvt.rzi.s32.f64 saturates, NaN -> 0
        int32_t s = (int32_t)((uint32_t)x + (uint32_t)y + (uint32_t)z);
        tex = (s % 2 == 0) ? t.tex1 : t.tex2;
    }
    return mk(0, 0, 0);
}

// ---------------------------------------------------------------- bsdf/mod.rs:61-97
PT_D d3 tint(d3 c) { return luminance(c) > 0.0 ? c / luminance(c) : mk(1, 1, 1); }
PT_D double r0f(double eta) { return powi2((eta - 1.0) / (eta + 1.0)); }
PT_D double fresnel_dielectric(d3 w, d3 h, double eta_i, double eta_o) {
    double c = fabs(dot(w, h));
    double g2 = powi2(eta_o / eta_i) - 1.0 + c * c;
    if (g2 < 0.0) return 1.0;
    double g = sqrt(g2);
    double gmc = g - c, gpc = g + c;
    double x = (c * gpc - 1.0) / (c * gmc + 1.0);
    return 0.5 * (gmc * gmc) / (gpc * gpc) * (1.0 + x * x);
}
PT_D d3 fresnel_schlick(d3 r0v, double angle) { return r0v + sub_from(1.0, r0v) * powi5(1.0 - angle); }
PT_D double schlick_weight(double x) { return powi5(clampd(1.0 - x, 0.0, 1.0)); }

// ---------------------------------------------------------------- bsdf/sampling.rs
PT_D d3 cosine_sample_hemisphere(Rng& rng) {  // :18-24
    double phi = rng.next() * (2.0 * kPi);
    double r2 = rng.next();
    double r2s = sqrt(r2);
    double sp, cp; pt_sincos(phi, sp, cp);
    return mk(r2s * cp, r2s * sp, sqrt(1.0 - r2));
}
PT_D double ggx_D(d3 h, double roughness) {  // :38-43
    double ct = fmax(h.z, 0.001);
    double a2 = fmax(roughness * roughness, 0.001);
    double denom = (a2 - 1.0) * (ct * ct) + 1.0;
    return a2 / (kPi * denom * denom);
}
PT_D double ggx_G1(d3 w, double roughness) {  // :51-55
    double a2 = fmax(roughness * roughness, 0.001);
    double c = fabs(w.z);
    return 2.0 * c / (c + sqrt(c * c * (1.0 - a2) + a2));
}
PT_D double ggx_G(d3 v, d3 l, double roughness) { return ggx_G1(v, roughness) * ggx_G1(l, roughness); }
PT_D d3 ggx_sample_normal(d3 v_in, double roughness, Rng& rng) {  // :57-94 (stretch = roughness^2, Q15)
    double a2 = roughness * roughness;
    d3 v = normalize(mk(v_in.x * a2, v_in.y * a2, v_in.z));
    d3 t1 = v.z < 0.9999 ? normalize(cross(v, mk(0, 0, 1))) : mk(1, 0, 0);
    d3 t2 = cross(t1, v);
    double e1 = rng.next(), e2 = rng.next();
    double a = 1.0 / (1.0 + v.z);
    double r = sqrt(e1);
    double phi = e2 < a ? e2 / a * kPi : kPi + (e2 - a) / (1.0 - a) * kPi;
    double sp, cp; pt_sincos(phi, sp, cp);
    double p1 = r * cp;
    double p2 = r * sp * (e2 < a ? 1.0 : v.z);
    d3 n = p1 * t1 + p2 * t2 + sqrt(fmax(1.0 - p1 * p1 - p2 * p2, 0.0)) * v;
    d3 h = normalize(mk(a2 * n.x, a2 * n.y, fmax(n.z, 0.0)));
    return h.z < 0.0 ? -h : h;
}
PT_D double gtr1_D(double abs_cos, double alpha_g) {  // :121-125 (log2, Q16)
    double a2 = alpha_g * alpha_g;
    double t = 1.0 + (a2 - 1.0) * abs_cos * abs_cos;
    return (a2 - 1.0) / (kPi * t * log2(a2));
}
PT_D d3 gtr1_sample_normal(double alpha, Rng& rng) {  // :127-142
    double e1 = rng.next(), e2 = rng.next();
    double a2 = alpha * alpha;
    double ct = (1.0 - pow(a2, 1.0 - e1)) / (1.0 - a2);
    double st = sqrt(fmax(1.0 - ct * ct, 0.0));
    double phi = 2.0 * kPi * e2;
    double sp, cp; pt_sincos(phi, sp, cp);
    d3 h = mk(st * cp, st * sp, ct);
    return h.z < 0.0 ? -h : h;
}

// ---------------------------------------------------------------- shared glass pieces (glass.rs / principled.rs)
PT_D d3 generalized_half(d3 v, d3 l, double eta_i, double eta_o, bool refl) {
    if (refl) return normalize(l + v) * signum(v.z);
    return -normalize(l * eta_o + v * eta_i);
}
PT_D double glass_pdf(d3 v, d3 l, d3 h, double rough, double eta_i, double eta_o, bool refl) {
    double pdf_h = ggx_G1(v, rough) * fabs(dot(v, h)) * ggx_D(h, rough) / fabs(v.z);
    double f = fresnel_dielectric(v, h, eta_i, eta_o);
    double jac;
    if (refl) jac = f * 1.0 / (4.0 * fabs(dot(l, h)));
    else {
        double vh = dot(v, h), lh = dot(l, h);
        jac = (1.0 - f) * (eta_o * eta_o * fabs(lh)) / powi2(eta_i * vh + eta_o * lh);
    }
    return pdf_h * jac;
}
PT_D double glass_factor(d3 v, d3 l, d3 h, double rough, double eta_i, double eta_o, bool refl) {
    double d = ggx_D(h, rough), g = ggx_G(v, l, rough);
    double f = fresnel_dielectric(v, h, eta_i, eta_o);
    if (refl) return f * g * d / (4.0 * fabs(l.z) * fabs(v.z));
    double lh = dot(l, h), vh = dot(v, h);
    double term1 = fabs((lh * vh) / (l.z * v.z));
    double term2 = (eta_o * eta_o) / powi2(eta_i * vh + eta_o * lh);
    return term1 * term2 * (1.0 - f) * g * d;
}
PT_D d3 glass_sample_local(d3 v, double rough, double eta_i, double eta_o, Rng& rng) {
    d3 h = ggx_sample_normal(v, rough, rng);
    double f = fresnel_dielectric(v, h, eta_i, eta_o);
    if (rng.next() < f) return reflect(-v, h);
    d3 t = refract(-v, h, eta_i / eta_o);
    if (t.x == 0.0 && t.y == 0.0 && t.z == 0.0) t = reflect(-v, h);
    return t;
}

// ---------------------------------------------------------------- per-material sample / pdf / eval / emitted
struct PrincipledLobes { double dw, sw, gw, cw, dp, sp, gp, cp; };
PT_D PrincipledLobes principled_lobes(const DMaterial& m) {  // principled.rs:79-100
    PrincipledLobes L;
    double metallic = m.p[PT_P_METALLIC], st = m.p[PT_P_SPEC_TRANS];
    L.dw = (1.0 - metallic) * (1.0 - st);
    L.sw = 1.0 - st * (1.0 - metallic);
    L.gw = st * (1.0 - metallic);
    L.cw = 0.25 * m.p[PT_P_CLEARCOAT];
    double inv_total = 1.0 / (L.dw + L.sw + L.gw + L.cw);
    L.dp = L.dw * inv_total; L.sp = L.sw * inv_total; L.gp = L.gw * inv_total; L.cp = L.cw * inv_total;
    return L;
}
PT_D double principled_alpha_g(const DMaterial& m) { double g = m.p[PT_P_CLEARCOAT_GLOSS]; return (1.0 - g) * 0.1 + g * 0.001; }

// What BOTH BxDFMaterial::sample and ::{eval, pdf} derive from the hit alone — the rotation of the material's frame normal onto +z
// (sampling.rs:8-16: shading normal; the principled BSDF uses the geometric normal, Q18), the view direction in that frame and the
// roughness texture value.  The reference recomputes them in each of its three calls; the per-class shade kernels compute them once
// (bsdf_prepare) and hand them to both functions: same expressions on the same inputs, hence the same bits.  cx == nullptr (the
// parity entry points, mix materials): every function derives them itself as before.
#ifndef PT_BSDF_CTX
#define PT_BSDF_CTX 1
#endif
struct BsdfCtx { q4 q; d3 v; double rough; };
PT_D d3 to_local_q(const q4& q, d3 w) { return quat_mul(q, w); }
PT_D d3 to_world_q(q4 q, d3 w) { q.x = -q.x; q.y = -q.y; q.z = -q.z; return quat_mul(q, w); }
template <int K>
PT_D BsdfCtx bsdf_prepare(const DScene& S, const DMaterial& m, d3 view_dir, const HitInfoD& h) {
    BsdfCtx c;
    c.q = rotation_to_z(K == PT_MAT_PRINCIPLED ? h.gn : h.sn);
    c.v = K == PT_MAT_DIFFUSE ? mk(0, 0, 0) : to_local_q(c.q, view_dir);
    c.rough = (K == PT_MAT_METAL || K == PT_MAT_GLASS) ? texture_value(S, m.roughness_tex, h.u, h.v, h.point).x : 0.0;
    return c;
}

// BxDFMaterial::sample (leaf materials). ray_dir = incoming ray direction; returns false for None.
// K >= 0: the material kind is known at compile time (per-class shade kernels) and the switch folds away.
template <int K = -1>
PT_D bool bsdf_sample_leaf(const DScene& S, const DMaterial& m, d3 ray_dir, const HitInfoD& h, Rng& rng, d3& out, const BsdfCtx* cx = nullptr) {
    switch (K >= 0 ? (uint32_t)K : m.kind) {
        case PT_MAT_DIFFUSE: {  // diffuse.rs:51-54
            const d3 w = cosine_sample_hemisphere(rng);
            out = cx ? to_world_q(cx->q, w) : to_world(h.sn, w);
            return true;
        }
        case PT_MAT_METAL: {  // metal.rs:39-54
            d3 v = cx ? cx->v : to_local(h.sn, -ray_dir);
            double rough = cx ? cx->rough : texture_value(S, m.roughness_tex, h.u, h.v, h.point).x;
            d3 hh = ggx_sample_normal(v, rough, rng);
            out = cx ? to_world_q(cx->q, reflect(-v, hh)) : to_world(h.sn, reflect(-v, hh));
            return !(dot(out, h.sn) <= 0.0);
        }
        case PT_MAT_GLASS: {  // glass.rs:66-90
            d3 v = cx ? cx->v : to_local(h.sn, -ray_dir);
            double rough = cx ? cx->rough : texture_value(S, m.roughness_tex, h.u, h.v, h.point).x;
            double ior = m.p[PT_P_IOR];
            double eta_i = h.front_face ? 1.0 : ior, eta_o = h.front_face ? ior : 1.0;
            const d3 w = glass_sample_local(v, rough, eta_i, eta_o, rng);
            out = cx ? to_world_q(cx->q, w) : to_world(h.sn, w);
            return true;
        }
        case PT_MAT_PRINCIPLED: {  // principled.rs:262-277 (geometric normal, Q18)
            PrincipledLobes L = principled_lobes(m);
            double r = rng.next();
            d3 n = h.gn;
            if (r < L.dp) { const d3 w = cosine_sample_hemisphere(rng); out = cx ? to_world_q(cx->q, w) : to_world(n, w); return true; }
            d3 v = cx ? cx->v : to_local(n, -ray_dir);
            double rough = m.p[PT_P_ROUGHNESS];
            if (r < L.dp + L.sp) {
                d3 hh = ggx_sample_normal(v, rough, rng);
                out = cx ? to_world_q(cx->q, reflect(-v, hh)) : to_world(n, reflect(-v, hh));
                return !(dot(out, n) <= 0.0);
            }
            if (r < L.dp + L.sp + L.gp) {
                double ior = m.p[PT_P_IOR];
                double eta_i = h.front_face ? 1.0 : ior, eta_o = h.front_face ? ior : 1.0;
                const d3 w = glass_sample_local(v, rough, eta_i, eta_o, rng);
                out = cx ? to_world_q(cx->q, w) : to_world(n, w);
                return true;
            }
            d3 hh = gtr1_sample_normal(0.25, rng);  // fixed alpha 0.25 (Q16)
            out = cx ? to_world_q(cx->q, reflect(-v, hh)) : to_world(n, reflect(-v, hh));
            return !(dot(out, n) <= 0.0);
        }
        case PT_MAT_SHEEN: out = to_world(h.gn, cosine_sample_hemisphere(rng)); return true;  // sheen.rs:26-29
        case PT_MAT_CLEARCOAT: {  // clearcoat.rs:23-35
            d3 v = to_local(h.sn, -ray_dir);
            d3 hh = gtr1_sample_normal(0.25, rng);
            out = to_world(h.sn, reflect(-v, hh));
            return !(dot(out, h.sn) <= 0.0);
        }
        case PT_MAT_ISOTROPIC: {  // phase function of a medium (pt_volume; ours): uniform on the sphere
            const double u1 = rng.next(), u2 = rng.next();
            const double z = 1.0 - 2.0 * u1;
            const double r = sqrt(fmax(0.0, 1.0 - z * z));
            const double phi = 2.0 * kPi * u2;
            double sp, cp; pt_sincos(phi, sp, cp);
            out = mk(r * cp, r * sp, z);
            return true;
        }
        default: return false;  // DiffuseLight::sample -> None (material.rs:168-170)
    }
}

PT_D double clearcoat_pdf(d3 v, d3 l, d3 h, double alpha_g) {
    double pdf_h = ggx_G1(v, 0.25) * fabs(dot(v, h)) * gtr1_D(fabs(dot(l, h)), alpha_g) / fabs(v.z);
    return pdf_h * (1.0 / (4.0 * fabs(dot(l, h))));
}
PT_D d3 clearcoat_eval(d3 v, d3 l, d3 h, double alpha_g) {  // extra |l.z| (Q17)
    double d = gtr1_D(fabs(dot(l, h)), alpha_g);
    double g = ggx_G(v, l, 0.25);
    d3 f = fresnel_schlick(splat(r0f(1.5)), dot(l, h));
    return fabs(l.z) * (f * d * g / (4.0 * fabs(l.z) * fabs(v.z)));
}

// BxDFMaterial::{pdf, eval} for leaf materials, computed together (they share frames and half vectors).
template <int K = -1>
PT_D void bsdf_eval_pdf_leaf(const DScene& S, const DMaterial& m, d3 view_dir, d3 light_dir, const HitInfoD& hi, d3& f_out, double& pdf_out, const BsdfCtx* cx = nullptr) {
    switch (K >= 0 ? (uint32_t)K : m.kind) {
        case PT_MAT_DIFFUSE: {  // diffuse.rs:56-65
            d3 color = texture_value(S, m.base_color_tex, hi.u, hi.v, hi.point);
            d3 l = cx ? to_local_q(cx->q, light_dir) : to_local(hi.sn, light_dir);
            pdf_out = fabs(l.z) / kPi;
            f_out = fabs(l.z) * (color / kPi);
            return;
        }
        case PT_MAT_METAL: {  // metal.rs:56-80
            d3 v = cx ? cx->v : to_local(hi.sn, view_dir), l = cx ? to_local_q(cx->q, light_dir) : to_local(hi.sn, light_dir);
            d3 h = normalize(v + l);
            double rough = cx ? cx->rough : texture_value(S, m.roughness_tex, hi.u, hi.v, hi.point).x;
            d3 color = texture_value(S, m.base_color_tex, hi.u, hi.v, hi.point);
            double d = ggx_D(h, rough);
            double pdf_h = ggx_G1(v, rough) * fabs(dot(v, h)) * d / fabs(v.z);
            pdf_out = pdf_h * (1.0 / (4.0 * fabs(dot(l, h))));
            double g = ggx_G(v, l, rough);
            d3 f = fresnel_schlick(color, dot(l, h));
            f_out = fabs(l.z) * (f * g * d / (4.0 * fabs(l.z) * fabs(v.z)));
            return;
        }
        case PT_MAT_GLASS: {  // glass.rs:92-163 (colourless, Q19)
            d3 v = cx ? cx->v : to_local(hi.sn, view_dir), l = cx ? to_local_q(cx->q, light_dir) : to_local(hi.sn, light_dir);
            bool refl = l.z * v.z > 0.0;
            double ior = m.p[PT_P_IOR];
            double eta_i = hi.front_face ? 1.0 : ior, eta_o = hi.front_face ? ior : 1.0;
            d3 h = generalized_half(v, l, eta_i, eta_o, refl);
            double rough = cx ? cx->rough : texture_value(S, m.roughness_tex, hi.u, hi.v, hi.point).x;
            pdf_out = glass_pdf(v, l, h, rough, eta_i, eta_o, refl);
            f_out = splat(glass_factor(v, l, h, rough, eta_i, eta_o, refl)) * fabs(l.z);
            return;
        }
        case PT_MAT_PRINCIPLED: {  // principled.rs:279-366
            d3 color = texture_value(S, m.base_color_tex, hi.u, hi.v, hi.point);
            PrincipledLobes L = principled_lobes(m);
            d3 v = cx ? cx->v : to_local(hi.gn, view_dir), l = cx ? to_local_q(cx->q, light_dir) : to_local(hi.gn, light_dir);
            bool refl = l.z * v.z > 0.0;
            double ior = m.p[PT_P_IOR], rough = m.p[PT_P_ROUGHNESS];
            double eta_i = hi.front_face ? 1.0 : ior, eta_o = hi.front_face ? ior : 1.0;
            d3 h = generalized_half(v, l, eta_i, eta_o, refl);
            double pdf = 0.0; d3 brdf = mk(0, 0, 0);
            // The specular and the glass lobe evaluate the same GGX terms (principled.rs:215-260 and glass.rs:92-163 call ggx_d, ggx_g1,
            // fresnel_dielectric with the same arguments), the diffuse and the specular lobe the same tint: each is computed once here —
            // same expression, same inputs, same bits as the reference's repeated calls.
            const bool spec_on = L.sp > 0.0 && refl, glass_on = L.gp > 0.0;
            double ggx_d_h = 0.0, g1_v = 0.0, g1_l = 0.0, f_diel = 0.0, pdf_h = 0.0;
            if (spec_on || glass_on) {
                ggx_d_h = ggx_D(h, rough); g1_v = ggx_G1(v, rough); g1_l = ggx_G1(l, rough);
                f_diel = fresnel_dielectric(v, h, eta_i, eta_o);
                pdf_h = g1_v * fabs(dot(v, h)) * ggx_d_h / fabs(v.z);
            }
            const d3 tint_c = refl && (L.dp > 0.0 || L.sp > 0.0) ? tint(color) : mk(1, 1, 1);
            if (L.dp > 0.0 && refl) {
                pdf += L.dp * (fabs(l.z) / kPi);
                d3 c_sheen = lerp3(mk(1, 1, 1), tint_c, m.p[PT_P_SHEEN_TINT]);
                d3 sheen_term = m.p[PT_P_SHEEN] * c_sheen * schlick_weight(fabs(dot(l, h)));
                // eval_diffuse, principled.rs:196-213
                double lh = dot(l, h);
                double rr = 2.0 * rough * lh * lh;
                double fl = schlick_weight(l.z), fv = schlick_weight(v.z);
                double f_retro = rr * (fl + fv + fl * fv * (rr - 1.0));
                double f_d = (1.0 - 0.5 * fl) * (1.0 - 0.5 * fv);
                double fss90 = 0.5 * rr;
                double f_ss = lerp1(1.0, fss90, fl) * lerp1(1.0, fss90, fv);
                double ss = 1.25 * (f_ss * (1.0 / (l.z + v.z) - 0.5) + 0.5);
                d3 diffuse_term = color / kPi * lerp1(f_d + f_retro, ss, m.p[PT_P_SUBSURFACE]);
                brdf = brdf + L.dw * (diffuse_term + sheen_term);
            }
            if (spec_on) {
                pdf += L.sp * (pdf_h * (1.0 / (4.0 * fabs(dot(l, h)))));
                d3 ks = lerp3(mk(1, 1, 1), tint_c, m.p[PT_P_SPECULAR_TINT]);
                d3 c0 = lerp3(m.p[PT_P_SPECULAR] * r0f(eta_i / eta_o) * ks, color, m.p[PT_P_METALLIC]);
                d3 mf = fresnel_schlick(c0, dot(l, h));
                d3 df = splat(f_diel);
                d3 fr = lerp3(df, mf, m.p[PT_P_METALLIC]);
                double g = g1_v * g1_l;
                brdf = brdf + L.sw * (fr * g * ggx_d_h / (4.0 * fabs(l.z) * fabs(v.z)));
            }
            if (glass_on) {  // glass_pdf / glass_factor with the shared terms
                double jac, factor;
                const double g = g1_v * g1_l;
                if (refl) {
                    jac = f_diel * 1.0 / (4.0 * fabs(dot(l, h)));
                    factor = f_diel * g * ggx_d_h / (4.0 * fabs(l.z) * fabs(v.z));
                } else {
                    double vh = dot(v, h), lh = dot(l, h);
                    jac = (1.0 - f_diel) * (eta_o * eta_o * fabs(lh)) / powi2(eta_i * vh + eta_o * lh);
                    double term1 = fabs((lh * vh) / (l.z * v.z));
                    double term2 = (eta_o * eta_o) / powi2(eta_i * vh + eta_o * lh);
                    factor = term1 * term2 * (1.0 - f_diel) * g * ggx_d_h;
                }
                pdf += L.gp * (pdf_h * jac);
                brdf = brdf + L.gw * splat(factor);
            }
            if (L.cp > 0.0 && refl) {
                double ag = principled_alpha_g(m);
                pdf += L.cp * clearcoat_pdf(v, l, h, ag);
                brdf = brdf + L.cw * clearcoat_eval(v, l, h, ag);
            }
            pdf_out = pdf;
            f_out = brdf * fabs(l.z);
            return;
        }
        case PT_MAT_LIGHT: pdf_out = 1.0; f_out = mk(1, 1, 1); return;  // material.rs:172-178
        case PT_MAT_SHEEN: {  // sheen.rs:31-43
            d3 v = to_local(hi.gn, view_dir), l = to_local(hi.gn, light_dir);
            d3 h = normalize(v + l);
            pdf_out = fabs(l.z) / kPi;
            d3 c_sheen = lerp3(mk(1, 1, 1), tint(mk(m.p[PT_P_COLOR_R], m.p[PT_P_COLOR_G], m.p[PT_P_COLOR_B])), m.p[PT_P_SHEEN_TINT]);
            f_out = c_sheen * powi5(1.0 - fabs(dot(l, h))) * fabs(l.z);
            return;
        }
        case PT_MAT_CLEARCOAT: {  // clearcoat.rs:37-61
            d3 v = to_local(hi.sn, view_dir), l = to_local(hi.sn, light_dir);
            d3 h = normalize(v + l);
            pdf_out = clearcoat_pdf(v, l, h, m.p[PT_P_ALPHA_G]);
            f_out = clearcoat_eval(v, l, h, m.p[PT_P_ALPHA_G]);
            return;
        }
        case PT_MAT_ISOTROPIC:
            pdf_out = 1.0 / (4.0 * kPi);
            f_out = texture_value(S, m.base_color_tex, hi.u, hi.v, hi.point) * (1.0 / (4.0 * kPi));
            return;
        default: pdf_out = 0.0; f_out = mk(0, 0, 0); return;
    }
}

// MixBxDf (mix.rs:24-45) is a binary tree over leaf materials; walked with a small explicit stack.
constexpr int kMixDepth = 8;
template <int K = -1>
PT_D bool bsdf_sample(const DScene& S, uint32_t mat, d3 ray_dir, const HitInfoD& h, Rng& rng, d3& out, const BsdfCtx* cx = nullptr) {
    if constexpr (K >= 0) {
        return bsdf_sample_leaf<K>(S, S.materials[mat], ray_dir, h, rng, out, cx);
    } else {
        for (int d = 0; d < kMixDepth; d++) {
            const DMaterial& m = S.materials[mat];
            if (m.kind != PT_MAT_MIX) return bsdf_sample_leaf(S, m, ray_dir, h, rng, out);
            double p = rng.next();
            mat = (m.p[PT_P_MIX_T] < p) ? m.mix_a : m.mix_b;  // mix.rs:26-31
        }
        return false;
    }
}
template <int K = -1>
PT_D void bsdf_eval_pdf(const DScene& S, uint32_t mat, d3 view_dir, d3 light_dir, const HitInfoD& h, d3& f_out, double& pdf_out, const BsdfCtx* cx = nullptr) {
    const DMaterial& m0 = S.materials[mat];
    if (K >= 0) { bsdf_eval_pdf_leaf<K>(S, m0, view_dir, light_dir, h, f_out, pdf_out, cx); return; }
    if (m0.kind != PT_MAT_MIX) { bsdf_eval_pdf_leaf(S, m0, view_dir, light_dir, h, f_out, pdf_out); return; }
    // post-order evaluation of w1 = (1-t)*f1, w2 = t*f2, w1 + w2 (mix.rs:34-44)
    struct Frame { uint32_t mat; int state; d3 f1; double p1; };
    Frame st[kMixDepth]; int sp = 0;
    st[0].mat = mat; st[0].state = 0; sp = 1;
    d3 rf = mk(0, 0, 0); double rp = 0.0;  // result of the most recently completed subtree
    while (sp > 0) {
        Frame& fr = st[sp - 1];
        const DMaterial& m = S.materials[fr.mat];
        if (m.kind != PT_MAT_MIX) { bsdf_eval_pdf_leaf(S, m, view_dir, light_dir, h, rf, rp); sp--; continue; }
        if (fr.state == 0) { fr.state = 1; if (sp < kMixDepth) { st[sp].mat = m.mix_a; st[sp].state = 0; sp++; } else { rf = mk(0, 0, 0); rp = 0.0; } continue; }
        if (fr.state == 1) { fr.f1 = rf; fr.p1 = rp; fr.state = 2; if (sp < kMixDepth) { st[sp].mat = m.mix_b; st[sp].state = 0; sp++; } else { rf = mk(0, 0, 0); rp = 0.0; } continue; }
        double t = m.p[PT_P_MIX_T];
        d3 w1 = (1.0 - t) * fr.f1, w2 = t * rf;
        double p1 = (1.0 - t) * fr.p1, p2 = t * rp;
        rf = w1 + w2; rp = p1 + p2;
        sp--;
    }
    f_out = rf; pdf_out = rp;
}
PT_D d3 bsdf_emitted(const DScene& S, uint32_t mat, double u, double v, d3 p) {  // only DiffuseLight overrides emitted()
    const DMaterial& m = S.materials[mat];
    if (m.kind == PT_MAT_LIGHT) return texture_value(S, m.base_color_tex, u, v, p);
    return mk(0, 0, 0);
}

// ---------------------------------------------------------------- World.lights.{sample,pdf} (list.rs:78-96)
// Hittable::sample for one non-instance object (sphere.rs:110-121, quad.rs:80-86, cuboid.rs:74-76 -> list.rs:78-84,
// mesh.rs:122-129,213-215).  Uniform picks follow the RNG contract: index = min(floor(U*n), n-1).
template <bool GENERAL> PT_D bool light_sample_object(const DScene& S, uint32_t kind, uint32_t index, d3 origin, double time, Rng& rng, d3& dir) {
    if (GENERAL && kind == PT_OBJ_CUBOID) {  // sides.sample: one of the six quads
        uint32_t i = (uint32_t)(rng.next() * 6.0); if (i > 5) i = 5;
        kind = PT_PRIM_QUAD; index = S.cuboids[index].first_quad + i;
    }
    if (kind == PT_PRIM_QUAD) {
        const DQuad& q = S.quads[index];
        double a = rng.next(), b = rng.next();
        d3 point = mk(q.q[0], q.q[1], q.q[2]) + mk(q.u[0], q.u[1], q.u[2]) * a + mk(q.v[0], q.v[1], q.v[2]) * b;
        dir = normalize(point - origin);
        return true;
    }
    if (kind == PT_PRIM_SPHERE) {
        const DSphere& s = S.spheres[index];
        double u = rng.next(), v = rng.next();
        double theta = 2.0 * kPi * u;
        double phi = acos(2.0 * v - 1.0);
        double sph, cph, sth, cth; pt_sincos(phi, sph, cph); pt_sincos(theta, sth, cth);
        double x = sph * cth, y = sph * sth, z = cph;
        d3 p1 = mk(s.p1[0], s.p1[1], s.p1[2]), p2 = mk(s.p2[0], s.p2[1], s.p2[2]);
        d3 point = (p1 + (p2 - p1) * time) + mk(x, y, z) * fmax(s.radius, 0.0);
        dir = normalize(point - origin);
        return true;
    }
    if (GENERAL && kind == PT_OBJ_MESH) {  // triangles.sample: a uniformly chosen triangle, then mesh.rs:122-129
        const DMesh& m = S.meshes[index];
        if (m.n_tri == 0 || !S.tri_verts) return false;
        uint32_t i = (uint32_t)(rng.next() * (double)m.n_tri); if (i >= m.n_tri) i = m.n_tri - 1;
        const double* tv = S.tri_verts + 9ull * (m.first_tri + i);
        double u = rng.next(), v = rng.next();
        double w = 1.0 - u - v;  // not area-uniform (w may be negative): kept as the reference has it
        d3 point = mk(tv[0], tv[1], tv[2]) * w + mk(tv[3], tv[4], tv[5]) * u + mk(tv[6], tv[7], tv[8]) * v;
        dir = normalize(point - origin);
        return true;
    }
    return false;
}
// Hittable::pdf for one non-instance object (sphere.rs:123-135, quad.rs:88-98, mesh.rs:131-141; lists average, list.rs:86-96).
template <bool GENERAL> PT_D double light_pdf_prim(const DScene& S, uint32_t kind, uint32_t index, d3 origin, d3 direction, double time) {
    RayD ray = make_ray(origin, direction, time);  // Ray::new re-normalises (quad.rs:89, sphere.rs:125, mesh.rs:132)
    if (kind == PT_PRIM_QUAD) {
        const DQuad& q = S.quads[index];
        double t, a, b;
        if (!quad_t(q, ray, 0.0, t, a, b)) return 0.0;
        HitInfoD h;
        const d3 un = mk(q.un[0], q.un[1], q.un[2]);
        finish_hit(S, ray, ray_at(ray, t), mk(q.n[0], q.n[1], q.n[2]), t, S.quad_material[index], a, b, h, &un);
        double area = q.area;  // |u x v|, derived on upload
        double cos_theta = fabs(dot(ray.d, h.sn));  // shading normal (Q9)
        return (t * t) / (cos_theta * area);
    }
    if (kind == PT_PRIM_SPHERE) {
        const DSphere& s = S.spheres[index];
        double t;
        if (!sphere_t(s, ray, 0.0, t) || !(t < __longlong_as_double(0x7ff0000000000000ll))) return 0.0;
        double rad = fmax(s.radius, 0.0);
        double r2 = rad * rad;
        d3 p1 = mk(s.p1[0], s.p1[1], s.p1[2]), p2 = mk(s.p2[0], s.p2[1], s.p2[2]);
        d3 c = p1 + (p2 - p1) * time;
        d3 dc = c - origin;
        double solid_angle = 2.0 * kPi * sqrt(1.0 - r2 / dot(dc, dc));
        return 1.0 / solid_angle;
    }
    if (GENERAL && kind == PT_PRIM_TRIANGLE) {
        const DTri& tr = S.tris[index];
        double t, u, v;
        if (!tri_t(tr, ray, 0.0, t, u, v)) return 0.0;
        HitInfoD h;
        reconstruct_hit<true>(S, ray, ref_pack(PT_PRIM_TRIANGLE, index), kInstNone, t, h);
        double area = 0.5 * length(cross(mk(tr.e1[0], tr.e1[1], tr.e1[2]), mk(tr.e2[0], tr.e2[1], tr.e2[2])));  // mesh.rs:42-46
        double cos_theta = fabs(dot(direction, h.sn));  // the un-normalised `direction` (mesh.rs:136)
        return t * t / (cos_theta * area);
    }
    return 0.0;
}
template <bool GENERAL> PT_D double light_pdf_object(const DScene& S, uint32_t kind, uint32_t index, d3 origin, d3 direction, double time) {
    if (GENERAL && kind == PT_OBJ_CUBOID) {
        const uint32_t fq = S.cuboids[index].first_quad;
        double s = 0.0;
        for (uint32_t k = 0; k < 6; k++) s += light_pdf_prim<GENERAL>(S, PT_PRIM_QUAD, fq + k, origin, direction, time);
        return s / 6.0;
    }
    if (GENERAL && kind == PT_OBJ_MESH) {  // O(triangles) per evaluation, exactly like the reference's list average
        const DMesh& m = S.meshes[index];
        if (m.n_tri == 0) return 0.0;
        double s = 0.0;
        for (uint32_t k = 0; k < m.n_tri; k++) s += light_pdf_prim<GENERAL>(S, PT_PRIM_TRIANGLE, m.first_tri + k, origin, direction, time);
        return s / (double)m.n_tri;
    }
    return light_pdf_prim<GENERAL>(S, kind, index, origin, direction, time);
}
template <bool GENERAL> PT_D bool light_sample_one(const DScene& S, DRef rf, d3 origin, double time, Rng& rng, d3& dir) {
    const uint32_t kind = ref_kind(rf.kind_index), index = ref_index(rf.kind_index);
    if (!GENERAL || kind != PT_OBJ_INSTANCE) return light_sample_object<GENERAL>(S, kind, index, origin, time, rng, dir);
    const DInstance& in = S.instances[index];  // instance.rs:64-69
    d3 local;
    if (!light_sample_object<GENERAL>(S, in.child_kind, in.child_index, xform_point(in.inv, origin), time, rng, local)) return false;
    dir = xform_vector(in.fwd, local);
    return true;
}
template <bool GENERAL> PT_D double light_pdf_one(const DScene& S, DRef rf, d3 origin, d3 direction, double time) {
    const uint32_t kind = ref_kind(rf.kind_index), index = ref_index(rf.kind_index);
    if (!GENERAL || kind != PT_OBJ_INSTANCE) return light_pdf_object<GENERAL>(S, kind, index, origin, direction, time);
    const DInstance& in = S.instances[index];  // instance.rs:71-75
    return light_pdf_object<GENERAL>(S, in.child_kind, in.child_index, xform_point(in.inv, origin), xform_vector(in.inv, direction), time);
}
template <bool GENERAL = true> PT_D bool lights_sample(const DScene& S, d3 origin, double time, Rng& rng, d3& dir) {  // list.rs:78-84
    if (S.n_lights == 0) return false;
    uint32_t n = S.n_lights;
    uint32_t i = (uint32_t)(rng.next() * (double)n);  // RNG contract: index = min(floor(U*n), n-1)
    if (i >= n) i = n - 1;
    return light_sample_one<GENERAL>(S, S.lights[i], origin, time, rng, dir);
}
template <bool GENERAL = true> PT_D double lights_pdf(const DScene& S, d3 origin, d3 direction, double time) {  // list.rs:86-96
    if (S.n_lights == 0) return 0.0;
    double s = 0.0;
    for (uint32_t i = 0; i < S.n_lights; i++) s += light_pdf_one<GENERAL>(S, S.lights[i], origin, direction, time);
    return s / (double)S.n_lights;
}

}  // namespace ptd
