#!/usr/bin/env python3
"""Developer diagnostic (GPU box): device vs oracle on closest-hit batches, BSDF batches and small renders,
printing mismatch details.  The pytest suite (tests/ -m gpu) asserts the same things."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pt = ge.load_package()
orc = ge.load_oracle()


def ulp_diff(a, b):
    ai = a.view(np.int64)
    bi = b.view(np.int64)
    return np.abs(ai - bi)


def check_trace(name, dev, ora, rays, t_min=1e-3):
    t0 = time.time(); a = dev.trace_closest(rays, t_min); t1 = time.time(); b = ora.trace_closest(rays, t_min); t2 = time.time()
    bad = (a["hit"] != b["hit"])
    both = (a["hit"] == 1) & (b["hit"] == 1)
    for f in ["prim_kind", "prim_index", "instance", "is_light", "material", "front_face"]:
        bad |= both & (a[f] != b[f])
    tu = ulp_diff(a["t"][both], b["t"][both])
    pe = np.abs(a["point"][both] - b["point"][both]).max() if both.any() else 0
    ne = np.abs(a["geometric_normal"][both] - b["geometric_normal"][both]).max() if both.any() else 0
    se = np.abs(a["shading_normal"][both] - b["shading_normal"][both]).max() if both.any() else 0
    ue = max(np.abs(a["u"][both] - b["u"][both]).max(), np.abs(a["v"][both] - b["v"][both]).max()) if both.any() else 0
    print(f"[trace {name}] n={len(rays)} hits={int(b['hit'].sum())} id-mismatch={int(bad.sum())} t-ulp-max={int(tu.max()) if tu.size else 0} "
          f"t-neq={int((tu > 0).sum())} point-err={pe:.2e} gn-err={ne:.2e} sn-err={se:.2e} uv-err={ue:.2e} dev {t1 - t0:.3f}s orc {t2 - t1:.3f}s")
    if bad.any():
        i = np.nonzero(bad)[0][:5]
        for k in i:
            print("   ray", k, rays[k], "\n     dev", a[k], "\n     orc", b[k])
    return int(bad.sum()), int(tu.max()) if tu.size else 0


def rand_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def check_bsdf(name, dev, ora, n_materials, rng, n=4096):
    worst = 0
    for m in range(n_materials):
        q = np.zeros(n, dtype=pt.BSDF_QUERY_DTYPE)
        gn = rand_dirs(rng, n)
        sn = gn + 0.2 * rand_dirs(rng, n); sn /= np.linalg.norm(sn, axis=1, keepdims=True)
        q["geometric_normal"], q["shading_normal"] = gn, sn
        q["view_dir"], q["light_dir"] = rand_dirs(rng, n), rand_dirs(rng, n)
        q["point"] = rng.uniform(-5, 5, size=(n, 3)); q["u"] = rng.uniform(size=n); q["v"] = rng.uniform(size=n)
        q["front_face"] = rng.integers(0, 2, size=n)
        a, b = dev.bsdf_eval_pdf(m, q), ora.bsdf_eval_pdf(m, q)
        def rel(x, y):
            d = np.abs(x - y); s = np.maximum(np.abs(y), 1e-300)
            ok = np.isfinite(x) & np.isfinite(y)
            same_nonfinite = (~ok) & ((np.isnan(x) & np.isnan(y)) | (x == y))
            r = np.where(ok, d / s, np.where(same_nonfinite, 0.0, np.inf))
            r = np.where(ok & (d < 1e-300), 0.0, r)
            return r.max()
        e = max(rel(a["eval"], b["eval"]), rel(a["pdf"], b["pdf"]), rel(a["emitted"], b["emitted"]))
        u8 = rng.uniform(size=(n, 8))
        sa, sb = dev.bsdf_sample(m, q, u8), ora.bsdf_sample(m, q, u8)
        vm = int((sa["valid"] != sb["valid"]).sum()); um = int((sa["n_uniforms"] != sb["n_uniforms"]).sum())
        both = (sa["valid"] == 1) & (sb["valid"] == 1)
        de = np.abs(sa["dir"][both] - sb["dir"][both]).max() if both.any() else 0
        print(f"[bsdf {name}] mat {m}: eval/pdf max rel err {e:.2e}; sample valid-mismatch {vm} uniforms-mismatch {um} dir-err {de:.2e}")
        worst = max(worst, e)
    return worst


def main():
    rng = np.random.default_rng(1)
    ctx = pt.Context(0)
    scenes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["3", "1", "7", "6"])]
    for sid in scenes:
        scene = pt.Scene.build(sid, width=160, spp=4, seed=1)
        dev = ctx.upload(scene)
        t0 = time.time(); ora = orc.OracleScene(scene.desc, pt); print(f"scene {sid}: oracle build {time.time() - t0:.1f}s, device bytes {dev.device_bytes}")
        h = scene.image_height(); w = scene.camera.image_width
        rows, cols = np.divmod(np.arange(w * h, dtype=np.uint32), w)
        cr_o = orc.camera_rays(scene.camera, 7, rows, cols, np.zeros_like(rows), pt)
        cr_d = pt.camera_rays(ctx, scene.camera, 7, rows, cols, np.zeros_like(rows))
        print(f"[camera {sid}] origin err {np.abs(cr_o['origin'] - cr_d['origin']).max():.2e} dir err {np.abs(cr_o['direction'] - cr_d['direction']).max():.2e} time err {np.abs(cr_o['time'] - cr_d['time']).max():.2e}")
        check_trace(f"s{sid} camera", dev, ora, cr_o)
        pr = ora.dump_path_rays(scene.camera, 11, 3, 2, 1, 200000)
        check_trace(f"s{sid} bounce", dev, ora, pr)
        check_bsdf(f"s{sid}", dev, ora, _n_materials(scene), rng, 2048)
        for policy in (pt.PT_NAN_DROP, pt.PT_NAN_REFERENCE):
            t0 = time.time(); img, st = dev.render(spp=4, seed=5, nan_policy=policy); t1 = time.time()
            ref, ost = ora.render(scene.camera, 4, seed=5, nan_policy=policy); t2 = time.time()
            d = np.abs(img - ref); fin = np.isfinite(ref) & np.isfinite(img)
            print(f"[render s{sid} policy {policy}] dev {t1 - t0:.3f}s ({st.device_ms:.1f} ms, {st.segments / st.device_ms / 1e3:.1f} Mrays/s, iters {st.iterations}) orc {t2 - t1:.2f}s; "
                  f"segments dev {st.segments} orc {ost.segments}; nonfinite dev {st.nonfinite} orc {ost.nonfinite}; "
                  f"mean ref {ref[fin].mean():.4f} mean|diff| {d[fin].mean():.3e} max {d[fin].max():.3e} px>1e-3: {int((d.max(axis=2) > 1e-3).sum())}/{w * h}")
        dev.close(); ora.close()
    ctx.close()


def _n_materials(scene):
    import ctypes as C
    # n_materials is the 4th uint32 of pt_scene_desc
    return C.cast(scene.desc, C.POINTER(C.c_uint32))[3]


if __name__ == "__main__":
    main()
