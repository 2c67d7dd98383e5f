import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge, numpy as np
pt = ge.load_package(); orc = ge.load_oracle()
ctx = pt.Context(0)
for sid in [6, 1]:
    scene = pt.Scene.build(sid, width=320, spp=4, seed=1)
    dev = ctx.upload(scene); ora = orc.OracleScene(scene.desc, pt)
    h = scene.image_height(); w = scene.camera.image_width
    rows, cols = np.divmod(np.arange(w*h, dtype=np.uint32), w)
    rays = orc.camera_rays(scene.camera, 7, rows, cols, np.zeros_like(rows), pt)
    a = dev.trace_closest(rays)
    pairs = a['work'] & 0xFFFF; prims = a['work'] >> 16
    print(sid, 'camera rays: pairs mean %.1f max %d p99 %d | prims mean %.1f max %d p99 %d' % (pairs.mean(), pairs.max(), np.percentile(pairs,99), prims.mean(), prims.max(), np.percentile(prims,99)))
    img = pairs.reshape(h,w)
    ys, xs = np.unravel_index(np.argsort(pairs)[-5:], (h,w)); print('  worst pixels', list(zip(ys,xs)), pairs[np.argsort(pairs)[-5:]])
    pr = ora.dump_path_rays(scene.camera, 11, 7, 2, 1, 100000)
    b = dev.trace_closest(pr)
    pairs = b['work'] & 0xFFFF; prims = b['work'] >> 16
    print(sid, 'bounce rays: pairs mean %.1f max %d p99 %d | prims mean %.1f max %d p99 %d' % (pairs.mean(), pairs.max(), np.percentile(pairs,99), prims.mean(), prims.max(), np.percentile(prims,99)))
