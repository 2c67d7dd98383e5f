#!/usr/bin/env python3
"""Developer probe (GPU box): histograms of the traversal work per ray (top-level pass) and per mesh visit (mesh rounds).
usage: hist_probe.py scene:width:spp ..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pt = ge.load_package()
NAMES = ["top: node fetches / ray", "top: ref boxes / ray", "top: f64 tests / ray", "top: mesh visits queued / ray",
         "mesh: wide nodes / visit", "mesh: ref boxes / visit", "mesh: f64 tri tests / visit", "mesh: [dropped, walked, improved]"]


def main():
    ctx = pt.Context(0)
    for spec in sys.argv[1:]:
        sid, width, spp = [int(x) for x in spec.split(":")[:3]]
        scene = pt.Scene.build(sid, width=width, spp=spp, seed=1)
        dev = ctx.upload(scene)
        ctx.set_profiling(2)
        ctx.histograms(reset=True)
        img, st = dev.render(spp=spp, seed=2, nan_policy=pt.PT_NAN_DROP)
        h = ctx.histograms()
        print(f"== scene {sid} {width}px {spp}spp: {st.segments} segments, {st.paths} paths, two-pass iterations {st.two_pass_iterations}")
        for k, name in enumerate(NAMES):
            row = h[k].astype(np.float64)
            tot = row.sum()
            if tot == 0:
                continue
            mean = (row * np.arange(64)).sum() / tot
            cdf = np.cumsum(row) / tot
            pct = {q: int(np.searchsorted(cdf, q)) for q in (0.5, 0.9, 0.99)}
            head = " ".join(f"{int(i)}:{row[i] / tot:.3f}" for i in range(16) if row[i])
            print(f"  {name}: n={int(tot)} mean {mean:.2f} p50 {pct[0.5]} p90 {pct[0.9]} p99 {pct[0.99]} | {head}")
        dev.close()
    ctx.close()


if __name__ == "__main__":
    main()
