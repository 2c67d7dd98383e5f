#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/collect_evidence.sh).
usage: launch_summary.py LAUNCHES.csv "title" > summary.md"""
import csv, re, sys, collections

def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi: continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("ptd::", "")
        name = re.sub(r"\((int|bool)\)", "", name)
        v = float(r[vi].replace(",", "")); u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        # the counting variants (profiling level 2: k_top<_, 1>, k_mesh_*<1>, k_trace<_, _, 1, ...>) belong to bench.py's counter pass, not to a step
        if re.match(r"k_top<\d, 1>|k_mesh_(walk|enter)<1>|k_trace<\d, \d, 1", name): name = "(counting variants of the profiling pass, excluded from the shares)"
        tot[name][0] += 1; tot[name][1] += us
    excl = tot.pop("(counting variants of the profiling pass, excluded from the shares)", None)
    total = sum(v[1] for v in tot.values())
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised, so only the SHARES are meaningful\n"
          "(small launches are inflated most: every kernel starts with an empty L2).\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, us) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {us:.0f} | {100 * us / total:.1f} % |")
    if excl: print(f"| (counting variants of bench.py's counter pass, not part of a step) | {excl[0]} | {excl[1]:.0f} | excluded |")
    def share(pred): return 100 * sum(v[1] for k, v in tot.items() if pred(k)) / total
    trav = share(lambda k: k.startswith(("k_top", "k_mesh", "k_trace", "k_generate")))
    print(f"\nBy stage: traversal (k_top + k_mesh_enter + k_mesh_walk + k_trace) {trav:.1f} %, k_shade<*> {share(lambda k: k.startswith('k_shade')):.1f} %, "
          f"k_tail {share(lambda k: k.startswith('k_tail')):.1f} %, other {share(lambda k: not k.startswith(('k_top', 'k_mesh', 'k_trace', 'k_generate', 'k_shade', 'k_tail'))):.1f} %.")

if __name__ == "__main__":
    main()
