#!/usr/bin/env python3
"""Joins an `ncu --page source --csv` export (per-SASS-instruction stall samples) with `nvdisasm --print-line-info` of the SAME
build, so that samples and executed instructions can be summed per source line of the kernels.
usage: ncu_source_lines.py SRC.csv[.gz] DISASM.txt KERNEL_SUBSTRING [launch_index] [top_n]
The join is by instruction index inside the function; the opcode of every row is checked against the disassembly."""
import csv, gzip, re, sys, collections

def load_sections(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    out = []
    for k, s in enumerate(idx):
        e = idx[k + 1] if k + 1 < len(idx) else len(rows)
        out.append((rows[s][1], rows[s + 1], [r for r in rows[s + 2:e] if len(r) > 5]))
    return out

def load_disasm(path, mangled_sub):
    cur = None; line = ("?", 0); out = []
    for l in open(path):
        if l.startswith("//----") and ".text." in l:
            cur = l.split(".text.")[1].split()[0]
            continue
        if cur is None or mangled_sub not in cur: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: out.append((int(m.group(1), 16), m.group(2).strip(), line))
    return out

def main():
    src, dis, ksub = sys.argv[1:4]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    mang = sys.argv[6] if len(sys.argv) > 6 else ksub
    secs = [s for s in load_sections(src) if ksub in s[0]]
    name, hdr, rows = secs[launch]
    ci = {n: i for i, n in enumerate(hdr)}
    d = load_disasm(dis, mang)
    assert len(d) >= len(rows), (len(d), len(rows))
    bad = 0
    agg = collections.defaultdict(lambda: [0, 0, 0.0, collections.Counter()])
    tot_s = tot_i = 0
    for k, r in enumerate(rows):
        sass = r[ci["Source"]].strip()
        opc = sass.split()[1] if sass.startswith("@") else sass.split()[0]
        dop = d[k][1].split()[1] if d[k][1].startswith("@") else d[k][1].split()[0]
        if opc.split(".")[0] != dop.split(".")[0]: bad += 1
        s = int(r[ci["# Samples"]]); ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]])
        a = agg[d[k][2]]; a[0] += s; a[1] += ie; a[2] += te
        for st in ("stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_barrier", "stall_lg", "stall_not_selected", "stall_dispatch", "stall_selected"):
            a[3][st] += int(r[ci[st]] or 0)
        tot_s += s; tot_i += ie
    print(f"# {name[:80]}  rows {len(rows)}  opcode mismatches {bad}  samples {tot_s}  warp instr {tot_i}")
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, l), a in agg.items(): byfile[f][0] += a[0]; byfile[f][1] += a[1]
    for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]): print(f"  file {f:24s} samples {100 * a[0] / tot_s:5.1f}%  instr {100 * a[1] / tot_i:5.1f}%")
    allst = collections.Counter()
    for a in agg.values(): allst.update(a[3])
    print("  stalls:", " ".join(f"{k[6:]} {100 * v / max(1, sum(allst.values())):.1f}%" for k, v in allst.most_common()))
    for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
        top = " ".join(f"{k[6:]} {v}" for k, v in a[3].most_common(3) if v)
        print(f"  {f}:{l:<5d} samples {100 * a[0] / tot_s:5.2f}%  instr {100 * a[1] / tot_i:5.2f}%  lanes {a[2] / max(1, a[1]):4.1f}  [{top}]")

if __name__ == "__main__":
    main()
