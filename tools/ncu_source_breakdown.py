"""Executed-instruction share per source function of one kernel: joins `ncu --page source --csv` (SASS view, per
instruction "Instructions Executed") with the line table of the same binary (`nvdisasm -g`), by instruction offset.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    python tools/ncu_source_breakdown.py sass.csv <lib.so> <kernel mangled-name substring> [core source file]
"""
import bisect, collections, csv, os, re, subprocess, sys, tempfile

sass_csv, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
srcfile = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(lib), "csrc", "mpc_core.cuh")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    start = next((i for i, l in enumerate(txt) if l.startswith(".text.") and kname in l), None)
    if start is None:
        continue
    cur = None
    for l in txt[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    break
rows = list(csv.reader(open(sass_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
names = rows[hdr]
ia, ie, it = names.index("Address"), names.index("Instructions Executed"), names.index("Thread Instructions Executed")
data = [r for r in rows[hdr + 1:] if len(r) > it and r[ia].startswith("0x")]
base = int(data[0][ia], 16)
src = open(srcfile).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:MPC_HD|MPC_NOINLINE|__device__ __forceinline__)\s+[\w<>:, \*&]+?\s+(\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
starts = [f[0] for f in funcs]
agg_w, agg_t = collections.Counter(), collections.Counter()
for r in data:
    off = int(r[ia], 16) - base
    key = line_of.get(off)
    w, t = int(r[ie] or 0), int(r[it] or 0)
    if key is None:
        name = "?"
    elif key[0] == os.path.basename(srcfile):
        j = bisect.bisect_right(starts, key[1]) - 1
        name = funcs[j][1] if j >= 0 else "?"
    else:
        name = key[0]
    agg_w[name] += w
    agg_t[name] += t
tot_w, tot_t = sum(agg_w.values()), sum(agg_t.values())
print(f"| source function | warp instructions | share | avg active lanes |\n|---|---|---|---|")
for k, v in agg_w.most_common(22):
    print(f"| `{k}` | {v:.3e} | {100 * v / tot_w:.1f} % | {agg_t[k] / max(v, 1):.1f} |")
print(f"| total | {tot_w:.3e} | 100 % | {tot_t / tot_w:.1f} |")
