"""Opcode histogram per kernel of the in-tree libmpcb200.so (cuobjdump -sass), written as markdown.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpc-rl_for_avs_b200", "libmpcb200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kern = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kern[cur][m.group(1).split(".")[0]] += 1
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.split("\n")
print("# SASS summary of libmpcb200.so (round 2)\n")
print(f"`cuobjdump -sass` of the in-tree library; architectures present: {', '.join(archs)} (sm_100a only: `-gencode arch=compute_100a,code=sm_100a`).")
print("This path has no dense contraction, so there is no `UTC*MMA` (tcgen05.mma) by design; tensor memory is used as a per-thread")
print("scratchpad for the feedback gains: `LDTM` / `STTM` (tcgen05.ld / st) and `UTCATOMSWS` (tcgen05.alloc / dealloc) in `k_solve_tmem`.\n")
print("| kernel | regs | stack B | static smem B | instructions | FFMA | FMUL+FADD | MUFU | LDS | STS | LDTM | STTM | UTCATOMSWS | DFMA+DMUL+DADD | BAR | top opcodes |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, c), dn in zip(kern.items(), demangle):
    short = re.sub(r"\(.*", "", dn).replace("mpcb::", "")
    if not any(k in short for k in ("k_solve", "k_prepare", "k_select", "k_rollout", "k_fma", "k_env")):
        continue
    if "k_solve" in short and not any(t in short for t in ("<256>", "<192>", "<128>")):
        continue
    u = usage.get(name, ("?", "?", "?"))
    tot = sum(c.values())
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(6))
    print(f"| `{short}` | {u[0]} | {u[1]} | {u[2]} | {tot} | {c['FFMA']} | {c['FMUL'] + c['FADD']} | {c['MUFU']} | {c['LDS']} | {c['STS']} | {c['LDTM']} | {c['STTM']} | "
          f"{c['UTCATOMSWS']} | {c['DFMA'] + c['DMUL'] + c['DADD']} | {c['BAR']} | {top} |")
