"""Per-kernel SASS evidence of the shipped library (cuobjdump -sass): which kernels issue DMMA, bulk-TMA
(UBLKCP), mbarrier traffic (SYNCS), setmaxnreg (USETMAXREG), and how many -- runs without a GPU.

usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "metalquicha_b200", "libmqcb200.so")
MNEMONICS = ["DMMA", "UBLKCP", "SYNCS", "USETMAXREG", "DFMA", "LDG", "STG", "LDS", "STS", "BAR", "UTMALDG", "UTCHMMA", "HMMA"]


def _short(nice):
    return re.sub(r"\(.*", "", nice).replace("mqcb200::", "").replace("void ", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for mn in MNEMONICS:
                if op.split(".")[0] == mn:
                    kernels[name][mn] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence, round 2 -- `cuobjdump -sass metalquicha_b200/libmqcb200.so`\n")
    print(f"Architectures in the fatbinary: {', '.join(arch)} (built with `-gencode arch=compute_100a,code=sm_100a`).\n")
    print("FP64 has no tcgen05 kind, so the tensor path of this library is `DMMA.8x8x4` fed by 1-D bulk TMA "
          "(`UBLKCP`) on mbarriers (`SYNCS`), with `USETMAXREG` moving registers from the producer to the consumer "
          "warps; no `UTCHMMA`/`UTMALDG`/`HMMA` is expected anywhere.\n")
    print("| kernel | instructions | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    tot = collections.Counter()
    seen = set()
    for (mangled, c), nice in zip(kernels.items(), demangled):
        short = _short(nice)
        if short in seen:
            continue
        seen.add(short)
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
        tot.update(c)
    print(f"| **total** | {tot['_total']} | " + " | ".join(str(tot[m]) for m in MNEMONICS) + " |")
    spills = []
    bdir = os.path.join(ROOT, "metalquicha_b200", "csrc", "build")
    for f in sorted(os.listdir(bdir)):
        if f.endswith(".ptxas.log"):
            txt = open(os.path.join(bdir, f)).read()
            for fn, body in re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n(.*?)(?=ptxas info    : Compiling|\Z)", txt, re.S):
                m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
                r = re.search(r"Used (\d+) registers", body)
                if m and (int(m.group(2)) or int(m.group(3))):
                    spills.append((fn, r.group(1) if r else "?", m.group(1), m.group(2), m.group(3)))
    print("\n## Register spills reported by `ptxas -v`\n")
    if not spills:
        print("None.")
    else:
        names = subprocess.run(["c++filt"], input="\n".join(s[0] for s in spills), capture_output=True, text=True).stdout.splitlines()
        print("| kernel | registers | stack frame (B) | spill stores (B) | spill loads (B) |\n|---|---|---|---|---|")
        for s, nm in zip(spills, names):
            print(f"| `{_short(nm)}` | {s[1]} | {s[2]} | {s[3]} | {s[4]} |")


if __name__ == "__main__":
    main()
