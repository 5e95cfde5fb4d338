"""Count the Blackwell-native SASS mnemonics per kernel of libcddpm_b200.so (cuobjdump -sass): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA loads), HMMA (legacy mma.sync), plus registers from --dump-resource-usage.
   python tools/sass_summary.py > profiles/<tag>_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "conditioned-diffusion-models-uad_b200", "lib", "libcddpm_b200.so")
PATS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "DMMA", "REDG", "ATOMG", "ACQBULK"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["INSTR"] += 1
            for p in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "DMMA"):
                if op.startswith(p):
                    counts[cur][p] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                counts[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in op:
                counts[cur]["UTMALDG.2CTA"] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?STACK:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    print("| kernel | instr | UTCHMMA (of which .2CTA) | LDTM | UTMALDG (.2CTA) | HMMA (mma.sync) | regs | stack |")
    print("|---|---|---|---|---|---|---|---|")
    tot = collections.Counter()
    for fn, c in counts.items():
        tot.update(c)
        if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["HMMA"]) and "--all" not in sys.argv:
            continue
        name = demangle(fn).replace("cddpm::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name.replace("void ", ""))
        r = regs.get(fn, ("?", "?", "?"))
        print(f"| `{name[:70]}` | {c['INSTR']} | {c['UTCHMMA']} ({c['UTCHMMA.2CTA']}) | {c['LDTM']} | "
              f"{c['UTMALDG']} ({c['UTMALDG.2CTA']}) | {c['HMMA']} | {r[0]} | {r[1]} |")
    print(f"\nTotals over {len(counts)} kernels: UTCHMMA {tot['UTCHMMA']} (.2CTA {tot['UTCHMMA.2CTA']}), LDTM {tot['LDTM']}, "
          f"UTMALDG {tot['UTMALDG']} (.2CTA {tot['UTMALDG.2CTA']}), HMMA {tot['HMMA']}.")


if __name__ == "__main__":
    main()
