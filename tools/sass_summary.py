"""Static summary of the built library (no GPU needed): per kernel the registers / stack / static shared memory that
ptxas assigned (cuobjdump --dump-resource-usage) and the count of the SASS mnemonics that tell which hardware path a
kernel uses (B200_PROFILING.md: UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
UBLKCP / UTMALDG = bulk / tensor TMA copies, LDGSTS = cp.async, SYNCS = mbarrier, FFMA2 = packed FP32).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ad-yolo_b200", "lib", "libadyolo_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS",
             "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "HFMA2", "DFMA", "MUFU", "I2F", "LDS", "STS", "LDG", "STG",
             "ATOMS", "ATOMG", "RED", "SHFL", "BAR", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*", "", name)                      # drop the argument list
    return name.replace("void ", "").replace("ady::", "")


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first (python -c 'import __graft_entry__ as g; g.build()')")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage, arch = {}, set()
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, total = collections.defaultdict(collections.Counter), collections.Counter()
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            for mn in MNEMONICS:
                if op == mn:
                    counts[fn][mn] += 1
    names = demangle(sorted(usage))
    print(f"# {os.path.relpath(LIB, ROOT)}: cubin architectures {sorted(arch)}; {len(usage)} kernels")
    print("# columns: registers / stack bytes / static shared bytes (dynamic shared memory is set at launch) / SASS instructions,")
    print("# then the non-zero counts of the mnemonics listed in tools/sass_summary.py (static counts, not executed ones)")
    for mangled in sorted(usage, key=lambda k: names[k]):
        u = usage[mangled]
        c = counts[mangled]
        mix = " ".join(f"{k}={c[k]}" for k in MNEMONICS if c[k])
        print(f"{short(names[mangled])}\n    REG {u.get('REG')}  STACK {u.get('STACK')}  SHARED {u.get('SHARED')}  SASS {total[mangled]}  | {mix}")
    tc = sorted({short(names[k]) for k in usage if counts[k]["UTCHMMA"] or counts[k]["UTCQMMA"]})
    tma = sorted({short(names[k]) for k in usage if counts[k]["UBLKCP"] or counts[k]["UTMALDG"]})
    print(f"# kernels issuing tcgen05.mma: {tc or 'none'}")
    print(f"# kernels issuing TMA copies (UBLKCP / UTMALDG): {tma or 'none'}")


if __name__ == "__main__":
    main()
