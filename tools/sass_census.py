"""Per-kernel SASS opcode census of the shipped library (no GPU needed): which kernels use the 5th-generation
tensor cores (UTC*MMA), TMEM loads / stores (LDTM / STTM), TMA-engine bulk copies (UBLKCP), mbarriers (SYNCS),
FP64 / legacy tensor MMA (DMMA / HMMA), cp.async (LDGSTS).  Writes a markdown table (committed under profiles/)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "srcfinder_b200", "libcmf_b200.so")
OPS = ["UTCHMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "HMMA", "LDGSTS", "DFMA",
       "FFMA", "FFMA2", "LDG", "STG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fn[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            fn[cur][op] += 1
            fn[cur]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(fn), capture_output=True, text=True).stdout.splitlines()
    groups = collections.OrderedDict()
    for mangled, name in zip(fn, demangle):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("cmf::", ""))
        short = re.sub(r"^void ", "", short)
        base = re.sub(r"<.*", "", short)
        g = groups.setdefault(base, {"n": 0, "ops": collections.Counter(), "ex": short})
        g["n"] += 1
        g["ops"].update(fn[mangled])
    out = ["# SASS opcode census of srcfinder_b200/libcmf_b200.so (sm_100a), per kernel (summed over template instances)", "",
           "| kernel | instances | instructions | " + " | ".join(OPS) + " |", "|---|---|---|" + "---|" * len(OPS)]
    for base, g in sorted(groups.items(), key=lambda kv: -kv[1]["ops"]["_total"]):
        out.append("| `%s` | %d | %d | %s |" % (base, g["n"], g["ops"]["_total"],
                                               " | ".join(str(g["ops"][o]) if g["ops"][o] else "" for o in OPS)))
    text = "\n".join(out) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_census.md")
    with open(dst, "w") as fh:
        fh.write(text)
    print(text)


if __name__ == "__main__":
    main()
