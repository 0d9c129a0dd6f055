"""Static instruction mix of every kernel in libraym0nade_b200.so, from `cuobjdump -sass` (no GPU needed):
global loads by width, warp votes / shuffles, atomics, local-memory traffic (spills + the traversal stack's overflow),
barriers, and any tensor-core instruction (there must be none: nothing on this path is a dense contraction).

    python scripts/sass_mix.py [lib.so] > profiles/rNN_sass_mix.txt

tests/test_cpu_host.py::test_sass_has_the_fetch_widths_and_warp_primitives_the_design_states reads the same table.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["inst", "LDG.32", "LDG.64", "LDG.128", "LDG.256", "STG", "LDS", "STS", "LDL", "STL", "VOTE", "SHFL", "MATCH", "ATOM/RED", "BAR", "MUFU", "TENSOR"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def classify(op):
    """column of one SASS opcode (with its dot suffixes), or None"""
    base = op.split(".")[0]
    if base in ("LDG", "LD"):
        for w in ("256", "128", "64"):
            if "." + w in op:
                return "LDG." + w
        return "LDG.32"
    if base in ("STG", "ST"):
        return "STG"
    if base in ("LDS", "LDSM"):
        return "LDS"
    if base == "STS":
        return "STS"
    if base in ("LDL", "STL", "VOTE", "SHFL", "MATCH", "BAR", "MUFU"):
        return base
    if base in ("ATOM", "ATOMG", "ATOMS", "RED", "REDG"):
        return "ATOM/RED"
    if base in ("HMMA", "IMMA", "DMMA", "QMMA", "OMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCOMMA", "UTCMMA", "HGMMA", "IGMMA", "QGMMA"):
        return "TENSOR"
    return None


def mix(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur["inst"] += 1
            c = classify(m.group(1))
            if c:
                cur[c] += 1
    return kernels


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("rm::", "")
    return re.sub(r"\(.*", "", name)                      # drop the parameter list


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raym0nade_b200", "libraym0nade_b200.so")
    k = mix(lib)
    names = demangle(list(k))
    print("static SASS mix per kernel (sm_100a), %s" % os.path.relpath(lib, ROOT))
    print("%-52s" % "kernel" + "".join("%9s" % c for c in COLS))
    for n, c in sorted(k.items(), key=lambda kv: -kv[1]["inst"]):
        print("%-52s" % short(names[n])[:52] + "".join("%9d" % c[col] for col in COLS))


if __name__ == "__main__":
    main()
