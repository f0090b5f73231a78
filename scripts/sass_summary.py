"""Counts the Blackwell-specific SASS opcodes per kernel of the built objects (cuobjdump -sass):
tcgen05 MMA (UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMA bulk copies (UBLKCP), tensor-core and
transaction barriers (UTCBAR, SYNCS), cp.async (LDGSTS), FP64 FMA / MMA (DFMA, DMMA).

  python scripts/sass_summary.py > profiles/r02_sass_opcodes.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "LDGSTS",
       "DFMA", "DMUL", "DADD", "DMMA", "MUFU.RSQ64H", "RED", "ATOM", "BAR.SYNC", "LDS", "SHFL", "ST.E", "LD.E", "LDG", "STG"]


def main():
    print("# cuobjdump -sass of cge_jl_b200/csrc/build/*.o (sm_100a): opcode counts per kernel (static SASS)")
    print("# kernel".ljust(74) + " ".join(o.rjust(11) for o in OPS if o not in ("LD.E", "ST.E")))
    for obj in sorted(glob.glob(os.path.join(ROOT, "cge_jl_b200", "csrc", "build", "*.o"))):
        if re.search(r"cge_inst[1-7]\.o$", obj):
            continue  # the same kernels for other exponents
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, counts = None, collections.OrderedDict()
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = name.replace("(anonymous namespace)::", "")
                name = re.sub(r"\(.*", "", name).replace("void ", "").replace("cge::", "")
                counts[name] = collections.Counter()
                continue
            if name is None:
                continue
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                for o in OPS:
                    if op == o or op.startswith(o + ".") or (o == "MUFU.RSQ64H" and op.startswith(o)):
                        counts[name][o] += 1
        shown = 0
        for k, c in counts.items():
            if not any(c.values()):
                continue
            if os.path.basename(obj) == "cge_inst0.o" and not re.search(r"<1,|<\(int\)1,|k_bfp<2>", k):
                continue  # one exponent is enough
            if shown == 0:
                print(f"## {os.path.basename(obj)}")
            shown += 1
            print(k[:72].ljust(74) + " ".join(str(c.get(o, 0)).rjust(11) for o in OPS if o not in ("LD.E", "ST.E")))


if __name__ == "__main__":
    main()
