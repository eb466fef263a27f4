"""CPU-only: per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA use in the shipped
library (`cuobjdump -sass ctrl-v_b200/libctrlv_b200.so`).  Writes profiles/r02_sass_opcodes.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ctrl-v_b200", "libctrlv_b200.so")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCATOM", "MUFU.EX2", "MUFU.TANH",
        "FFMA2", "FMUL2", "FADD2", "STG.E.ENL2.256", "LDG.E.ENL2.256", "SYNCS", "ELECT", "USETMAXREG", "ATOMG", "REDG", "RED."]


def main(out_path):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    lines = ["# cuobjdump -sass ctrl-v_b200/libctrlv_b200.so: per-kernel opcode counts (sm_100a); scripts/sass_opcodes.py",
             "# (first matching prefix per instruction; UTCHMMA.2CTA is counted apart from UTCHMMA)"]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        except OSError:
            pass
        c = collections.Counter()
        for line in f.split("\n"):
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                for p in PATS:
                    if m.group(2).startswith(p):
                        c[p] += 1
                        break
        if c:
            lines.append(name.split("(")[0] + ": " + ", ".join(f"{k}={v}" for k, v in sorted(c.items())))
    with open(out_path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt"))
