#!/usr/bin/env python
"""Static opcode census of one kernel from `cuobjdump -sass` (no GPU needed):

    python profiles/sass_census.py [kernel-substring] [library.so]

Complements the dynamic numbers of the ncu captures: which instruction classes the compiled kernel consists of, whether the
asynchronous-copy / mbarrier path is really there (UBLKCP, SYNCS), how many FP64 pipe instructions (DFMA/DMUL/DADD/DSETP) there are
against integer, predicate and shared-memory work, and whether ptxas spilled (STL/LDL)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax-in-cell_b200", "jaxincell_b200", "libjic_b200.so")
CLASSES = [("fp64 pipe", r"^(DFMA|DMUL|DADD|DSETP|DMNMX|MUFU\.RCP64H|F2F|I2F\.F64|F2I\.F64|FRND\.F64)"),
           ("shared memory", r"^(LDS|STS|ATOMS)"), ("global / local memory", r"^(LDG|STG|LD\b|ST\b|LDL|STL|ATOMG|RED|ATOM\b)"),
           ("bulk copy + mbarrier", r"^(UBLKCP|SYNCS|UTMALDG|UTMASTG)"), ("warp vote / shuffle / match", r"^(VOTE|SHFL|MATCH|POPC|BREV|FLO|REDUX)"),
           ("branch / convergence", r"^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|CALL|RET|BRX|JMP|NANOSLEEP|YIELD|BAR|BREAK)"),
           ("predicate / select", r"^(ISETP|PLOP3|SEL|FSEL|P2R|R2P|PSETP|UISETP|UPLOP3|USEL|FSETP)"),
           ("integer / address", r"^(IMAD|IADD3|IADD|LEA|LOP3|SHF\.|SHL|SHR|PRMT|MOV|UMOV|UIADD3|UIMAD|ULEA|ULOP3|USHF|S2R|S2UR|CS2R|R2UR|UR2R|I2I|IABS|IMNMX|VIADD|UPRMT|ULDC|LDC|LDCU)")]


def census(kernel, lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    hits = [b for b in blocks[1:] if kernel in b.split("\n", 1)[0]]
    out = []
    for b in hits:
        name, body = b.split("\n", 1)
        ops = collections.Counter()
        for line in body.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                ops[m.group(1)] += 1
        out.append((name.strip(), ops))
    return out


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "k_pushIdLb0ELb0"
    lib = sys.argv[2] if len(sys.argv) > 2 else LIB
    for name, ops in census(kernel, lib):
        total = sum(ops.values())
        print(f"== {name}: {total} SASS instructions (static)")
        rest = collections.Counter(ops)
        for label, pat in CLASSES:
            sel = {o: n for o, n in ops.items() if re.match(pat, o)}
            for o in sel:
                rest.pop(o, None)
            top = ", ".join(f"{o} {n}" for o, n in sorted(sel.items(), key=lambda kv: -kv[1])[:8])
            print(f"  {label:28s} {sum(sel.values()):6d}  ({top})")
        top = ", ".join(f"{o} {n}" for o, n in rest.most_common(10))
        print(f"  {'other':28s} {sum(rest.values()):6d}  ({top})")
        print(f"  spills: STL {sum(n for o, n in ops.items() if o.startswith('STL'))}, LDL {sum(n for o, n in ops.items() if o.startswith('LDL'))}")


if __name__ == "__main__":
    main()
