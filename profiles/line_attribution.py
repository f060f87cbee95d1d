#!/usr/bin/env python
"""Instructions executed per particle and stall samples per CUDA source line of k_push<double,0,0>, from an .ncu-rep captured with
--import-source on and the cubin's line info:  python profiles/line_attribution.py gpurun_out/push_r1_final.ncu-rep [n_particles [mangled-kernel-prefix]]
(another kernel of the library: e.g. _ZN3jic16k_cn_push_sortedIdE)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax-in-cell_b200", "jaxincell_b200", "libjic_b200.so")
KERNEL = "_ZN3jic6k_pushIdLb0ELb0E"


def main(rep, n_particles, KERNEL=KERNEL):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    lines = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(lines) if l.startswith(".text." + KERNEL)][0]
    end = [i for i, l in enumerate(lines[start + 1:], start + 1) if l.strip().startswith(".section")][0]
    cur, amap = None, {}
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            amap[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    first = [i for i, r in enumerate(rows) if "Address" in r][0]
    hdr, data = rows[first], [r for r in rows[first + 1:] if len(r) == len(rows[first]) and re.fullmatch(r"(0x)?[0-9a-f]+", r[rows[first].index("Address")])]
    iS, iE, iP, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    base = int(data[0][iA], 16)
    per, samp, fp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        loc = amap.get(int(r[iA], 16) - base)
        n = int(r[iE])
        per[loc] += n
        samp[loc] += int(r[iP])
        op = r[iS].strip().split()
        op = op[1] if op[0].startswith("@") else op[0]
        if op[:2] in ("DF", "DA", "DM", "DS"):
            fp[loc] += n
    src = {f: open(os.path.join(ROOT, "jax-in-cell_b200", "csrc", f)).read().split("\n") for f in ("jic_push.cuh", "jic_binned.cuh", "jic_device.cuh", "jic_cn.cuh", "jic_cn_sorted.cuh", "jic_kernels.cuh")}
    tot, tots = sum(per.values()), sum(samp.values())
    k = 32.0 / n_particles
    print(f"# {rep}: {tot} warp instructions = {tot * k:.1f} per particle, {sum(fp.values()) * k:.1f} of them FP64; {tots} stall samples")
    print("# file:line   instr/particle   FP64/particle   stall samples   source")
    for loc, n in per.most_common(48):
        f, ln = loc if loc else ("?", 0)
        text = src[f][ln - 1].strip()[:95] if f in src and ln > 0 else ""
        print(f"{f[:14]:14s}:{ln:4d} {n * k:7.2f} {fp[loc] * k:6.2f} {100 * samp[loc] / tots:5.1f}%  {text}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1e8, *(sys.argv[3:4]))
