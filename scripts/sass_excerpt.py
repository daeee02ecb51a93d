"""Instruction-level evidence for the insert kernels without a GPU: cuobjdump -sass of libgenome_b200.so, per kernel the opcode
histogram, registers / shared memory (cuobjdump -res-usage) and every memory, atomic, vote / match and barrier instruction.
    python scripts/sass_excerpt.py > profiles/sass_insert_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genome_b200", "libgenome_b200.so")
WANT = sys.argv[1:] or ["bucket_slabs_kernelILb1ELb0ELi5ELb0E", "bucket_slabs_kernelILb1ELb0ELi7ELb1E", "insert_slabs_kernelIjLb0E", "insert_slabs_kernelIjLb1E",
                        "part_scatter_kernelILb1ELb0ELb0ELb1E", "part_count_kernelILb1ELb0ELb0E", "insert_keys_kernelILb0EjE",
                        "insert_reads_kernelILb1ELb0E", "place_distinct_kernel", "compact_survivors_kernel", "init_table_kernel", "masks_kernelILb0E",
                        "jump_kernel", "items_kernelINS0_7ProbeOpE", "items_kernelINS0_7PartsOpE", "items_kernelINS0_9CombineOpE"]
MEM = re.compile(r"\b(LDG|STG|LDS|STS|LDSM|ATOM|ATOMG|ATOMS|RED|REDG|REDUX|MATCH|VOTE|BAR|LDC|LDGSTS|UBLKCP|UTMALDG|CCTL|MEMBAR|SHFL|WARPSYNC)\b")


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    name, body = None, []
    kernels = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                kernels[name] = body
            name, body = m.group(1), []
        elif name:
            mm = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
            if mm:
                body.append((mm.group(1), mm.group(2)))
    if name:
        kernels[name] = body
    print("# SASS excerpt of libgenome_b200.so (sm_100a), made by scripts/sass_excerpt.py -- no tensor-core or TMA instructions are expected:")
    print("# nothing on this path is a dense contraction, and tiles are staged by 128-bit LDG.E.128 + STS (extract.cuh stage_tile)")
    for n, body in kernels.items():
        if not any(w in n for w in WANT):
            continue
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0] for _, ins in body)
        print("\n" + "=" * 120)
        print(demangle(n))
        print("  %d instructions; %s" % (len(body), usage.get(n, "")))
        print("  opcodes: " + ", ".join("%s x%d" % kv for kv in ops.most_common(40)))
        for addr, ins in body:
            if MEM.search(ins) and not ins.startswith("LDC") and "SHFL" not in ins:
                print("    /*%s*/ %s" % (addr, ins))


if __name__ == "__main__":
    main()
