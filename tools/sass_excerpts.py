"""SASS evidence for the hot kernels (no GPU needed): per-kernel opcode histogram of the
compiled sm_100a code, register / shared-memory footprint, and the mnemonics that show the
hardware paths the design relies on (TMA bulk copies + mbarrier waits, the 64-bit reciprocal
seed of the lean cell, multimem stores of the sharded accept step).

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "naima_b200", "libnaima_b200.so")
WANT = ["contract_kernelILi8ELi2", "contract_kernelILi8ELi0", "synchrotron_fused_kernel",
        "walker_prep_kernel", "combine_lnprob_kernel", "ssc_inner_kernelILi8",
        "ssc_outer_kernel"]
PROOF = ["UBLKCP", "SYNCS", "MUFU.RCP64H", "MUFU.RSQ64H", "DFMA", "DMUL", "DADD", "DSETP", "LDS",
         "LDG", "STG", "STG.E.64.STRONG.SYS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "REDG", "ATOMG", "MEMBAR", "LDL", "STL", "BAR.SYNC",
         "SHFL", "CCTL", "ERRBAR", "FENCE"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True,
                     text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1).strip())

print("# SASS of the hot kernels (sm_100a, `cuobjdump -sass naima_b200/libnaima_b200.so`)\n")
print("Static instruction counts of the compiled code (not executed counts: those are in the "
      "ncu summaries).  Produced by `tools/sass_excerpts.py`.  `UBLKCP` = `cp.async.bulk` (TMA "
      "bulk copy), `SYNCS.*` = mbarrier arrive / try_wait, `MUFU.RCP64H` = the 64-bit reciprocal "
      "seed of the lean cell, `STG.E.64.STRONG.SYS` in the combine kernel = `multimem.st` / peer "
      "stores of the sharded accept step (the multicast mapping is in the address), one "
      "`MEMBAR.*.SYS` = the single system-scope release per half-step.\n")
for want in WANT:
    for fn, ins in funcs.items():
        if want not in fn:
            continue
        ops = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op] += 1
        reg, stack, shared = usage.get(fn, (0, 0, 0))
        print("## `%s`\n" % fn)
        print("%d instructions, %d registers, %d B stack, %d B static shared memory\n"
              % (len(ins), reg, stack, shared))
        print("| mnemonic family | static count |\n|---|---|")
        for p in PROOF:
            n = sum(v for k, v in ops.items() if k.startswith(p) or ("." + p) in k)
            if n:
                print("| %s | %d |" % (p, n))
        print("\ntop opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(14)) + "\n")
        # the first bulk copy and the reciprocal seed, verbatim
        shown = 0
        for i in ins:
            if any(k in i for k in ("UBLKCP", "MUFU.RCP64H", "SYNCS.ARRIVE", "SYNCS.PHASECHK",
                                    "STRONG.SYS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS")) and shown < 6:
                print("    " + i)
                shown += 1
        print()
