#!/usr/bin/env python
"""SASS evidence for the tensor-core / TMA claims, from the shipped library: per kernel the counts of the Blackwell-specific
mnemonics (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier
ops, UTCATOMSWS = tensor-memory allocation), plus the level loop of the fused tree kernel that tests/test_abi.py parses.

  python scripts/sass_evidence.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "alphazero-implementation_b200", "libaz_engine.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "UTCATOMSWS", "NANOSLEEP", "HMMA", "DFMA", "SHFL", "FADD2", "F2FP", "LDS", "STS", "LDG", "STG"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)[:110]
        kernels[name] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
    if m and name:
        kernels[name].append((int(m.group(1), 16), m.group(2)))

print("# cuobjdump -sass alphazero-implementation_b200/libaz_engine.so  (sm_100a) - mnemonic counts per kernel")
print(f"{'kernel':112s} {'instr':>6s} " + " ".join(f"{o:>9s}" for o in OPS))
for k, ins in kernels.items():
    c = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") and len(t.split()) > 1 else t.split()[0]
        for o in OPS:
            if op.startswith(o):
                c[o] += 1
    print(f"{k:112s} {len(ins):6d} " + " ".join(f"{c[o]:9d}" for o in OPS))

# the level loop of the fused tree kernel (variant the config-2 bench runs)
ins = next((v for k, v in kernels.items() if "k_run_sims<4, 1, true, true, true>" in k), None)
if ins:
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            body = [(a, t) for a, t in ins if int(m.group(1), 16) <= a <= addr]
            if any("SHFL.BFLY" in t for _, t in body):
                loops.append(body)
    level = min(loops, key=len)
    print(f"\n# k_run_sims<4, uniform, latency, move, tables in shared memory>: the level loop ({len(level)} instructions)")
    for a, t in level:
        print(f"  /*{a:04x}*/  {t}")
