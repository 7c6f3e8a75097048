"""Per-kernel SASS evidence for the built library: counts of the instructions that prove the Blackwell path (DMMA.8x8x4 = the FP64
tensor instruction, UTMALDG = tensor-map TMA, UBLKCP = bulk copy, SYNCS = mbarrier, STTM / LDTM = tensor-memory store / load),
registers, spill and stack bytes (cuobjdump -res-usage).

    python tools/sass_counts.py [projected_langevin_sampling_b200/libpls_b200.so] > profiles/sass_gen_gemm_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "projected_langevin_sampling_b200", "libpls_b200.so")
OPS = ["DMMA", "DFMA", "DADD", "DMUL", "UTMALDG", "UBLKCP", "SYNCS", "STTM", "LDTM", "LDS", "STG", "LDL", "STL"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: dict(zip(names, subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines())) if names else {}

counts, arch, cur = collections.OrderedDict(), {}, None
cur_arch = "?"
for line in sass.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        cur_arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        arch[cur] = cur_arch
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        if op in OPS:
            counts[cur][op] += 1
usage = {}
fn = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    if fn and "REG:" in line:
        usage[fn] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
        fn = None
names = demangle(list(counts))
print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels / device functions; archs: {sorted(set(arch.values()))}")
print("# gen_gemm_kernel<NKD, BACKWARD, KSRC (0 linear, 1 RBF generated, 2 cached Gram), RT, EPI (-1 backward), NS (2 = second accumulator set parked in tensor memory), CL (2 = CTA pairs sharing the Gram generation over DSMEM)>")
print(f"# {'REG':>4s} {'STACK':>5s} {'instr':>7s} " + " ".join(f"{o:>7s}" for o in OPS) + "  kernel")
tot = collections.Counter()
for f, c in counts.items():
    u = usage.get(f, {})
    name = names.get(f, f)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|pls::|void |\((?:int|bool)\)", "", name).split("(")[0]
    print(f"  {u.get('REG', 0):4d} {u.get('STACK', 0):5d} {c['total']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS) + f"  {name}")
    tot.update(c)
print(f"# totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS) + "; UTC*MMA / tcgen05.mma: none (no FP64 kind exists)")
