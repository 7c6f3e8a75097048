"""Summarise ncu artefacts brought back from the GPU box into small text/JSON files under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/launches_r01g.csv profiles/launches_r01.txt
    python tools/summarise_ncu.py full gpurun_out/prof_bench_r01g.ncu-rep profiles/ncu_gen_gemm_r01.txt [profiles/roofline_traffic.json c4]

`launches`: per-kernel totals and shares of the `--metrics gpu__time_duration.sum` launch list.
`full`: the metrics DESIGN.md quotes from one `ncu --set full` capture (read here with `ncu -i ... --page raw --csv`).
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warps_active.avg.per_cycle_active",
    # the FP64 tensor sub-pipe (DMMA) and the pipe it shares with DFMA/DADD/DMUL ("shared pipe"): busy cycles and instruction counts
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__inst_executed_pipe_fp64.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0][-90:]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[name][0] += 1
        tot[name][1] += v
    all_ms = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# {src}: per-kernel totals of gpu__time_duration.sum (ncu --clock-control none; cold-cache, serialised: compare SHARES)\n")
        f.write(f"# total {all_ms:.3f} ms over {sum(v[0] for v in tot.values())} launches\n")
        for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{ms:12.3f} ms  {100 * ms / all_ms:6.2f} %  x{n:<4d} {name}\n")
    print(open(dst).read())


def full(src, dst, traffic_json=None, tag=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        out.append(f"== {name}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] not in ("", "nan", "-nan"):
                out.append(f"   {k:85s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "nan", "-nan")),
                        reverse=True)
        out.append("   warp stall reasons (cycles per issued instruction): " +
                   ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, k in stalls[:7]))

        def val(key):
            v = float(d[key].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]

        try:
            import re
            margs = re.search(r"gen_gemm_kernel<([^>]*)>", name)
            targs = [a.strip().split(")")[-1] for a in margs.group(1).split(",")] if margs else []
            role = "backward" if len(targs) > 1 and targs[1] in ("1", "true") else "forward"
            traffic[role] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            out.append(f"   dram traffic per launch: {traffic[role] / 1e9:.3f} GB  ({role} role)")
        except Exception as e:  # noqa: BLE001
            out.append(f"   dram traffic: unavailable ({e})")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    if traffic_json and traffic:
        try:
            cur = json.load(open(traffic_json))
        except Exception:  # noqa: BLE001
            cur = {}
        cur[tag] = {"per_launch_bytes_mean": sum(traffic.values()) / len(traffic), "per_role_bytes": traffic, "source": src.split("/")[-1]}
        json.dump(cur, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
