#!/usr/bin/env python3
"""Turn an `ncu --set full` report into the text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "free-text header" > profiles/rNN_x.txt

Per kernel launch in the report: duration, DRAM bytes, issue / pipe utilisation, shared-memory
wavefronts, occupancy limits; then, per kernel, the SASS instruction count split at BAR.SYNC
boundaries (which stage of the kernel the issue slots go to).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    drop = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # leading launches of the report to leave out (a window that
    print(sys.argv[2] if len(sys.argv) > 2 else rep)      # starts in the middle of a step)
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    names = []
    for r in rows[2 + drop:]:
        name = r[kcol].split("(")[0]
        names.append(name)
        print(f"\n== launch {r[0]}: {r[kcol]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:70s} {r[i]:>16s} {units[i]}")
    for name in dict.fromkeys(names):
        base = name.replace("void ", "").split("<")[0].strip()   # templates: ncu matches the base name
        out = ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:^" + base + "$"])
        srows = list(csv.reader(io.StringIO(out)))
        try:
            h = next(r for r in srows if "Instructions Executed" in r)
        except StopIteration:
            continue
        ie, sm = h.index("Instructions Executed"), h.index("# Samples")
        body = []
        for r in srows[srows.index(h) + 1:]:
            if r and r[0] == "Kernel Name":  # the page repeats per matching launch: keep the first
                break
            if len(r) > ie and r[ie].isdigit():
                body.append(r)
        stalls = {c: sum(int(r[i] or 0) for r in body if len(r) > i) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c}
        st_tot = sum(stalls.values()) or 1
        tot = sum(int(r[ie]) for r in body) or 1
        ts = sum(int(r[sm]) for r in body) or 1
        print(f"\n== {name}: warp stall samples, first launch in report: " +
              ", ".join(f"{k[6:]} {100 * v / st_tot:.0f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]))
        print(f"== {name}: SASS instructions executed by segment (split at BAR.SYNC), first launch in report")
        seg = acc = accs = start = 0
        for i, r in enumerate(body):
            acc += int(r[ie])
            accs += int(r[sm])
            if "BAR.SYNC" in r[1] or i == len(body) - 1:
                print(f"segment {seg:2d}  sass[{start:4d}..{i:4d}]  inst {acc:12d} ({100 * acc / tot:5.1f} %)  "
                      f"stall samples {accs:7d} ({100 * accs / ts:5.1f} %)")
                seg += 1
                acc = accs = 0
                start = i + 1


if __name__ == "__main__":
    main()
