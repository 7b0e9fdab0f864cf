"""Summarise ncu outputs into small text files for profiles/ (the raw .ncu-rep / csv stay in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/launches_cfg3.csv > profiles/r01_launches_cfg3.txt
    python tools/ncu_summary.py full gpurun_out/fvp_cfg3.ncu-rep  > profiles/r01_fvp_cfg3_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        ms = val / 1e6 if unit.startswith("n") else (val / 1e3 if unit.startswith("u") else val)
        a = agg.setdefault(d["Kernel Name"], [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)")
    print("%-86s %6s %11s %7s %10s  %s" % ("kernel", "n", "total ms", "share", "ms/launch", "grid x block"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-86s %6d %11.3f %6.1f%% %10.4f  %s x %s" % (k[:86], v[0], v[1], 100 * v[1] / tot, v[1] / v[0], v[2], v[3]))
    print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none, one block per captured launch")
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for k in KEYS:
            if k in idx:
                print("   %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        stalls = sorted(((float(r[idx[h]] or 0), h) for h in hdr
                         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)[:6]
        for v, h in stalls:
            print("   stall %-66s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
