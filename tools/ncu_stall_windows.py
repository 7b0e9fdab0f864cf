"""Per-region warp-stall summary of one kernel from an ncu report captured with --set full --import-source on:
walks the SASS in windows of W instructions and prints each window's share of the warp-state samples, its top stall reasons
and the landmark instructions it contains (tcgen05.mma, mbarrier waits, bar.sync, TMEM loads, MUFU ...).

    python tools/ncu_stall_windows.py gpurun_out/fvp_cfg3.ncu-rep [W=64] > profiles/r02_fvp_stalls.txt
"""
import collections
import csv
import io
import subprocess
import sys

MARKS = ("UTCHMMA", "UTCBAR", "SYNCS", "BAR.SYNC", "LDTM", "STTM", "MUFU", "UBLKCP", "REDG", "LDG", "STG")


def main():
    rep = sys.argv[1]
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    print("# %s" % (rows[hdr_i - 1][1] if hdr_i > 0 and len(rows[hdr_i - 1]) > 1 else rep))
    hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:] if len(r) == len(rows[hdr_i])]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    base = int(data[0][ix["Address"]], 16)
    print("# %d SASS instructions, %d warp-state samples; windows of %d instructions holding >= 1 %% of the samples" % (len(data), tot, W))
    print("# offset range      share  top stall reasons (samples)                      landmarks")
    for w0 in range(0, len(data), W):
        blk = data[w0:w0 + W]
        s = sum(int(r[ix["# Samples"]]) for r in blk)
        if s < tot * 0.01:
            continue
        st = {c: sum(int(r[ix[c]]) for r in blk) for c in stall_cols}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        marks = collections.Counter()
        for r in blk:
            op = r[ix["Source"]].strip().split()
            op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
            for m in MARKS:
                if op.startswith(m):
                    marks[m] += 1
        print("%6x-%6x  %5.1f%%  %-52s %s" % (int(blk[0][ix["Address"]], 16) - base, int(blk[-1][ix["Address"]], 16) - base, 100.0 * s / tot,
                                             "  ".join("%s:%d" % (k[6:], v) for k, v in top), " ".join("%s=%d" % kv for kv in sorted(marks.items()))))


if __name__ == "__main__":
    main()
