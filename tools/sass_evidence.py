"""Static SASS evidence per kernel of libmjrl_b200.so (no GPU needed): cuobjdump -sass, count the mnemonics that prove the
Blackwell paths (tcgen05 / TMEM / TMA bulk copies / mbarriers / packed fp32 / fp64 FMA / system-scope stores).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "UBLKCP", "LDTM", "STTM", "SYNCS", "STAS", "UCGABAR", "FFMA2", "FMUL2", "DFMA", "MUFU",
        "LDGSTS", "REDG", "ELECT", "MEMBAR", "CCTL"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "mjrl_b200/libmjrl_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    print("# SASS evidence (cuobjdump -sass %s, sm_100a): static instruction counts per kernel" % so)
    print("# UTCHMMA = tcgen05.mma kind::f16 | UTCBAR = tcgen05.commit | UTCATOMSWS = tcgen05.alloc/dealloc | LDTM/STTM = tcgen05.ld/st")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy) | SYNCS = mbarrier ops | UCGABAR = barrier.cluster | FFMA2/FMUL2 = packed fp32x2")
    print("# DFMA = fp64 FMA (ridge baselines) | .SYS = system-scope loads / stores (peer-memory all-reduce) | ELECT = elect.sync")
    print()
    name, counts, total, rows = None, None, 0, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, total, counts))
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total += 1
            base = op.split(".")[0]
            if base in KEYS:
                counts[base] += 1
            if ".SYS" in op and base in ("LD", "ST", "LDG", "STG"):
                counts[base + ".SYS"] += 1
    if name:
        rows.append((name, total, counts))
    for name, total, counts in rows:
        d = demangle(name)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"mjb::", "", d)
        keys = [k for k in KEYS + ["LD.SYS", "ST.SYS", "LDG.SYS", "STG.SYS"] if counts.get(k)]
        print("%-78s %6d instr  %s" % (d[:78], total, "  ".join("%s=%d" % (k, counts[k]) for k in keys)))


if __name__ == "__main__":
    main()
