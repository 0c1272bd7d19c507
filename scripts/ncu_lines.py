"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report
(--set full --import-source on) and the cubin's line table (nvdisasm -gi).
Usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX LIB.so SOURCE.cu [min_pct] [--outer]
--outer attributes inlined code to the line of the outermost frame inside SOURCE.cu."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern, lib, src = sys.argv[1:5]
    thr = float(sys.argv[5]) if len(sys.argv) > 5 and not sys.argv[5].startswith("--") else 0.5
    outer = "--outer" in sys.argv
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    base = os.path.basename(src)
    cub = [f for f in os.listdir(tmp) if f.startswith(base.split(".")[0] + ".")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(sass) if l.startswith(".text.") and re.search(kern, l)][0]
    chain, fresh, addr2 = [], True, {}
    for l in sass[start + 1:]:
        if l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            addr2[int(m.group(1), 16)] = list(chain)
            fresh = True
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hi[0]]
    end = hi[1] if len(hi) > 1 else len(rows)           # first captured launch only
    ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    a0 = int(rows[hi[0] + 1][0], 16)
    inst, samp = collections.Counter(), collections.Counter()
    for r in rows[hi[0] + 1:end]:
        if len(r) <= ii or not r[0].startswith("0x"):
            continue
        ch = addr2.get(int(r[0], 16) - a0) or [("?", 0)]
        key = ch[0]
        if outer:
            own = [c for c in ch if c[0] == base]
            key = own[-1] if own else ch[-1]
        if "--fn" in sys.argv:                          # innermost frame inside [lo, hi) of SOURCE.cu
            lo, hi = map(int, sys.argv[sys.argv.index("--fn") + 1].split("-"))
            own = [c for c in ch if c[0] == base and lo <= c[1] < hi]
            key = own[0] if own else ("(outside)", 0)
        inst[key] += int(r[ii])
        samp[key] += int(r[isamp])
    ti, ts = sum(inst.values()), sum(samp.values())
    text = open(src).read().split("\n")
    print("kernel %s: %d warp instructions, %d samples" % (kern, ti, ts))
    for k in sorted(inst):
        if inst[k] * 100 >= ti * thr or samp[k] * 100 >= ts * thr:
            t = text[k[1] - 1].strip()[:100] if k[0] == base else ""
            print("%-28s %5d %6.2f%% inst %6.2f%% samp  %s" % (k[0], k[1], 100 * inst[k] / ti, 100 * samp[k] / max(ts, 1), t))


if __name__ == "__main__":
    main()
