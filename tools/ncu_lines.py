"""Per-source-line stall samples / executed instructions of one kernel from an .ncu-rep captured with --import-source on:
    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_percent]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, per = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            continue
        if hdr is None or r[0] in ("Function Name",):
            continue
        if r[0] != "" and r[2] == "-":   # a source line summary row
            key = (cur_file, int(r[0]), r[1].strip()[:100])
            s, i = int(r[i_s] or 0), int(r[i_i] or 0)
            a = per.setdefault(key, [0, 0])
            a[0] += s
            a[1] += i
    ts = sum(v[0] for v in per.values()) or 1
    ti = sum(v[1] for v in per.values()) or 1
    print("samples %d, warp instructions %d" % (ts, ti))
    for (f, ln, src), (s, i) in sorted(per.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if 100.0 * s / ts >= thr or 100.0 * i / ti >= thr:
            print("%-12s %4d  stall %5.1f%%  inst %5.1f%%  %s" % (f, ln, 100.0 * s / ts, 100.0 * i / ti, src))


if __name__ == "__main__":
    main()
