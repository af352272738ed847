"""Aggregates the ncu source page (ncu -i X --page source --csv -k regex:K --print-source cuda,sass) by CUDA source line:
stall samples and executed instructions. Usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex [topN]"""
import collections
import csv
import subprocess
import sys


def main(rep, rx, top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # several kernels/files are concatenated: find header rows
    agg = collections.OrderedDict()
    cur_file, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr, r))
        src = r[1]
        key = (cur_file, line)

        def num(k):
            try:
                return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError:
                return 0.0
        a = agg.setdefault(key, [src.strip()[:110], 0.0, 0.0])
        a[1] += num("# Samples")
        a[2] += num("Instructions Executed")
    tot_s = sum(a[1] for a in agg.values()) or 1.0
    tot_i = sum(a[2] for a in agg.values()) or 1.0
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    for (f, line), (src, smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * smp / tot_s, 100 * ins / tot_i, f, line, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
