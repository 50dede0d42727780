"""Dynamic instruction mix of a profiled kernel from `ncu -i X.ncu-rep --page source --csv`.

  python tools/ncu_dynmix.py <report.ncu-rep> <warps> <steps>

Sums the per-SASS-instruction "Instructions Executed" column by class and prints warp-instructions
per (warp x step), the FP64 pipe cost under the measured B200 issue costs (tools/microbench2.cu:
DFMA with three vector-register sources 3 cycles, other FP64 ~2) and the top stall lines.
"""
import collections
import csv
import io
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from sass_mix import family


def main():
    rep, warps, steps = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    cnt = collections.Counter()
    stalls = []
    tot_samples = 0
    for row in rd:
        src = row["Source"].strip()
        try:
            ex = float(row["Instructions Executed"])
        except (ValueError, KeyError):
            continue
        parts = src.split()
        if not parts:
            continue
        k = 1 if parts[0].startswith("@") else 0
        fam = family(parts[k], " ".join(parts[k + 1:]))
        cnt[fam] += ex
        smp = float(row.get("# Samples") or 0)
        tot_samples += smp
        stalls.append((smp, src, ex))
    unit = warps * steps
    tot = sum(cnt.values())
    fp64 = sum(v for k, v in cnt.items() if k.startswith("fp64"))
    cyc = sum(v * (3.0 if k.endswith("3reg") else 2.1 if "DFMA" in k else 2.0) for k, v in cnt.items() if k.startswith("fp64"))
    print("warp-instructions per warp-step: %.1f total, %.1f fp64; fp64 pipe cost %.1f cycles per warp-step"
          % (tot / unit, fp64 / unit, cyc / unit))
    for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:24]:
        print("    %-28s %8.1f" % (k, v / unit))
    if "--stalls" in sys.argv:
        stalls.sort(reverse=True)
        print("top sampled instructions (%d samples):" % tot_samples)
        for smp, src, ex in stalls[:25]:
            print("    %5.2f%%  %s" % (100.0 * smp / max(tot_samples, 1), src))


if __name__ == "__main__":
    main()
