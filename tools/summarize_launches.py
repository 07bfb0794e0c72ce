"""Turn the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` into the short per-kernel
launch summary kept under profiles/ (per-launch times under ncu are cold-cache and serialised: compare the
kernels' SHARES of the step with bench.py's stage_ms, not the absolute times)."""
import collections
import csv
import sys


def main(path, out, command=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
        name = r[ik].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    step = sum(v for k, v in tot.items() if not k.startswith("k_probe"))
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none --csv: " + command,
             "(per-launch times are cold-cache and serialised: compare shares, not absolutes)"]
    for k, v in tot.most_common():
        tag = "roofline probe (cpx_probe_peaks), not part of a step" if k.startswith("k_probe") else "%5.1f%% of step kernels" % (100 * v / step)
        lines.append("%-62s launches %4d  total %12.0f ns  %s" % (k, cnt[k], v, tag))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
