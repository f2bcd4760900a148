"""Rank CUDA source lines of an ncu --set full --import-source on capture by warp-stall samples and
executed instructions: python profiles/rank_source.py prof.ncu-rep [top_n] [kernel regex]"""
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, out = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            d["file"] = cur
            out.append(d)

    def f(d, k):
        try:
            return float(d.get(k) or 0)
        except ValueError:
            return 0.0

    ti = sum(f(d, "Instructions Executed") for d in out)
    ts = sum(f(d, "# Samples") for d in out)
    print("total warp instructions %.0f, stall samples %.0f" % (ti, ts))
    out.sort(key=lambda d: -f(d, "# Samples"))
    print("%-18s %5s %7s %6s %9s %6s  %s" % ("file", "line", "samples", "%", "inst", "%", "source"))
    for d in out[:top]:
        reasons = sorted(((f(d, k), k[6:]) for k in d if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
        rs = ",".join("%s=%.0f" % (k, v) for v, k in reasons if v > 0)
        print("%-18s %5s %7.0f %5.1f%% %9.0f %5.1f%%  %s  [%s]" % (d["file"][:18], d["Line No"], f(d, "# Samples"),
              100 * f(d, "# Samples") / max(ts, 1), f(d, "Instructions Executed"), 100 * f(d, "Instructions Executed") / max(ti, 1),
              d["Source"].strip()[:90], rs))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
