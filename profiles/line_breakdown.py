"""Per-source-line executed warp instructions of one kernel in an ncu --set full --import-source on capture:
python profiles/line_breakdown.py prof.ncu-rep <kernel regex> [min_pct]"""
import csv, subprocess, sys


def main(rep, kern, minpct=0.4):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    cur = hdr = None
    out = []
    for r in csv.reader(txt.splitlines()):
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
    print("total warp instructions %.0f" % ti)
    out.sort(key=lambda d: (d["file"], int(d["Line No"])))
    for d in out:
        i = f(d, "Instructions Executed")
        if i > ti * minpct / 100:
            print("%-16s %4s %9.0f %5.1f%% lanes %4.1f smp %4.0f  %s" % (d["file"][:16], d["Line No"], i, 100 * i / ti,
                  f(d, "Thread Instructions Executed") / max(i, 1), f(d, "# Samples"), d["Source"].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.4)
