"""SASS evidence from the shipped libjpgb200.so: per product kernel the counts of the mnemonics that matter
(TMA, packed FP32x2, mbarrier, cp.async, ...) and the first lines of each.
python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

LIB = "jpeg-encoder-and-decoder_b200/libjpgb200.so"
KERNELS = ["k1_transform_pcILi0ELb0", "k1_transform_pctILi0ELb0", "k_encode_chunksILi0ELb0", "k_encode_chunksILi0ELb1", "k_count_ff", "k_scatter", "k1e_fusedILi0ELb0"]
WANT = ["UTMASTG", "UTMALDG", "UBLKCP", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "SYNCS", "USETMAXREG", "HSET2", "RED", "ATOMS", "ATOMG", "BAR.SYNC", "LDS", "STS", "FLO", "PRMT", "VOTE", "SHFL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    print("SASS evidence from the shipped libjpgb200.so (cuobjdump -sass; sm_100a cubin only), default build.")
    print("Per kernel: counts of the mnemonics that matter and the first lines of each.\n")
    for k in KERNELS:
        for name, ins in funcs.items():
            if k not in name:
                continue
            cnt = collections.Counter()
            first = collections.defaultdict(list)
            for i in ins:
                body = re.sub(r"^@!?U?P\d+\s+", "", i)
                for w in WANT:
                    if body.startswith(w):
                        cnt[w] += 1
                        if len(first[w]) < 2:
                            first[w].append(i)
            print("== %s  (%d instructions)" % (name, len(ins)))
            print("   " + "  ".join("%s=%d" % (w, cnt[w]) for w in WANT if cnt[w]))
            for w in ("UTMASTG", "UTMALDG", "USETMAXREG", "FFMA2", "LDGSTS", "SYNCS", "HSET2"):
                for i in first[w]:
                    print("      " + i)
            print()


if __name__ == "__main__":
    main()
