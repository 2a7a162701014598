"""Groups the SASS lines of an `ncu --page source --csv` export into runs of equal execution count (basic-block-like)
and prints each run's share of the executed instructions and of the stall samples, with its top stall reasons.
  python tools/ncu_blocks.py source.csv [min_share_pct]"""
import collections
import csv
import sys


def main(path, min_share=0.5):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    ci, si, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[ci] or 0) for r in data)
    tots = sum(float(r[si] or 0) for r in data)
    print(f"total warp instructions {tot:.0f}, samples {tots:.0f}, SASS lines {len(data)}")
    blocks = []
    for n, r in enumerate(data):
        c, s = float(r[ci] or 0), float(r[si] or 0)
        if blocks and abs(blocks[-1]["c"] - c) < 0.5:
            b = blocks[-1]
            b["n"] += 1
            b["s"] += s
            b["end"] = n
        else:
            blocks.append(dict(c=c, n=1, s=s, start=n, end=n))
    for b in blocks:
        share = b["c"] * b["n"] / tot
        if 100 * share > min_share or 100 * b["s"] / tots > 2 * min_share:
            acc = collections.Counter()
            for r in data[b["start"]:b["end"] + 1]:
                for i in stall:
                    if r[i]:
                        acc[hdr[i][6:]] += float(r[i])
            t = sum(acc.values()) or 1
            top = {k: round(100 * v / t) for k, v in acc.most_common(4)}
            print(f"sass {b['start']:4d}-{b['end']:4d} n={b['n']:3d} exec={b['c']:9.0f} inst {100 * share:5.1f}% "
                  f"samples {100 * b['s'] / tots:5.1f}%  {data[b['start']][src].strip()[:32]:32s} {top}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
