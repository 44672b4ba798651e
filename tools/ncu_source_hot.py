#!/usr/bin/env python
"""Hot source lines of a `ncu --page source --csv --print-source cuda,sass` export: stall samples per CUDA line.
Usage: tools/ncu_source_hot.py X_src.csv [instance=0] [top=25]"""
import csv
import sys


def main(path, inst=0, top=25):
    rows = list(csv.reader(open(path)))
    blocks, cur, fpath = [], None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fpath = r[1]
        elif r and r[0] == "Function Name":
            cur = {"name": r[1], "head": None, "lines": [], "file": fpath}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Line No":
            cur["head"] = r
        elif cur is not None and cur["head"] is not None and r and r[0] not in ("", "File Path"):
            cur["lines"].append(r)
    # blocks come per (kernel instance, source file); keep the .cu blocks and pick the instance among those
    own = [b for b in blocks if b.get("file", "").endswith(".cu")] or blocks
    b = own[inst]
    h = b["head"]
    L = len(h)
    # source text may contain quotes / commas that break the CSV quoting: numeric columns are aligned from the END
    i_s = h.index("# Samples") - L
    i_x = h.index("Instructions Executed") - L
    stall = [i - L for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    lines = [r for r in b["lines"] if len(r) >= L and r[i_s].isdigit()]
    tot = sum(int(r[i_s]) for r in lines)
    print(b["name"], "| .cu instances:", len(own), "| total samples:", tot)
    for r in sorted(lines, key=lambda r: -int(r[i_s]))[:top]:
        st = sorted(((int(r[i]), h[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        src = ",".join(r[1:len(r) - L + 2]).strip()
        print("%5s %6.1f%% x%-8s %-100s %s" % (r[0], 100.0 * int(r[i_s]) / max(tot, 1), r[i_x], src[:100],
                                              " ".join(f"{n[6:]}={v}" for v, n in st)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
