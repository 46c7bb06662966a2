"""Warp-stall samples of an .ncu-rep (ncu --set full --import-source on), aggregated from the
SASS source page: totals per stall reason, per opcode, the loop body size (instructions by
execution count) and the hottest instructions.  usage: ncu_stalls.py rep [kernel-id] >> profiles/x.txt"""
import collections
import csv
import subprocess
import sys


def main(rep):
    if rep.endswith(".csv"):  # a saved `ncu -i rep --page source --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    by_op = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    by_exec = collections.Counter()
    n_by_exec = collections.Counter()
    top = []
    for k, r in enumerate(data):
        n = int(r[ix["# Samples"]])
        ie = int(r[ix["Instructions Executed"]])
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
        by_op[op][0] += n
        by_op[op][1] += ie
        for s in stalls:
            v = int(r[ix[s]])
            tot[s] += v
            by_op[op][2][s] += v
        by_exec[ie] += n
        n_by_exec[ie] += 1
        top.append((n, k, src, ie, r))
    total = sum(by_exec.values())
    print(f"\n# warp-stall samples, total {total}; SASS instructions {len(data)}")
    print("# by reason: " + ", ".join(f"{s[6:]} {v}" for s, v in tot.most_common(8)))
    print("# by execution count (instructions with that count, their samples): " +
          ", ".join(f"{ie}: {n_by_exec[ie]} instr / {v}" for ie, v in by_exec.most_common(6)))
    print("# by opcode (samples, warp instructions executed, top stall reasons)")
    for op, (n, ie, c) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{op:12s} {n:6d} {ie:10d}  {[(s[6:], v) for s, v in c.most_common(3)]}")
    print("# hottest instructions (samples, line, SASS, executed, top stall reasons)")
    top.sort(key=lambda t: -t[0])
    for n, k, src, ie, r in top[:16]:
        st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"{n:6d} {k:5d} {src[:64]:64s} {ie:9d} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
