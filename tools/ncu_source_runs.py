#!/usr/bin/env python3
"""Where a kernel's warp instructions and stall samples go, from the source page of an ncu report.

    python tools/ncu_source_runs.py gpurun_out/r2_full.ncu-rep 6 [--min-share 0.01]

The second argument is the 1-based position of the launch in the report (`ncu -i REP --page raw --csv` lists them).
Consecutive SASS instructions with the same execution count form a run (a loop body, a prologue, ...); for every run
that holds at least --min-share of the kernel's warp instructions or stall samples: its length, how often it ran, its
share of the instructions and of the samples, and the average number of active threads.  No GPU needed.
"""
import argparse
import collections
import csv
import io
import subprocess


def source_rows(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{launch}"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?")
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "Warp Stall Sampling (All Samples)", "Avg. Threads Executed")}
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["Warp Stall Sampling (All Samples)"]]),
                         float(r[col["Avg. Threads Executed"]])))
        except (ValueError, IndexError):
            pass
    half = len(data) // 2
    if half and data[:half] == data[half:]:      # (the page lists a kernel's code twice when the report holds two views of it)
        data = data[:half]
    return name, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("launch", type=int)
    ap.add_argument("--min-share", type=float, default=0.01)
    args = ap.parse_args()
    name, data = source_rows(args.report, args.launch)
    tot = sum(d[1] for d in data)
    samples = sum(d[2] for d in data)
    print(f"== {name}\n   {tot} warp instructions, {samples} stall samples, {len(data)} SASS instructions")
    runs = []
    for i, (_, e, w, _) in enumerate(data):
        if runs and abs(runs[-1][2] - e) <= max(2, 0.02 * e):
            runs[-1][1] = i; runs[-1][3] += e; runs[-1][4] += w
        else:
            runs.append([i, i, e, e, w])
    print("   SASS lines   len   executed  instr%  stall%  threads  first .. last instruction")
    for a, b, e, te, w in runs:
        if te >= tot * args.min_share or w >= samples * args.min_share * 1.5:
            thr = sum(d[3] * d[1] for d in data[a:b + 1]) / max(1, te)
            print(f"   {a:5d}-{b:5d} {b - a + 1:5d} {e:10d}  {100 * te / tot:5.1f}  {100 * w / max(1, samples):5.1f}   {thr:5.1f}   {data[a][0][:36]} .. {data[b][0][:36]}")
    ops = collections.Counter()
    for s, e, _, _ in data:
        parts = s.split()
        if parts:
            ops[parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]] += e
    print("   top opcodes: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(12)))


if __name__ == "__main__":
    main()
