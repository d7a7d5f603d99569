#!/usr/bin/env python
"""Developer tool: per-source-line instruction counts and stall samples of one kernel.

Joins `ncu -i REP --page source --csv` (per-SASS-instruction metrics) with the line table of
the matching cubin (`nvdisasm -g`), by instruction offset.  Works without a GPU.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep shade k_shade [--top 40] [--inlined]
                              (cubin stem) (kernel name substring)
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.abspath(os.environ.get("PAR_B200_LIB") or os.path.join(ROOT, "pixel-art-raytracer_b200", "par_b200", "libpar_b200.so"))  # (override: A/B builds)


def line_table(stem, kernel):
    """offset -> (file, line, inlined_at_line) for the kernel in <stem>.sm_100a.cubin."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.startswith(stem + ".")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], check=True,
                             capture_output=True, text=True).stdout
    table, cur, outer, active = {}, None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = kernel in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            outer = (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, outer, m.group(2).strip())
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("stem")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--outer", action="store_true", help="attribute inlined code to its call site")
    ap.add_argument("--sass", action="store_true", help="list the hottest SASS instructions too")
    ap.add_argument("--by", default="samples", choices=["samples", "inst"], help="sort key")
    ap.add_argument("--symbol", default=None, help="substring of the mangled name in the cubin (default: the kernel "
                                                   "name), e.g. k_tileILb0 for k_tile<false>")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "-k", "regex:" + a.kernel], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    col = {name: h.index(name) for name in ("Address", "Source", "# Samples", "Instructions Executed",
                                            "Thread Instructions Executed")}
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    table = line_table(a.stem, a.symbol or a.kernel)
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    sass = []
    for r in rows[hdr + 1:]:
        try:
            addr = int(r[col["Address"]], 16)
        except (ValueError, IndexError):
            continue
        base = addr if base is None else base
        cur, outer, _ = table.get(addr - base, (None, None, ""))
        key = (outer if (a.outer and outer) else cur) or ("?", 0)
        inst, thr, smp = int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]])
        e = per_line[key]
        e[0] += inst
        e[1] += thr
        e[2] += smp
        for i, n in stall_cols:
            if r[i] not in ("", "0"):
                e[3][n] += int(r[i])
        sass.append((inst, smp, key, r[col["Source"]].strip()))
    ti = sum(e[0] for e in per_line.values()) or 1
    ts = sum(e[2] for e in per_line.values()) or 1
    tt = sum(e[1] for e in per_line.values())
    print(f"warp instructions {ti:,}  thread instructions {tt:,}  (avg {tt / ti:.1f} active lanes)  samples {ts:,}")
    src_cache = {}
    for key, e in sorted(per_line.items(), key=lambda kv: -kv[1][2 if a.by == 'samples' else 0])[:a.top]:
        f, ln = key
        if f not in src_cache:
            p = next((os.path.join(dp, f) for dp, _, fs in os.walk(ROOT) if f in fs and "build" not in dp), None)
            src_cache[f] = open(p).read().splitlines() if p else []
        text = src_cache[f][ln - 1].strip() if 0 < ln <= len(src_cache[f]) else ""
        top = ", ".join(f"{n[6:]} {c * 100 // max(e[2], 1)}%" for n, c in e[3].most_common(3))
        print(f"{e[0] / ti * 100:5.1f}% inst {e[2] / ts * 100:5.1f}% smp  lanes {e[1] / max(e[0], 1):4.1f}  {f}:{ln:<4} {text[:70]:70s} [{top}]")
    if a.sass:
        print("--- hottest SASS by samples")
        for inst, smp, key, s in sorted(sass, key=lambda x: -x[1])[:a.top]:
            print(f"{inst / ti * 100:5.1f}% inst {smp / ts * 100:5.1f}% smp  {key[0]}:{key[1]:<4} {s}")


if __name__ == "__main__":
    main()
