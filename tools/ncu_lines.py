"""Per-source-line view of an ncu capture, for kernels whose `--page source --print-source cuda --csv` export carries no metrics.

    python tools/ncu_lines.py REPORT.ncu-rep LIB.so MANGLED_SUBSTRING [regions.json] > summary.md

Joins ncu's SASS rows (`--page source --csv`: one row per instruction, in address order) with `nvdisasm -g` of the same
cubin (`//## File "...", line N` markers; inlined frames carry `inlined at` lines, the innermost line is used) by instruction
index.  Prints warp-stall samples and executed instructions per source line and per region.  Development tool: it reads
reports under gpurun_out/ and is not part of any timed path.
"""
import csv
import json
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, needle, main_file=None):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, file_, line, home = None, None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            if cur_fn is None or needle not in cur_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                file_, line = os.path.basename(m.group(1)), int(m.group(2))
                if main_file is None or file_ == main_file:
                    home = (file_, line)
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                # instructions of inlined helpers from other files (plain -lineinfo keeps only the innermost line) are
                # charged to the last line of the main file seen before them
                k = home if (main_file is not None and file_ != main_file and home is not None) else (file_, line)
                out.append((int(m.group(1), 16), m.group(2).strip(), k[0], k[1]))
        if out:
            break
    return out


def main():
    rep, lib, needle = sys.argv[1:4]
    regions = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else []
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: hdr.index(n) for n in hdr}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    sl = sass_lines(lib, needle, os.environ.get("MAIN_FILE"))
    if len(sl) != len(body):
        print(f"warning: {len(sl)} disassembled instructions vs {len(body)} profiled rows", file=sys.stderr)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    per = {}
    for (off, ins, f, ln), r in zip(sl, body):
        k = (f, ln)
        d = per.setdefault(k, {"samples": 0, "inst": 0, "stalls": {}, "n_sass": 0})
        d["samples"] += int(r[col["# Samples"]])
        d["inst"] += int(r[col["Instructions Executed"]])
        d["n_sass"] += 1
        for s in stall_cols:
            v = int(r[col[s]])
            if v:
                d["stalls"][s] = d["stalls"].get(s, 0) + v
    tot = sum(d["samples"] for d in per.values()) or 1
    tot_i = sum(d["inst"] for d in per.values()) or 1
    print(f"# {os.path.basename(rep)}: {needle}\n\ntotal samples {tot}, warp instructions {tot_i}\n")
    if regions:
        print("| region | lines | samples | % | warp instructions | % |\n|---|---|---|---|---|---|")
        for name, f, a, b in regions:
            s = sum(d["samples"] for (ff, ln), d in per.items() if ff == f and ln is not None and a <= ln <= b)
            i = sum(d["inst"] for (ff, ln), d in per.items() if ff == f and ln is not None and a <= ln <= b)
            print(f"| {name} | {f}:{a}-{b} | {s} | {100 * s / tot:.1f} | {i} | {100 * i / tot_i:.1f} |")
        print()
    print("| line | samples | % | warp inst | SASS | top stalls |\n|---|---|---|---|---|---|")
    for (f, ln), d in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:int(os.environ.get("TOP", "60"))]:
        st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"| {f}:{ln} | {d['samples']} | {100 * d['samples'] / tot:.1f} | {d['inst']} | {d['n_sass']} | {st} |")


if __name__ == "__main__":
    main()
