#!/usr/bin/env python
"""Per-function stall-sample profile of one kernel of an .ncu-rep (ncu --set full --import-source on).

    python tools/ncu_by_function.py rep.ncu-rep lib.so kernel_substring [out.txt]

Joins the source page of the report (per-SASS-instruction samples) with the line info of the library that was
profiled (`nvdisasm -gi`: innermost source line of every instruction, also inside inlined device functions) and
attributes every instruction to the function whose definition contains that line in csrc/*.cu."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def function_table(path):
    """[(first_line, name)] of the function definitions in a .cu/.h file (column-0 heuristics)."""
    out = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|__host__\s+|__device__\s+|__global__\s+|__forceinline__\s+|__noinline__\s+)+"
                     r"[\w:<>\s\*&]*?\b(\w+)\s*\(")
    lines = open(path, errors="replace").read().split("\n")
    for i, ln in enumerate(lines, 1):
        mk = re.match(r"^\s+(\w+_kernel)\(const Workspace", ln)  # kernel name on the line after its __launch_bounds__
        if mk:
            out.append((i, mk.group(1)))
            continue
        if ln[:1] in (" ", "\t", "/", "#", "}", ""):
            continue
        m = pat.match(ln)
        if m and m.group(1) not in ("if", "for", "while", "switch", "__launch_bounds__"):
            out.append((i, m.group(1)))
        elif "__launch_bounds__" in ln:
            m2 = re.search(r"\)\s*(\w+)\s*\(", ln)
            if m2:
                out.append((i, m2.group(1)))
    return out


def main(rep, lib, kernel_sub, out=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    line_of = {}   # offset -> (file, line)
    for cub in sorted(os.listdir(tmp)):
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cur_fn, cur_line, inside = None, None, False
        for ln in dis.split("\n"):
            if ln.startswith("\t.section\t.text."):
                cur_fn = ln.split(".text.", 1)[1].split(",")[0]
                inside = kernel_sub in cur_fn and not line_of
                pending = None
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                if "inlined at" in ln and pending is not None and pending[2]:
                    continue
                # the innermost location comes first; its "inlined at" continuation lines follow
                if pending is None or not pending[2]:
                    pending = (m.group(1), int(m.group(2)), True)
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
            if m:
                off = int(m.group(1), 16)
                if pending is not None:
                    cur_line = (pending[0], pending[1])
                    pending = (pending[0], pending[1], False)
                line_of[off] = cur_line
        if line_of:
            break
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # find the kernel's block
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and kernel_sub.split("IL")[0].replace("_ZN5ilcc2", "").lstrip("0123456789") in r[1].replace("::", ""):
            start = i
            break
    if start is None:
        start = 0
    hdr = rows[start + 1]
    col = {h: i for i, h in enumerate(hdr)}
    tables = {}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: defaultdict(float))
    by_line = defaultdict(float)
    base = None
    total = 0.0
    ninstr = 0
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        off = addr - base
        loc = line_of.get(off)
        fn = "?"
        if loc:
            f, l = loc
            if f not in tables:
                tables[f] = function_table(f) if os.path.exists(f) else []
            for first, name in tables[f]:
                if first <= l:
                    fn = name
                else:
                    break
            if not tables[f]:
                fn = os.path.basename(f)
        smp = float(r[col["# Samples"]] or 0)
        agg[fn]["samples"] += smp
        agg[fn]["inst"] += float(r[col["Instructions Executed"]] or 0)
        agg[fn]["sass"] += 1
        for s in stalls:
            agg[fn][s] += float(r[col[s]] or 0)
        if loc:
            by_line[(os.path.basename(loc[0]), loc[1])] += smp
        total += smp
        ninstr += 1
    keys = ["stall_wait", "stall_long_sb", "stall_no_inst", "stall_selected", "stall_branch_resolving", "stall_short_sb",
            "stall_barrier", "stall_not_selected", "stall_math", "stall_dispatch"]
    lines = [f"total samples {total:.0f}  SASS instructions in kernel {ninstr}",
             f"{'function':26s} {'samples':>8s} {'%':>6s} {'Minst':>8s} {'sass':>5s} | " + " ".join(k.replace('stall_', '')[:6].rjust(6) for k in keys)]
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] < 0.003 * total:
            continue
        lines.append(f"{fn:26s} {a['samples']:8.0f} {100 * a['samples'] / total:5.1f}% {a['inst'] / 1e6:8.1f} {a['sass']:5.0f} | "
                     + " ".join(f"{a[k]:6.0f}" for k in keys))
    src_cache = {}
    for (f, l), s in sorted(by_line.items(), key=lambda kv: -kv[1])[:24]:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lidar_camera_calibration_b200", "csrc", f)
        if f not in src_cache:
            src_cache[f] = open(p, errors="replace").read().split("\n") if os.path.exists(p) else []
        text = src_cache[f][l - 1].strip()[:110] if 0 < l <= len(src_cache[f]) else ""
        lines.append(f"({f}:{l}) {s:.0f}  {text}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
