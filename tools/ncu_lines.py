#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go, by source line, without bringing the (35 MB) report home.

    python tools/ncu_lines.py <report.ncu-rep> <library.so> [top N] > profiles/<tag>_lines.txt

The SASS page of the report (`ncu --page source --csv`) carries per-instruction counters; `nvdisasm -g` of the same
cubin (extracted from the library with cuobjdump) carries the file:line of every instruction (-lineinfo).  The two are
joined by instruction offset and summed per source line (innermost inlined frame)."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return kernels


def _norm(name):
    n = re.sub(r"\s+", "", name.replace("void ", "", 1))
    n = re.sub(r"\((int|bool)\)", "", n)
    return n.replace("false", "0").replace("true", "1")


def demangled_to_lines(so, kernel_name):
    """offset -> (file, line) for the kernel whose demangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        fn, line, table, tables = None, None, None, {}
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                fn = m.group(1)
                table = tables.setdefault(fn, {})
                line = None
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and table is not None:
                table[int(m.group(1), 16)] = (line, m.group(2).strip())
        if not tables:
            continue
        names = subprocess.run(["c++filt"] + list(tables), capture_output=True, text=True).stdout.splitlines()
        for mangled, dem in zip(tables, names):
            if _norm(dem) == _norm(kernel_name):
                return tables[mangled]
    return {}


def main():
    rep, so = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    for k in sass_rows(rep):
        hdr = k["hdr"]
        ia, ii, ith, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        lines = demangled_to_lines(so, k["name"])
        base = int(k["rows"][0][ia], 16)
        per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), 0])
        tot_i = tot_s = 0
        for r in k["rows"]:
            off = int(r[ia], 16) - base
            loc, _ = lines.get(off, (None, ""))
            key = loc or ("?", 0)
            n, th, smp = int(r[ii] or 0), int(r[ith] or 0), int(r[isamp] or 0)
            p = per[key]
            p[0] += n; p[1] += th; p[2] += smp; p[4] += 1
            for i, h in stall_cols:
                p[3][h] += int(r[i] or 0)
            tot_i += n; tot_s += smp
        print("kernel:", k["name"])
        print("  warp-instructions %d, stall samples %d, SASS instructions %d, lines with line info %d" % (tot_i, tot_s, len(k["rows"]), len(per)))
        print("  %-28s %6s %7s %6s %7s  %s" % ("file:line", "SASS", "inst %", "lanes", "smpl %", "top stalls"))
        for key, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
            st = ", ".join("%s %.0f%%" % (h[6:], 100.0 * c / max(p[2], 1)) for h, c in p[3].most_common(3) if c)
            print("  %-28s %6d %7.2f %6.1f %7.2f  %s" % ("%s:%d" % key, p[4], 100.0 * p[0] / max(tot_i, 1), p[1] / max(p[0], 1), 100.0 * p[2] / max(tot_s, 1), st))


if __name__ == "__main__":
    main()
