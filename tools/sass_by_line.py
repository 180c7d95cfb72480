"""Attribute the executed SASS instructions of one profiled kernel to CUDA source lines.

    python tools/sass_by_line.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [units]

Joins `ncu --page source --csv` (per-SASS-instruction executed counts, keyed by address) with
`nvdisasm -g` line tables of the same build (compile with -lineinfo).  `units` divides the counts
(e.g. number of warp-tiles) so the output reads "instructions per thread per tile"."""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def main():
    rep, obj, sub = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, check=True, capture_output=True)
        cubin = next(Path(td).glob("*.cubin"))
        sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith("//--------------------- .text.") and sub in l)
    end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith("//--------------------- .text.")), len(sass))
    cur, ins = None, []
    for l in sass[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    ismp = hdr.index("# Samples")
    data = [r for r in rows[2:] if r and r[ia].startswith("0x")]
    base = int(data[0][ia], 16)
    prof = {int(r[ia], 16) - base: (r[isrc].strip(), int(r[iex]), int(r[ismp] or 0)) for r in data}
    op = lambda t: [w for w in t.split() if not w.startswith("@")][0]
    agg, aggop, mism, total = collections.Counter(), collections.defaultdict(collections.Counter), 0, 0
    smp = collections.Counter()
    for off, txt, line in ins:
        if off in prof:
            ptxt, ex, ns = prof[off]
            smp[line] += ns
            mism += op(ptxt) != op(txt)
            agg[line] += ex
            aggop[line][op(txt).split(".")[0]] += ex
            total += ex
    print(f"{len(ins)} SASS instructions, {mism} opcode mismatches vs the report (must be 0: same build?), total {total / units:.1f} per unit")
    cache = {}
    for (f, l), ex in sorted(agg.items(), key=lambda kv: -kv[1])[:80]:
        if f not in cache:
            try:
                cache[f] = Path(f).read_text().split("\n")
            except Exception:
                cache[f] = []
        s = cache[f][l - 1].strip()[:90] if l - 1 < len(cache[f]) else ""
        ops = " ".join(f"{k}:{v / units:.0f}" for k, v in aggop[(f, l)].most_common(4))
        print(f"{ex / units:7.1f} {100.0 * smp[(f, l)] / max(1, sum(smp.values())):5.1f}%smp {Path(f).name}:{l}: {s}   [{ops}]")


if __name__ == "__main__":
    main()
