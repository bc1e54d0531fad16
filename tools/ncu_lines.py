"""Per-source-line totals of an .ncu-rep: joins the SASS source page (instructions executed, stall samples) with the
line table of the shipped cubin (nvdisasm -g), so the hottest CUDA lines of each kernel can be read without the GUI.

    python tools/ncu_lines.py report.ncu-rep [top] [kernel substring]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble():
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "libzseek_b200", "libzseek_b200.so")], cwd=d, capture_output=True)
    cub = [c for c in glob.glob(os.path.join(d, "*.cubin")) if "zsk_cuda" in c][0]
    return subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()


def line_table(dis, kname):
    start = next(i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and kname in l)
    m, cur = {}, None
    for l in dis[start + 1:]:
        if l.startswith("//--------------------- "):
            break
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if mm:
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    only = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    dis = disassemble()
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = [r[1], None, []]
            sections.append(cur)
        elif r and r[0] == "Address":
            cur[1] = r
        elif cur is not None and cur[1] is not None and len(r) == len(cur[1]):
            cur[2].append(r)
    srcs = {}
    for name, hdr, data in sections:
        kname = name.split("(")[0]
        if only not in kname:
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        m = line_table(dis, kname)
        base = int(data[0][0], 16)
        agg = collections.defaultdict(lambda: [0.0, 0.0, 0, 0.0])
        stall_cols = [h for h in hdr if h.startswith("stall") and "Not Issued" not in h]
        stalls = collections.defaultdict(float)
        for r in data:
            key = m.get(int(r[0], 16) - base)
            ie, sm, te = (float(r[ix[c]] or 0) for c in ("Instructions Executed", "# Samples", "Thread Instructions Executed"))
            a = agg[key]
            a[0] += ie; a[1] += sm; a[2] += 1; a[3] += te
            for h in stall_cols:
                stalls[h] += float(r[ix[h]] or 0)
        tot = sum(v[0] for v in agg.values()) or 1
        stot = sum(v[1] for v in agg.values()) or 1
        print(f"===== {kname}: {tot:.3e} warp instructions, {stot:.0f} samples")
        st = sum(stalls.values()) or 1
        print("stalls: " + ", ".join(f"{h[6:]} {100 * v / st:.0f}%" for h, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]))
        for key, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
            if key is None:
                print(f"{'(no line)':22s} inst {100 * v[0] / tot:5.1f}% samp {100 * v[1] / stot:5.1f}%")
                continue
            fn, ln = key
            if fn not in srcs:
                try:
                    srcs[fn] = open(os.path.join(ROOT, "libzseek_b200", "csrc", fn)).read().splitlines()
                except OSError:
                    srcs[fn] = []
            text = srcs[fn][ln - 1].strip()[:96] if ln - 1 < len(srcs[fn]) else ""
            print(f"{fn[:16]:16s}{ln:5d} inst {100 * v[0] / tot:5.1f}% samp {100 * v[1] / stot:5.1f}% sass {v[2]:4d} thr/inst {v[3] / max(v[0], 1):5.1f} | {text}")


if __name__ == "__main__":
    main()
