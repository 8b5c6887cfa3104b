#!/usr/bin/env python3
"""Instruction ledger of one profiled kernel: warp-instructions executed per CUDA function and per category, from the
source page of an .ncu-rep captured with --import-source on (read here, no GPU).
usage: tools/ncu_ledger.py <rep> <frames in the launch> <git commit the profiled library was built from>"""
import csv, io, re, subprocess, sys

rep, frames, commit = sys.argv[1], float(sys.argv[2]), sys.argv[3]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))

CATEGORY = [
    ("flat spans (per pixel)", r"draw_plane_warp|flat_offset|plane_dir|plane_row|plane_u"),
    ("wall columns (per pixel)", r"wall_fast_loop|wall_sel|wall_advance|wall_acc29|pick_byte|draw_wall_warp|wall_row|wall_tbase|lit_index|tex_interleaved"),
    ("framebuffer stores", r"put_px|store_batch|row_mask|lds_u32"),
    ("texel loads", r"tex_ld|__ldg|ldg"),
    ("per (strip, seg) set-up: projection, clipping, light", r"column_eval|yrow|clampv|light_row|udiv|div|scale|recip|norm|bitlen|seg_|mul"),
    ("sky / void", r"draw_sky_warp|fill_void_warp|sky"),
    ("kernel frame: work list, windows, colormap staging", r"b2d_raster_kernel|masked"),
]


def functions_of(path):
    text = subprocess.run(["git", "show", "%s:%s" % (commit, path)], capture_output=True, text=True).stdout.split("\n")
    starts = []
    for i, line in enumerate(text, 1):
        if re.match(r"^(template|__device__|__global__|B2D_HD|static|inline|constexpr [a-z_0-9]+ \w+\()", line) and "(" in line or \
           (i > 1 and re.match(r"^template", text[i - 2]) and "(" in line):
            m = re.search(r"(\w+)\s*\(", line.split("__launch_bounds__")[0] if "__launch_bounds__" in line else line)
            if "__global__" in line:
                j = i
                while j < len(text) and "(" not in text[j].replace("__launch_bounds__(", ""):
                    j += 1
                m = re.search(r"(\w+)\s*\(", text[j].replace("__launch_bounds__(", "")) if j < len(text) else m
            if m and m.group(1) not in ("defined", "if", "for", "while", "switch", "sizeof"):
                starts.append((i, m.group(1)))
    return starts


per_fn = {}
cur_file, starts = None, []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        rel = cur_file.split("/root/repo/")[-1] if "/root/repo/" in cur_file else None
        starts = functions_of(rel) if rel else []
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    v = r[hdr.index("Instructions Executed")]
    n = int(v) if v.strip().isdigit() else 0
    if not n:
        continue
    line = int(r[0])
    fn = "(%s)" % cur_file.split("/")[-1]
    for s, name in starts:
        if s <= line:
            fn = name
    per_fn[fn] = per_fn.get(fn, 0) + n

tot = sum(per_fn.values())
cats = {}
for fn, n in per_fn.items():
    cat = next((c for c, pat in CATEGORY if re.search(pat, fn)), "other (intrinsics headers)")
    cats.setdefault(cat, []).append((n, fn))
print("| category | warp-instructions per frame | share | functions (warp-instructions per frame) |")
print("|---|---|---|---|")
for cat, items in sorted(cats.items(), key=lambda kv: -sum(n for n, _ in kv[1])):
    s = sum(n for n, _ in items)
    print("| %s | %.0f | %.1f %% | %s |" % (cat, s / frames, 100.0 * s / tot, ", ".join("`%s` %.0f" % (fn, n / frames) for n, fn in sorted(items, reverse=True))))
print("| **total attributed** | **%.0f** | 100 %% | |" % (tot / frames))
