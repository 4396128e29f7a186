"""Instruction / stall-sample shares of one kernel per SOURCE REGION (line ranges) from an ncu report.
usage: python tools/ncu_regions.py <report> <cubin-substring> <kernel-substring> <source.cu> name:lo-hi [name:lo-hi ...]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, cubin_sub, kern_sub, src_path = sys.argv[1:5]
regions = []
for a in sys.argv[5:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "diffusionhandles_b200", "lib", "libdiffhandles_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cubin_sub in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern_sub in l and l.rstrip().endswith(":")][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//-----")), len(dis))
cur, off2line = None, {}
for l in dis[start:end]:
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        # inlined code: attribute to the outermost call site in the kernel's own file when nvdisasm gives "inlined at"
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m2 = re.findall(r'inlined at "(.*?)", line (\d+)', l)
        if m2:
            cur = (os.path.basename(m2[-1][0]), int(m2[-1][1]))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi_]
ii, si = h.index("Instructions Executed"), h.index("# Samples")
base = int(rows[hi_ + 1][0], 16)
acc = {n: [0, 0] for n, _, _ in regions}
acc["(other)"] = [0, 0]
base_src = os.path.basename(src_path)
for r in rows[hi_ + 1:]:
    try:
        off, n, s = int(r[0], 16) - base, int(r[ii]), int(r[si])
    except (ValueError, IndexError):
        continue
    ln = off2line.get(off)
    key = "(other)"
    if ln and ln[0] == base_src:
        for name, lo, hi in regions:
            if lo <= ln[1] <= hi:
                key = name
                break
    acc[key][0] += n; acc[key][1] += s
tot, tots = sum(v[0] for v in acc.values()), max(1, sum(v[1] for v in acc.values()))
print(f"total warp instructions {tot}, samples {tots}")
for k, (n, s) in acc.items():
    print(f"{k:>16}: {100 * n / tot:5.1f}% inst ({n / 1e6:6.2f} M)  {100 * s / tots:5.1f}% samples")
