"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report (needs -lineinfo).
usage: python tools/ncu_lines.py <report.ncu-rep> <cubin-name-substring> <mangled-kernel-substring> <source.cu> [top]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, cubin_sub, kern_sub, src_path = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 35
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
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
kfilter = os.environ.get("NCU_KERNEL")          # e.g. regex:loss_flat when the report holds several kernels
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", kfilter] if kfilter else [])
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ii, si = h.index("Instructions Executed"), h.index("# Samples")
base = int(rows[hi + 1][0], 16)
cnt, smp = {}, {}
for r in rows[hi + 1:]:
    try:
        off, n, s = int(r[0], 16) - base, int(r[ii]), int(r[si])
    except (ValueError, IndexError):
        continue
    ln = off2line.get(off)
    cnt[ln] = cnt.get(ln, 0) + n
    smp[ln] = smp.get(ln, 0) + s
tot, tots = sum(cnt.values()), max(sum(smp.values()), 1)
src = open(src_path).read().split("\n")
print(f"total warp instructions {tot}, samples {tots}")
for ln, n in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    text = src[ln[1] - 1].strip()[:100] if ln and ln[0] == os.path.basename(src_path) else str(ln)
    print(f"{100 * n / tot:5.1f}% inst {100 * smp[ln] / tots:5.1f}% smp  {ln[1] if ln else '?':>4}: {text}")
