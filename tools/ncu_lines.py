"""Executed instructions and warp-stall samples per SOURCE LINE of a third-generation SPARC kernel.
ncu's source page lists SASS instructions in program order without line numbers when read as CSV; nvdisasm -g gives the
line of every SASS instruction of the same cubin: join the two by position.
  ncu -i gpurun_out/X.ncu-rep --page source --csv > /tmp/src.csv        (capture taken with --import-source on)
  python tools/ncu_lines.py bwd3 /tmp/src.csv [n_ctas]"""
import collections, csv, os, re, subprocess, sys, tempfile

which, src_csv = sys.argv[1], sys.argv[2]
ncta = float(sys.argv[3]) if len(sys.argv) > 3 else 148.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(root, "clip_finegrained_alignment_b200", "csrc")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "clip_finegrained_alignment_b200", "libcfa_b200.so")],
               cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, "sparc_tc_%s.sm_100a.cubin" % which)
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
kname = "sparc_%s_kernelILi80" % which                      # flagship instantiation <80, 208, 512, *>
body = [f for f in re.split(r"\n\s*\.text\.", out) if kname in f.split("\n")[0] and "Lb0E" in f.split("\n")[0]][0]
line, seq = None, []
for ln in body.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", ln):
        seq.append(line)
rows = list(csv.reader(open(src_csv)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "ins": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 6:
        cur["ins"].append(r)
k = [k for k in kernels if ("sparc_%s" % which) in k["name"]][0]
h = k["hdr"]
ie, ss = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
n = min(len(seq), len(k["ins"]))
print("SASS instructions: nvdisasm %d, ncu %d" % (len(seq), len(k["ins"])))
agg, samp = collections.Counter(), collections.Counter()
for i in range(n):
    agg[seq[i]] += int(k["ins"][i][ie])
    samp[seq[i]] += int(k["ins"][i][ss])
tot, ts = sum(agg.values()), sum(samp.values())
print("warp instructions executed per CTA: %.0f   stall samples: %d" % (tot / ncta, ts))
cache = {}
def text(f, l):
    p = os.path.join(csrc, f)
    if f not in cache:
        cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
for key, c in sorted(agg.items(), key=lambda kv: -samp[kv[0]])[:40]:
    f, l = key if key else ("?", 0)
    print(f"{c / ncta:9.0f} inst/CTA {100 * c / tot:5.1f}%   samples {100 * samp[key] / ts:5.1f}%   {f}:{l}  {text(f, l)}")
