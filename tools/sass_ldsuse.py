"""Static check for load->use serialisation: for every LDS/LDL/LD in a kernel, the distance (in SASS instructions) to the first
instruction that reads its destination register.  Distance 1-2 = the shared-memory latency (~29 cycles) is fully exposed.
   python tools/sass_ldsuse.py KERNEL [lib]"""
import collections, re, subprocess, sys, os, tempfile
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
kernel = sys.argv[1]; lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "myosuite_b200", "libmyo_b200.so")
tmp = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
src = {f: open(os.path.join(root, "myosuite_b200", "csrc", f)).read().split("\n") for f in ("myo_b200.cu", "myo_device.cuh", "myo_solver.cuh")}
def func(f, ln):
    if f not in src: return f
    for i in range(min(ln, len(src[f])) - 1, -1, -1):
        mm = re.match(r"\s*(template.*)?(extern \"C\" )?__(device|global)__.*?(\w+)\s*\(", src[f][i])
        if mm and not src[f][i].strip().startswith("//"): return mm.group(4)
    return "?"
ins, cur, kern = [], None, None
for l in dis.splitlines():
    if l.startswith(".text."): kern = l.strip()[6:-1]
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", l)
    if m and kern == kernel: ins.append((m.group(1), m.group(2), cur))
def regs(s, width):
    out = set()
    for r in re.findall(r"\bR(\d+)\b", s):
        out.add(int(r))
    return out
stat = collections.defaultdict(lambda: [0, 0, 0])
for i, (op, args, cur) in enumerate(ins):
    if not op.startswith(("LDS", "LDL")): continue
    dst = int(re.match(r"R(\d+)", args).group(1)) if re.match(r"R(\d+)", args) else None
    if dst is None: continue
    w = 2 if ".64" in op else (4 if ".128" in op else 1)
    d = {dst + k for k in range(w)}
    dist = None
    for j in range(i + 1, min(i + 40, len(ins))):
        a = ins[j][1]
        srcs = a.split(",", 1)[1] if "," in a and not ins[j][0].startswith(("ST", "BRA")) else a
        if regs(srcs, 1) & d: dist = j - i; break
    f = func(*cur) if cur else "?"
    stat[f][0] += 1
    if dist is not None and dist <= 2: stat[f][1] += 1
    if dist is not None and dist <= 6: stat[f][2] += 1
print("%-24s %6s %8s %8s" % ("function", "loads", "use<=2", "use<=6"))
for f, (n, a, b) in sorted(stat.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-24s %6d %8d %8d" % (f, n, a, b))
