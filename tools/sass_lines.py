"""Attribute SASS opcodes of one kernel to source functions/lines (needs -lineinfo; no GPU).
   python tools/sass_lines.py KERNEL OPCODE_REGEX [--lines]"""
import collections, re, subprocess, sys, os, tempfile
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.environ.get("SASS_LIB") or os.path.join(root, "myosuite_b200", "libmyo_b200.so")
kernel, opre = sys.argv[1], re.compile(sys.argv[2]); by_line = "--lines" in sys.argv
tmp = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
src = {f: open(os.path.join(root, "myosuite_b200", "csrc", f)).read().split("\n") for f in ("myo_b200.cu", "myo_device.cuh", "myo_solver.cuh")}
def func(f, ln):
    if f not in src: return f
    for i in range(min(ln, len(src[f])) - 1, -1, -1):
        s = src[f][i]
        mm = re.match(r"\s*(template.*)?(extern \"C\" )?__(device|global)__.*?(\w+)\s*\(", s)
        if mm and not s.strip().startswith("//"): return mm.group(4)
    return "?"
cur, kern, cnt = None, None, collections.Counter()
for l in dis.splitlines():
    if l.startswith(".text."): kern = l.strip()[6:-1]
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur and kern == kernel and opre.match(m.group(1)):
        cnt[(cur if by_line else (func(*cur),))] += 1
for k, v in cnt.most_common(40): print(v, *k)
