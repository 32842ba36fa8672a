"""Static SASS statistics of the built library (no GPU needed): per kernel instruction count and the load/store mix.
   python tools/sass_stats.py [path/to/libmyo_b200.so]"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "myosuite_b200", "libmyo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, stats = None, collections.defaultdict(collections.Counter)
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: kern = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and kern:
        op = m.group(1); stats[kern]["total"] += 1
        base = op.split(".")[0]
        if base in ("LD", "ST", "LDS", "STS", "LDL", "STL", "LDG", "STG", "LDC", "SHFL", "BAR", "WARPSYNC", "IMAD", "LEA", "DFMA", "DMUL", "DADD", "MUFU", "CALL", "BRA", "FSEL", "SEL"):
            stats[kern][base] += 1
        if op.startswith("LD.E") or op.startswith("ST.E"): stats[kern]["generic"] += 1
for k, c in stats.items():
    print(k, dict(sorted(c.items(), key=lambda kv: -kv[1])))
