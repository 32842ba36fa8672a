"""Summarise an .ncu-rep (run here, no GPU needed): headline metrics, stall mix, top source lines, per-file-range shares."""
import collections, csv, subprocess, sys
csv.field_size_limit(10**9)
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]
W = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__registers_per_thread",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
     "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_local_ld.sum"]
for w in W:
    if w in h:
        print("%-70s %s %s" % (w, [r[h.index(w)] for r in rows[2:]], rows[1][h.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()]); text = {}; cur = None; hdr = None
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); st = [(x, i) for i, x in enumerate(hdr) if x.startswith("stall_") and "Not" not in x]; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    k = (cur, int(r[0])); text[k] = r[1]
    try: agg[k][0] += int(r[iS]); agg[k][1] += int(r[iI])
    except ValueError: continue
    for x, i in st:
        try: agg[k][2][x] += int(r[i])
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
alls = collections.Counter()
for v in agg.values(): alls.update(v[2])
print("samples", tot, "warp-instructions", toti)
print("stall mix:", {k: round(100 * v / tot, 1) for k, v in alls.most_common(9)})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%-15s %4d samp %5.1f%% inst %5.1f%% %s | %s" % (f, l, 100 * v[0] / tot, 100 * v[1] / toti, dict(v[2].most_common(2)), text[(f, l)].strip()[:95]))
