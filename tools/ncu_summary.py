"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + consumer/producer split of the source page."""
import csv, subprocess, sys, re, collections, io
rep = sys.argv[1]
nsym_per_cta = float(sys.argv[2]) if len(sys.argv) > 2 else None   # symbols per consumer warp (for per-symbol counts)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit", "sm__inst_executed_pipe_fma", "smsp__issue_active.avg.pct",
        "sm__throughput.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__waves_per_multiprocessor",
        "sm__pipe_fma_cycles_active", "sm__inst_executed_pipe_fp64", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "sm__pipe_alu_cycles_active", "smsp__inst_executed_pipe_lsu", "dram__throughput.avg.pct", "launch__shared_mem_per_block_dynamic"]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and "per_second" not in h:
        print("%-75s %-14s %s" % (h, rows[1][i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, E, A = ix['# Samples'], ix['Instructions Executed'], ix['Avg. Threads Executed']
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "total warp instr", sum(int(r[E]) for r in data))
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in data:
    for h in stall:
        agg[h] += int(r[ix[h]])
print("stalls:", [(k, v) for k, v in agg.most_common(8)])
top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for i in sorted(top):
    r = data[i]
    st = sorted(((h, int(r[ix[h]])) for h in stall if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:2]
    print("%4d %-56s thr=%-5s exec=%-10s samp=%-8s %s" % (i, r[1].strip()[:56], r[A], r[E], r[S], st))
