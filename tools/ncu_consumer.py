import csv, subprocess, sys, re, collections, io
rep, nsym = sys.argv[1], float(sys.argv[2])
maxthr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, E, A = ix['# Samples'], ix['Instructions Executed'], ix['Avg. Threads Executed']
ops = collections.Counter(); osamp = collections.Counter(); tot = 0; ts = 0
for r in data:
    if int(r[E]) > 0 and float(r[A] or 0) <= maxthr:
        m = re.match(r'\s*(@!?U?P\d\s+)?([A-Z0-9_]+)', r[1]); op = m.group(2) if m else '?'
        ops[op] += int(r[E])/nsym; osamp[op] += int(r[S]); tot += int(r[E])/nsym; ts += int(r[S])
print("consumer instr/symbol %.1f, samples %d" % (tot, ts))
for op, c in ops.most_common(30):
    print("%-8s %6.1f  samples %7d (%.1f%%)" % (op, c, osamp[op], 100.0*osamp[op]/ts))
if len(sys.argv) > 4:
    for i, r in enumerate(data):
        if int(r[E]) > 0 and float(r[A] or 0) <= maxthr:
            print("%4d %-64s %.2f %s" % (i, r[1].strip()[:64], int(r[E])/nsym, r[S]))
