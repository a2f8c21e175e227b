"""Top instructions of a kernel by stall samples, with the dominant stall reason, from an ncu source-page CSV."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit(): continue
    n = int(r[ix['# Samples']] or 0)
    st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols}
    data.append((n, r[ix['Source']].strip(), int(r[ix['Instructions Executed']] or 0), st))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
# opcode histogram
ops = collections.Counter(); opsamp = collections.Counter()
for n, src, ex, st in data:
    t = src.split(); op = (t[1] if t[0].startswith('@') else t[0])
    ops[op.split('.')[0]] += ex; opsamp[op.split('.')[0]] += n
texec = sum(ops.values())
print("opcode: executed% samples%")
for op, c in ops.most_common(14):
    print(f"  {op:10s} {100*c/texec:5.1f}  {100*opsamp[op]/tot:5.1f}")
# by-line agg of long scoreboard etc
for reason in sys.argv[2:]:
    print("top instrs for stall", reason)
    for n, src, ex, st in sorted(data, key=lambda d: -d[3].get(reason, 0))[:8]:
        print(f"   {st.get(reason,0):6d} {src[:90]}")
