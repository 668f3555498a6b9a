"""Hot instructions of one kernel from an ncu report: python tools/ncu_hot.py report.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 150
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = [r for r in rows if r and r[0] == 'Address'][0]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[rows.index(hdr) + 1:]
tot = sum(int(r[ci['# Samples']]) for r in data)
stall = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in data) for h in stall}
print("samples", tot, "instr", sum(int(r[ci['Instructions Executed']]) for r in data))
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  ", h, v, round(100 * v / tot, 1))
for r in data:
    sm = int(r[ci['# Samples']])
    if sm >= thr:
        st = {h: int(r[ci[h]] or 0) for h in stall}
        k = max(st, key=st.get)
        print(r[ci['Address']][-4:], r[ci['Source']][:58].ljust(58), sm, int(r[ci['Instructions Executed']]) // 1000, k, st[k])
