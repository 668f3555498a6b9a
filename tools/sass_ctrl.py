"""Print a kernel's SASS with decoded scheduling control fields (Volta+ 128-bit encoding):
stall count, yield, write/read scoreboard index, wait mask.  usage: sass_ctrl.py lib.so kernel_substring"""
import re, subprocess, sys
so, name = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout.split('\n')
on = False; pend = None
for l in txt:
    if 'Function :' in l:
        on = name in l
        continue
    if not on: continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', l)
    if m:
        pend = (m.group(1), m.group(2)); continue
    m = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', l)
    if m and pend:
        hi = int(m.group(1), 16)
        stall = (hi >> 41) & 0xF; yld = (hi >> 45) & 1; wb = (hi >> 46) & 7; rb = (hi >> 49) & 7; wait = (hi >> 52) & 0x3F
        print(f"{pend[0]} st{stall:2d} {'Y' if yld else ' '} W{wb if wb != 7 else '-'} R{rb if rb != 7 else '-'} wait{wait:06b}  {pend[1][:90]}")
        pend = None
