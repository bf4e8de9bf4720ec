"""Top CUDA source lines of an .ncu-rep by executed warp instructions / stall samples.
usage: python tools/ncu_lines.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []; fname = '?'
H = len(next(r for r in rows if 'Instructions Executed' in r))
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if len(r) > 8 and r[0].isdigit():
        try:   # source text with embedded quotes shifts the leading columns: index from the end
            lines.append((fname, int(r[0]), r[1], int(r[4 - H] or 0), int(r[7 - H] or 0)))
        except (ValueError, IndexError):
            pass
tot = sum(l[4] for l in lines); tots = max(1, sum(l[3] for l in lines))
print('total warp instr', tot, 'stall samples', tots)
for f, ln, src, st, ins in sorted(lines, key=lambda l: -l[4])[:N]:
    print(f'{ins / tot * 100:5.1f}% inst {st / tots * 100:5.1f}% stall  {f}:{ln:<4d} {src.strip()[:105]}')
