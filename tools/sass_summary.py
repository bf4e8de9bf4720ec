"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA, packed fp32 -> FFMA2.
usage: python tools/sass_summary.py [hfa_gp_b200/libhfagp_sm100.so] > profiles/rN_sass_summary.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else 'hfa_gp_b200/libhfagp_sm100.so'
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'FFMA2', 'FFMA', 'MUFU', 'LDGSTS', 'RED', 'ATOM']
fn, counts, total = None, collections.OrderedDict(), {}
for ln in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', ln)
    if m:
        fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r'\(.*', '', fn).replace('void ', '').replace('hfagp::', '')
        counts[fn] = collections.Counter(); total[fn] = 0
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
    if m and fn:
        op = m.group(2)
        total[fn] += 1
        for p in pats:
            if op.split('.')[0] == p or (p in ('UTCHMMA', 'UTCBAR', 'UTMALDG', 'SYNCS') and op.startswith(p)):
                counts[fn][p] += 1
print(f'# cuobjdump -sass {so}: static instruction counts per kernel (sm_100a)')
print(f'{"kernel":44s} {"instrs":>7s} ' + ' '.join(f'{p:>7s}' for p in pats))
for f, c in counts.items():
    print(f'{f[:44]:44s} {total[f]:7d} ' + ' '.join(f'{c[p]:7d}' for p in pats))
