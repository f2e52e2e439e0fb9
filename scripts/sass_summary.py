"""Per-kernel SASS opcode summary of the built library (evidence that the hot kernels use tcgen05 / tensor memory):
   python scripts/sass_summary.py [lib] > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'quinn_b200', 'lib', 'libquinn_b200.so')
WATCH = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'HMMA', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'DFMA', 'MUFU', 'F2FP', 'FHADD',
         'SYNCS', 'BAR', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = {}
names = re.findall(r'Function : (\S+)', out)
dm = subprocess.run(['cu++filt'] + names, capture_output=True, text=True).stdout.split('\n') if names else []
for n, d in zip(names, dm):
    demangle[n] = re.sub(r'\(.*', '', re.sub(r'\((int|bool)\)', '', d)).replace('void ', '')
cur = None
counts = collections.OrderedDict()
for line in out.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);', line)
    if m and cur:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        counts[cur][op.split('.')[0]] += 1
        counts[cur]['_total'] += 1
print(f'# {os.path.basename(lib)}: SASS opcode counts per kernel (static).  tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR,')
print('# tcgen05.ld/st = LDTM/STTM, mixed-precision fp16 subtract = FHADD, cvt.f16x2 = F2FP, spills = LDL/STL.')
print(f'{"kernel":58s} {"total":>7s} ' + ' '.join(f'{w:>7s}' for w in WATCH))
for k, c in counts.items():
    name = demangle.get(k, k)[:58]
    print(f'{name:58s} {c["_total"]:7d} ' + ' '.join(f'{c[w]:7d}' for w in WATCH))
