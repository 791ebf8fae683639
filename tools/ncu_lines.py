"""Joins an ncu source-page CSV (per-SASS-instruction samples / executed counts) with `nvdisasm -gi`
line info (inline chains) and aggregates by the innermost source line of the per-block program.

usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call('cd %s && cuobjdump -xelf all %s > /dev/null' % (tmp, os.path.abspath(lib)), shell=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run('nvdisasm -gi -c %s' % os.path.join(tmp, cubin), shell=True, capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and kname in l][0]
end = start + 1
while end < len(dis) and not dis[end].startswith('//--------------------- .'):
    end += 1
src_path = os.path.join(os.path.dirname(os.path.abspath(lib)), 'sqair_device.cuh')
src = open(src_path).read().split('\n')
lblock = [i for i, l in enumerate(src) if l.startswith('struct Block')][0] + 1
chain, seq, pending = [], [], []
for l in dis[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if pending:
            chain, pending = pending, []
        seq.append((chain, m.group(2)))
csvtxt = subprocess.run('ncu -i %s --page source --csv' % rep, shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.split('\n')))
hdr = [r for r in rows if '# Samples' in r][0]
data = rows[rows.index(hdr) + 1:]
data = [r for r in data if len(r) == len(hdr)]
isamp, iex = hdr.index('# Samples'), hdr.index('Instructions Executed')
assert len(data) == len(seq), (len(data), len(seq))
agg_s, agg_e = collections.Counter(), collections.Counter()
for (ch, _), r in zip(seq, data):
    key = None
    for f, ln in ch:                      # innermost first
        if f == 'sqair_device.cuh' and ln >= lblock:
            key = ln
            break
    if key is None:
        key = ch[0][1] if ch and ch[0][0] == 'sqair_device.cuh' else -1
    agg_s[key] += int(r[isamp] or 0)
    agg_e[key] += int(r[iex] or 0)
ts, te = sum(agg_s.values()), sum(agg_e.values())
print('total samples %d, instructions %d' % (ts, te))
for k, v in agg_s.most_common(top):
    print('%6.2f%% samp %6.2f%% inst  L%-5d %s' % (100 * v / ts, 100 * agg_e[k] / te, k, src[k - 1].strip()[:100] if k > 0 else '?'))
