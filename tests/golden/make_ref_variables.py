"""Extracts the reference's own variable listing (name, shape, per-scope counts, total) from
`/root/reference/notebooks/play.ipynb` (cell output at :239-362) into ref_variables.json.

Run once in the build container (the GPU box has no /root/reference):
    python tests/golden/make_ref_variables.py
"""
import json
import os
import re

SRC = '/root/reference/notebooks/play.ipynb'
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_variables.json')

nb = json.load(open(SRC))
text = None
for cell in nb['cells']:
    for out in cell.get('outputs', []):
        t = ''.join(out.get('text', []))
        if 'Trainable Variables:' in t:
            text = t
assert text is not None
variables, scopes = {}, {}
for line in text.splitlines():
    m = re.match(r'^\t(\S+):0 \[(.*)\]$', line)
    if m:
        shape = [int(s) for s in m.group(2).split(',')] if m.group(2).strip() else []
        variables[m.group(1)] = shape
    m = re.match(r'^(\w+) scope params = ([\d ]+)$', line)
    if m:
        scopes.setdefault(m.group(1), 0)
        scopes[m.group(1)] += int(m.group(2).replace(' ', ''))
    m = re.match(r'^Number of trainable parameters: ([\d ]+)$', line)
    if m:
        total = int(m.group(1).replace(' ', ''))
# the notebook prints the last `sequence` variable in a separate block without a count line
json.dump(dict(source='akosiorek/sqair@474f5d0 notebooks/play.ipynb:239-362', n_steps_per_image=3,
               variables=variables, scope_counts_printed=scopes, total=total), open(DST, 'w'), indent=1)
print(len(variables), 'variables, total', total, scopes)
