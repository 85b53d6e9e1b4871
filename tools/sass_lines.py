"""Development aid: executed warp-instructions per source line of DeblockTiledKernel.
usage: sass_lines.py <nvdisasm -g -c output> <ncu --page source --csv output> [top N] [mangled-name substring]
The two listings have the same instruction order; the nvdisasm one carries the line info."""
import collections
import csv
import os
import re
import sys

dis, srccsv = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
which = sys.argv[4] if len(sys.argv) > 4 else ''  
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('//--------------------- .text._ZN3hdb18DeblockTiledKernel') and which in l][0]
seq, cur = [], None
for l in lines[start + 1:]:
    if l.startswith('//---------------------'):
        break
    if l.strip().startswith('//## File'):
        m = re.findall(r'"([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m[0][0]), int(m[0][1]))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia, it = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
isrc = hdr.index('Source')
data = rows[2:]
assert len(data) == len(seq), (len(data), len(seq))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
pipe = collections.Counter()
for (loc, txt), r in zip(seq, data):
    a, t = int(r[ia]), int(r[it])
    tot += a
    agg[loc][0] += a
    agg[loc][1] += t
    op = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip()).split()[0].rstrip(';')
    agg[loc][2][op.split('.')[0]] += a
    pipe[op] += a
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcs = {}
def text(loc):
    if not loc:
        return ''
    f, n = loc
    if f not in srcs:
        for d in ('gpu_video_codec_b200/csrc',):
            p = os.path.join(root, d, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().split('\n')
                break
        else:
            srcs[f] = []
    s = srcs[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ''
print('total warp-instructions', tot)
for loc, (a, t, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.2f%% lanes %4.1f  %-22s %s   [%s]' % (100.0 * a / tot, t / max(a, 1), '%s:%d' % loc if loc else '?', text(loc),
                                                   ' '.join('%s:%d' % (k, 1000 * v // tot) for k, v in ops.most_common(4))))
print('--- opcodes (per mille of executed warp-instructions)')
for op, a in pipe.most_common(40):
    print('%6.1f  %s' % (1000.0 * a / tot, op))
