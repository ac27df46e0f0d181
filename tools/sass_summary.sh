#!/bin/bash
# SASS evidence per kernel (runs HERE, no GPU): mnemonic histogram of every kernel in the built objects -> profiles/<tag>_sass_<file>.txt
#   bash tools/sass_summary.sh r02
R=${1:-r02}
for f in conv_tc conv_tc4 conv_tc5 conv_s2_tc bottleneck_tc stem_tc dsconv_tc conv_mma dsconv_mma head_fused detect_post conv_direct elementwise hypergraph; do
  o=yolo_dbl_b200/build/$f.o
  [ -f $o ] || continue
  cuobjdump -sass $o 2>/dev/null | python3 -c "
import sys,re,collections
cur=None; hist=collections.OrderedDict()
for line in sys.stdin:
    m=re.match(r'\s*Function : (\S+)', line)
    if m:
        cur=m.group(1); hist[cur]=collections.Counter(); continue
    m=re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)', line)
    if m and cur: hist[cur][m.group(2)]+=1
KEY=('UTCHMMA','UTCQMMA','UTMALDG','UTMASTG','LDTM','STTM','UTCBAR','UTCCP','SYNCS','HMMA','LDGSTS','LDSM','FFMA2','FFMA','MUFU','REDUX','ATOM','RED','LDG','STG','LDS','STS','BAR','ELECT')
print('# SASS mnemonic counts per kernel of $f.o (cuobjdump -sass, sm_100a); Blackwell data-path instructions: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, SYNCS = mbarrier; HMMA = legacy mma.sync, LDGSTS = cp.async')
import subprocess
for k,h in hist.items():
    name=subprocess.run(['c++filt',k],capture_output=True,text=True).stdout.strip()
    tot=sum(h.values())
    agg=collections.Counter()
    for op,c in h.items():
        for key in KEY:
            if op==key or op.startswith(key+'.'): agg[key]+=c; break
    print(f'{name[:150]}\n    {tot} instructions: '+', '.join(f'{a} {agg[a]}' for a in KEY if agg[a]))
" > profiles/${R}_sass_$f.txt
done
ls -la profiles/${R}_sass_*
