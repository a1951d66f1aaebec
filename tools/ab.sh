#!/bin/bash
# A/B two builds of libsnb.so on the SAME box: tools/ab.sh <rounds>; expects lib/libsnb_A.so and lib/libsnb_B.so
L=semantic-nerf-for-satellite-data_b200/lib
R=${1:-2}
cp $L/libsnb.so /tmp/libsnb_keep.so
for r in $(seq 1 $R); do
  for v in A B; do
    cp $L/libsnb_$v.so $L/libsnb.so
    rm -f /tmp/ab_dump.txt
    SNB_PROF_DUMP=/tmp/ab_dump.txt python bench.py --steps 4 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ch=[l.split() for l in open('/tmp/ab_dump.txt') if l[0]!='#']
n=len(ch)//2
c=[float(x[7]) for x in ch[:n] if int(x[1])>=100]
w=[float(x[7]) for x in ch[:n] if int(x[1])<100]
print('$v round $r: rays/s %.0f  ms/step %.3f  frac %.4f  chains(us) %s  sum %.0f  other gemms %.0f  render %.1fM' % (d['value'], d['ms_per_step'], d['roofline']['frac'], ' '.join('%.0f'%x for x in c), sum(c), sum(w), d['render']['samples_per_s']/1e6))"
  done
done
cp /tmp/libsnb_keep.so $L/libsnb.so
