export BENCH_NO_CPU=1
for n in 0 32 40 48 64; do
  echo "== la-sms $n"; timeout 300 python bench.py --la-sms $n 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    j=json.loads(l); print('value',round(j['value'],1),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value'],1),'roof',round(j['roofline']['frac'],3), j['config'].get('sm_partition'))
except Exception as e: print('ERR',l[-600:])
"
done
for n in 32 48; do for p in main la; do echo "== la-sms $n part $p"; BENCH_PARTS=$p timeout 300 python bench.py --la-sms $n 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('ms',round(j['ms_per_step'],3))"; done; done
