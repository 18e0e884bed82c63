mkdir -p gpurun_out
run() { echo "== $*"; timeout 400 python bench.py "$@" 2>gpurun_out/err.txt | tail -1 | python -c "
import sys,json
s=sys.stdin.read().strip()
try:
    d=json.loads(s); print({k:d.get(k) for k in ('value','unit','ms_per_step','final_loss')}, d.get('e2e',{}).get('value'), d.get('round'), (d.get('sweep') or '')[:8] if isinstance(d.get('sweep'),list) else '')
except Exception as e:
    print('PARSE FAIL', s[:300])
"; tail -3 gpurun_out/err.txt | cut -c1-300; }
run --config 3 --steps 20
run --config 3 --steps 5 --impl reference
run --config 4 --steps 5
run --config 4 --steps 2 --impl reference
run --config 5 --sweep-dims 1e5,1e7 --sweep-reps 10
run --config 5 --sweep-dims 1e5,1e7 --sweep-reps 10 --impl reference
run --steps 50
run --steps 10 --impl reference
