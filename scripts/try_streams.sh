cd /root/repo
for a in "--streams 1" "--streams 2" "--streams 2 --frames-per-step 4"; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline $a 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', round(d['value'],2), round(d['e2e']['value'],2), d['roofline']['frac'], d['gpu_launches'])"
done
