python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" _mb4 _mb5 _mb6 _mb8; do
  CRCL_LIB_PATH=/root/repo/caracal_b200/libcaracal_gpu$v.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant$v', 'ms/step %.2f'%d['ms_per_step'], 'value %.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'])"
done
