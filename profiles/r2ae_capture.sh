# round 2, capture AE (1 GPU): trajectories per CTA of the headline unit with the tensor-core transform in place (the 256-thread
# choice was made on the FMA form, r2n/r2q): 128 / 256 / 512 threads per CTA, two passes each
set -x
O=gpurun_out/r2ae
mkdir -p $O
for pass in 1 2; do
for v in _ctpb128 "" _ctpb512; do   # (run one variant at a time: three 219 MB libraries exceed the snapshot limit)
  CRCL_LIB_PATH=/root/repo/caracal_b200/libcaracal_gpu$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v.$pass.json 2> $O/bench$v.$pass.err
  python - <<PY
import json
d = json.loads(open("$O/bench$v.$pass.json").read().strip().splitlines()[-1])
print("variant '$v' pass $pass ms/step %.3f value %.4e" % (d["ms_per_step"], d["value"]))
PY
done
done
