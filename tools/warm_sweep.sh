# warm-start experiments (round 2): c3 / c5 time and tree hash, and the stress test of the certificate, per configuration
for cfg in "2048 2e-4" "512 2e-4" "8192 2e-4"; do
  set -- $cfg
  export TMC_WARM_CHILD_ROWS=$1 TMC_CERT_S0=$2
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/sw_c3.json 2> gpurun_out/sw_c3.err
  python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/sw_c5.json 2> gpurun_out/sw_c5.err
  python tests/manual/warm_stress.py 24 > gpurun_out/sw_stress.log 2>&1
  python - <<PY
import json
r=[]
for f in ("gpurun_out/sw_c3.json","gpurun_out/sw_c5.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r.append((round(d["value"],1), d["tree"]["lanczos_steps"], d["tree"]["sweeps"], d["tree"]["hash"][:8]))
print("warm child rows $1 s0 $2:", r, open("gpurun_out/sw_stress.log").read().strip().splitlines()[-1][:120], flush=True)
PY
done
