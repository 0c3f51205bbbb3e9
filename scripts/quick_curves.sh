for v in fused torch; do
  F=""; [ $v = torch ] && F="--torch-ops"
  timeout 200 python -m marl_responsible_nav_b200.train $F --config custom_fear_10 --envs 4096 --steps 9000 --report 1000 2>&1 | grep env_steps > gpurun_out/r1o_curve_${v}_ops.log
done
python - <<'PY'
import json
for v in ("fused","torch"):
    print(v, [(d['env_steps'], round(d['agent_steps_per_s']/1e6,1), round(d['mean_return'],2), round(d['apples_per_episode'],3)) for d in map(json.loads, open(f'gpurun_out/r1o_curve_{v}_ops.log'))])
PY
