#!/bin/bash
# round-2 GPU call A: full GPU test suite, smoke, bench (N=1), 3-group experiment, ncu launch list + DRAM traffic
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02a_pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r02a_smoke.log
echo "== bench"; timeout 600 python bench.py > $O/r02a_bench_cfg2.json 2> $O/r02a_bench_cfg2.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02a_bench_cfg2.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "launches", d["gpu_launches"])
    print("eager", d.get("gpu_eager_baseline"))
    for k in d["kernels"][:14]: print("  %-34s n=%d %.1f us/step %.0f GB/s %.0f TF" % (k["name"], k["launches_per_step"], 1e3*k["ms_per_step"], k["GBps"], k["TFLOPs"]))
except Exception as e:
    print("bench parse failed", e)
PY
echo "== 3-group experiment (guarded TMEM)"; CIDNET_LIB=$PWD/hvi-cidnet_b200/libcidnet_b200_g3.so timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py tests/test_lca_gpu.py -m gpu -q > $O/r02a_g3_pytest.log 2>&1; echo "g3 rc=$?"; tail -5 $O/r02a_g3_pytest.log
CIDNET_LIB=$PWD/hvi-cidnet_b200/libcidnet_b200_g3.so timeout 300 python bench.py --steps 20 > $O/r02a_bench_g3.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02a_bench_g3.json').read().strip().splitlines()[-1]); print('g3 ms/step', d['ms_per_step'])"
echo "== ncu launch list + traffic"
export CIDNET_NO_GRAPH=1
timeout 300 python scripts/prof_forward.py 1 640 1120 3 $O/r02a_marks.txt > $O/r02a_prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 178 -c 89 --csv \
    --log-file $O/r02a_launches_traffic.csv python scripts/prof_forward.py 1 640 1120 3 > $O/r02a_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02a_ncu.log
