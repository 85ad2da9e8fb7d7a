#!/bin/bash
# Round-2 GPU batch: parity tests, the headline bench, the ncu launch list of the step and one --set full capture of
# every hot kernel.  Results under gpurun_out/ (tools/summarize_ncu.py turns them into profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 120 python tools/probe_rope.py > gpurun_out/${TAG}_probe_rope.log 2>&1; cat gpurun_out/${TAG}_probe_rope.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python tools/run_kernels.py > gpurun_out/${TAG}_run_kernels.log 2>&1; echo "run_kernels rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-variants > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
# the report of ~25 kernels with source exceeds what gpurun brings back (64 MiB): keep it on the box, bring the raw page as CSV
timeout 1200 ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_${TAG} -f python tools/run_kernels.py \
  > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null; ls -la /tmp/prof_${TAG}.ncu-rep gpurun_out/prof_${TAG}_raw.csv
# attention core alone, with source (small report)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bf16_kp_kernel --launch-skip 2 -c 1 -o gpurun_out/fa_${TAG} -f \
  python tools/run_attn_bf16.py 12 32760 32760 3 > gpurun_out/${TAG}_ncu_fa.log 2>&1; echo "ncu fa rc=$?"
du -sh gpurun_out
