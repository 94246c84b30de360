#!/bin/bash
# One GPU call: parity tests, smoke, bench (ours + reference arm), ncu launch list + full captures.
#   scripts/gpu_round.sh [ncu] [notests]
set -u
mkdir -p gpurun_out
if [[ " $* " != *" notests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
fi
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [[ " $* " == *" ncu "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 20 --warmup 3 --no-extra --no-cpu --no-parity > gpurun_out/bench_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:matvec_panel -s 5 -c 2 -f -o gpurun_out/prof_matvec \
      python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-parity > gpurun_out/ncu_matvec.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gf_fault_fault|gf_fault_mantle|gf_mantle_fault|gf_mantle_mantle" \
      -c 12 -f -o gpurun_out/prof_assembly python scripts/assembly_probe.py > gpurun_out/ncu_assembly.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:class_matvec -c 3 -f -o gpurun_out/prof_classmv \
      python scripts/coupled_scaling.py --form classes --gf11 fft --steps 1 --warmup 1 --check-sources 0 > gpurun_out/ncu_classmv.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
