#!/bin/bash
# One parameterised runner for every GPU session (replaces round 1's per-session scripts):
#
#   gpurun --timeout 2400 -- 'bash tools/gpu_run.sh tests smoke bench launches'
#
# Each argument is one stage, run in order; everything lands in gpurun_out/ (merged back by gpurun).
#   tests            pytest -m gpu (whole suite)                     -> gpu_tests.log
#   tests:<expr>     pytest -m gpu -k <expr>                         -> gpu_tests_<expr>.log
#   smoke            __graft_entry__.smoke()                         -> smoke.log
#   bench[:args]     python bench.py <args> (',' separates args)     -> bench_<args>.json / .log
#   bench2[:args]    the same under torchrun with 2 ranks (needs gpurun --gpus 2); bench4 / bench8 likewise
#   ref              python bench.py --impl reference                -> bench_ref.json
#   launches         ncu launch list (gpu__time_duration) of tools/profile_step.py -> launches.csv
#   ncufull:<regex>  ncu --set full on the kernels matching <regex> in tools/profile_step.py -> ncu_full_<regex>.csv
#                    (the .ncu-rep stays in /tmp on the box: gpurun copies back at most 64 MiB)
#   sweep[:ENV=V,..] tools/sweep_gemm.py under the given environment       -> gemm_sweep<env>.txt
#   micro:<what>     tools/microbench.py <what>                      -> micro_<what>.txt
#   py:<script>      python tools/<script>.py, 300 s cap                  -> py_<script>.txt
#   sass             cuobjdump opcode histogram per kernel           -> sass_summary.txt
#   ab:<ENV=V,...>   bench.py --no-cpu --no-library $AB_ARGS under the given environment (A/B switches), 150 s cap -> ab.txt
# Every stage runs under its own `timeout`: a hung kernel costs minutes, not the whole GPU budget.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for stage in "$@"; do
  name="${stage%%:*}"; arg=""; [[ "$stage" == *:* ]] && arg="${stage#*:}"
  tag="$(echo "$arg" | tr -c 'A-Za-z0-9_.\n-' '_')"
  echo "=== stage $stage ($(date +%T))"
  case "$name" in
    tests)
      if [ -z "$arg" ]; then timeout -s KILL 900 python -m pytest tests -m gpu -q -s > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/gpu_tests.log; grep "^\[parity\]\|^\[dropin\]" gpurun_out/gpu_tests.log > gpurun_out/parity_lines.txt
      else timeout -s KILL 600 python -m pytest tests -m gpu -x -q -s -k "$arg" > "gpurun_out/gpu_tests_$tag.log" 2>&1; echo "pytest exit $?"; tail -n 5 "gpurun_out/gpu_tests_$tag.log"; fi ;;
    smoke) timeout -s KILL 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log ;;
    bench) timeout -s KILL ${BENCH_TIMEOUT:-420} python bench.py $(echo "$arg" | tr ',' ' ') > "gpurun_out/bench_$tag.json" 2> "gpurun_out/bench_$tag.log"; echo "bench exit $?"; head -c 1500 "gpurun_out/bench_$tag.json"; echo ;;
    bench2|bench4|bench8)
      n="${name#bench}"
      timeout -s KILL ${BENCH_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$n" $(echo "$arg" | tr ',' ' ') > "gpurun_out/bench${n}_$tag.json" 2> "gpurun_out/bench${n}_$tag.log"; echo "bench$n exit $?"; head -c 1500 "gpurun_out/bench${n}_$tag.json"; echo ;;
    ref) timeout -s KILL 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "ref exit $?"; head -c 1200 gpurun_out/bench_ref.json; echo ;;
    launches)
      PROF_LLM_LAYERS=${PROF_LLM_LAYERS:-2} timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "ncu exit $?"; tail -n 2 gpurun_out/prof_step.log; wc -l gpurun_out/launches.csv ;;
    ncufull)
      PROF_LLM_LAYERS=${PROF_LLM_LAYERS:-2} timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k "regex:$arg" -c ${NCU_COUNT:-12} -s ${NCU_SKIP:-0} -o "/tmp/ncu_full_$tag" -f python tools/profile_step.py > "gpurun_out/ncu_full_$tag.log" 2>&1; echo "ncu exit $?"
      ncu -i "/tmp/ncu_full_$tag.ncu-rep" --page raw --csv > "gpurun_out/ncu_full_$tag.raw.csv" 2>/dev/null; python tools/ncu_summary.py "gpurun_out/ncu_full_$tag.raw.csv" > "gpurun_out/ncu_full_$tag.csv"; head -n 6 "gpurun_out/ncu_full_$tag.csv" ;;
    sweep) env $(echo "$arg" | tr ',' ' ') timeout -s KILL 900 python tools/sweep_gemm.py > "gpurun_out/gemm_sweep$tag.txt" 2>&1; echo "sweep exit $?"; cat "gpurun_out/gemm_sweep$tag.txt" ;;
    micro) timeout -s KILL 900 python tools/microbench.py $(echo "$arg" | tr ',' ' ') > "gpurun_out/micro_$tag.txt" 2>&1; echo "micro exit $?"; tail -n 40 "gpurun_out/micro_$tag.txt" ;;
    ab) echo "$arg $(env $(echo "$arg" | tr ',' ' ') timeout -s KILL 150 python bench.py --no-cpu --no-library --steps ${AB_STEPS:-8} ${AB_ARGS:-} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],3), {k: round(v,2) for k,v in d['stages'].items() if k.endswith('_ms')})" 2>&1 | tail -n 1)" | tee -a gpurun_out/ab.txt ;;
    py) timeout -s KILL 300 python "tools/$arg.py" > "gpurun_out/py_$tag.txt" 2>&1; echo "py exit $?"; tail -n 30 "gpurun_out/py_$tag.txt" ;;
    sass) python tools/sass_summary.py > gpurun_out/sass_summary.txt 2>&1; tail -n 30 gpurun_out/sass_summary.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu_info.txt 2>&1
echo "=== done ($(date +%T))"
