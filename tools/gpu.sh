#!/bin/bash
# tools/gpu.sh <task> [args...] -- the one parametrised runner behind every `gpurun` call of this repo.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu.sh tests'            all GPU tests
#   ... 'bash tools/gpu.sh tests tests/test_gpu_pair.py -k bins'                      a subset (pytest args)
#   ... 'bash tools/gpu.sh bench [bench.py args]'                                     bench.py -> gpurun_out/bench_<tag>.json
#   ... 'bash tools/gpu.sh step  [tools/dev_step.py args]'                            per-phase timing of the fused step
#   ... 'bash tools/gpu.sh launches <tag> [bench.py args]'                            ncu launch list of a bench command
#   ... 'bash tools/gpu.sh ncu <tag> <kernel regex> [bench.py args]'                  one --set full capture of a kernel
#   ... 'bash tools/gpu.sh sanitize <memcheck|racecheck> [pytest args]'               compute-sanitizer over small tests
# Everything it writes goes to gpurun_out/ (merged back by gpurun); summaries worth keeping are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
task=${1:-tests}; shift || true
case "$task" in
  tests)
    if [ $# -eq 0 ]; then set -- tests; fi
    timeout 1500 python -m pytest "$@" -m gpu ${RSB_X:+-x} -q --timeout 900 2>&1 | tee gpurun_out/pytest_gpu.log | grep -E "^E  |FAILED|ERROR|passed|failed|skipped|Error" | cut -c1-400 | tail -40 ;;
  bench)
    tag=${RSB_TAG:-run}
    timeout 1500 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
    echo "rc=$?"; tail -3 gpurun_out/bench_$tag.err | cut -c1-300; cat gpurun_out/bench_$tag.json | cut -c1-3000 ;;
  step)
    timeout 1200 python tools/dev_step.py "$@" 2> gpurun_out/dev_step.err | tee -a gpurun_out/dev_step.jsonl; tail -3 gpurun_out/dev_step.err | cut -c1-300 ;;
  launches)
    tag=$1; shift
    timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --no-cpu --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_launches.log 2>&1
    echo "rc=$?"; tail -2 gpurun_out/${tag}_launches.log | cut -c1-300 ;;
  ncu)
    tag=$1; kern=$2; shift 2
    timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s ${RSB_SKIP:-3} -c 1 -f -o gpurun_out/${tag} \
        python bench.py --no-cpu --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_ncu.log 2>&1
    echo "rc=$?"; tail -2 gpurun_out/${tag}_ncu.log | cut -c1-300
    ncu -i gpurun_out/${tag}.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
    ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null ;;
  sanitize)
    tool=${1:-memcheck}; shift || true
    if [ $# -eq 0 ]; then set -- tests/test_gpu_pair.py -k "golden or random or edge or bins or padding"; fi
    timeout 1700 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -m gpu -x -q --timeout 1600 \
        > gpurun_out/sanitizer_$tool.log 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -5 ;;
  *) echo "unknown task $task"; exit 2 ;;
esac
