#!/usr/bin/env bash
# One GPU-box session that re-establishes the state of the repo on hardware, in the order that
# matters if the session is cut short.  Run from the repo root, e.g.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh r2a'
# Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
tag="${1:-run}"
out=gpurun_out
mkdir -p "$out"
t0=$(date +%s)
note() { echo "[gpu_session +$(( $(date +%s) - t0 ))s] $*" | tee -a "$out/session_$tag.log"; }

note "nvidia-smi"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv | tee -a "$out/session_$tag.log"

# 1. smoke + the paths written after the last GPU session (online G=1, replay kernel)
note "smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a "$out/session_$tag.log"
note "first-run tests"; timeout 600 python -m pytest tests/test_zz_gpu_online.py tests/test_zz_gpu_replay.py -q -rxX --runxfail 2>&1 | tail -15 | tee "$out/zz_$tag.txt"

# 2. the bench lines (own arm, then the CPU reference arm)
note "bench playout"; timeout 600 python bench.py > "$out/bench_$tag.json" 2> "$out/bench_$tag.err"; tail -c 600 "$out/bench_$tag.json"
note "bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > "$out/bench_ref_$tag.json" 2> "$out/bench_ref_$tag.err"
note "bench search (fake net, 800 rollouts)"; timeout 900 python bench.py --workload mcts --fake-net --rollouts 800 --steps 3 --warmup 1 \
    > "$out/bench_mcts_fake800_$tag.json" 2> "$out/bench_mcts_fake800_$tag.err"

# 3. the full GPU suite
note "pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee "$out/pytest_gpu_$tag.txt"

# 4. launch list of the default bench command (per-launch times: share of the step, not absolutes)
note "ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$out/launches_$tag.csv" python bench.py --steps 2 --warmup 1 > "$out/b_ncu_$tag.log" 2>&1
note "done"
