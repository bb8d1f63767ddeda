#!/bin/bash
# Round-2 evidence captures (run on the GPU box; everything lands in gpurun_out/, summaries are made from the raw CSVs afterwards).
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (per-launch durations, cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-raycast --no-job > gpurun_out/r2_launches_bench.log 2>&1
# 2. --set full of every trace launch of ONE step (warm-up step skipped: 32 trace launches per step)
ncu --set full --clock-control none --import-source on -k regex:traceKernel --launch-skip 32 -c 32 -f -o /tmp/r2_trace python scripts/profile_step.py 1 > gpurun_out/r2_trace.log 2>&1
ncu -i /tmp/r2_trace.ncu-rep --page raw --csv > gpurun_out/r2_trace.raw.csv 2>/dev/null
ncu -i /tmp/r2_trace.ncu-rep --page source --csv --kernel-name regex:traceKernel --launch-skip 0 --launch-count 1 > gpurun_out/r2_trace_bounce0.source.csv 2>/dev/null
# 3. the other kernels of the step: generate, bin*, shade, passTotals (first bounces of the captured step)
ncu --set full --clock-control none --import-source on -k regex:"shadeKernel|binCountKernel|binScanKernel|binScatterKernel|generateKernel|passTotalsKernel" \
    --launch-skip 66 -c 14 -f -o /tmp/r2_shade python scripts/profile_step.py 1 > gpurun_out/r2_shade.log 2>&1
ncu -i /tmp/r2_shade.ncu-rep --page raw --csv > gpurun_out/r2_shade.raw.csv 2>/dev/null
# 4. BVH build kernels (SAH build of the 10 M-triangle scene)
ncu --set full --clock-control none -k regex:"sahSplitKernel|sahClusterRootsKernel|sahSeedItemsKernel|sahGraftTopKernel|sahOrderedCountsKernel|collapseLevelKernel|radixTreeKernel|fitBoxesKernel|mortonKernel|gatherTrianglesKernel" \
    -c 120 -f -o /tmp/r2_build python scripts/profile_step.py 0 > gpurun_out/r2_build.log 2>&1
ncu -i /tmp/r2_build.ncu-rep --page raw --csv > gpurun_out/r2_build.raw.csv 2>/dev/null
# 5. VCM kernels (one iteration after a warm-up iteration)
ncu --set full --clock-control none -k regex:"vcm" --launch-skip 0 -c 200 -f -o /tmp/r2_vcm python scripts/vcm_profile_run.py 1 > gpurun_out/r2_vcm.log 2>&1
ncu -i /tmp/r2_vcm.ncu-rep --page raw --csv > gpurun_out/r2_vcm.raw.csv 2>/dev/null
# 6. curve variant of the trace kernel (grass scene, primary rays)
ls -la gpurun_out/ | head -40
find gpurun_out -type f -size +14M -print -delete
